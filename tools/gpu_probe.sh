#!/bin/bash
# Run each GPU test file in its own process (a trapped kernel poisons only that process), logs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in "$@"; do
  n=$(basename $f .py)
  timeout 600 python -m pytest $f -q -m gpu -s --timeout 300 > gpurun_out/$n.log 2>&1
  echo "== $n exit $?"; tail -n 25 gpurun_out/$n.log
done
