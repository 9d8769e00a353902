#!/bin/bash
# Flakiness soak: the GPU parity suite several times over (fresh process each), then smoke.
mkdir -p gpurun_out
for i in 1 2 3 4 5; do
  python -m pytest tests -q -m gpu -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_soak_$i.log 2>&1; echo "soak $i rc=$? $(tail -n 1 gpurun_out/pytest_soak_$i.log)"
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
