#!/bin/bash
# Pair-kernel bring-up runs: conv parity first (default + pairs everywhere), then the full suite and both benches.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -x --timeout 400 > gpurun_out/pytest_conv.log 2>&1; rc=$?; echo "conv pytest rc=$rc"; tail -n 25 gpurun_out/pytest_conv.log | cut -c1-220
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python -m pytest tests -q -m gpu -x --timeout 400 --deselect tests/test_gpu_conv.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
python bench.py --steps 200 --warmup 20 --secondary none --layer-profile gpurun_out/layers_s_b1.csv > gpurun_out/bench_s_b1.json 2> gpurun_out/bench_s_b1.err; cut -c1-250 gpurun_out/bench_s_b1.json; tail -n 3 gpurun_out/bench_s_b1.err
python bench.py --workload yolov5l_b16 --secondary none --steps 20 --warmup 5 --layer-profile gpurun_out/layers_l_b16.csv > gpurun_out/bench_l_b16.json 2> gpurun_out/bench_l_b16.err; cut -c1-250 gpurun_out/bench_l_b16.json; tail -n 3 gpurun_out/bench_l_b16.err
