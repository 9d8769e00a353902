#!/bin/bash
# Green check: the whole GPU parity suite (no -x), smoke, and both bench workloads.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
python bench.py --workload yolov5l_b16 --secondary none --steps 20 --warmup 5 > gpurun_out/bench_l_b16.json 2> gpurun_out/bench_l_b16.err; cut -c1-600 gpurun_out/bench_l_b16.json; tail -n 3 gpurun_out/bench_l_b16.err
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_s_b1.json 2> gpurun_out/bench_s_b1.err; cut -c1-600 gpurun_out/bench_s_b1.json; tail -n 3 gpurun_out/bench_s_b1.err
