#!/bin/bash
# Quick GPU session: parity tests + bench (primary + secondary workload) + ncu launch list (durations only).
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/pytest_gpu.log
python bench.py --steps 200 --warmup 20 --layer-profile gpurun_out/layers_s_b1.csv > gpurun_out/bench_s_b1.json 2> gpurun_out/bench_s_b1.err; cut -c1-300 gpurun_out/bench_s_b1.json; tail -n 3 gpurun_out/bench_s_b1.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_s_b1.csv python tools/profile_step.py --workload yolov5s_b1 --steps 2 > gpurun_out/ncu1.log 2>&1; tail -n 1 gpurun_out/ncu1.log
