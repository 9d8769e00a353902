#!/bin/bash
# Diagnostic session: parity tests, per-geometry probe timings, benches.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/pytest_gpu.log
timeout 600 python tools/conv_probe.py --dbg 0,1,3 --out gpurun_out/conv_probe.json > gpurun_out/conv_probe.log 2>&1; echo "probe rc=$?"
python bench.py --steps 200 --warmup 20 --layer-profile gpurun_out/layers_s_b1.csv > gpurun_out/bench_s_b1.json 2> gpurun_out/bench_s_b1.err; cut -c1-300 gpurun_out/bench_s_b1.json; tail -n 3 gpurun_out/bench_s_b1.err
python bench.py --workload yolov5l_b16 --secondary none --steps 20 --warmup 5 --layer-profile gpurun_out/layers_l_b16.csv > gpurun_out/bench_l_b16.json 2> gpurun_out/bench_l_b16.err; cut -c1-300 gpurun_out/bench_l_b16.json; tail -n 3 gpurun_out/bench_l_b16.err
