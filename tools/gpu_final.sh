#!/bin/bash
# Round-end evidence on one GPU: parity suite, smoke, reference arm, both benches with layer profiles, DMFF sweep.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
python bench.py --steps 200 --warmup 20 --layer-profile gpurun_out/layers_s_b1.csv > gpurun_out/bench_s_b1.json 2> gpurun_out/bench_s_b1.err; cut -c1-300 gpurun_out/bench_s_b1.json; tail -n 3 gpurun_out/bench_s_b1.err
python bench.py --workload yolov5l_b16 --secondary none --steps 20 --warmup 5 --layer-profile gpurun_out/layers_l_b16.csv > gpurun_out/bench_l_b16.json 2> gpurun_out/bench_l_b16.err; cut -c1-300 gpurun_out/bench_l_b16.json; tail -n 3 gpurun_out/bench_l_b16.err
timeout 900 python tools/dmff_sweep.py --out gpurun_out/dmff_sweep.json > gpurun_out/dmff_sweep.log 2>&1; echo "sweep rc=$?"; tail -n 3 gpurun_out/dmff_sweep.log
