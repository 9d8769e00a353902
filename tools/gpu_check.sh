#!/bin/bash
# Green check on the shipped tree: the whole GPU parity suite (no -x), smoke, then the driver's own bench invocations.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 900 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -n 5; grep -E "^\[" gpurun_out/pytest_gpu.log | head -n 80
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_ref.json
python bench.py --gpus 1 --steps 20 --warmup 5 --layer-profile gpurun_out/layers_l_b16.csv > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench_n1.err
