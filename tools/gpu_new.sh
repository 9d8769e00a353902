#!/bin/bash
# Quick GPU pass over the newest tests only.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nms.py tests/test_gpu_datasets.py -q -m gpu --timeout 600 -s > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; grep -E "^\[|passed|failed|Error|error|assert" gpurun_out/pytest_new.log | tail -n 20
python bench.py --gpus 1 --steps 20 --warmup 5 --secondary none > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('e2e_detect'))"
