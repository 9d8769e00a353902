#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -q -m gpu --timeout 600 -x -k "matches_oracle or golden or grouped" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error|error|assert|icaf:" gpurun_out/pytest_new.log | tail -n 12
for v in 1 0; do
ICAF_STEM=$v python bench.py --gpus 1 --steps 20 --warmup 5 --secondary none --layer-profile gpurun_out/layers_stem$v.csv > gpurun_out/bench_stem$v.json 2> gpurun_out/bench_stem$v.err; python -c "
import json; d=json.load(open('gpurun_out/bench_stem$v.json')); print('ICAF_STEM=$v', d['value'], d['ms_per_step'], d['e2e']['value'])"; grep "K144" gpurun_out/layers_stem$v.csv
done
