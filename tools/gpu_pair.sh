#!/bin/bash
# CTA-pair kernel: stress parity (pairs wherever eligible), default parity, then threshold sweep on the benches.
mkdir -p gpurun_out
ICAF_PAIR_BN=64 ICAF_PAIR_MIN=1 timeout 600 python -m pytest tests -q -m gpu -x --timeout 300 > gpurun_out/pytest_pair_all.log 2>&1; rc=$?; echo "pair-everywhere pytest rc=$rc"; tail -n 12 gpurun_out/pytest_pair_all.log | cut -c1-200
timeout 600 python -m pytest tests -q -m gpu -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu.log
for cfg in 256_148 256_37 128_148 128_37 64_148 64_37; do
  bn=${cfg%_*}; mn=${cfg#*_}
  ICAF_PAIR_BN=$bn ICAF_PAIR_MIN=$mn python bench.py --workload yolov5l_b16 --secondary none --steps 20 --warmup 5 --layer-profile gpurun_out/layers_l_b16_$cfg.csv > gpurun_out/bench_l_b16_$cfg.json 2> gpurun_out/bench_l_b16_$cfg.err; echo "$cfg $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_l_b16_$cfg.json | head -1)"
done
for cfg in 256_148 128_37 64_37 64_16; do
  bn=${cfg%_*}; mn=${cfg#*_}
  ICAF_PAIR_BN=$bn ICAF_PAIR_MIN=$mn python bench.py --secondary none --steps 200 --warmup 20 > gpurun_out/bench_s_b1_$cfg.json 2> gpurun_out/bench_s_b1_$cfg.err; echo "s_b1 $cfg $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_s_b1_$cfg.json | head -1)"
done
