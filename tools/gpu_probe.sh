#!/bin/bash
# Probe build (stage switches + forced tile width) -> per-geometry timings; the shipped library is rebuilt afterwards.
mkdir -p gpurun_out
ICAF_PROBE=1 python -m icafusion_b200.build --force > gpurun_out/probe_build.log 2>&1; tail -n 1 gpurun_out/probe_build.log
timeout 600 python tools/conv_probe.py --dbg 0 --bns 0,64,128,256 --out gpurun_out/conv_probe_bn.json > gpurun_out/conv_probe_bn.log 2>&1; echo "probe rc=$?"; cat gpurun_out/conv_probe_bn.log | cut -c1-120
python -m icafusion_b200.build --force > /dev/null 2>&1
