#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attn.py tests/test_gpu_dmff.py -q -m gpu -x -s --timeout 300 > gpurun_out/pytest_attn.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error|error" gpurun_out/pytest_attn.log | tail -n 5
timeout 300 python tools/attn_probe.py > gpurun_out/attn_probe_tma.log 2>&1; tail -n 16 gpurun_out/attn_probe_tma.log
