#!/bin/bash
# Quick GPU pass over the newest tests only.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_loss.py tests/test_gpu_datasets.py tests/test_gpu_aux.py -q -m gpu --timeout 600 -s > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; grep -E "^\[|passed|failed|Error|error|assert" gpurun_out/pytest_new.log | tail -n 40
