#!/bin/bash
# Quick GPU pass over the newest tests only.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nms.py tests/test_gpu_aux.py -q -m gpu --timeout 600 -x > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/pytest_new.log
