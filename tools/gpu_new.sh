#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_model.py -k graphed -q -m gpu --timeout 600 -s > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; grep -E "^\[|passed|failed|Error|error|assert|icaf:" gpurun_out/pytest_new.log | tail -n 20
