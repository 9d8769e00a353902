#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on the shared-memory reductions) over the training kernels at small shapes.
mkdir -p gpurun_out
SEL='small_training or attention_backward or attention_dropout or dmff_pool or conv_bn_act or bn_silu'
timeout 1000 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_loss.py -q -m gpu -k "$SEL or loss_backward" -x --timeout 900 > gpurun_out/sanitize_train_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_train_memcheck.log | tail -n 4
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_train_ops.py -q -m gpu -k "small_training or bn_silu or dmff_pool" -x --timeout 900 > gpurun_out/sanitize_train_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/sanitize_train_racecheck.log | tail -n 4
