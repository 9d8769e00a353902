#!/bin/bash
# compute-sanitizer over a few launches of every conv kernel variant (default dispatch, then pairs everywhere).
mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_probe.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 6 gpurun_out/sanitize_memcheck.log | cut -c1-200
ICAF_PAIR=all timeout 800 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_probe.py > gpurun_out/sanitize_memcheck_pairs.log 2>&1; echo "memcheck(pairs) rc=$?"; tail -n 6 gpurun_out/sanitize_memcheck_pairs.log | cut -c1-200
ICAF_PAIR=all timeout 800 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_probe.py > gpurun_out/sanitize_synccheck_pairs.log 2>&1; echo "synccheck(pairs) rc=$?"; tail -n 6 gpurun_out/sanitize_synccheck_pairs.log | cut -c1-200
