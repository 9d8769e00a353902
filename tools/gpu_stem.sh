#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm_ -c 3 -o gpurun_out/full_stem_l_b16 python tools/profile_step.py --workload yolov5l_b16 --steps 2 > gpurun_out/ncu_stem.log 2>&1; tail -n 1 gpurun_out/ncu_stem.log
