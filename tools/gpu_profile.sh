#!/bin/bash
# ncu evidence for profiles/: (1) per-launch duration + DRAM bytes of one step, both workloads (metrics-only, one pass);
# (2) --set full captures (with source) of a few launches of the dominant kernels.  Outputs stay well below 64 MiB.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed
for wl in yolov5s_b1 yolov5l_b16; do
  timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_$wl.csv python tools/profile_step.py --workload $wl --steps 2 > gpurun_out/ncu_step_$wl.log 2>&1; tail -n 1 gpurun_out/ncu_step_$wl.log
done
# yolov5l batch 16: the first four conv launches (stem = pair<64> halo-2, stride-2 down conv, two 1x1 of the first C3) ...
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm_ -c 4 -o gpurun_out/full_conv_l_b16_head python tools/profile_step.py --workload yolov5l_b16 --steps 2 > gpurun_out/ncu_full_l1.log 2>&1; tail -n 1 gpurun_out/ncu_full_l1.log
# ... and three CTA-pair launches from the P3/P4 stages (3x3 halo layers, BN = 128 / 256)
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm_pair -s 8 -c 3 -o gpurun_out/full_conv_l_b16_pair python tools/profile_step.py --workload yolov5l_b16 --steps 2 > gpurun_out/ncu_full_l2.log 2>&1; tail -n 1 gpurun_out/ncu_full_l2.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm_pair -s 16 -c 2 -o gpurun_out/full_conv_l_b16_pair2 python tools/profile_step.py --workload yolov5l_b16 --steps 2 > gpurun_out/ncu_full_l3.log 2>&1; tail -n 1 gpurun_out/ncu_full_l3.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm_ -s 18 -c 3 -o gpurun_out/full_conv_s_b1 python tools/profile_step.py --workload yolov5s_b1 --steps 2 > gpurun_out/ncu_full_s.log 2>&1; tail -n 1 gpurun_out/ncu_full_s.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:cross_attn_ -c 3 -o gpurun_out/full_attn_l_b16 python tools/profile_step.py --workload yolov5l_b16 --steps 2 > gpurun_out/ncu_full_a.log 2>&1; tail -n 1 gpurun_out/ncu_full_a.log
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
