#!/bin/bash
# Evidence pass: DMFF sweep (BASELINE configs[4], complete), per-launch ncu step lists of both workloads, and --set full
# captures of the dominant kernels.  Summaries: python tools/summarize_profiles.py r02
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 python tools/dmff_sweep.py > gpurun_out/dmff_sweep.log 2>&1; tail -n 2 gpurun_out/dmff_sweep.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed
for wl in yolov5l_b16 yolov5s_b1; do
  timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_$wl.csv python tools/profile_step.py --workload $wl --steps 2 > gpurun_out/ncu_step_$wl.log 2>&1; tail -n 1 gpurun_out/ncu_step_$wl.log
done
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_ -c 4 -o gpurun_out/full_conv_l_b16_head python tools/profile_step.py --workload yolov5l_b16 --steps 2 > gpurun_out/ncu_full_l1.log 2>&1; tail -n 1 gpurun_out/ncu_full_l1.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm_pair -s 8 -c 3 -o gpurun_out/full_conv_l_b16_pair python tools/profile_step.py --workload yolov5l_b16 --steps 2 > gpurun_out/ncu_full_l2.log 2>&1; tail -n 1 gpurun_out/ncu_full_l2.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:cross_attn_ -c 3 -o gpurun_out/full_attn_l_b16 python tools/profile_step.py --workload yolov5l_b16 --steps 2 > gpurun_out/ncu_full_a.log 2>&1; tail -n 1 gpurun_out/ncu_full_a.log
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
