#!/bin/bash
# Evidence for the training step: ncu launch list (time, DRAM bytes, tensor pipe) of one yolov5l b16 step + a --set full capture of wgrad_kernel.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/train_step.csv python tools/train_one.py > gpurun_out/train_step.log 2>&1; tail -n 1 gpurun_out/train_step.log
python tools/ncu_kernel_sum.py gpurun_out/train_step.csv 12
