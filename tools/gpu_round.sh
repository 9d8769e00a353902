#!/bin/bash
# One GPU session: parity tests, smoke, both bench arms, ncu launch list + full capture of the dominant kernel.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_s_b1.json 2> gpurun_out/bench_s_b1.err; cat gpurun_out/bench_s_b1.json; tail -n 3 gpurun_out/bench_s_b1.err
python bench.py --workload yolov5l_b16 --steps 20 --warmup 5 > gpurun_out/bench_l_b16.json 2> gpurun_out/bench_l_b16.err; cat gpurun_out/bench_l_b16.json; tail -n 3 gpurun_out/bench_l_b16.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_s_b1.csv python tools/profile_step.py --workload yolov5s_b1 --steps 2 > gpurun_out/ncu1.log 2>&1; tail -n 2 gpurun_out/ncu1.log
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm_tc -c 20 -o gpurun_out/prof_conv_s_b1 python tools/profile_step.py --workload yolov5s_b1 --steps 2 > gpurun_out/ncu2.log 2>&1; tail -n 2 gpurun_out/ncu2.log
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm_tc -s 10 -c 10 -o gpurun_out/prof_conv_l_b16 python tools/profile_step.py --workload yolov5l_b16 --steps 2 > gpurun_out/ncu3.log 2>&1; tail -n 2 gpurun_out/ncu3.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_l_b16.csv python tools/profile_step.py --workload yolov5l_b16 --steps 2 > gpurun_out/ncu4.log 2>&1; tail -n 2 gpurun_out/ncu4.log
ls -la gpurun_out
