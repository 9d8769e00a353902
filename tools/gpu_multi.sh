#!/bin/bash
# 2-GPU check of the torchrun bench path (ours + reference arm), both workloads.
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 20 --secondary none > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-260 gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --secondary none > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; cut -c1-260 gpurun_out/bench_n2.json; tail -n 5 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload yolov5l_b16 --steps 20 --warmup 5 --secondary none > gpurun_out/bench_l_b16_n2.json 2> gpurun_out/bench_l_b16_n2.err; cut -c1-260 gpurun_out/bench_l_b16_n2.json; tail -n 3 gpurun_out/bench_l_b16_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; cut -c1-200 gpurun_out/bench_ref_n2.json; tail -n 3 gpurun_out/bench_ref_n2.err
