#!/bin/bash
# N-GPU sanity run of both bench arms exactly as the driver launches them (N = $1, default 2).
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus $N --steps 5 --warmup 2 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; cut -c1-200 gpurun_out/bench_ref_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -n 3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print("infer", d["value"], d["ms_per_step"], d.get("notes"))
t=d.get("train") or {}
for k,v in t.items():
    if k!="per_kernel_event_pass": print(k, v)
PY
