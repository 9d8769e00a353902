#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_train_model.py -q -m gpu --timeout 600 -s > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error|error|assert|icaf:|attention" gpurun_out/pytest_new.log | tail -n 14
timeout 900 python bench.py --steps 5 --warmup 3 --secondary none > gpurun_out/bench_train1.json 2> gpurun_out/bench_train1.err; tail -2 gpurun_out/bench_train1.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_train1.json").read().strip().splitlines()[-1])
print(d.get("notes")); t=d["train"]
for k,v in t.items():
    if k not in("per_kernel_event_pass","config","metric","e2e"): print(k, v)
print("infer", d["value"], d["ms_per_step"])
for k,v in sorted(t["per_kernel_event_pass"].items(), key=lambda kv:-kv[1]["ms"])[:12]: print(f"{k:32s} {v['launches']:4d} {v['ms']:8.3f} ms")
PY
