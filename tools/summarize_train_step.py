"""gpurun_out/train_step.csv (ncu launch list of ONE training step: time, DRAM bytes, tensor-pipe / SM throughput) ->
profiles/<tag>_train_step_yolov5l_b16.md: per kernel totals, achieved DRAM rate, tensor-pipe activity.

    python tools/summarize_train_step.py r02
"""
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
src = os.path.join(ROOT, "gpurun_out", "train_step.csv")
lines = [l for l in open(src) if not l.startswith("==")]
launch = collections.OrderedDict()
for row in csv.DictReader(lines):
    d = launch.setdefault(row["ID"], {"k": re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("icaf::", "").strip(), "grid": row["Grid Size"]})
    v = float(row["Metric Value"].replace(",", ""))
    n, u = row["Metric Name"], row["Metric Unit"]
    if n.startswith("gpu__time_duration"):
        d["us"] = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    elif n.startswith("dram__bytes"):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        d["dram"] = d.get("dram", 0.0) + v * mult
    elif "pipe_tensor" in n:
        d["tensor"] = v
    elif n.startswith("sm__throughput"):
        d["sm"] = v
agg = collections.OrderedDict()
for d in launch.values():
    a = agg.setdefault(d["k"], {"n": 0, "us": 0.0, "dram": 0.0, "tensor_us": 0.0, "sm_us": 0.0})
    a["n"] += 1
    a["us"] += d.get("us", 0.0)
    a["dram"] += d.get("dram", 0.0)
    a["tensor_us"] += d.get("tensor", 0.0) * d.get("us", 0.0)
    a["sm_us"] += d.get("sm", 0.0) * d.get("us", 0.0)
tot = sum(a["us"] for a in agg.values())
ours = sum(a["us"] for k, a in agg.items() if not k.startswith(("at::", "nccl")))
out = [f"# Training step, yolov5l_Transfusion_kaist, 16 pairs of 640x512, one B200 -- ncu launch list ({tag})", "",
       "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active...,sm__throughput... "
       "--clock-control none --profile-from-start off python tools/train_one.py` (one eager step after two warm-up steps; per-launch times are "
       "serialised and cold-cache, so the SHARES are the evidence, not the sum: the graph-replayed step runs the two backbone streams "
       "concurrently).", "",
       "NOTE: this capture was cut by its 900 s limit after the launches below (a full eager step is ~2 660 launches, see "
       "`r02_train_launches_time_only.md`): the forward pass and roughly the first 70 % of the backward pass are in, the early-layer "
       "(largest-map) BatchNorm-backward / wgrad launches at the end of the backward pass are missing, so the shares of `wgrad_kernel`, "
       "`chan_partial_kernel<1>` and `bn_bwd_apply_kernel` are understated here.", "",
       f"{len(launch)} launches, {tot / 1e3:.2f} ms summed ({ours / 1e3:.2f} ms in libicaf_b200 kernels, {(tot - ours) / 1e3:.2f} ms in torch's: "
       "gradient accumulation at fan-outs, optimiser, GradScaler, copies).", "",
       "| kernel | launches | ms | share | avg us | DRAM GB (r+w) | DRAM TB/s | tensor pipe % (time-weighted) | SM throughput % |", "|---|---|---|---|---|---|---|---|---|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    if a["us"] < 0.1e3 and len(out) > 40:
        continue
    out.append(f"| `{k[:70]}` | {a['n']} | {a['us'] / 1e3:.3f} | {100 * a['us'] / tot:.1f} % | {a['us'] / a['n']:.1f} | {a['dram'] / 1e9:.2f} | "
               f"{a['dram'] / max(a['us'], 1e-9) / 1e6:.2f} | {a['tensor_us'] / max(a['us'], 1e-9):.1f} | {a['sm_us'] / max(a['us'], 1e-9):.1f} |")
path = os.path.join(ROOT, "profiles", f"{tag}_train_step_yolov5l_b16.md")
open(path, "w").write("\n".join(out) + "\n")
print("wrote", path)
