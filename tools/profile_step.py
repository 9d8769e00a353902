"""Run a few eager (non-graph) forward steps of the detector -- the command ncu wraps.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python tools/profile_step.py --workload yolov5s_b1 --steps 3
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from icafusion_b200 import Model, synth  # noqa: E402
from icafusion_b200.synth import load_synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="yolov5s_b1")
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
wl = WORKLOADS[a.workload]
dev = torch.device("cuda:0")
model = Model(f"yolov5{wl['size']}_Transfusion_kaist").eval()
load_synth(model, 0)
model = model.fuse().half().to(dev)
rgb, ir = [(t * 255).to(torch.uint8).to(dev) for t in synth.synth_images(wl["batch"], wl["H"], wl["W"], 0)]
with torch.no_grad():
    for _ in range(a.steps):            # warm-up: filter packing, kernel attribute setup (outside the profiled range)
        model(rgb, ir)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()         # ncu --profile-from-start off
    model(rgb, ir)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
