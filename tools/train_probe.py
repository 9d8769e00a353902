"""Diagnostics: per-layer deviation of the training-mode forward from the fp32 oracle, next to the deviation of the oracle
itself when it runs on the GPU under fp16 autocast (the reference's own training regime, train.py:334)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from icafusion_b200 import Model, autograd
from icafusion_b200.cfg import load_cfg
from icafusion_b200.synth import load_synth
from oracle import icaf_oracle as O, synth

size, B, H, W, seed = sys.argv[1] if len(sys.argv) > 1 else "s", 2, 320, 320, 1234
cfg = load_cfg(f"yolov5{size}_Transfusion_kaist")
rgb, ir = synth.synth_images(B, H, W, seed)
sd = synth.synth_state_dict(synth.model_param_shapes(cfg), seed)
ref = []
with torch.no_grad():
    O.model_forward({k: v.clone() for k, v in sd.items()}, cfg, rgb, ir, training=True, taps=ref)
amp = []
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    O.model_forward({k: v.clone().cuda() for k, v in sd.items()}, cfg, rgb.cuda(), ir.cuda(), training=True, taps=amp)
model = Model(f"yolov5{size}_Transfusion_kaist")
load_synth(model, seed)
model = model.cuda().train()
for m in model.modules():
    if isinstance(m, torch.nn.Dropout):
        m.p = 0.0
ours = []
with torch.no_grad():
    autograd.model_forward(model, rgb.cuda(), ir.cuda(), taps=ours)
e = lambda a, b: float((a.float().cpu() - b).abs().max() / b.abs().max())
for i, (r, a, o) in enumerate(zip(ref, amp, ours)):
    if isinstance(r, (list, tuple)):
        for j in range(len(r)):
            print(f"layer {i:2d}[{j}] {type(model.model[i]).__name__:24s} ours {e(o[j], r[j]):.2e}   autocast oracle {e(a[j], r[j]):.2e}")
    else:
        print(f"layer {i:2d}    {type(model.model[i]).__name__:24s} ours {e(o.permute(0, 3, 1, 2), r):.2e}   autocast oracle {e(a, r):.2e}")
