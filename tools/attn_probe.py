"""Time icaf_cross_attention alone (both directions, 8 heads) over the DMFF shapes: CUDA events around 20 back-to-back
launches, median of 5 rounds.  ICAF_ATTN=legacy python tools/attn_probe.py  -> the round-1 cp.async kernels."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from icafusion_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
CASES = [(1, 400, 128), (1, 400, 256), (1, 256, 512), (1, 100, 1024), (16, 400, 256), (16, 256, 512), (16, 100, 1024),
         (1, 1280, 128), (1, 1280, 256), (1, 5120, 128), (1, 5120, 256), (1, 5120, 512), (16, 1280, 256), (16, 1280, 512), (4, 5120, 512)]
rows = []
for B, N, C in CASES:
    n_pad = ops.round_up(N, 8)
    g = torch.Generator(device="cpu").manual_seed(1)
    qk_v, qk_i = [torch.randn(B, n_pad, 2 * C, generator=g).half().to(dev) for _ in range(2)]
    vt_v, vt_i = [torch.randn(C, B * n_pad, generator=g).half().to(dev) for _ in range(2)]
    for _ in range(3):
        ops.cross_attention(qk_v, qk_i, vt_v, vt_i, B, N, n_pad, C, 8)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.cross_attention(qk_v, qk_i, vt_v, vt_i, B, N, n_pad, C, 8)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 20)
    ms = sorted(ts)[2]
    fl = 8.0 * B * N * N * C
    exps = 2.0 * B * 8 * N * N
    rows.append(dict(B=B, N=N, C=C, d=C // 8, us=round(ms * 1e3, 2), tflops=round(fl / ms / 1e9, 1), gexp_s=round(exps / ms / 1e6, 1)))
    print(rows[-1], flush=True)
out = os.path.join(ROOT, "gpurun_out", f"attn_probe_{os.environ.get('ICAF_ATTN', 'tma')}.json")
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump(rows, open(out, "w"), indent=1)
