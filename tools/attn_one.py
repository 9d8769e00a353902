"""One launch of icaf_cross_attention per shape (for ncu --set full): python tools/attn_one.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from icafusion_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
for B, N, C in ((1, 5120, 128), (1, 5120, 256), (1, 5120, 512), (16, 400, 256)):
    n_pad = ops.round_up(N, 8)
    g = torch.Generator(device="cpu").manual_seed(1)
    qkv = [torch.randn(B, n_pad, 3 * C, generator=g).half().to(dev) for _ in range(2)]
    for _ in range(2):
        ops.cross_attention(qkv[0], qkv[1], None, None, B, N, n_pad, C, 8)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ops.cross_attention(qkv[0], qkv[1], None, None, B, N, n_pad, C, 8)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
