"""A handful of launches that cover every conv kernel variant, for compute-sanitizer (memcheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_probe.py
Geometries are small enough for the instrumented run; each result is compared with the CUDA-core reference kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from icafusion_b200 import ops  # noqa: E402

CASES = [
    # B, Cin, H, W, Cout, k, s, p    (ICAF_PAIR=all in the environment routes the TMA-able ones through the pair kernel)
    (2, 64, 32, 40, 128, 3, 1, 1),    # halo copies, BN=128 pairs
    (2, 128, 16, 20, 256, 3, 1, 1),   # halo copies, tiles overhang the 20-wide map
    (2, 64, 32, 40, 128, 3, 2, 1),    # stride 2: tap boxes
    (8, 16, 64, 80, 32, 3, 1, 1),     # 16-channel map: halo copies of the zero-filled 64-channel view, N = 32 < BN
    (2, 64, 32, 40, 96, 1, 1, 0),     # 1x1, ragged N
    (1, 8, 9, 11, 40, 3, 1, 1),       # gather path (one-tile kernel)
    (4, 64, 64, 80, 64, 1, 1, 0),     # many tiles, short K
]


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    worst = 0.0
    for B, Cin, H, W, Cout, k, s, p in CASES:
        x = torch.randn(B, H, W, Cin, generator=g).half().to(dev)
        w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
        pk = ops.pack_conv_weight(w, torch.randn(Cout, generator=g) * 0.1, s, p, ops.ACT_SILU, dev)
        y = ops.conv2d([x], [pk])[0]
        r = ops.conv2d([x], [pk], simt=True)[0]
        torch.cuda.synchronize()
        e = float((y.float() - r.float()).abs().max() / r.float().abs().max())
        worst = max(worst, e)
        print(f"{(B, Cin, H, W, Cout, k, s, p)}  err vs CUDA-core reference {e:.2e}", flush=True)
    assert worst < 2e-3, worst
    print("ok")


if __name__ == "__main__":
    main()
