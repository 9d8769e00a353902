"""Per-geometry timing of icaf_conv2d_fwd with the probe switches of the persistent kernel (tools only; needs a probe
build of the library: ICAF_PROBE=1 python -m icafusion_b200.build --force -- the shipped build has no such switches).
    python tools/conv_probe.py [--out gpurun_out/conv_probe.json]
Every geometry is timed behind an L2 flush with CUDA events (median of 7), for dbg in DBG and forced BN in BNS."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from icafusion_b200 import _lib, ops  # noqa: E402

# (name, B, H, W, Cin, Cout, k, s, p, n_io, res)
GEOMS = [
    ("stem_s2d_l", 16, 256, 320, 16, 64, 3, 1, 1, 2, False),
    ("1x1_N64_K64", 16, 128, 160, 64, 64, 1, 1, 0, 2, False),
    ("1x1_N128_K128", 16, 128, 160, 128, 128, 1, 1, 0, 2, False),
    ("1x1_N256_K256", 16, 64, 80, 256, 256, 1, 1, 0, 2, False),
    ("3x3_N64_K576", 16, 128, 160, 64, 64, 3, 1, 1, 2, True),
    ("3x3_N128_K1152", 16, 64, 80, 128, 128, 3, 1, 1, 2, True),
    ("3x3_N256_K2304", 16, 32, 40, 256, 256, 3, 1, 1, 2, True),
    ("3x3_N512_K4608", 16, 16, 20, 512, 512, 3, 1, 1, 2, True),
    ("3x3s2_N1024_K4608", 16, 32, 40, 512, 1024, 3, 2, 1, 2, False),
    ("3x3_N256_K2304_x1", 16, 32, 40, 256, 256, 3, 1, 1, 1, False),
    ("3x3_N128_K1152_x1", 16, 64, 80, 128, 128, 3, 1, 1, 1, False),
    ("3x3s2_N256_K1152", 16, 128, 160, 128, 256, 3, 2, 1, 2, False),
    ("3x3s2_N128_K576", 16, 256, 320, 64, 128, 3, 2, 1, 2, False),
    ("1x1_N128_K128_P3", 16, 64, 80, 128, 128, 1, 1, 0, 2, False),
    ("1x1_N256_K256_P4", 16, 32, 40, 256, 256, 1, 1, 0, 2, False),
    ("1x1_N512_K512_P5", 16, 16, 20, 512, 512, 1, 1, 0, 2, False),
    ("1x1_N256_K256_P3cv3", 16, 64, 80, 256, 256, 1, 1, 0, 2, False),
    ("1x1_N512_K512_P4cv3", 16, 32, 40, 512, 512, 1, 1, 0, 2, False),
    ("1x1_N1024_K1024_P5cv3", 16, 16, 20, 1024, 1024, 1, 1, 0, 2, False),
    ("3x3_N512_K4608_x1", 16, 16, 20, 512, 512, 3, 1, 1, 1, False),
]
DBG = [0, 1, 2, 3, 8, 16, 24, 32, 35]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/conv_probe.json")
    ap.add_argument("--bns", default="0")
    ap.add_argument("--dbg", default=",".join(map(str, DBG)))
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _lib.lib()
    if not hasattr(lib, "icaf_debug_set"):
        sys.exit("conv_probe needs a probe build: ICAF_PROBE=1 python -m icafusion_b200.build --force")
    lib.icaf_debug_set.argtypes = [C.c_int, C.c_int]
    lib.icaf_debug_set.restype = None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device="cpu").manual_seed(0)
    rows = []
    for name, B, H, W, Cin, Cout, k, s, p, n, res in GEOMS:
        xs = [torch.randn(B, H, W, Cin, generator=g).half().to(dev) for _ in range(n)]
        w = torch.randn(Cout, Cin, k, k, generator=g) * (1.0 / (Cin * k * k) ** 0.5)
        pk = [ops.pack_conv_weight(w, torch.zeros(Cout), s, p, ops.ACT_SILU, dev) for _ in range(n)]
        Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        outs = [torch.empty(B, Ho, Wo, Cout, dtype=torch.float16, device=dev) for _ in range(n)]
        rs = [torch.randn(B, Ho, Wo, Cout, generator=g).half().to(dev) for _ in range(n)] if res else None
        flops = 2.0 * B * Ho * Wo * Cout * Cin * k * k * n
        byts = 2.0 * n * (B * H * W * Cin + B * Ho * Wo * Cout * (2 if res else 1))
        for bn in [int(v) for v in a.bns.split(",")]:
            if bn and bn > max(32, Cout):
                continue
            for dbg in [int(v) for v in a.dbg.split(",")]:
                lib.icaf_debug_set(dbg, bn)
                ts = []
                for it in range(9):
                    flush.fill_(it)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    ops.conv2d(xs, pk, outs, rs)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                ts = sorted(ts[2:])
                us = ts[len(ts) // 2]
                rows.append({"geom": name, "bn": bn, "dbg": dbg, "us": round(us, 2), "tflops": round(flops / us / 1e6, 1),
                             "gbs": round(byts / us / 1e3, 1)})
                print(f"{name:20s} bn={bn:3d} dbg={dbg:2d}  {us:8.1f} us  {flops / us / 1e6:7.1f} TF/s  {byts / us / 1e3:7.1f} GB/s", flush=True)
    lib.icaf_debug_set(0, 0)
    with open(a.out, "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
