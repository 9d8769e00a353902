"""What bounds the image stem?  Time the stem geometry (16 x 256 x 320 map, 64 filters, 3x3/s1, two streams) with 16, 32 and
64 input channels: 9 / 18 / 36 MMAs and 1x / 2x / 4x the operand bytes per tile, same output."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from icafusion_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator().manual_seed(0)
for B, H, W, Cout in ((16, 256, 320, 64), (1, 256, 320, 32)):
    for Cin in (16, 32, 64):
        xs = [torch.randn(B, H, W, Cin, generator=g).half().to(dev) for _ in range(2)]
        w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
        pk = [ops.pack_conv_weight(w, torch.zeros(Cout), 1, 1, ops.ACT_SILU, dev) for _ in range(2)]
        outs = [torch.empty(B, H, W, Cout, dtype=torch.float16, device=dev) for _ in range(2)]
        ts = []
        for it in range(9):
            flush.fill_(it)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv2d(xs, pk, outs)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts = sorted(ts[2:])
        print(f"B={B} Cin={Cin:3d} Cout={Cout}: {ts[len(ts) // 2]:8.1f} us", flush=True)
