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
    round2(dev, g)
    print("ok")


def round2(dev, g):
    """Round-2 kernels: TMA attention (split and fused-qkv forms, every head dim), LayerNorm-folded / statistics-emitting
    linears on the three conv kernels, the 5-launch DMFF block, token pooling with statistics, NMS, loss, letterbox."""
    import numpy as np
    from icafusion_b200 import TransformerFusionBlock
    from icafusion_b200.datasets import letterbox
    from icafusion_b200.loss import ComputeLoss
    import types
    for B, N, C in ((2, 100, 128), (1, 200, 256), (1, 130, 512), (1, 100, 1024)):
        n_pad = ops.round_up(N, 8)
        qkv = [torch.randn(B, n_pad, 3 * C, generator=g).half().to(dev) for _ in range(2)]
        o = ops.cross_attention(qkv[0], qkv[1], None, None, B, N, n_pad, C, 8)
        r = ops.cross_attention(qkv[0], qkv[1], None, None, B, N, n_pad, C, 8, simt=True)
        qk = [t[:, :, :2 * C].contiguous() for t in qkv]
        vt = [t[:, :, 2 * C:].permute(2, 0, 1).reshape(C, B * n_pad).contiguous() for t in qkv]
        o2 = ops.cross_attention(qk[0], qk[1], vt[0], vt[1], B, N, n_pad, C, 8)
        torch.cuda.synchronize()
        e = max(float((a.float() - b.float()).abs().max()) for a, b in zip(o + o2, r + r))
        print(f"attention B{B} N{N} C{C}: max abs diff vs CUDA-core reference {e:.2e}", flush=True)
        assert e < 5e-3
    for M, K, Nn in ((400, 256, 768), (6400, 256, 768), (20480, 512, 1536)):
        x = [torch.randn(M, K, generator=g).half().to(dev) for _ in range(2)]
        w = torch.randn(Nn, K, generator=g) / K ** 0.5
        pk = [ops.pack_linear_ln(w, torch.zeros(Nn), torch.ones(K), torch.zeros(K), 1e-5, ops.ACT_GELU, device=dev) for _ in range(2)]
        st = ops.row_stats(x[0], x[1])
        y = ops.linear(x, pk, ln_stats=list(st))
        pk2 = [ops.pack_linear(torch.randn(K, Nn, generator=g) / Nn ** 0.5, torch.zeros(K), device=dev) for _ in range(2)]
        so = [torch.empty(M, (K + 31) // 32, 2, device=dev) for _ in range(2)]
        al = torch.ones(2, device=dev)
        ops.linear(y, pk2, res=x, scaled=[(al[0:1], al[1:2])] * 2, stats_out=so)
        torch.cuda.synchronize()
        print(f"LN-folded + statistics linears M{M} K{K} N{Nn}: finite {bool(torch.isfinite(so[0]).all())}", flush=True)
    blk = TransformerFusionBlock(128, 10, 10).eval().half().to(dev)
    blk.crosstransformer[0].loops = 2
    with torch.no_grad():
        out = blk([torch.randn(2, 128, 16, 20, generator=g).half().to(dev), torch.randn(2, 128, 16, 20, generator=g).half().to(dev)])
    torch.cuda.synchronize()
    print("DMFF block (2 loops):", tuple(out.shape), bool(torch.isfinite(out.float()).all()), flush=True)
    z = torch.rand(2, 2000, 8, generator=g).half().to(dev)
    z[..., :4] *= 300
    det, cnt = ops.nms(z, 0.25, 0.45)
    det2, cnt2 = ops.nms(z, 0.001, 0.6, agnostic=True, classes=[0, 2])
    torch.cuda.synchronize()
    print("nms counts", cnt.tolist(), cnt2.tolist(), flush=True)
    anchors = np.array([[[1.25, 1.6], [2.0, 3.75], [4.1, 2.9]], [[1.9, 3.8], [3.9, 2.8], [3.7, 7.4]]], dtype=np.float32)
    stub = types.SimpleNamespace(hyp=dict(box=0.05, obj=1.0, cls=0.5, cls_pw=1.0, obj_pw=1.0, anchor_t=4.0, fl_gamma=0.0), gr=1.0,
                                 model=[types.SimpleNamespace(na=3, nc=2, nl=2, anchors=torch.from_numpy(anchors))])
    p = [torch.randn(2, 3, 16, 20, 7, generator=g).to(dev), torch.randn(2, 3, 8, 10, 7, generator=g).to(dev)]
    t = torch.rand(9, 6, generator=g)
    t[:, 0] = torch.randint(0, 2, (9,), generator=g).float()
    t[:, 1] = torch.randint(0, 2, (9,), generator=g).float()
    loss, items = ComputeLoss(stub)(p, t.to(dev))
    torch.cuda.synchronize()
    print("loss", loss.tolist(), items.tolist(), flush=True)
    fr = torch.randint(0, 256, (2, 120, 200, 3), generator=g, dtype=torch.uint8).to(dev)
    lb = letterbox(fr, (160, 256))[0]
    torch.cuda.synchronize()
    print("letterbox", tuple(lb.shape), flush=True)


if __name__ == "__main__":
    main()
