"""DMFF-block throughput sweep (BASELINE.json configs[4]): C in {128,256,512}, (H,W) in {(64,80),(32,40),(16,20)}
un-pooled (N = H*W tokens) and as shipped (pooled token grids), loops in {1,2,4}, fp16, batch in {1,16}.
Each case: CUDA-graph replay of TransformerFusionBlock.run on device-resident NHWC maps, L2 flushed between replays,
GFLOP/s from F_dmff = B*[L*(48 N C^2 + 8 N^2 C) + 4 H W C^2] against the measured tensor peak and
the ideal-bytes HBM bound.

    python tools/dmff_sweep.py [--quick] [--out gpurun_out/dmff_sweep.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from icafusion_b200 import TransformerFusionBlock  # noqa: E402
from icafusion_b200.synth import load_synth  # noqa: E402


def dmff_flops(B, C, H, W, N, loops=1):
    """F_dmff = B*[L*(48 N C^2 + 8 N^2 C) + 4 H W C^2]  """
    return B * (loops * (48 * N * C * C + 8 * N * N * C) + 4 * H * W * C * C)


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return p["bf16_tflops"], p["hbm_gbs"]        # burst: each block is timed alone (sub-ms to a few ms)
    except Exception:  # noqa: BLE001
        return 1590.0, 6650.0


def time_block(C, H, W, va, ha, loops, B, dev, reps=20):
    blk = TransformerFusionBlock(C, va, ha).eval()
    blk.crosstransformer[0].loops = loops
    load_synth(blk, 1, "blk.")
    blk = blk.half().to(dev)
    rgb = torch.randn(B, H, W, C, device=dev).half()
    ir = torch.randn(B, H, W, C, device=dev).half()
    st = torch.cuda.Stream(dev)
    with torch.no_grad(), torch.cuda.stream(st):
        for _ in range(2):
            blk.run(rgb, ir)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            out = blk.run(rgb, ir)
    st.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        g.replay()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for s, e in ev:
        flush.zero_()
        s.record()
        g.replay()
        e.record()
    torch.cuda.synchronize()
    ms = sorted(s.elapsed_time(e) for s, e in ev)[len(ev) // 2]
    del g, out
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "dmff_sweep.json"))
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    tf_peak, hbm = peaks()
    rows = []
    grids = {(64, 80): (20, 20), (32, 40): (16, 16), (16, 20): (10, 10)}
    Cs = (128, 256, 512)
    loops_set = (1,) if a.quick else (1, 2, 4)
    for B in (1, 16):
        for C in Cs:
            for (H, W), shipped in grids.items():
                for mode, (va, ha) in (("unpooled", (H, W)), ("pooled", shipped)):
                    for L in loops_set:
                        N = va * ha
                        if mode == "pooled" and L != 1:
                            continue
                        ms = time_block(C, H, W, va, ha, L, B, dev, reps=8 if (B == 16 and N >= 5120) else 20)
                        F = dmff_flops(B, C, H, W, N, L)
                        bytes_ideal = 2.0 * (3 * B * C * H * W + 2 * N * C + 26 * C * C)
                        t_bound = max(F / (tf_peak * 1e12), bytes_ideal / (hbm * 1e9))
                        rows.append(dict(B=B, C=C, H=H, W=W, tokens=N, mode=mode, loops=L, ms=round(ms, 4),
                                         gflops=round(F / ms / 1e6, 1), frac_tensor_peak=round(F / (ms * 1e-3) / (tf_peak * 1e12), 4),
                                         frac_roofline=round(t_bound / (ms * 1e-3), 4),
                                         bound="tensor" if F / (tf_peak * 1e12) > bytes_ideal / (hbm * 1e9) else "hbm"))
                        print(rows[-1], flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump({"peaks": {"tensor_tflops": tf_peak, "hbm_gbs": hbm}, "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
