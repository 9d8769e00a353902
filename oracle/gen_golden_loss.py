"""Generate tests/golden/loss_cases.npz by executing the REAL reference's utils/loss.py:ComputeLoss (build container only).

    python -m oracle.gen_golden_loss

Predictions are seeded random Detect training outputs (rebuilt in the tests from the same numpy PCG64 streams, not stored);
targets are seeded random boxes (stored).  Cases: KAIST shape nc=1 (three levels of a 512x640 frame, batch 4), a 3-class
variant with label smoothing and non-unit BCE weights, gr = 0.5, and an empty target list.  Each case also stores the
reference's own backward (loss.backward()): the coarsest level's gradient in full and a 3-number fingerprint per level.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_shim import REF_ROOT, load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
HYP = dict(box=0.05, obj=1.0, cls=0.5, cls_pw=1.0, obj_pw=1.0, anchor_t=4.0, fl_gamma=0.0)
CASES = [  # name, nc, B, nt, gr, hyp overrides
    ("kaist_nc1", 1, 4, 37, 1.0, {}),
    ("nc3_smooth", 3, 2, 25, 1.0, dict(label_smoothing=0.1, cls_pw=1.5, obj_pw=0.7, box=0.07)),
    ("gr_half", 1, 2, 16, 0.5, {}),
    ("no_targets", 1, 2, 0, 1.0, {}),
]
LEVELS = [(64, 80), (32, 40), (16, 20)]


def synth_case(name, nc, B, nt, seed=3):
    """Seeded predictions / targets shared by the generator and the tests."""
    import zlib
    g = np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))
    p = [(1.5 * g.standard_normal((B, 3, ny, nx, nc + 5), dtype=np.float32)).astype(np.float32) for ny, nx in LEVELS]
    t = np.zeros((nt, 6), dtype=np.float32)
    if nt:
        t[:, 0] = g.integers(0, B, nt)
        t[:, 1] = g.integers(0, nc, nt)
        t[:, 2:4] = g.uniform(0.02, 0.98, (nt, 2))
        t[:, 4:6] = np.exp(g.uniform(np.log(0.01), np.log(0.6), (nt, 2)))
        t[: min(4, nt), 2] = [0.001, 0.999, 0.5, 0.0125][: min(4, nt)]          # boxes at the frame edge / on a cell boundary
    return p, t


def grad_fingerprint(g: np.ndarray, lvl: int) -> np.ndarray:
    """[sum |g|, <g, w1>, <g, w2>] in float64 with seeded Gaussian w: compact stand-in for the full gradient of a level."""
    r = np.random.Generator(np.random.PCG64([17, lvl]))
    g64 = g.astype(np.float64).reshape(-1)
    return np.array([np.abs(g64).sum(), g64 @ r.standard_normal(g64.size), g64 @ r.standard_normal(g64.size)], dtype=np.float64)


def main():
    _, yolo = load_reference()
    from utils.loss import ComputeLoss
    arrays, meta = {}, {"cases": []}
    for name, nc, B, nt, gr, over in CASES:
        cfg = os.path.join(REF_ROOT, "models", "transformer", "yolov5s_Transfusion_kaist.yaml")
        model = yolo.Model(cfg, ch=3, nc=nc)
        hyp = dict(HYP, **over)
        model.hyp, model.gr = hyp, gr
        loss_fn = ComputeLoss(model)
        p, t = synth_case(name, nc, B, nt)
        pt = [torch.from_numpy(x).requires_grad_(True) for x in p]
        loss, items = loss_fn(pt, torch.from_numpy(t))
        loss.sum().backward()                                                   # train.py:344 on the reference's own graph
        for lvl, x in enumerate(pt):
            arrays[f"{name}_gproj{lvl}"] = grad_fingerprint(x.grad.numpy(), lvl)
        arrays[f"{name}_grad2"] = pt[2].grad.numpy().astype(np.float32)         # coarsest level in full
        arrays[f"{name}_targets"] = t
        arrays[f"{name}_out"] = np.concatenate([loss.detach().numpy().reshape(1), items.numpy()]).astype(np.float32)
        arrays[f"{name}_anchors"] = model.model[-1].anchors.numpy().astype(np.float32)
        meta["cases"].append(dict(name=name, nc=nc, B=B, nt=nt, gr=gr, hyp=hyp))
        print(name, arrays[f"{name}_out"])
    meta["reference"] = "utils/loss.py:325-463 ComputeLoss(model)(p, targets), fp32 CPU"
    meta["torch"] = torch.__version__
    np.savez_compressed(os.path.join(OUT, "loss_cases.npz"), meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)
    print("wrote", os.path.join(OUT, "loss_cases.npz"))


if __name__ == "__main__":
    main()
