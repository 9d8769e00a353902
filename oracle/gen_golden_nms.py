"""Generate tests/golden/nms_*.npz by executing the REAL reference's utils/general.py:non_max_suppression
(build container only; needs /root/reference).

    python -m oracle.gen_golden_nms

Input predictions are built from the committed detector goldens (tests/golden/yolov5s_*.npz: the reference's own decoded
output `z`), rounded to fp16 -- the dtype the device path hands to NMS -- with objectness / class scores re-spread so that
both the sparse (conf 0.25) and the dense (conf 0.001, test.py's setting) regimes hold a few hundred to a few thousand
candidates with heavy overlap; a 3-class variant exercises the class offset.  Stored: the fp16 predictions and, per
setting, the reference's kept rows (fp32) for every image.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_shim import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SETTINGS = [  # name, conf, iou, agnostic, classes
    ("detect", 0.25, 0.45, False, None),           # detect_twostream.py defaults
    ("test", 0.001, 0.6, False, None),             # test.py:  conf_thres=0.001, iou_thres=0.6
    ("agnostic", 0.25, 0.45, True, None),
    ("class1", 0.1, 0.45, False, [1]),
]


def build_predictions(seed=5):
    g = np.random.Generator(np.random.PCG64(seed))
    d = np.load(os.path.join(OUT, "yolov5s_512x640.npz"))
    z = np.asarray(d["z_fused"], dtype=np.float32)            # (1, 20160, 6) decoded boxes of the real reference
    z = np.concatenate([z, z[:, ::-1] * np.array([1, 1, 1.1, 0.9, 1, 1], dtype=np.float32)], 0)   # second image: other order / sizes
    B, R, _ = z.shape
    nc = 3
    obj = g.beta(0.35, 2.2, size=(B, R)).astype(np.float32)   # image 0: thousands above 0.25 (saturates max_det = 300)
    obj[1] = g.beta(0.08, 6.0, size=R).astype(np.float32)     # image 1: sparse -- a few dozen detections survive
    obj[1, ::7] = 0.0
    cls = g.uniform(0.05, 1.0, size=(B, R, nc)).astype(np.float32)
    pred = np.concatenate([z[..., :4], obj[..., None], cls], -1)
    pred[..., 2:4] = np.clip(pred[..., 2:4] * 3.0, 4.0, 400.0)   # larger boxes -> many overlaps
    return pred.astype(np.float16)


def main():
    load_reference()
    from utils.general import non_max_suppression      # the reference's own function
    pred16 = build_predictions()
    arrays, meta = {"pred": pred16}, {"settings": []}
    for name, conf, iou, agn, classes in SETTINGS:
        out = non_max_suppression(torch.from_numpy(pred16).float(), conf, iou, classes=classes, agnostic=agn)
        counts = [int(o.shape[0]) for o in out]
        meta["settings"].append(dict(name=name, conf=conf, iou=iou, agnostic=agn, classes=classes, counts=counts))
        for b, o in enumerate(out):
            arrays[f"{name}_{b}"] = o.numpy().astype(np.float32)
        print(name, counts)
    meta["reference"] = "utils/general.py:518-607 non_max_suppression on pred.float() (CPU, torchvision.ops.nms)"
    meta["torch"] = torch.__version__
    path = os.path.join(OUT, "nms_cases.npz")
    np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
