"""Generate tests/golden/*.npz by executing the REAL reference (build container only).

    python -m oracle.gen_golden            # needs /root/reference (read-only, never copied)

Each file stores the reference's fp32 CPU output for one seeded case plus the case
description; weights and inputs are NOT stored -- they are rebuilt from `oracle/synth.py`
(numpy PCG64 streams keyed by seed + parameter name).  tests/test_oracle_golden.py replays
every file against `oracle/icaf_oracle.py`; the `-m gpu` tests replay them against the CUDA
path.  The reference cannot travel to the GPU box, these vectors can.
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import synth  # noqa: E402
from oracle.ref_shim import REF_ROOT, load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

DMFF_CASES = [
    # name,            B, C,   H,  W,  va, ha, loops
    ("dmff_p5_pooled", 2, 128, 16, 20, 10, 10, 1),     # k=(7,2) s=(1,2) overlapping windows
    ("dmff_unpooled_l2", 1, 128, 16, 20, 16, 20, 2),   # identity pool, N=320, iterative loop
    ("dmff_cfg1_pooled", 1, 256, 32, 40, 16, 16, 1),   # BASELINE config 1 as shipped (N=256)
    ("dmff_cfg1_unpooled", 1, 256, 32, 40, 32, 40, 1),  # BASELINE config 1 unpooled (N=1280)
    ("dmff_p3_pooled_l4", 1, 128, 64, 80, 20, 20, 4),  # P3 geometry k=(7,4) s=(3,4), 4 loops
]
MODEL_CASES = [
    # name,           size, B, H,   W
    ("yolov5s_320", "s", 2, 320, 320),
    ("yolov5s_512x640", "s", 1, 512, 640),
    ("yolov5l_512x640", "l", 1, 512, 640),
]


def _save(name, meta, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)
    print(f"wrote {path}  ({os.path.getsize(path) / 1e6:.2f} MB)")


def _load_synth(module, seed):
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    sd = synth.synth_state_dict(shapes, seed)
    missing = module.load_state_dict(sd, strict=False)
    assert all(k.endswith(("anchors", "anchor_grid")) for k in missing.missing_keys), missing
    assert not missing.unexpected_keys


def main():
    warnings.filterwarnings("ignore")
    torch.set_grad_enabled(False)
    common, yolo = load_reference()
    seed = 1234

    for name, B, C, H, W, va, ha, loops in DMFF_CASES:
        blk = common.TransformerFusionBlock(C, va, ha).eval()
        blk.crosstransformer[0].loops = loops
        pre = "blk"
        shapes = synth.dmff_param_shapes(C, va * ha, pre)
        ref_shapes = {f"{pre}.{k}": tuple(v.shape) for k, v in blk.state_dict().items()}
        assert ref_shapes == {k: tuple(v) for k, v in shapes.items()}, "state_dict layout drifted"
        sd = synth.synth_state_dict(shapes, seed)
        blk.load_state_dict({k[len(pre) + 1:]: v for k, v in sd.items()}, strict=True)
        rgb, ir = synth.synth_features(B, C, H, W, seed)
        out = blk([rgb, ir])
        # intermediate: token streams after the cross transformer (pins a2-a5 separately from the tail)
        r = blk.vis_coefficient(blk.avgpool(rgb), blk.maxpool(rgb)).flatten(2).permute(0, 2, 1) + blk.pos_emb_vis
        i = blk.ir_coefficient(blk.avgpool(ir), blk.maxpool(ir)).flatten(2).permute(0, 2, 1) + blk.pos_emb_ir
        tr, ti = blk.crosstransformer([r, i])
        try:   # how far the reference's own fp16 path (detect_twostream.py:40-41) sits from its fp32 path
            o16 = blk.half()([rgb.half(), ir.half()]).float()
            dev16 = float((o16 - out).abs().max() / out.abs().max())
            blk.float()
        except Exception as e:  # noqa: BLE001
            dev16 = None
            print("fp16 CPU run failed:", e)
        meta = dict(kind="dmff", B=B, C=C, H=H, W=W, va=va, ha=ha, loops=loops, seed=seed, bn_eps=1e-5,
                    ref_fp16_self_dev=dev16,
                    reference="models/common.py:762-865 TransformerFusionBlock.eval()", torch=torch.__version__)
        _save(name, meta, out=out.numpy(), tok_vis=tr.numpy(), tok_ir=ti.numpy())

    for name, size, B, H, W in MODEL_CASES:
        cfg = os.path.join(REF_ROOT, "models", "transformer", f"yolov5{size}_Transfusion_kaist.yaml")
        model = yolo.Model(cfg, ch=3, nc=1).eval()
        _load_synth(model, seed)
        rgb, ir = synth.synth_images(B, H, W, seed)
        z, logits, xs = model(rgb, ir)
        fused = yolo.Model(cfg, ch=3, nc=1).eval()
        _load_synth(fused, seed)
        fused.fuse()
        zf = fused(rgb, ir)[0]
        dev16 = None
        if size == "s":
            try:
                z16 = fused.half()(rgb.half(), ir.half())[0].float()
                dev16 = float((z16 - zf).abs().max() / zf.abs().max())
            except Exception as e:  # noqa: BLE001
                print("fp16 CPU run failed:", e)
        meta = dict(kind="model", size=size, B=B, H=H, W=W, seed=seed, ref_fp16_self_dev=dev16,
                    reference="models/yolo_test.py Model(...).eval() forward, plus .fuse() variant",
                    torch=torch.__version__)
        _save(name, meta, z=z.numpy(), z_fused=zf.numpy(), logits=logits.numpy(),
              x0=xs[0].numpy().astype(np.float16), x1=xs[1].numpy().astype(np.float16),
              x2=xs[2].numpy().astype(np.float16))


if __name__ == "__main__":
    main()
