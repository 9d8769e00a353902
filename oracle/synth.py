"""Deterministic "trained-like" synthetic weights and inputs -- TEST INFRASTRUCTURE.

No checkpoints ship with the reference (weights are off-line links, README.md:41-51) and
its default inits are degenerate for testing (CrossAttention Linear std=1e-3 common.py:637,
pos_emb zeros :773-774, coefficients 1.0, LearnableWeights 0.5), which would hide bugs in
softmax / pos-emb / coefficient handling.  Values here are drawn from numpy PCG64 streams
keyed by (seed, parameter name) so the same state_dict can be rebuilt bit-identically in the
build container (to generate golden vectors with the real reference) and on the GPU box
(to replay them) without shipping the weights.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict
from typing import Dict, Mapping, Sequence

import numpy as np
import torch


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))


def synth_tensor(name: str, shape: Sequence[int], seed: int) -> torch.Tensor:
    g = _rng(seed, name)
    shape = tuple(shape)

    def normal(std, mean=0.0):
        return (mean + std * g.standard_normal(shape, dtype=np.float32)).astype(np.float32)

    leaf = name.rsplit(".", 1)[-1]
    if name.endswith("num_batches_tracked"):
        a = np.zeros(shape, dtype=np.int64)
    elif ".bn." in name:
        if leaf == "weight":
            a = normal(0.1, 1.0)
        elif leaf == "bias":
            a = normal(0.1)
        elif leaf == "running_mean":
            a = normal(0.1)
        else:  # running_var
            a = g.uniform(0.5, 1.5, shape).astype(np.float32)
    elif "pos_emb" in name:
        a = normal(0.5)
    elif "_coefficient.w" in name:          # LearnableWeights
        a = normal(0.1, 0.5)
    elif ".coefficient" in name:            # LearnableCoefficient
        a = normal(0.2, 1.0)
    elif any(t in name for t in (".LN1.", ".LN2.", ".ln_input.", ".ln_output.")):
        a = normal(0.1, 1.0) if leaf == "weight" else normal(0.1)
    elif leaf == "weight" and len(shape) == 4:   # conv (Detect's m.* included)
        # gain 1.2 keeps the 100-conv-deep yolov5l stream at O(1) activations (sqrt(2) explodes to
        # 1e4 through the residual C3 stacks and would overflow fp16); Detect's plain Conv2d gets
        # gain 3 so its logits spread over the sigmoid's non-linear range.
        fan_in = shape[1] * shape[2] * shape[3]
        gain = 1.2 if name.endswith(".conv.weight") else 3.0
        a = normal(float(gain / np.sqrt(fan_in)))
    elif leaf == "weight" and len(shape) == 2:   # linear
        a = normal(float(1.0 / np.sqrt(shape[1])))
    elif leaf == "bias":
        a = normal(0.1)
    elif leaf in ("anchors", "anchor_grid"):
        raise KeyError(name)                 # buffers, never synthesised
    else:
        a = normal(0.1)
    return torch.from_numpy(a)


def synth_state_dict(shapes: Mapping[str, Sequence[int]], seed: int) -> "OrderedDict[str, torch.Tensor]":
    """`shapes`: parameter/buffer name -> shape (anchors buffers are skipped)."""
    out = OrderedDict()
    for k, shp in shapes.items():
        if k.endswith("anchors") or k.endswith("anchor_grid"):
            continue
        out[k] = synth_tensor(k, shp, seed)
    return out


def synth_images(B: int, H: int, W: int, seed: int):
    """Two (B,3,H,W) fp32 tensors in [0,1): stand-ins for `/255` RGB and IR frames
    (train.py:295-297, detect_twostream.py:70-80)."""
    g = _rng(seed, "images")
    rgb = g.random((B, 3, H, W), dtype=np.float32)
    ir = g.random((B, 3, H, W), dtype=np.float32)
    return torch.from_numpy(rgb), torch.from_numpy(ir)


def synth_features(B: int, C: int, H: int, W: int, seed: int):
    """Two (B,C,H,W) fp32 feature maps ~ N(0,1) (post-SiLU-like scale) for DMFF tests."""
    g = _rng(seed, "features")
    a = g.standard_normal((B, C, H, W), dtype=np.float32)
    b = g.standard_normal((B, C, H, W), dtype=np.float32)
    return torch.from_numpy(a), torch.from_numpy(b)


def dmff_param_shapes(C: int, N: int, pre: str = "blk", h: int = 8, block_exp: int = 4) -> Dict[str, tuple]:
    """state_dict layout of TransformerFusionBlock (common.py:762-807), dead params included."""
    s: Dict[str, tuple] = OrderedDict()
    s[f"{pre}.pos_emb_vis"] = (1, N, C)
    s[f"{pre}.pos_emb_ir"] = (1, N, C)
    for m in ("vis", "ir"):
        s[f"{pre}.{m}_coefficient.w1"] = (1,)
        s[f"{pre}.{m}_coefficient.w2"] = (1,)
    t = f"{pre}.crosstransformer.0"
    for ln in ("ln_input", "ln_output"):
        s[f"{t}.{ln}.weight"] = (C,)
        s[f"{t}.{ln}.bias"] = (C,)
    for m in ("vis", "ir"):
        for p in ("que", "key", "val"):
            s[f"{t}.crossatt.{p}_proj_{m}.weight"] = (C, C)
            s[f"{t}.crossatt.{p}_proj_{m}.bias"] = (C,)
    for m in ("vis", "ir"):
        s[f"{t}.crossatt.out_proj_{m}.weight"] = (C, C)
        s[f"{t}.crossatt.out_proj_{m}.bias"] = (C,)
    for ln in ("LN1", "LN2"):
        s[f"{t}.crossatt.{ln}.weight"] = (C,)
        s[f"{t}.crossatt.{ln}.bias"] = (C,)
    for m in ("mlp_vis", "mlp_ir", "mlp"):
        s[f"{t}.{m}.0.weight"] = (block_exp * C, C)
        s[f"{t}.{m}.0.bias"] = (block_exp * C,)
        s[f"{t}.{m}.2.weight"] = (C, block_exp * C)
        s[f"{t}.{m}.2.bias"] = (C,)
    for ln in ("LN1", "LN2"):
        s[f"{t}.{ln}.weight"] = (C,)
        s[f"{t}.{ln}.bias"] = (C,)
    for j in range(1, 9):
        s[f"{t}.coefficient{j}.bias"] = (1,)
    s[f"{pre}.conv1x1_out.conv.weight"] = (C, 2 * C, 1, 1)
    for b, shp in (("weight", (C,)), ("bias", (C,)), ("running_mean", (C,)), ("running_var", (C,)),
                   ("num_batches_tracked", ())):
        s[f"{pre}.conv1x1_out.bn.{b}"] = shp
    return s


def model_param_shapes(cfg: dict) -> "OrderedDict[str, tuple]":
    """state_dict layout (names -> shapes) of models/yolo_test.py Model(cfg) for the module
    subset the Transfusion configs use; checked against the real reference by gen_golden."""
    from oracle.icaf_oracle import parse_layers
    s: "OrderedDict[str, tuple]" = OrderedDict()

    def conv(pre, c1, c2, k):
        s[pre + ".conv.weight"] = (c2, c1, k, k)
        for b, shp in (("weight", (c2,)), ("bias", (c2,)), ("running_mean", (c2,)),
                       ("running_var", (c2,)), ("num_batches_tracked", ())):
            s[f"{pre}.bn.{b}"] = shp

    for L in parse_layers(cfg):
        pre, t = f"model.{L['i']}", L["type"]
        if t == "Conv":
            conv(pre, L["c1"], L["c2"], L["args"][0] if L["args"] else 1)
        elif t == "C3":
            c1, c2, n = L["c1"], L["c2"], L["n"]
            c_ = c2 // 2
            conv(pre + ".cv1", c1, c_, 1)
            conv(pre + ".cv2", c1, c_, 1)
            conv(pre + ".cv3", 2 * c_, c2, 1)
            for j in range(n):
                conv(f"{pre}.m.{j}.cv1", c_, c_, 1)
                conv(f"{pre}.m.{j}.cv2", c_, c_, 3)
        elif t == "SPPF":
            c1, c2 = L["c1"], L["c2"]
            conv(pre + ".cv1", c1, c1 // 2, 1)
            conv(pre + ".cv2", 2 * c1, c2, 1)
        elif t == "DMFF":
            s.update(dmff_param_shapes(L["c"], L["va"] * L["ha"], pre))
        elif t == "Detect":
            na = len(L["anchors"][0]) // 2
            for j, c in enumerate(L["ch"]):
                s[f"{pre}.m.{j}.weight"] = (na * (L["nc"] + 5), c, 1, 1)
                s[f"{pre}.m.{j}.bias"] = (na * (L["nc"] + 5),)
    return s
