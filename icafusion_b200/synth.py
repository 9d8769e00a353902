"""Deterministic "trained-like" synthetic weights and inputs (benchmarks, tests, golden-vector generation).

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


def load_synth(module, seed: int, prefix: str = ""):
    """Fill `module` with the seeded synthetic state_dict for its own parameter names/shapes (Detect anchors untouched)."""
    own = module.state_dict()
    shapes = {prefix + k: tuple(v.shape) for k, v in own.items() if not k.endswith(("anchors", "anchor_grid"))}
    sd = synth_state_dict(shapes, seed)
    res = module.load_state_dict({k[len(prefix):]: v for k, v in sd.items()}, strict=False)
    assert not res.unexpected_keys
    assert all(k.endswith(("anchors", "anchor_grid")) for k in res.missing_keys), res.missing_keys
    return sd
