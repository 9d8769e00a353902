"""Shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations

import numpy as np
import torch

from oracle import synth


def load_synth(module: torch.nn.Module, seed: int, prefix: str = ""):
    """Fill `module` with the seeded synthetic state_dict (same values gen_golden loaded into the reference)."""
    own = module.state_dict()
    shapes = {prefix + k: tuple(v.shape) for k, v in own.items() if not k.endswith(("anchors", "anchor_grid"))}
    sd = synth.synth_state_dict(shapes, seed)
    res = module.load_state_dict({k[len(prefix):]: v for k, v in sd.items()}, strict=False)
    assert not res.unexpected_keys
    assert all(k.endswith(("anchors", "anchor_grid")) for k in res.missing_keys), res.missing_keys
    return sd


def nhwc(x_nchw: torch.Tensor) -> torch.Tensor:
    return x_nchw.permute(0, 2, 3, 1).contiguous()


def nchw(x_nhwc: torch.Tensor) -> torch.Tensor:
    return x_nhwc.permute(0, 3, 1, 2).contiguous()


def err(a, b) -> float:
    """max|a-b| / max|b|"""
    a = a.detach().float().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float32)
    b = b.detach().float().cpu().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float32)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.isfinite(a).all(), "non-finite values in the CUDA result"
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / max(np.abs(b).max(), 1e-30))
