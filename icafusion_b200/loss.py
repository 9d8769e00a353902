"""Detection loss with the reference's name and call signature (utils/loss.py:325-463), forward on the device.

``ComputeLoss(model)(p, targets)`` returns ``(loss * batch_size, cat(lbox, lobj, lcls, lrk))`` like the reference.  This is
the use test.py:132-133 makes of it (validation loss from the Detect training outputs); the returned tensors carry no
autograd graph -- the backward pass of the training step is not built in icafusion_b200.  Target
assignment, CIoU, both BCE terms and the reductions run as three kernels of libicaf_b200 (csrc/loss.cu) with no host sync.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch

from . import _lib, ops


def smooth_BCE(eps: float = 0.1):
    """reference: utils/loss.py:15-17"""
    return 1.0 - 0.5 * eps, 0.5 * eps


class ComputeLoss:
    def __init__(self, model, autobalance: bool = False):
        if autobalance:
            raise NotImplementedError("ComputeLoss: autobalance needs a host read per level and step; not built")
        m = model.module if hasattr(model, "module") else model
        h = m.hyp                                            # train.py:229 attaches the hyper-parameter dict to the model
        if h.get("fl_gamma", 0.0) > 0:
            raise NotImplementedError("ComputeLoss: focal loss (fl_gamma > 0) is not built")
        det = m.model[-1]
        self.na, self.nc, self.nl = det.na, det.nc, det.nl
        self.anchors = det.anchors.detach().float().cpu().reshape(-1).tolist()      # grid units, (nl, na, 2)
        cp, cn = smooth_BCE(eps=h.get("label_smoothing", 0.0))
        balance = {3: [4.0, 1.0, 0.4]}.get(det.nl, [4.0, 1.0, 0.25, 0.06, 0.02])    # loss.py:346
        self.hyp = _lib.LossHyp(float(h["box"]), float(h["obj"]), float(h["cls"]), float(h["cls_pw"]), float(h["obj_pw"]),
                                float(h["anchor_t"]), float(h.get("fl_gamma", 0.0)), float(getattr(m, "gr", 1.0)), cp, cn,
                                (C.c_float * 5)(*(balance + [0.0] * 5)[:5]))

    def __call__(self, p: Sequence[torch.Tensor], targets: torch.Tensor):
        nl = len(p)
        if nl != self.nl:
            raise ValueError(f"ComputeLoss: {nl} prediction levels, Detect has {self.nl}")
        dev = p[0].device
        if dev.type != "cuda":
            raise RuntimeError("icafusion_b200 runs on CUDA tensors only (no CPU fallback)")
        dt = p[0].dtype
        if dt not in (torch.float16, torch.float32) or any(t.dtype != dt for t in p):
            raise ValueError("ComputeLoss: predictions must all be fp16 or all fp32")
        ps = [t.detach().contiguous() for t in p]
        B, na, _, _, no = ps[0].shape
        if na != self.na or no != self.nc + 5:
            raise ValueError(f"ComputeLoss: predictions are (B, {na}, ny, nx, {no}), Detect has na={self.na}, nc={self.nc}")
        tg = targets.detach().to(dev, torch.float32).contiguous()
        nt = int(tg.shape[0])
        ny = (C.c_int * nl)(*[t.shape[2] for t in ps])
        nx = (C.c_int * nl)(*[t.shape[3] for t in ps])
        need = int(_lib.lib().icaf_loss_workspace_bytes(B, na, nt, ny, nx, nl))
        ws = torch.empty((need + 7) // 8, dtype=torch.int64, device=dev)
        out = torch.empty(5, dtype=torch.float32, device=dev)
        ptrs = (C.c_void_p * nl)(*[t.data_ptr() for t in ps])
        anch = (C.c_float * len(self.anchors))(*self.anchors)
        ops._call("icaf_compute_loss_fwd", _lib.lib().icaf_compute_loss_fwd,
                  (ptrs, 1 if dt == torch.float32 else 0, ny, nx, nl, B, na, no, ops._ptr(tg if nt else None), nt, anch, C.byref(self.hyp),
                   ops._ptr(out), ops._ptr(ws), C.c_size_t(ws.numel() * 8)), {"bytes": float(sum(t.numel() * t.element_size() for t in ps))})
        return out[0:1], out[1:5]
