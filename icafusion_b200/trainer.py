"""The training step of train.py:117-349 around the icafusion_b200 model: optimiser groups, DDP wrap, GradScaler, one step.

This is the caller side of the hot path: what the reference's ``train_rgb_ir`` does between building
the model and ``ema.update`` -- minus data loading, logging, checkpoints and evaluation.  The model's forward and backward are
the autograd nodes of ``icafusion_b200.autograd`` (our kernels); the optimiser, the GradScaler and DistributedDataParallel
are torch's, exactly the objects train.py constructs, so the one exchange step of the data-parallel path is DDP's bucketed
NCCL all-reduce of the gradients (481 MB fp32 for yolov5l), overlapped with the rest of the backward pass.

Two defects of the reference's multi-GPU mode are repaired here, outside the model code:
  * 30 DMFF parameters per model never receive a gradient (``ln_input``, ``ln_output``, the block-level ``mlp`` and ``LN1``
    of every CrossTransformerBlock: common.py:701-702,716-721,724) while train.py:233 builds DDP with
    ``find_unused_parameters=False`` -> DDP raises in the second iteration.  ``freeze_dead_parameters`` takes them out of the
    graph (``requires_grad_(False)``) before the wrap; they stay in the state_dict.
  * train.py:583-588 only calls ``train_rgb_ir`` on rank 0; ``TrainStep`` is constructed and called on every rank.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .common import CrossTransformerBlock
from .loss import ComputeLoss

HYP_SCRATCH = dict(lr0=0.01, momentum=0.937, weight_decay=0.0005, box=0.05, cls=0.5, cls_pw=1.0, obj=1.0, obj_pw=1.0, anchor_t=4.0,
                   fl_gamma=0.0)      # data/hyp.scratch.yaml (the keys the step reads)


def dead_parameters(model: nn.Module) -> List[str]:
    """Names of the parameters the reference's forward never touches (they get no gradient in the reference either)."""
    names = []
    for mod_name, m in model.named_modules():
        if isinstance(m, CrossTransformerBlock):
            for sub in ("ln_input", "ln_output", "mlp", "LN1"):
                for k, _ in getattr(m, sub).named_parameters():
                    names.append(f"{mod_name}.{sub}.{k}" if mod_name else f"{sub}.{k}")
    return names


def freeze_dead_parameters(model: nn.Module) -> List[str]:
    names = dead_parameters(model)
    params = dict(model.named_parameters())
    for k in names:
        params[k].requires_grad_(False)
    return names


def param_groups(model: nn.Module):
    """train.py:124-131: (BatchNorm weights [no decay], other weights [decay], biases), trainable parameters only."""
    pg0, pg1, pg2 = [], [], []
    for _, v in model.named_modules():
        if hasattr(v, "bias") and isinstance(v.bias, nn.Parameter) and v.bias.requires_grad:
            pg2.append(v.bias)
        if isinstance(v, nn.BatchNorm2d):
            if v.weight.requires_grad:
                pg0.append(v.weight)
        elif hasattr(v, "weight") and isinstance(v.weight, nn.Parameter) and v.weight.requires_grad:
            pg1.append(v.weight)
    # (like the reference, parameters that are neither .weight nor .bias -- pos_emb, LearnableWeights.w1/w2 -- join no group;
    #  LearnableCoefficient.bias lands in the bias group)
    return pg0, pg1, pg2


class TrainStep:
    """model -> (optional DDP) -> loss -> scaled backward -> optimiser step, per call (train.py:334-349)."""

    def __init__(self, model: nn.Module, hyp: Optional[Dict[str, float]] = None, total_batch_size: int = 64, world_size: int = 1,
                 local_rank: Optional[int] = None, imgsz: int = 640, amp_scale: bool = True):
        hyp = dict(HYP_SCRATCH if hyp is None else hyp)
        det = model.model[-1]
        nl, nc = det.nl, det.nc
        nbs = 64
        accumulate = max(round(nbs / total_batch_size), 1)
        hyp["weight_decay"] *= total_batch_size * accumulate / nbs                    # train.py:121
        self.dead = freeze_dead_parameters(model)
        pg0, pg1, pg2 = param_groups(model)
        self.optimizer = torch.optim.SGD(pg0, lr=hyp["lr0"], momentum=hyp["momentum"], nesterov=True)      # train.py:136
        self.optimizer.add_param_group({"params": pg1, "weight_decay": hyp["weight_decay"]})
        self.optimizer.add_param_group({"params": pg2})
        hyp["box"] *= 3.0 / nl                                                        # train.py:238-240
        hyp["cls"] *= nc / 80.0 * 3.0 / nl
        hyp["obj"] *= (imgsz / 640) ** 2 * 3.0 / nl
        model.nc, model.hyp, model.gr = nc, hyp, 1.0                                  # train.py:242-244
        self.raw_model = model
        self.world_size = world_size
        if world_size > 1:
            from torch.nn.parallel import DistributedDataParallel as DDP
            model = DDP(model, device_ids=[local_rank], output_device=local_rank, find_unused_parameters=False)   # train.py:233
        self.model = model
        self.scaler = torch.amp.GradScaler("cuda", enabled=amp_scale)                 # train.py:282
        self.compute_loss = ComputeLoss(self.raw_model)                               # train.py:284
        self.hyp = hyp

    def __call__(self, rgb: torch.Tensor, ir: torch.Tensor, targets: torch.Tensor):
        """rgb / ir: (B,3,H,W) uint8 (scaled by 1/255 inside the stem staging, train.py:297-298) or float images on the device;
        targets (nt, 6).  Returns (loss, loss_items) of this rank."""
        pred = self.model(rgb, ir)                                                    # train.py:336
        loss, items = self.compute_loss(pred, targets)                                # train.py:337
        if self.world_size > 1:
            loss = loss * self.world_size                                             # train.py:339
        self.scaler.scale(loss).backward()                                            # train.py:344
        self.scaler.step(self.optimizer)                                              # train.py:348-350
        self.scaler.update()
        self.optimizer.zero_grad(set_to_none=True)
        return loss.detach(), items
