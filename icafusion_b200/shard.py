"""Batch-dimension sharding of RGB+IR pairs across ranks (one process per GPU).

The hot path has no cross-sample operation in inference (BatchNorm uses running statistics, attention is per
image), so multi-GPU execution is N independent replicas over disjoint slices of the batch -- no collective on
the data path.  The only exchange is the optional gather of the per-rank predictions for a host-side consumer
(NMS / evaluation on rank 0), mirroring how the reference evaluates on rank 0 only (train.py:375-381).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_pairs: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first `n_pairs % world` ranks get one extra pair."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_pairs, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_pairs(rgb: torch.Tensor, ir: torch.Tensor, world: Optional[int] = None, rank: Optional[int] = None):
    """This rank's slice of a global batch of pairs (views, no copy)."""
    world = dist.get_world_size() if world is None else world
    rank = dist.get_rank() if rank is None else rank
    if rgb.shape[0] != ir.shape[0]:
        raise ValueError("rgb and ir batches differ")            # reference asserts the same (common.py:740,812)
    lo, hi = shard_bounds(rgb.shape[0], world, rank)
    return rgb[lo:hi], ir[lo:hi]


def gather_predictions(z_local: torch.Tensor, n_pairs: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Collect per-rank decoded predictions (b_local, rows, no) on `dst` in global batch order.
    Ragged shards (n_pairs % world != 0) are padded to the largest shard for the collective and trimmed after."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(n_pairs, world, r) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    buf = z_local.new_zeros((cap,) + tuple(z_local.shape[1:]))
    buf[: z_local.shape[0]] = z_local
    outs: Optional[List[torch.Tensor]] = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, outs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], 0)
