"""The N>1 host path (batch sharding + prediction gather) on CPU with gloo, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from icafusion_b200.shard import gather_predictions, shard_bounds, shard_pairs


def test_shard_bounds_cover_and_balance():
    for n in (1, 2, 7, 16, 129):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        rgb, ir = torch.rand(n_pairs, 3, 8, 8, generator=g), torch.rand(n_pairs, 3, 8, 8, generator=g)
        a, b = shard_pairs(rgb, ir)
        z_local = torch.stack([a.sum((1, 2, 3)), b.sum((1, 2, 3))], 1).unsqueeze(-1)     # stand-in per-pair "prediction"
        z = gather_predictions(z_local, n_pairs, dst=0)
        if rank == 0:
            want = torch.stack([rgb.sum((1, 2, 3)), ir.sum((1, 2, 3))], 1).unsqueeze(-1)
            q.put(bool(torch.equal(z, want)))
        else:
            assert z is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [4, 5])
def test_shard_and_gather_world2(n_pairs):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
