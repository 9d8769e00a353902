"""Host side of the training path on CPU: the reference's dead parameters, optimiser groups, and the N>1 path -- DDP over gloo,
world_size 2 -- with a stand-in forward (the CUDA kernels cannot run here; what is under test is the wrap, the dead-parameter
repair and the gradient averaging that train.py:231-235,339 rely on)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden


def test_dead_parameters_and_groups_match_the_reference():
    """The 30 parameters that get no gradient in the REAL reference's training step (golden meta) are exactly the ones
    freeze_dead_parameters removes; optimiser groups follow train.py:124-131."""
    from icafusion_b200 import Model
    from icafusion_b200.trainer import dead_parameters, freeze_dead_parameters, param_groups
    m, _ = load_golden("train_yolov5s_320")
    model = Model("yolov5s_Transfusion_kaist")
    assert sorted(dead_parameters(model)) == sorted(m["dead_params"])
    assert len(m["dead_params"]) == 30
    freeze_dead_parameters(model)
    live = [k for k, p in model.named_parameters() if p.requires_grad]
    assert sorted(live) == sorted(m["params"])
    pg0, pg1, pg2 = param_groups(model)
    n_bn = sum(isinstance(x, torch.nn.BatchNorm2d) for x in model.modules())
    assert len(pg0) == n_bn
    grouped = {id(p) for g in (pg0, pg1, pg2) for p in g}
    left_out = sorted(k for k, p in model.named_parameters() if p.requires_grad and id(p) not in grouped)
    # reference quirk kept: pos_emb_* and LearnableWeights.w1/w2 are neither `.weight` nor `.bias` of a module -> in no group
    assert all(k.split(".")[-1] in ("pos_emb_vis", "pos_emb_ir", "w1", "w2") for k in left_out) and len(left_out) == 3 * 6


def _scenario(rank, world, freeze):
    from icafusion_b200 import Model
    from icafusion_b200 import trainer
    model = Model("yolov5s_Transfusion_kaist")
    keep = trainer.freeze_dead_parameters
    if not freeze:
        trainer.freeze_dead_parameters = lambda m: []                      # the reference as written (train.py:233)
    try:
        ts = trainer.TrainStep(model, None, total_batch_size=4, world_size=world, local_rank=None, imgsz=320)
    finally:
        trainer.freeze_dead_parameters = keep
    live = [p for p in model.parameters() if p.requires_grad]
    dead = set(trainer.dead_parameters(model))
    used = [p for k, p in model.named_parameters() if k not in dead]

    def stand_in(rgb, ir):                                                 # touches exactly what the real forward touches
        return sum(p.sum() for p in used) * float(rank + 1)
    model.forward = stand_in
    try:
        for _ in range(2):                                                 # the reference's DDP failure shows up in iteration 2
            loss = ts.model(None, None) * world                            # train.py:339
            loss.backward()
            g = [p.grad.clone() for p in live if p.grad is not None]
            ts.zero_grad()
    except RuntimeError as e:
        return ("error", str(e)[:80])
    # mean over ranks of world * (rank + 1) = world * (world + 1) / 2 for every element
    want = world * (world + 1) / 2
    bad = [(float(x.min()), float(x.max())) for x in g if not torch.allclose(x, torch.full_like(x, want))]
    return ("ok", (len(bad), len(g), len(live), bad[:3]))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = [_scenario(rank, world, True), _scenario(rank, world, False)]
        if rank == 0:
            q.put(res)
    finally:
        dist.destroy_process_group()


def test_ddp_world2_gradient_average_and_dead_parameter_repair():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(400)
        assert p.exitcode == 0
    (kind, val), (kind2, val2) = q.get(timeout=5)
    assert kind == "ok" and val[0] == 0 and val[1] == val[2], val          # repaired: averaged gradients on every live parameter
    # without the repair DDP refuses the second iteration: parameters that never got a gradient
    assert kind2 == "error" and "Expected to have finished reduction" in val2, (kind2, val2)


def test_model_ema_matches_the_reference_loop():
    """ModelEMA.update == the reference's per-tensor loop (utils/torch_utils.py:305-315), on a small DMFF block."""
    import math
    from copy import deepcopy
    from icafusion_b200.common import TransformerFusionBlock
    from icafusion_b200.trainer import ModelEMA
    torch.manual_seed(0)
    m = TransformerFusionBlock(64, 4, 4)
    ema = ModelEMA(m)
    ref = {k: v.clone() for k, v in deepcopy(m).state_dict().items()}
    for step in range(1, 4):
        with torch.no_grad():
            for p in m.parameters():
                p.add_(0.1 * torch.randn_like(p))
            m.conv1x1_out.bn.running_mean.add_(0.05)
            m.conv1x1_out.bn.num_batches_tracked += 1
        ema.update(m)
        d = 0.9999 * (1 - math.exp(-step / 2000))
        for k, v in m.state_dict().items():
            if v.dtype.is_floating_point:
                ref[k] = ref[k] * d + (1.0 - d) * v
    assert ema.updates == 3
    for k, v in ema.ema.state_dict().items():
        if v.dtype.is_floating_point:
            assert torch.allclose(v, ref[k], rtol=1e-6, atol=1e-7), k
        else:
            assert torch.equal(v, ref[k]), k              # integer buffers are left alone, like in the reference
