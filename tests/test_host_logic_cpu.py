"""Host-side logic that needs no GPU: config generation, graph building, state_dict parity with the reference,
filter packing layout, BN folding, pooling geometry."""
import os

import pytest
import torch

from icafusion_b200 import Model, TransformerFusionBlock, ops
from icafusion_b200.cfg import load_cfg, transfusion_kaist_cfg
from icafusion_b200.common import AdaptivePool2d, Conv
from icafusion_b200.yolo_test import fuse_conv_and_bn
from oracle import icaf_oracle as O
from oracle import synth
from oracle.ref_shim import REF_ROOT, reference_available


@pytest.mark.parametrize("size", ["s", "l"])
def test_state_dict_layout_matches_reference(size):
    """Keys/shapes equal oracle.synth.model_param_shapes, which gen_golden loaded into the real reference with strict
    matching -- so reference checkpoints load here with strict=True."""
    if size == "l":
        with torch.device("meta"):
            m = Model(f"yolov5{size}_Transfusion_kaist")
    else:
        m = Model(f"yolov5{size}_Transfusion_kaist")
    own = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.endswith(("anchors", "anchor_grid"))}
    want = {k: tuple(v) for k, v in synth.model_param_shapes(load_cfg(f"yolov5{size}_Transfusion_kaist")).items()}
    assert own == want
    assert m._ir_start == 10 and len(m.model) == 38
    n_params = sum(p.numel() for p in m.parameters())
    assert abs(n_params / 1e6 - (23.26 if size == "s" else 120.25)) < 0.01


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("size", ["s", "l"])
def test_generated_cfg_equals_reference_yaml(size):
    import yaml
    with open(os.path.join(REF_ROOT, "models", "transformer", f"yolov5{size}_Transfusion_kaist.yaml")) as f:
        ref = yaml.safe_load(f)
    mine = transfusion_kaist_cfg(size)
    for k in ("nc", "depth_multiple", "width_multiple", "anchors", "backbone", "head"):
        assert mine[k] == ref[k], k


def test_dmff_block_state_dict_and_defaults():
    blk = TransformerFusionBlock(128, 10, 10)
    assert {f"blk.{k}": tuple(v.shape) for k, v in blk.state_dict().items()} == \
        {k: tuple(v) for k, v in synth.dmff_param_shapes(128, 100, "blk").items()}
    assert blk.crosstransformer[0].loops == 1 and blk.crosstransformer[0].crossatt.h == 8   # common.py:691,763


def test_pack_conv_weight_layout():
    w = torch.randn(20, 16, 3, 3)
    pk = ops.pack_conv_weight(w, torch.zeros(20), 1, 1, ops.ACT_SILU)
    assert pk.w.shape == (32, 192) and pk.cin == 16 and pk.cout == 20
    # K order is (ky, kx, c): element [n, (ky*3+kx)*16 + c] == w[n, c, ky, kx]
    assert torch.equal(pk.w[:20, :144].float().view(20, 3, 3, 16), w.half().float().permute(0, 2, 3, 1))
    assert float(pk.w[20:].abs().max()) == 0 and float(pk.w[:, 144:].abs().max()) == 0
    stem = ops.pack_conv_weight(torch.randn(32, 3, 6, 6), None, 2, 2, ops.ACT_SILU)
    assert stem.cin == 4 and stem.w.shape == (32, 192)
    assert float(stem.w[:, :144].view(32, 36, 4)[..., 3].abs().max()) == 0        # padded 4th input channel


def test_fuse_conv_and_bn_matches_oracle_fold():
    c = Conv(16, 24, 3, 1).eval()
    sd = synth.synth_state_dict({f"m.{k}": tuple(v.shape) for k, v in c.state_dict().items()}, 1)
    c.load_state_dict({k[2:]: v for k, v in sd.items()})
    c.bn.eps = 1e-3
    f = fuse_conv_and_bn(c.conv, c.bn)
    folded = O.fold_bn(sd, 1e-3)
    assert torch.allclose(f.weight, folded["m.conv.weight"], atol=1e-6)
    assert torch.allclose(f.bias, folded["m.conv.bias"], atol=1e-6)


def test_adaptive_pool_geometry():
    p = AdaptivePool2d(20, 20)
    assert p.out_size(64, 80) == (20, 20) and p.out_size(16, 20) == (16, 20) and p.out_size(20, 20) == (20, 20)
    with pytest.raises(ValueError):
        AdaptivePool2d(16, 16).out_size(8, 20)        # the reference divides by zero here (common.py:880)


def test_model_fuse_removes_bn_and_rebinds_forward():
    m = Model("yolov5s_Transfusion_kaist").eval().fuse()
    convs = [x for x in m.modules() if type(x) is Conv]
    assert convs and not any(hasattr(c, "bn") for c in convs) and all(c.conv.bias is not None for c in convs)
    assert not any(".bn." in k for k in m.state_dict())
