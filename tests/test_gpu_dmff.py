"""DMFF block (TransformerFusionBlock / CrossTransformerBlock) on the GPU vs
 (a) golden vectors produced by the real reference (tests/golden/dmff_*.npz) and
 (b) the CPU oracle on fresh seeded inputs, incl. BASELINE config 1 and the iterative loops."""
import glob
import os

import pytest
import torch

from conftest import GOLDEN, load_golden
from helpers import err, load_synth
from oracle import icaf_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu

# north_star: the DMFF forward matches the reference "within 1e-3 fp16 relative tolerance" (norm-wise max|a-b|/max|b|,
# SURVEY.md 7.2) -- that is TOL_OUT, applied to the block output against the reference's fp32 result.  (The reference's
# OWN fp16 path deviates 0.8e-3 .. 1.5e-3 from its fp32 path on these cases, meta['ref_fp16_self_dev'].)
# The token streams after the cross transformer are an internal intermediate (fp16 residual stream, rounded twice per
# loop); they are checked against a looser bound so that a semantic bug is still caught early.
TOL_OUT = 1e-3
TOL_TOKENS = 2e-3


NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "dmff_*.npz")))


def _build(m, device):
    from icafusion_b200 import TransformerFusionBlock
    blk = TransformerFusionBlock(m["C"], m["va"], m["ha"]).eval()
    blk.crosstransformer[0].loops = m["loops"]
    load_synth(blk, m["seed"], "blk.")
    return blk.to(device)


@pytest.mark.parametrize("name", NAMES)
def test_dmff_matches_reference_golden(cuda_device, name):
    m, d = load_golden(name)
    blk = _build(m, cuda_device)
    rgb, ir = synth.synth_features(m["B"], m["C"], m["H"], m["W"], m["seed"])
    with torch.no_grad():
        out = blk([rgb.to(cuda_device).half(), ir.to(cuda_device).half()])
        # token streams after the cross transformer (a2-a5 without the tail)
        from icafusion_b200 import ops
        from icafusion_b200.common import to_nhwc
        nh, nw = blk.avgpool.out_size(m["H"], m["W"])
        pv, pi, mix = blk._front()
        r, i = ops.dmff_pool_tokens(to_nhwc(rgb.to(cuda_device).half()), to_nhwc(ir.to(cuda_device).half()), pv, pi, mix, nh, nw)
        r, i = blk.crosstransformer[0].run(r, i, nh * nw)
    torch.cuda.synchronize()
    N = nh * nw
    e_tok = max(err(r[:, :N], d["tok_vis"]), err(i[:, :N], d["tok_ir"]))
    e_out = err(out, d["out"])
    print(f"\n[{name}] tokens {e_tok:.2e}  out {e_out:.2e}  (reference's own fp16 path: {m.get('ref_fp16_self_dev')})")
    assert tuple(out.shape) == d["out"].shape
    assert e_out < TOL_OUT and e_tok < TOL_TOKENS


@pytest.mark.parametrize("B,C,H,W,va,ha,loops", [
    (1, 256, 40, 32, 16, 16, 1),       # BASELINE config 1 read literally as NCHW (H=40, W=32), pooled
    (1, 256, 40, 32, 40, 32, 1),       # ... unpooled N=1280
    (3, 512, 16, 20, 10, 10, 2),       # yolov5s P5 geometry, batch 3, two iterations (weights shared, K/V recomputed)
    (1, 1024, 16, 20, 10, 10, 1),      # yolov5l P5: head_dim 128
    (2, 128, 80, 80, 20, 20, 1),       # 640x640 letterboxed input: P3 80x80 -> k=s=(4,4)
])
def test_dmff_matches_oracle(cuda_device, B, C, H, W, va, ha, loops):
    from icafusion_b200 import TransformerFusionBlock
    seed = 99
    blk = TransformerFusionBlock(C, va, ha).eval()
    blk.crosstransformer[0].loops = loops
    sd = load_synth(blk, seed, "blk.")
    blk = blk.to(cuda_device)
    rgb, ir = synth.synth_features(B, C, H, W, seed)
    with torch.no_grad():
        out = blk([rgb.to(cuda_device).half(), ir.to(cuda_device).half()])
        ref = O.dmff_block(rgb.half().float(), ir.half().float(), sd, "blk", va, ha, loops, bn_eps=1e-5)
    e = err(out, ref)
    print(f"\n[C={C} {H}x{W}->{va}x{ha} L={loops}] {e:.2e}")
    assert e < TOL_OUT


def test_cross_transformer_block_token_api(cuda_device):
    """CrossTransformerBlock.forward([r, i]) with (B,N,C) tokens, N not a multiple of 8 (reference API, common.py:737)."""
    from icafusion_b200 import CrossTransformerBlock
    C, N, B = 128, 100, 2
    blk = CrossTransformerBlock(C, C, C, 8, 4, 0.1, 0.1, loops_num=2).eval()
    sd = load_synth(blk, 5, "t.")
    blk = blk.to(cuda_device)
    g = torch.Generator().manual_seed(0)
    r, i = torch.randn(B, N, C, generator=g).half(), torch.randn(B, N, C, generator=g).half()
    with torch.no_grad():
        o_r, o_i = blk([r.to(cuda_device), i.to(cuda_device)])
        rr, ri = O.cross_transformer_block(r.float(), i.float(), sd, "t", loops=2)
    assert err(o_r, rr) < 2e-3 and err(o_i, ri) < 2e-3


def test_training_mode_takes_the_autograd_path(cuda_device):
    """In train() the module's forward runs the training nodes (batch-statistics BN, dropout, nearest tail; parity in
    tests/test_gpu_train_model.py); the fused inference entry point `run` refuses to run with training semantics pending."""
    from icafusion_b200 import TransformerFusionBlock
    blk = TransformerFusionBlock(128, 10, 10).to(cuda_device).train()
    x = torch.randn(2, 128, 16, 20, device=cuda_device).half()
    y = blk([x, x])
    assert tuple(y.shape) == (2, 128, 16, 20) and y.requires_grad and torch.isfinite(y).all()
    with pytest.raises(NotImplementedError):
        blk.run(x.permute(0, 2, 3, 1).contiguous(), x.permute(0, 2, 3, 1).contiguous())
