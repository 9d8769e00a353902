"""Training-step building blocks (operator level): weight / data / bias gradients of the Conv2d and Linear layers of the
hot path against torch autograd (CPU, fp32, on the same fp16-rounded operands)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import err, nhwc

pytestmark = pytest.mark.gpu


def _grads(x, w, dy, s, p):
    x = x.float().requires_grad_(True)
    w = w.float().requires_grad_(True)
    y = F.conv2d(x, w, None, stride=s, padding=p)
    y.backward(dy.float())
    return x.grad, w.grad


@pytest.mark.parametrize("B,Cin,H,W,Cout,k,s,p", [
    (2, 64, 16, 20, 64, 1, 1, 0), (2, 64, 32, 40, 128, 3, 1, 1), (3, 128, 16, 20, 256, 3, 2, 1), (1, 256, 16, 20, 128, 1, 1, 0),
    (4, 64, 33, 47, 64, 3, 1, 1), (2, 192, 16, 24, 64, 3, 1, 1), (8, 16, 64, 80, 32, 3, 1, 1), (2, 32, 32, 40, 64, 3, 2, 1),
    (16, 128, 64, 80, 128, 3, 1, 1), (2, 512, 16, 20, 24, 1, 1, 0)])
def test_conv_weight_and_data_gradients(cuda_device, B, Cin, H, W, Cout, k, s, p):
    from icafusion_b200 import ops
    g = torch.Generator().manual_seed(B * 100 + Cin + k)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    x = torch.randn(B, Cin, H, W, generator=g).half()
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).half()
    dy = (torch.randn(B, Cout, Ho, Wo, generator=g) * 0.1).half()
    dx_ref, dw_ref = _grads(x, w, dy, s, p)
    xv, dyv = nhwc(x).to(cuda_device), nhwc(dy).to(cuda_device)
    dw = ops.conv2d_wgrad(xv, dyv, k, k, s, p)
    acc = dw.clone()
    ops.conv2d_wgrad(xv, dyv, k, k, s, p, scale=0.5, out=acc)            # accumulate with a scale: .grad += 0.5 dW
    dx = ops.conv2d_dgrad(dyv, w.float().to(cuda_device), s, p, (H, W))
    torch.cuda.synchronize()
    e_w = err(dw, dw_ref)
    print(f"\n[grads B{B} {Cin}->{Cout} {H}x{W} k{k}s{s}] dW {e_w:.2e}" + ("" if dx is None else f"  dX {err(dx.permute(0, 3, 1, 2), dx_ref):.2e}"))
    assert e_w < 1e-3
    assert err(acc, dw_ref * 1.5) < 1e-3
    if dx is not None:
        assert err(dx.permute(0, 3, 1, 2), dx_ref) < 1.5e-3          # fp16 output
    dw2 = ops.conv2d_wgrad(xv, dyv, k, k, s, p)
    assert torch.equal(dw, dw2), "wgrad is not deterministic"


def test_linear_and_bias_gradients(cuda_device):
    from icafusion_b200 import ops
    g = torch.Generator().manual_seed(9)
    for rows, K, N in ((400, 256, 768), (6400, 256, 1024), (1664, 1024, 64), (104, 4096, 1024)):
        x = torch.randn(rows, K, generator=g).half()
        dy = (torch.randn(rows, N, generator=g) * 0.1).half()
        dw = ops.linear_wgrad(x.to(cuda_device), dy.to(cuda_device))
        db = ops.colsum(dy.to(cuda_device))
        torch.cuda.synchronize()
        assert err(dw, dy.float().t() @ x.float()) < 1e-3, (rows, K, N)
        assert err(db, dy.float().sum(0)) < 1e-4


def test_batchnorm_silu_training_forward_backward(cuda_device):
    """Conv.forward in training mode after the convolution (models/common.py:56-57): BatchNorm2d with batch statistics + SiLU,
    forward (incl. the running-statistics update) and backward vs torch autograd."""
    from icafusion_b200 import ops
    g = torch.Generator().manual_seed(4)
    for B, C, H, W, act in ((4, 64, 32, 40, 1), (2, 256, 16, 20, 1), (16, 128, 64, 80, 1), (3, 64, 9, 11, 0)):
        x = (torch.randn(B, C, H, W, generator=g) * 1.3 + 0.2).half()
        dy = (torch.randn(B, C, H, W, generator=g) * 0.1).half()
        gamma, beta = 1 + 0.2 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g)
        bn = torch.nn.BatchNorm2d(C, eps=1e-3, momentum=0.03)
        with torch.no_grad():
            bn.weight.copy_(gamma); bn.bias.copy_(beta)
        xr = x.float().requires_grad_(True)
        yr = bn(xr)
        yr = F.silu(yr) if act else yr
        yr.backward(dy.float())
        rm, rv = torch.zeros(C, device=cuda_device), torch.ones(C, device=cuda_device)
        gd, bd = gamma.to(cuda_device), beta.to(cuda_device)
        xv = nhwc(x).to(cuda_device)
        y, sm, si = ops.bn_act_fwd(xv, gd, bd, rm, rv, 1e-3, 0.03, act)
        dg, db = torch.zeros(C, device=cuda_device), torch.zeros(C, device=cuda_device)
        dx = ops.bn_act_bwd(xv, nhwc(dy).to(cuda_device), gd, bd, sm, si, act, dg, db)
        torch.cuda.synchronize()
        assert err(y.permute(0, 3, 1, 2), yr) < 1e-3
        assert err(rm, bn.running_mean) < 1e-4 and err(rv, bn.running_var) < 1e-4
        assert err(dx.permute(0, 3, 1, 2), xr.grad) < 2e-3
        assert err(dg, bn.weight.grad) < 1e-3 and err(db, bn.bias.grad) < 1e-3


def test_small_training_kernels(cuda_device):
    """GELU forward/backward, LayerNorm backward, dot products, nearest-upsample and SPPF max-pool backward, dropout."""
    from icafusion_b200 import ops
    g = torch.Generator().manual_seed(6)
    x = torch.randn(400, 256, generator=g).half()
    dy = (torch.randn(400, 256, generator=g) * 0.1).half()
    xr = x.float().requires_grad_(True)
    F.gelu(xr).backward(dy.float())
    xd, dyd = x.to(cuda_device), dy.to(cuda_device)
    assert err(ops.eltwise(0, xd), F.gelu(x.float())) < 1e-3
    assert err(ops.eltwise(1, xd, dyd), xr.grad) < 1e-3
    gamma, beta = 1 + 0.2 * torch.randn(256, generator=g), 0.2 * torch.randn(256, generator=g)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    F.layer_norm(xr, (256,), gr, br, 1e-5).backward(dy.float())
    dg, db = torch.zeros(256, device=cuda_device), torch.zeros(256, device=cuda_device)
    dx = ops.layernorm_bwd(xd, dyd, gamma.to(cuda_device), 1e-5, dg, db)
    torch.cuda.synchronize()
    assert err(dx, xr.grad) < 1.5e-3 and err(dg, gr.grad) < 1e-3 and err(db, br.grad) < 1e-3
    assert abs(float(ops.dot(xd, dyd)) - float((x.float() * dy.float()).sum())) < 1e-2
    # nearest 2x up-sampling and one SPPF pool stage
    m = torch.randn(2, 64, 16, 20, generator=g).half()
    d2 = torch.randn(2, 64, 32, 40, generator=g).half()
    mr = m.float().requires_grad_(True)
    F.interpolate(mr, scale_factor=2, mode="nearest").backward(d2.float())
    assert err(ops.upsample2x_bwd(nhwc(d2).to(cuda_device)).permute(0, 3, 1, 2), mr.grad) < 1e-3
    d1 = torch.randn(2, 64, 16, 20, generator=g).half()
    mr = m.float().requires_grad_(True)
    F.max_pool2d(mr, 5, 1, 2).backward(d1.float())
    assert err(ops.maxpool5_bwd(nhwc(m).to(cuda_device), nhwc(d1).to(cuda_device)).permute(0, 3, 1, 2), mr.grad) < 1e-3
    # dropout: keeps ~ (1 - p), scales by 1 / (1 - p), same mask for the same seed
    ones = torch.ones(1 << 16, dtype=torch.float16, device=cuda_device)
    a, b2, c2 = ops.eltwise(2, ones, p=0.1, seed=7), ops.eltwise(2, ones, p=0.1, seed=7), ops.eltwise(2, ones, p=0.1, seed=8)
    keep = float((a > 0).float().mean())
    assert torch.equal(a, b2) and not torch.equal(a, c2) and abs(keep - 0.9) < 0.01 and abs(float(a.max()) - 1 / 0.9) < 1e-3
