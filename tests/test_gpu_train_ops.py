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
