"""HBM-bound helper kernels vs the CPU oracle (torch fp32 ops on the same fp16-rounded inputs)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import err, nchw, nhwc
from oracle import icaf_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1.2e-3


@pytest.mark.parametrize("dtype,scale", [(torch.float16, 1.0), (torch.float32, 1.0), (torch.uint8, 1 / 255.0)])
def test_pack_image(cuda_device, dtype, scale):
    from icafusion_b200 import ops
    g = torch.Generator().manual_seed(0)
    img = (torch.rand(2, 3, 10, 12, generator=g) * (255 if dtype == torch.uint8 else 1)).to(dtype)
    out = ops.pack_image(img.to(cuda_device), scale).cpu()
    assert out.shape == (2, 10, 12, 4)
    ref = (img.float() * scale).half()
    assert torch.equal(out[..., :3], ref.permute(0, 2, 3, 1)) and float(out[..., 3].abs().max()) == 0


def test_pack_image_space_to_depth(cuda_device):
    from icafusion_b200 import ops
    img = (torch.rand(2, 3, 12, 16) * 255).to(torch.uint8)
    out = ops.pack_image(img.to(cuda_device), 1 / 255.0, s2d=True).cpu()
    assert out.shape == (2, 6, 8, 16)
    ref = (img.float() / 255.0).half()                                    # (b,c,H,W)
    ref4 = torch.cat([ref, torch.zeros(2, 1, 12, 16, dtype=torch.float16)], 1)
    want = ref4.view(2, 4, 6, 2, 8, 2).permute(0, 2, 4, 3, 5, 1).reshape(2, 6, 8, 16)    # channel = (dy*2+dx)*4 + c
    assert torch.equal(out, want)


@pytest.mark.parametrize("H,W", [(16, 20), (20, 20), (36, 40)])      # <=1024 px: smem kernel; 36x40: direct-window kernel
def test_sppf_pool_bit_exact(cuda_device, H, W):
    from icafusion_b200 import ops
    x = torch.randn(2, 64, H, W).half()
    cat = torch.zeros(2, H, W, 256, dtype=torch.float16, device=cuda_device)
    cat[..., :64] = nhwc(x).to(cuda_device)
    ops.sppf_pool(cat[..., :64], cat[..., 64:128], cat[..., 128:192], cat[..., 192:])
    y1 = F.max_pool2d(x.float(), 5, 1, 2); y2 = F.max_pool2d(y1, 5, 1, 2); y3 = F.max_pool2d(y2, 5, 1, 2)   # common.py:259-266
    ref = torch.cat([x.float(), y1, y2, y3], 1)
    assert torch.equal(nchw(cat).float().cpu(), ref)


def test_upsample_and_concat_bit_exact(cuda_device):
    from icafusion_b200 import ops
    from icafusion_b200.common import Concat
    x = torch.randn(2, 32, 5, 7).half()
    up = ops.upsample2x(nhwc(x).to(cuda_device))
    assert torch.equal(nchw(up).cpu(), F.interpolate(x.float(), scale_factor=2.0, mode="nearest").half())
    y = torch.randn(2, 16, 10, 14).half()
    cat = Concat.run([up, nhwc(y).to(cuda_device)])
    assert torch.equal(nchw(cat).cpu(), torch.cat([nchw(up).cpu(), y], 1))


@pytest.mark.parametrize("H,W,nh,nw", [(64, 80, 20, 20), (32, 40, 16, 16), (16, 20, 10, 10), (16, 20, 16, 20),
                                       (40, 40, 20, 20), (20, 20, 16, 16)])
def test_dmff_pool_tokens(cuda_device, H, W, nh, nw):
    from icafusion_b200 import ops
    B, C = 2, 64
    g = torch.Generator().manual_seed(1)
    xv, xi = torch.randn(B, C, H, W, generator=g).half(), torch.randn(B, C, H, W, generator=g).half()
    N = nh * nw
    sd = {"b.pos_emb_vis": torch.randn(1, N, C, generator=g).half().float(), "b.pos_emb_ir": torch.randn(1, N, C, generator=g).half().float(),
          "b.vis_coefficient.w1": torch.tensor([0.6]), "b.vis_coefficient.w2": torch.tensor([0.3]),
          "b.ir_coefficient.w1": torch.tensor([0.45]), "b.ir_coefficient.w2": torch.tensor([0.7])}
    mix = torch.tensor([0.6, 0.3, 0.45, 0.7], device=cuda_device)
    tv, ti = ops.dmff_pool_tokens(nhwc(xv).to(cuda_device), nhwc(xi).to(cuda_device), sd["b.pos_emb_vis"][0].half().to(cuda_device),
                                  sd["b.pos_emb_ir"][0].half().to(cuda_device), mix, nh, nw)
    rv, a, b = O.dmff_tokens(xv.float(), sd, "b", "vis", nh, nw)
    ri, _, _ = O.dmff_tokens(xi.float(), sd, "b", "ir", nh, nw)
    assert (a, b) == (nh, nw)
    assert err(tv[:, :N], rv) < TOL and err(ti[:, :N], ri) < TOL
    assert float(tv[:, N:].abs().max() if tv.shape[1] > N else 0) == 0


@pytest.mark.parametrize("C", [128, 512, 1024])
def test_layernorm(cuda_device, C):
    from icafusion_b200 import ops
    g = torch.Generator().manual_seed(2)
    x0, x1 = (torch.randn(3, 104, C, generator=g) * 2 + 0.5).half(), torch.randn(3, 104, C, generator=g).half()
    g0, b0, g1, b1 = (torch.randn(C, generator=g) for _ in range(4))
    y0, y1 = ops.layernorm(x0.to(cuda_device), g0.to(cuda_device), b0.to(cuda_device), x1.to(cuda_device), g1.to(cuda_device),
                           b1.to(cuda_device))
    assert err(y0, F.layer_norm(x0.float(), (C,), g0, b0, 1e-5)) < TOL
    assert err(y1, F.layer_norm(x1.float(), (C,), g1, b1, 1e-5)) < TOL


@pytest.mark.parametrize("H,W,nh,nw,mode", [(64, 80, 20, 20, 0), (32, 40, 16, 16, 0), (16, 20, 10, 10, 0), (16, 20, 16, 20, 0),
                                            (64, 80, 20, 20, 1), (16, 20, 10, 10, 1)])
def test_dmff_upsample_cat(cuda_device, H, W, nh, nw, mode):
    from icafusion_b200 import ops
    B, C = 2, 64
    N, n_pad = nh * nw, ops.round_up(nh * nw, 8)
    g = torch.Generator().manual_seed(4)
    tv, ti = torch.randn(B, n_pad, C, generator=g).half(), torch.randn(B, n_pad, C, generator=g).half()
    xv, xi = torch.randn(B, C, H, W, generator=g).half(), torch.randn(B, C, H, W, generator=g).half()
    out = ops.dmff_upsample_cat(tv.to(cuda_device), ti.to(cuda_device), nhwc(xv).to(cuda_device), nhwc(xi).to(cuda_device), nh, nw, mode)

    def up(t):   # common.py:827-837
        t = t[:, :N].float().reshape(B, nh, nw, C).permute(0, 3, 1, 2)
        return F.interpolate(t, size=(H, W), mode="nearest" if mode else "bilinear")
    ref = torch.cat([up(tv) + xv.float(), up(ti) + xi.float()], 1)
    assert err(nchw(out), ref) < TOL


def test_detect_decode(cuda_device):
    from icafusion_b200 import ops
    B, ny, nx, na, no = 2, 8, 10, 3, 6
    g = torch.Generator().manual_seed(5)
    p = (torch.randn(B, ny, nx, na * no, generator=g) * 2).half()
    anchors = [10., 13., 16., 30., 33., 23.]
    z = torch.zeros(B, 50 + na * ny * nx, no, dtype=torch.float16, device=cuda_device)
    lg = torch.zeros(B, 50 + na * ny * nx, no - 5, dtype=torch.float16, device=cuda_device)
    x = ops.detect_decode(p.to(cuda_device), na, no, z, lg, 50, 16.0, anchors)
    # oracle: yolo_test.py:49-65 on the same conv output
    xr = p.float().permute(0, 3, 1, 2).reshape(B, na, no, ny, nx).permute(0, 1, 3, 4, 2)
    yv, xv = torch.meshgrid(torch.arange(ny), torch.arange(nx), indexing="ij")
    grid = torch.stack((xv, yv), 2).view(1, 1, ny, nx, 2).float()
    y = xr.sigmoid()
    xy = (y[..., 0:2] * 2. - 0.5 + grid) * 16.0
    wh = (y[..., 2:4] * 2) ** 2 * torch.tensor(anchors).view(1, na, 1, 1, 2)
    zr = torch.cat((xy, wh, y[..., 4:]), -1).reshape(B, -1, no)
    assert torch.equal(x.cpu().float(), xr)
    assert err(z[:, 50:], zr) < TOL
    assert torch.equal(lg[:, 50:].cpu().float(), xr[..., 5:].reshape(B, -1, no - 5))
    assert float(z[:, :50].abs().max()) == 0
