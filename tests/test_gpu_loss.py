"""Device ComputeLoss forward (icaf_compute_loss_fwd) vs golden outputs of the real reference's utils/loss.py and vs the
CPU oracle, fp32 and fp16 predictions, determinism."""
import types

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import icaf_oracle as O
from oracle.gen_golden_loss import synth_case

pytestmark = pytest.mark.gpu


def _model_stub(nc, anchors, hyp, gr, device):
    """What ComputeLoss reads from the model: .hyp, .gr and the Detect attributes (loss.py:330-352)."""
    det = types.SimpleNamespace(na=anchors.shape[1], nc=nc, nl=anchors.shape[0], anchors=torch.from_numpy(anchors).to(device))
    return types.SimpleNamespace(hyp=hyp, gr=gr, model=[det])


def test_compute_loss_matches_reference_golden(cuda_device):
    from icafusion_b200.loss import ComputeLoss
    m, d = load_golden("loss_cases")
    for cs in m["cases"]:
        p, t = synth_case(cs["name"], cs["nc"], cs["B"], cs["nt"])
        anchors = d[f"{cs['name']}_anchors"]
        fn = ComputeLoss(_model_stub(cs["nc"], anchors, cs["hyp"], cs["gr"], cuda_device))
        loss, items = fn([torch.from_numpy(x).to(cuda_device) for x in p], torch.from_numpy(t).to(cuda_device))
        got = torch.cat([loss, items]).cpu().numpy()
        want = d[f"{cs['name']}_out"]
        print(f"\n[loss {cs['name']}] device {got}  reference {want}")
        assert np.allclose(got, want, rtol=3e-5, atol=2e-6), cs["name"]
        loss2, items2 = fn([torch.from_numpy(x).to(cuda_device) for x in p], torch.from_numpy(t).to(cuda_device))
        assert torch.equal(torch.cat([loss, items]), torch.cat([loss2, items2])), "not deterministic"
        # fp16 predictions: same values after rounding, arithmetic still fp32 -> equals the oracle on the rounded inputs
        p16 = [torch.from_numpy(x).half() for x in p]
        l16, i16 = fn([x.to(cuda_device) for x in p16], torch.from_numpy(t).to(cuda_device))
        lo, io = O.compute_loss([x.float() for x in p16], torch.from_numpy(t), torch.from_numpy(anchors), cs["hyp"], cs["gr"])
        assert np.allclose(torch.cat([l16, i16]).cpu().numpy(), np.concatenate([lo.numpy().reshape(1), io.numpy()]), rtol=3e-5, atol=2e-6)


def test_compute_loss_on_detector_training_outputs(cuda_device):
    """test.py:132-133: validation loss from the model's own Detect training outputs (the third element of the eval-mode
    return value), here against the oracle on the same fp16 outputs."""
    from helpers import load_synth
    from icafusion_b200 import Model
    from icafusion_b200.loss import ComputeLoss
    from oracle import synth
    model = Model("yolov5s_Transfusion_kaist").eval()
    load_synth(model, 13)
    model = model.fuse().to(cuda_device)
    model.hyp = dict(box=0.05, obj=1.0, cls=0.5, cls_pw=1.0, obj_pw=1.0, anchor_t=4.0, fl_gamma=0.0)
    model.gr = 1.0
    rgb, ir = synth.synth_images(2, 512, 640, 13)
    with torch.no_grad():
        _, _, train_out = model(rgb.to(cuda_device), ir.to(cuda_device))
    _, t = synth_case("kaist_nc1", 1, 2, 20)
    loss, items = ComputeLoss(model)([x.float() for x in train_out], torch.from_numpy(t).to(cuda_device))
    lo, io = O.compute_loss([x.float().cpu() for x in train_out], torch.from_numpy(t), model.model[-1].anchors.cpu(), model.hyp, 1.0)
    assert np.allclose(torch.cat([loss, items]).cpu().numpy(), np.concatenate([lo.numpy().reshape(1), io.numpy()]), rtol=5e-5, atol=2e-6)


@pytest.mark.parametrize("layout", ["reference", "nhwc", "nhwc_padded"])
def test_compute_loss_backward(cuda_device, layout):
    """d loss / d predictions (icaf_compute_loss_bwd) against autograd through the oracle restatement of utils/loss.py, for
    the reference's contiguous (B,na,ny,nx,no) tensors and for permuted views of the head's own NHWC maps."""
    from icafusion_b200.loss import ComputeLoss
    m, d = load_golden("loss_cases")
    for cs in m["cases"]:
        p, t = synth_case(cs["name"], cs["nc"], cs["B"], cs["nt"])
        anchors = d[f"{cs['name']}_anchors"]
        fn = ComputeLoss(_model_stub(cs["nc"], anchors, cs["hyp"], cs["gr"], cuda_device))
        ref = [torch.from_numpy(x).requires_grad_(True) for x in p]
        lo, _ = O.compute_loss(ref, torch.from_numpy(t), torch.from_numpy(anchors), cs["hyp"], cs["gr"])
        (lo.sum() * 3.0).backward()
        dev = []
        for x in p:
            x = torch.from_numpy(x).to(cuda_device)
            B, na, ny, nx, no = x.shape
            if layout != "reference":
                ld = na * no if layout == "nhwc" else (na * no + 7) // 8 * 8 + 8
                buf = torch.zeros(B, ny, nx, ld, device=cuda_device)
                buf[..., :na * no] = x.permute(0, 2, 3, 1, 4).reshape(B, ny, nx, na * no)
                x = buf.as_strided((B, na, ny, nx, no), (ny * nx * ld, no, nx * ld, ld, 1))
                assert not x.is_contiguous()
            dev.append(x.requires_grad_(True))
        loss, items = fn(dev, torch.from_numpy(t).to(cuda_device))
        assert np.allclose(torch.cat([loss.detach(), items]).cpu().numpy(), d[f"{cs['name']}_out"], rtol=3e-5, atol=2e-6)
        (loss.sum() * 3.0).backward()
        for lvl, (g, r) in enumerate(zip(dev, ref)):
            e = float((g.grad.cpu() - r.grad).abs().max() / r.grad.abs().max())
            print(f"\n[loss bwd {cs['name']} {layout} level {lvl}] rel err {e:.2e}")
            assert e < 2e-5, (cs["name"], lvl)
        g2 = d[f"{cs['name']}_grad2"] * 3.0                  # the REAL reference's loss.backward(), coarsest level
        assert np.abs(dev[2].grad.cpu().numpy() - g2).max() <= 2e-5 * np.abs(g2).max(), cs["name"]
