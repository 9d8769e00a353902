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
