"""Device NMS (icaf_nms) vs golden rows from the real reference's non_max_suppression and vs the CPU oracle; the
stand-alone DMFF operator forwards (LearnableCoefficient / LearnableWeights / AdaptivePool2d / CrossAttention)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import err, load_synth
from oracle import icaf_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu


def test_nms_matches_reference_golden_bit_exact(cuda_device):
    from icafusion_b200 import ops
    from icafusion_b200.general import non_max_suppression
    m, d = load_golden("nms_cases")
    pred = torch.from_numpy(d["pred"]).to(cuda_device)
    for st in m["settings"]:
        det, count = ops.nms(pred, st["conf"], st["iou"], st["agnostic"], st["classes"])
        torch.cuda.synchronize()
        assert count.tolist() == st["counts"], (st["name"], count.tolist())
        for b, n in enumerate(st["counts"]):
            want = d[f"{st['name']}_{b}"]
            got = det[b, :n].cpu().numpy()
            assert np.array_equal(got, want), (st["name"], b, np.abs(got - want).max())
        lst = non_max_suppression(pred, st["conf"], st["iou"], classes=st["classes"], agnostic=st["agnostic"])
        assert [int(t.shape[0]) for t in lst] == st["counts"]


def test_nms_edge_cases(cuda_device):
    from icafusion_b200 import ops
    # nothing above the threshold; a single box; ties in confidence keep row order; max_det cut
    z = torch.zeros(3, 64, 6, dtype=torch.float16, device=cuda_device)
    z[1, 5] = torch.tensor([100, 100, 20, 20, 0.9, 1.0], dtype=torch.float16)
    z[2, :, 0] = torch.arange(64, device=cuda_device).half() * 50      # 64 disjoint boxes, identical confidence
    z[2, :, 1] = 30
    z[2, :, 2:4] = 10
    z[2, :, 4] = 0.5
    z[2, :, 5] = 1.0
    det, count = ops.nms(z, 0.25, 0.45, max_det=10)
    torch.cuda.synchronize()
    assert count.tolist() == [0, 1, 10]
    assert det[1, 0].tolist() == [90.0, 90.0, 110.0, 110.0, pytest.approx(0.9, abs=1e-3), 0.0]
    assert det[2, :10, 0].tolist() == [50.0 * i - 5.0 for i in range(10)]
    ref = O.non_max_suppression(z.cpu(), 0.25, 0.45, max_det=10)
    for b in range(3):
        assert np.array_equal(det[b, :int(count[b])].cpu().numpy(), ref[b].numpy())


def test_nms_on_detector_output_vs_oracle(cuda_device):
    """End of the real pipeline: the detector's own decoded predictions through the device NMS == the oracle's NMS of the
    same fp16 predictions (test.py's dense setting: ~20 k candidates per image)."""
    from icafusion_b200 import Model, ops
    model = Model("yolov5s_Transfusion_kaist").eval()
    load_synth(model, 9)
    model = model.fuse().to(cuda_device)
    rgb, ir = synth.synth_images(2, 512, 640, 9)
    with torch.no_grad():
        z = model(rgb.to(cuda_device), ir.to(cuda_device))[0]
    for conf, iou in ((0.25, 0.45), (0.001, 0.6)):
        det, count = ops.nms(z, conf, iou)
        torch.cuda.synchronize()
        ref = O.non_max_suppression(z.cpu(), conf, iou)
        assert count.tolist() == [int(r.shape[0]) for r in ref]
        for b, r in enumerate(ref):
            assert np.array_equal(det[b, :r.shape[0]].cpu().numpy(), r.numpy()), (conf, b)


def test_standalone_dmff_operators(cuda_device):
    """The small DMFF operator classes called on their own, like a reference-side ablation would (common.py:569-587,
    868-891, 641-687), against the CPU oracle."""
    from icafusion_b200.common import AdaptivePool2d, CrossAttention, LearnableCoefficient, LearnableWeights
    g = torch.Generator().manual_seed(3)
    x1, x2 = torch.randn(2, 64, 16, 20, generator=g), torch.randn(2, 64, 16, 20, generator=g)
    lc, lw = LearnableCoefficient().to(cuda_device), LearnableWeights().to(cuda_device)
    with torch.no_grad():
        lc.bias.fill_(1.37)
        lw.w1.fill_(0.3)
        lw.w2.fill_(0.9)
        assert err(lc(x1.to(cuda_device)), x1.half().float() * 1.37) < 1e-3
        assert err(lw(x1.to(cuda_device), x2.to(cuda_device)), x1.half().float() * 0.3 + x2.half().float() * 0.9) < 1e-3
        for kind in ("avg", "max"):
            y = AdaptivePool2d(10, 10, kind)(x1.half().to(cuda_device))
            ref = O.adaptive_pool(x1.half().float(), 10, 10, kind)
            assert tuple(y.shape) == tuple(ref.shape) and err(y, ref) < 1e-3, kind
        assert AdaptivePool2d(16, 20, "avg")(x1.half().to(cuda_device)).shape == x1.shape      # identity when not larger
        ca = CrossAttention(128, 128, 128, 8).eval()
        sd = load_synth(ca, 4, "ca.")
        r, i = torch.randn(2, 100, 128, generator=g), torch.randn(2, 100, 128, generator=g)
        ov, oi = ca.to(cuda_device)([r.to(cuda_device), i.to(cuda_device)])
        rv, ri = O.cross_attention(r.half().float(), i.half().float(), sd, "ca")
        assert tuple(ov.shape) == (2, 100, 128)
        assert err(ov, rv) < 1e-3 and err(oi, ri) < 1e-3


def test_graphed_detector_with_captured_nms(cuda_device):
    """GraphedDetector(nms=...) replays forward + NMS as one graph; its detections equal the oracle's NMS of its own z."""
    from icafusion_b200 import Model
    from icafusion_b200.engine import GraphedDetector
    model = Model("yolov5s_Transfusion_kaist").eval()
    load_synth(model, 5)
    model = model.fuse().half().to(cuda_device)
    eng = GraphedDetector(model, 2, 320, 320, in_dtype=torch.uint8, device=cuda_device, nms=dict(conf_thres=0.25, iou_thres=0.45))
    for sd in (31, 32):
        a, b = synth.synth_images(2, 320, 320, sd)
        a, b = (a * 255).to(torch.uint8).pin_memory(), (b * 255).to(torch.uint8).pin_memory()
        det, count = eng.infer_detections(a, b)
        ref = O.non_max_suppression(eng.z.cpu(), 0.25, 0.45)
        assert count.tolist() == [int(r.shape[0]) for r in ref]
        for i, r in enumerate(ref):
            assert np.array_equal(det[i, :r.shape[0]].numpy(), r.numpy())


def test_nms_idempotent_at_full_size(cuda_device):
    """Size-independent property at the full 20160-row prediction size: suppressing the survivors again (as predictions with
    objectness = confidence, class score 1) keeps every one of them, in the same order."""
    from icafusion_b200 import ops
    g = torch.Generator().manual_seed(2)
    B, R = 4, 20160
    z = torch.zeros(B, R, 6)
    z[..., 0:2] = torch.rand(B, R, 2, generator=g) * 600
    z[..., 2:4] = torch.rand(B, R, 2, generator=g) * 80 + 4
    z[..., 4] = torch.rand(B, R, generator=g) ** 4
    z[..., 5] = 1.0
    z = z.half().to(cuda_device)
    det, cnt = ops.nms(z, 0.05, 0.5)
    n = cnt.tolist()
    z2 = torch.zeros(B, 304, 6, device=cuda_device)
    for b in range(B):
        d = det[b, :n[b]]
        z2[b, :n[b], 0] = (d[:, 0] + d[:, 2]) / 2
        z2[b, :n[b], 1] = (d[:, 1] + d[:, 3]) / 2
        z2[b, :n[b], 2] = d[:, 2] - d[:, 0]
        z2[b, :n[b], 3] = d[:, 3] - d[:, 1]
        z2[b, :n[b], 4] = d[:, 4]
        z2[b, :n[b], 5] = 1.0
    det2, cnt2 = ops.nms(z2.half(), 0.04, 0.5 + 2e-2)       # survivors overlap at most 0.5 (+ fp16 re-rounding of the boxes)
    torch.cuda.synchronize()
    assert min(n) > 50 and cnt2.tolist() == n
