"""Device letterbox (icaf_letterbox) vs the reference's cv2 pipeline (utils/datasets.py:1404-1427 + :238), bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


def _reference_letterbox(img, new_shape=(640, 640), color=(114, 114, 114), scaleup=True):
    """utils/datasets.py:1404-1427 restated with cv2 (the reference's own dependency), then datasets.py:238."""
    shape = img.shape[:2]
    r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
    if not scaleup:
        r = min(r, 1.0)
    new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = (new_shape[1] - new_unpad[0]) / 2, (new_shape[0] - new_unpad[1]) / 2
    if shape[::-1] != new_unpad:
        img = cv2.resize(img, new_unpad, interpolation=cv2.INTER_LINEAR)
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    img = cv2.copyMakeBorder(img, top, bottom, left, right, cv2.BORDER_CONSTANT, value=color)
    return np.ascontiguousarray(img[:, :, ::-1].transpose(2, 0, 1))


@pytest.mark.parametrize("H0,W0,new_shape,scaleup", [(512, 640, (640, 640), True),      # KAIST frame: no resize, 64-row bands
                                                     (480, 640, (640, 640), True), (1080, 1920, (640, 640), True),
                                                     (300, 400, (640, 640), True), (300, 400, (640, 640), False),
                                                     (333, 517, (512, 640), True), (720, 1280, (384, 672), True)])
def test_letterbox_matches_cv2_bit_exact(cuda_device, H0, W0, new_shape, scaleup):
    from icafusion_b200.datasets import letterbox
    g = np.random.Generator(np.random.PCG64(H0 * 7 + W0))
    frames = g.integers(0, 256, (3, H0, W0, 3), dtype=np.uint8)
    frames[1] = np.clip(np.add.outer(np.arange(H0), np.arange(W0))[..., None] % 256 + np.arange(3) * 40, 0, 255).astype(np.uint8)   # smooth ramp
    out, ratio, pad = letterbox(torch.from_numpy(frames).to(cuda_device), new_shape, scaleup=scaleup)
    torch.cuda.synchronize()
    for b in range(3):
        want = _reference_letterbox(frames[b], new_shape, scaleup=scaleup)
        got = out[b].cpu().numpy()
        assert got.shape == want.shape, (got.shape, want.shape)
        assert np.array_equal(got, want), (b, int(np.abs(got.astype(int) - want.astype(int)).max()))


def test_letterboxed_pair_through_the_detector(cuda_device):
    """detect_twostream.py:70-84 end to end on the device: raw BGR frames -> letterbox -> staging (/255, fp16) -> model."""
    from helpers import load_synth
    from icafusion_b200 import Model
    from icafusion_b200.datasets import letterbox
    model = Model("yolov5s_Transfusion_kaist").eval()
    load_synth(model, 2)
    model = model.fuse().half().to(cuda_device)
    g = np.random.Generator(np.random.PCG64(5))
    rgb0, ir0 = g.integers(0, 256, (1, 512, 640, 3), dtype=np.uint8), g.integers(0, 256, (1, 512, 640, 3), dtype=np.uint8)
    a, _, _ = letterbox(torch.from_numpy(rgb0).to(cuda_device))
    b, _, _ = letterbox(torch.from_numpy(ir0).to(cuda_device))
    assert tuple(a.shape) == (1, 3, 640, 640)
    with torch.no_grad():
        z = model(a, b)[0]
        wa = torch.from_numpy(_reference_letterbox(rgb0[0])[None]).to(cuda_device)
        wb = torch.from_numpy(_reference_letterbox(ir0[0])[None]).to(cuda_device)
        zr = model(wa, wb)[0]
    assert tuple(z.shape) == (1, 25200, 6) and torch.equal(z, zr)


def test_graphed_detect_loop_body(cuda_device):
    """GraphedDetector(frame_hw=..., nms=...): raw BGR frames -> letterbox -> forward -> NMS as ONE graph; equals the separate
    calls (device letterbox, eager model, device NMS) on the same frames."""
    from helpers import load_synth
    from icafusion_b200 import Model, ops
    from icafusion_b200.datasets import letterbox
    from icafusion_b200.engine import GraphedDetector
    model = Model("yolov5s_Transfusion_kaist").eval()
    load_synth(model, 6)
    model = model.fuse().half().to(cuda_device)
    eng = GraphedDetector(model, 2, 384, 640, in_dtype=torch.uint8, device=cuda_device, nms=dict(conf_thres=0.25, iou_thres=0.45),
                          frame_hw=(300, 500))
    g = np.random.Generator(np.random.PCG64(8))
    for _ in range(2):
        a = torch.from_numpy(g.integers(0, 256, (2, 300, 500, 3), dtype=np.uint8)).pin_memory()
        b = torch.from_numpy(g.integers(0, 256, (2, 300, 500, 3), dtype=np.uint8)).pin_memory()
        det, count = eng.infer_frames(a, b)
        with torch.no_grad():
            la, lb = letterbox(a.to(cuda_device), (384, 640))[0], letterbox(b.to(cuda_device), (384, 640))[0]
            z = model(la, lb)[0]
            d2, c2 = ops.nms(z, 0.25, 0.45)
        torch.cuda.synchronize()
        assert count.tolist() == c2.tolist() and torch.equal(det, d2.cpu())
