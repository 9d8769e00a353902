"""Whole two-stream detector on the GPU vs golden vectors from the real reference and operator-level checks
(Conv / C3 / SPPF modules vs the CPU oracle)."""
import pytest
import torch

from conftest import load_golden
from helpers import err, load_synth
from oracle import icaf_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu

# Whole-model tolerance: ~100 fp16 layers deep.  The reference's own fp16 path sits 1.1e-3 .. 1.2e-3 from its fp32
# path on these inputs (meta['ref_fp16_self_dev']); z is additionally fp16-rounded at magnitudes up to ~1e3.
TOL_MODEL = 3e-3


@pytest.mark.parametrize("name", ["yolov5s_320", "yolov5s_512x640", "yolov5l_512x640"])
@pytest.mark.parametrize("fused", [False, True])
def test_model_matches_reference_golden(cuda_device, name, fused):
    from icafusion_b200 import Model
    m, d = load_golden(name)
    model = Model(f"yolov5{m['size']}_Transfusion_kaist").eval()
    load_synth(model, m["seed"])
    if fused:
        model.fuse()
    model = model.to(cuda_device)
    rgb, ir = synth.synth_images(m["B"], m["H"], m["W"], m["seed"])
    with torch.no_grad():
        z, logits, xs = model(rgb.to(cuda_device), ir.to(cuda_device))
    torch.cuda.synchronize()
    ez = err(z, d["z_fused" if fused else "z"])
    el = err(logits, d["logits"])
    ex = max(err(xs[j], d[f"x{j}"].astype("float32")) for j in range(3))
    print(f"\n[{name} fused={fused}] z {ez:.2e} logits {el:.2e} x {ex:.2e}  (reference fp16 self-dev: {m.get('ref_fp16_self_dev')})")
    assert tuple(z.shape) == d["z"].shape and len(xs) == 3
    assert ez < TOL_MODEL and el < TOL_MODEL and ex < TOL_MODEL


def test_operator_modules_vs_oracle(cuda_device):
    """Stand-alone Conv / Bottleneck / C3 / SPPF modules called like the reference's forward_once calls them."""
    from icafusion_b200 import C3, SPPF, Conv
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 64, 16, 20, generator=g).half()
    for make, fn in ((lambda: Conv(64, 128, 3, 2), lambda sd, t: O.conv_bn_silu(t, sd, "m", 3, 2, bn_eps=1e-5)),
                     (lambda: C3(64, 64, 2), lambda sd, t: O.c3(t, sd, "m", 2, True, bn_eps=1e-5)),
                     (lambda: C3(64, 128, 1, False), lambda sd, t: O.c3(t, sd, "m", 1, False, bn_eps=1e-5)),
                     (lambda: SPPF(64, 64, 5), lambda sd, t: O.sppf(t, sd, "m", 5, bn_eps=1e-5))):
        mod = make().eval()
        sd = load_synth(mod, 11, "m.")
        mod = mod.to(cuda_device)
        with torch.no_grad():
            y = mod(x.to(cuda_device))
            ref = fn(sd, x.float())
        assert y.shape == ref.shape and err(y, ref) < 2e-3, type(mod).__name__


def test_batch_independence(cuda_device):
    """Size-independent property at full 512x640 size: each pair's result does not depend on its batch neighbours
    (what makes batch-dim sharding across GPUs valid).  Equality is up to fp32 summation order only: the split-K factor
    of the small deep layers is chosen from the grid size, i.e. from the batch."""
    from icafusion_b200 import Model
    model = Model("yolov5s_Transfusion_kaist").eval()
    load_synth(model, 3)
    model = model.fuse().to(cuda_device)
    rgb, ir = synth.synth_images(3, 512, 640, 3)
    rgb, ir = rgb.to(cuda_device), ir.to(cuda_device)
    with torch.no_grad():
        z_all = model(rgb, ir)[0]
        z_1 = model(rgb[1:2], ir[1:2])[0]
    assert err(z_all[1:2], z_1) < 1e-3
    model.__dict__["_icaf_concurrent"] = False          # same walk without the side-stream forks
    with torch.no_grad():
        z_seq = model(rgb[1:2], ir[1:2])[0]
    assert torch.equal(z_seq, z_1)


def test_graph_engines_match_eager(cuda_device):
    """CUDA-graph replay (GraphedDetector) and the depth-2 streaming front end (PipelinedDetector) return exactly what the
    eager forward returns, frame after frame, from host uint8 frames."""
    from icafusion_b200 import Model
    from icafusion_b200.engine import GraphedDetector, PipelinedDetector
    model = Model("yolov5s_Transfusion_kaist").eval()
    load_synth(model, 5)
    model = model.fuse().half().to(cuda_device)
    frames = []
    for sd in range(5):
        a, b = synth.synth_images(1, 320, 320, 100 + sd)
        frames.append(((a * 255).to(torch.uint8).pin_memory(), (b * 255).to(torch.uint8).pin_memory()))
    with torch.no_grad():
        want = [model(a.to(cuda_device), b.to(cuda_device))[0].cpu() for a, b in frames]
    eng = GraphedDetector(model, 1, 320, 320, in_dtype=torch.uint8, device=cuda_device)
    for (a, b), w in zip(frames, want):
        assert torch.equal(eng.infer_to_host(a, b).clone(), w)
    pipe = PipelinedDetector(model, 1, 320, 320, device=cuda_device, depth=2)
    got = [z.clone() for z in pipe.infer_stream(frames)]
    assert len(got) == len(want) and all(torch.equal(g, w) for g, w in zip(got, want))
    assert eng.launches_per_step > 50


def test_letterboxed_640x640_vs_oracle(cuda_device):
    """The reference's detect_twostream.py feeds letterboxed 640x640 frames (utils/datasets.py:1404-1444): P3 80x80 pools
    with k=s=(4,4), P4 40x40 -> 16x16 with k=(10,10) s=(2,2), P5 20x20 with k=s=(2,2).  Checked against the CPU oracle."""
    from icafusion_b200 import Model
    from icafusion_b200.cfg import load_cfg
    cfg = load_cfg("yolov5s_Transfusion_kaist")
    model = Model(cfg).eval()
    sd = load_synth(model, 21)
    model = model.fuse().to(cuda_device)
    rgb, ir = synth.synth_images(1, 640, 640, 21)
    with torch.no_grad():
        z = model(rgb.to(cuda_device), ir.to(cuda_device))[0]
        zr = O.model_forward(O.fold_bn(sd), cfg, rgb, ir)[0]
    e = err(z, zr)
    print(f"\n[yolov5s 640x640] z {e:.2e}")
    assert tuple(z.shape) == (1, 25200, 6) and e < TOL_MODEL


def test_bench_shape_yolov5l_b16_through_graph(cuda_device):
    """The benchmark's own configuration (BASELINE configs[2]: yolov5l, batch 16, 512x640, uint8 frames through
    GraphedDetector): pairs 0 and 15 against the fp32 CPU oracle, and every pair against the same model run eagerly at
    batch 1 (batch independence; equal up to the fp32 summation order the batch-dependent tile plans choose)."""
    from icafusion_b200 import Model
    from icafusion_b200.cfg import load_cfg
    from icafusion_b200.engine import GraphedDetector
    cfg = load_cfg("yolov5l_Transfusion_kaist")
    model = Model(cfg).eval()
    sd = load_synth(model, 0)
    model = model.fuse().half().to(cuda_device)
    B = 16
    rgb, ir = synth.synth_images(B, 512, 640, 0)
    rgb_u8, ir_u8 = (rgb * 255).to(torch.uint8), (ir * 255).to(torch.uint8)
    eng = GraphedDetector(model, B, 512, 640, in_dtype=torch.uint8, device=cuda_device)
    z = eng.infer_to_host(rgb_u8.pin_memory(), ir_u8.pin_memory()).clone()
    assert tuple(z.shape) == (B, 20160, 6) and torch.isfinite(z.float()).all()
    with torch.no_grad():
        for j in (0, 15):
            a, b = rgb_u8[j:j + 1].float() / 255.0, ir_u8[j:j + 1].float() / 255.0
            zr = O.model_forward(O.fold_bn(sd), cfg, a, b)[0]
            e = err(z[j:j + 1], zr)
            print(f"\n[yolov5l b16 graph, pair {j}] z vs fp32 oracle {e:.2e}")
            assert e < TOL_MODEL
        worst = 0.0
        for j in range(B):
            z1 = model(rgb_u8[j:j + 1].to(cuda_device), ir_u8[j:j + 1].to(cuda_device))[0]
            worst = max(worst, err(z[j:j + 1], z1))
        print(f"[yolov5l b16 graph] worst pair vs the eager batch-1 forward {worst:.2e}")
        # two fp16 runs of ~100 layers with different tile plans / split-K factors (the plans depend on the batch): each sits
        # ~1e-3 from the fp32 oracle (above), so their mutual distance can reach the sum of the two
        assert worst < 2.5e-3
