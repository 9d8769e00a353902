"""Drop-in proof for INTEGRATION.md section 2 (CPU only, needs the reference tree): a full-object checkpoint written by
the REAL reference (``torch.save({'model': model})``, train.py:424-435) is unpickled into the shadow modules
(``sys.modules['models.common'] = icafusion_b200.common`` etc.), goes through what ``attempt_load`` does
(models/experimental.py:113-121: ``ckpt['model'].float().fuse().eval()``), loads the reference's state_dict with
``strict=True`` and walks the product path (dry run: meta tensors, every kernel launch planned, none issued)."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

from conftest import ROOT
from oracle.ref_shim import REF_ROOT, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")


def _write_reference_checkpoint(path, sd_path):
    """Runs in a child process so the reference's `models` package never shares sys.modules with the shadow modules."""
    code = textwrap.dedent(f"""
        import sys, torch, os
        sys.path.insert(0, {ROOT!r})
        from oracle import synth
        from oracle.ref_shim import load_reference, REF_ROOT
        common, yolo = load_reference()
        cfg = os.path.join(REF_ROOT, "models", "transformer", "yolov5s_Transfusion_kaist.yaml")
        model = yolo.Model(cfg, ch=3, nc=1)
        shapes = {{k: tuple(v.shape) for k, v in model.state_dict().items()}}
        model.load_state_dict(synth.synth_state_dict(shapes, 77), strict=False)
        model.half()                                   # train.py:427 saves the half() model object
        torch.save({{"epoch": 3, "model": model, "optimizer": None}}, {path!r})
        torch.save(model.float().state_dict(), {sd_path!r})
        print(type(model).__module__, len(shapes))
    """)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.split()[0] == "models.yolo_test"


def test_reference_checkpoint_unpickles_into_shadow_modules(tmp_path):
    ckpt, sdp = str(tmp_path / "last.pt"), str(tmp_path / "sd.pt")
    _write_reference_checkpoint(ckpt, sdp)
    code = textwrap.dedent(f"""
        import sys, torch
        sys.path.insert(0, {ROOT!r})
        import icafusion_b200.common as C, icafusion_b200.yolo_test as Y
        import types
        pkg = types.ModuleType("models"); pkg.__path__ = []
        sys.modules["models"] = pkg                    # INTEGRATION.md section 2: shadow before anything imports the reference
        sys.modules["models.common"] = C
        sys.modules["models.yolo_test"] = Y
        from icafusion_b200 import ops
        ck = torch.load({ckpt!r}, map_location="cpu", weights_only=False)
        m = ck["model"]
        assert type(m) is Y.Model and type(m.model[0]) is C.Conv and type(m.model[-1]) is Y.Detect, type(m)
        assert type(m.model[20]) is C.TransformerFusionBlock and type(m.model[20].crosstransformer[0].crossatt) is C.CrossAttention
        m = m.float().fuse().eval()                    # models/experimental.py:118
        assert not hasattr(m.model[0], "bn") and m.model[0].conv.bias is not None
        # the reference's own state_dict (unfused layout) loads strictly into a freshly built shadow model
        fresh = Y.Model("yolov5s_Transfusion_kaist")
        sd = torch.load({sdp!r}, map_location="cpu")
        missing = fresh.load_state_dict(sd, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        fresh = fresh.eval().fuse()
        for (ka, va), (kb, vb) in zip(sorted(m.state_dict().items()), sorted(fresh.state_dict().items())):
            assert ka == kb and va.shape == vb.shape and torch.allclose(va.float(), vb.float(), atol=2e-3, rtol=2e-3), ka
        # the unpickled object drives the product path: dry-run walk (nothing launched), every conv plan accepted
        from icafusion_b200 import _lib
        import ctypes
        img = torch.empty(1, 3, 512, 640, dtype=torch.uint8, device="meta")
        with torch.no_grad(), ops.dry_run() as dr:
            z, logits, xs = m.half()(img, img)
        assert tuple(z.shape) == (1, 20160, 6) and len(xs) == 3
        convs = [w for n, a, w in dr.records if n == "icaf_conv2d_fwd"]
        assert len(convs) >= 60
        for w in convs:
            pl = _lib.ConvPlan()
            assert _lib.lib().icaf_conv2d_plan(ctypes.byref(w["geom"]), w["n_io"], 148, -1, ctypes.byref(pl)) == 0
        print("ok", len(dr.records))
    """)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout[-1000:], out.stderr[-3000:])
    assert out.stdout.strip().startswith("ok")


def test_shadow_modules_export_the_reference_names():
    """Every class the Transfusion YAMLs / pickles name exists in the shadow modules with the reference's constructor
    signature (parameter names and defaults)."""
    import inspect
    import icafusion_b200.common as C
    import icafusion_b200.yolo_test as Y
    from oracle.ref_shim import load_reference
    src_common = open(os.path.join(REF_ROOT, "models", "common.py")).read()
    for name in ("Conv", "Bottleneck", "C3", "SPPF", "Concat", "TransformerFusionBlock", "CrossTransformerBlock", "CrossAttention",
                 "LearnableCoefficient", "LearnableWeights", "AdaptivePool2d"):
        assert f"class {name}(" in src_common and hasattr(C, name), name
    rc, ry = load_reference()
    try:
        for name in ("Conv", "Bottleneck", "C3", "SPPF", "Concat", "TransformerFusionBlock", "CrossTransformerBlock", "CrossAttention",
                     "AdaptivePool2d"):
            a, b = inspect.signature(getattr(rc, name).__init__), inspect.signature(getattr(C, name).__init__)
            assert [(p.name, p.default) for p in a.parameters.values()] == [(p.name, p.default) for p in b.parameters.values()], name
        for name in ("Model", "Detect"):
            assert hasattr(Y, name)
        a, b = inspect.signature(ry.Detect.__init__), inspect.signature(Y.Detect.__init__)
        assert [p.name for p in a.parameters.values()] == [p.name for p in b.parameters.values()]
    finally:
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils.")]:
            del sys.modules[k]      # keep the reference's packages out of the other tests' namespace
