"""CPU-side checks of the boundary: the shared library loads, exports every symbol include/icaf_b200.h declares,
argument validation works without a GPU, and the product path refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "icaf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(icaf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from icafusion_b200 import _lib
    L = _lib.lib()
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/icaf_b200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)
    assert L.icaf_version() >= 100


def test_argument_validation_needs_no_gpu():
    """Bad arguments are rejected on the host before any launch; the error string is readable through the ABI."""
    from icafusion_b200 import _lib
    L = _lib.lib()
    g = _lib.ConvGeom(1, 8, 8, 12, 8, 8, 16, 1, 1, 1, 0, 64, 32, 0, 0)      # Cin=12: not 4 / multiple of 8
    io = (_lib.ConvIO * 1)()
    rc = L.icaf_conv2d_fwd(ctypes.byref(g), io, 1, None)
    assert rc == 2 and b"Cin" in L.icaf_last_error()
    rc = L.icaf_cross_attention(None, None, None, None, None, None, 1, 10, 16, 128, 8, None)
    assert rc == 1
    with pytest.raises(_lib.IcafError):
        _lib.check(rc, "icaf_cross_attention")


def test_no_cpu_fallback():
    from icafusion_b200 import Conv, Model, TransformerFusionBlock
    x = torch.randn(1, 64, 8, 8)
    with pytest.raises(RuntimeError):
        Conv(64, 64).eval()(x)
    with pytest.raises(RuntimeError):
        TransformerFusionBlock(128, 4, 4).eval()([torch.randn(1, 128, 8, 8)] * 2)
    m = Model("yolov5s_Transfusion_kaist").eval()
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 3, 320, 320), torch.rand(1, 3, 320, 320))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from icafusion_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.IcafError):
        _lib.lib()


def test_shipped_library_has_no_probe_hooks_and_counts_launches():
    """The stage switches of tools/conv_probe.py exist only in -DICAF_PROBE builds; the shipped library exports exactly the
    header's symbols.  The launch tally behind bench.py's `gpu_launches` starts at zero and does not move on rejected calls."""
    from icafusion_b200 import _lib
    L = _lib.lib()
    assert not hasattr(L, "icaf_debug_set")
    n0 = L.icaf_kernel_launches()
    g = _lib.ConvGeom(1, 8, 8, 12, 8, 8, 16, 1, 1, 1, 0, 64, 32, 0, 0)      # rejected on the host (Cin = 12)
    io = (_lib.ConvIO * 1)()
    assert L.icaf_conv2d_fwd(ctypes.byref(g), io, 1, None) != 0
    assert L.icaf_kernel_launches() == n0
