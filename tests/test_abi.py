"""CPU-only checks of the drop-in boundary: both C-ABI libraries load and export every symbol their headers declare,
and compute entry points fail loudly (never silently fall back) when no GPU is present."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared(header):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rtnw_[a-z0-9_]+)\s*\(", text)))


def test_device_library_exports_every_declared_symbol(rtnw):
    lib = rtnw.device_lib()
    names = _declared("rtnw.h")
    assert set(names) == set(rtnw.DEVICE_SYMBOLS)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.rtnw_abi_version() == rtnw.RTNW_ABI_VERSION


def test_host_library_exports_every_declared_symbol(rtnw):
    lib = rtnw.host_lib()
    names = _declared("rtnw_host.h")
    assert set(names) == set(rtnw.HOST_SYMBOLS)
    for n in names:
        assert hasattr(lib, n), n


def test_struct_sizes_match_the_header(rtnw):
    assert C.sizeof(rtnw.Prim) == 32 and C.sizeof(rtnw.XformOp) == 16 and C.sizeof(rtnw.BvhNode) == 64
    assert C.sizeof(rtnw.Item) == 48 and C.sizeof(rtnw.Material) == 32 and C.sizeof(rtnw.Texture) == 32
    assert rtnw.RAY_DTYPE.itemsize == 32 and rtnw.HIT_DTYPE.itemsize == 48


def test_no_cpu_fallback_without_a_gpu(rtnw):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert rtnw.device_count() == 0
    with pytest.raises(rtnw.RtnwError) as e:
        rtnw.Context(0)
    assert e.value.code == rtnw.RTNW_ERR_CUDA


def test_unknown_scene_and_unsupported_nesting_are_reported(rtnw):
    with pytest.raises(rtnw.RtnwError):
        rtnw.HostScene("no_such_scene")


def test_quantize_matches_reference_epilogue(rtnw):
    import numpy as np
    # PSC/main.cpp:315-325: col/ns, sqrt, int(255.99*c), clamp to 255; output top row first
    sums = np.array([[[0.0, 50.0, 100.0], [400.0, 25.0, 1.0]], [[100.0, 100.0, 100.0], [9.0, 16.0, 36.0]]], dtype=np.float32)
    q = rtnw.quantize(sums, 100, clamp255=True)
    want_bottom = [[0, int(255.99 * np.sqrt(np.float32(0.5))), 255], [255, int(255.99 * 0.5), int(255.99 * np.float32(0.1))]]
    assert q[1].tolist() == want_bottom
    assert q[0][0].tolist() == [255, 255, 255]
    q2 = rtnw.quantize(sums, 100, clamp255=False)
    assert q2[1][1][0] == int(255.99 * 2.0)
