"""The C++ host side above the C-ABI: rtnw_main (the reference's main() with the sample loop replaced by the library) and
the CUDA bridge that serves the reference API's virtuals (hit / scatter / emitted / value) from the GPU."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "peter-shirley-ray-tracing-the-next-week_b200" / "bin" / "rtnw_main"


def test_driver_is_built_and_fails_loudly_without_gpu():
    import torch
    assert BIN.exists(), "run __graft_entry__.build()"
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([str(BIN), "--selftest-bridge"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_driver_renders_the_same_image_as_the_python_binding(rtnw, ctx, tmp_path):
    out = tmp_path / "cb.ppm"
    r = subprocess.run([str(BIN), "cornell_box", "40", "40", "8", str(out), "--seed", "5"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    tok = out.read_text().split()
    assert tok[:4] == ["P3", "40", "40", "255"]
    img = np.array(tok[4:], dtype=np.int32).reshape(40, 40, 3)
    hs = rtnw.HostScene("cornell_box")
    ds = ctx.upload(hs.desc_ptr)
    sums, _ = ds.render(hs.camera(40, 40), hs.params(nx=40, ny=40, ns=8, seed=5))
    assert np.array_equal(img, rtnw.quantize(sums, 8, clamp255=True))
    ds.close()


@pytest.mark.gpu
def test_reference_api_virtuals_are_served_by_the_gpu(rtnw, ctx):
    r = subprocess.run([str(BIN), "--selftest-bridge"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    hits = [l for l in lines if l.startswith("hit depth")]
    assert len(hits) >= 1
    # the first world->hit() of the self test must equal rtnw_trace on the same ray
    m = re.match(r"hit depth 0 t (\S+) p (\S+) (\S+) (\S+)", hits[0])
    hs = rtnw.HostScene("cornell_box")
    ds = ctx.upload(hs.desc_ptr)
    ray = np.zeros(1, dtype=rtnw.RAY_DTYPE)
    ray["origin"] = [278, 278, -800]; ray["direction"] = [0.1, -0.2, 1.0]; ray["time"] = 0.5
    h = ds.trace(ray)[0]
    assert np.float32(float(m.group(1))) == h["t"]
    assert np.allclose([float(m.group(k)) for k in (2, 3, 4)], h["p"], rtol=1e-6)
    ds.close()
    chk = [l for l in lines if l.startswith("checker")][0]
    assert chk == "checker 0.2 0.3 0.1 | 0.2 0.3 0.1" or "0.9" in chk
    assert any(l.startswith("noise ") for l in lines) and any(l.startswith("radiance ") for l in lines)
