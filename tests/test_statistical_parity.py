"""Statistical image parity against the reference running on its own generator (glibc drand48): the GPU's samples are
independent of the reference's, so the comparison is in distribution:
  (i)   global mean radiance per channel within 1.5 % (+ 3 standard errors);
  (ii)  >= 98.5 % of pixel channels with |mean_gpu - mean_ref| <= 4 * sqrt(var_ref/K + var_gpu/K), the variances estimated
        from K = 8 independent batches on each side (a perfect match gives ~99.99 %; heavy-tailed pixels near the small
        light make the batch variance itself noisy, hence the slack);
  (iii) PSNR of the gamma-encoded 8-bit images (PSC/main.cpp:315-325) no worse than 1.5 dB below the noise floor, i.e.
        the PSNR between two independent reference renders of the same sample count (second half of the fixture).
The reference batches are committed fixtures (tests/golden/stat_*.npz, tests/golden/make_stat_golden.py)."""
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden"


def _encode(mean):
    return np.minimum((255.99 * np.sqrt(np.maximum(mean, 0))).astype(np.int64), 255).astype(np.float64)


def _psnr(a, b):
    mse = np.mean((_encode(a) - _encode(b)) ** 2)
    return 10 * np.log10(255.0 ** 2 / max(mse, 1e-12))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cornell_box", "final_northstar", "ch01_random"])
def test_image_statistics_match_reference_drand48(rtnw, ctx, name):
    g = np.load(GOLD / f"stat_{name}.npz")
    nx, ny, spp, K = int(g["nx"]), int(g["ny"]), int(g["spp"]), int(g["k"])
    ref_a, ref_b = g["batches"][:K].astype(np.float64), g["batches"][K:].astype(np.float64)
    hs = rtnw.HostScene(name)
    ds = ctx.upload(hs.desc_ptr)
    cam = hs.camera(nx, ny)
    gpu = np.stack([ds.render(cam, hs.params(nx=nx, ny=ny, ns=spp, seed=900 + k))[0] / spp for k in range(K)]).astype(np.float64)
    ds.close()
    m_gpu, m_ref = gpu.mean(0), ref_a.mean(0)
    se = np.sqrt(gpu.var(0, ddof=1) / K + ref_a.var(0, ddof=1) / K)
    # (i) global means
    for c in range(3):
        gm, rm = m_gpu[..., c].mean(), m_ref[..., c].mean()
        se_glob = np.sqrt((gpu[..., c].mean(axis=(1, 2)).var(ddof=1) + ref_a[..., c].mean(axis=(1, 2)).var(ddof=1)) / K)
        assert abs(gm - rm) <= 0.015 * rm + 3 * se_glob, (name, c, gm, rm, se_glob)
    # (ii) per-pixel z-scores
    z = np.abs(m_gpu - m_ref) / np.maximum(se, 1e-9)
    frac = (z[se > 0] <= 4).mean()
    assert frac >= 0.985, f"{name}: only {frac:.4f} of pixel channels within 4 sigma"
    # (iii) PSNR against the noise floor of the reference itself
    floor = _psnr(ref_b.mean(0), m_ref)
    got = _psnr(m_gpu, m_ref)
    assert got >= floor - 1.5, f"{name}: PSNR {got:.2f} dB vs reference-vs-reference {floor:.2f} dB"


def test_fixtures_are_self_consistent():
    """the two halves of each fixture are two independent reference renders: their difference defines the noise floor and
    must itself pass the z-score criterion (guards the test's statistics, CPU only)"""
    for name in ["cornell_box", "final_northstar", "ch01_random"]:
        g = np.load(GOLD / f"stat_{name}.npz")
        K = int(g["k"])
        a, b = g["batches"][:K].astype(np.float64), g["batches"][K:].astype(np.float64)
        se = np.sqrt(a.var(0, ddof=1) / K + b.var(0, ddof=1) / K)
        z = np.abs(a.mean(0) - b.mean(0)) / np.maximum(se, 1e-9)
        assert (z[se > 0] <= 4).mean() >= 0.985
        assert _psnr(a.mean(0), b.mean(0)) > 10
