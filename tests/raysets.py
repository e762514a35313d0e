"""Deterministic ray sets shared by the parity tests and the golden-fixture generator (needs the reference oracle)."""
import numpy as np

import ref_oracle as ro

FLT_MAX = float(np.finfo(np.float32).max)


def make_rays(name, n_primary=3000, seed=7):
    """primary camera rays + secondary rays leaving real hit points + stress rays (axis-parallel, zero components,
    origins inside the scene volume).  Returns (RefScene, rays)."""
    rng = np.random.default_rng(seed)
    v = ro.view_of(name)
    nx, ny = 120, 90
    ij = np.stack([rng.integers(0, nx, n_primary), rng.integers(0, ny, n_primary)], axis=1)
    prim = ro.camera_rays(v, nx, ny, ij, rng.integers(0, 50, n_primary), seed=seed)
    rs = ro.RefScene(name, tagged=True)
    h = rs.trace(prim, 0.001, FLT_MAX, seed=seed)
    hit = h["prim_id"] >= 0
    sec = np.zeros(int(hit.sum()) * 2, dtype=ro.RAY_DTYPE)
    p = np.repeat(h["p"][hit], 2, axis=0)
    d = rng.normal(size=p.shape).astype(np.float32)
    d[::2] = d[::2] + h["normal"][hit]  # half of them lambertian-like
    sec["origin"] = p
    sec["direction"] = d
    sec["time"] = rng.random(len(sec)).astype(np.float32)
    sec["key"] = rng.integers(0, 2**31, len(sec), dtype=np.uint32)
    lo, hi = h["p"][hit].min(axis=0) - 1, h["p"][hit].max(axis=0) + 1
    m = max(n_primary // 2, 16)
    st = np.zeros(m, dtype=ro.RAY_DTYPE)
    st["origin"] = (lo + rng.random((m, 3)) * (hi - lo)).astype(np.float32)
    dd = rng.normal(size=(m, 3)).astype(np.float32)
    dd[np.arange(m), rng.integers(0, 3, m)] *= (rng.random(m) < 0.5)  # zero one component in half of them
    axis = rng.random(m) < 0.25
    dd[axis] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, int(axis.sum()))] * rng.choice([-1.0, 1.0], int(axis.sum()))[:, None]
    st["direction"] = dd
    st["time"] = rng.random(m).astype(np.float32)
    st["key"] = rng.integers(0, 2**31, m, dtype=np.uint32)
    return rs, np.concatenate([prim, sec, st])


def assert_hits_equal(got, want, uv_tol=1e-5):
    """leaf ids / faces exact; t, p, normal bit-exact (NaN-aware); u, v within uv_tol (0 = bit-exact)"""
    assert np.array_equal(got["prim_id"], want["prim_id"]), f"{(got['prim_id'] != want['prim_id']).sum()} leaf ids differ"
    hit = want["prim_id"] >= 0
    assert np.array_equal(got["sub_id"][hit], want["sub_id"][hit])
    for f in ("t", "p", "normal"):
        a, b = got[f][hit], want[f][hit]
        same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b)) | ((a == 0) & (b == 0))
        assert same.all(), f"{f}: {(~same).sum()} of {same.size} values differ, max abs {np.nanmax(np.abs(a - b))}"
    for f in ("u", "v"):
        a, b = got[f][hit], want[f][hit]
        ok = (a == b) | (np.isnan(a) & np.isnan(b)) if uv_tol == 0 else np.isclose(a, b, rtol=uv_tol, atol=uv_tol) | (np.isnan(a) & np.isnan(b))
        assert ok.all(), f"{f}: max abs diff {np.nanmax(np.abs(a - b))}"
