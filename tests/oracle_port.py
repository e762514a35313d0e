"""ctypes bindings of oracle/_ref/librtnw_oracle.so — the plain-C restatement of the reference path (oracle/rtnw_oracle.c),
operating on the flattened tables of include/rtnw.h.  TEST INFRASTRUCTURE (see the header of the C file)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
PORT_SO = ROOT / "oracle" / "_ref" / "librtnw_oracle.so"
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not PORT_SO.exists():
            subprocess.run(["make", "-C", str(ROOT / "oracle"), "-s", "port"], check=True)
        L = C.CDLL(str(PORT_SO))
        L.rtnw_oracle_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_float, C.c_uint64, C.c_void_p]
        L.rtnw_oracle_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.rtnw_oracle_camera_rays.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p]
        L.rtnw_oracle_eval_texture.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]
        L.rtnw_oracle_eval_perlin.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]
        L.rtnw_oracle_scatter.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _desc_ptr(desc):
    return C.cast(desc, C.c_void_p) if isinstance(desc, C._Pointer) else C.cast(C.pointer(desc), C.c_void_p)


def trace(rtnw, desc, rays, t_min=0.001, t_max=None, seed=1):
    t_max = rtnw.FLT_MAX if t_max is None else t_max
    rays = np.ascontiguousarray(rays, dtype=rtnw.RAY_DTYPE)
    out = np.zeros(rays.shape[0], dtype=rtnw.HIT_DTYPE)
    rc = lib().rtnw_oracle_trace(_desc_ptr(desc), rays.ctypes.data, rays.shape[0], t_min, t_max, seed, out.ctypes.data)
    assert rc == 0
    return out


def render(rtnw, desc, cam, params, out=None):
    """returns (sums[ny,nx,3], stats dict); pass `out` to accumulate into / keep untouched pixels of an existing image"""
    if out is None:
        out = np.zeros((params.ny, params.nx, 3), dtype=np.float32)
    st = np.zeros(5, dtype=np.float64)
    rc = lib().rtnw_oracle_render(_desc_ptr(desc), C.cast(C.pointer(cam), C.c_void_p), C.cast(C.pointer(params), C.c_void_p),
                                  out.ctypes.data, st.ctypes.data)
    assert rc == 0
    return out, dict(zip(["paths", "rays", "box_tests", "prim_tests", "seconds"], st.tolist()))


def camera_rays(rtnw, cam, nx, ny, ij, sample, seed=1):
    ij = np.ascontiguousarray(ij, dtype=np.int32).reshape(-1, 2)
    sample = np.ascontiguousarray(sample, dtype=np.int32)
    out = np.zeros(ij.shape[0], dtype=rtnw.RAY_DTYPE)
    assert lib().rtnw_oracle_camera_rays(C.cast(C.pointer(cam), C.c_void_p), nx, ny, ij.ctypes.data, sample.ctypes.data, ij.shape[0],
                                         seed, out.ctypes.data) == 0
    return out


def eval_texture(rtnw, desc, tex, uvp):
    uvp = np.ascontiguousarray(uvp, dtype=np.float32).reshape(-1, 5)
    out = np.zeros((uvp.shape[0], 3), dtype=np.float32)
    assert lib().rtnw_oracle_eval_texture(_desc_ptr(desc), tex, uvp.ctypes.data, uvp.shape[0], out.ctypes.data) == 0
    return out


def eval_perlin(rtnw, desc, which, xyz):
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    out = np.zeros(xyz.shape[0], dtype=np.float32)
    assert lib().rtnw_oracle_eval_perlin(_desc_ptr(desc), which, xyz.ctypes.data, xyz.shape[0], out.ctypes.data) == 0
    return out


def scatter(rtnw, desc, rays_in, hits, seed=1):
    rays_in = np.ascontiguousarray(rays_in, dtype=rtnw.RAY_DTYPE)
    hits = np.ascontiguousarray(hits, dtype=rtnw.HIT_DTYPE)
    n = rays_in.shape[0]
    sc = np.zeros(n, dtype=rtnw.RAY_DTYPE)
    att = np.zeros((n, 3), dtype=np.float32)
    em = np.zeros((n, 3), dtype=np.float32)
    flag = np.zeros(n, dtype=np.int32)
    assert lib().rtnw_oracle_scatter(_desc_ptr(desc), rays_in.ctypes.data, hits.ctypes.data, n, seed, sc.ctypes.data, att.ctypes.data,
                                     em.ctypes.data, flag.ctypes.data) == 0
    return sc, att, em, flag
