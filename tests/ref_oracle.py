"""ctypes bindings of oracle/_ref/libref_oracle.so — the UNMODIFIED reference renderer compiled behind a C ABI
(oracle/ref_harness.cpp, built by oracle/Makefile from /root/reference; the binary travels to the GPU box).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF_SO = ROOT / "oracle" / "_ref" / "libref_oracle.so"

RAY_DTYPE = np.dtype([("origin", np.float32, 3), ("direction", np.float32, 3), ("time", np.float32), ("key", np.uint32)])
HIT_DTYPE = np.dtype([("prim_id", np.int32), ("sub_id", np.int32), ("t", np.float32), ("p", np.float32, 3),
                      ("normal", np.float32, 3), ("u", np.float32), ("v", np.float32), ("mat_id", np.int32)])

_lib = None


def available() -> bool:
    return REF_SO.exists()


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(str(REF_SO))
        L.ref_scene_build.argtypes = [C.c_char_p, C.c_int]
        L.ref_scene_build.restype = C.c_void_p
        L.ref_scene_leaf_count.argtypes = [C.c_void_p]
        L.ref_scene_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_perlin_tables.argtypes = [C.c_void_p] * 4
        L.ref_perlin_tables.restype = None
        L.ref_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_float, C.c_uint64, C.c_void_p]
        L.ref_trace.restype = None
        L.ref_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
        L.ref_render.restype = None
        L.ref_camera_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                                      C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p]
        L.ref_camera_rays.restype = None
        L.ref_eval_texture.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.ref_eval_texture.restype = None
        L.ref_eval_perlin.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        L.ref_eval_perlin.restype = None
        L.ref_scatter.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]
        L.ref_scatter.restype = None
        L.ref_epilogue.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ref_epilogue.restype = None
        _lib = L
    return _lib


# camera + integrator settings per scene name, mirroring rtnw_scenes::view_* (host/scenes/chapter_scenes.cpp)
VIEWS = {
    "ch01_random": dict(lookfrom=(13, 2, 3), lookat=(0, 0, 0), vfov=20, aperture=0.1, t_min=0.001, sky=1, emit=0, denan=0),
    "two_perlin": dict(lookfrom=(13, 2, 3), lookat=(0, 0, 0), vfov=20, aperture=0.0, t_min=0.001, sky=1, emit=0, denan=0),
    "simple_light": dict(lookfrom=(13, 2, 3), lookat=(0, 0, 0), vfov=20, aperture=0.0, t_min=0.001, sky=0, emit=1, denan=0),
    "cornell_box": dict(lookfrom=(278, 278, -800), lookat=(278, 278, 0), vfov=40, aperture=0.0, t_min=0.001, sky=0, emit=1, denan=1),
    "cornell_smoke": dict(lookfrom=(278, 278, -800), lookat=(278, 278, 0), vfov=40, aperture=0.0, t_min=0.001, sky=0, emit=1, denan=1),
    "earth": dict(lookfrom=(278, 278, -800), lookat=(278, 278, 0), vfov=40, aperture=0.0, t_min=0.001, sky=0, emit=1, denan=1),
    "two_spheres": dict(lookfrom=(278, 278, -800), lookat=(278, 278, 0), vfov=40, aperture=0.0, t_min=0.001, sky=1, emit=1, denan=1),
    "random_scene": dict(lookfrom=(13, 2, 3), lookat=(0, 0, 0), vfov=20, aperture=0.1, t_min=0.001, sky=1, emit=1, denan=0),
    "test": dict(lookfrom=(13, 2, 3), lookat=(0, 0, 0), vfov=20, aperture=0.0, t_min=0.001, sky=0, emit=1, denan=0),
    "final": dict(lookfrom=(228, 278, -800), lookat=(278, 278, 0), vfov=40, aperture=0.0, t_min=0.001, sky=0, emit=1, denan=1),
    "perlin_v1": dict(lookfrom=(13, 2, 3), lookat=(0, 0, 0), vfov=20, aperture=0.1, t_min=0.0, sky=1, emit=0, denan=0),
    "perlin_v2": dict(lookfrom=(13, 2, 3), lookat=(0, 0, 0), vfov=20, aperture=0.1, t_min=0.0, sky=1, emit=0, denan=0),
    "perlin_v3": dict(lookfrom=(13, 2, 3), lookat=(0, 0, 0), vfov=20, aperture=0.1, t_min=0.0, sky=1, emit=0, denan=0),
    "final_northstar": dict(lookfrom=(228, 278, -800), lookat=(278, 278, 0), vfov=40, aperture=0.0, t_min=0.001, sky=0, emit=1, denan=1),
}


def view_of(name: str) -> dict:
    base, _, snapshot = name.partition(":")
    v = dict(VIEWS[base.split("+")[0]])
    v.update(focus_dist=10.0, time0=0.0, time1=1.0)
    if snapshot in ("ch01", "ch03"):    # TNW/Chapter01_Motion Blur.cpp:16, TNW/Chapter03_Soild Texture.cpp:16
        v.update(t_min=0.0, aperture=0.1, sky=1, emit=0, denan=0)
    elif snapshot in ("ch07", "ch08"):  # TNW/Chapter07_Instance.cpp:25,173-178, TNW/Chapter08_Volume.cpp:26
        v.update(t_min=0.01, aperture=0.1, lookfrom=(278, 278, -800), denan=0)
    elif snapshot:
        raise KeyError(snapshot)
    return v


class RefScene:
    """A scene built by the reference's own classes; tagged=True wraps leaves so hits report leaf ids."""

    def __init__(self, name: str, tagged: bool = True):
        self.name = name
        self._h = lib().ref_scene_build(name.partition(":")[0].encode(), int(tagged))
        if not self._h:
            raise ValueError(f"unknown scene {name}")

    @property
    def leaf_count(self) -> int:
        return lib().ref_scene_leaf_count(self._h)

    def dump(self) -> np.ndarray:
        out = np.zeros((self.leaf_count, 24), dtype=np.float32)
        lib().ref_scene_dump(self._h, out.ctypes.data, self.leaf_count)
        return out

    def trace(self, rays: np.ndarray, t_min=0.001, t_max=float(np.finfo(np.float32).max), seed=1) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        out = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        lib().ref_trace(self._h, rays.ctypes.data, rays.shape[0], t_min, t_max, seed, out.ctypes.data)
        return out

    def render(self, nx, ny, ns, seed=1, rng_mode=1, sample_begin=0, sample_stride=1, max_depth=50, view=None):
        """Returns (sums[ny,nx,3], stats dict).  rng_mode 1 = the framework's Philox stream, 0 = glibc drand48."""
        v = view or view_of(self.name)
        f3 = C.c_float * 3
        accum = np.zeros((ny, nx, 3), dtype=np.float32)
        stats = np.zeros(9, dtype=np.float64)
        lib().ref_render(self._h, f3(*v["lookfrom"]), f3(*v["lookat"]), v["vfov"], v["aperture"], v["focus_dist"], v["time0"],
                         v["time1"], nx, ny, sample_begin, ns, sample_stride, max_depth, v["t_min"], v["sky"], v["emit"],
                         v["denan"], rng_mode, seed, accum.ctypes.data, stats.ctypes.data)
        keys = ["paths", "rays", "draws", "aabb", "sphere", "moving_sphere", "rect", "medium", "seconds"]
        return accum, dict(zip(keys, stats.tolist()))


def perlin_tables():
    rv = np.zeros(768, dtype=np.float32)
    px, py, pz = (np.zeros(256, dtype=np.int32) for _ in range(3))
    lib().ref_perlin_tables(rv.ctypes.data, px.ctypes.data, py.ctypes.data, pz.ctypes.data)
    return rv, px, py, pz


def camera_rays(view: dict, nx, ny, ij, sample, seed=1) -> np.ndarray:
    f3 = C.c_float * 3
    ij = np.ascontiguousarray(ij, dtype=np.int32).reshape(-1, 2)
    sample = np.ascontiguousarray(sample, dtype=np.int32)
    out = np.zeros(ij.shape[0], dtype=RAY_DTYPE)
    lib().ref_camera_rays(f3(*view["lookfrom"]), f3(*view["lookat"]), view["vfov"], view["aperture"], view["focus_dist"],
                          view["time0"], view["time1"], nx, ny, ij.ctypes.data, sample.ctypes.data, ij.shape[0], seed,
                          out.ctypes.data)
    return out


def eval_texture(which: int, consts, uvp: np.ndarray) -> np.ndarray:
    """which: 0 constant(c0..2), 1 checker(even c0..2, odd c3..5), 2 noise(scale c0), 3 synthetic-earth image"""
    c = np.zeros(6, dtype=np.float32)
    c[:len(consts)] = consts
    uvp = np.ascontiguousarray(uvp, dtype=np.float32).reshape(-1, 5)
    out = np.zeros((uvp.shape[0], 3), dtype=np.float32)
    lib().ref_eval_texture(which, c.ctypes.data, uvp.ctypes.data, uvp.shape[0], out.ctypes.data)
    return out


def eval_perlin(which: int, xyz: np.ndarray) -> np.ndarray:
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    out = np.zeros(xyz.shape[0], dtype=np.float32)
    lib().ref_eval_perlin(which, xyz.ctypes.data, xyz.shape[0], out.ctypes.data)
    return out


def scatter(mat: np.ndarray, rays_in: np.ndarray, hits: np.ndarray, seed=1):
    """mat: n x 8 float32 {kind, tex_kind(0 const / 2 noise), r, g, b, fuzz_or_ri, scale, 0}"""
    mat = np.ascontiguousarray(mat, dtype=np.float32).reshape(-1, 8)
    rays_in = np.ascontiguousarray(rays_in, dtype=RAY_DTYPE)
    hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
    n = mat.shape[0]
    sc = np.zeros(n, dtype=RAY_DTYPE)
    att = np.zeros((n, 3), dtype=np.float32)
    em = np.zeros((n, 3), dtype=np.float32)
    flag = np.zeros(n, dtype=np.int32)
    lib().ref_scatter(mat.ctypes.data, rays_in.ctypes.data, hits.ctypes.data, n, seed, sc.ctypes.data, att.ctypes.data,
                      em.ctypes.data, flag.ctypes.data)
    return sc, att, em, flag


def epilogue(sums: np.ndarray, ns: int, clamp255: bool = True) -> np.ndarray:
    """PSC/main.cpp:315-325 (the reference's own lines) on (ny, nx, 3) float32 sums; returns (ny, nx, 3) int32, top row first"""
    sums = np.ascontiguousarray(sums, dtype=np.float32)
    ny, nx, _ = sums.shape
    out = np.zeros((ny, nx, 3), dtype=np.int32)
    lib().ref_epilogue(sums.ctypes.data, nx, ny, ns, int(clamp255), out.ctypes.data)
    return out
