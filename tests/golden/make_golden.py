#!/usr/bin/env python
"""Generate the committed golden fixtures from the REFERENCE ITSELF (oracle/_ref/libref_oracle.so, built from
/root/reference by oracle/Makefile: the reference's sources compiled unmodified + the F2 aabb fix).

    python tests/golden/make_golden.py        # run in the build container; needs /root/reference

Per scene:  trace_<scene>.npz   rays (primary + secondary + stress), reference hits for two t-ranges
            render_<scene>.npz  per-pixel float sums of a small render driven by the shared sample stream
plus        units.npz           perlin noise/turb, checker/noise/image texture values, camera rays, scatter vectors
The fixtures travel to the GPU box, where /root/reference does not exist."""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import ref_oracle as ro  # noqa: E402
from raysets import FLT_MAX, make_rays  # noqa: E402

TRACE = ["ch01_random", "two_perlin", "cornell_box", "cornell_smoke", "final", "final+bvh", "final_northstar", "earth",
         "simple_light", "cornell_smoke+bvh", "random_scene+bvh", "test", "two_spheres"]
RENDER = [("ch01_random", 32, 16, 4), ("two_perlin", 32, 16, 4), ("cornell_box", 24, 24, 6), ("cornell_smoke", 24, 24, 6),
          ("final", 20, 20, 2), ("final+bvh", 24, 24, 3), ("final_northstar", 48, 48, 4), ("simple_light", 32, 16, 4), ("earth", 20, 20, 3), ("random_scene", 40, 20, 4), ("test", 32, 16, 4), ("two_spheres", 24, 24, 4),
          ("perlin_v1", 32, 16, 4), ("perlin_v2", 32, 16, 4), ("perlin_v3", 32, 16, 4)]


def fname(kind, scene):
    return HERE / f"{kind}_{scene.replace('+', '_')}.npz"


def main(only=None):
    assert ro.available(), "build oracle/_ref/libref_oracle.so first (make -C oracle ref)"
    for name in TRACE:
        if only and name not in only:
            continue
        rs, rays = make_rays(name, n_primary=220, seed=21)
        np.savez_compressed(fname("trace", name), rays=rays, hits_a=rs.trace(rays, 0.001, FLT_MAX, seed=13),
                            hits_b=rs.trace(rays, 0.0, 400.0, seed=13), seed=13)
    for name, nx, ny, ns in RENDER:
        if only and name not in only:
            continue
        sums, st = ro.RefScene(name, tagged=True).render(nx, ny, ns, seed=4242, rng_mode=1)
        np.savez_compressed(fname("render", name), sums=sums, nx=nx, ny=ny, ns=ns, seed=4242, rays=st["rays"], aabb=st["aabb"])
    if not only or only & {"perlin_v1", "perlin_v2", "perlin_v3", "readme_units"}:
        # the Chapter 4 noise drafts (README.md:516-630, restated in the harness): texture values at fixed points
        rng = np.random.default_rng(6)
        ro.RefScene("perlin_v1", tagged=False)  # their tables: ranfloat + permutations of the never-seeded stream
        xyz = np.concatenate([rng.normal(scale=3, size=(600, 3)), rng.normal(scale=300, size=(300, 3)), rng.integers(-4, 5, (100, 3))]).astype(np.float32)
        uvp = np.concatenate([np.zeros((len(xyz), 2), np.float32), xyz], axis=1)
        np.savez_compressed(HERE / "units_readme_noise.npz", uvp=uvp, v1=ro.eval_texture(4, [], uvp), v2=ro.eval_texture(5, [], uvp),
                            v3=ro.eval_texture(6, [], uvp))
    if only:
        return
    rng = np.random.default_rng(5)
    ro.RefScene("two_perlin", tagged=False)  # perlin tables of the never-seeded drand48 stream
    xyz = np.concatenate([rng.normal(scale=3, size=(400, 3)), rng.normal(scale=300, size=(400, 3))]).astype(np.float32)
    uvp = np.concatenate([rng.random((len(xyz), 2)).astype(np.float32), xyz], axis=1)
    v = ro.view_of("ch01_random")
    ij = np.stack([rng.integers(0, 200, 300), rng.integers(0, 100, 300)], axis=1).astype(np.int32)
    smp = rng.integers(0, 100, 300).astype(np.int32)
    n = 300
    rays = np.zeros(n, dtype=ro.RAY_DTYPE)
    rays["origin"] = rng.normal(size=(n, 3)); rays["direction"] = rng.normal(size=(n, 3)); rays["time"] = rng.random(n)
    hits = np.zeros(n, dtype=ro.HIT_DTYPE)
    nrm = rng.normal(size=(n, 3)); hits["normal"] = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
    hits["p"] = rng.normal(scale=3, size=(n, 3)); hits["t"] = 1; hits["u"] = rng.random(n); hits["v"] = rng.random(n)
    mats = {"lambertian": [0, 0, 0.4, 0.2, 0.1, 0, 0, 0], "metal": [1, 0, 0.8, 0.8, 0.9, 0.3, 0, 0], "dielectric": [2, 0, 0, 0, 0, 1.5, 0, 0],
            "light": [3, 0, 7, 7, 7, 0, 0, 0], "isotropic": [4, 0, 0.2, 0.4, 0.9, 0, 0, 0]}
    out = dict(xyz=xyz, noise=ro.eval_perlin(0, xyz), turb=ro.eval_perlin(1, xyz), uvp=uvp,
               checker=ro.eval_texture(1, [0.2, 0.3, 0.1, 0.9, 0.9, 0.9], uvp), noise_tex=ro.eval_texture(2, [4.0], uvp),
               image=ro.eval_texture(3, [], uvp), cam_ij=ij, cam_sample=smp, cam_rays=ro.camera_rays(v, 200, 100, ij, smp, seed=77),
               sc_rays=rays, sc_hits=hits)
    for k, row in mats.items():
        sc, att, em, flag = ro.scatter(np.tile(np.array(row, dtype=np.float32), (n, 1)), rays, hits, seed=31)
        out.update({f"{k}_row": np.array(row, dtype=np.float32), f"{k}_sc": sc, f"{k}_att": att, f"{k}_em": em, f"{k}_flag": flag})
    np.savez_compressed(HERE / "units.npz", **out)
    print("wrote", sorted(p.name for p in HERE.glob("*.npz")))


if __name__ == "__main__":
    main(set(sys.argv[1:]) or None)  # optional scene names: regenerate only those trace/render fixtures
