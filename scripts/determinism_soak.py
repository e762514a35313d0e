#!/usr/bin/env python
"""Soak test for races: render the same frame many times and compare bits (tuning / debugging aid).
usage: determinism_soak.py [scene] [reps] [nx] [ns] [flags]"""
import importlib, sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
rtnw = importlib.import_module("peter-shirley-ray-tracing-the-next-week_b200")
name = sys.argv[1] if len(sys.argv) > 1 else "final+bvh"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
nx = ny = int(sys.argv[3]) if len(sys.argv) > 3 else 160
ns = int(sys.argv[4]) if len(sys.argv) > 4 else 6
flags = int(sys.argv[5]) if len(sys.argv) > 5 else 0
ctx = rtnw.Context(0)
hs = rtnw.HostScene(name)
ds = ctx.upload(hs.desc_ptr)
cam = hs.camera(nx, ny)
a, sa = ds.render(cam, hs.params(nx=nx, ny=ny, ns=ns, seed=77, flags_extra=flags))
bad = 0
for i in range(reps):
    b, sb = ds.render(cam, hs.params(nx=nx, ny=ny, ns=ns, seed=77, flags_extra=flags))
    d = a.view(np.uint32) != b.view(np.uint32)
    if d.any() or sa.rays != sb.rays:
        bad += 1
        idx = np.argwhere(d.any(axis=2))
        print(f"rep {i}: {d.any(axis=2).sum()} pixels differ, rays {sa.rays} vs {sb.rays}, first {idx[:3].tolist()}, "
              f"max rel {np.nanmax(np.abs(a - b) / (np.abs(a) + 1e-9)):.3g}")
print(f"{name} {nx}x{ny}x{ns} flags {flags}: {bad} of {reps} repetitions differ")
