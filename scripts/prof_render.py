#!/usr/bin/env python
"""Render one scene once or a few times and print throughput; the command ncu wraps (see profiles/README.md)."""
import argparse, importlib, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rtnw = importlib.import_module("peter-shirley-ray-tracing-the-next-week_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="final_northstar")
ap.add_argument("--nx", type=int, default=1000)
ap.add_argument("--ny", type=int, default=1000)
ap.add_argument("--ns", type=int, default=8)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--flags", type=int, default=0)
a = ap.parse_args()
ctx = rtnw.Context(0)
hs = rtnw.HostScene(a.scene)
ds = ctx.upload(hs.desc_ptr)
cam = hs.camera(a.nx, a.ny)
for r in range(a.reps):
    out, st = ds.render(cam, hs.params(nx=a.nx, ny=a.ny, ns=a.ns, seed=5, flags_extra=a.flags))
    extra = f" box/ray {st.box_tests / st.rays:.1f} prim/ray {st.prim_tests / st.rays:.1f}" if a.flags & 8 else ""
    print(f"{a.scene} {a.nx}x{a.ny}x{a.ns}: kernel {st.kernel_ms:.2f} ms  {st.paths / st.kernel_ms / 1e3:.1f} Mpaths/s  "
          f"{st.rays / st.kernel_ms / 1e3:.1f} Mrays/s  rays/path {st.rays / st.paths:.2f}{extra}")
