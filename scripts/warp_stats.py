#!/usr/bin/env python
"""Summarise RTNW_DEBUG_WARPS output (per-warp scheduler statistics of k_render, tuning aid)."""
import sys, numpy as np
a = np.loadtxt(sys.argv[1], dtype=np.float64, ndmin=2)
it, cyc = a[:, 0], a[:, 1]
print(f"warps {len(a)}  iters mean {it.mean():.0f} max {it.max():.0f}  cycles mean {cyc.mean():.3e} max {cyc.max():.3e}  "
      f"cycles/iter mean {(cyc / it).mean():.0f}")
names = ["node", "sphere", "box", "misc", "shade"]
steps, lanes = a[:, 2:7].sum(0), a[:, 7:12].sum(0)
for n, s, l in zip(names, steps, lanes):
    print(f"  {n:7s} steps {s / steps.sum() * 100:5.1f}%  avg lanes {l / max(s, 1):5.1f}")
order = np.argsort(-cyc)[:5]
for w in order:
    print("  straggler", a[w].astype(np.int64).tolist())
