#!/usr/bin/env python
"""Statistical image fixtures from the reference running on ITS OWN generator (glibc drand48, rng_mode 0).

Per scene: two independent halves A and B, each K = 8 batches of 256-1024 samples per pixel (distinct srand48 seeds), i.e.
>= 2048 spp per half at >= 100x100 pixels (tests/test_statistical_parity.py: the GPU renders the same scene with the
framework's stream — independent samples, so agreement is statistical; half B is the CPU-vs-CPU rerun that states the
noise floor).  Stored per half: pixel mean over the batches, variance of the batch means, and the per-batch global
channel means.  Run here (needs oracle/_ref/libref_oracle.so, i.e. /root/reference); ~2 min on 8 cores."""
import multiprocessing as mp
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import ref_oracle as ro  # noqa: E402

# (fixture name, reference scene, nx, ny): the Ch01 scene is rendered through the F2-patched bvh_node on the reference side
# (same scene, same distribution, 5x faster than its flat list)
# samples per batch: the Cornell scenes (small light, heavy-tailed pixels) get more, so that the standard error of their
# global mean is well inside the 0.5 % bound
CASES = [("cornell_box", "cornell_box", 100, 100, 1024), ("cornell_smoke", "cornell_smoke", 100, 100, 512),
         ("two_perlin", "two_perlin", 144, 72, 256), ("final_northstar", "final_northstar", 100, 100, 256),
         ("ch01_random", "ch01_random+bvh", 144, 72, 256)]
K = 8


def _batch(args):
    scene, nx, ny, spp, k = args
    rs = ro.RefScene(scene, tagged=False)
    return (rs.render(nx, ny, spp, seed=7000 + 31 * k, rng_mode=0)[0] / spp).astype(np.float64)


def main():
    only = set(sys.argv[1:])
    with mp.get_context("fork").Pool(min(16, mp.cpu_count())) as pool:
        for name, scene, nx, ny, spp in CASES:
            if only and name not in only:
                continue
            b = np.stack(pool.map(_batch, [(scene, nx, ny, spp, k) for k in range(2 * K)]))
            out = dict(nx=nx, ny=ny, spp=spp, k=K)
            for tag, h in (("a", b[:K]), ("b", b[K:])):
                out[f"mean_{tag}"] = h.mean(0).astype(np.float32)
                out[f"var_{tag}"] = h.var(0, ddof=1).astype(np.float32)      # variance of the K batch means, per pixel channel
                out[f"glob_{tag}"] = h.mean(axis=(1, 2)).astype(np.float64)   # (K, 3) global channel means per batch
            np.savez_compressed(HERE / f"stat_{name}.npz", **out)
            print(name, b.shape, out["glob_a"].mean(0), out["glob_b"].mean(0), flush=True)


if __name__ == "__main__":
    main()
