#!/usr/bin/env python
"""Statistical image fixtures from the reference running on ITS OWN generator (glibc drand48, rng_mode 0): K independent
batches per scene (distinct srand48 seeds), stored as per-batch pixel means.  Used by tests/test_statistical_parity.py
where the GPU renders the same scene with the framework's stream — independent samples, so agreement is statistical."""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import ref_oracle as ro  # noqa: E402

CASES = [("cornell_box", 40, 40, 64), ("final_northstar", 40, 40, 32), ("ch01_random", 48, 24, 64)]
K = 8


def main():
    for name, nx, ny, spp in CASES:
        rs = ro.RefScene(name, tagged=False)
        batches = np.stack([rs.render(nx, ny, spp, seed=5000 + 17 * k, rng_mode=0)[0] / spp for k in range(2 * K)])
        np.savez_compressed(HERE / f"stat_{name}.npz", batches=batches.astype(np.float32), nx=nx, ny=ny, spp=spp, k=K)
        print(name, batches.shape, float(batches.mean()))


if __name__ == "__main__":
    main()
