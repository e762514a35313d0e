#!/usr/bin/env python
"""Golden vectors for the epilogue of the sample loop (PSC/main.cpp:315-325): float sums -> 8-bit values, computed by the
reference's own lines (oracle/ref_harness.cpp: ref_epilogue).  Needs oracle/_ref/libref_oracle.so."""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import ref_oracle as ro  # noqa: E402


def sums_for(ns, rng, n=4096):
    """per-pixel sums covering: dark, mid, > 1 (lights), exact quantisation boundaries +- 1 ulp, zero, huge"""
    v = np.concatenate([rng.random(n) * ns, rng.random(n) ** 4 * ns, rng.random(n // 4) * 40 * ns, [0.0, ns, 4 * ns, 1e30]])
    k = np.arange(0, 300, dtype=np.float64)
    edge = ((k / 255.99) ** 2 * ns).astype(np.float32)  # 255.99 * sqrt(x / ns) lands on an integer (up to rounding)
    v = np.concatenate([v.astype(np.float32), edge, np.nextafter(edge, np.float32(np.inf)), np.nextafter(edge, np.float32(-np.inf))])
    v = v[v >= 0]
    pad = (-len(v)) % 48
    v = np.concatenate([v, np.zeros(pad, np.float32)])
    return v.reshape(-1, 16, 3)


def main():
    rng = np.random.default_rng(20181025)
    out = {}
    for ns in (1, 7, 10, 100, 1000):
        s = sums_for(ns, rng)
        out[f"sums_{ns}"] = s
        out[f"clamped_{ns}"] = ro.epilogue(s, ns, True)
        out[f"raw_{ns}"] = ro.epilogue(s, ns, False)
    np.savez_compressed(HERE / "epilogue.npz", **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
