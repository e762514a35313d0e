#!/usr/bin/env python
"""Known-answer fixtures from the images the reference ships (TNW/*.ppm, the only expected outputs in the repo):
per file its size and the mean of its 8-bit R, G, B values as written (unclipped: the Ch06-Ch08 snapshots do not clamp).
Writes tests/golden/shipped_ppm_means.json.  Needs /root/reference."""
import json
from pathlib import Path

import numpy as np

TNW = Path("/root/reference/The-Next-Week")
HERE = Path(__file__).resolve().parent


def read_p3(path):
    tok = path.read_text().split()
    assert tok[0] == "P3"
    nx, ny = int(tok[1]), int(tok[2])
    px = np.array(tok[4:4 + 3 * nx * ny], dtype=np.int64).reshape(ny, nx, 3)
    return nx, ny, px


def main():
    out = {}
    for f in sorted(TNW.glob("*.ppm")):
        nx, ny, px = read_p3(f)
        out[f.name] = dict(nx=nx, ny=ny, mean_rgb=[round(float(x), 4) for x in px.reshape(-1, 3).mean(0)], max=int(px.max()))
        print(f.name, out[f.name])
    (HERE / "shipped_ppm_means.json").write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
