#!/usr/bin/env python
"""DRAM traffic of k_render per bench config, from ncu (bench.py reports it as roofline.traffic).

On the GPU box, after `python scripts/capture_traffic.py --run` has exited 0 without a profiler:
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_render \
        --csv --log-file gpurun_out/traffic.csv python scripts/capture_traffic.py --run
Here:  python scripts/capture_traffic.py --parse gpurun_out/traffic.csv   -> profiles/round2_dram_traffic.json
--run launches k_render once per (config, GPU count) in the fixed order of PLAN, each exactly as bench.py's rank 0 does
(same scene, size, spp, seed and sample split), so launch k of the capture is PLAN[k]."""
import argparse
import csv
import importlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

PLAN = [(c, 1) for c in ("1", "2", "3", "4", "5R", "5R+bvh", "5N")] + [("5N", 2), ("5N", 4), ("5N", 8)]


def run():
    rtnw = importlib.import_module("peter-shirley-ray-tracing-the-next-week_b200")
    ctx = rtnw.Context(0)
    for config, world in PLAN:
        scene, nx, ny, ns, _, _ = bench.CONFIGS[config]
        hs = rtnw.HostScene(scene)
        ds = ctx.upload(hs.desc_ptr)
        extra = rtnw.F_ROTATE_SAMPLES if world > 1 else 0
        _, st = ds.render(hs.camera(nx, ny), hs.params(nx=nx, ny=ny, ns=ns, seed=bench.SEED, sample_begin=0, sample_stride=world, flags_extra=extra))
        print(f"{config}@{world}: {st.paths} paths, kernel {st.kernel_ms:.2f} ms, {st.sample_ranges} sample ranges", flush=True)
        ds.close()


def parse(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    head = next(i for i, r in enumerate(rows) if "Metric Name" in r)
    ix = {n: i for i, n in enumerate(rows[head])}
    per = {}
    for r in rows[head + 1:]:
        if "k_render" not in r[ix["Kernel Name"]]:
            continue
        e = per.setdefault(int(r[ix["ID"]]), {})
        v, unit = float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1)  # bytes / ms
        e[r[ix["Metric Name"]]] = v * scale
    ids = sorted(per)
    assert len(ids) == len(PLAN), f"{len(ids)} k_render launches in the capture, {len(PLAN)} planned"
    out = {"source_id": bench.source_id(), "tool": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none", "configs": {}}
    for (config, world), i in zip(PLAN, ids):
        rd, wr = int(per[i]["dram__bytes_read.sum"]), int(per[i]["dram__bytes_write.sum"])
        out["configs"][f"{config}@{world}"] = {"read": rd, "write": wr, "bytes": rd + wr, "kernel_ms_under_ncu": round(per[i].get("gpu__time_duration.sum", 0.0), 3)}
    (ROOT / "profiles" / "round2_dram_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--run", action="store_true")
    ap.add_argument("--parse")
    a = ap.parse_args()
    if a.parse:
        parse(a.parse)
    else:
        run()
