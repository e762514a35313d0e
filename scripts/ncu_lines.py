#!/usr/bin/env python
"""Join an ncu SASS profile of k_render with nvdisasm line info: instruction / stall shares per source line.
usage: ncu_lines.py <report.ncu-rep> <lib.so> [kernel-substring] [top]"""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile
rep, lib = sys.argv[1], os.path.abspath(sys.argv[2])
kname = sys.argv[3] if len(sys.argv) > 3 else "k_renderILb0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c", glob.glob(tmp + "/*.cubin")[0]], capture_output=True, text=True).stdout
lines, cur, on = [], ("?", 0), False
for ln in dis.splitlines():
    if ln.startswith(".text."):
        on = kname in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = rows[2:]; ix = {n: i for i, n in enumerate(h)}
assert len(data) == len(lines), (len(data), len(lines))
def f(r, n):
    try: return float(r[ix[n]])
    except Exception: return 0.0
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for r, loc in zip(data, lines):
    a = agg[loc]; a[0] += f(r, "Instructions Executed"); a[1] += f(r, "Thread Instructions Executed"); a[2] += f(r, "# Samples")
ti = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values())
text = {}
for fn in set(l[0] for l in agg):
    for root in ("peter-shirley-ray-tracing-the-next-week_b200/csrc", "."):
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", root, fn)
        if os.path.exists(p):
            text[fn] = open(p).read().splitlines()
print(f"{'file:line':28s} {'inst%':>6s} {'thr':>5s} {'stall%':>6s}  source")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    t = text.get(loc[0], [])
    s = t[loc[1] - 1].strip()[:90] if 0 < loc[1] <= len(t) else ""
    print(f"{loc[0] + ':' + str(loc[1]):28s} {a[0] / ti * 100:6.2f} {a[1] / max(a[0], 1):5.1f} {a[2] / ts * 100:6.2f}  {s}")
# ---- shares by line range of rtnw_device.cuh (regions given as name:lo-hi,... in env NCU_REGIONS)
reg = os.environ.get("NCU_REGIONS")
if reg:
    print("\nregions of rtnw_device.cuh:")
    for item in reg.split(","):
        name, rng = item.split(":"); lo, hi = map(int, rng.split("-"))
        sel = [a for l, a in agg.items() if l[0] == "rtnw_device.cuh" and lo <= l[1] <= hi]
        a0, a1, a2 = (sum(a[q] for a in sel) for q in range(3))
        print(f"  {name:16s} inst {a0 / ti * 100:6.2f}%  thr {a1 / max(a0, 1):5.1f}  stall {a2 / ts * 100:6.2f}%")
    byf = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
    for l, a in agg.items():
        if l[0] != "rtnw_device.cuh":
            for q in range(3): byf[l[0]][q] += a[q]
    for fn, a in byf.items():
        print(f"  {fn:32s} inst {a[0] / ti * 100:6.2f}%  thr {a[1] / max(a[0], 1):5.1f}  stall {a[2] / ts * 100:6.2f}%")
