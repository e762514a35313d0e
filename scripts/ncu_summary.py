#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + instruction mix / SIMT width / top stall lines (needs ncu on PATH)."""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__sass_average_branch_targets_threads_uniform.pct", "sm__cycles_elapsed.max"]
for h, u, v in zip(hdr, units, vals):
    if h in keys or h.startswith("smsp__average_warps_issue_stalled") and float(v or 0) > 0.05:
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = rows[2:]; ix = {n: i for i, n in enumerate(h)}
def f(r, n):
    try: return float(r[ix[n]])
    except Exception: return 0.0
ti = sum(f(r, "Instructions Executed") for r in data); tt = sum(f(r, "Thread Instructions Executed") for r in data)
ts = sum(f(r, "# Samples") for r in data)
print(f"\nSASS instructions {len(data)}  warp-inst {ti:.3e}  thread-inst {tt:.3e}  avg threads/inst {tt / ti:.2f}")
by = collections.defaultdict(lambda: [0, 0, 0])
for r in data:
    s = r[ix["Source"]].split()
    op = (s[1] if s and s[0].startswith("@") and len(s) > 1 else (s[0] if s else "?")).split(".")[0]
    by[op][0] += f(r, "Instructions Executed"); by[op][1] += f(r, "Thread Instructions Executed"); by[op][2] += f(r, "# Samples")
print("opcode mix (share of warp instructions, avg active threads, share of stall samples):")
for op, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:18]:
    print(f"  {op:10s} {v[0] / ti * 100:5.1f}%  thr {v[1] / max(v[0], 1):5.1f}  samples {v[2] / ts * 100:5.1f}%")
print(f"top {topn} SASS lines by stall samples:")
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:topn]:
    st = sorted(((f(r, n), n) for n in stall_cols), reverse=True)[:2]
    print(f"  {r[ix['Address']][-5:]} {f(r, '# Samples') / ts * 100:5.2f}% thr {f(r, 'Avg. Threads Executed'):4.1f} "
          f"x{f(r, 'Instructions Executed'):.2e} {r[ix['Source']][:60]:60s} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}")
