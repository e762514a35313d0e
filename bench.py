#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: Mpaths/s (and Mrays/s) of the path-tracing sample loop on the final scene,
1000x1000 at 100 spp (config 5N), on N B200s of one box, with the reference's CPU renderer beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one full render of the workload: every (pixel, sample) path traced by ONE launch of k_render per GPU.
N > 1 (torchrun, one rank per GPU): rank g renders samples g, g+N, ... of every pixel, then a single NCCL reduce
(sum) of the float accumulation buffer to rank 0 — strong scaling, the partition the north star names.
value  = paths per second with the scene resident in HBM (CUDA events around the step on the launching stream).
e2e    = the same through the host-buffer C-ABI: scene tables uploaded from pinned host memory, render,
         accumulation buffer copied back to pinned host memory, every step.
"""
import argparse
import importlib
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

WORKLOAD = dict(scene="final_northstar", nx=1000, ny=1000, ns=100)
SEED = 20181025
# Algorithmic FP32 operations per ray on the reference topology: 18/aabb test, 25/sphere, 30/moving
# sphere, 10/rect, 60/medium, +100 shading (+800 per noise_texture hit, +60 per checker hit).
F_OPS = dict(aabb=18.0, sphere=25.0, moving_sphere=30.0, rect=10.0, medium=60.0, shade=100.0)
# per-ray test counts of config 5N measured with the reference's own counters (oracle/ref_harness.cpp ref_cnt) at
# 200x200x4 (sphere/rect counts include media boundaries and box faces); refreshed from the live cpu_baseline leg when
# it runs.  See DESIGN.md §7.
DEFAULT_COUNTS = dict(aabb=24.13, sphere=9.96, moving_sphere=1.0, rect=19.49, medium=2.0)


def flops_per_ray(c):
    return (F_OPS["aabb"] * c["aabb"] + F_OPS["sphere"] * c["sphere"] + F_OPS["moving_sphere"] * c["moving_sphere"] +
            F_OPS["rect"] * c["rect"] + F_OPS["medium"] * c["medium"] + F_OPS["shade"])


# ------------------------------------------------------------------------------------------------ reference arm
def _ref_worker(args):
    scene, nx, ny, ns, seed, begin = args
    import ref_oracle as ro
    rs = ro.RefScene(scene, tagged=False)
    _, st = rs.render(nx, ny, ns, seed=seed, rng_mode=0, sample_begin=begin)  # glibc drand48, as shipped
    return st


def run_reference(scene, nx, ny, ns_total, procs):
    """The reference's own CPU implementation (oracle/_ref/libref_oracle.so = its headers compiled unmodified, +F2 fix)
    on `procs` processes, each rendering ns_total/procs samples of every pixel with its own srand48 seed (the reference
    is single-threaded and drand48 is process-global, so processes are the only way to use more cores)."""
    import multiprocessing as mp
    per = max(1, ns_total // procs)
    jobs = [(scene, nx, ny, per, 1000 + k, k * per) for k in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        stats = [_ref_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            stats = pool.map(_ref_worker, jobs)
    wall = time.perf_counter() - t0
    paths = sum(s["paths"] for s in stats)
    rays = sum(s["rays"] for s in stats)
    loop_s = max(s["seconds"] for s in stats)  # the render loops run concurrently; scene construction excluded
    counts = {k: sum(s[k] for s in stats) / rays for k in ("aabb", "sphere", "moving_sphere", "rect", "medium")}
    return dict(paths=paths, rays=rays, seconds=loop_s, wall=wall, counts=counts)


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import ref_oracle as ro
    if not ro.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_oracle.so was not built"}))
        return 0
    procs = os.cpu_count() or 1
    # bounded sample of the workload: same scene and camera, 200x200 pixels, 24 samples per pixel per process and step
    # (about 6 s of CPU work per step at ~160 kpaths/s per core)
    nx = ny = 200
    ns = 24 * procs
    times, paths, rays = [], 0, 0
    for it in range(args.warmup + args.steps):
        r = run_reference(WORKLOAD["scene"], nx, ny, ns, procs)
        if it >= args.warmup:
            times.append(r["seconds"])
            paths += r["paths"]
            rays += r["rays"]
    total = sum(times)
    v = paths / total / 1e6
    line = {
        "impl": "reference", "metric": "Mpaths/s", "value": v, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "final_northstar 1000x1000x100spp (config 5N)", "sample": f"{nx}x{ny}x{ns}spp per step"},
        "mrays_per_s": rays / total / 1e6,
        "cpu_baseline": {"value": v, "unit": "Mpaths/s", "cores": procs, "kind": "reference",
                         "sample": f"{nx}x{ny} pixels x {ns} spp of the same scene/camera, {procs} processes, glibc drand48"},
        "e2e": {"value": v, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        top = sorted(self.samples)[len(self.samples) // 2:]  # upper half = samples taken under load
        return {"sm_mhz": statistics.median(top), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="rtnw", choices=["rtnw", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--flags", type=int, default=0, help="extra RTNW_F_* render flags (4 = RTNW_F_FAST_BVH, the non-reference-exact fast traversal mode)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_main(args)

    import numpy as np
    import torch
    rtnw = importlib.import_module("peter-shirley-ray-tracing-the-next-week_b200")
    mg = importlib.import_module("peter-shirley-ray-tracing-the-next-week_b200.multi_gpu")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    ctx = rtnw.Context(local)
    hs = rtnw.HostScene(WORKLOAD["scene"])
    ds = ctx.upload(hs.desc_ptr)
    nx, ny, ns = WORKLOAD["nx"], WORKLOAD["ny"], WORKLOAD["ns"]
    cam = hs.camera(nx, ny)
    plan = mg.partition_plan(ns, nx * ny, world, rank)  # one launch; even for any ns (RTNW_F_ROTATE_SAMPLES)

    def launch_params(launch):
        extra = args.flags | (rtnw.F_ROTATE_SAMPLES if launch["rotate"] else 0) | (rtnw.F_ACCUMULATE if launch["accumulate"] else 0)
        return hs.params(nx=nx, ny=ny, ns=launch["sample_count"], seed=SEED, sample_begin=launch["sample_begin"],
                         sample_stride=launch["sample_stride"], flags_extra=extra)
    params = launch_params(plan[0])  # world == 1: the whole frame in one launch
    accum = torch.empty(ny, nx, 3, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    last = {}

    def step():
        last["rays"], last["kernel_ms"], last["launches"] = 0, 0.0, 0

        def render(**launch):
            st = ds.render_device(cam, launch_params(launch), accum.data_ptr(), stream.cuda_stream)
            last["rays"] += st.rays; last["kernel_ms"] += st.kernel_ms; last["launches"] += st.kernel_launches; last["ranges"] = st.sample_ranges
        mg.render_partitioned(render, accum, ns, dist=dist, dst=0)  # k_render launches of this rank, then one NCCL reduce(sum)
        return last

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    rays = 0
    kernel_ms = []
    for a, b in ev:
        flush.zero_()  # L2 flush between timed iterations (outside the timed events)
        barrier()
        a.record(stream)
        st = step()
        b.record(stream)
        rays += st["rays"]
        kernel_ms.append(st["kernel_ms"])
        launches_per_step = st["launches"]
        sample_ranges = st["ranges"]
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    clocks = sampler.result()
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    tot_rays = torch.tensor([float(rays)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_rays, op=dist.ReduceOp.SUM)
    total_s = total_ms.item() / 1e3
    # per-rank kernel time and step time (diagnoses a slow GPU / a slow reduce when the max over ranks is off)
    mine = torch.tensor([sum(kernel_ms) / len(kernel_ms), sum(step_ms) / len(step_ms)], dtype=torch.float64, device=dev)
    per_rank = [torch.zeros_like(mine) for _ in range(world)]
    if dist is not None:
        dist.all_gather(per_rank, mine)
    else:
        per_rank = [mine]
    per_rank = [[round(float(x), 3) for x in t.tolist()] for t in per_rank]
    paths_per_step = nx * ny * ns
    value = paths_per_step * args.steps / total_s / 1e6
    mrays = tot_rays.item() / total_s / 1e6

    # ---- e2e: host buffers through the C-ABI, copies inside the timed region
    desc_bytes = rtnw_desc_bytes(rtnw, hs.desc)
    host_accum = torch.empty(ny, nx, 3, dtype=torch.float32).pin_memory()
    host_np = host_accum.numpy()

    e2e_kernel_ms = []

    def e2e_step():
        _t0 = time.perf_counter()
        s2 = ctx.upload(hs.desc_ptr)  # H2D of every scene table
        _t1 = time.perf_counter()
        if dist is None:
            _, st2 = s2.render(cam, params, out=host_np)  # render + D2H into pinned host memory
            e2e_kernel_ms.append(st2.kernel_ms)
        else:
            tot = {"ms": 0.0}

            def render2(**launch):
                tot["ms"] += s2.render_device(cam, launch_params(launch), accum.data_ptr(), stream.cuda_stream).kernel_ms
            mg.render_partitioned(render2, accum, ns, dist=dist, dst=0)
            e2e_kernel_ms.append(tot["ms"])
            if rank == 0:
                host_accum.copy_(accum, non_blocking=False)
        _t2 = time.perf_counter()
        s2.close()
        if os.environ.get("RTNW_BENCH_DEBUG"):
            print(f"e2e step: upload {1e3 * (_t1 - _t0):.2f} render {1e3 * (_t2 - _t1):.2f} close {1e3 * (time.perf_counter() - _t2):.2f} ms "
                  f"kernel {e2e_kernel_ms[-1]:.2f}", file=sys.stderr)

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = paths_per_step * args.steps / e2e_s.item() / 1e6

    # ---- FP32-issue roofline of k_render (the dominant and only kernel of the step)
    counts = dict(DEFAULT_COUNTS)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import ref_oracle as ro
        if ro.available():
            procs = os.cpu_count() or 1
            r = run_reference(WORKLOAD["scene"], 200, 200, 64 * procs, procs)  # ~16 s of CPU work per core
            counts = r["counts"]
            cpu = {"value": r["paths"] / r["seconds"] / 1e6, "unit": "Mpaths/s", "cores": procs, "kind": "reference",
                   "mrays_per_s": r["rays"] / r["seconds"] / 1e6,
                   "sample": f"200x200 pixels x {64 * procs} spp of the same scene/camera, {procs} processes "
                             f"(the reference is single-threaded), glibc drand48; {r['seconds']:.1f} s"}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        info = ctx.info()
        kernel_s = sum(kernel_ms) / 1e3 / len(kernel_ms)  # all k_render launches of one step on rank 0
        rays_per_launch = rays / args.steps
        f_ray = flops_per_ray(counts)
        sm_mhz = clocks["sm_mhz"] or (info["clock_khz"] / 1e3)
        fp32_nominal = info["sm_count"] * 128 * 2 * sm_mhz * 1e6 / 1e12  # TFLOP/s at the clock seen under load (FMA = 2)
        fp32_peak = ctx.fp32_peak_tflops()  # measured on this device: independent FFMA chains (rtnw_measure_fp32_peak)
        achieved = rays_per_launch * f_ray / kernel_s / 1e12
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # scene tables read once + one plane of partial sums per sample range written by k_render, read back and reduced to the
        # image by k_sum_chunks (0.09 ms of the step)
        hbm_bytes = desc_bytes + (2 * sample_ranges + 1) * nx * ny * 3 * 4
        line = {
            "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "final_northstar 1000x1000x100spp (config 5N: BVH over 1024 floor boxes + "
                                   "translate(rotate_y(BVH over 1000 spheres)) + media + perlin + image texture)",
                       "paths_per_step": paths_per_step, "partition": f"spp split over {world} GPU(s) (sample ownership rotates with the pixel index), 1 NCCL reduce; "
                                    f"{sample_ranges} sample ranges per pixel per launch",
                       "l2": "flushed between timed steps (256 MiB write)", "traversal": "fast (RTNW_F_FAST_BVH)" if args.flags & 4 else "reference-exact",
                       "seed": SEED},
            "mrays_per_s": mrays, "rays_per_path": tot_rays.item() / (paths_per_step * args.steps),
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": desc_bytes * world,
                    "d2h_bytes_per_step": nx * ny * 3 * 4, "ms_per_step": 1e3 * e2e_s.item() / args.steps,
                    "kernel_ms_per_step": sum(e2e_kernel_ms[-args.steps:]) / args.steps},
            "gpu_launches": args.steps * launches_per_step,
            "per_rank_ms": {"kernel": [r[0] for r in per_rank], "step": [r[1] for r in per_rank]},
            "clocks": clocks,
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one k_render launch of THIS workload (ncu capture of
                         # `bench.py --steps 1 --warmup 1`, profiles/round1_r15_bench_dram_traffic.csv): 16 MB read (scene tables,
                         # image texture) + 284 MB written (27 planes of per-sample-range partial sums, 324 MB, less what is
                         # still dirty in the 126 MB L2 when the kernel ends)
                         "traffic": 299885312, "kernel": "k_render", "kernel_ms": 1e3 * kernel_s, "flop_per_ray": f_ray,
                         "peak_source": f"measured FFMA microbenchmark on this GPU (nominal {fp32_nominal:.1f} = {info['sm_count']} SM x "
                                        f"128 lanes x 2 x {sm_mhz:.0f} MHz)",
                         "note": "FP32-issue bound, not HBM/tensor: the scene (<2 MB) lives in L1/L2; see the hbm sub-object",
                         "hbm": {"achieved": hbm_bytes / kernel_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": hbm_bytes / kernel_s / 1e9 / hbm_peak, "algorithmic_bytes": hbm_bytes,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    ds.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def rtnw_desc_bytes(rtnw, d):
    import ctypes as C
    return (d.n_items * C.sizeof(rtnw.Item) + d.n_nodes * C.sizeof(rtnw.BvhNode) + d.n_prim_slots * (C.sizeof(rtnw.Prim) + 4) +
            d.n_xform_ops * C.sizeof(rtnw.XformOp) + d.n_materials * C.sizeof(rtnw.Material) +
            d.n_textures * C.sizeof(rtnw.Texture) + int(d.image_bytes) + 768 * 4 + 3 * 256 * 4)


if __name__ == "__main__":
    sys.exit(main())
