"""Multi-GPU partition of the sample loop: one process per GPU, rank g of G renders samples
g, g+G, g+2G, ... of EVERY pixel into a private float accumulation buffer, then ONE reduce(sum) to rank 0 before the
host epilogue (mean, gamma, quantise, PPM — PSC/main.cpp:315-330).  The sample stream is keyed by (seed, pixel, sample),
so the set of paths is the same for any G; results differ only by float summation order.  There is no other exchange
step on this path, hence no other collective.
"""
from __future__ import annotations


def partition_plan(ns: int, npix: int, world: int, rank: int) -> list[dict]:
    """The launches of `rank` (keyword overrides for the render parameters).  One launch: with RTNW_F_ROTATE_SAMPLES the
    ownership of the samples rotates with the pixel index — rank g renders, for pixel p, the samples s in [0, ns) with
    s = (g - p) mod world (+ k*world) — so every rank traces floor or ceil(ns/world) samples of every pixel and the
    remainder ns % world is spread evenly over the pixels: per-rank work is even for ANY ns, in a single kernel launch.
    (For world == 1 this is the plain render.)"""
    if not (0 <= rank < world) or ns < 0 or npix < 0:
        raise ValueError("bad partition arguments")
    if ns == 0 or npix == 0:
        return []
    if world == 1:
        return [dict(sample_begin=0, sample_count=ns, sample_stride=1, rotate=False, accumulate=False)]
    return [dict(sample_begin=rank, sample_count=ns, sample_stride=world, rotate=True, accumulate=False)]


def samples_of(launch: dict, pixel: int) -> list[int]:
    """the sample indices a launch of partition_plan() renders for `pixel` (what the kernel computes per pixel)"""
    g = launch["sample_stride"]
    if not launch["rotate"]:
        return [launch["sample_begin"] + k * g for k in range(launch["sample_count"])]
    return list(range((launch["sample_begin"] - pixel) % g, launch["sample_count"], g))


def render_partitioned(render_fn, accum, ns: int, dist=None, dst: int = 0, npix: int | None = None):
    """render_fn(**launch) must render one launch of partition_plan() into `accum` (a torch tensor: nx*ny*3 float32
    sums, on the GPU for NCCL, on the CPU for gloo), overwriting unless launch['accumulate'].  Returns the reduced
    tensor (valid on dst).  A rank with nothing to do contributes zeros."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    if npix is None:
        npix = accum.numel() // 3
    plan = partition_plan(ns, npix, world, rank)
    if not plan:
        accum.zero_()
    for launch in plan:
        render_fn(**launch)
    if dist is not None and world > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
