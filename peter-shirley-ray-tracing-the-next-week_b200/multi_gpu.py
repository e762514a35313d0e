"""Multi-GPU partition of the sample loop: one process per GPU, rank g of G renders samples
g, g+G, g+2G, ... of EVERY pixel into a private float accumulation buffer, then ONE reduce(sum) to rank 0 before the
host epilogue (mean, gamma, quantise, PPM — PSC/main.cpp:315-330).  The sample stream is keyed by (seed, pixel, sample),
so the set of paths is the same for any G; results differ only by float summation order.  There is no other exchange
step on this path, hence no other collective.
"""
from __future__ import annotations


def sample_partition(ns: int, world: int, rank: int) -> tuple[int, int, int]:
    """(sample_begin, sample_count, sample_stride) of `rank`: samples rank, rank+world, ... below ns."""
    if not (0 <= rank < world) or ns < 0:
        raise ValueError("bad partition arguments")
    return rank, len(range(rank, ns, world)), world


def partition_plan(ns: int, npix: int, world: int, rank: int) -> list[dict]:
    """The launches of `rank` so that every rank does ns/world samples' worth of work for ANY ns:
      1. the samples that divide evenly, split by sample index: rank, rank+world, ... below world*(ns//world);
      2. the ns % world left-over samples, split by interleaved PIXELS: pixel p is done by rank p % world, added into the
         same buffer (RTNW_F_ACCUMULATE).
    Each dict holds keyword overrides for the render parameters; `accumulate` is True for launches after the first."""
    if not (0 <= rank < world) or ns < 0 or npix < 0:
        raise ValueError("bad partition arguments")
    base, rem = divmod(ns, world)
    plan = []
    if base > 0:
        plan.append(dict(sample_begin=rank, sample_count=base, sample_stride=world, pixel_begin=0, pixel_stride=1, pixel_count=0,
                         accumulate=False))
    if rem > 0:
        count = len(range(rank, npix, world))
        if count > 0:
            plan.append(dict(sample_begin=base * world, sample_count=rem, sample_stride=1, pixel_begin=rank, pixel_stride=world,
                             pixel_count=count, accumulate=bool(plan)))
    return plan


def render_partitioned(render_fn, accum, ns: int, dist=None, dst: int = 0, npix: int | None = None):
    """render_fn(**launch) must render one launch of partition_plan() into `accum` (a torch tensor: nx*ny*3 float32
    sums, on the GPU for NCCL, on the CPU for gloo), overwriting unless launch['accumulate'].  Returns the reduced
    tensor (valid on dst).  A rank with nothing to do contributes zeros."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    if npix is None:
        npix = accum.numel() // 3
    plan = partition_plan(ns, npix, world, rank)
    if not plan or plan[0]["pixel_count"] != 0:
        accum.zero_()  # nothing rendered, or only a pixel subset: the rest of the buffer must be zero
        for launch in plan:
            launch["accumulate"] = True
    for launch in plan:
        render_fn(**launch)
    if dist is not None and world > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
