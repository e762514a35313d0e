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


def render_partitioned(render_fn, accum, ns: int, dist=None, dst: int = 0):
    """render_fn(sample_begin, sample_count, sample_stride) must fill `accum` (a torch tensor: nx*ny*3 float32 sums,
    on the GPU for NCCL, on the CPU for gloo) with this rank's samples; returns the reduced tensor (valid on dst).
    A rank whose share is empty (world > ns) contributes zeros."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    begin, count, stride = sample_partition(ns, world, rank)
    if count > 0:
        render_fn(begin, count, stride)
    else:
        accum.zero_()
    if dist is not None and world > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
