"""The (pixel, sample range) work items of rtnw_render: rtnw_plan_sample_ranges is host arithmetic, so the
partition the kernel walks is checked here on the CPU by replaying the kernel's integer formulas
(k_render: k in [cum[c]*m/total, cum[c+1]*m/total), with m from RTNW_F_ROTATE_SAMPLES where set)."""
import importlib
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rtnw = importlib.import_module("peter-shirley-ray-tracing-the-next-week_b200")


def _params(nx, ny, ns, begin=0, stride=1, flags=0, ranges=0):
    return rtnw.RenderParams(nx, ny, begin, ns, stride, 50, 0.001, rtnw.FLT_MAX, 0, flags, 1, 0, 1, 0, ranges)


def _pixel_sample_count(p, pix):
    if p.flags & rtnw.F_ROTATE_SAMPLES:  # k_render: ownership of the samples rotates with the pixel index
        g = p.sample_stride
        b = ((p.sample_begin - pix) % g + g) % g
        return (p.sample_count - b + g - 1) // g if b < p.sample_count else 0
    return p.sample_count


@pytest.mark.parametrize("ns", [1, 2, 3, 4, 5, 7, 8, 11, 12, 13, 16, 37, 100, 101, 256, 1000, 5000])
def test_ranges_partition_every_pixels_samples(ns):
    p = _params(200, 100, ns)
    cum = rtnw.plan_sample_ranges(p)
    n = len(cum) - 1
    assert 1 <= n <= 64 and cum[0] == 0 and cum[-1] == ns and np.all(np.diff(cum) >= 1)
    sizes = np.diff(cum)
    assert sizes[-1] == 1 or ns < 2          # the kernel's tail is one short item
    assert sizes.max() <= max(4, -(-ns // 28)) + 7
    # replay: the ranges of a pixel with m samples are disjoint, in order, and cover [0, m)
    for m in {ns, max(ns - 1, 0)}:
        edges = [int(c) * m // ns for c in cum]
        assert edges[0] == 0 and edges[-1] == m and all(a <= b for a, b in zip(edges, edges[1:]))


@pytest.mark.parametrize("ns,world", [(100, 8), (12, 8), (7, 3), (3, 8), (64, 2)])
def test_rotated_split_ranges_cover_each_ranks_samples(ns, world):
    """multi-GPU: rank g owns, for pixel p, the samples s = (g - p) mod G + k*G below ns; its ranges partition them"""
    seen = {}
    for g in range(world):
        p = _params(40, 30, ns, begin=g, stride=world, flags=rtnw.F_ROTATE_SAMPLES)
        cum = rtnw.plan_sample_ranges(p)
        total = int(cum[-1])
        assert total == -(-ns // world)
        for pix in range(0, 40 * 30, 7):
            m = _pixel_sample_count(p, pix)
            b = ((g - pix) % world + world) % world
            ks = []
            for c in range(len(cum) - 1):
                ks += list(range(int(cum[c]) * m // total, int(cum[c + 1]) * m // total))
            assert ks == list(range(m))
            for k in ks:
                s = b + k * world
                assert 0 <= s < ns and (pix, s) not in seen
                seen[(pix, s)] = g
    assert len(seen) == len(range(0, 40 * 30, 7)) * ns  # every (pixel, sample) exactly once over the ranks


def test_forced_count_and_errors():
    """rtnw_render_params.sample_ranges forces N equal ranges; the schedule no longer depends on the image size (one
    fixed-point plane whatever the number of ranges)"""
    cum = rtnw.plan_sample_ranges(_params(64, 48, 37, ranges=5))
    assert len(cum) == 6 and cum[-1] == 37 and np.diff(cum).min() >= 7
    assert list(rtnw.plan_sample_ranges(_params(64, 48, 37, ranges=1))) == [0, 37]
    assert len(rtnw.plan_sample_ranges(_params(64, 48, 37, ranges=1000))) - 1 == 37  # capped by the sample count (and 64)
    big, small = rtnw.plan_sample_ranges(_params(8192, 8192, 100)), rtnw.plan_sample_ranges(_params(64, 48, 100))
    assert np.array_equal(big, small) and big[-1] == 100
    with pytest.raises(rtnw.RtnwError):
        rtnw.plan_sample_ranges(_params(64, 48, 0))
    with pytest.raises(rtnw.RtnwError):
        rtnw.plan_sample_ranges(_params(64, 48, 8, ranges=-1))
