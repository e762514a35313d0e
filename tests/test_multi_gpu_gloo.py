"""world_size-2 test of the multi-GPU partition on CPU (gloo): each rank renders its sample share with the C oracle
(the CPU stand-in for the kernel in this test only), one reduce to rank 0, compared with the single-process render."""
import importlib
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ns, out_path):
    import torch
    import torch.distributed as dist
    for p in (str(ROOT), str(ROOT / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    rtnw = importlib.import_module("peter-shirley-ray-tracing-the-next-week_b200")
    mg = importlib.import_module("peter-shirley-ray-tracing-the-next-week_b200.multi_gpu")
    import oracle_port as op
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hs = rtnw.HostScene("cornell_box")
    nx = ny = 24
    cam = hs.camera(nx, ny)
    accum = torch.zeros(ny, nx, 3, dtype=torch.float32)

    def render(**launch):
        p = hs.params(nx=nx, ny=ny, ns=launch["sample_count"], seed=17, sample_begin=launch["sample_begin"],
                      sample_stride=launch["sample_stride"], flags_extra=rtnw.F_ROTATE_SAMPLES if launch["rotate"] else 0)
        op.render(rtnw, hs.desc_ptr, cam, p, out=accum.numpy())  # the C oracle stands in for the kernel; same parameters

    mg.render_partitioned(render, accum, ns, dist=dist, dst=0)
    if rank == 0:
        np.save(out_path, accum.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("ns", [6, 1, 5])
def test_two_ranks_partition_and_reduce(rtnw, tmp_path, ns):
    import torch.multiprocessing as mp
    import oracle_port as op
    out = tmp_path / "accum.npy"
    mp.spawn(_worker, args=(2, _free_port(), ns, str(out)), nprocs=2, join=True)
    got = np.load(out)
    hs = rtnw.HostScene("cornell_box")
    want, _ = op.render(rtnw, hs.desc_ptr, hs.camera(24, 24), hs.params(nx=24, ny=24, ns=ns, seed=17))
    assert np.allclose(got, want, rtol=1e-6, atol=1e-6)
    assert got.sum() > 0


def test_partition_plan_covers_every_path_exactly_once_and_evenly():
    mg = importlib.import_module("peter-shirley-ray-tracing-the-next-week_b200.multi_gpu")
    npix = 37
    for ns in (0, 1, 7, 100):
        for world in (1, 2, 3, 8):
            seen = {}
            work = []
            for r in range(world):
                w = 0
                for l in mg.partition_plan(ns, npix, world, r):
                    for p in range(npix):
                        for smp in mg.samples_of(l, p):
                            seen[(p, smp)] = seen.get((p, smp), 0) + 1
                            w += 1
                work.append(w)
            assert len(seen) == npix * ns and set(seen.values()) <= {1}
            assert max(work) - min(work) <= max(1, npix % world)  # even: at most a few paths apart
    with pytest.raises(ValueError):
        mg.partition_plan(4, 10, 2, 2)
