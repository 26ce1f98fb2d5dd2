"""Multi-rank host logic on CPU (gloo, world size 2): the sample split covers every sample
exactly once for any world size, and the one collective of the path (SUM-reduce of the
accumulators) reproduces the single-rank image."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def test_rank_split_partitions_samples():
    d = importlib.import_module(PKG + ".distributed")
    for spp in (0, 1, 7, 64, 400, 1024):
        for world in (1, 2, 3, 4, 8):
            seen = np.concatenate([d.local_sample_indices(spp, r, world) for r in range(world)])
            assert sorted(seen.tolist()) == list(range(spp))
            sizes = [d.rank_split(spp, r, world)[2] for r in range(world)]
            assert sum(sizes) == spp and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        d.rank_split(4, 2, 2)


def fake_accumulators(samples, w, h):
    """Stand-in for a rank's render: a deterministic function of (pixel, sample) summed over
    the rank's samples — exactly the structure of the real accumulators."""
    pix = np.arange(w * h, dtype=np.float64).reshape(h, w, 1)
    out = np.zeros((h, w, 4), np.float32)
    for s in samples:
        out[..., :3] += np.sin(0.37 * pix + 1.3 * s + np.arange(3)).astype(np.float32) ** 2
    return out


def _worker(rank, world, port, spp, w, h, q):
    import sys
    sys.path.insert(0, ROOT)
    d = importlib.import_module(PKG + ".distributed")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    acc = torch.from_numpy(fake_accumulators(d.local_sample_indices(spp, rank, world), w, h))
    d.reduce_sum(acc, dst=0)
    if rank == 0:
        q.put(acc.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduce_matches_single_rank():
    spp, w, h, world = 13, 16, 8, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, spp, w, h, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = fake_accumulators(range(spp), w, h)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-5)


def test_plan_split_covers_every_sample_of_every_pixel_once():
    d = importlib.import_module(PKG + ".distributed")
    for spp, height, world in ((400, 600, 8), (1, 64, 8), (2, 10, 8), (3, 7, 6), (16, 5, 4), (5, 9, 1)):
        seen = np.zeros((height, spp), int)
        for rank in range(world):
            pl = d.plan_split(spp, height, rank, world)
            assert pl["sample_stride"] * pl["row_stride"] == world
            rows = d.local_rows(height, pl["row_offset"], pl["row_stride"])
            smp = np.arange(pl["sample_offset"], spp, pl["sample_stride"])
            seen[np.ix_(rows, smp)] += 1
        assert (seen == 1).all(), (spp, height, world)
    pl = d.plan_split(400, 600, 3, 8)
    assert pl == {"sample_offset": 3, "sample_stride": 8, "row_offset": 0, "row_stride": 1}   # enough samples: no row split
    with pytest.raises(ValueError):
        d.plan_split(1, 2, 0, 8)
