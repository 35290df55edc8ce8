"""CPU tests of the N>1 path: partitioning and the gather of the per-point totals over gloo (world 2 and 3)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from transport_b200 import shard


def test_split_and_plan_cover_everything_once():
    for n, w in ((1000, 8), (10, 3), (7, 7), (5, 8)):
        seen = np.zeros(n, dtype=int)
        for r in range(w):
            a, b = shard.split_range(n, w, r)
            seen[a:b] += 1
        assert (seen == 1).all()
    # fewer parameters than ranks -> the energy axis is split, parameters replicated
    (p, e) = shard.plan(1, 100_000, 8, 3)
    assert p == (0, 1) and e == shard.split_range(100_000, 8, 3)
    (p, e) = shard.plan(1000, 1000, 8, 7)
    assert p == (875, 1000) and e == (0, 1000)
    with pytest.raises(ValueError):
        shard.split_range(10, 2, 2)


def _worker(rank, world, port, n_params, n_energies, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full_ref = torch.arange(n_params * n_energies, dtype=torch.float64).reshape(n_params, n_energies)
    (p0, p1), (e0, e1) = shard.plan(n_params, n_energies, world, rank)
    local = full_ref[p0:p1, e0:e1].clone()
    got = shard.gather_totals(local, world, rank, n_params, n_energies)
    ret[rank] = bool(torch.equal(got, full_ref))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_params,n_energies", [(2, 10, 7), (3, 2, 11), (2, 5, 4)])
def test_gather_totals_gloo(world, n_params, n_energies):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_params, n_energies, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret[r] for r in range(world))
