"""World-size-2 test of the multi-GPU plumbing on CPU (gloo): slab partition + the single table broadcast.
The data path itself has no collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fimex_b200.slab import broadcast_tables, slab_range


def test_slab_range_partitions_the_stack():
    for n, world in ((3288, 1), (3288, 2), (3288, 8), (10, 4), (3, 8), (0, 2)):
        covered = []
        for r in range(world):
            b, e = slab_range(n, r, world)
            assert 0 <= b <= e <= n
            covered.extend(range(b, e))
        assert covered == list(range(n))
        sizes = [slab_range(n, r, world)[1] - slab_range(n, r, world)[0] for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    assert slab_range(3288, 7, 8) == (2877, 3288)  # 411 levels per GPU at G = 8
    with pytest.raises(ValueError):
        slab_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if rank == 0:
            rng = np.random.default_rng(1)
            px = torch.from_numpy(rng.uniform(0, 100, n))
            py = torch.from_numpy(rng.uniform(0, 50, n))
            geom = [331, 207, 512, 77]
        else:
            px = torch.empty(n, dtype=torch.float64)
            py = torch.empty(n, dtype=torch.float64)
            geom = [0, 0, 0, 0]
        px, py, geom = broadcast_tables(px, py, geom, src=0)
        b, e = slab_range(3288, rank, world)
        q.put((rank, float(px.sum()), float(py.sum()), geom, (b, e)))
    finally:
        dist.destroy_process_group()


def test_table_broadcast_world_size_2():
    world, n = 2, 4096
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] and res[0][2] == res[1][2]  # identical tables on both ranks, bit for bit
    assert res[0][3] == res[1][3] == [331, 207, 512, 77]
    assert res[0][4] == (0, 1644) and res[1][4] == (1644, 3288)
