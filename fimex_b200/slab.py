"""Multi-GPU plumbing: one process per GPU, contiguous slabs of the (time x level x ensemble) stack, ONE broadcast of
the cached tables.

The path shards naturally: every [y][x] level is independent given the index tables
(/root/reference/src/CachedInterpolation.cc:133-141 has no cross-level state; the reference's own MPI mode is a
rank-modulo split of the time axis, src/NetCDF_CDMWriter.cc:632-645).  So there is no data-path collective: rank 0
computes the tables on its GPU, the two fp64 position tables (2 x 8 B x outX*outY; 64 MB at 2000^2) go out once with
torch.distributed.broadcast (NCCL over NVLink on GPUs, gloo in the CPU tests), and every rank compiles its own
gather table from them on its own GPU.
"""
from __future__ import annotations

from typing import List, Optional, Tuple


def slab_range(n_levels: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slab [begin, end) of the flattened level stack owned by `rank` (sizes differ by at most one)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_levels, world)
    begin = rank * base + min(rank, rem)
    end = begin + base + (1 if rank < rem else 0)
    return begin, end


def broadcast_tables(px, py, geom: List[int], src: int = 0):
    """Broadcast (px, py) in place and the small integer geometry list; works on CPU (gloo) and CUDA (NCCL) tensors."""
    import torch
    import torch.distributed as dist

    g = torch.tensor(geom, dtype=torch.int64, device=px.device)
    dist.broadcast(g, src=src)
    dist.broadcast(px, src=src)
    dist.broadcast(py, src=src)
    return px, py, [int(v) for v in g.tolist()]


def broadcast_cached_interpolation(ci, geom: Optional[List[int]], method: int, outX: int, outY: int, rank: int, device):
    """Rank 0 holds a CachedInterpolation (tables on its GPU); afterwards every rank holds an equivalent one on its own
    GPU.  `geom` = [inX, inY, xMin, yMin] after createReducedDomain.  Returns (ci, geom)."""
    import torch
    import torch.distributed as dist

    from .cached import CachedInterpolation

    n = outX * outY
    if rank == 0:
        px, py = ci.device_points()
        assert px.numel() == n
        geom_t = list(geom)
    else:
        px = torch.empty(n, dtype=torch.float64, device=device)
        py = torch.empty(n, dtype=torch.float64, device=device)
        geom_t = [0, 0, 0, 0]
    px, py, geom_t = broadcast_tables(px, py, geom_t, src=0)
    if rank != 0:
        ci = CachedInterpolation("x", "y", method, px, py, geom_t[0], geom_t[1], outX, outY)
    dist.barrier()
    return ci, geom_t
