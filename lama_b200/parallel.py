"""Batch sharding across the GPUs of one box.

Inference has no cross-sample coupling (eval-mode BN, per-plane FFTs), so the path shards by image:
the global batch is cut into contiguous per-rank shards, weights are replicated, and there is no
collective inside the model.  These helpers cover the only exchange BASELINE config 4 has — rank 0
holds the whole batch, scatters the (B,4,S,S) inputs and gathers the (B,3,S,S) outputs — on top of
``torch.distributed`` (NCCL over NVLink on GPUs; gloo in the CPU tests).  bench.py uses the
weak-scaling form instead (every rank generates its own shard), which needs no exchange at all.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [start, end) of every rank; the first ``batch % world`` ranks get one extra image."""
    base, extra = divmod(batch, world)
    out, s = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((s, s + n))
        s += n
    return out


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    s, e = shard_bounds(x.shape[0], world)[rank]
    return x[s:e]


def scatter_batch(x: Optional[torch.Tensor], shape_tail: Tuple[int, ...], batch: int, *, src: int = 0,
                  device=None, dtype=torch.float32, group=None) -> torch.Tensor:
    """Rank ``src`` passes the full ``(batch, *shape_tail)`` tensor, the others ``None``; every rank
    returns its contiguous shard.  Uneven shards are handled by padding to the largest shard."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(batch, world)
    big = max(e - s for s, e in bounds)
    recv = torch.empty((big,) + tuple(shape_tail), dtype=dtype, device=device)
    parts = None
    if rank == src:
        assert x is not None and tuple(x.shape) == (batch,) + tuple(shape_tail)
        parts = []
        for s, e in bounds:
            p = torch.zeros_like(recv)
            p[: e - s].copy_(x[s:e])
            parts.append(p)
    dist.scatter(recv, parts, src=src, group=group)
    s, e = bounds[rank]
    return recv[: e - s]


def gather_batch(y: torch.Tensor, batch: int, *, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Inverse of :func:`scatter_batch`: rank ``dst`` returns the ``(batch, ...)`` tensor, others ``None``."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(batch, world)
    big = max(e - s for s, e in bounds)
    send = torch.zeros((big,) + tuple(y.shape[1:]), dtype=y.dtype, device=y.device)
    send[: y.shape[0]].copy_(y)
    parts = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, parts, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([p[: e - s] for p, (s, e) in zip(parts, bounds)], dim=0)


def sharded_apply(fn, x_full: Optional[torch.Tensor], shape_tail, batch: int, *, device=None, group=None):
    """scatter -> ``fn`` on the local shard -> gather (rank 0 returns the full result)."""
    local = scatter_batch(x_full, shape_tail, batch, device=device, group=group)
    out = fn(local) if local.shape[0] > 0 else local.new_zeros((0,) + tuple(fn(local.new_zeros((1,) + tuple(shape_tail))).shape[1:]))
    return gather_batch(out, batch, group=group)
