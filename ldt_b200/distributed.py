"""Multi-GPU sharding for the two partitionable units of the hot path: independent samples and
independent rows of the Chamfer matrix.  One process per GPU; no data-path collective -- only the final gather
of generated points / matrix row blocks goes through torch.distributed (NCCL on the GPU box, gloo in CPU tests).
The reference has no distributed code at all (README.md:53), so there is nothing to be wire-compatible with.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of `total` units for `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_rows(local: torch.Tensor, total_rows: int, group=None) -> torch.Tensor:
    """All-gather row blocks produced with shard_range back into [total_rows, ...] on every rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(total_rows, world, r) for r in range(world)]
    max_rows = max(e - b for b, e in sizes)
    pad = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: e - b] for r, (b, e) in enumerate(sizes)], dim=0)


def sharded_pairwise_cd(a: torch.Tensor, b: torch.Tensor, group=None) -> torch.Tensor:
    """Chamfer matrix [na, nb] with rows split over ranks, gathered everywhere."""
    from . import metrics
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    rows = shard_range(a.shape[0], world, rank)
    local = metrics._pairwise_CD_(a, b, rows=rows)
    return gather_rows(local, a.shape[0], group)
