"""Multi-GPU sharding for the two partitionable units of the hot path: independent samples and
independent rows of the Chamfer matrix.  One process per GPU; no data-path collective -- only the final gather
of generated points / matrix row blocks goes through torch.distributed (NCCL on the GPU box, gloo in CPU tests).
The reference has no distributed code at all (README.md:53), so there is nothing to be wire-compatible with.

Row assignment:
* a general [na, nb] matrix (M_rs): contiguous, balanced row blocks (``shard_range``);
* a set against itself (M_rr, M_ss): only the upper triangle is evaluated, so row i costs n - i pairs; rows are
  INTERLEAVED over ranks in a snake (in every block of 2*world rows rank r owns rows r and 2*world-1-r, whose costs add
  up to the same total for every rank), which balances the triangle to within one row per rank (``interleaved_rows``);
  the gathered triangle is mirrored.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def _world_rank(group=None) -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of `total` units for `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def interleaved_rows(total: int, world: int, rank: int) -> list:
    """Rows of a symmetric matrix's upper triangle owned by ``rank``: r and 2*world-1-r (mod 2*world), ascending."""
    step = 2 * world
    return sorted(list(range(rank, total, step)) + list(range(step - 1 - rank, total, step)))


def upper_pairs(total: int, world: int, rank: int) -> int:
    """Cloud pairs (j >= i) rank evaluates under the interleaved assignment."""
    return sum(total - i for i in interleaved_rows(total, world, rank))


def _all_gather(t: torch.Tensor, world: int, group=None) -> list:
    """all_gather of equally shaped tensors; gloo has no CUDA all_gather, so CUDA tensors are staged through the host
    there (CPU tests and the 2-process single-GPU test); NCCL gathers in place on the device."""
    if dist.get_backend(group) == "gloo" and t.is_cuda:
        host = t.cpu()
        bufs = [torch.empty_like(host) for _ in range(world)]
        dist.all_gather(bufs, host, group=group)
        return [b.to(t.device) for b in bufs]
    bufs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(bufs, t, group=group)
    return bufs


def _gather_blocks(local: torch.Tensor, counts: list, group=None) -> list:
    """All-gather per-rank row blocks of different heights (padded to the tallest); returns the trimmed blocks."""
    world = len(counts)
    max_rows = max(counts)
    pad = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = _all_gather(pad, world, group)
    return [bufs[r][: counts[r]] for r in range(world)]


def gather_rows(local: torch.Tensor, total_rows: int, group=None) -> torch.Tensor:
    """All-gather row blocks produced with shard_range back into [total_rows, ...] on every rank."""
    world, _ = _world_rank(group)
    if world == 1:
        return local
    counts = [e - b for b, e in (shard_range(total_rows, world, r) for r in range(world))]
    return torch.cat(_gather_blocks(local, counts, group), dim=0)


def sharded_pairwise_cd(a: torch.Tensor, b: torch.Tensor, group=None) -> torch.Tensor:
    """Chamfer matrix [na, nb] (evaluation/evaluation_metrics.py:165-198) with its rows split over ranks and gathered
    everywhere.  ``a`` and ``b`` the same set: upper triangle on interleaved rows, mirrored after the gather."""
    from . import metrics, ops
    world, rank = _world_rank(group)
    if metrics._same_set(a, b):
        x = a.contiguous().float()
        n = x.shape[0]
        if world == 1:
            return ops.mirror_upper(ops.pairwise_cd_upper(x))
        # this rank's two strided row sequences (j >= i filled, zeros elsewhere) into one [n, n] buffer
        upper = ops.pairwise_cd_upper(x, rank, 2 * world)
        ops.pairwise_cd_upper(x, 2 * world - 1 - rank, 2 * world, out=upper)
        owned = [interleaved_rows(n, world, r) for r in range(world)]
        mine = torch.tensor(owned[rank], dtype=torch.long, device=x.device)
        blocks = _gather_blocks(upper.index_select(0, mine), [len(o) for o in owned], group)
        upper = torch.empty_like(upper)
        for o, blk in zip(owned, blocks):
            upper[torch.tensor(o, dtype=torch.long, device=x.device)] = blk
        return ops.mirror_upper(upper)
    rows = shard_range(a.shape[0], world, rank)
    local = metrics._pairwise_CD_(a, b, rows=rows)
    return gather_rows(local, a.shape[0], group)


def sharded_compute_CD_metrics(sample_pcs: torch.Tensor, ref_pcs: torch.Tensor, group=None) -> dict:
    """compute_CD_metrics (evaluation_metrics.py:299-318) with the three matrices row-sharded over the ranks of ``group``;
    every rank returns the same MMD-CD / COV-CD / 1-NNA-CD dictionary."""
    from . import metrics
    results = {}
    ref_pcs, sample_pcs = ref_pcs.cuda(), sample_pcs.cuda()
    M_rs = sharded_pairwise_cd(ref_pcs, sample_pcs, group)
    metrics._update(results, metrics.lgan_mmd_cov(M_rs.t()), "CD")
    M_rr = sharded_pairwise_cd(ref_pcs, ref_pcs, group)
    M_ss = sharded_pairwise_cd(sample_pcs, sample_pcs, group)
    metrics._update(results, metrics.knn(M_rr, M_rs, M_ss, 1, sqrt=False), "CD", only_acc=True)
    return results
