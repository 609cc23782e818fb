"""N>1 host logic on CPU: 2-rank gloo process group (the GPU box uses NCCL with the same code path).

Covers the two shardings of SURVEY.md 8(e): the sample batch (independent samples, final all-gather of the generated
points) and the rows of the Chamfer matrix (row blocks gathered into the full matrix).  The per-rank compute is
stood in for by the C oracle (tests may use it as a checker); what is under test is shard_range / gather_rows and the
ragged-size handling, which is what runs unchanged on the GPU box.
"""
import ctypes as C
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ldt_b200.distributed import gather_rows, interleaved_rows, shard_range, upper_pairs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle():
    L = C.CDLL(os.path.join(ROOT, "oracle", "liboracle_nn.so"))
    L.oracle_pairwise_cd.argtypes = [C.c_int] * 4 + [C.c_void_p] * 2 + [C.c_int] * 2 + [C.c_void_p, C.c_int]
    return L


def _worker(rank, world, port, na, nb, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) Chamfer rows sharded over ranks, gathered into the full matrix
        g = torch.Generator().manual_seed(5)
        a = torch.randn((na, 64, 3), generator=g)
        b = torch.randn((nb, 48, 3), generator=g)
        r0, r1 = shard_range(na, world, rank)
        local = torch.empty((r1 - r0, nb))
        L = _oracle()
        if r1 > r0:
            L.oracle_pairwise_cd(na, nb, 64, 48, a.data_ptr(), b.data_ptr(), r0, r1, local.data_ptr(), 1)
        full = gather_rows(local, na)
        # (2) sample batch sharded over ranks: every rank "generates" its slice, all ranks end with all clouds in order
        total = 2 * na + 1
        s0, s1 = shard_range(total, world, rank)
        mine = torch.arange(s0, s1, dtype=torch.float32).view(-1, 1, 1).expand(-1, 4, 3).contiguous()
        clouds = gather_rows(mine, total)
        # (3) a set against itself: upper triangle on interleaved rows, gathered and mirrored.  The per-rank kernel call is
        # stood in for by the oracle (this is the host logic; tests/test_gpu_distributed.py runs the real kernel per rank)
        from ldt_b200 import distributed, ops

        def fake_upper(x, row_first=0, row_step=1, out=None):
            n = x.shape[0]
            u = torch.zeros((n, n)) if out is None else out
            for i in range(row_first, n, row_step):
                row = torch.empty((1, n))
                L.oracle_pairwise_cd(n, n, x.shape[1], x.shape[1], x.data_ptr(), x.data_ptr(), i, i + 1, row.data_ptr(), 1)
                u[i, i:] = row[0, i:]
            return u

        ops.pairwise_cd_upper = fake_upper
        sym = distributed.sharded_pairwise_cd(a, a)
        torch.save({"full": full, "clouds": clouds, "sym": sym}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_shard_range_is_a_balanced_partition():
    for total in (0, 1, 7, 8, 2048, 2049):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_interleaved_rows_balance_the_upper_triangle():
    for total in (1, 7, 8, 2048, 2049):
        for world in (1, 2, 3, 8):
            rows = [list(interleaved_rows(total, world, r)) for r in range(world)]
            assert sorted(i for rr in rows for i in rr) == list(range(total))
            pairs = [upper_pairs(total, world, r) for r in range(world)]
            assert sum(pairs) == total * (total + 1) // 2
            assert max(pairs) - min(pairs) <= total        # within one row of each other
    # BASELINE configs[3]: 2048 x 2048 over 8 ranks -> every rank evaluates exactly the same number of pairs
    pairs = [upper_pairs(2048, 8, r) for r in range(8)]
    assert max(pairs) == min(pairs)


@pytest.mark.parametrize("na,nb", [(5, 3), (8, 8), (1, 4)])
def test_two_rank_gloo_row_sharded_matrix_and_sample_gather(tmp_path, na, nb):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, na, nb, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(5)
    a = torch.randn((na, 64, 3), generator=g)
    b = torch.randn((nb, 48, 3), generator=g)
    ref = torch.empty((na, nb))
    _oracle().oracle_pairwise_cd(na, nb, 64, 48, a.data_ptr(), b.data_ptr(), 0, na, ref.data_ptr(), 1)
    ref_sym = torch.empty((na, na))
    _oracle().oracle_pairwise_cd(na, na, 64, 64, a.data_ptr(), a.data_ptr(), 0, na, ref_sym.data_ptr(), 1)
    assert torch.equal(ref_sym, ref_sym.t())
    total = 2 * na + 1
    want = torch.arange(total, dtype=torch.float32).view(-1, 1, 1).expand(-1, 4, 3)
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"))
        assert torch.equal(got["full"], ref)          # bit-identical to the unsharded matrix, on every rank
        assert torch.equal(got["clouds"], want)
        assert torch.equal(got["sym"], ref_sym)       # triangle on interleaved rows, gathered, mirrored == full matrix
