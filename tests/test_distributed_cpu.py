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

from ldt_b200.distributed import gather_rows, shard_range

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
        torch.save({"full": full, "clouds": clouds}, os.path.join(out_dir, f"rank{rank}.pt"))
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
    total = 2 * na + 1
    want = torch.arange(total, dtype=torch.float32).view(-1, 1, 1).expand(-1, 4, 3)
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"))
        assert torch.equal(got["full"], ref)          # bit-identical to the unsharded matrix, on every rank
        assert torch.equal(got["clouds"], want)
