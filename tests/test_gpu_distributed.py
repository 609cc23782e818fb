"""Row-sharded Chamfer matrix (BASELINE configs[3]) on the GPU: ``distributed.sharded_pairwise_cd`` with
the REAL kernel on every rank.  Single rank in-process; two ranks as two processes -- NCCL when the box has two GPUs,
else gloo with both ranks on cuda:0 (NCCL refuses two ranks on one device; the kernel path is the same)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _clouds(n, pts, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((n, pts, 3), generator=g)
    x = x - x.mean(1, keepdim=True)
    return x / x.norm(dim=-1).amax(1)[:, None, None]


def test_sharded_pairwise_cd_single_rank():
    from ldt_b200 import ops
    from ldt_b200.distributed import sharded_compute_CD_metrics, sharded_pairwise_cd
    from ldt_b200.metrics import compute_CD_metrics
    dev = torch.device("cuda:0")
    a, b = _clouds(13, 2048, 7).to(dev), _clouds(9, 2048, 8).to(dev)
    assert torch.equal(sharded_pairwise_cd(a, b), ops.pairwise_cd(a, b))
    assert torch.equal(sharded_pairwise_cd(a, a), ops.pairwise_cd(a, a))
    r1, r2 = sharded_compute_CD_metrics(b, a), compute_CD_metrics(b, a)
    assert all(torch.equal(r1[k], r2[k]) for k in r2)


def _worker(rank, world, port, backend, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ldt_b200 import ops
        from ldt_b200.distributed import sharded_compute_CD_metrics, sharded_pairwise_cd
        n0 = ops.launch_count()
        a, b = _clouds(11, 1024, 7).to(dev), _clouds(6, 1024, 8).to(dev)
        M_ab = sharded_pairwise_cd(a, b)
        M_aa = sharded_pairwise_cd(a, a)
        res = sharded_compute_CD_metrics(b, a)
        torch.save({"M_ab": M_ab.cpu(), "M_aa": M_aa.cpu(), "res": {k: v.cpu() for k, v in res.items()},
                    "launches": ops.launch_count() - n0}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_sharded_pairwise_cd_two_ranks_run_the_real_kernel(tmp_path):
    from ldt_b200 import ops
    from ldt_b200.metrics import compute_CD_metrics
    world = 2
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    mp.spawn(_worker, args=(world, _free_port(), backend, str(tmp_path)), nprocs=world, join=True)
    dev = torch.device("cuda:0")
    a, b = _clouds(11, 1024, 7).to(dev), _clouds(6, 1024, 8).to(dev)
    want_ab, want_aa = ops.pairwise_cd(a, b).cpu(), ops.pairwise_cd(a, a).cpu()
    want = compute_CD_metrics(b, a)
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"))
        assert got["launches"] > 0                               # the CUDA kernel ran in this rank's process
        assert torch.equal(got["M_ab"], want_ab) and torch.equal(got["M_aa"], want_aa)
        assert all(torch.equal(got["res"][k], want[k].cpu()) for k in want)
