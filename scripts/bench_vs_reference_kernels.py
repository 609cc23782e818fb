#!/usr/bin/env python
"""Side-by-side timing, on the same B200, of the evaluation kernels against the REFERENCE's own CUDA kernels compiled
from its sources (oracle/_ref, built by `make -C oracle`): how `_pairwise_CD_` / `_pairwise_EMD_CD_`
(evaluation/evaluation_metrics.py:112-198) would run if the reference extension were simply recompiled for sm_100a.

The reference computes one ROW of the pairwise matrix per call: the row cloud expanded to a batch of `bs` copies
against `bs` column clouds (:126-146).  We time exactly that call pattern (kernel launches only, no expand / mean /
cat, which favours the reference) and our one-launch-per-matrix kernels, on `n` x `n` clouds of 2048 points.
Prints one JSON line.  This is a measurement tool, not product code.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ldt_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clouds", type=int, default=64)
    ap.add_argument("--emd-clouds", type=int, default=32)
    ap.add_argument("--points", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=64, help="column batch of the reference loop (compute_all_metrics: 64)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    P = a.points
    nnd = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_nnd.so"))
    emd = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_emd.so"))
    f_nn = getattr(nnd, "_Z10nndistanceiiPKfiS0_PfPiS1_S2_P11CUstream_st")
    f_nn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_void_p] * 5
    f_am = getattr(emd, "_Z11approxmatchiiiPKfS0_PfS1_P11CUstream_st")
    f_mc = getattr(emd, "_Z9matchcostiiiPKfS0_PfS1_P11CUstream_st")
    f_am.argtypes = [C.c_int] * 3 + [C.c_void_p] * 5
    f_mc.argtypes = [C.c_int] * 3 + [C.c_void_p] * 5
    g = torch.Generator().manual_seed(7)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, reps=2):
        fn()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(reps):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / 1e3 / reps

    out = {}
    # ---- Chamfer ----
    n, bs = a.clouds, min(a.batch, a.clouds)
    A = torch.rand((n, P, 3), generator=g).to(dev)
    B = torch.rand((n, P, 3), generator=g).to(dev)
    d1, d2 = torch.empty((bs, P), device=dev), torch.empty((bs, P), device=dev)
    i1 = torch.empty((bs, P), dtype=torch.int32, device=dev)
    i2 = torch.empty((bs, P), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def ref_cd():
        for i in range(n):
            row = A[i:i + 1].expand(bs, -1, -1).contiguous()
            for j0 in range(0, n, bs):
                f_nn(bs, P, row.data_ptr(), P, B[j0:j0 + bs].data_ptr(), d1.data_ptr(), i1.data_ptr(), d2.data_ptr(), i2.data_ptr(), st)

    t_ref = timed(ref_cd)
    t_our = timed(lambda: ops.pairwise_cd(A, B), reps=5)
    out["cd"] = {"matrix": f"{n}x{n}", "reference_kernel_pairs_per_s": n * n / t_ref, "ldt_b200_pairs_per_s": n * n / t_our,
                 "speedup": t_ref / t_our}
    # ---- approximate EMD ----
    n = a.emd_clouds
    bs = min(a.batch, n)
    A, B = A[:n].contiguous(), B[:n].contiguous()
    match = torch.zeros((bs, P, P), device=dev)
    temp = torch.zeros((bs, 4 * P), device=dev)
    cost = torch.zeros((bs,), device=dev)

    def ref_emd():
        for i in range(n):
            row = A[i:i + 1].expand(bs, -1, -1).contiguous()
            for j0 in range(0, n, bs):
                f_am(bs, P, P, row.data_ptr(), B[j0:j0 + bs].data_ptr(), match.data_ptr(), temp.data_ptr(), None)
                f_mc(bs, P, P, row.data_ptr(), B[j0:j0 + bs].data_ptr(), match.data_ptr(), cost.data_ptr(), None)

    t_ref = timed(ref_emd, reps=1)
    t_our = timed(lambda: ops.pairwise_emd(A, B), reps=2)
    out["emd"] = {"matrix": f"{n}x{n}", "reference_kernel_pairs_per_s": n * n / t_ref, "ldt_b200_pairs_per_s": n * n / t_our,
                  "speedup": t_ref / t_our}
    out["note"] = ("reference = its nndistance.cu / approxmatch.cu compiled unchanged for sm_100a, driven with the call pattern of "
                   "evaluation_metrics.py:112-198 (one matrix row per call, column batch %d); kernels only" % a.batch)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
