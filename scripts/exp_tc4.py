#!/usr/bin/env python
"""Cluster-of-4 multicast GEMM (backend 4) against the CTA-pair kernel (backend 3): exact-input correctness, then
sustained (graph replay, interleaved A B A B) throughput per shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from ldt_b200 import ops  # noqa: E402
from scripts.exp_gemm_limits import graph_of, timed_with_clocks  # noqa: E402

dev = torch.device("cuda:0")


def check():
    g = torch.Generator().manual_seed(3)
    for M, N, K in [(512, 256, 64), (512, 256, 512), (1024, 512, 1024), (768, 256, 256), (8192, 1024, 1024), (8192, 4096, 1024),
                    (2048 + 256, 1024, 4096)]:
        A = torch.randint(-4, 5, (M, K), generator=g).float().to(dev).bfloat16()
        W = torch.randint(-4, 5, (N, K), generator=g).float().to(dev).bfloat16()
        bias = torch.randint(-8, 9, (N,), generator=g).float().to(dev)
        ref = torch.empty((M, N), device=dev)
        ops.gemm(A, W, bias, ref, 0, backend=3)
        out = torch.full((M, N), 7.0, device=dev)
        ops.gemm(A, W, bias, out, 0, backend=4)
        torch.cuda.synchronize()
        ok = torch.equal(out, ref)
        print(f"backend 4 vs 3 at {M}x{N}x{K}: {'bit-identical' if ok else 'MISMATCH max ' + str((out - ref).abs().max().item())}", flush=True)
        if not ok:
            bad = (out != ref).nonzero()
            print("  first mismatches:", bad[:5].tolist(), " rows with mismatch:", bad[:, 0].unique()[:16].tolist(), flush=True)
            return False
    return True


def main():
    if not check():
        sys.exit(1)
    M = 8192
    g = torch.Generator().manual_seed(0)
    for name, N, K, epi in [("qkv", 3072, 1024, 1), ("fc_o", 1024, 1024, 3), ("fc1", 4096, 1024, 2), ("fc2", 1024, 4096, 3)]:
        A = (torch.randn((M, K), generator=g) * 0.5).to(dev).bfloat16()
        W = [(torch.randn((N, K), generator=g) / K ** 0.5).to(dev).bfloat16() for _ in range(24)]
        bias = torch.randn((N,), generator=g).to(dev)
        gate = torch.randn((1, N), generator=g).to(dev)
        out = torch.zeros((M, N), dtype=torch.float32 if epi == 3 else torch.bfloat16, device=dev)
        kw = dict(resid=out, gate=gate, gate_stride=0, rows_per_gate=32) if epi == 3 else {}
        graphs = {}
        for be in (3, 4):
            def body(be=be):
                for i in range(24):
                    ops.gemm(A, W[i], bias, out, epi, backend=be, **kw)
            graphs[be] = graph_of(body)
        tot = {3: 0.0, 4: 0.0}
        for r in range(2):
            for be in (3, 4):
                ms, clk, pw = timed_with_clocks(graphs[be].replay, 1.0)
                tot[be] += ms / 2
        fl = 2.0 * M * N * K
        for be in (3, 4):
            us = tot[be] / 24 * 1e3
            print(f"{name:5s} backend {be}: {us:8.2f} us {fl / us / 1e6:8.1f} TFLOP/s (sustained, interleaved)", flush=True)


if __name__ == "__main__":
    main()
