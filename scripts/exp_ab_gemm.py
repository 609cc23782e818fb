#!/usr/bin/env python
"""Sustained interleaved A/B of one GEMM shape under two ldt_debug_set_gemm_mode values (set at graph-capture time).
usage: python scripts/exp_ab_gemm.py modeA,modeB [shape ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from ldt_b200 import _lib, ops  # noqa: E402
from scripts.exp_gemm_limits import graph_of, timed_with_clocks  # noqa: E402

dev = torch.device("cuda:0")
M = 8192
SH = {"qkv": (3072, 1024, 1), "fc_o": (1024, 1024, 3), "fc1": (4096, 1024, 2), "fc2": (1024, 4096, 3)}


def main():
    modes = [int(m) for m in sys.argv[1].split(",")]
    names = sys.argv[2:] or ["fc1", "qkv"]
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    for name in names:
        N, K, epi = SH[name]
        A = (torch.randn((M, K), generator=g) * 0.5).to(dev).bfloat16()
        W = [(torch.randn((N, K), generator=g) / K ** 0.5).to(dev).bfloat16() for _ in range(24)]
        bias = torch.randn((N,), generator=g).to(dev)
        gate = torch.randn((1, N), generator=g).to(dev)
        out = torch.zeros((M, N), dtype=torch.float32 if epi == 3 else torch.bfloat16, device=dev)
        kw = dict(resid=out, gate=gate, gate_stride=0, rows_per_gate=32) if epi == 3 else {}
        graphs, outs = {}, {}
        for m in modes:
            lib.ldt_debug_set_gemm_mode(m)

            def body():
                for i in range(24):
                    ops.gemm(A, W[i], bias, out, epi, backend=3, **kw)
            if epi == 3:
                out.copy_(torch.arange(M * N, device=dev, dtype=torch.float32).view(M, N) * 1e-6)
                ops.gemm(A, W[0], bias, out, epi, backend=3, **kw)
            else:
                ops.gemm(A, W[0], bias, out, epi, backend=3)
            outs[m] = out.clone()
            graphs[m] = graph_of(body)
        lib.ldt_debug_set_gemm_mode(0)
        if len(outs) > 1:
            ks = list(outs)
            print(f"{name}: outputs of mode {ks[0]} and {ks[1]} identical: {torch.equal(outs[ks[0]], outs[ks[1]])}", flush=True)
        tot = {m: 0.0 for m in modes}
        rounds = 3
        for r in range(rounds):
            for m in modes:
                ms, clk, pw = timed_with_clocks(graphs[m].replay, 0.8)
                tot[m] += ms / rounds
        fl = 2.0 * M * N * K
        for m in modes:
            us = tot[m] / 24 * 1e3
            print(f"{name:5s} mode {m:4d}: {us:8.2f} us {fl / us / 1e6:8.1f} TFLOP/s (sustained, interleaved x{rounds})", flush=True)


if __name__ == "__main__":
    main()
