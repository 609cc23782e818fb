#!/usr/bin/env python
"""Sustained (power-capped) throughput of the CTA-pair GEMM variants: each configuration is 24 launches captured in a
CUDA graph and replayed for ~1.2 s while NVML samples SM clock and board power.  At the 1 kW cap every kernel of the step
draws the same power, so time per launch is proportional to ENERGY per launch: this is the measurement that decides
what to optimise (isolated 50-launch timings run at boost clocks and reward idle-time removal that the cap takes back).

usage: python scripts/exp_sustained.py [shape,...]     shapes: qkv fc_o fc1 fc2
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from ldt_b200 import _lib, ops  # noqa: E402
from scripts.exp_gemm_limits import graph_of, timed_with_clocks  # noqa: E402

dev = torch.device("cuda:0")
M = 8192
SHAPES = {"qkv": (3072, 1024), "fc_o": (1024, 1024), "fc1": (4096, 1024), "fc2": (1024, 4096)}


def main():
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else list(SHAPES)
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    for name in names:
        N, K = SHAPES[name]
        A = (torch.randn((M, K), generator=g) * 0.5).to(dev).bfloat16()
        W = [(torch.randn((N, K), generator=g) / K ** 0.5).to(dev).bfloat16() for _ in range(24)]
        bias = torch.randn((N,), generator=g).to(dev)
        gate = torch.randn((1, N), generator=g).to(dev)
        o32 = torch.zeros((M, N), dtype=torch.float32, device=dev)
        o16 = torch.zeros((M, N), dtype=torch.bfloat16, device=dev)
        fl = 2.0 * M * N * K

        def run(tag, fn, mode=0):
            lib.ldt_debug_set_gemm_mode(mode)

            def body():
                for i in range(24):
                    fn(i)
            gr = graph_of(body)
            ms, clk, pw = timed_with_clocks(gr.replay, 1.2)
            lib.ldt_debug_set_gemm_mode(0)
            us = ms / 24 * 1e3
            print(f"{name:5s} {tag:38s}: {us:8.2f} us {fl / us / 1e6:8.1f} TFLOP/s  [SM {clk} MHz, {pw:.0f} W]  "
                  f"{fl / us / 1e6 / max(pw, 1):.3f} TFLOP/J", flush=True)

        run("cuBLAS bf16 out", lambda i: torch.matmul(A, W[i].t(), out=o16))
        run("ours bias -> bf16", lambda i: ops.gemm(A, W[i], bias, o16, 1, backend=3))
        run("ours bias+GELU -> bf16", lambda i: ops.gemm(A, W[i], bias, o16, 2, backend=3))
        run("ours bias -> f32", lambda i: ops.gemm(A, W[i], bias, o32, 0, backend=3))
        run("ours gate*acc+resid -> f32", lambda i: ops.gemm(A, W[i], bias, o32, 3, resid=o32, gate=gate, gate_stride=0,
                                                             rows_per_gate=32, backend=3))
        run("ours no epilogue (mode 4)", lambda i: ops.gemm(A, W[i], bias, o16, 1, backend=3), 4)
        run("ours no A loads (mode 1), bf16 out", lambda i: ops.gemm(A, W[i], bias, o16, 1, backend=3), 1)
        run("ours no A loads, no epilogue (mode 5)", lambda i: ops.gemm(A, W[i], bias, o16, 1, backend=3), 5)
        run("single-CTA tiles bias -> bf16", lambda i: ops.gemm(A, W[i], bias, o16, 1, backend=2))


if __name__ == "__main__":
    main()
