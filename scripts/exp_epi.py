#!/usr/bin/env python
"""fc1-shaped GEMM with the bias+GELU->bf16 epilogue taken apart at COMPILE time (diagnostics build of the CTA-pair kernel,
ldt_debug_set_gemm_mode): sustained time per launch and the MMA thread's cycle counters for each variant."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from ldt_b200 import _lib, ops  # noqa: E402
from scripts.exp_gemm_limits import graph_of, timed_with_clocks  # noqa: E402

dev = torch.device("cuda:0")
M, N, K = 8192, 4096, 1024
MODES = [(0, "full epilogue"), (4, "no epilogue at all"), (64, "no GELU math"), (32, "no global stores"), (16, "no smem staging"),
         (8, "no TMEM loads"), (48, "TMEM + math only"), (56, "math only"), (112, "bias/cvt only (no TMEM/stg/st/GELU)"),
         (96, "no GELU, no global stores")]


def main():
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    A = (torch.randn((M, K), generator=g) * 0.5).to(dev).bfloat16()
    W = [(torch.randn((N, K), generator=g) / K ** 0.5).to(dev).bfloat16() for _ in range(24)]
    bias = torch.randn((N,), generator=g).to(dev)
    out = torch.zeros((M, N), dtype=torch.bfloat16, device=dev)
    buf = torch.zeros((148 * 8,), dtype=torch.int64, device=dev)
    lib.ldt_debug_set_gemm_counters(buf.data_ptr())   # selects the diagnostics build for every variant, mode 0 included
    graphs = {}
    for mode, _ in MODES:
        lib.ldt_debug_set_gemm_mode(mode)

        def body():
            for i in range(24):
                ops.gemm(A, W[i], bias, out, 2, backend=3)
        graphs[mode] = graph_of(body)
    tot = {m: 0.0 for m, _ in MODES}
    rounds = 2
    for r in range(rounds):
        for mode, _ in MODES:
            ms, clk, pw = timed_with_clocks(graphs[mode].replay, 0.8)
            tot[mode] += ms / rounds
    fl = 2.0 * M * N * K
    for mode, tag in MODES:
        lib.ldt_debug_set_gemm_mode(mode)
        ops.gemm(A, W[0], bias, out, 2, backend=3)
        torch.cuda.synchronize()
        c = buf.view(148, 8).cpu().double()
        lead = c[0::2]
        us = tot[mode] / 24 * 1e3
        print(f"{tag:38s}: {us:7.2f} us {fl / us / 1e6:7.1f} TFLOP/s sustained | MMA warp {lead[:, 0].mean():6.0f} clk, wait TMA "
              f"{lead[:, 1].mean():6.0f}, wait acc {lead[:, 2].mean():6.0f} | epilogue warp total {c[:, 3].mean():6.0f}, its wait {c[:, 4].mean():6.0f}",
              flush=True)
    lib.ldt_debug_set_gemm_mode(0)
    lib.ldt_debug_set_gemm_counters(None)


if __name__ == "__main__":
    main()
