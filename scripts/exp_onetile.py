#!/usr/bin/env python
"""Timeline of a minimal CTA-pair GEMM launch (one tile, one or few k-blocks) from the kernel's per-role clock64 counters."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ldt_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
for M, N, K, epi in ((256, 256, 64, 1), (256, 256, 1024, 1), (256, 256, 64, 3), (2048, 1024, 1024, 3)):
    A = torch.randn((M, K), device=dev).bfloat16()
    W = torch.randn((N, K), device=dev).bfloat16()
    b = torch.randn((N,), device=dev)
    out = torch.zeros((M, N), dtype=torch.float32 if epi == 3 else torch.bfloat16, device=dev)
    kw = dict(resid=out, gate=None) if epi == 3 else {}
    for _ in range(3):
        ops.gemm(A, W, b, out, epi, backend=3, **kw)
    buf = torch.zeros((148 * 8,), dtype=torch.int64, device=dev)
    _lib.load().ldt_debug_set_gemm_counters(buf.data_ptr())
    ops.gemm(A, W, b, out, epi, backend=3, **kw)
    torch.cuda.synchronize()
    _lib.load().ldt_debug_set_gemm_counters(None)
    c = buf.view(148, 8).cpu()[0]
    print(f"M={M} N={N} K={K} epi={epi}: MMA role total {int(c[0])} clk (wait TMA {int(c[1])}, wait acc {int(c[2])}) | epilogue role total "
          f"{int(c[3])} (wait acc {int(c[4])}) | producer total {int(c[6])} (wait stage {int(c[5])})", flush=True)
