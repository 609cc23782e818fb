#!/usr/bin/env python
"""Per-unit timeline of the f32 gate/residual epilogue of ONE tile (variant build -DLDT_EPI_STAMPS): clock64 of the first
epilogue warp of CTA 0 at: 0 before / 1 after the accumulator wait, then per 32-column unit u: 2+4u before tcgen05.ld,
3+4u after the staging stores, 4+4u after issuing the next unit's tcgen05.ld, 5+4u after the 8 store rounds + wait::ld."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ldt_b200 import _lib, build  # noqa: E402

lib_path = build.build_variant("stamps", ["LDT_EPI_STAMPS"])
_lib.LIB_PATH = lib_path
from ldt_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
for M, N, K in ((256, 256, 64), (256, 256, 1024), (8192, 1024, 1024)):
    A = torch.randn((M, K), device=dev).bfloat16()
    W = torch.randn((N, K), device=dev).bfloat16()
    b = torch.randn((N,), device=dev)
    out = torch.zeros((M, N), device=dev)
    gate = torch.randn((M // 32, N), device=dev)
    kw = dict(resid=out, gate=gate, gate_stride=N, rows_per_gate=32)
    for _ in range(3):
        ops.gemm(A, W, b, out, 3, backend=3, **kw)
    buf = torch.zeros((148 * 8,), dtype=torch.int64, device=dev)
    _lib.load().ldt_debug_set_gemm_counters(buf.data_ptr())
    ops.gemm(A, W, b, out, 3, backend=3, **kw)
    torch.cuda.synchronize()
    _lib.load().ldt_debug_set_gemm_counters(None)
    t = buf.cpu()[64:64 + 18].tolist()
    rel = [x - t[1] for x in t]
    print(f"M={M} N={N} K={K}: wait for accumulator {t[1] - t[0]} clk; relative to its arrival:", flush=True)
    for u in range(4):
        a, b_, c, d = rel[2 + 4 * u: 6 + 4 * u]
        print(f"   unit {u}: start {a:6d}  +staging stores {b_ - a:5d}  +issue next tcgen05.ld {c - b_:5d}  +8 store rounds, wait::ld {d - c:5d}   (done at {d})", flush=True)
