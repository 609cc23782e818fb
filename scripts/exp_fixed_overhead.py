#!/usr/bin/env python
"""Fixed cost of one GEMM launch inside a replayed graph: the same output shape with K = 64 (one k-block) against the real
K, for the shapes of the token pass at M = 2048 and M = 8192.  usage: python scripts/exp_fixed_overhead.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ldt_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def graph_time(fn, n=24, reps=200):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    for _ in range(10):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / n * 1e3


for M in (512, 2048, 8192):
    for name, N, K, epi in (("fc_o", 1024, 1024, 3), ("fc2", 1024, 4096, 3), ("fc1", 4096, 1024, 2), ("bias_bf16 N=1024", 1024, 1024, 1)):
        out_dt = torch.float32 if epi in (0, 3) else torch.bfloat16
        res = []
        for k in (64, K):
            A = torch.randn((M, k), device=dev).bfloat16()
            W = (torch.randn((N, k), device=dev) / k ** 0.5).bfloat16()
            b = torch.randn((N,), device=dev)
            out = torch.zeros((M, N), dtype=out_dt, device=dev)
            gate = torch.randn((M // 32, N), device=dev)
            kw = dict(resid=out, gate=gate, gate_stride=N, rows_per_gate=32) if epi == 3 else {}
            res.append(graph_time(lambda: ops.gemm(A, W, b, out, epi, **kw)))
        print(f"M={M:5d} {name:18s}: K=64 {res[0]:6.2f} us   K={K} {res[1]:6.2f} us   -> mainloop part {res[1] - res[0]:6.2f} us", flush=True)

# floors: an empty-ish kernel, and a one-tile GEMM (one CTA pair, one k-block)
step = torch.zeros(1, dtype=torch.int32, device=dev)
print(f"graph launch floor (1-thread kernel): {graph_time(lambda: ops.advance_step(step)):.2f} us", flush=True)
x32 = torch.randn((2048, 1024), device=dev)
a16 = torch.empty((2048, 1024), dtype=torch.bfloat16, device=dev)
print(f"LayerNorm M=2048: {graph_time(lambda: ops.layernorm_mod(x32, a16)):.2f} us", flush=True)
for M, N in ((256, 256), (256, 1024), (2048, 1024)):
    A = torch.randn((M, 64), device=dev).bfloat16()
    W = torch.randn((N, 64), device=dev).bfloat16()
    b = torch.randn((N,), device=dev)
    o16 = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
    print(f"one-k-block GEMM M={M} N={N} bias->bf16: {graph_time(lambda: ops.gemm(A, W, b, o16, 1)):.2f} us", flush=True)
