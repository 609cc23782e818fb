#!/usr/bin/env python
"""Micro-benchmark of the dense-contraction core on the score net's four GEMM shapes (M = batch*32 rows).

usage: python scripts/bench_gemm.py [--batch 256] [--backends 2,3] [--reps 50]
Prints one line per (shape, backend): microseconds (CUDA events on the launching stream, L2 flushed by cycling
through enough distinct operand sets to exceed the 126 MB L2 is NOT done here on purpose: inside the sampler the
activations of one step are L2-resident, so this measures the same regime) and TFLOP/s.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from ldt_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--backends", default="2,3")
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--extra", action="store_true")
    ap.add_argument("--counters", action="store_true", help="print the CTA-pair kernel's stall counters (backend 3)")
    a = ap.parse_args()
    bench_fused(a.batch, a.reps)
    dev = torch.device("cuda:0")
    M = a.batch * 32
    shapes = [("qkv", 3072, 1024, 1), ("fc_o", 1024, 1024, 3), ("fc1", 4096, 1024, 2), ("fc2", 1024, 4096, 3)]
    if a.extra:
        shapes = [("fc1/bias-only", 4096, 1024, 1), ("fc1/gelu", 4096, 1024, 2), ("fc1/f32", 4096, 1024, 0)]
    g = torch.Generator().manual_seed(0)
    for name, N, K, epi in shapes:
        A = (torch.randn((M, K), generator=g) * 0.5).to(dev).bfloat16()
        W = (torch.randn((N, K), generator=g) / K ** 0.5).to(dev).bfloat16()
        bias = torch.randn((N,), generator=g).to(dev)
        gate = torch.randn((1, N), generator=g).to(dev)
        outs = {}
        for be in [int(x) for x in a.backends.split(",")]:
            dt = torch.float32 if epi == 3 else torch.bfloat16
            out = torch.zeros((M, N), dtype=dt, device=dev)
            kw = dict(resid=out, gate=gate, gate_stride=0, rows_per_gate=32) if epi == 3 else {}
            for _ in range(5):
                ops.gemm(A, W, bias, out, epi, backend=be, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                ops.gemm(A, W, bias, out, epi, backend=be, **kw)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / a.reps
            print(f"{name:5s} M={M} N={N} K={K} backend={be}: {us:8.2f} us  {2.0 * M * N * K / us / 1e6:8.1f} TFLOP/s", flush=True)
            if a.counters and be == 3:
                from ldt_b200 import _lib
                buf = torch.zeros((148 * 8,), dtype=torch.int64, device=dev)
                _lib.load().ldt_debug_set_gemm_counters(buf.data_ptr())
                ops.gemm(A, W, bias, out, epi, backend=be, **kw)
                torch.cuda.synchronize()
                _lib.load().ldt_debug_set_gemm_counters(None)
                c = buf.view(148, 8).cpu().double()
                lead, both = c[0::2], c
                print(f"      leader MMA thread: total {lead[:,0].mean():.0f} clk, wait TMA {lead[:,1].mean():.0f}, wait free acc "
                      f"{lead[:,2].mean():.0f}, tiles {lead[:,7].mean():.2f} | epilogue warp: total {both[:,3].mean():.0f}, wait acc "
                      f"{both[:,4].mean():.0f} | producer: total {both[:,6].mean():.0f}, wait free stage {both[:,5].mean():.0f}")
            if a.check:
                out.zero_()
                ops.gemm(A, W, bias, out, epi, backend=be, **kw)
                outs[be] = out.float().clone()
        if a.check and len(outs) > 1:
            ks = list(outs)
            for k in ks[1:]:
                d = (outs[k] - outs[ks[0]]).abs().max().item()
                print(f"   max |backend {k} - backend {ks[0]}| = {d:.3e}")


def bench_fused(batch, reps):
    dev = torch.device("cuda:0")
    H, dh, Hd = 16, 64, 1024
    M = batch * 32
    g = torch.Generator().manual_seed(1)
    A = torch.randn((M, Hd), generator=g).to(dev).bfloat16()
    Wp = (torch.randn((3 * Hd, Hd), generator=g) / 32).to(dev).bfloat16()
    bp = torch.randn((3 * Hd,), generator=g).to(dev)
    o = torch.empty((M, Hd), dtype=torch.bfloat16, device=dev)
    for _ in range(5):
        ops.qkv_attention(batch, H, A, Wp, bp, o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.qkv_attention(batch, H, A, Wp, bp, o)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    fl = 2.0 * M * 3 * Hd * Hd + 4.0 * batch * H * 32 * 32 * dh
    print(f"fused qkv+attention M={M}: {us:8.2f} us  {fl / us / 1e6:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
