#!/usr/bin/env python
"""Sustained interleaved A/B of the product library against a variant build (ldt_b200.build.build_variant) on one GEMM
shape.  Build the variant first (CPU container):  python -c "from ldt_b200.build import build_variant as b; b('alt', ['LDT_GELU_AS'])"
usage: python scripts/exp_ab_lib.py alt [shape ...]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from ldt_b200 import _lib, ops  # noqa: E402
from scripts.exp_gemm_limits import graph_of, timed_with_clocks  # noqa: E402

dev = torch.device("cuda:0")
M = int(os.environ.get("LDT_AB_M", "8192"))
SH = {"qkv": (3072, 1024, 1), "fc_o": (1024, 1024, 3), "fc1": (4096, 1024, 2), "fc2": (1024, 4096, 3),
      "fc1_nogelu": (4096, 1024, 1), "fc2_noresid": (1024, 4096, 0), "fc_o_noresid": (1024, 1024, 0)}   # epilogue cost probes


def load_variant(tag):
    lib = C.CDLL(os.path.join(os.path.dirname(_lib.LIB_PATH), f"libldt_b200_{tag}.so"))
    for name, (res, args) in _lib.PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def main():
    tag = sys.argv[1]
    names = sys.argv[2:] or ["fc1"]
    libs = {"product": _lib.load(), tag: load_variant(tag)}
    g = torch.Generator().manual_seed(0)
    for name in names:
        if name == "attn":   # the fused projection + attention kernel
            A = torch.randn((M, 1024), generator=g).to(dev).bfloat16()
            Wp = [(torch.randn((3072, 1024), generator=g) / 32).to(dev).bfloat16() for _ in range(24)]
            bp = torch.randn((3072,), generator=g).to(dev)
            o = torch.empty((M, 1024), dtype=torch.bfloat16, device=dev)
            graphs = {}
            for k, lib in libs.items():
                _lib._lib = lib

                def body():
                    for i in range(24):
                        ops.qkv_attention(M // 32, 16, A, Wp[i], bp, o)
                graphs[k] = graph_of(body)
            _lib._lib = libs["product"]
            tot = {k: 0.0 for k in libs}
            for r in range(3):
                for k in libs:
                    ms, clk, pw = timed_with_clocks(graphs[k].replay, 0.8)
                    tot[k] += ms / 3
            for k in libs:
                print(f"attn  {k:8s}: {tot[k] / 24 * 1e3:8.2f} us (sustained, interleaved x3)", flush=True)
            continue
        N, K, epi = SH[name]
        A = (torch.randn((M, K), generator=g) * 0.5).to(dev).bfloat16()
        W = [(torch.randn((N, K), generator=g) / K ** 0.5).to(dev).bfloat16() for _ in range(24)]
        bias = torch.randn((N,), generator=g).to(dev)
        gate = torch.randn((1, N), generator=g).to(dev)
        out = torch.zeros((M, N), dtype=torch.float32 if epi in (0, 3) else torch.bfloat16, device=dev)
        kw = dict(resid=out, gate=gate, gate_stride=0, rows_per_gate=32) if epi == 3 else {}
        graphs = {}
        for k, lib in libs.items():
            _lib._lib = lib

            def body():
                for i in range(24):
                    ops.gemm(A, W[i], bias, out, epi, backend=3, **kw)
            graphs[k] = graph_of(body)
        _lib._lib = libs["product"]
        tot = {k: 0.0 for k in libs}
        rounds = 3
        for r in range(rounds):
            for k in libs:
                ms, clk, pw = timed_with_clocks(graphs[k].replay, 0.8)
                tot[k] += ms / rounds
        fl = 2.0 * M * N * K
        for k in libs:
            us = tot[k] / 24 * 1e3
            print(f"{name:5s} {k:8s}: {us:8.2f} us {fl / us / 1e6:8.1f} TFLOP/s (sustained, interleaved x{rounds})", flush=True)


if __name__ == "__main__":
    main()
