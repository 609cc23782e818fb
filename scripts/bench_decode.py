#!/usr/bin/env python
"""Compressor.sample (decode of B clouds to 2048 points) alone: time and per-kernel-kind breakdown (CUDA events)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from ldt_b200 import Compressor, ops  # noqa: E402
from tests.helpers import airplane_config, ns  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
c = ns(airplane_config())
torch.manual_seed(0)
comp = Compressor(c.compressor).to(dev).eval()
eps = torch.randn((B, 32, 120), device=dev)
for _ in range(2):
    comp.sample((B, 2048), given_eps=eps)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    comp.sample((B, 2048), given_eps=eps)
e1.record()
torch.cuda.synchronize()
print(f"decode of {B} clouds: {e0.elapsed_time(e1) / 3:.2f} ms", flush=True)
with ops.profile() as rec:
    comp.sample((B, 2048), given_eps=eps)
    torch.cuda.synchronize()
    agg = {}
    for kind, a, b, note in rec:
        key = kind if kind != "gemm" else f"gemm M={note[0]} N={note[1]} K={note[2]} epi={note[3]}"
        n, t = agg.get(key, (0, 0.0))
        agg[key] = (n + 1, t + a.elapsed_time(b))
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"  {k:50s} x{n:3d} {t:8.3f} ms", flush=True)
