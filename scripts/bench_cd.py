#!/usr/bin/env python
"""Chamfer-matrix kernel alone: cloud pairs/s at 2048 x 2048 points (CUDA events, warm, 3 repetitions)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from ldt_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
P = 2048
g = torch.Generator().manual_seed(7)
a = torch.randn((n, P, 3), generator=g).to(dev)
b = torch.randn((n, P, 3), generator=g).to(dev)
ops.pairwise_cd(a, b)
torch.cuda.synchronize()
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.pairwise_cd(a, b)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 1e3 / 3
    print(f"{n}x{n} clouds: {n * n / t / 1e6:.4f} M cloud pairs/s, {n * n * P * P / t / 1e12:.3f} T point pairs/s, "
          f"{n * n * P * P * 8 / t / 1e12:.1f} TFLOP/s (8 flop per pair)", flush=True)
