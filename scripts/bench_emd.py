#!/usr/bin/env python
"""Approximate-EMD matrix kernel alone: cloud pairs/s at 2048 x 2048 points (LDT_EMD_SCALAR=1 selects the scalar kernel)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from ldt_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
P = 2048
g = torch.Generator().manual_seed(7)
a = torch.rand((n, P, 3), generator=g).to(dev)
b = torch.rand((n, P, 3), generator=g).to(dev)
out = ops.pairwise_emd(a, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = ops.pairwise_emd(a, b)
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 1e3
print(f"{n}x{n} clouds: {n * n / t:.0f} cloud pairs/s  (checksum {out.double().sum().item():.6f})", flush=True)
