#!/usr/bin/env python
"""ldt_nn_distance (the NmDistanceKernel drop-in, both directions with indices) alone: cloud pairs/s at 2048 x 2048 points."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from ldt_b200 import _lib, ops  # noqa: E402

if len(sys.argv) > 2:   # another build of the library (A/B)
    _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), sys.argv[2])

dev = torch.device("cuda:0")
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator().manual_seed(7)
a = torch.randn((bs, 2048, 3), generator=g).to(dev)
b = torch.randn((bs, 2048, 3), generator=g).to(dev)
ops.nn_distance_idx(a, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.nn_distance_idx(a, b)
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 1e3 / 10
print(f"batch {bs}: {bs / t / 1e3:.1f} k cloud pairs/s (both directions, distances + indices)", flush=True)
