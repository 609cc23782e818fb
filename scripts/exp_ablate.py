#!/usr/bin/env python
"""In-graph ablation table (ldt_b200.profiling.ablate_score_step) for a batch size: python scripts/exp_ablate.py [batch]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ldt_b200 import Score, profiling  # noqa: E402
from tests.helpers import airplane_config, ns  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model = Score(ns(airplane_config()).score).to("cuda:0").eval()
r = profiling.ablate_score_step(model, B)
print(f"batch {B}: token pass {r['token_pass_ms']:.4f} ms (first/last {r['token_pass_ms_first_last']}), sum of marginals "
      f"{r['sum_marginal_ms']:.4f} ms, residual {r['residual_ms']:.4f} ms")
for k, e in r["classes"].items():
    rate = f"{e['tflops']:.0f} TFLOP/s" if e.get("tflops") else f"{e['gbytes_per_s']:.0f} GB/s"
    print(f"  {k:14s} {e['launches']:3d} launches  {e['marginal_ms']:.4f} ms  {e['us_per_launch']:.2f} us/launch  {rate}")
