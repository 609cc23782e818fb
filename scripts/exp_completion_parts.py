#!/usr/bin/env python
"""Where a completion step (BASELINE configs[4] shape, 64 clouds per GPU) spends its time beyond the unconditional token
pass: the per-step adaLN GEMM [B, t_dim] x [149504, t_dim]^T, the conditioning kernel, the cross-attention blocks."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ldt_b200 import DiffusionVPSDE, Score, ops  # noqa: E402
from ldt_b200.sampler import StepGraph  # noqa: E402
from tests.helpers import airplane_config, ns  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
c = ns(airplane_config())
c.score.condition = True
torch.manual_seed(0)
model = Score(c.score).to(dev).eval()
sde = DiffusionVPSDE(c.sde, device=dev)


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for per_sample_c, cross in ((False, False), (True, False), (False, True), (True, True)):
    sg = StepGraph(model, sde, B, 1000, "ancestral", 1e-6, False, dev, True, per_sample_c=per_sample_c, cross_attention=cross)
    if per_sample_c or cross:
        sg.set_condition(torch.randn((B, 1024, 32), device=dev) if cross else None,
                         torch.randn((B, 1024), device=dev) * 0.5 if per_sample_c else None)
    x0 = torch.randn((B, 32, 120), device=dev)
    sg.N = 1
    sg.run(x0, 1, 0)       # captures
    ms = timed(lambda: (sg.step.zero_(), sg.graph.replay()))
    print(f"B={B} per_sample_c={per_sample_c} cross_attention={cross}: {ms:.4f} ms per step ({sg.launches_per_step} launches)", flush=True)
    if per_sample_c and not cross:
        ws = sg.ws
        ms_g = timed(lambda: ops.gemm(ws.sc, sg.P["w_ada"], sg.P["b_ada"], ws.mod, ops.EPI_BIAS_F32))
        print(f"   adaLN GEMM [{B} x 149504 x 1024] alone: {ms_g * 1e3:.1f} us", flush=True)
    del sg
