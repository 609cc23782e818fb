#!/usr/bin/env python
"""The completion workload alone (BASELINE configs[4] shape: ConditionNet prologue + conditional sampling + decode), for
profiling under ncu.  usage: python scripts/run_completion.py [sde_steps] [batch]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from ldt_b200 import Compressor, DiffusionVPSDE, Score  # noqa: E402
from tests.helpers import airplane_config, ns  # noqa: E402

dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
Bc = int(sys.argv[2]) if len(sys.argv) > 2 else 64
c = ns(airplane_config())
c.score.condition = True
torch.manual_seed(0)
model = Score(c.score).to(dev).eval()
comp = Compressor(c.compressor).to(dev).eval()
sde = DiffusionVPSDE(c.sde, device=dev)
g = torch.Generator().manual_seed(99)
views = torch.rand((Bc, 3, 224, 224), generator=g).to(dev)
part = torch.randn((Bc, 2048, 3), generator=g)
part = (part / part.norm(dim=-1).max(dim=1)[0][:, None, None]).to(dev)


class Trainer:   # the slice of completion_trainer/Latent_SDE_Trainer.py the sampler sees
    def __init__(self):
        self.model, self.SDE = model, sde

    def score_fn(self, t, x, label=None, condition=None):
        t = t.to(x)
        params = self.model(x, t, label=label, condition=condition)
        return -params / torch.sqrt(self.SDE.var(t))[:, None, None], params


tr = Trainer()
score_fn = tr.score_fn


def step():
    with torch.no_grad():
        condition = model.c_net({"img": views, "pts": part})
        eps = sde.sample_discrete(score_fn=score_fn, N=N, corrector=None, predictor=c.sde.predictor, corrector_steps=1,
                                  shape=(c.score.z_scale, c.score.z_dim), time_eps=c.sde.sample_time_eps, label=None,
                                  denoise=c.sde.denoise, device=dev, num_samples=Bc, probability_flow=False, snr=c.sde.snr,
                                  condition=condition)
        return comp.sample((Bc, 2048), given_eps=eps)


step()
torch.cuda.synchronize()
t0 = time.time()
step()
torch.cuda.synchronize()
dt = time.time() - t0
print(f"completion: {Bc} clouds, {N} steps: {dt * 1e3:.1f} ms -> {Bc / dt:.2f} clouds/s", flush=True)
