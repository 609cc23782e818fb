#!/usr/bin/env python
"""Experiment: run the sampler as `chains` independent sub-batch chains on separate CUDA streams (each its own
captured step graph) and report clouds/s.  Used to decide whether kernel-level concurrency hides wave quantisation,
launch gaps and exposed epilogue tails.  usage: python scripts/exp_streams.py --batch 256 --chains 1,2,4 --sde-steps 200
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from ldt_b200 import DiffusionVPSDE, Score  # noqa: E402
from ldt_b200.sampler import StepGraph  # noqa: E402
from tests.helpers import airplane_config, ns  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--chains", default="1,2,4")
    ap.add_argument("--sde-steps", type=int, default=200)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    c = ns(airplane_config())
    torch.manual_seed(0)
    model = Score(c.score).to(dev).eval()
    sde = DiffusionVPSDE(c.sde, device=dev)
    N = a.sde_steps
    for nch in [int(x) for x in a.chains.split(",")]:
        b = a.batch // nch
        streams = [torch.cuda.Stream() for _ in range(nch)]
        sgs = []
        for i in range(nch):
            # separate Score workspaces per chain: StepGraph asks score._workspace(B, 1) keyed by B -> clone via dict hack
            model._ws = {}
            sg = StepGraph(model, sde, b, N, "ancestral", 1e-6, False, dev, use_graph=True)
            sg.ws = model._workspace(b, 1, dev)
            model._ws = {}
            sgs.append(sg)
        x0 = torch.randn(b, 32, 120, device=dev)
        # capture each chain's graph on its own stream
        for sg, st in zip(sgs, streams):
            with torch.cuda.stream(st):
                sg.N = 1
                sg.run(x0, 1, 0)   # captures + 1 replay
                sg.N = N
        torch.cuda.synchronize()

        def run_all():
            for sg, st in zip(sgs, streams):
                sg.step.zero_()
                st.wait_stream(torch.cuda.current_stream())
            for _ in range(N):
                for sg, st in zip(sgs, streams):
                    with torch.cuda.stream(st):
                        sg.graph.replay()
            for st in streams:
                torch.cuda.current_stream().wait_stream(st)

        run_all()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_all()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"chains={nch} sub-batch={b}: {ms / N:.3f} ms per SDE step -> {a.batch / (ms / N * 1000 / 1e3):.2f} clouds/s at 1000 steps",
              flush=True)


if __name__ == "__main__":
    main()
