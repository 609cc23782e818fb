#!/usr/bin/env python
"""A/B of whole token passes (24 blocks, batch 256) as replayed CUDA graphs, interleaved A B A B ... so that thermal /
power-cap drift hits both arms alike.  usage: python scripts/exp_step_ab.py attr=valueA,valueB [rounds]
   e.g. fused_mlp=0,1      (Score attribute toggled between the arms)
        env:LDT_X=0,1      (environment variable read at capture time)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from ldt_b200 import Score  # noqa: E402
from tests.helpers import airplane_config, ns  # noqa: E402
from scripts.exp_gemm_limits import timed_with_clocks  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("AB_BATCH", "256"))


def main():
    key, vals = sys.argv[1].split("=")
    vals = vals.split(",")
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    model = Score(ns(airplane_config()).score).to(dev).eval()
    P = model.packed()
    ws = model._workspace(B, 1, dev)
    x = torch.randn((B * 32, 120), device=dev)
    out = torch.empty_like(x)
    mod = torch.randn((1, ws.mod_len), device=dev) * 0.1
    graphs = []
    for v in vals:
        if key == "gemm_mode":   # ldt_debug_set_gemm_mode value, read at capture time
            from ldt_b200 import _lib
            _lib.load().ldt_debug_set_gemm_mode(int(v))
        elif key == "attn_backend":   # 0 = tcgen05 attention, 1 = mma.sync (ldt_debug_set_attention_backend), read at capture time
            from ldt_b200 import ops
            ops.set_attention_backend(int(v))
        elif key == "pdl":   # programmatic dependent launch on / off (ldt_set_pdl), read at capture time
            from ldt_b200 import _lib
            _lib.load().ldt_set_pdl(int(v))
        elif key.startswith("env:"):
            os.environ[key[4:]] = v
        else:
            setattr(model, key, type(getattr(model, key))(int(v)))
        model.run_tokens(P, ws, x, mod, 0, out)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            model.run_tokens(P, ws, x, mod, 0, out)
        graphs.append(g)
    tot = [0.0] * len(vals)
    for r in range(rounds):
        for i, g in enumerate(graphs):
            ms, clk, pw = timed_with_clocks(g.replay, 1.5)
            tot[i] += ms
            print(f"round {r} {key}={vals[i]}: {ms:.4f} ms per token pass  [SM {clk} MHz, {pw:.0f} W]", flush=True)
    for i, v in enumerate(vals):
        print(f"mean {key}={v}: {tot[i] / rounds:.4f} ms", flush=True)


if __name__ == "__main__":
    main()
