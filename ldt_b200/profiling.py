"""Per-kernel timing of one reverse-SDE step with CUDA events on the launching stream (used by bench.py for the
roofline object; numbers taken under ncu are never reported)."""
from __future__ import annotations

import torch

from . import ops

FLOP_ATTN_PER_SAMPLE_STEP = 24 * 2 * 2 * 16 * 32 * 32 * 64             # QK^T + PV of the 24 blocks
FLOP_GEMM_PER_SAMPLE_STEP = 19.44e9 - FLOP_ATTN_PER_SAMPLE_STEP        # SURVEY.md 8(d) minus QK^T/PV


def profile_score_step(score, B: int, reps: int = 3) -> dict:
    """Run the per-step token path eagerly `reps` times with an event pair around every launch.

    Returns total GEMM time per step (ms), the algorithmic GEMM FLOPs per step, and a per-kind breakdown."""
    dev = score.ln_in.weight.device
    P = score.packed()
    ws = score._workspace(B, 1, dev)
    x = torch.randn((B * score.z_scale, score.z_dim), device=dev)
    out = torch.empty_like(x)
    mod = torch.randn((1, ws.mod_len), device=dev) * 0.1
    score.run_tokens(P, ws, x, mod, 0, out)  # warm
    torch.cuda.synchronize()
    by_kind: dict = {}
    gemm_ms = total_ms = 0.0
    gemm_launches = 0
    fused = False
    for _ in range(reps):
        with ops.profile() as rec:
            score.run_tokens(P, ws, x, mod, 0, out)
            torch.cuda.synchronize()
            for kind, e0, e1, note in rec:
                ms = e0.elapsed_time(e1)
                key = kind if kind != "gemm" else f"gemm N={note[1]} K={note[2]}"
                by_kind[key] = by_kind.get(key, 0.0) + ms / reps
                total_ms += ms / reps
                if kind in ("gemm", "qkv_attention", "mlp"):   # the tcgen05 contraction kernels
                    gemm_ms += ms / reps
                    gemm_launches += 1
                if kind == "qkv_attention":
                    fused = True
    flop = B * (FLOP_GEMM_PER_SAMPLE_STEP + (FLOP_ATTN_PER_SAMPLE_STEP if fused else 0))
    return {"gemm_ms": gemm_ms, "total_ms": total_ms, "gemm_flop": flop,
            "gemm_launches": gemm_launches // reps, "by_kind": {k: round(v, 4) for k, v in by_kind.items()}}
