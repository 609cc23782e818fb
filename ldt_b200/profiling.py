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


# ------------------------------------------------------------------------------------------------
# In-graph ablation: what each kernel class costs INSIDE the replayed step, at the clocks the step runs at
# ------------------------------------------------------------------------------------------------
CLASSES = ("qkv_attention", "fc_o", "fc1_gelu", "fc2", "layernorm_mod")


def _class_of(kind: str, args: tuple, kwargs: dict):
    if kind == "layernorm_mod":
        return "layernorm_mod"
    if kind == "qkv_attention":
        return "qkv_attention"
    if kind == "gemm":
        A, W, epi = args[0], args[1], args[4]
        if epi == ops.EPI_BIAS_GELU_BF16:
            return "fc1_gelu"
        if epi == ops.EPI_GATE_RESID_F32:
            return "fc_o" if A.shape[1] == W.shape[0] else "fc2"
    return None


def ablate_score_step(score, B: int, replays: int = 60, warm: int = 15) -> dict:
    """Marginal in-graph cost of every kernel class of the score-net token pass.

    The token pass (ln_in -> 24 blocks -> ln_out, the 150 launches a sampler step replays) is captured in a CUDA graph and
    timed over ``replays`` back-to-back replays with CUDA events; then re-captured with ONE class of launches left out.
    marginal(class) = full - without(class).  Unlike event pairs around eager launches (which add ~10 us of launch gap to
    every short kernel and run at boost clocks), the marginals are measured at the power-capped clocks of the real step
    and sum to at most the step: sum(marginals) + residual (launch-to-launch drain, ln_in / ln_out, cast) = full.
    Outputs of an ablated pass are garbage by construction; nothing reads them."""
    dev = score.ln_in.weight.device
    P = score.packed()
    ws = score._workspace(B, 1, dev)
    x = torch.randn((B * score.z_scale, score.z_dim), device=dev)
    out = torch.empty_like(x)
    mod = torch.randn((1, ws.mod_len), device=dev) * 0.1
    real = {"layernorm_mod": ops.layernorm_mod, "gemm": ops.gemm, "qkv_attention": ops.qkv_attention}
    counts: dict = {}
    counting = [False]

    def install(skip):
        def wrap(kind):
            fn = real[kind]

            def inner(*a, **k):
                c = _class_of(kind, a, k)
                if c is not None and counting[0]:
                    counts[c] = counts.get(c, 0) + 1
                if c is not None and c == skip:
                    return a[3] if kind == "gemm" else None
                return fn(*a, **k)
            return inner
        ops.layernorm_mod, ops.gemm, ops.qkv_attention = wrap("layernorm_mod"), wrap("gemm"), wrap("qkv_attention")

    def timed(skip):
        install(skip)
        try:
            if skip is None:
                counts.clear()
                counting[0] = True          # launches per token pass: counted on the eager warm run only
            score.run_tokens(P, ws, x, mod, 0, out)
            counting[0] = False
            torch.cuda.synchronize()
            n0 = ops.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                score.run_tokens(P, ws, x, mod, 0, out)
            ops.add_launches(-(ops.launch_count() - n0))
        finally:
            ops.layernorm_mod, ops.gemm, ops.qkv_attention = real["layernorm_mod"], real["gemm"], real["qkv_attention"]
        for _ in range(warm):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(replays):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / replays

    c_path, score.c_path = score.c_path, False   # the ablation hooks wrap the per-kernel Python calls
    try:
        return _ablate(score, B, timed, counts)
    finally:
        score.c_path = c_path


def _ablate(score, B, timed, counts) -> dict:
    full = timed(None)
    launches = dict(counts)
    # interleave: full is re-measured after the ablated passes and the two are averaged (thermal drift hits both)
    marg = {}
    for c in CLASSES:
        if launches.get(c, 0):
            marg[c] = timed(c)
    full2 = timed(None)
    full_avg = 0.5 * (full + full2)
    H, T = score.hidden_size, score.z_scale
    M = B * T
    flop = {"qkv_attention": 2.0 * M * H * 3 * H + 2.0 * 2 * B * score.num_heads * T * T * (H // score.num_heads),
            "fc_o": 2.0 * M * H * H, "fc1_gelu": 2.0 * M * H * 4 * H, "fc2": 2.0 * M * 4 * H * H}
    res = {"token_pass_ms": full_avg, "token_pass_ms_first_last": [full, full2], "classes": {}}
    total = 0.0
    for c, t_without in marg.items():
        m_ms = max(full_avg - t_without, 0.0)
        total += m_ms
        n = launches[c]
        e = {"launches": n, "marginal_ms": m_ms, "us_per_launch": 1e3 * m_ms / n}
        if c in flop:
            e["tflops"] = flop[c] * n / (m_ms * 1e-3) / 1e12 if m_ms > 0 else None
        else:   # LayerNorm + modulate: f32 row in, bf16 row out (algorithmic bytes), modulation rows broadcast
            e["gbytes_per_s"] = n * M * H * 6.0 / (m_ms * 1e-3) / 1e9 if m_ms > 0 else None
        res["classes"][c] = e
    res["sum_marginal_ms"] = total
    res["residual_ms"] = full_avg - total
    res["tensor_flop"] = sum(flop[c] * launches[c] for c in flop if c in launches)
    res["tensor_ms"] = sum(res["classes"][c]["marginal_ms"] for c in flop if c in res["classes"])
    return res
