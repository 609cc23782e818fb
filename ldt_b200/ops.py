"""Tensor-level wrappers over the C ABI.  Each function validates what the reference's CHECK_INPUT macro
validates (CUDA + contiguous, evaluation/pytorch_structural_losses/src/structural_loss.cpp:10-12), allocates
outputs with torch (caller-owned memory; the kernels never allocate) and launches on the current stream.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_BIAS_BF16, EPI_BIAS_F32, EPI_BIAS_GELU_BF16, EPI_BIAS_GELU_F32, EPI_BIAS_RELU_F32, EPI_GATE_RESID_F32, EPI_RESID_RELU_F32, GemmArgs, MlpArgs, check, load,
                   ptr, stream_ptr)


# ---- launch accounting / optional per-launch CUDA-event profiling (bench.py roofline) ----------------------
_LAUNCHES = [0]
_PROFILE = None  # None, or a list receiving (kind, start_event, end_event, note)


def launch_count() -> int:
    """Number of ldt_b200 kernels launched so far by this process (graph replays included)."""
    return _LAUNCHES[0]


def add_launches(n: int) -> None:
    _LAUNCHES[0] += n


class profile:
    """Context manager: record a CUDA-event pair around every kernel launch made through this module."""

    def __enter__(self):
        global _PROFILE
        _PROFILE = []
        return _PROFILE

    def __exit__(self, *a):
        global _PROFILE
        _PROFILE = None


class _launch:
    __slots__ = ("kind", "n", "note", "ev")

    def __init__(self, kind, n=1, note=None):
        self.kind, self.n, self.note, self.ev = kind, n, note, None

    def __enter__(self):
        if _PROFILE is not None:
            self.ev = torch.cuda.Event(enable_timing=True)
            self.ev.record()
        return self

    def __exit__(self, et, ev, tb):
        _LAUNCHES[0] += self.n
        if _PROFILE is not None and et is None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            _PROFILE.append((self.kind, self.ev, e, self.note))
        return False


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")


def nn_distance_idx(a: torch.Tensor, b: torch.Tensor):
    """(dist1, idx1, dist2, idx2) -- StructuralLossesBackend.NNDistance (structural_loss.cpp:80-99)."""
    _req(a, torch.float32, "set_d")
    _req(b, torch.float32, "set_q")
    if a.dim() != 3 or b.dim() != 3 or a.shape[2] != 3 or b.shape[2] != 3 or a.shape[0] != b.shape[0]:
        raise RuntimeError(f"expected [b,n,3] and [b,m,3], got {tuple(a.shape)} and {tuple(b.shape)}")
    bs, n, m = a.shape[0], a.shape[1], b.shape[1]
    dist1 = torch.empty((bs, n), dtype=torch.float32, device=a.device)
    idx1 = torch.empty((bs, n), dtype=torch.int32, device=a.device)
    dist2 = torch.empty((bs, m), dtype=torch.float32, device=a.device)
    idx2 = torch.empty((bs, m), dtype=torch.int32, device=a.device)
    with torch.cuda.device(a.device), _launch("nn_distance", 2):
        check(load().ldt_nn_distance(bs, n, ptr(a), m, ptr(b), ptr(dist1), ptr(idx1), ptr(dist2), ptr(idx2),
                                     stream_ptr()), "ldt_nn_distance")
    return dist1, idx1, dist2, idx2


def pairwise_cd(a: torch.Tensor, b: torch.Tensor, row_begin: int = 0, row_end: int | None = None) -> torch.Tensor:
    """Rows [row_begin,row_end) of the [na,nb] Chamfer matrix between cloud sets a [na,pa,3] and b [nb,pb,3]."""
    _req(a, torch.float32, "a")
    _req(b, torch.float32, "b")
    if a.dim() != 3 or b.dim() != 3 or a.shape[2] != 3 or b.shape[2] != 3:
        raise RuntimeError(f"expected [na,pa,3] and [nb,pb,3], got {tuple(a.shape)} and {tuple(b.shape)}")
    na, pa = a.shape[0], a.shape[1]
    nb, pb = b.shape[0], b.shape[1]
    row_end = na if row_end is None else row_end
    out = torch.empty((row_end - row_begin, nb), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device), _launch("pairwise_cd"):
        check(load().ldt_pairwise_cd(na, nb, pa, pb, ptr(a), ptr(b), row_begin, row_end, ptr(out), stream_ptr()),
              "ldt_pairwise_cd")
    return out


def pairwise_cd_upper(a: torch.Tensor, row_first: int = 0, row_step: int = 1, out: torch.Tensor | None = None) -> torch.Tensor:
    """Upper triangle (j >= i) of the [n,n] Chamfer matrix of cloud set a [n,p,3] against itself, for the rows
    i = row_first, row_first + row_step, ...; written into ``out`` [n,n] (zeros where nothing is computed).  Mirror with
    ``mirror_upper``: the kernel's (i,j) and (j,i) values are bit-identical."""
    _req(a, torch.float32, "a")
    if a.dim() != 3 or a.shape[2] != 3:
        raise RuntimeError(f"expected [n,p,3], got {tuple(a.shape)}")
    n, p = a.shape[0], a.shape[1]
    if out is None:
        out = torch.zeros((n, n), dtype=torch.float32, device=a.device)
    else:
        _req(out, torch.float32, "out")
        if tuple(out.shape) != (n, n):
            raise RuntimeError(f"out must be [{n},{n}], got {tuple(out.shape)}")
    with torch.cuda.device(a.device), _launch("pairwise_cd"):
        check(load().ldt_pairwise_cd_upper(n, p, ptr(a), row_first, row_step, ptr(out), stream_ptr()),
              "ldt_pairwise_cd_upper")
    return out


def measure_fma_peak(device, packed: bool, iters: int = 4096, blocks_per_sm: int = 8, reps: int = 5) -> float:
    """Measured FP32 FMA rate in TFLOP/s on ``device`` (scalar FFMA or packed FFMA2 chains; csrc/diag.cu)."""
    out = torch.zeros(1, dtype=torch.float32, device=device)
    flop = C.c_longlong(0)
    best = 0.0
    with torch.cuda.device(device):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(reps + 1):
            e0.record()
            with _launch("fma_peak"):
                check(load().ldt_debug_fma_peak(iters, int(packed), blocks_per_sm, ptr(out), C.byref(flop), stream_ptr()),
                      "ldt_debug_fma_peak")
            e1.record()
            torch.cuda.synchronize()
            best = max(best, flop.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def mirror_upper(u: torch.Tensor) -> torch.Tensor:
    """Full symmetric matrix from its upper triangle (entries below the diagonal of ``u`` are ignored)."""
    return torch.triu(u) + torch.triu(u, 1).t()


def _check_sets(a, b, who):
    _req(a, torch.float32, "set_d")
    _req(b, torch.float32, "set_q")
    if a.dim() != 3 or b.dim() != 3 or a.shape[2] != 3 or b.shape[2] != 3 or a.shape[0] != b.shape[0]:
        raise RuntimeError(f"{who}: expected [b,n,3] and [b,m,3], got {tuple(a.shape)} and {tuple(b.shape)}")


def match_cost(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """cost [b] of the approximate EMD matching (StructuralLosses.match_cost forward), match never materialised."""
    _check_sets(a, b, "match_cost")
    cost = torch.empty((a.shape[0],), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device), _launch("match_cost"):
        check(load().ldt_match_cost(a.shape[0], a.shape[1], b.shape[1], ptr(a), ptr(b), ptr(cost), stream_ptr()),
              "ldt_match_cost")
    return cost


def approx_match(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Dense match [b,m,n] as StructuralLossesBackend.ApproxMatch returns it (structural_loss.cpp:14-44)."""
    _check_sets(a, b, "approx_match")
    match = torch.empty((a.shape[0], b.shape[1], a.shape[1]), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device), _launch("approx_match"):
        check(load().ldt_approx_match(a.shape[0], a.shape[1], b.shape[1], ptr(a), ptr(b), ptr(match), stream_ptr()),
              "ldt_approx_match")
    return match


def match_cost_from_match(a: torch.Tensor, b: torch.Tensor, match: torch.Tensor) -> torch.Tensor:
    """StructuralLossesBackend.MatchCost (structural_loss.cpp:46-78)."""
    _check_sets(a, b, "match_cost_from_match")
    _req(match, torch.float32, "match")
    cost = torch.empty((a.shape[0],), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device), _launch("match_cost_from_match"):
        check(load().ldt_match_cost_from_match(a.shape[0], a.shape[1], b.shape[1], ptr(a), ptr(b), ptr(match), ptr(cost),
                                               stream_ptr()), "ldt_match_cost_from_match")
    return cost


def pairwise_emd(a: torch.Tensor, b: torch.Tensor, row_begin: int = 0, row_end: int | None = None) -> torch.Tensor:
    """Rows [row_begin,row_end) of the [na,nb] approximate-EMD matrix (match cost / points) between cloud sets."""
    _req(a, torch.float32, "a")
    _req(b, torch.float32, "b")
    if a.dim() != 3 or b.dim() != 3 or a.shape[2] != 3 or b.shape[2] != 3 or a.shape[1] != b.shape[1]:
        raise RuntimeError(f"expected [na,p,3] and [nb,p,3] with equal point counts, got {tuple(a.shape)} and {tuple(b.shape)}")
    row_end = a.shape[0] if row_end is None else row_end
    out = torch.empty((row_end - row_begin, b.shape[0]), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device), _launch("pairwise_emd"):
        check(load().ldt_pairwise_emd(a.shape[0], b.shape[0], a.shape[1], ptr(a), ptr(b), row_begin, row_end, ptr(out),
                                      stream_ptr()), "ldt_pairwise_emd")
    return out


def gemm(A: torch.Tensor, W: torch.Tensor, bias, out: torch.Tensor, epilogue: int, *, N: int | None = None,
         K: int | None = None, resid=None, gate=None, gate_stride: int = 0, rows_per_gate: int = 1,
         backend: int = 0, split_operands: bool = False) -> torch.Tensor:
    """out = epilogue(A @ W[:N,:K].T + bias).  A bf16 [M, lda], W bf16 [>=N, ldw]; strides taken from tensors.
    Both f32 (holding TF32-rounded values): the kind::tf32 contraction of the TF32 parity mode (operand_type 1);
    ``split_operands``: both laid out by :func:`split_tf32` (operand_type 2: the GELU epilogue then keeps full fp32)."""
    assert A.dtype == W.dtype and A.dtype in (torch.bfloat16, torch.float32) and A.dim() == 2 and W.dim() == 2
    assert A.stride(1) == 1 and W.stride(1) == 1 and out.stride(-1) == 1
    M = A.shape[0]
    K = A.shape[1] if K is None else K
    N = W.shape[0] if N is None else N
    out2 = out.view(-1, out.shape[-1]) if out.dim() != 2 else out
    args = GemmArgs(M=M, N=N, K=K, A=ptr(A), lda=A.stride(0), W=ptr(W), ldw=W.stride(0), bias=ptr(bias),
                    out=ptr(out2), ldo=out2.stride(0), epilogue=epilogue, resid=ptr(resid), gate=ptr(gate),
                    gate_stride=gate_stride, rows_per_gate=rows_per_gate, backend=backend,
                    operand_type=(2 if split_operands else 1) if A.dtype == torch.float32 else 0)
    with torch.cuda.device(A.device), _launch("gemm", 1, (M, N, K, epilogue)):
        check(load().ldt_gemm_bf16(C.byref(args), stream_ptr()), "ldt_gemm_bf16")
    return out


def mlp_sync_buffer(M: int, device) -> torch.Tensor:
    """Zeroed per-256-row completion counters for :func:`mlp` (self-cleaning: allocate once per workspace)."""
    return torch.zeros((load().ldt_mlp_sync_words(M),), dtype=torch.int32, device=device)


def mlp_supported(M: int, Cc: int, inner: int) -> bool:
    """Shapes the fused MLP kernel takes (whole 256-wide tiles; enough rows to fill CTA pairs)."""
    return M >= 1024 and Cc % 256 == 0 and inner % 256 == 0


def mlp(A: torch.Tensor, W1: torch.Tensor, b1, hidden: torch.Tensor, W2: torch.Tensor, b2, out: torch.Tensor, sync: torch.Tensor,
        *, resid=None, gate=None, gate_stride: int = 0, rows_per_gate: int = 1) -> torch.Tensor:
    """out = resid + gate * (GELU(A @ W1.T + b1) @ W2.T + b2) in one launch (MLP + gated residual, layers.py:110-133,219).
    A bf16 [M, C]; W1 bf16 [inner, C]; hidden bf16 [M, inner] scratch; W2 bf16 [C, inner]; out/resid f32 [M, C]."""
    assert A.dtype == torch.bfloat16 and W1.dtype == torch.bfloat16 and W2.dtype == torch.bfloat16
    assert hidden.dtype == torch.bfloat16 and out.dtype == torch.float32 and sync.dtype == torch.int32
    M, Cc = A.shape
    inner = W1.shape[0]
    assert W2.shape[0] == Cc and hidden.shape[0] >= M and hidden.shape[1] >= inner
    assert sync.numel() >= load().ldt_mlp_sync_words(M)
    resid = out if resid is None else resid
    args = MlpArgs(M=M, C=Cc, inner=inner, A=ptr(A), lda=A.stride(0), W1=ptr(W1), ldw1=W1.stride(0), bias1=ptr(b1),
                   hidden=ptr(hidden), ldh=hidden.stride(0), W2=ptr(W2), ldw2=W2.stride(0), bias2=ptr(b2),
                   resid=ptr(resid), out=ptr(out), ldo=out.stride(0), gate=ptr(gate), gate_stride=gate_stride,
                   rows_per_gate=rows_per_gate, sync=ptr(sync))
    with torch.cuda.device(A.device), _launch("mlp", 1, (M, Cc, inner)):
        check(load().ldt_mlp_bf16(C.byref(args), stream_ptr()), "ldt_mlp_bf16")
    return out


def cast_pad_bf16(x: torch.Tensor, ld_out: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """f32 [rows, cols] -> bf16 [rows, ld_out] zero-padded."""
    _req(x, torch.float32, "x")
    x2 = x.view(-1, x.shape[-1])
    rows, cols = x2.shape
    if out is None:
        out = torch.empty((rows, ld_out), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device), _launch("cast"):
        check(load().ldt_cast_pad_bf16(rows, cols, ptr(x2), x2.stride(0), ptr(out), ld_out, stream_ptr()),
              "ldt_cast_pad_bf16")
    return out


def pack_weight(w: torch.Tensor, ld_out: int | None = None) -> torch.Tensor:
    """Conv1d(k=1)/Linear weight [out, in(,1)] f32 -> bf16 [out, ld_out] K-major, zero-padded to a multiple of 64."""
    w2 = w.detach().reshape(w.shape[0], -1).contiguous().float()
    rows, cols = w2.shape
    ld_out = ((cols + 63) // 64) * 64 if ld_out is None else ld_out
    out = torch.empty((rows, ld_out), dtype=torch.bfloat16, device=w.device)
    with torch.cuda.device(w.device), _launch("pack"):
        check(load().ldt_pack_weights(rows, cols, ptr(w2), cols, ptr(out), ld_out, stream_ptr()), "ldt_pack_weights")
    return out


def layernorm_mod(x: torch.Tensor, out: torch.Tensor, *, shift=None, scale=None, mod_stride: int = 0,
                  rows_per_mod: int = 1, weight=None, bias=None, eps: float = 1e-6) -> torch.Tensor:
    rows, Cc = x.shape
    with torch.cuda.device(x.device), _launch("layernorm"):
        check(load().ldt_layernorm_mod_bf16(rows, Cc, ptr(x), ptr(shift), ptr(scale), mod_stride, rows_per_mod,
                                            ptr(weight), ptr(bias), eps, ptr(out), stream_ptr()),
              "ldt_layernorm_mod_bf16")
    return out


def time_embedding(t, freq, w0, b0, w1, b1, extra, c_out, silu_out, scratch) -> None:
    R, half, D = t.shape[0], freq.shape[0], w1.shape[0]
    with torch.cuda.device(t.device), _launch("time_embedding", 3):
        check(load().ldt_time_embedding(R, half, D, ptr(t), ptr(freq), ptr(w0), ptr(b0), ptr(w1), ptr(b1), ptr(extra),
                                        ptr(c_out), ptr(silu_out), ptr(scratch), stream_ptr()), "ldt_time_embedding")


def attention_nk32(B: int, H: int, Nq: int, dh: int, q, ldq: int, k, v, ldkv: int, o) -> None:
    with torch.cuda.device(o.device), _launch("attention"):
        check(load().ldt_attention_nk32(B, H, Nq, dh, ptr(q), ldq, ptr(k), ptr(v), ldkv, ptr(o), stream_ptr()),
              "ldt_attention_nk32")


def round_pad_tf32(x: torch.Tensor, ld_out: int | None = None, out: torch.Tensor | None = None, silu: bool = False) -> torch.Tensor:
    """f32 [rows, cols] -> f32 [rows, ld_out]: (SiLU, then) rounded to the nearest TF32 value, columns zero-padded."""
    _req(x, torch.float32, "x")
    rows, cols = x.shape
    ld_out = cols if ld_out is None else ld_out
    if out is None:
        out = torch.empty((rows, ld_out), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device), _launch("round_pad_tf32"):
        check(load().ldt_round_pad_tf32(rows, cols, ptr(x), x.stride(0), ptr(out), out.stride(0), int(silu), stream_ptr()),
              "ldt_round_pad_tf32")
    return out


def split_tf32(x: torch.Tensor, ld_part: int | None = None, weight_side: bool = False, out: torch.Tensor | None = None,
               silu: bool = False) -> torch.Tensor:
    """f32 [rows, cols] -> f32 [rows, 3*ld_part], the error-compensated TF32 operand layout ("3xTF32"): activations
    [hi | hi | lo], weights [hi | lo | hi]; one kind::tf32 contraction over 3*ld_part then has fp32-grade accuracy.
    ``silu`` applies SiLU before the split (the adaLN input)."""
    _req(x, torch.float32, "x")
    rows, cols = x.shape
    ld_part = ((cols + 31) // 32) * 32 if ld_part is None else ld_part
    if out is None:
        out = torch.empty((rows, 3 * ld_part), dtype=torch.float32, device=x.device)
    elif out.shape != (rows, 3 * ld_part) or out.dtype != torch.float32 or not out.is_contiguous():
        raise RuntimeError(f"split_tf32: out must be contiguous f32 {(rows, 3 * ld_part)}")
    with torch.cuda.device(x.device), _launch("split_tf32"):
        check(load().ldt_split_tf32(rows, cols, ptr(x), x.stride(0), ptr(out), ld_part, int(weight_side), int(silu), stream_ptr()),
              "ldt_split_tf32")
    return out


def layernorm_mod_f32(x: torch.Tensor, out: torch.Tensor, *, shift=None, scale=None, mod_stride: int = 0,
                      rows_per_mod: int = 1, weight=None, bias=None, eps: float = 1e-6, round_tf32: bool = True) -> torch.Tensor:
    rows, Cc = x.shape
    with torch.cuda.device(x.device), _launch("layernorm"):
        check(load().ldt_layernorm_mod_f32(rows, Cc, ptr(x), ptr(shift), ptr(scale), mod_stride, rows_per_mod,
                                           ptr(weight), ptr(bias), eps, ptr(out), int(round_tf32), stream_ptr()), "ldt_layernorm_mod_f32")
    return out


def attention_nk32_f32(B: int, H: int, Nq: int, dh: int, q, ldq: int, k, v, ldkv: int, o, round_tf32: bool = True) -> None:
    with torch.cuda.device(o.device), _launch("attention"):
        check(load().ldt_attention_nk32_f32(B, H, Nq, dh, ptr(q), ldq, ptr(k), ptr(v), ldkv, ptr(o), int(round_tf32), stream_ptr()),
              "ldt_attention_nk32_f32")


def attention_longkv_f32(B: int, H: int, Nq: int, Nk: int, dh: int, q, ldq: int, k, v, ldkv: int, o) -> None:
    with torch.cuda.device(o.device), _launch("attention"):
        check(load().ldt_attention_longkv_f32(B, H, Nq, Nk, dh, ptr(q), ldq, ptr(k), ptr(v), ldkv, ptr(o), stream_ptr()),
              "ldt_attention_longkv_f32")


def score_forward(plan, launches: int, x_tokens, mod, mod_stride: int, out) -> None:
    """The whole token pass as one C call (ldt_score_forward); ``plan`` is a _lib.ScorePlan, ``launches`` its kernel count."""
    with torch.cuda.device(out.device), _launch("score_forward", launches):
        check(load().ldt_score_forward(C.byref(plan), ptr(x_tokens), ptr(mod), mod_stride, ptr(out), stream_ptr()),
              "ldt_score_forward")


def sample_loop(args, launches: int) -> None:
    """ldt_sample_loop on the current stream; ``launches`` = kernels the call issues (for the launch counter)."""
    with _launch("sample_loop", launches):
        check(load().ldt_sample_loop(C.byref(args), stream_ptr()), "ldt_sample_loop")


def set_attention_backend(backend: int) -> None:
    """0 = tcgen05 attention kernel where it applies (default), 1 = warp-level mma.sync kernels (cross-check)."""
    check(load().ldt_debug_set_attention_backend(int(backend)), "ldt_debug_set_attention_backend")


def attention_longkv(B: int, H: int, Nq: int, Nk: int, dh: int, q, ldq: int, k, v, ldkv: int, o) -> None:
    with torch.cuda.device(o.device), _launch("attention_longkv"):
        check(load().ldt_attention_longkv(B, H, Nq, Nk, dh, ptr(q), ldq, ptr(k), ptr(v), ldkv, ptr(o), stream_ptr()),
              "ldt_attention_longkv")


def qkv_attention(B: int, H: int, A, Wp, bias_p, out) -> None:
    """Fused head-major QKV projection + 32-token self-attention (ldt_qkv_attention_bf16)."""
    K = A.shape[1]
    with torch.cuda.device(out.device), _launch("qkv_attention", 1, (B * 32, 3 * 64 * H, K, -1)):
        check(load().ldt_qkv_attention_bf16(B, H, K, ptr(A), A.stride(0), ptr(Wp), Wp.stride(0), ptr(bias_p), ptr(out),
                                            stream_ptr()), "ldt_qkv_attention_bf16")


def sde_step(predictor: int, x, params, z, coef_table, step_index, seed: int, offset: int, offset_per_step: int,
             rng_grid: int, x_next, x_mean, rng_state=None) -> None:
    """rng_state: device int64[2] = {seed, base offset} (as unsigned bit patterns) added to seed / offset in the kernel."""
    with torch.cuda.device(x.device), _launch("sde_step"):
        check(load().ldt_sde_step(predictor, x.numel(), ptr(x), ptr(params), ptr(z), ptr(coef_table), ptr(step_index),
                                  seed, offset, offset_per_step, ptr(rng_state), rng_grid, ptr(x_next), ptr(x_mean),
                                  stream_ptr()), "ldt_sde_step")


def pndm_transfer(x, et, coef, out) -> None:
    with torch.cuda.device(x.device), _launch("pndm_transfer"):
        check(load().ldt_pndm_transfer(x.numel(), ptr(x), ptr(et), ptr(coef), ptr(out), stream_ptr()), "ldt_pndm_transfer")


def lincomb4(coefs, tensors, scale: float, out) -> None:
    (c0, c1, c2, c3), (a0, a1, a2, a3) = coefs, tensors
    with torch.cuda.device(out.device), _launch("lincomb4"):
        check(load().ldt_lincomb4(out.numel(), c0, ptr(a0), c1, ptr(a1), c2, ptr(a2), c3, ptr(a3), scale, ptr(out),
                                  stream_ptr()), "ldt_lincomb4")


def batch_mean_norm(x: torch.Tensor) -> torch.Tensor:
    """0-dim tensor: mean over the batch of the per-sample L2 norms (torch.norm(x.reshape(B,-1), dim=-1).mean())."""
    _req(x, torch.float32, "x")
    B = x.shape[0]
    norms = torch.empty((B,), dtype=torch.float32, device=x.device)
    out = torch.empty((1,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device), _launch("batch_mean_norm", 2):
        check(load().ldt_batch_mean_norm(B, x.numel() // B, ptr(x), ptr(norms), ptr(out), stream_ptr()),
              "ldt_batch_mean_norm")
    return out[0]


def advance_step(step_index) -> None:
    with torch.cuda.device(step_index.device), _launch("advance_step"):
        check(load().ldt_advance_step(ptr(step_index), stream_ptr()), "ldt_advance_step")


def select_row(table, step_index, out) -> None:
    with torch.cuda.device(out.device), _launch("select_row"):
        check(load().ldt_select_row(ptr(table), table.shape[1], ptr(step_index), ptr(out), stream_ptr()),
              "ldt_select_row")


def cond_silu(table, step_index, extra, c_out, silu_out) -> None:
    R, D = silu_out.shape
    with torch.cuda.device(silu_out.device), _launch("cond_silu"):
        check(load().ldt_cond_silu(R, D, ptr(table), ptr(step_index), ptr(extra), ptr(c_out), ptr(silu_out), stream_ptr()),
              "ldt_cond_silu")


def furthest_point_sample(xyz: torch.Tensor, m: int, min_sq_norm: float = 1e-3) -> torch.Tensor:
    """idx [b,m] int32 -- pointnet2_utils.furthest_point_sample(xyz [b,n,3], m) (see include/ldt_b200.h)."""
    _req(xyz, torch.float32, "xyz")
    if xyz.dim() != 3 or xyz.shape[2] != 3:
        raise RuntimeError(f"expected [b,n,3], got {tuple(xyz.shape)}")
    b, n = xyz.shape[0], xyz.shape[1]
    idx = torch.empty((b, m), dtype=torch.int32, device=xyz.device)
    with torch.cuda.device(xyz.device), _launch("fps"):
        check(load().ldt_furthest_point_sample(b, n, m, ptr(xyz), min_sq_norm, ptr(idx), stream_ptr()),
              "ldt_furthest_point_sample")
    return idx


def knn_indices(k: int, xyz: torch.Tensor, centers: torch.Tensor) -> torch.Tensor:
    """idx [b,s,k] int32 -- knn_point(k, xyz [b,n,3], centers [b,s,3]) (model/Compressor/layers.py:86-98)."""
    _req(xyz, torch.float32, "xyz")
    _req(centers, torch.float32, "centers")
    if xyz.dim() != 3 or centers.dim() != 3 or xyz.shape[2] != 3 or centers.shape[2] != 3 or xyz.shape[0] != centers.shape[0]:
        raise RuntimeError(f"expected [b,n,3] and [b,s,3], got {tuple(xyz.shape)} and {tuple(centers.shape)}")
    b, n, s = xyz.shape[0], xyz.shape[1], centers.shape[1]
    idx = torch.empty((b, s, k), dtype=torch.int32, device=xyz.device)
    with torch.cuda.device(xyz.device), _launch("knn"):
        check(load().ldt_knn_indices(b, n, s, k, ptr(xyz), ptr(centers), ptr(idx), stream_ptr()), "ldt_knn_indices")
    return idx



_NORMALIZE = {None: 0, "center": 1, "anchor": 2}


def group_features(xyz: torch.Tensor, fea: torch.Tensor, center_idx: torch.Tensor, group_idx: torch.Tensor, normalize,
                   alpha, beta, ld_out: int | None = None) -> torch.Tensor:
    """LocalGrouper's normalised group features (model/Compressor/layers.py:300-317) as rows [b*s*k, ld_out] f32:
    the A operand of PreExtraction's first 1x1 convolution.  xyz [b,n,3], fea [b,n,d] f32; center_idx [b,s], group_idx [b,s,k] i32."""
    _req(xyz, torch.float32, "xyz")
    _req(fea, torch.float32, "fea")
    _req(center_idx, torch.int32, "center_idx")
    _req(group_idx, torch.int32, "group_idx")
    b, n, d = fea.shape
    s, k = group_idx.shape[1], group_idx.shape[2]
    if xyz.shape != (b, n, 3) or center_idx.shape != (b, s):
        raise RuntimeError(f"group_features: inconsistent shapes {tuple(xyz.shape)} {tuple(fea.shape)} {tuple(center_idx.shape)} {tuple(group_idx.shape)}")
    mode = _NORMALIZE[normalize]
    ld_out = ((2 * d + 3 + 31) // 32) * 32 if ld_out is None else ld_out
    out = torch.empty((b * s * k, ld_out), dtype=torch.float32, device=fea.device)
    partial = torch.empty((b, s, 2), dtype=torch.float64, device=fea.device) if mode else None
    if mode:
        alpha = alpha.detach().reshape(-1).float().contiguous()
        beta = beta.detach().reshape(-1).float().contiguous()
    with torch.cuda.device(fea.device), _launch("group_features", 2 if mode else 1):
        check(load().ldt_group_features(b, n, s, k, d, ptr(xyz), ptr(fea), ptr(center_idx), ptr(group_idx), mode,
                                        ptr(alpha) if mode else None, ptr(beta) if mode else None, ptr(partial), ptr(out), ld_out,
                                        stream_ptr()), "ldt_group_features")
    return out


def group_max(x: torch.Tensor, k: int, c: int | None = None) -> torch.Tensor:
    """x f32 [groups*k, ldx] -> [groups, c]: max over each group's k rows."""
    _req(x, torch.float32, "x")
    c = x.shape[1] if c is None else c
    groups = x.shape[0] // k
    out = torch.empty((groups, c), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device), _launch("group_max"):
        check(load().ldt_group_max(groups, k, c, ptr(x), x.stride(0), ptr(out), out.stride(0), stream_ptr()), "ldt_group_max")
    return out


__all__ = [n for n in dir() if not n.startswith("_")] + ["_lib"]
