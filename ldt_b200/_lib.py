"""ctypes binding of ``libldt_b200.so`` (the C ABI declared in ``include/ldt_b200.h``).

There is deliberately no fallback: if the library is missing or a call fails, a ``RuntimeError`` is raised.
PyTorch is used only to own device memory and streams; every pointer crossing this boundary is a raw
``data_ptr()``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libldt_b200.so")
ABI_VERSION = 2   # == LDT_ABI_VERSION of include/ldt_b200.h (checked at load and in tests/test_abi_and_host.py)

EPI_BIAS_F32 = 0
EPI_BIAS_BF16 = 1
EPI_BIAS_GELU_BF16 = 2
EPI_GATE_RESID_F32 = 3
EPI_BIAS_GELU_F32 = 4
EPI_BIAS_RELU_F32 = 5
EPI_RESID_RELU_F32 = 6

PRED_ANCESTRAL = 0
PRED_REVERSE_DIFFUSION = 1
PRED_EULER_MARUYAMA = 2
PRED_DDIM = 3
PRED_CORRECTOR = 4
SDE_COEF_STRIDE = 8


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("A", C.c_void_p), ("lda", C.c_int),
        ("W", C.c_void_p), ("ldw", C.c_int),
        ("bias", C.c_void_p),
        ("out", C.c_void_p), ("ldo", C.c_int),
        ("epilogue", C.c_int),
        ("resid", C.c_void_p),
        ("gate", C.c_void_p),
        ("gate_stride", C.c_longlong),
        ("rows_per_gate", C.c_int),
        ("backend", C.c_int),
        ("operand_type", C.c_int),
    ]


class ScoreBlock(C.Structure):
    _fields_ = [("w_qkv_packed", C.c_void_p), ("b_qkv_packed", C.c_void_p), ("w_o", C.c_void_p), ("b_o", C.c_void_p),
                ("w_fc1", C.c_void_p), ("b_fc1", C.c_void_p), ("w_fc2", C.c_void_p), ("b_fc2", C.c_void_p)]


class ScorePlan(C.Structure):
    _fields_ = [("batch", C.c_int), ("tokens", C.c_int), ("z_dim", C.c_int), ("z_pad", C.c_int), ("hidden", C.c_int),
                ("heads", C.c_int), ("mlp_hidden", C.c_int), ("num_blocks", C.c_int),
                ("w_in", C.c_void_p), ("b_in", C.c_void_p), ("w_out", C.c_void_p), ("b_out", C.c_void_p),
                ("blocks", C.POINTER(ScoreBlock)),
                ("ws_xa", C.c_void_p), ("ws_h", C.c_void_p), ("ws_a", C.c_void_p), ("ws_att", C.c_void_p), ("ws_hid", C.c_void_p)]


class DecoderLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w_ln", "b_ln", "w_kv", "b_kv", "w_q", "b_q", "w_o", "b_o", "w_fc1", "b_fc1",
                                           "w_fc2", "b_fc2", "norm1_w", "norm1_b", "norm2_w", "norm2_b")]


class DecoderPlan(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("batch", "num_points", "z_dim", "z_pad", "hidden", "heads", "mlp_hidden", "n_layers")] + \
               [("layers", C.POINTER(DecoderLayer)), ("w_out", C.c_void_p), ("b_out", C.c_void_p)] + \
               [(n, C.c_void_p) for n in ("ws_e", "ws_x", "ws_kv", "ws_a", "ws_q", "ws_att", "ws_hid")]


class SampleArgs(C.Structure):
    _fields_ = [("score", C.POINTER(ScorePlan)), ("predictor", C.c_int), ("num_steps", C.c_int), ("use_graph", C.c_int),
                ("mod_table", C.c_void_p), ("mod_len", C.c_longlong), ("mod_cur", C.c_void_p), ("coef", C.c_void_p),
                ("step", C.c_void_p), ("rng_state", C.c_void_p), ("offset_per_step", C.c_ulonglong), ("rng_grid", C.c_int),
                ("x", C.c_void_p), ("x_mean", C.c_void_p), ("params", C.c_void_p)]


class MlpArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("C", C.c_int), ("inner", C.c_int),
        ("A", C.c_void_p), ("lda", C.c_int),
        ("W1", C.c_void_p), ("ldw1", C.c_int), ("bias1", C.c_void_p),
        ("hidden", C.c_void_p), ("ldh", C.c_int),
        ("W2", C.c_void_p), ("ldw2", C.c_int), ("bias2", C.c_void_p),
        ("resid", C.c_void_p), ("out", C.c_void_p), ("ldo", C.c_int),
        ("gate", C.c_void_p), ("gate_stride", C.c_longlong), ("rows_per_gate", C.c_int),
        ("sync", C.c_void_p),
    ]


# name -> (restype, argtypes); the single source of truth used by tests to check exported symbols
PROTOTYPES = {
    "ldt_abi_version": (C.c_int, []),
    "ldt_last_error_string": (C.c_char_p, []),
    "ldt_device_sm_count": (C.c_int, []),
    "ldt_set_pdl": (C.c_int, [C.c_int]),
    "ldt_nn_distance": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldt_pairwise_cd": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p]),
    "ldt_debug_set_attention_backend": (C.c_int, [C.c_int]),
    "ldt_debug_get_attention_backend": (C.c_int, []),
    "ldt_score_forward": (C.c_int, [C.POINTER(ScorePlan), C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]),
    "ldt_decoder_forward": (C.c_int, [C.POINTER(DecoderPlan), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldt_sample_loop": (C.c_int, [C.POINTER(SampleArgs), C.c_void_p]),
    "ldt_round_pad_tf32": (C.c_int, [C.c_longlong, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ldt_layernorm_mod_f32": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_void_p]),
    "ldt_attention_nk32_f32": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "ldt_attention_longkv_f32": (C.c_int, [C.c_int] * 5 + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ldt_debug_fma_peak": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_longlong), C.c_void_p]),
    "ldt_pairwise_cd_upper": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ldt_match_cost": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldt_approx_match": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldt_match_cost_from_match": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    "ldt_pairwise_emd": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p]),
    "ldt_gemm_bf16": (C.c_int, [C.POINTER(GemmArgs), C.c_void_p]),
    "ldt_mlp_bf16": (C.c_int, [C.POINTER(MlpArgs), C.c_void_p]),
    "ldt_mlp_sync_words": (C.c_int, [C.c_int]),
    "ldt_mlp_schedule_item": (C.c_int, [C.c_int] * 8 + [C.POINTER(C.c_int)] * 4),
    "ldt_debug_set_gemm_counters": (C.c_int, [C.c_void_p]),
    "ldt_debug_set_gemm_mode": (C.c_int, [C.c_int]),
    "ldt_debug_get_gemm_mode": (C.c_int, []),
    "ldt_cast_pad_bf16": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "ldt_pack_weights": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "ldt_layernorm_mod_bf16": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong,
                                         C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "ldt_time_embedding": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "ldt_attention_nk32": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ldt_attention_longkv": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ldt_qkv_attention_bf16": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p]),
    "ldt_sde_step": (C.c_int, [C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_ulonglong, C.c_ulonglong, C.c_ulonglong, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p]),
    "ldt_pndm_transfer": (C.c_int, [C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldt_lincomb4": (C.c_int, [C.c_longlong, C.c_float, C.c_void_p, C.c_float, C.c_void_p, C.c_float, C.c_void_p, C.c_float,
                              C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "ldt_batch_mean_norm": (C.c_int, [C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldt_advance_step": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ldt_select_row": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldt_cond_silu": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldt_furthest_point_sample": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "ldt_knn_indices": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldt_group_features": (C.c_int, [C.c_int] * 5 + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p]),
    "ldt_split_tf32": (C.c_int, [C.c_longlong, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ldt_group_max": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load the library (once).  Raises RuntimeError with build instructions when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
                "Build it with `python -m ldt_b200.build` (needs nvcc, cross-compiles for sm_100a).")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError here means header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if lib.ldt_abi_version() != ABI_VERSION:
            raise RuntimeError("libldt_b200.so ABI version mismatch; rebuild with `python -m ldt_b200.build --force`")
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().ldt_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def stream_ptr(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()
