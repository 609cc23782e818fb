"""Shared test helpers: the shipped experiment configuration (values of the reference's
experiments/Latent_Diffusion_Trainer/airplane/config.yaml:46-113, the sections the hot path reads), a reduced
score config for fast CPU runs, golden-fixture loading and tolerance helpers."""
import os
from types import SimpleNamespace

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def airplane_config() -> dict:
    return {
        "score": dict(num_steps=1000, z_dim=120, z_scale=32, hidden_size=1024, num_heads=16, num_blocks=24,
                      num_categorys=1, c_dim=0.0, t_dim=1024, dropout=0.0, norm="layer_norm", learn_sigma=False,
                      act="swish", unet=False, AdaLN=True, condition=False, graphconv=False),
        "compressor": dict(pretrain_path=None, outsize=2048, max_outputs=2048, input_dim=3, z_dim=20, z_scales=32,
                           p_dim=256, n_layers=6, hidden_dim=128, num_heads=4, activation="swish",
                           encoder_dropout_p=0.0, decoder_dropout_p=0.0, norm="layer_norm", neighbors=128,
                           encoder_layers=2, mlp_ratio=4.0, min_sigma=-30, cluster_norm="anchor", norm_input=False,
                           pre_group=False, decoder_act=None, ActNorm=True, AdaLN=True, pos_embedding="center",
                           class_condition=False),
        "sde": dict(beta_start=0.1, beta_end=20, sde_type="vpsde", sigma2_0=0, iw_sample_p_mode="drop_all_iw",
                    iw_sample_q_mode="drop_all_iw", time_eps=0.01, ode_tol=0.00001, sample_time_eps=0.000001,
                    sample_mode="discrete", predictor="ancestral", corrector=None, train_N=1000, sample_N=1000,
                    snr=0.01, corrector_steps=1, denoise=True, probability_flow=False, alpha=1.0),
    }


def ns(d: dict) -> SimpleNamespace:
    return SimpleNamespace(**{k: (ns(v) if isinstance(v, dict) else v) for k, v in d.items()})


def small_score_cfg() -> SimpleNamespace:
    d = dict(airplane_config()["score"])
    d.update(hidden_size=128, num_heads=2, num_blocks=2, t_dim=128)
    return SimpleNamespace(**d)


def golden(name: str) -> dict:
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


def rel_rms_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / rms(b): the tolerance measure stated in SURVEY.md section 8(d)."""
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.pow(2).mean().sqrt().clamp_min(1e-30))


def shapes_of(module) -> dict:
    return {k: tuple(v.shape) for k, v in module.state_dict().items()}


def rms_rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """rms(a-b) / rms(b)."""
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-30))


def assert_bf16_close(got: torch.Tensor, ref: torch.Tensor, extra_rel: float = 0.0, what: str = "") -> None:
    """Element-wise bound for a value that was rounded once to bf16: half an ulp is 2^-9 relative; allow 2^-8 plus
    `extra_rel` of the tensor's rms for fp32 accumulation-order differences upstream of the rounding."""
    got, ref = got.double().cpu(), ref.double().cpu()
    rms = ref.pow(2).mean().sqrt()
    bound = ref.abs() * 2.0 ** -8 + (extra_rel + 1e-6) * rms
    bad = (got - ref).abs() > bound
    assert not bad.any(), f"{what}: {int(bad.sum())} elements outside the bf16 bound, worst {(got - ref).abs().max():.3e}"


_ORACLE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")


def _pointops_oracle():
    import ctypes as C
    lib = C.CDLL(os.path.join(_ORACLE_DIR, "liboracle_pointops.so"))
    lib.oracle_fps.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p]
    lib.oracle_knn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def oracle_fps(xyz: torch.Tensor, m: int, min_sq_norm: float = 1e-3) -> torch.Tensor:
    """C-oracle furthest point sampling: xyz [B,N,3] (CPU f32) -> long [B,m]."""
    x = xyz.detach().cpu().float().contiguous()
    idx = torch.empty((x.shape[0], m), dtype=torch.int32)
    _pointops_oracle().oracle_fps(x.shape[0], x.shape[1], m, x.data_ptr(), min_sq_norm, idx.data_ptr())
    return idx.long()


def oracle_knn(k: int, xyz: torch.Tensor, centers: torch.Tensor) -> torch.Tensor:
    x, c = xyz.detach().cpu().float().contiguous(), centers.detach().cpu().float().contiguous()
    idx = torch.empty((x.shape[0], c.shape[1], k), dtype=torch.int32)
    _pointops_oracle().oracle_knn(x.shape[0], x.shape[1], c.shape[1], k, x.data_ptr(), c.data_ptr(), idx.data_ptr())
    return idx.long()


def small_cond_score_cfg() -> SimpleNamespace:
    """Reduced completion score config: the shipped score config with ``condition: True`` (no completion yaml ships
    in the reference's experiments/), narrowed like small_score_cfg()."""
    c = small_score_cfg()
    c.condition = True
    return c


def small_unet_score_cfg() -> SimpleNamespace:
    """Reduced UNet score config (``unet: True`` is the default of the reference's model/scorenet/config.yaml)."""
    c = small_score_cfg()
    c.unet = True
    c.num_blocks = 4
    return c
