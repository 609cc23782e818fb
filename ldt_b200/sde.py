"""Host mirror of the reference's VP-SDE object and discrete sampler entry point.

Mirrors ``DiffusionVPSDE`` (reference diffusion/diffusion_continuous.py:626-678) and ``sample_discrete``
(:133-338): same constructor keys (``cfg.sde``), same method names, same argument meaning and the same
errors for unknown predictors / correctors.  The arithmetic of the hot loop -- score conversion plus
predictor update plus noise injection -- runs in one sm_100a kernel per step (``ldt_sde_step``); when the
``score_fn`` is the Trainer closure over an ``ldt_b200.Score`` the whole N-step loop runs as a replayed CUDA
graph (``ldt_b200.sampler``).  PyTorch is used for the tiny per-step scalar tables (computed with the same op
sequence as the reference so they are bit-identical) and for the initial CPU-generator noise (:237).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._lib import PRED_ANCESTRAL, PRED_DDIM, PRED_EULER_MARUYAMA, PRED_REVERSE_DIFFUSION, SDE_COEF_STRIDE

_PRED_CODES = {
    "ancestral": PRED_ANCESTRAL,
    "reversediffusion": PRED_REVERSE_DIFFUSION,
    "eulermaruyama": PRED_EULER_MARUYAMA,
    "ddim": PRED_DDIM,
}


def torch_randn_launch_geometry(numel: int, device) -> tuple[int, int]:
    """(grid, philox offset advance) torch's CUDA ``normal_`` kernel uses for a float tensor of ``numel``
    elements: 256-thread blocks, grid capped at SMs * (max threads per SM / 256), 4 values per engine call."""
    props = torch.cuda.get_device_properties(device)
    block = 256
    unroll = 4
    blocks_per_sm = props.max_threads_per_multi_processor // block
    grid = min(props.multi_processor_count * blocks_per_sm, (numel + block - 1) // block)
    grid = max(grid, 1)
    per_call = ((numel - 1) // (block * grid * unroll) + 1) * 4
    return grid, per_call


class DiffusionVPSDE:
    """VP-SDE with linear beta(t); see reference diffusion/diffusion_continuous.py:626-678."""

    def __init__(self, args, device="cuda"):
        self.sigma2_0 = args.sigma2_0
        self.sde_type = args.sde_type
        self.time_eps = args.time_eps
        self.sample_time_eps = args.sample_time_eps
        self.beta_start = args.beta_start
        self.beta_end = args.beta_end
        self.device = torch.device(device)
        self.train_N = args.train_N
        # auxiliary constants (:638-646); kept because trainers read them
        self.delta_beta_half = torch.tensor(0.5 * (self.beta_end - self.beta_start), device=self.device)
        self.beta_frac = torch.tensor(self.beta_start / (self.beta_end - self.beta_start), device=self.device)
        if args.sample_mode == "discrete":
            self.N = args.sample_N
            # float64 linspace cast to the float32 of delta_beta_half, exactly as :648-651
            self.betas = torch.from_numpy(
                np.linspace(self.beta_start / self.N, self.beta_end / self.N, self.N, dtype=np.float64)
            ).to(self.delta_beta_half)
            self.alpha = 1.0 - self.betas
            self.alphas_cump = self.alpha.cumprod(dim=0)

    # coefficient functions (:655-678)
    def f(self, t):
        return -0.5 * self.g2(t)

    def g2(self, t):
        return self.beta_start + (self.beta_end - self.beta_start) * t

    def discrete(self, idx):
        return self.betas.index_select(0, idx), self.alpha.index_select(0, idx)

    def var(self, t):
        return 1.0 - (1.0 - self.sigma2_0) * torch.exp(
            -self.beta_start * t - 0.5 * (self.beta_end - self.beta_start) * t * t)

    def std(self, t):
        return torch.sqrt(self.var(t))

    def e2int_f(self, t):
        return torch.exp(-0.5 * self.beta_start * t - 0.25 * (self.beta_end - self.beta_start) * t * t)

    def inv_var(self, var):
        c = torch.log((1 - var) / (1 - self.sigma2_0))
        a = self.beta_end - self.beta_start
        return (-self.beta_start + torch.sqrt(np.square(self.beta_start) - 2 * a * c)) / a

    def sample_q(self, x_init, noise, var_t, m_t):
        return m_t * x_init + torch.sqrt(var_t) * noise

    # ------------------------------------------------------------------------------------------
    def step_coefficients(self, predictor: str, N: int, time_eps: float, probability_flow: bool, device,
                          raw_score: bool = False) -> tuple[torch.Tensor, torch.Tensor]:
        """Per-step scalar table [N, 8] f32 for ``ldt_sde_step`` and the timesteps [N].

        Every entry is produced by the same torch op sequence the reference applies per step (on the same
        device), so the values are bit-identical to what its predictors use.  ``raw_score`` makes entry 0
        equal -1 so the kernel's ``-params / c0`` passes a caller-computed score through unchanged.
        """
        timesteps = torch.linspace(1.0, time_eps, N, device=device)  # :238
        t = timesteps
        coef = torch.zeros((N, SDE_COEF_STRIDE), dtype=torch.float32, device=device)
        coef[:, 0] = -1.0 if raw_score else torch.sqrt(self.var(t))  # Trainer.score_fn :59-60
        betas = self.betas.to(device)
        if predictor == "ancestral":  # :152-162
            idx = (t * (N - 1) / 1.0).long()
            beta = betas[idx]
            coef[:, 1] = beta
            coef[:, 2] = torch.sqrt(1.0 - beta)
            coef[:, 3] = torch.sqrt(beta)
        elif predictor == "reversediffusion":  # :141-150
            dt = torch.tensor((1 - time_eps) / N, device=device)
            g2 = self.g2(t)
            coef[:, 1] = self.f(t)
            coef[:, 2] = g2 * (0.5 if probability_flow else 1.0)
            coef[:, 3] = dt
            coef[:, 4] = torch.zeros_like(g2) if probability_flow else torch.sqrt(g2)
            coef[:, 5] = torch.sqrt(dt)
        elif predictor == "eulermaruyama":  # :182-191
            dt = -1.0 / N
            g2 = self.g2(t)
            coef[:, 1] = self.f(t)
            coef[:, 2] = g2 * (0.5 if probability_flow else 1.0)
            coef[:, 3] = dt
            coef[:, 4] = torch.zeros_like(g2) if probability_flow else torch.sqrt(g2) * np.sqrt(-dt)
        elif predictor == "ddim":  # :164-180
            idx = (t * (N - 1) / 1.0).long()
            ac = self.alphas_cump.to(device)
            at = ac[idx]
            at_next = torch.where(idx - 1 < 0, torch.ones_like(at), ac[(idx - 1).clamp(min=0)])
            coef[:, 1] = at_next.sqrt()
            coef[:, 2] = (1 - at).sqrt()
            coef[:, 3] = at.sqrt()
            coef[:, 4] = (1 - at_next).sqrt()
        else:
            raise NotImplementedError("preditor not Implemented")
        return coef, timesteps

    def sample_model_ode(self, *a, **k):
        raise NotImplementedError(
            "continuous (ODE / torchdiffeq RK45) sampling is outside the ldt_b200 hot path; "
            "the shipped configs use sample_mode: discrete")

    def sample_discrete(self, score_fn, num_samples, N, predictor, corrector, corrector_steps, shape, time_eps,
                        probability_flow, denoise, snr, device, condition=None, label=None, print_steps=None):
        """Reverse-SDE sampling; signature and semantics of diffusion_continuous.py:133-338."""
        if predictor == "pndm":
            raise NotImplementedError("pndm predictor: not part of the ldt_b200 hot path yet (SURVEY.md 8f2)")
        if predictor is not None and predictor not in _PRED_CODES:
            raise NotImplementedError("preditor not Implemented")
        if corrector is not None:
            if corrector in ("langevin", "ancestral"):
                raise NotImplementedError(f"corrector '{corrector}': not part of the ldt_b200 hot path yet (SURVEY.md 8f2)")
            raise NotImplementedError("corrector not Implemented")
        device = torch.device(device)
        with torch.no_grad():
            # initial sample from the CPU generator, then H2D (:237)
            x = torch.randn((num_samples,) + tuple(shape)).to(device)
            if predictor is None:
                return x  # no predictor, no corrector: the loop is the identity (:243-249)

            from .sampler import fused_sample_loop, find_score_module  # late import (sampler imports Score)
            score_mod = find_score_module(score_fn, self)
            if score_mod is not None and print_steps is None and not isinstance(condition, dict):
                # the whole loop as one replayed graph; condition = (tokens | None, vector | 0.) as ConditionNet
                # returns it (completion_trainer/Latent_SDE_Trainer.py:150-151), label -> embedding (score.py:125-126)
                cond_tokens, extra = None, None
                if label is not None:
                    extra = score_mod.LabelEmbedding(label.to(device))
                if condition is not None:
                    if torch.is_tensor(condition[0]):
                        cond_tokens = condition[0].to(device)
                    if label is None and torch.is_tensor(condition[1]):
                        extra = condition[1].to(device)
                return fused_sample_loop(score_mod, self, x, N, predictor, time_eps, probability_flow, denoise,
                                         cond_tokens=cond_tokens, extra=extra)

            # generic path: arbitrary score_fn called once per step, fused update kernel in between
            coef, timesteps = self.step_coefficients(predictor, N, time_eps, probability_flow, device, raw_score=True)
            code = _PRED_CODES[predictor]
            zero = torch.zeros(1, dtype=torch.int32, device=device)
            x_mean = torch.empty_like(x)
            out_list, steps = None, None
            if print_steps is not None:
                out_list = [x]
                steps = (N - 1) // (print_steps - 2)
            for i in range(N):
                vec_t = torch.ones((num_samples,), device=device) * timesteps[i]
                score, params = score_fn(vec_t, x, label=label, condition=condition)
                z = torch.randn_like(x)  # same CUDA-generator draw as the reference (:160)
                src = (params if predictor == "ddim" else score).contiguous()
                x_next = torch.empty_like(x)
                ops.sde_step(code, x.contiguous(), src, z, coef[i:i + 1], zero, 0, 0, 0, 0, x_next, x_mean)
                x = x_next
                if out_list is not None and (i + 1) % steps == 0:
                    out_list.append(x_mean.clone())
            if out_list is not None:
                out_list.append(x_mean if denoise else x)
                return out_list
            return x_mean if denoise else x
