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
from ._lib import (PRED_ANCESTRAL, PRED_CORRECTOR, PRED_DDIM, PRED_EULER_MARUYAMA, PRED_REVERSE_DIFFUSION,
                   SDE_COEF_STRIDE)

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

    def corrector_coefficients(self, N: int, time_eps: float, snr: float, device, raw_score: bool = False) -> torch.Tensor:
        """Per-step scalar table [N, 8] of the AncestralCorrector update (:212-229) for ``ldt_sde_step`` with
        LDT_PRED_CORRECTOR: [0] sqrt(var(t)) (or -1), [1] step_size, [2] sqrt(2 * step_size).

        ``alpha`` is 1: the reference tests ``self.__class__ in ["DiffusionVPSDE", ...]`` (:195,214), a class against
        strings, which is always False, so it takes ``alpha = torch.ones_like(t)``."""
        t = torch.linspace(1.0, time_eps, N, device=device)
        coef = torch.zeros((N, SDE_COEF_STRIDE), dtype=torch.float32, device=device)
        coef[:, 0] = -1.0 if raw_score else torch.sqrt(self.var(t))
        alpha = torch.ones_like(t)
        step_size = (snr * self.std(t)) ** 2 * 2 * alpha
        coef[:, 1] = step_size
        coef[:, 2] = torch.sqrt(step_size * 2)
        return coef

    def _pndm(self, score_fn, x, time_eps, condition, label):
        """The ``predictor == "pndm"`` branch (:260-316): pseudo linear multistep with a Runge-Kutta warm-up, driven by
        ``self.N`` / ``self.train_N`` (not the N argument), deterministic.  The reference's index arithmetic is kept
        as written, including ``timesteps[t_next * 2 - 1]`` with ``t_next == 0`` on the last step, which Python wraps
        to the LAST entry (t = 1).  The reference's ``at.view(-1, 1)`` broadcast only type-checks for batch sizes 1
        and 32 (all entries are equal, so the value is the scalar); here the scalar form runs for any batch."""
        device = x.device
        train_N = self.train_N
        timesteps = torch.linspace(time_eps, 1.0, self.N * 2)
        betas = torch.from_numpy(np.linspace(self.beta_start / train_N, self.beta_end / train_N, train_N,
                                             dtype=np.float64)).to(self.delta_beta_half)
        alphas_cump = torch.cat((torch.ones(1, device=device), (1.0 - betas).cumprod(dim=0).to(device)))
        B = x.shape[0]

        def tvec(i):
            return timesteps[i].view(-1).expand(B).to(x)

        def transfer(x, t, t_next, et):
            ti = (train_N * (t[:1] - time_eps) + 1).long()
            tn = (train_N * (t_next[:1] - time_eps) + 1).long()
            at, at_next = alphas_cump[ti], alphas_cump[tn]
            coef = torch.cat([at_next - at,
                              1 / (at.sqrt() * (at.sqrt() + at_next.sqrt())),
                              1 / (at.sqrt() * (((1 - at_next) * at).sqrt() + ((1 - at) * at_next).sqrt()))]).contiguous()
            out = torch.empty_like(x)
            ops.pndm_transfer(x, et.contiguous(), coef, out)
            return out

        def eps_of(t, x):
            return score_fn(t, x, condition=condition, label=label)[1].contiguous()

        ets = []
        for idx in range(self.N, 0, -1):
            t_next = idx - 1
            t_list = [idx, (idx + t_next) / 2, t_next]
            if len(ets) > 2:
                ets.append(eps_of(tvec(idx * 2 - 1), x))
                noise = torch.empty_like(x)
                ops.lincomb4((55.0, -59.0, 37.0, -9.0), (ets[-1], ets[-2], ets[-3], ets[-4]), 1 / 24, noise)
            else:
                t1, t2, t3 = tvec(t_list[0] * 2 - 1), tvec(int(t_list[1] * 2) - 1), tvec(int(t_list[2] * 2) - 1)
                e1 = eps_of(t1, x)
                ets.append(e1)
                e2 = eps_of(t2, transfer(x, t1, t2, e1))
                e3 = eps_of(t2, transfer(x, t1, t2, e2))
                e4 = eps_of(t3, transfer(x, t1, t3, e3))
                noise = torch.empty_like(x)
                ops.lincomb4((1.0, 2.0, 2.0, 1.0), (e1, e2, e3, e4), 1 / 6, noise)
            x = transfer(x, tvec(idx * 2 - 1), tvec(t_next * 2 - 1), noise)
        return x

    def sample_discrete(self, score_fn, num_samples, N, predictor, corrector, corrector_steps, shape, time_eps,
                        probability_flow, denoise, snr, device, condition=None, label=None, print_steps=None):
        """Reverse-SDE sampling; signature and semantics of diffusion_continuous.py:133-338."""
        if predictor is not None and predictor != "pndm" and predictor not in _PRED_CODES:
            raise NotImplementedError("preditor not Implemented")
        if corrector is not None and corrector not in ("langevin", "ancestral"):
            raise NotImplementedError("corrector not Implemented")
        device = torch.device(device)
        with torch.no_grad():
            # initial sample from the CPU generator, then H2D (:237)
            x = torch.randn((num_samples,) + tuple(shape)).to(device)
            if predictor == "pndm":
                return self._pndm(score_fn, x, time_eps, condition, label)
            if predictor is None and corrector is None:
                if print_steps is not None:
                    steps = (N - 1) // (print_steps - 2)
                    return [x] * (1 + N // steps + 1)
                return x  # no predictor, no corrector: the loop is the identity (:243-249)

            from .sampler import fused_sample_loop, find_score_module  # late import (sampler imports Score)
            fusable = (predictor is not None and (corrector is None or corrector == "ancestral")
                       and not isinstance(condition, dict))
            owner_model = getattr(getattr(score_fn, "__self__", None), "model", None)
            if getattr(owner_model, "precision", "bf16") in ("tf32", "fp32") and (condition is not None or label is not None):
                fusable = False   # the TF32 parity mode samples conditionally through the per-step path
            score_mod = None
            if fusable:   # one real call of score_fn must reproduce what the fused step hard-wires (sampler.py)
                probe_t = torch.ones((num_samples,), device=device) * torch.linspace(1.0, time_eps, N, device=device)[0]
                score_mod = find_score_module(score_fn, self, probe=(probe_t, x, label, condition))
            if score_mod is not None:
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
                                         cond_tokens=cond_tokens, extra=extra, print_steps=print_steps,
                                         corrector_steps=corrector_steps if corrector is not None else 0, snr=snr)

            # generic path: arbitrary score_fn called once per (predictor | corrector) step, fused update kernel in between
            timesteps = torch.linspace(1.0, time_eps, N, device=device)
            zero = torch.zeros(1, dtype=torch.int32, device=device)
            if predictor is not None:
                coef, _ = self.step_coefficients(predictor, N, time_eps, probability_flow, device, raw_score=True)
                code = _PRED_CODES[predictor]
            if corrector == "ancestral":
                ccoef = self.corrector_coefficients(N, time_eps, snr, device, raw_score=True)
            x_mean = x
            out_list, steps = None, None
            if print_steps is not None:
                out_list = [x]
                steps = (N - 1) // (print_steps - 2)
            for i in range(N):
                vec_t = torch.ones((num_samples,), device=device) * timesteps[i]
                x_mean = x
                if predictor is not None:
                    score, params = score_fn(vec_t, x, label=label, condition=condition)
                    z = torch.randn_like(x)  # same CUDA-generator draw as the reference (:160)
                    src = (params if predictor == "ddim" else score).contiguous()
                    x_next, x_mean = torch.empty_like(x), torch.empty_like(x)
                    ops.sde_step(code, x.contiguous(), src, z, coef[i:i + 1], zero, 0, 0, 0, 0, x_next, x_mean)
                    x = x_next
                if corrector is not None:
                    for _ in range(corrector_steps):
                        grad, params = score_fn(vec_t, x, label=label, condition=condition)
                        grad = grad.contiguous()
                        noise = torch.randn_like(x)
                        if corrector == "langevin":
                            # (:193-210); the reference's `step_size[:, None] * grad` only broadcasts for batch sizes 1
                            # and 32 -- every entry of step_size is the same scalar, which is what is applied here
                            row = torch.zeros((1, SDE_COEF_STRIDE), dtype=torch.float32, device=device)
                            step_size = (snr * ops.batch_mean_norm(noise) / ops.batch_mean_norm(grad)) ** 2 * 2
                            row[0, 0], row[0, 1], row[0, 2] = -1.0, step_size, torch.sqrt(step_size * 2)
                        else:
                            row = ccoef[i:i + 1]
                        x_next, x_mean = torch.empty_like(x), torch.empty_like(x)
                        ops.sde_step(PRED_CORRECTOR, x.contiguous(), grad, noise, row, zero, 0, 0, 0, 0, x_next, x_mean)
                        x = x_next
                if out_list is not None and (i + 1) % steps == 0:
                    out_list.append(x_mean.clone())
            if out_list is not None:
                out_list.append(x_mean if denoise else x)
                return out_list
            return x_mean if denoise else x
