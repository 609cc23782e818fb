"""CPU oracle for the LDT sampling hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional restatement (plain torch-CPU tensor arithmetic over a ``state_dict``; float32 by default,
float64 on request) of the reference's algorithm for: the time embedding, the AdaLN transformer score
network, the Compressor decoder, the VP-SDE tables and discrete predictors, and the evaluation metrics
built on the Chamfer matrix.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; ``ldt_b200`` itself never does.

Parity pin: ``tests/test_oracle_golden.py`` checks every function here against golden vectors produced by
importing the UNMODIFIED reference from /root/reference (``tests/golden/make_golden.py``), so the oracle is
pinned to the reference's own outputs, and against the reference's only known-answer criterion
(ChamferDistancePytorch/unit_test.py:22-33) for the NN kernel.

Every function cites the reference lines it follows (paths relative to the reference root).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

LN_EPS = 1e-6  # tools/utils.py:130

# Optional operand-rounding hook.  With QUANT = None (default) this file is the plain fp32/fp64 restatement of the
# reference.  Tests may set QUANT = bf16_round to emulate WHERE the B200 path rounds operands to bf16 (GEMM inputs,
# stored q/k/v, softmax probabilities, MLP hidden) while keeping the reference's algorithm, to separate "bf16
# operand noise" from "wrong algorithm" when comparing against the fp32 reference.
QUANT = None


def bf16_round(x):
    return x.to(torch.bfloat16).to(x.dtype)


def tf32_round(x):
    """fp32 -> nearest TF32 value (10 mantissa bits, ties away from zero: PTX cvt.rna.tf32.f32), kept in fp32.  Tests set
    QUANT = tf32_round to emulate where the TF32 parity mode of the B200 path rounds its contraction operands."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def _q(x):
    return x if QUANT is None else QUANT(x)


# --------------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------------
def conv1x1(x, w, b):
    """nn.Conv1d(kernel_size=1) on channels-first [B, C, N]; w is [out, in, 1]."""
    return torch.einsum("oc,bcn->bon", _q(w[:, :, 0]), _q(x)) + b[None, :, None]


def layer_norm_cf(x, weight=None, bias=None):
    """tools/utils.py:127-133: nn.LayerNorm(C, eps=1e-6) over channels of a channels-first tensor."""
    xt = x.transpose(1, 2)
    y = F.layer_norm(xt, (xt.shape[-1],), weight, bias, LN_EPS)
    return y.transpose(1, 2)


def time_embedding(sd, t, prefix="TimeEmbedding."):
    """model/layers.py:14-41.  dim_embed = mlp.0.weight.shape[1]; t is the raw continuous time."""
    w0 = sd[prefix + "mlp.0.weight"]
    half = w0.shape[1] // 2
    freq = torch.exp(torch.arange(half) * -(np.log(10000) / (half - 1))).to(t.device)  # :28-30 (float32)
    arg = t.unsqueeze(1).to(torch.float32) * freq
    emb = torch.cat((torch.sin(arg), torch.cos(arg)), 1).to(w0.dtype)
    h = F.silu(F.linear(emb, w0, sd[prefix + "mlp.0.bias"]))
    return F.linear(h, sd[prefix + "mlp.2.weight"], sd[prefix + "mlp.2.bias"])


def time_freq(dim_embed: int) -> torch.Tensor:
    half = dim_embed // 2
    return torch.exp(torch.arange(half) * -(np.log(10000) / (half - 1)))


def attention(sd, prefix, x, y, num_heads):
    """ResidualBlock.compute_attention, model/layers.py:183-200, including the head-layout quirk at :197."""
    if y is None:
        y = x
    query = _q(conv1x1(x, sd[prefix + "fc_q.weight"], sd[prefix + "fc_q.bias"]))
    kv = _q(conv1x1(y, sd[prefix + "fc_kv.weight"], sd[prefix + "fc_kv.bias"]))
    B, Cc, N = query.shape
    key, value = kv[:, :Cc, :], kv[:, Cc:, :]
    M = key.shape[2]
    dh = Cc // num_heads
    q = query.reshape(B, num_heads, dh, N).permute(0, 1, 3, 2)
    k = key.reshape(B, num_heads, dh, M).permute(0, 1, 3, 2)
    v = value.reshape(B, num_heads, dh, M).permute(0, 1, 3, 2)
    w = (q @ k.transpose(-2, -1)) * (dh ** -0.5)
    w = w.softmax(dim=-1)
    if QUANT is not None:  # the kernel rounds un-normalised probabilities exp(s - max) and divides afterwards
        sc_ = (q @ k.transpose(-2, -1)) * (dh ** -0.5)
        e = torch.exp(sc_ - sc_.amax(-1, keepdim=True))
        w = _q(e) / e.sum(-1, keepdim=True)
    att = (w @ v).reshape(B, N, Cc).transpose(1, 2)  # :197 -- heads are NOT permuted back
    return conv1x1(att, sd[prefix + "fc_o.weight"], sd[prefix + "fc_o.bias"])


def mlp(sd, prefix, x):
    """MLP with n_hidden=1, exact-erf GELU, no residual: model/layers.py:110-133."""
    h = F.gelu(conv1x1(x, sd[prefix + "fc.0.0.weight"], sd[prefix + "fc.0.0.bias"]))
    return conv1x1(h, sd[prefix + "out.weight"], sd[prefix + "out.bias"])


def residual_block_adaln(sd, prefix, x, y, c, num_heads):
    """ResidualBlock.forward, AdaLN branch with dim_in == dim_out: model/layers.py:211-219."""
    cc = c[:, None, :] if c.dim() == 2 else c.transpose(1, 2)
    mod = F.linear(_q(F.silu(cc)), _q(sd[prefix + "adaLN.1.weight"]), sd[prefix + "adaLN.1.bias"]).transpose(1, 2)
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = mod.chunk(6, dim=1)
    h = layer_norm_cf(x) * (1 + scale_msa) + shift_msa
    x = x + gate_msa * attention(sd, prefix, h, y, num_heads)
    h = layer_norm_cf(x) * (1 + scale_mlp) + shift_mlp
    return x + gate_mlp * mlp(sd, prefix + "mlp.", h)


def residual_block_adaln_down(sd, prefix, x, y, c, num_heads):
    """ResidualBlock.forward, AdaLN branch with dim_in = 2*dim_out (UNet Down blocks): model/layers.py:153-155,
    173-175, 216-219 -- Conv1d shortcut on the raw input, adaLN1 -> (shift, scale) of the dim_in-wide norm, adaLN2 ->
    (gate_msa, shift_mlp, scale_mlp, gate_mlp)."""
    cc = c[:, None, :] if c.dim() == 2 else c.transpose(1, 2)
    m1 = F.linear(_q(F.silu(cc)), _q(sd[prefix + "adaLN1.1.weight"]), sd[prefix + "adaLN1.1.bias"]).transpose(1, 2)
    m2 = F.linear(_q(F.silu(cc)), _q(sd[prefix + "adaLN2.1.weight"]), sd[prefix + "adaLN2.1.bias"]).transpose(1, 2)
    shift_msa, scale_msa = m1.chunk(2, dim=1)
    gate_msa, shift_mlp, scale_mlp, gate_mlp = m2.chunk(4, dim=1)
    h = layer_norm_cf(x) * (1 + scale_msa) + shift_msa
    x = conv1x1(x, sd[prefix + "shortcut.weight"], sd[prefix + "shortcut.bias"]) + gate_msa * attention(sd, prefix, h, y, num_heads)
    h = layer_norm_cf(x) * (1 + scale_mlp) + shift_mlp
    return x + gate_mlp * mlp(sd, prefix + "mlp.", h)


def score_forward_unet(sd, cfg, x, t, cond_vec=None):
    """Score.forward with ``unet: True`` (model/scorenet/score.py:138-146), unconditional or image-vector only."""
    c = time_embedding(sd, t.to(x.dtype))
    if cond_vec is not None:
        c = c + cond_vec
    h = conv1x1(x.transpose(1, 2), sd["ln_in.weight"], sd["ln_in.bias"])
    n = cfg.num_blocks // 2
    saved = [h]
    for i in range(n):
        h = residual_block_adaln(sd, f"Transformer_Up.{i}.", h, None, c, cfg.num_heads)
        saved.append(h)
    h = residual_block_adaln(sd, "Transformer_Mid.", h, None, c, cfg.num_heads)
    for i in range(n):
        h = torch.cat((h, saved.pop()), dim=1)
        h = residual_block_adaln_down(sd, f"Transformer_Down.{i}.", h, None, c, cfg.num_heads)
    mod = F.linear(_q(F.silu(c[:, None, :])), _q(sd["ln_out.adaLN.1.weight"]), sd["ln_out.adaLN.1.bias"]).transpose(1, 2)
    shift, scale = mod.chunk(2, dim=1)
    h = layer_norm_cf(h) * (1 + scale) + shift
    return conv1x1(h, sd["ln_out.ln.weight"], sd["ln_out.ln.bias"]).transpose(1, 2)


def residual_block_plain(sd, prefix, x, y, num_heads):
    """ResidualBlock.forward with c=None and act=Identity (decoder): model/layers.py:224-226."""
    h = layer_norm_cf(x, sd[prefix + "norm1.norm.weight"], sd[prefix + "norm1.norm.bias"])
    x = x + attention(sd, prefix, h, y, num_heads)
    h = layer_norm_cf(x, sd[prefix + "norm2.norm.weight"], sd[prefix + "norm2.norm.bias"])
    return x + mlp(sd, prefix + "mlp.", h)


# --------------------------------------------------------------------------------------------------
# score network (model/scorenet/score.py:117-151, non-unet branch)
# --------------------------------------------------------------------------------------------------
def score_forward(sd, cfg, x, t, cond_tokens=None, cond_vec=None, return_blocks=False):
    """x [B, z_scale, z_dim], t [B] -> eps prediction [B, z_scale, z_dim].

    cond_tokens [B, hidden, z_scale] / cond_vec [B, t_dim] are the two outputs of ConditionNet
    (score.py:31-44) when present; even-indexed blocks cross-attend to cond_tokens (:148-149).
    """
    c = time_embedding(sd, t.to(x.dtype))
    if cond_vec is not None:
        c = c + cond_vec  # :135
    h = conv1x1(x.transpose(1, 2), sd["ln_in.weight"], sd["ln_in.bias"])  # :136-137
    blocks = []
    for i in range(cfg.num_blocks):
        y = cond_tokens if (i % 2 == 0) else None
        h = residual_block_adaln(sd, f"Transformer.{i}.", h, y, c, cfg.num_heads)
        if return_blocks:
            blocks.append(h)
    # FinalLayer, model/layers.py:240-245
    mod = F.linear(_q(F.silu(c[:, None, :])), _q(sd["ln_out.adaLN.1.weight"]), sd["ln_out.adaLN.1.bias"]).transpose(1, 2)
    shift, scale = mod.chunk(2, dim=1)
    h = layer_norm_cf(h) * (1 + scale) + shift
    out = conv1x1(h, sd["ln_out.ln.weight"], sd["ln_out.ln.bias"]).transpose(1, 2)
    return (out, blocks) if return_blocks else out


# --------------------------------------------------------------------------------------------------
# Compressor decoder (model/Compressor/Network.py:251-268, layers.py:26-37, ops.py:6-14)
# --------------------------------------------------------------------------------------------------
def sample_mask(batch: int, num_points: int, max_size: int) -> torch.Tensor:
    """ops.py:6-14 -- consumes `batch` CPU randperms exactly like the reference."""
    presence = [torch.randperm(max_size) < num_points for _ in range(batch)]
    return ~torch.stack(presence, dim=0)


def decoder_sample(sd, cfg, given_eps, num_points, mask=None):
    """given_eps [B, z_scales, n_layers*z_dim] -> points [B, num_points, 3]."""
    B = given_eps.shape[0]
    prior = sd["init_set.prior"]
    if mask is None:
        mask = sample_mask(B, num_points, cfg.max_outputs)
    xs = prior[None].expand(B, -1, -1)
    o = xs[~mask, :].view(B, num_points, prior.shape[1]).transpose(1, 2)  # layers.py:35-37
    eps_cf = given_eps.transpose(1, 2)
    chunks = torch.split(eps_cf, [cfg.z_dim] * cfg.n_layers, dim=1)  # Network.py:261-262
    for idx in range(cfg.n_layers):
        l = cfg.n_layers - 1 - idx  # reversed(self.decoder), :263
        p = f"decoder.{l}."
        xx = conv1x1(chunks[idx], sd[p + "ln.weight"], sd[p + "ln.bias"])  # DecoderBlock.forward :80-83
        o = residual_block_plain(sd, p + "att1.", o, xx, cfg.num_heads)
    o = conv1x1(o, sd["output.weight"], sd["output.bias"]).transpose(1, 2)  # :266
    return o  # postprocess is the identity for 3-D points (:273-274)


# --------------------------------------------------------------------------------------------------
# VP-SDE (diffusion/diffusion_continuous.py:626-678) and discrete predictors (:141-191)
# --------------------------------------------------------------------------------------------------
class VPSDE:
    def __init__(self, beta_start, beta_end, sigma2_0, N, dtype=torch.float32):
        self.beta_start, self.beta_end, self.sigma2_0, self.N = beta_start, beta_end, sigma2_0, N
        # :647-653 -- float64 linspace cast to float32
        self.betas = torch.from_numpy(np.linspace(beta_start / N, beta_end / N, N, dtype=np.float64)).to(dtype)
        self.alpha = 1.0 - self.betas
        self.alphas_cump = self.alpha.cumprod(dim=0)

    def g2(self, t):  # :658-659
        return self.beta_start + (self.beta_end - self.beta_start) * t

    def f(self, t):  # :655-656
        return -0.5 * self.g2(t)

    def var(self, t):  # :664-666
        return 1.0 - (1.0 - self.sigma2_0) * torch.exp(-self.beta_start * t - 0.5 * (self.beta_end - self.beta_start) * t * t)

    def e2int_f(self, t):  # :671-672
        return torch.exp(-0.5 * self.beta_start * t - 0.25 * (self.beta_end - self.beta_start) * t * t)


def score_from_params(sde, params, t):
    """Trainer.score_fn, trainer/Latent_SDE_Trainer.py:57-61."""
    return -params / torch.sqrt(sde.var(t))[:, None, None]


def ancestral_step(sde, x, t, params, noise, N):
    """diffusion_continuous.py:152-162."""
    idx = (t * (N - 1) / 1.0).long()
    beta = sde.betas[idx]
    score = score_from_params(sde, params, t)
    x_mean = (x + beta[:, None, None] * score) / torch.sqrt(1.0 - beta)[:, None, None]
    return x_mean + torch.sqrt(beta)[:, None, None] * noise, x_mean


def reverse_diffusion_step(sde, x, t, params, noise, N, time_eps, probability_flow=False):
    """:141-150."""
    dt = torch.tensor((1 - time_eps) / N)
    f, g2 = sde.f(t)[:, None, None] * x, sde.g2(t)[:, None, None]
    score = score_from_params(sde, params, t)
    dx = (f - g2 * score * (0.5 if probability_flow else 1.0)) * dt
    g = torch.zeros_like(g2) if probability_flow else torch.sqrt(g2)
    x_mean = x - dx
    return x_mean + g * noise * torch.sqrt(dt), x_mean


def euler_maruyama_step(sde, x, t, params, noise, N, probability_flow=False):
    """:182-191."""
    dt = -1.0 / N
    f, g2 = sde.f(t)[:, None, None] * x, sde.g2(t)[:, None, None]
    score = score_from_params(sde, params, t)
    f = f - g2 * score * (0.5 if probability_flow else 1.0)
    x_mean = x + f * dt
    g2 = torch.zeros(1).to(x) if probability_flow else g2
    return x_mean + torch.sqrt(g2) * np.sqrt(-dt) * noise, x_mean


def ddim_step(sde, x, t, params, N):
    """:164-180 (sigma = 0)."""
    idx = (t * (N - 1) / 1.0).long()
    at = sde.alphas_cump[idx][:, None, None]
    at_next = torch.ones_like(at) if idx[0] - 1 < 0 else sde.alphas_cump[idx - 1][:, None, None]
    x_mean = at_next.sqrt() * (x - (1 - at).sqrt() * params) / at.sqrt() + (1 - at_next).sqrt() * params
    return x_mean, x_mean


def ancestral_corrector_step(sde, x, t, params, noise, snr):
    """AncestralCorrector, :212-229.  alpha = 1: `self.__class__ in ["DiffusionVPSDE", ...]` is never true."""
    grad = score_from_params(sde, params, t)
    alpha = torch.ones_like(t)
    step_size = (snr * torch.sqrt(sde.var(t))) ** 2 * 2 * alpha
    x_mean = x + step_size[:, None, None] * grad
    return x_mean + noise * torch.sqrt(step_size * 2)[:, None, None], x_mean


def langevin_corrector_step(sde, x, t, params, noise, snr):
    """LangevinCorrector, :193-210, with the scalar the reference's `step_size[:, None]` broadcast stands for."""
    grad = score_from_params(sde, params, t)
    grad_norm = torch.norm(grad.reshape(grad.shape[0], -1), dim=-1).mean()
    noise_norm = torch.norm(noise.reshape(noise.shape[0], -1), dim=-1).mean()
    step_size = (snr * noise_norm / grad_norm) ** 2 * 2
    x_mean = x + step_size * grad
    return x_mean + torch.sqrt(step_size * 2) * noise, x_mean


def sample_discrete(sde, score_net, x0, N, time_eps, noises, predictor="ancestral", denoise=True, corrector=None,
                    corrector_steps=1, snr=0.01, print_steps=None):
    """pc_sampling loop, :231-258, with the noise supplied by the caller in draw order (noises[j] ~ randn_like)."""
    x = x0
    timesteps = torch.linspace(1.0, time_eps, N)
    x_mean = x
    it = iter(noises)
    out_list = [x] if print_steps is not None else None
    steps = (N - 1) // (print_steps - 2) if print_steps is not None else None
    for i in range(N):
        vec_t = torch.ones((x.shape[0],)) * timesteps[i]
        x_mean = x
        if predictor is not None:
            params = score_net(x, vec_t)
            if predictor == "ancestral":
                x, x_mean = ancestral_step(sde, x, vec_t, params, next(it), N)
            elif predictor == "reversediffusion":
                x, x_mean = reverse_diffusion_step(sde, x, vec_t, params, next(it), N, time_eps)
            elif predictor == "eulermaruyama":
                x, x_mean = euler_maruyama_step(sde, x, vec_t, params, next(it), N)
            elif predictor == "ddim":
                next(it)  # DDIM draws a noise it multiplies by sigma = 0 (:178-179)
                x, x_mean = ddim_step(sde, x, vec_t, params, N)
            else:
                raise NotImplementedError(predictor)
        if corrector is not None:
            for _ in range(corrector_steps):
                params = score_net(x, vec_t)
                fn = {"ancestral": ancestral_corrector_step, "langevin": langevin_corrector_step}[corrector]
                x, x_mean = fn(sde, x, vec_t, params, next(it), snr)
        if out_list is not None and (i + 1) % steps == 0:
            out_list.append(x_mean)
    if out_list is not None:
        out_list.append(x_mean if denoise else x)
        return out_list
    return x_mean if denoise else x


def pndm_sample(sde, score_net, x0, sample_N, train_N, time_eps):
    """predictor == "pndm", :260-316: Runge-Kutta warm-up for the first three steps, then 4-term linear multistep.
    Index arithmetic as written in the reference (timesteps[-1] on the last step included)."""
    timesteps = torch.linspace(time_eps, 1.0, sample_N * 2)
    betas = torch.from_numpy(np.linspace(sde.beta_start / train_N, sde.beta_end / train_N, train_N, dtype=np.float64)).float()
    alphas_cump = torch.cat((torch.ones(1), (1.0 - betas).cumprod(dim=0)))
    B = x0.shape[0]

    def tv(i):
        return timesteps[i].view(-1).expand(B).to(x0)

    def transfer(x, t, t_next, et):
        at = alphas_cump[(train_N * (t - time_eps) + 1).long()][0]
        at_next = alphas_cump[(train_N * (t_next - time_eps) + 1).long()][0]
        x_delta = (at_next - at) * ((1 / (at.sqrt() * (at.sqrt() + at_next.sqrt()))) * x - 1 / (at.sqrt() * (
            ((1 - at_next) * at).sqrt() + ((1 - at) * at_next).sqrt())) * et)
        return x + x_delta

    x, ets = x0, []
    for idx in range(sample_N, 0, -1):
        t_next = idx - 1
        t_list = [idx, (idx + t_next) / 2, t_next]
        if len(ets) > 2:
            ets.append(score_net(x, tv(idx * 2 - 1)))
            noise = (1 / 24) * (55 * ets[-1] - 59 * ets[-2] + 37 * ets[-3] - 9 * ets[-4])
        else:
            t1, t2, t3 = tv(t_list[0] * 2 - 1), tv(int(t_list[1] * 2) - 1), tv(int(t_list[2] * 2) - 1)
            e1 = score_net(x, t1)
            ets.append(e1)
            e2 = score_net(transfer(x, t1, t2, e1), t2)
            e3 = score_net(transfer(x, t1, t2, e2), t2)
            e4 = score_net(transfer(x, t1, t3, e3), t3)
            noise = (1 / 6) * (e1 + 2 * e2 + 2 * e3 + e4)
        x = transfer(x, tv(idx * 2 - 1), tv(t_next * 2 - 1), noise)
    return x


def nn_distance_f64(a, b):
    """ChamferDistancePytorch/chamfer_python.py:18-39 style brute force in float64 (the check of
    unit_test.py:22-33): returns dist1, idx1, dist2, idx2."""
    a64, b64 = a.double(), b.double()
    d = ((a64[:, :, None, :] - b64[:, None, :, :]) ** 2).sum(-1)
    d1, i1 = d.min(2)
    d2, i2 = d.min(1)
    return d1, i1, d2, i2


def pairwise_cd(a, b, nn_fn):
    """_pairwise_CD_, evaluation_metrics.py:165-198: M[i,j] = mean(dl)+mean(dr) for cloud i of a, j of b."""
    rows = []
    for i in range(a.shape[0]):
        ai = a[i].view(1, -1, 3).expand(b.shape[0], -1, -1).contiguous()
        dl, dr = nn_fn(ai, b)
        rows.append((dl.mean(dim=1) + dr.mean(dim=1)).view(1, -1))
    return torch.cat(rows, dim=0)


def lgan_mmd_cov(all_dist):
    """evaluation_metrics.py:234-246."""
    n_ref = all_dist.size(1)
    _, min_idx = torch.min(all_dist, dim=1)
    min_val, _ = torch.min(all_dist, dim=0)
    return {"mmd": min_val.mean(), "cov": torch.tensor(float(min_idx.unique().view(-1).size(0)) / float(n_ref))}


def knn_1nna(Mxx, Mxy, Myy, k=1):
    """evaluation_metrics.py:202-231 (sqrt=False)."""
    n0, n1 = Mxx.size(0), Myy.size(0)
    label = torch.cat((torch.ones(n0), torch.zeros(n1))).to(Mxx)
    M = torch.cat((torch.cat((Mxx, Mxy), 1), torch.cat((Mxy.transpose(0, 1), Myy), 1)), 0)
    val, idx = (M + torch.diag(float("inf") * torch.ones(n0 + n1).to(Mxx))).topk(k, 0, False)
    count = torch.zeros(n0 + n1).to(Mxx)
    for i in range(k):
        count = count + label.index_select(0, idx[i])
    pred = torch.ge(count, (float(k) / 2) * torch.ones(n0 + n1).to(Mxx)).float()
    s = {"tp": (pred * label).sum(), "fp": (pred * (1 - label)).sum(), "fn": ((1 - pred) * label).sum(),
         "tn": ((1 - pred) * (1 - label)).sum()}
    s.update({"precision": s["tp"] / (s["tp"] + s["fp"] + 1e-10), "recall": s["tp"] / (s["tp"] + s["fn"] + 1e-10),
              "acc": torch.eq(label, pred).float().mean()})
    return s


def compute_cd_metrics(sample_pcs, ref_pcs, nn_fn):
    """compute_CD_metrics, evaluation_metrics.py:299-318."""
    res = {}
    M_rs = pairwise_cd(ref_pcs, sample_pcs, nn_fn)
    res.update({f"{k}-CD": v for k, v in lgan_mmd_cov(M_rs.t()).items()})
    M_rr = pairwise_cd(ref_pcs, ref_pcs, nn_fn)
    M_ss = pairwise_cd(sample_pcs, sample_pcs, nn_fn)
    res.update({f"1-NN-CD-{k}": v for k, v in knn_1nna(M_rr, M_rs, M_ss, 1).items() if "acc" in k})
    return res


# --------------------------------------------------------------------------------------------------
# deterministic synthetic weights shared by the golden generator and the tests
# --------------------------------------------------------------------------------------------------
def synth_state_dict(shapes: dict, seed: int, gain: float = 1.5) -> dict:
    """Fill {name: shape} with seeded values in sorted-key order (independent of module registration order).

    Conv/Linear weights ~ N(0, 1/fan_in) * gain (1.5 by default; the deep stochastic encoder test uses a smaller gain so
    that the map stays well conditioned), biases ~ N(0, 0.1), norm weights ~ 1 + N(0, 0.1), the decoder
    prior ~ U(0,1) like its initialiser (Compressor/layers.py:24).  Same torch build => same bits everywhere.
    """
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros(shape, dtype=torch.long)
        elif name.endswith("running_var") or name.endswith("initialized"):
            out[name] = torch.ones(shape)
        elif name.endswith("running_mean"):
            out[name] = torch.zeros(shape)
        elif name.endswith("prior"):
            out[name] = torch.rand(shape, generator=g)
        elif "norm" in name and name.endswith("weight") and len(shape) == 1:
            out[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("weight") and len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            out[name] = torch.randn(shape, generator=g) * (gain / math.sqrt(fan_in))
        else:
            out[name] = 0.1 * torch.randn(shape, generator=g)
    return out


def cfg_ns(**kw):
    return SimpleNamespace(**kw)


# --------------------------------------------------------------------------------------------------
# completion prologue: ConditionNet (model/scorenet/score.py:13-44) over LocalGrouper
# (model/Compressor/layers.py:288-319).  `fps_fn(xyz [B,N,3], m) -> long [B,m]` stands in for the un-vendored
# pointnet2_ops call (tests pass oracle/pointops_oracle.c::oracle_fps through ctypes).
# --------------------------------------------------------------------------------------------------
def index_points(points, idx):
    """points [B,N,C], idx [B,...] -> [B,...,C]  (Compressor/layers.py:46-62)."""
    B = points.shape[0]
    batch = torch.arange(B).view(B, *([1] * (idx.dim() - 1))).expand_as(idx)
    return points[batch, idx, :]


def knn_point(k, xyz, new_xyz):
    """Indices [B,S,k] of the k nearest points, by the reference's expanded distance (layers.py:65-98)."""
    d = -2 * torch.matmul(new_xyz, xyz.permute(0, 2, 1))
    d = d + torch.sum(new_xyz ** 2, -1)[:, :, None] + torch.sum(xyz ** 2, -1)[:, None, :]
    return torch.topk(d, k, dim=-1, largest=False, sorted=False)[1]


def _bn_eval(sd, prefix, x):
    return F.batch_norm(x, sd[prefix + "running_mean"], sd[prefix + "running_var"], sd[prefix + "weight"],
                        sd[prefix + "bias"], training=False, eps=1e-5)


def local_grouper(sd, prefix, xyz, feature, groups, k, normalize, fps_fn, use_xyz=True):
    """xyz [B,3,N], feature [B,D,N] -> (new_xyz [B,3,S], [B,D,S]); eval-mode BatchNorm (layers.py:288-319,163-192)."""
    pts, fea = xyz.transpose(1, 2), feature.transpose(1, 2)
    B = pts.shape[0]
    fps_idx = fps_fn(pts.contiguous(), groups)
    new_xyz = index_points(pts, fps_idx)
    idx = knn_point(k, pts, new_xyz)
    new_feature = index_points(fea, fps_idx)
    grouped = index_points(fea, idx)
    if use_xyz:
        grouped = torch.cat([grouped, index_points(pts, idx)], dim=-1)
    if normalize == "center":
        mean = grouped.mean(dim=2, keepdim=True)
    else:  # anchor
        mean = (torch.cat([new_feature, new_xyz], dim=-1) if use_xyz else new_feature).unsqueeze(-2)
    std = torch.std((grouped - mean).reshape(B, -1), dim=-1, keepdim=True)[:, :, None, None]
    grouped = sd[prefix + "affine_alpha"] * ((grouped - mean) / (std + 1e-5)) + sd[prefix + "affine_beta"]
    x = torch.cat([grouped, new_feature[:, :, None, :].expand(-1, -1, k, -1)], dim=-1)
    b, s, kk, d = x.shape
    x = x.permute(0, 1, 3, 2).reshape(-1, d, kk)
    e = prefix + "extraction."
    x = F.relu(_bn_eval(sd, e + "transfer.net.1.", F.conv1d(x, sd[e + "transfer.net.0.weight"], sd[e + "transfer.net.0.bias"])))
    o = e + "operation.0."
    y = F.relu(_bn_eval(sd, o + "net1.1.", F.conv1d(x, sd[o + "net1.0.weight"], sd[o + "net1.0.bias"])))
    x = F.relu(F.conv1d(y, sd[o + "net2.0.weight"], sd[o + "net2.0.bias"]) + x)
    x = x.max(dim=-1)[0].reshape(b, s, -1).permute(0, 2, 1)
    return new_xyz.transpose(1, 2), x


def _basic_block(sd, p, x, stride):
    """torchvision BasicBlock in eval mode (the ResNet18 trunk of score.py:24-25)."""
    y = F.relu(_bn_eval(sd, p + "bn1.", F.conv2d(x, sd[p + "conv1.weight"], None, stride=stride, padding=1)))
    y = _bn_eval(sd, p + "bn2.", F.conv2d(y, sd[p + "conv2.weight"], None, stride=1, padding=1))
    if (p + "downsample.0.weight") in sd:
        x = _bn_eval(sd, p + "downsample.1.", F.conv2d(x, sd[p + "downsample.0.weight"], None, stride=stride))
    return F.relu(y + x)


def condition_net(sd, prefix, img, pts, patch_size, fps_fn):
    """(pts_cond [B,hidden,patch_size], img_cond [B,p_dim])  (model/scorenet/score.py:31-44)."""
    r = prefix + "resnet."
    x = F.relu(_bn_eval(sd, r + "1.", F.conv2d(img, sd[r + "0.weight"], None, stride=2, padding=3)))
    x = F.max_pool2d(x, 3, stride=2, padding=1)
    x = _basic_block(sd, r + "4.0.", x, 1)
    x = _basic_block(sd, r + "4.1.", x, 1)
    x = _basic_block(sd, r + "5.0.", x, 2)
    x = _basic_block(sd, r + "5.1.", x, 1)
    img_cond = F.linear(F.adaptive_max_pool2d(x, 1).squeeze(), sd[prefix + "ln.weight"], sd[prefix + "ln.bias"])
    p = pts.transpose(1, 2)
    f = F.conv1d(p, sd[prefix + "pc_conv_in.weight"], sd[prefix + "pc_conv_in.bias"])
    _, f = local_grouper(sd, prefix + "group.", p, f, patch_size, f.shape[1] // patch_size * 2, "center", fps_fn)
    pts_cond = F.conv1d(f, sd[prefix + "pc_conv_out.weight"], sd[prefix + "pc_conv_out.bias"])
    return pts_cond, img_cond


# --------------------------------------------------------------------------------------------------
# Compressor bidirectional inference: bottom_up + top_down (model/Compressor/Network.py:188-249)
# --------------------------------------------------------------------------------------------------
def mini_pointnet(sd, prefix, x):
    """MiniPointnet.forward in eval mode (Network.py:86-101): [B,3,S] -> [B, p_dim]."""
    x = F.relu(_bn_eval(sd, prefix + "bn1.", F.conv1d(x, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"])))
    x = F.relu(_bn_eval(sd, prefix + "bn2.", F.conv1d(x, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"])))
    return F.linear(x.max(dim=2)[0], sd[prefix + "fc.weight"], sd[prefix + "fc.bias"])


def compressor_forward(sd, cfg, pts, fps_fn, num_points=None):
    """Compressor.forward (eval mode, no label): pts [B,N,3] -> dict like the reference's.  Consumes the CPU generator
    in the reference's order: B randperms (InitialSet), then one randn(mu.shape) per decoder layer (sample(), :26-29)."""
    H = cfg.hidden_dim
    x = pts.transpose(1, 2)
    f = F.conv1d(x, sd["input.weight"], sd["input.bias"])
    center, g = local_grouper(sd, "group.", x, f, cfg.z_scales, x.shape[2] // cfg.z_scales * 2, cfg.cluster_norm.lower(), fps_fn)
    pos = mini_pointnet(sd, "pos_embedding.", center)
    if cfg.ActNorm is not None:  # ActNorm.forward, model/layers.py:103-107
        g = ((g.transpose(1, 2) - sd["conv_in.shift"]) * torch.exp(-sd["conv_in.log_scale"])).transpose(1, 2)
    outputs = []
    for i in range(cfg.n_layers):  # Encoder.forward, Network.py:41-45: layer(x, x, pos) -> K/V from the raw x
        for j in range(cfg.encoder_layers):
            g = residual_block_adaln(sd, f"encoder.{i}.atts.{j}.", g, g, pos, cfg.num_heads)
        p = f"encoder.{i}.conv_out."
        mod = F.linear(_q(F.silu(pos[:, None, :])), _q(sd[p + "adaLN.1.weight"]), sd[p + "adaLN.1.bias"]).transpose(1, 2)
        shift, scale = mod.chunk(2, dim=1)
        outputs.append(conv1x1(layer_norm_cf(g) * (1 + scale) + shift, sd[p + "ln.weight"], sd[p + "ln.bias"]))
    # top_down, :211-233
    B = pts.shape[0]
    N = cfg.outsize if num_points is None else num_points
    prior = sd["init_set.prior"]
    mask = sample_mask(B, N, cfg.max_outputs)
    o = prior[None].expand(B, -1, -1)[~mask, :].view(B, N, H).transpose(1, 2)
    res = {"posteriors": [(o, None, None)], "kls": [], "all_eps": [], "all_logqz": []}
    for idx in range(cfg.n_layers):
        l = cfg.n_layers - 1 - idx
        p = f"decoder.{l}."
        xe = outputs[l]
        h = residual_block_plain(sd, p + "att.", xe, o if idx != 0 else xe, cfg.num_heads)   # compute_posterior :62-77
        post = conv1x1(F.silu(h), sd[p + "prior.1.weight"], sd[p + "prior.1.bias"])
        mu = post[:, :post.shape[1] // 2, :]
        logvar = post[:, post.shape[1] // 2:, :].clamp(cfg.min_sigma, 10.0)
        eps = mu + torch.exp(logvar / 2.0) * torch.randn(mu.shape).to(mu)
        logqz = -0.5 * torch.square(eps - mu) / torch.exp(logvar) - 0.5 * logvar - 0.9189385332
        logpz = -0.5 * torch.square(eps) - 0.9189385332
        xx = conv1x1(eps, sd[p + "ln.weight"], sd[p + "ln.bias"])
        o = residual_block_plain(sd, p + "att1.", o, xx, cfg.num_heads)
        res["all_eps"].append(eps)
        res["posteriors"].append((eps, mu, logvar))
        res["kls"].append(logqz - logpz)
        res["all_logqz"].append(logqz)
    res["set"] = conv1x1(o, sd["output.weight"], sd["output.bias"]).transpose(1, 2)
    res["all_eps"] = torch.cat(res["all_eps"], dim=1).transpose(1, 2)
    res["max"] = g.max()
    return res
