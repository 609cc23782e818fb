"""Host mirror of the reference ``Compressor`` for the sampling path (reference
model/Compressor/Network.py:105-279).

``Compressor(cfg.compressor)`` takes the same config keys, exposes ``sample(shape, given_eps=None)``,
``init()``, ``postprocess`` and a ``state_dict`` with exactly the reference's keys and shapes -- including the
encoder / grouper / position-embedding parameters that sampling never touches -- so stage-1 checkpoints
load with ``strict=True`` (trainer/Latent_SDE_Trainer.py:269-273).  Only the decoder is executed, on the
sm_100a kernels: per layer  Conv1d(z_dim->hidden) on the 32 latent tokens, K/V projection, LayerNorm(affine),
Q projection of the 2048 query rows, 2048x32 cross-attention, output projection + residual, LayerNorm, MLP with
GELU epilogue + residual (DecoderBlock.forward :80-83 -> ResidualBlock.forward c=None branch,
model/layers.py:224-226); finally Conv1d(hidden->3).  The encoder (bottom_up/top_down/forward) is training /
reconstruction only and is out of scope: ``forward`` raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from ._lib import EPI_BIAS_BF16, EPI_BIAS_F32, EPI_BIAS_GELU_BF16, EPI_GATE_RESID_F32
from .score import _PtrView, _pad_to


def _register_tree(root: nn.Module, spec: dict) -> None:
    """Create nested holder modules so that root.state_dict() has exactly the keys of ``spec``.

    spec: dotted name -> (shape, kind) with kind in {"param", "buffer", "long_buffer"}.
    """
    for name, (shape, kind) in spec.items():
        parts = name.split(".")
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, nn.Module())
            mod = mod._modules[p]
        if kind == "param":
            mod.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape)))
        elif kind == "long_buffer":
            mod.register_buffer(parts[-1], torch.zeros(shape, dtype=torch.long))
        else:
            mod.register_buffer(parts[-1], torch.zeros(shape))


def _conv(spec, name, cout, cin):
    spec[name + ".weight"] = ((cout, cin, 1), "param")
    spec[name + ".bias"] = ((cout,), "param")


def _linear(spec, name, cout, cin):
    spec[name + ".weight"] = ((cout, cin), "param")
    spec[name + ".bias"] = ((cout,), "param")


def _bn(spec, name, c):
    spec[name + ".weight"] = ((c,), "param")
    spec[name + ".bias"] = ((c,), "param")
    spec[name + ".running_mean"] = ((c,), "buffer")
    spec[name + ".running_var"] = ((c,), "buffer")
    spec[name + ".num_batches_tracked"] = ((), "long_buffer")


def _resblock(spec, name, dim, dim_c, mlp_ratio):
    """ResidualBlock(dim, dim, dim_c) parameters (model/layers.py:140-181) with dim_in == dim_out."""
    _conv(spec, name + ".fc_q", dim, dim)
    _conv(spec, name + ".fc_kv", 2 * dim, dim)
    _conv(spec, name + ".fc_o", dim, dim)
    if dim_c is None:
        for n in ("norm1", "norm2"):
            spec[f"{name}.{n}.norm.weight"] = ((dim,), "param")
            spec[f"{name}.{n}.norm.bias"] = ((dim,), "param")
    else:
        _linear(spec, name + ".adaLN.1", 6 * dim, dim_c)
    hid = int(mlp_ratio * dim)
    _conv(spec, name + ".mlp.fc.0.0", hid, dim)
    _conv(spec, name + ".mlp.out", dim, hid)


def _grouper(spec, name, dim, normalize):
    """LocalGrouper(dim, use_xyz=True, normalize) parameters (model/Compressor/layers.py:264-283, 178-199)."""
    if normalize is not None and str(normalize).lower() in ("center", "anchor"):
        spec[name + ".affine_alpha"] = ((1, 1, 1, dim + 3), "param")
        spec[name + ".affine_beta"] = ((1, 1, 1, dim + 3), "param")
    _conv(spec, name + ".extraction.transfer.net.0", dim, 3 + 2 * dim)
    _bn(spec, name + ".extraction.transfer.net.1", dim)
    _conv(spec, name + ".extraction.operation.0.net1.0", dim, dim)
    _bn(spec, name + ".extraction.operation.0.net1.1", dim)
    _conv(spec, name + ".extraction.operation.0.net2.0", dim, dim)


def compressor_param_spec(cfg) -> dict:
    """Every state_dict entry of reference Compressor(cfg) (Network.py:106-159), in its registration order."""
    H, P = cfg.hidden_dim, cfg.p_dim
    spec: dict = {}
    _conv(spec, "input", H, cfg.input_dim)
    if cfg.ActNorm is not None:
        shp = (1, 1, H) if cfg.ActNorm == "set" else (1, cfg.z_scales, H)
        spec["conv_in.shift"] = (shp, "param")
        spec["conv_in.log_scale"] = (shp, "param")
        spec["conv_in.initialized"] = ((1,), "buffer")
    label_dim = P if cfg.class_condition else None
    for i in range(cfg.n_layers):
        for j in range(cfg.encoder_layers):
            _resblock(spec, f"encoder.{i}.atts.{j}", H, P, cfg.mlp_ratio)
        _linear(spec, f"encoder.{i}.conv_out.adaLN.1", 2 * H, P)
        _conv(spec, f"encoder.{i}.conv_out.ln", H, H)
    for i in range(cfg.n_layers):
        _resblock(spec, f"decoder.{i}.att", H, label_dim, cfg.mlp_ratio)
        _conv(spec, f"decoder.{i}.prior.1", 2 * cfg.z_dim, H)
        _resblock(spec, f"decoder.{i}.att1", H, label_dim, cfg.mlp_ratio)
        _conv(spec, f"decoder.{i}.ln", H, cfg.z_dim)
    _grouper(spec, "group", H, cfg.cluster_norm)
    if cfg.pos_embedding == "mlp":
        _conv(spec, "pos_embedding.fc.0.0", P, 3)
        _conv(spec, "pos_embedding.out", P, P)
    else:  # MiniPointnet(3, p_dim), Network.py:86-101
        _conv(spec, "pos_embedding.conv1", 128, 3)
        _conv(spec, "pos_embedding.conv2", 256, 128)
        _bn(spec, "pos_embedding.bn1", 128)
        _bn(spec, "pos_embedding.bn2", 256)
        _linear(spec, "pos_embedding.fc", P, 256)
    if cfg.class_condition:
        spec["LabelEmbedding.label_emb.weight"] = ((cfg.num_categorys, P), "param")
        _linear(spec, "LabelEmbedding.mlp.0", P, P)
        _linear(spec, "LabelEmbedding.mlp.2", P, P)
    _conv(spec, "output", 3, H)
    if cfg.max_outputs is not None:
        spec["init_set.prior"] = ((cfg.max_outputs, H), "param")
    if cfg.pre_group:
        _grouper(spec, "pre_grouper", H, cfg.cluster_norm)
    return spec


class Compressor(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.input_dim = cfg.input_dim
        self.max_outputs = cfg.max_outputs
        self.n_layers = cfg.n_layers
        self.z_dim = cfg.z_dim
        self.hidden_dim = cfg.hidden_dim
        self.num_heads = cfg.num_heads
        self.norm = cfg.norm
        self.z_scales = cfg.z_scales
        self.p_dim = cfg.p_dim
        self.ActNorm = cfg.ActNorm
        self.outsize = cfg.outsize
        self.mlp_ratio = cfg.mlp_ratio
        self.decoder_act = cfg.decoder_act
        self.class_condition = cfg.class_condition
        if cfg.max_outputs is None:
            raise NotImplementedError("ldt_b200.Compressor: the mixture-of-Gaussians InitialSet (max_outputs: null) is not supported")
        if cfg.decoder_act is not None:
            raise NotImplementedError("ldt_b200.Compressor: decoder_act must be null (identity), as in the shipped configs")
        if cfg.class_condition:
            raise NotImplementedError("ldt_b200.Compressor: class-conditional decoder blocks are not supported")
        if cfg.norm != "layer_norm" or cfg.decoder_dropout_p != 0:
            raise NotImplementedError("ldt_b200.Compressor: needs norm: layer_norm and decoder_dropout_p: 0")
        if cfg.hidden_dim % 128 != 0 or cfg.hidden_dim // cfg.num_heads not in (32, 64) or cfg.z_scales != 32:
            raise NotImplementedError("ldt_b200.Compressor: needs hidden_dim % 128 == 0, head dim 32 or 64, z_scales 32")
        _register_tree(self, compressor_param_spec(cfg))
        self.reset_parameters()
        self._packed = None
        self._packed_key = None

    def reset_parameters(self):
        """Own initialiser (the reference relies on torch's per-layer defaults; checkpoints overwrite either)."""
        with torch.no_grad():
            for name, p in self.named_parameters():
                if name.endswith("prior"):
                    p.uniform_(0.0, 1.0)  # Compressor/layers.py:24
                elif name.endswith("affine_alpha") or (p.dim() == 1 and "norm" in name and name.endswith("weight")) \
                        or (".bn" in name and name.endswith("weight")) or name.endswith("net.1.weight") \
                        or name.endswith("net1.1.weight"):
                    p.fill_(1.0)
                elif p.dim() >= 2 and name.endswith("weight"):
                    fan_in = p[0].numel()
                    bound = (1.0 / fan_in) ** 0.5
                    p.uniform_(-bound, bound)
                elif name.endswith("bias") and p.dim() == 1:
                    p.uniform_(-0.05, 0.05)
            for name, b in self.named_buffers():
                if name.endswith("running_var"):
                    b.fill_(1.0)

    def init(self):
        """Network.py:163-165 -- marks ActNorm as initialised."""
        if self.ActNorm is not None:
            self.conv_in.initialized += 1.0

    def forward(self, x, num_points=None, label=None):
        raise NotImplementedError(
            "ldt_b200.Compressor implements the sampling decoder only; the encoder (bottom_up/top_down) is "
            "training/reconstruction code outside the B200 hot path (SURVEY.md section 2, row 4)")

    # ------------------------------------------------------------------------------------------
    def _fingerprint(self):
        return tuple((p.data_ptr(), p._version) for n, p in self.named_parameters()
                     if n.startswith(("decoder.", "output.", "init_set.")))

    def packed(self):
        key = self._fingerprint()
        if self._packed is not None and key == self._packed_key:
            return self._packed
        dev = self.output.weight.device
        if dev.type != "cuda":
            raise RuntimeError("ldt_b200.Compressor runs on CUDA only (no CPU fallback); call .to('cuda') first")
        H = self.hidden_dim
        P = {"layers": []}
        with torch.no_grad():
            for l in range(self.n_layers):
                d = self.decoder._modules[str(l)]
                a = d.att1
                f32 = lambda t: t.detach().float().contiguous()
                P["layers"].append({
                    "w_ln": ops.pack_weight(d.ln.weight), "b_ln": f32(d.ln.bias),
                    "w_q": ops.pack_weight(a.fc_q.weight), "b_q": f32(a.fc_q.bias),
                    "w_kv": ops.pack_weight(a.fc_kv.weight), "b_kv": f32(a.fc_kv.bias),
                    "w_o": ops.pack_weight(a.fc_o.weight), "b_o": f32(a.fc_o.bias),
                    "n1w": f32(a.norm1.norm.weight), "n1b": f32(a.norm1.norm.bias),
                    "n2w": f32(a.norm2.norm.weight), "n2b": f32(a.norm2.norm.bias),
                    "w_fc1": ops.pack_weight(a.mlp.fc._modules["0"]._modules["0"].weight),
                    "b_fc1": f32(a.mlp.fc._modules["0"]._modules["0"].bias),
                    "w_fc2": ops.pack_weight(a.mlp.out.weight), "b_fc2": f32(a.mlp.out.bias),
                })
            # output Conv1d(hidden -> 3): N padded to 8 rows of zeros so the GEMM core's N % 8 rule holds
            w_out = torch.zeros((8, H), dtype=torch.float32, device=dev)
            w_out[:3] = self.output.weight.detach().reshape(3, H)
            b_out = torch.zeros(8, dtype=torch.float32, device=dev)
            b_out[:3] = self.output.bias.detach()
            P["w_out"], P["b_out"] = ops.pack_weight(w_out), b_out
        self._packed, self._packed_key = P, key
        return P

    def initial_set(self, B, num_points):
        """InitialSet.forward with max_outputs set (Compressor/layers.py:26-37): per sample a random subset of the
        learned prior rows, ascending order kept.  Draws B CPU randperms exactly like ops.py:12 so the CPU
        generator ends where the reference's would.  Returns token-major f32 [B*num_points, hidden]."""
        prior = self.init_set.prior.detach().float()
        presence = [torch.randperm(self.max_outputs) < num_points for _ in range(B)]
        if num_points == self.max_outputs:
            return prior.unsqueeze(0).expand(B, -1, -1).reshape(B * num_points, -1).contiguous()
        keep = torch.stack(presence, dim=0).to(prior.device)
        x = prior.unsqueeze(0).expand(B, -1, -1)
        return x[keep, :].view(B * num_points, -1).contiguous()

    def sample(self, shape, given_eps=None):
        """Top-down generation (Network.py:251-268): given_eps [B, z_scales, n_layers*z_dim] -> [B, N, 3]."""
        B, num_points = shape[0], shape[1]
        if num_points is None:
            num_points = self.outsize
        P = self.packed()
        dev = self.output.weight.device
        H, T, Z, heads = self.hidden_dim, self.z_scales, self.z_dim, self.num_heads
        dh = H // heads
        with torch.no_grad():
            o = self.initial_set(B, num_points)  # f32 [B*N, H] residual stream
            if given_eps is None:
                given_eps = torch.randn((B, T, self.n_layers * Z)).to(o)  # :259-260
            eps = given_eps.detach().to(device=dev, dtype=torch.float32).contiguous().view(B * T, self.n_layers * Z)
            MQ, MT = B * num_points, B * T
            bf = torch.bfloat16
            zpad = _pad_to(Z, 64)
            e_a = torch.zeros((MT, zpad), dtype=bf, device=dev)
            xx = torch.empty((MT, H), dtype=bf, device=dev)
            kv = torch.empty((MT, 2 * H), dtype=bf, device=dev)
            a = torch.empty((MQ, H), dtype=bf, device=dev)
            q = torch.empty((MQ, H), dtype=bf, device=dev)
            att = torch.empty((MQ, H), dtype=bf, device=dev)
            hid = torch.empty((MQ, int(self.mlp_ratio * H)), dtype=bf, device=dev)
            for idx in range(self.n_layers):
                W = P["layers"][self.n_layers - 1 - idx]  # reversed(self.decoder), :263
                chunk = eps[:, idx * Z:(idx + 1) * Z]     # torch.split(...)[idx], :262
                _cast_strided(chunk, eps.stride(0), Z, e_a)
                ops.gemm(e_a, W["w_ln"], W["b_ln"], xx, EPI_BIAS_BF16)              # x = self.ln(eps)        :81
                ops.gemm(xx, W["w_kv"], W["b_kv"], kv, EPI_BIAS_BF16)               # kv = fc_kv(x)   layers.py:187
                ops.layernorm_mod(o, a, weight=W["n1w"], bias=W["n1b"])             # norm1 (affine)          :225
                ops.gemm(a, W["w_q"], W["b_q"], q, EPI_BIAS_BF16)                   # q = fc_q(norm1(o))
                ops.attention_nk32(B, heads, num_points, dh, q, H, kv, _PtrView(kv.data_ptr() + 2 * H), 2 * H, att)
                ops.gemm(att, W["w_o"], W["b_o"], o, EPI_GATE_RESID_F32, resid=o)   # o = o + fc_o(att)
                ops.layernorm_mod(o, a, weight=W["n2w"], bias=W["n2b"])             # norm2                   :226
                ops.gemm(a, W["w_fc1"], W["b_fc1"], hid, EPI_BIAS_GELU_BF16)
                ops.gemm(hid, W["w_fc2"], W["b_fc2"], o, EPI_GATE_RESID_F32, resid=o)
            ob = ops.cast_pad_bf16(o, H, out=a)
            pts8 = torch.empty((MQ, 8), dtype=torch.float32, device=dev)
            ops.gemm(ob, P["w_out"], P["b_out"], pts8, EPI_BIAS_F32)               # self.output(o)           :266
            out = pts8[:, :3].reshape(B, num_points, 3).contiguous()
        return self.postprocess(out)

    @staticmethod
    def postprocess(x):
        """Network.py:270-279."""
        if x.shape[-1] == 2:
            return (torch.tanh(x) + 1) / 2.0
        elif x.shape[-1] == 3:
            return x
        elif x.shape[-1] == 4:
            x = x.clone()
            x[..., -1] = (torch.tanh(x[..., -1]) + 1) / 2.0
            return x


def _cast_strided(view, ld_in, cols, out):
    """cast_pad on a column slice of a row-major f32 matrix (rows = view.shape[0], leading dim ld_in)."""
    from ._lib import check, load, ptr, stream_ptr
    with ops._launch("cast"):
        check(load().ldt_cast_pad_bf16(view.shape[0], cols, view.data_ptr(), ld_in, ptr(out), out.shape[1], stream_ptr()),
              "ldt_cast_pad_bf16")
