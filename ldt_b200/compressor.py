"""Host mirror of the reference ``Compressor`` for the sampling path (reference
model/Compressor/Network.py:105-279).

``Compressor(cfg.compressor)`` takes the same config keys, exposes ``sample(shape, given_eps=None)``,
``init()``, ``postprocess`` and a ``state_dict`` with exactly the reference's keys and shapes -- including the
encoder / grouper / position-embedding parameters that sampling never touches -- so stage-1 checkpoints
load with ``strict=True`` (trainer/Latent_SDE_Trainer.py:269-273).  Only the decoder is executed, on the
sm_100a kernels: per layer  Conv1d(z_dim->hidden) on the 32 latent tokens, K/V projection, LayerNorm(affine),
Q projection of the 2048 query rows, 2048x32 cross-attention, output projection + residual, LayerNorm, MLP with
GELU epilogue + residual (DecoderBlock.forward :80-83 -> ResidualBlock.forward c=None branch,
model/layers.py:224-226); finally Conv1d(hidden->3).

``forward(x)`` (bottom_up + top_down, Network.py:188-249: the inference path of the set-VAE encoder)
runs on the same kernels: FPS + k-NN grouping (``ldt_furthest_point_sample`` / ``ldt_knn_indices``), the AdaLN encoder
blocks on the 32 group tokens, the posterior blocks whose 32 tokens attend to the 2048 decoded points
(``ldt_attention_longkv``), and the decoder blocks above.  The point-wise layers around the grouping (input
Conv1d(3->hidden), the grouper's Conv + BatchNorm + ReLU stack, MiniPointnet) are error-compensated kind::tf32 contractions
("3xTF32", fp32-grade) of the same GEMM core with the eval-mode BatchNorm folded into the weights and the ReLU / residual in the epilogue; gather + normalise +
concatenate and the max over neighbours are ``ldt_group_features`` / ``ldt_group_max`` (``grouping.py``).  Only ActNorm's
per-channel affine is a torch expression.  Inference only: the KL terms are returned, nothing is differentiable.
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
        # "bf16": the product decoder (bf16 operands, fp32 accumulate); "fp32": the 3xTF32 parity mode of sample() (_decode_fp32)
        self.precision = "bf16"

    def reset_parameters(self):
        """torch's per-layer default initialisers drawn in the reference's CONSTRUCTION order (Network.py:125-157: input,
        conv_in, group, pos_embedding, then encoder.i / decoder.i interleaved, output, init_set, pre_grouper), so that
        ``torch.manual_seed(s); Compressor(cfg)`` yields the reference's random-init weights bit for bit (BASELINE
        configs[0]: "random-init score net + Compressor decoder"; pinned by tests/golden/init_hashes.json).
        Conv1d / Linear: kaiming_uniform(a=sqrt(5)) weight then U(-1/sqrt(fan_in), 1/sqrt(fan_in)) bias; LayerNorm /
        BatchNorm / affine_alpha: ones and zeros; ActNorm: zeros; InitialSet.prior: U(0, 1) (Compressor/layers.py:24)."""
        def group_of(name):
            top = name.split(".")[0]
            if top in ("encoder", "decoder"):
                return (5, int(name.split(".")[1]), 0 if top == "encoder" else 1)
            return ({"input": 0, "conv_in": 1, "group": 2, "pos_embedding": 3, "LabelEmbedding": 4, "output": 6,
                     "init_set": 7}.get(top, 8), 0, 0)

        params = dict(self.named_parameters())
        with torch.no_grad():
            for name in sorted(params, key=group_of):   # stable: registration order inside each group
                p = params[name]
                if name.endswith("prior"):
                    p.copy_(torch.rand(p.shape))
                elif name.endswith("label_emb.weight"):
                    p.normal_()
                elif name.endswith(".weight") and p.dim() >= 2:
                    nn.init.kaiming_uniform_(p, a=5 ** 0.5)
                    bias = params.get(name[:-len("weight")] + "bias")
                    if bias is not None:
                        bound = (1.0 / p[0].numel()) ** 0.5
                        bias.uniform_(-bound, bound)
                elif name.endswith("affine_alpha") or (p.dim() == 1 and name.endswith(".weight")):
                    p.fill_(1.0)      # LayerNorm / BatchNorm scale, grouper affine
                elif name.endswith(".bias") and (name[:-len("bias")] + "weight") in params \
                        and params[name[:-len("bias")] + "weight"].dim() >= 2:
                    pass              # drawn together with its weight above
                else:
                    p.zero_()         # norm biases, affine_beta, ActNorm shift / log_scale
            for name, b in self.named_buffers():
                if name.endswith("running_var"):
                    b.fill_(1.0)

    def init(self):
        """Network.py:163-165 -- marks ActNorm as initialised."""
        if self.ActNorm is not None:
            self.conv_in.initialized += 1.0

    def forward(self, x, num_points=None, label=None):
        """Bidirectional inference (Network.py:235-249): x [B, N, 3] -> dict(set, posteriors, kls, all_eps, all_logqz, max)."""
        if label is not None and self.class_condition:
            raise NotImplementedError("ldt_b200.Compressor: class-conditional encoding is not supported")
        with torch.no_grad():
            bup = self.bottom_up(x)
            tdn = self.top_down(bup["outputs"], num_points=num_points)
            all_eps = torch.cat(tdn["all_eps"], dim=1).transpose(1, 2)
            return {"set": self.postprocess(tdn["set"]), "posteriors": tdn["posteriors"], "kls": tdn["kls"],
                    "all_eps": all_eps, "all_logqz": tdn["all_logqz"], "max": bup["max"]}

    # ------------------------------------------------------------------------------------------
    def _fingerprint(self):
        return (getattr(self, "_generation", 0),) + tuple(
            (p.data_ptr(), p._version) for n, p in self.named_parameters()
            if n.startswith(("decoder.", "output.", "init_set.", "encoder.")))

    def invalidate_packed(self) -> None:
        """Drop the packed bf16 weights; needed only after in-place writes through ``.data`` (see Score.invalidate_packed)."""
        self._packed = None
        self._packed_key = None
        self._packed_pro = None
        self._packed_f32 = None
        self._generation = getattr(self, "_generation", 0) + 1

    def packed(self):
        key = self._fingerprint()
        if self._packed is not None and key == self._packed_key:
            return self._packed
        dev = self.output.weight.device
        if dev.type != "cuda":
            raise RuntimeError("ldt_b200.Compressor runs on CUDA only (no CPU fallback); call .to('cuda') first")
        H = self.hidden_dim
        P = {"layers": [], "post": [], "enc": []}
        with torch.no_grad():
            f32 = lambda t: t.detach().float().contiguous()

            def mlp_w(a):
                return {"w_fc1": ops.pack_weight(a.mlp.fc._modules["0"]._modules["0"].weight),
                        "b_fc1": f32(a.mlp.fc._modules["0"]._modules["0"].bias),
                        "w_fc2": ops.pack_weight(a.mlp.out.weight), "b_fc2": f32(a.mlp.out.bias)}

            def attn_w(a):
                return {"w_q": ops.pack_weight(a.fc_q.weight), "b_q": f32(a.fc_q.bias),
                        "w_kv": ops.pack_weight(a.fc_kv.weight), "b_kv": f32(a.fc_kv.bias),
                        "w_o": ops.pack_weight(a.fc_o.weight), "b_o": f32(a.fc_o.bias)}

            # encoder (Network.py:32-45): per layer `encoder_layers` AdaLN blocks + a FinalLayer; every adaLN consumes the
            # same SiLU(pos), so all of them are one GEMM against the row-concatenated weights
            ada_w, ada_b = [], []
            for l in range(self.n_layers):
                e = self.encoder._modules[str(l)]
                blocks = []
                for j in range(self.cfg.encoder_layers):
                    a = e.atts._modules[str(j)]
                    blocks.append({**attn_w(a), **mlp_w(a)})
                    ada_w.append(a.adaLN._modules["1"].weight.detach())
                    ada_b.append(a.adaLN._modules["1"].bias.detach())
                ada_w.append(e.conv_out.adaLN._modules["1"].weight.detach())
                ada_b.append(e.conv_out.adaLN._modules["1"].bias.detach())
                P["enc"].append({"blocks": blocks, "w_out": ops.pack_weight(e.conv_out.ln.weight), "b_out": f32(e.conv_out.ln.bias)})
            P["w_ada"] = ops.pack_weight(torch.cat(ada_w, dim=0))
            P["b_ada"] = torch.cat(ada_b).float().contiguous()
            for l in range(self.n_layers):
                d = self.decoder._modules[str(l)]
                a = d.att
                # posterior block `att` (c = None: affine LayerNorms) and the SiLU -> Conv1d prior head (:55,62-77)
                P["post"].append({**attn_w(a), **mlp_w(a),
                                  "n1w": f32(a.norm1.norm.weight), "n1b": f32(a.norm1.norm.bias),
                                  "n2w": f32(a.norm2.norm.weight), "n2b": f32(a.norm2.norm.bias),
                                  "w_prior": ops.pack_weight(d.prior._modules["1"].weight), "b_prior": f32(d.prior._modules["1"].bias)})
                a = d.att1
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
            bufs = (e_a, xx, kv, a, q, att, hid)
            pts8 = torch.empty((MQ, 8), dtype=torch.float32, device=dev)
            if getattr(self, "precision", "bf16") == "fp32":
                out = self._decode_fp32(B, num_points, eps, o)
                return self.postprocess(out)
            if getattr(self, "precision", "bf16") != "bf16":
                raise ValueError(f"ldt_b200.Compressor.precision must be 'bf16' or 'fp32', got {self.precision!r}")
            if getattr(self, "c_path", True):
                # the whole decoder as ONE call of the C entry point ldt_decoder_forward (same kernels, same order, same bits)
                from . import _lib
                layers = (_lib.DecoderLayer * self.n_layers)()
                for l, W in enumerate(P["layers"]):
                    for f, k in (("w_ln", "w_ln"), ("b_ln", "b_ln"), ("w_kv", "w_kv"), ("b_kv", "b_kv"), ("w_q", "w_q"), ("b_q", "b_q"),
                                 ("w_o", "w_o"), ("b_o", "b_o"), ("w_fc1", "w_fc1"), ("b_fc1", "b_fc1"), ("w_fc2", "w_fc2"),
                                 ("b_fc2", "b_fc2"), ("norm1_w", "n1w"), ("norm1_b", "n1b"), ("norm2_w", "n2w"), ("norm2_b", "n2b")):
                        setattr(layers[l], f, W[k].data_ptr())
                plan = _lib.DecoderPlan(batch=B, num_points=num_points, z_dim=Z, z_pad=zpad, hidden=H, heads=heads,
                                        mlp_hidden=hid.shape[1], n_layers=self.n_layers, layers=layers, w_out=P["w_out"].data_ptr(),
                                        b_out=P["b_out"].data_ptr(), ws_e=e_a.data_ptr(), ws_x=xx.data_ptr(), ws_kv=kv.data_ptr(),
                                        ws_a=a.data_ptr(), ws_q=q.data_ptr(), ws_att=att.data_ptr(), ws_hid=hid.data_ptr())
                with torch.cuda.device(dev), ops._launch("decoder_forward", 10 * self.n_layers + 2):
                    _lib.check(_lib.load().ldt_decoder_forward(_lib.C.byref(plan), eps.data_ptr(), o.data_ptr(), pts8.data_ptr(),
                                                               _lib.stream_ptr()), "ldt_decoder_forward")
            else:
                for idx in range(self.n_layers):
                    W = P["layers"][self.n_layers - 1 - idx]  # reversed(self.decoder), :263
                    chunk = eps[:, idx * Z:(idx + 1) * Z]     # torch.split(...)[idx], :262
                    self._decoder_block(W, B, num_points, chunk, eps.stride(0), o, bufs)
                ob = ops.cast_pad_bf16(o, H, out=a)
                ops.gemm(ob, P["w_out"], P["b_out"], pts8, EPI_BIAS_F32)               # self.output(o)           :266
            out = pts8[:, :3].reshape(B, num_points, 3).contiguous()
        return self.postprocess(out)

    def _packed_fp32(self):
        """3xTF32 ([hi | lo | hi]) copies of every contraction weight of the decoder, the posterior blocks and the encoder:
        the fp32 parity mode (``precision = "fp32"``) of sample() and forward()."""
        from . import grouping
        key = self._fingerprint()
        if getattr(self, "_packed_f32", None) is not None and self._packed_f32_key == key:
            return self._packed_f32
        f32 = lambda t: t.detach().float().contiguous()
        H = self.hidden_dim

        def block(a, affine):
            W = {"kv": grouping.pack_tf32(a.fc_kv), "q": grouping.pack_tf32(a.fc_q), "o": grouping.pack_tf32(a.fc_o),
                 "fc1": grouping.pack_tf32(a.mlp.fc._modules["0"]._modules["0"]), "fc2": grouping.pack_tf32(a.mlp.out)}
            if affine:
                W.update(n1w=f32(a.norm1.norm.weight), n1b=f32(a.norm1.norm.bias), n2w=f32(a.norm2.norm.weight), n2b=f32(a.norm2.norm.bias))
            return W

        with torch.no_grad():
            Q = {"layers": [], "post": [], "enc": []}
            ada_w, ada_b = [], []
            for l in range(self.n_layers):
                d = self.decoder._modules[str(l)]
                Q["layers"].append({"ln": grouping.pack_tf32(d.ln), **block(d.att1, True)})
                Q["post"].append({**block(d.att, True), "prior": grouping.pack_tf32(d.prior._modules["1"])})
                e = self.encoder._modules[str(l)]
                blocks = []
                for j in range(self.cfg.encoder_layers):
                    a = e.atts._modules[str(j)]
                    blocks.append(block(a, False))
                    ada_w.append(a.adaLN._modules["1"].weight.detach())
                    ada_b.append(a.adaLN._modules["1"].bias.detach())
                ada_w.append(e.conv_out.adaLN._modules["1"].weight.detach())
                ada_b.append(e.conv_out.adaLN._modules["1"].bias.detach())
                Q["enc"].append({"blocks": blocks, "out": grouping.pack_tf32(e.conv_out.ln)})
            wa = torch.cat(ada_w, dim=0).float().contiguous()
            Q["ada"] = (ops.split_tf32(wa, _pad_to(wa.shape[1], 32), weight_side=True), torch.cat(ada_b).float().contiguous())
            w_out = torch.zeros((8, H), dtype=torch.float32, device=wa.device)     # N padded to 8 rows (the GEMM core's N % 8 rule)
            w_out[:3] = self.output.weight.detach().reshape(3, H)
            b_out = torch.zeros(8, dtype=torch.float32, device=wa.device)
            b_out[:3] = self.output.bias.detach()
            Q["out"] = (ops.split_tf32(w_out, H, weight_side=True), b_out)
        self._packed_f32, self._packed_f32_key = Q, key
        return Q

    def _attn_block_fp32(self, W, B, x, kv_src, kv_tokens, n1, n2, gate1, gate2, mod_stride=0):
        """_attn_block in the fp32 parity mode: x f32 [B*32, H] in place, kv_src f32 [B*kv_tokens, H] (raw, un-normalised)."""
        from . import grouping
        from ._lib import EPI_BIAS_GELU_F32
        H, T, heads = self.hidden_dim, self.z_scales, self.num_heads
        a = torch.empty_like(x)
        att = torch.empty_like(x)
        ops.layernorm_mod_f32(x, a, round_tf32=False, **n1)
        q = grouping.conv_rows(a, W["q"])
        kv = grouping.conv_rows(kv_src, W["kv"])
        vptr = _PtrView(kv.data_ptr() + 4 * H)
        if kv_tokens == 32:
            ops.attention_nk32_f32(B, heads, T, H // heads, q, H, kv, vptr, 2 * H, att, round_tf32=False)
        else:
            ops.attention_longkv_f32(B, heads, T, kv_tokens, H // heads, q, H, kv, vptr, 2 * H, att)
        ops.gemm(ops.split_tf32(att, H), W["o"][0], W["o"][1], x, EPI_GATE_RESID_F32, resid=x, gate=gate1, gate_stride=mod_stride,
                 rows_per_gate=T, split_operands=True)
        ops.layernorm_mod_f32(x, a, round_tf32=False, **n2)
        hid = grouping.conv_rows(a, W["fc1"], EPI_BIAS_GELU_F32)
        ops.gemm(ops.split_tf32(hid, hid.shape[1]), W["fc2"][0], W["fc2"][1], x, EPI_GATE_RESID_F32, resid=x, gate=gate2,
                 gate_stride=mod_stride, rows_per_gate=T, split_operands=True)

    def _decoder_block_fp32(self, W, B, num_points, chunk, o):
        """DecoderBlock.forward (:80-83) in the fp32 parity mode: o f32 [B*N, H] (in place) attends to chunk f32 [B*32, z_dim]."""
        from . import grouping
        from ._lib import EPI_BIAS_GELU_F32
        H, heads = self.hidden_dim, self.num_heads
        a = torch.empty_like(o)
        att = torch.empty_like(o)
        xx = grouping.conv_rows(chunk, W["ln"])                                        # x = self.ln(eps)          :81
        kv = grouping.conv_rows(xx, W["kv"])                                           # fc_kv(x): raw, un-normalised x
        ops.layernorm_mod_f32(o, a, weight=W["n1w"], bias=W["n1b"], round_tf32=False)
        q = grouping.conv_rows(a, W["q"])
        ops.attention_nk32_f32(B, heads, num_points, H // heads, q, H, kv, _PtrView(kv.data_ptr() + 4 * H), 2 * H, att,
                               round_tf32=False)
        grouping.conv_rows(att, W["o"], EPI_GATE_RESID_F32, resid=o, out=o)             # o = o + fc_o(att)
        ops.layernorm_mod_f32(o, a, weight=W["n2w"], bias=W["n2b"], round_tf32=False)
        hid = grouping.conv_rows(a, W["fc1"], EPI_BIAS_GELU_F32)
        grouping.conv_rows(hid, W["fc2"], EPI_GATE_RESID_F32, resid=o, out=o)

    def _decode_fp32(self, B, num_points, eps, o):
        """The decoder in the fp32 parity mode (``precision = "fp32"``): the same layer sequence with fp32 activations, 3xTF32
        contractions (grouping.conv_rows), fp32 LayerNorm / attention / GELU without intermediate rounding.  A parity
        instrument like Score.precision = "fp32" (DESIGN.md 4.8); eps f32 [B*32, n_layers*z_dim], o f32 [B*N, H] in place."""
        from . import grouping
        Z = self.z_dim
        Q = self._packed_fp32()
        for idx in range(self.n_layers):
            W = Q["layers"][self.n_layers - 1 - idx]                                   # reversed(self.decoder)   :263
            chunk = eps[:, idx * Z:(idx + 1) * Z].contiguous()                         # torch.split(...)[idx]    :262
            self._decoder_block_fp32(W, B, num_points, chunk, o)
        pts8 = grouping.conv_rows(o, Q["out"])                                         # self.output(o)            :266
        return pts8[:, :3].reshape(B, num_points, 3).contiguous()

    def _decoder_block(self, W, B, num_points, chunk, ld_chunk, o, bufs):
        """DecoderBlock.forward (:80-83): o [B*N, H] f32 (in place) attends to the layer's latent chunk [B*32, z_dim]."""
        H, Z, heads = self.hidden_dim, self.z_dim, self.num_heads
        e_a, xx, kv, a, q, att, hid = bufs
        _cast_strided(chunk, ld_chunk, Z, e_a)
        ops.gemm(e_a, W["w_ln"], W["b_ln"], xx, EPI_BIAS_BF16)              # x = self.ln(eps)        :81
        ops.gemm(xx, W["w_kv"], W["b_kv"], kv, EPI_BIAS_BF16)               # kv = fc_kv(x)   layers.py:187
        ops.layernorm_mod(o, a, weight=W["n1w"], bias=W["n1b"])             # norm1 (affine)          :225
        ops.gemm(a, W["w_q"], W["b_q"], q, EPI_BIAS_BF16)                   # q = fc_q(norm1(o))
        ops.attention_nk32(B, heads, num_points, H // heads, q, H, kv, _PtrView(kv.data_ptr() + 2 * H), 2 * H, att)
        ops.gemm(att, W["w_o"], W["b_o"], o, EPI_GATE_RESID_F32, resid=o)   # o = o + fc_o(att)
        ops.layernorm_mod(o, a, weight=W["n2w"], bias=W["n2b"])             # norm2                   :226
        ops.gemm(a, W["w_fc1"], W["b_fc1"], hid, EPI_BIAS_GELU_BF16)
        ops.gemm(hid, W["w_fc2"], W["b_fc2"], o, EPI_GATE_RESID_F32, resid=o)

    # ------------------------------------------------------------------------------------------
    # encoder inference path
    # ------------------------------------------------------------------------------------------
    def _packed_prologue(self):
        """Folded TF32 weights of the point-wise layers in front of the encoder blocks (input Conv1d, the groupers'
        Conv + BatchNorm stacks, MiniPointnet); cached like :meth:`packed`, keyed on parameters AND BatchNorm buffers."""
        from . import grouping
        pre = ("input.", "group.", "pre_grouper.", "pos_embedding.")
        key = (getattr(self, "_generation", 0),) + tuple(
            (t.data_ptr(), t._version) for n, t in list(self.named_parameters()) + list(self.named_buffers()) if n.startswith(pre))
        if getattr(self, "_packed_pro", None) is not None and key == self._packed_pro_key:
            return self._packed_pro
        with torch.no_grad():
            P = {"input": grouping.pack_tf32(self.input), "group": grouping.pack_grouper(self.group),
                 "pos": grouping.pack_mini_pointnet(self.pos_embedding)}
            if self.cfg.pre_group:
                P["pre_grouper"] = grouping.pack_grouper(self.pre_grouper)
        self._packed_pro, self._packed_pro_key = P, key
        return P

    def _attn_block(self, W, B, x, kv_src, kv_tokens, n1, n2, gate1, gate2, bufs, mod_stride=0):
        """One ResidualBlock on the 32 group tokens with K/V taken from ``kv_src`` (bf16 [B*kv_tokens, H], NOT normalised:
        compute_attention receives the raw ``y``, layers.py:184-187).  n1 / n2 are the keyword arguments of the two
        LayerNorm passes (AdaLN shift/scale or affine weight/bias), gate1 / gate2 the AdaLN gates or None."""
        H, T, heads = self.hidden_dim, self.z_scales, self.num_heads
        a, q, kv, att, hid = bufs
        ops.layernorm_mod(x, a, **n1)
        ops.gemm(a, W["w_q"], W["b_q"], q, EPI_BIAS_BF16)
        ops.gemm(kv_src, W["w_kv"], W["b_kv"], kv, EPI_BIAS_BF16)
        vptr = _PtrView(kv.data_ptr() + 2 * H)
        if kv_tokens == 32:
            ops.attention_nk32(B, heads, T, H // heads, q, H, kv, vptr, 2 * H, att)
        else:
            ops.attention_longkv(B, heads, T, kv_tokens, H // heads, q, H, kv, vptr, 2 * H, att)
        ops.gemm(att, W["w_o"], W["b_o"], x, EPI_GATE_RESID_F32, resid=x, gate=gate1, gate_stride=mod_stride, rows_per_gate=T)
        ops.layernorm_mod(x, a, **n2)
        ops.gemm(a, W["w_fc1"], W["b_fc1"], hid, EPI_BIAS_GELU_BF16)
        ops.gemm(hid, W["w_fc2"], W["b_fc2"], x, EPI_GATE_RESID_F32, resid=x, gate=gate2, gate_stride=mod_stride, rows_per_gate=T)

    def encoder_prologue(self, pts):
        """Network.py:189-199: points [B,N,3] -> (group tokens [B, 32, H] f32 after ActNorm, position embedding [B, p_dim])."""
        cfg = self.cfg
        dev = self.output.weight.device
        H, T = self.hidden_dim, self.z_scales
        pts = pts.to(dev).float()
        if cfg.norm_input:
            pts = (pts - pts.mean(dim=1, keepdim=True)) / pts.std(dim=1, keepdim=True)
        B, N = pts.shape[0], pts.shape[1]
        from . import grouping
        Q = self._packed_prologue()
        pts = pts.contiguous()
        # input Conv1d(3 -> hidden) on every point (Network.py:192), rows [B*N, hidden]
        x = grouping.conv_rows(pts.reshape(B * N, 3), Q["input"]).reshape(B, N, H)
        if cfg.pre_group:
            pts, x = grouping.local_group(self.pre_grouper, Q["pre_grouper"], cfg.cluster_norm, pts, x, 256, 32)
            x = x.reshape(B, 256, H)
        center, x = grouping.local_group(self.group, Q["group"], cfg.cluster_norm, pts, x, T, pts.shape[1] // T * 2)
        pos = grouping.mini_pointnet(Q["pos"], center)                     # MiniPointnet, Network.py:86-101 -> [B, p_dim]
        x = x.reshape(B, T, H)                                             # [B, 32, H] token-major
        if self.ActNorm is not None:
            x = (x - self.conv_in.shift) * torch.exp(-self.conv_in.log_scale)   # ActNorm.forward, model/layers.py:103-107
        return x, pos

    def bottom_up(self, pts, label=None):
        """Network.py:188-209: points [B,N,3] -> per-layer encoder outputs (token-major f32 [B*32, H]) and max feature."""
        cfg = self.cfg
        dev = self.output.weight.device
        if dev.type != "cuda":
            raise RuntimeError("ldt_b200.Compressor runs on CUDA only (no CPU fallback); call .to('cuda') first")
        if cfg.pos_embedding == "mlp":
            raise NotImplementedError("ldt_b200.Compressor: pos_embedding 'mlp' (per-token conditioning) is not supported")
        if cfg.encoder_dropout_p != 0:
            raise NotImplementedError("ldt_b200.Compressor is an inference path: encoder_dropout_p must be 0")
        P = self.packed()
        H, T, Pd = self.hidden_dim, self.z_scales, self.p_dim
        x, pos = self.encoder_prologue(pts)
        B = x.shape[0]
        x = x.reshape(B * T, H).contiguous()                               # token-major residual stream
        if getattr(self, "precision", "bf16") == "fp32":
            return self._bottom_up_fp32(B, x, pos)
        # all adaLN rows of the encoder from SiLU(pos) in one GEMM
        sc = torch.empty((B, Pd), dtype=torch.bfloat16, device=dev)
        ops.cond_silu(torch.zeros((1, Pd), device=dev), None, pos, None, sc)
        L = cfg.encoder_layers
        row = (6 * L + 2) * H
        mod = torch.empty((B, self.n_layers * row), dtype=torch.float32, device=dev)
        ops.gemm(sc, P["w_ada"], P["b_ada"], mod, EPI_BIAS_F32)
        stride = mod.shape[1]
        mv = lambda off: _PtrView(mod.data_ptr() + 4 * off)
        bf = torch.bfloat16
        bufs = tuple(torch.empty((B * T, w), dtype=bf, device=dev) for w in (H, H, 2 * H, H, int(self.mlp_ratio * H)))
        xb = torch.empty((B * T, H), dtype=bf, device=dev)
        outputs = []
        for l in range(self.n_layers):
            for j, W in enumerate(P["enc"][l]["blocks"]):
                base = l * row + j * 6 * H
                ops.cast_pad_bf16(x, H, out=xb)                            # layer(x, x, pos): K/V from the raw x
                self._attn_block(W, B, x, xb, 32,
                                 dict(shift=mv(base), scale=mv(base + H), mod_stride=stride, rows_per_mod=T),
                                 dict(shift=mv(base + 3 * H), scale=mv(base + 4 * H), mod_stride=stride, rows_per_mod=T),
                                 mv(base + 2 * H), mv(base + 5 * H), bufs, mod_stride=stride)
            base = l * row + L * 6 * H                                     # FinalLayer: (shift, scale) then Conv1d
            ops.layernorm_mod(x, bufs[0], shift=mv(base), scale=mv(base + H), mod_stride=stride, rows_per_mod=T)
            o = torch.empty((B * T, H), dtype=torch.float32, device=dev)
            ops.gemm(bufs[0], P["enc"][l]["w_out"], P["enc"][l]["b_out"], o, EPI_BIAS_F32)
            outputs.append(o)
        return {"outputs": outputs, "max": x.max()}

    def _bottom_up_fp32(self, B, x, pos):
        """The encoder blocks of bottom_up in the fp32 parity mode (same sequence, 3xTF32 contractions)."""
        from . import grouping
        cfg = self.cfg
        Q = self._packed_fp32()
        H, T = self.hidden_dim, self.z_scales
        L = cfg.encoder_layers
        row = (6 * L + 2) * H
        mod = torch.empty((B, self.n_layers * row), dtype=torch.float32, device=x.device)
        ops.gemm(ops.split_tf32(pos.contiguous(), silu=True), Q["ada"][0], Q["ada"][1], mod, EPI_BIAS_F32, split_operands=True)
        stride = mod.shape[1]
        mv = lambda off: _PtrView(mod.data_ptr() + 4 * off)
        a = torch.empty_like(x)
        outputs = []
        for l in range(self.n_layers):
            for j, W in enumerate(Q["enc"][l]["blocks"]):
                base = l * row + j * 6 * H
                self._attn_block_fp32(W, B, x, x.clone(), 32,                  # layer(x, x, pos): K/V from the raw x
                                      dict(shift=mv(base), scale=mv(base + H), mod_stride=stride, rows_per_mod=T),
                                      dict(shift=mv(base + 3 * H), scale=mv(base + 4 * H), mod_stride=stride, rows_per_mod=T),
                                      mv(base + 2 * H), mv(base + 5 * H), mod_stride=stride)
            base = l * row + L * 6 * H
            ops.layernorm_mod_f32(x, a, shift=mv(base), scale=mv(base + H), mod_stride=stride, rows_per_mod=T, round_tf32=False)
            outputs.append(grouping.conv_rows(a, Q["enc"][l]["out"]))
        return {"outputs": outputs, "max": x.max()}

    def _top_down_fp32(self, encoder_out, N):
        """top_down in the fp32 parity mode; same generator consumption and returned dictionary as the bf16 path."""
        from . import grouping
        Q = self._packed_fp32()
        H, T, Z = self.hidden_dim, self.z_scales, self.z_dim
        B = encoder_out[0].shape[0] // T
        o = self.initial_set(B, N)
        MT = B * T
        cf = lambda t, c: t.view(B, T, c).transpose(1, 2)
        posteriors = [(o.view(B, N, H).transpose(1, 2).clone(), None, None)]
        kls, all_eps, all_logqz = [], [], []
        for idx in range(self.n_layers):
            Li = self.n_layers - 1 - idx
            W = Q["post"][Li]
            x = encoder_out[Li].clone()
            n1, n2 = dict(weight=W["n1w"], bias=W["n1b"]), dict(weight=W["n2w"], bias=W["n2b"])
            if idx != 0:
                self._attn_block_fp32(W, B, x, o, N, n1, n2, None, None)
            else:
                self._attn_block_fp32(W, B, x, x.clone(), 32, n1, n2, None, None)
            post = torch.empty((MT, 2 * Z), dtype=torch.float32, device=x.device)
            ops.gemm(ops.split_tf32(x, H, silu=True), W["prior"][0], W["prior"][1], post, EPI_BIAS_F32, split_operands=True)
            mu = cf(post[:, :Z], Z)
            logvar = cf(post[:, Z:], Z).clamp(self.cfg.min_sigma, 10.0)
            noise = torch.randn(mu.shape).to(mu)
            eps = mu + torch.exp(logvar / 2.0) * noise
            logqz = -0.5 * torch.square(eps - mu) / torch.exp(logvar) - 0.5 * logvar - 0.9189385332
            logpz = -0.5 * torch.square(eps) - 0.9189385332
            self._decoder_block_fp32(Q["layers"][Li], B, N, eps.transpose(1, 2).reshape(MT, Z).contiguous(), o)
            all_eps.append(eps)
            posteriors.append((eps, mu, logvar))
            kls.append(logqz - logpz)
            all_logqz.append(logqz)
        pts8 = grouping.conv_rows(o, Q["out"])
        return {"set": pts8[:, :3].reshape(B, N, 3).contiguous(), "posteriors": posteriors, "kls": kls,
                "all_logqz": all_logqz, "all_eps": all_eps}

    def top_down(self, encoder_out, num_points=None, label=None):
        """Stochastic top-down pass (Network.py:211-233).  Tensors in the returned dict use the reference's
        channels-first shapes ([B, z_dim, 32] latents, [B, N, 3] set)."""
        if getattr(self, "precision", "bf16") == "fp32":
            return self._top_down_fp32(encoder_out, num_points if num_points is not None else self.outsize)
        P = self.packed()
        dev = self.output.weight.device
        H, T, Z, heads = self.hidden_dim, self.z_scales, self.z_dim, self.num_heads
        B = encoder_out[0].shape[0] // T
        N = num_points if num_points is not None else self.outsize
        bf = torch.bfloat16
        o = self.initial_set(B, N)
        MQ, MT = B * N, B * T
        dbufs = (torch.zeros((MT, _pad_to(Z, 64)), dtype=bf, device=dev), torch.empty((MT, H), dtype=bf, device=dev),
                 torch.empty((MT, 2 * H), dtype=bf, device=dev), torch.empty((MQ, H), dtype=bf, device=dev),
                 torch.empty((MQ, H), dtype=bf, device=dev), torch.empty((MQ, H), dtype=bf, device=dev),
                 torch.empty((MQ, int(self.mlp_ratio * H)), dtype=bf, device=dev))
        pbufs = (torch.empty((MT, H), dtype=bf, device=dev), torch.empty((MT, H), dtype=bf, device=dev),
                 torch.empty((MQ, 2 * H), dtype=bf, device=dev), torch.empty((MT, H), dtype=bf, device=dev),
                 torch.empty((MT, int(self.mlp_ratio * H)), dtype=bf, device=dev))
        ob = torch.empty((MQ, H), dtype=bf, device=dev)
        xb = torch.empty((MT, H), dtype=bf, device=dev)
        zero_row = torch.zeros((1, H), device=dev)
        cf = lambda t, c: t.view(B, T, c).transpose(1, 2)                    # token-major -> [B, C, 32]
        posteriors = [(o.view(B, N, H).transpose(1, 2).clone(), None, None)]
        kls, all_eps, all_logqz = [], [], []
        for idx in range(self.n_layers):
            Li = self.n_layers - 1 - idx                                     # reversed(self.decoder)
            W = P["post"][Li]
            x = encoder_out[Li].clone()
            n1, n2 = dict(weight=W["n1w"], bias=W["n1b"]), dict(weight=W["n2w"], bias=W["n2b"])
            if idx != 0:                                                     # att(x, o, c): tokens attend to the current set
                ops.cast_pad_bf16(o, H, out=ob)
                self._attn_block(W, B, x, ob, N, n1, n2, None, None, pbufs)
            else:
                ops.cast_pad_bf16(x, H, out=xb)
                self._attn_block(W, B, x, xb, 32, n1, n2, None, None, pbufs)
            sx = pbufs[0]
            ops.cond_silu(zero_row, None, x, None, sx)                        # prior = Conv1d(SiLU(x))  :55
            post = torch.empty((MT, 2 * Z), dtype=torch.float32, device=dev)
            ops.gemm(sx, W["w_prior"], W["b_prior"], post, EPI_BIAS_F32)
            mu = cf(post[:, :Z], Z)
            logvar = cf(post[:, Z:], Z).clamp(self.cfg.min_sigma, 10.0)
            noise = torch.randn(mu.shape).to(mu)                             # sample(), Network.py:26-29 (CPU generator)
            eps = mu + torch.exp(logvar / 2.0) * noise
            logqz = -0.5 * torch.square(eps - mu) / torch.exp(logvar) - 0.5 * logvar - 0.9189385332
            logpz = -0.5 * torch.square(eps) - 0.9189385332
            eps_tok = eps.transpose(1, 2).reshape(MT, Z).contiguous()
            self._decoder_block(P["layers"][Li], B, N, eps_tok, Z, o, dbufs)
            all_eps.append(eps)
            posteriors.append((eps, mu, logvar))
            kls.append(logqz - logpz)
            all_logqz.append(logqz)
        pts8 = torch.empty((MQ, 8), dtype=torch.float32, device=dev)
        ops.gemm(ops.cast_pad_bf16(o, H, out=ob), P["w_out"], P["b_out"], pts8, EPI_BIAS_F32)
        return {"set": pts8[:, :3].reshape(B, N, 3).contiguous(), "posteriors": posteriors, "kls": kls,
                "all_logqz": all_logqz, "all_eps": all_eps}

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
