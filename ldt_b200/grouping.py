"""Host side of the point-set prologue shared by the Compressor encoder (reference model/Compressor/Network.py:188-199) and
``ConditionNet`` (model/scorenet/score.py:36-41): ``LocalGrouper`` = FPS centres + k-NN groups + normalised group features
(model/Compressor/layers.py:288-319) followed by ``PreExtraction`` = Conv1d/BatchNorm/ReLU layers and a max over the
neighbours (layers.py:163-192), and ``MiniPointnet`` (Network.py:86-101).

Everything dense runs on the library's own kernels: the 1x1 convolutions are ``ldt_gemm_bf16`` contractions on the tcgen05
kind::tf32 path with error-compensated operands ("3xTF32", ``ldt_split_tf32``: a = hi + lo, one contraction over 3 K
computes a_hi.w_hi + a_hi.w_lo + a_lo.w_hi), i.e. fp32-grade results as the reference's fp32 layers give, with the
eval-mode BatchNorm folded into the weight and bias and the ReLU (and the residual of ConvBNReLURes1D) in the epilogue
(``LDT_EPI_BIAS_RELU_F32`` / ``LDT_EPI_RESID_RELU_F32``); gather + normalise + concatenate is ``ldt_group_features``; the max
over neighbours ``ldt_group_max``.  torch only allocates.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import EPI_BIAS_F32, EPI_BIAS_RELU_F32, EPI_RESID_RELU_F32


def _pad32(k: int) -> int:
    return ((k + 31) // 32) * 32


def fold_conv_bn(conv, bn=None):
    """Conv1d(k=1) / Linear followed by an eval-mode BatchNorm1d -> (W [out,in], b [out]) f32 of the single affine map."""
    W = conv.weight.detach().float().reshape(conv.weight.shape[0], -1)
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(W.shape[0], device=W.device)
    if bn is not None:
        s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + getattr(bn, "eps", 1e-5))
        W = W * s[:, None]
        b = (b - bn.running_mean.detach().float()) * s + bn.bias.detach().float()
    return W.contiguous(), b.contiguous()


def pack_tf32(conv, bn=None):
    """(W f32 [out, 3*pad32(in)] in the [hi | lo | hi] split layout, bias f32 [out]) on the module's device."""
    W, b = fold_conv_bn(conv, bn)
    return ops.split_tf32(W, _pad32(W.shape[1]), weight_side=True), b


def conv_rows(x: torch.Tensor, packed, epilogue: int = EPI_BIAS_F32, resid=None) -> torch.Tensor:
    """rows [M, >= in] f32 -> [M, out] f32: split into [hi | hi | lo], then one kind::tf32 contraction over 3*pad32(in)."""
    W, b = packed
    ld_part = W.shape[1] // 3
    x3 = ops.split_tf32(x if x.shape[1] <= ld_part else x[:, :ld_part], ld_part)
    out = torch.empty((x.shape[0], W.shape[0]), dtype=torch.float32, device=x.device)
    return ops.gemm(x3, W, b, out, epilogue, resid=resid)


def pack_grouper(g) -> dict:
    """Folded TF32 weights of a LocalGrouper's PreExtraction sub-tree (parameter names of the reference)."""
    ex = g.extraction
    t = ex.transfer.net._modules            # works for nn.Sequential and for the bare parameter trees of compressor.py
    blocks = []
    for op in ex.operation._modules.values():
        n1, n2 = op.net1._modules, op.net2._modules
        if len(n2) != 1:
            raise NotImplementedError("ldt_b200: ConvBNReLURes1D with groups > 1 is not supported")
        blocks.append((pack_tf32(n1["0"], n1["1"]), pack_tf32(n2["0"])))
    return {"transfer": pack_tf32(t["0"], t["1"]), "blocks": blocks}


def local_group(g, packed: dict, normalize, pts: torch.Tensor, fea: torch.Tensor, groups: int, k: int):
    """LocalGrouper.forward on rows: pts [B,N,3], fea [B,N,D] f32 -> (centres [B,S,3], group features [B*S, D] f32)."""
    from .condition import cluster, gather_points
    pts = pts.contiguous().float()
    fea = fea.contiguous().float()
    new_xyz, fps_idx, idx = cluster(pts, groups, k)
    normalize = normalize.lower() if isinstance(normalize, str) else None
    if normalize not in ("center", "anchor"):
        normalize = None
    alpha = g.affine_alpha if normalize is not None else None
    beta = g.affine_beta if normalize is not None else None
    rows = ops.group_features(pts, fea, fps_idx.int().contiguous(), idx.int().contiguous(), normalize, alpha, beta)
    x = conv_rows(rows, packed["transfer"], EPI_BIAS_RELU_F32)           # transfer: Conv + BN + ReLU      layers.py:185
    for net1, net2 in packed["blocks"]:                                   # ConvBNReLURes1D                 layers.py:159-160
        y = conv_rows(x, net1, EPI_BIAS_RELU_F32)
        x = conv_rows(y, net2, EPI_RESID_RELU_F32, resid=x)
    return new_xyz, ops.group_max(x, k)                                   # adaptive_max_pool1d(x, 1)       layers.py:189


def pack_mini_pointnet(pe) -> dict:
    return {"conv1": pack_tf32(pe.conv1, pe.bn1), "conv2": pack_tf32(pe.conv2, pe.bn2), "fc": pack_tf32(pe.fc)}


def mini_pointnet(packed: dict, center: torch.Tensor) -> torch.Tensor:
    """MiniPointnet.forward (Network.py:93-100): centres [B,S,3] -> [B, output_dim]."""
    B, S, _ = center.shape
    x = conv_rows(center.reshape(B * S, 3).contiguous(), packed["conv1"], EPI_BIAS_RELU_F32)
    x = conv_rows(x, packed["conv2"], EPI_BIAS_RELU_F32)
    x = ops.group_max(x, S)                                               # torch.max(x, 2)
    return conv_rows(x, packed["fc"], EPI_BIAS_F32)
