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
import torch.nn.functional as F

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


def conv_rows(x: torch.Tensor, packed, epilogue: int = EPI_BIAS_F32, resid=None, out: torch.Tensor | None = None) -> torch.Tensor:
    """rows [M, <= pad32(in)] f32 -> [M, out] f32: split into [hi | hi | lo], then one kind::tf32 contraction over 3*pad32(in)."""
    W, b = packed
    x3 = ops.split_tf32(x, W.shape[1] // 3)
    if out is None:
        out = torch.empty((x.shape[0], W.shape[0]), dtype=torch.float32, device=x.device)
    return ops.gemm(x3, W, b, out, epilogue, resid=resid, split_operands=True)


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


# ------------------------------------------------------------------------------------------------
# ConditionNet's image branch (model/scorenet/score.py:24-26,33-35): torchvision ResNet18 stem + layer1 + layer2, global
# max-pool, Linear.  Every convolution is im2col (F.unfold: a copy, no arithmetic) followed by the same 3xTF32 contraction
# with the eval-mode BatchNorm folded in and ReLU / the BasicBlock's residual in the epilogue; both max-pools are
# ldt_group_max.  Activations are kept as channels-last rows [B*H*W, C].
# ------------------------------------------------------------------------------------------------
def _conv2d_rows(rows, shape, conv, bn, epilogue, resid=None):
    """rows [B*H*W, C] (+ shape (B, H, W)) -> rows [B*Ho*Wo, Cout] and (B, Ho, Wo) for one Conv2d + BatchNorm2d."""
    B, H, W = shape
    x = rows.reshape(B, H, W, -1).permute(0, 3, 1, 2)
    kh, kw = conv.kernel_size
    (sh, sw), (ph, pw) = conv.stride, conv.padding
    if conv.groups != 1 or conv.dilation != (1, 1):
        raise NotImplementedError("ldt_b200: grouped / dilated Conv2d is not supported")
    Ho, Wo = (H + 2 * ph - kh) // sh + 1, (W + 2 * pw - kw) // sw + 1
    cols = F.unfold(x, (kh, kw), padding=(ph, pw), stride=(sh, sw))            # [B, C*kh*kw, Ho*Wo], (C, kh, kw) order
    a = cols.transpose(1, 2).reshape(B * Ho * Wo, -1).contiguous()
    return conv_rows(a, pack_tf32(conv, bn), epilogue, resid=resid), (B, Ho, Wo)


def _maxpool2d_rows(rows, shape, pool):
    """nn.MaxPool2d on non-negative rows (after a ReLU, so F.unfold's zero padding never wins over a real element)."""
    B, H, W = shape
    k, s, p = pool.kernel_size, pool.stride, pool.padding
    C_ = rows.shape[1]
    x = rows.reshape(B, H, W, C_).permute(0, 3, 1, 2)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    cols = F.unfold(x, k, padding=p, stride=s).reshape(B, C_, k * k, Ho * Wo)   # window elements of every channel
    a = cols.permute(0, 3, 2, 1).reshape(B * Ho * Wo * k * k, C_).contiguous()
    return ops.group_max(a, k * k), (B, Ho, Wo)


def _basic_block_rows(rows, shape, blk):
    """torchvision BasicBlock: relu(bn2(conv2(relu(bn1(conv1(x))))) + (downsample(x) | x))."""
    y, shp = _conv2d_rows(rows, shape, blk.conv1, blk.bn1, EPI_BIAS_RELU_F32)
    identity = rows
    if blk.downsample is not None:
        identity, _ = _conv2d_rows(rows, shape, blk.downsample[0], blk.downsample[1], EPI_BIAS_F32)
    return _conv2d_rows(y, shp, blk.conv2, blk.bn2, EPI_RESID_RELU_F32, resid=identity)


def resnet_trunk_maxpool(trunk, img: torch.Tensor) -> torch.Tensor:
    """``adaptive_max_pool2d(trunk(img), 1)`` for trunk = Sequential(conv1, bn1, relu, maxpool, layer1, layer2):
    img [B,3,H,W] -> [B, 128] f32."""
    conv1, bn1, _, pool, layer1, layer2 = list(trunk.children())
    B, _, H, W = img.shape
    rows = img.float().permute(0, 2, 3, 1).reshape(B * H * W, -1).contiguous()
    rows, shp = _conv2d_rows(rows, (B, H, W), conv1, bn1, EPI_BIAS_RELU_F32)
    rows, shp = _maxpool2d_rows(rows, shp, pool)
    for blk in list(layer1) + list(layer2):
        rows, shp = _basic_block_rows(rows, shp, blk)
    return ops.group_max(rows, shp[1] * shp[2])
