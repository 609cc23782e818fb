"""Host mirror of the completion prologue: ``ConditionNet`` (reference model/scorenet/score.py:13-44) and the
``LocalGrouper`` it is built on (reference model/Compressor/layers.py:288-319, 225-256, 130-177).

Runs ONCE per ``sample()`` call, before the reverse-SDE loop (completion_trainer/Latent_SDE_Trainer.py:150-151), so
it is a prologue, not the hot loop.  What is B200-native here is the point-set part the reference cannot run without
its un-vendored ``pointnet2_ops`` dependency: furthest point sampling and k-NN grouping are sm_100a kernels behind the
C ABI (``ldt_furthest_point_sample``, ``ldt_knn_indices``), and the 1x1 convolutions with BatchNorm around them run as
error-compensated kind::tf32 contractions of the library's GEMM core (``grouping.py``: fp32-grade "3xTF32", BatchNorm folded, ReLU / residual in the epilogue,
``ldt_group_features`` / ``ldt_group_max``).  The ResNet18 stem + layer1 + layer2 on the image run the same way: im2col
(``F.unfold``, a copy) + the 3xTF32 contraction with folded BatchNorm, ReLU and the BasicBlock residual in the epilogue, both
max-pools on ``ldt_group_max`` -- no cuDNN on the path.  The modules keep the reference's parameter names, so reference
checkpoints load with ``strict=True``:

  c_net.pc_conv_in, c_net.group.{affine_alpha, affine_beta, extraction.transfer.net.{0,1},
  extraction.operation.0.{net1.{0,1}, net2.0}}, c_net.pc_conv_out, c_net.resnet.{0,1,4,5}.*, c_net.ln, c_net.conv_out
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


def gather_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """points [B,N,C], idx [B,...] (any integer dtype) -> [B,...,C]   (index_points, Compressor/layers.py:46-62)."""
    B = points.shape[0]
    flat = idx.reshape(B, -1).long()
    out = torch.gather(points, 1, flat.unsqueeze(-1).expand(-1, -1, points.shape[-1]))
    return out.reshape(*idx.shape, points.shape[-1])


def furthest_point_sample(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    """Drop-in for ``pointnet2_utils.furthest_point_sample(xyz [B,N,3], npoint) -> int32 [B,npoint]``."""
    return ops.furthest_point_sample(xyz.contiguous().float(), npoint)


def cluster(xyz: torch.Tensor, groups: int, k: int, center=None):
    """(new_xyz [B,S,3], center_idx [B,S] | None, group_idx [B,S,k]) -- Compressor/layers.py:101-112."""
    xyz = xyz.contiguous().float()
    if center is None:
        center_idx = furthest_point_sample(xyz, groups).long()
        new_xyz = gather_points(xyz, center_idx)
    else:
        new_xyz, center_idx = center.contiguous().float(), None
    group_idx = ops.knn_indices(k, xyz, new_xyz.contiguous()).long()
    return new_xyz, center_idx, group_idx


class _ConvBNAct(nn.Module):
    """Parameters of Conv1d(k=1) + BatchNorm1d + ReLU under the key ``net`` (ConvBNReLU1D, Compressor/layers.py:115-127).
    A parameter container: the arithmetic is ``grouping.local_group`` on the library's kernels."""

    def __init__(self, cin, cout):
        super().__init__()
        self.net = nn.Sequential(nn.Conv1d(cin, cout, 1), nn.BatchNorm1d(cout), nn.ReLU(inplace=True))


class _ConvBNActRes(nn.Module):
    """Parameters of the residual 1x1 block, keys ``net1`` / ``net2`` (ConvBNReLURes1D with groups=1, layers.py:130-160)."""

    def __init__(self, ch):
        super().__init__()
        self.net1 = nn.Sequential(nn.Conv1d(ch, ch, 1), nn.BatchNorm1d(ch), nn.ReLU(inplace=True))
        self.net2 = nn.Sequential(nn.Conv1d(ch, ch, 1))


class _PreExtraction(nn.Module):
    """Parameters of PreExtraction (Compressor/layers.py:163-192): ``transfer`` + one residual block."""

    def __init__(self, channels, out_channels, use_xyz=True):
        super().__init__()
        self.transfer = _ConvBNAct((3 if use_xyz else 0) + 2 * channels, out_channels)
        self.operation = nn.Sequential(_ConvBNActRes(out_channels))


class LocalGrouper(nn.Module):
    """FPS centres + k-NN groups + normalised group features -> per-group feature (Compressor/layers.py:288-319)."""

    def __init__(self, in_channels, use_xyz=True, normalize="anchor"):
        super().__init__()
        self.use_xyz = use_xyz
        self.normalize = normalize.lower() if normalize is not None else None
        if self.normalize not in ("center", "anchor"):
            self.normalize = None
        if self.normalize is not None:
            add = 3 if use_xyz else 0
            self.affine_alpha = nn.Parameter(torch.ones([1, 1, 1, in_channels + add]))
            self.affine_beta = nn.Parameter(torch.zeros([1, 1, 1, in_channels + add]))
        self.extraction = _PreExtraction(in_channels, in_channels)

    def forward(self, xyz, feature, groups, k):
        """xyz [B,3,N], feature [B,D,N] -> (new_xyz [B,3,S], new_feature [B,D,S]); eval mode only (folded BatchNorm)."""
        from . import grouping
        if self.training:
            raise RuntimeError("ldt_b200.LocalGrouper is an inference path: call .eval() (BatchNorm is folded into the weights)")
        B = xyz.shape[0]
        with torch.no_grad():
            packed = grouping.pack_grouper(self)
            new_xyz, x = grouping.local_group(self, packed, self.normalize, xyz.transpose(1, 2), feature.transpose(1, 2), groups, k)
        return new_xyz.transpose(1, 2), x.reshape(B, groups, -1).permute(0, 2, 1)


class ConditionNet(nn.Module):
    """(pts_cond [B,hidden,32], img_cond [B,p_dim]) from {'img': [B,3,H,W], 'pts': [B,N,3]} (score.py:13-44)."""

    def __init__(self, hidden_size, p_dim, patch_size=16, img_condition=True, pt_condition=True):
        super().__init__()
        self.hidden_size, self.patch_size = hidden_size, patch_size
        self.img_condition, self.pt_condition = img_condition, pt_condition
        if pt_condition:
            self.pc_conv_in = nn.Conv1d(3, 128, 1)
            self.group = LocalGrouper(128, True, normalize="center")
            self.pc_conv_out = nn.Conv1d(128, hidden_size, 1)
        if img_condition:
            from torchvision import models
            trunk = models.resnet18(weights=None)
            self.resnet = nn.Sequential(*list(trunk.children())[:-4])   # stem + layer1 + layer2 -> 128 channels
            self.ln = nn.Linear(128, p_dim)
        self.conv_out = nn.Conv1d(hidden_size, hidden_size, 1)          # present in checkpoints, unused (score.py:29)

    def forward(self, condition):
        dev = self.conv_out.weight.device
        if dev.type != "cuda":
            raise RuntimeError("ldt_b200.ConditionNet runs on CUDA only (FPS / k-NN kernels have no CPU fallback)")
        pts_cond, img_cond = 0.0, 0.0
        if "img" in condition and self.img_condition:
            from . import grouping
            if self.resnet.training:
                raise RuntimeError("ldt_b200.ConditionNet is an inference path: call .eval() (BatchNorm is folded into the weights)")
            img = condition["img"].to(dev)
            with torch.no_grad():
                pooled = grouping.resnet_trunk_maxpool(self.resnet, img)             # adaptive_max_pool2d(resnet(img), 1)
                img_cond = grouping.conv_rows(pooled, grouping.pack_tf32(self.ln))    # self.ln(...)
            if img_cond.shape[0] == 1:
                img_cond = img_cond.squeeze(0)                                        # the reference's .squeeze()  score.py:35
        if "pts" in condition and self.pt_condition:
            from . import grouping
            if self.group.training:
                raise RuntimeError("ldt_b200.ConditionNet is an inference path: call .eval() (BatchNorm is folded into the weights)")
            pts = condition["pts"].to(dev).float().contiguous()            # [B, N, 3]
            B, N = pts.shape[0], pts.shape[1]
            with torch.no_grad():
                x = grouping.conv_rows(pts.reshape(B * N, 3), grouping.pack_tf32(self.pc_conv_in))
                k = x.shape[1] // self.patch_size * 2                          # score.py:40
                _, x = grouping.local_group(self.group, grouping.pack_grouper(self.group), self.group.normalize, pts,
                                            x.reshape(B, N, -1), self.patch_size, k)
                x = grouping.conv_rows(x, grouping.pack_tf32(self.pc_conv_out))   # [B*S, hidden]
            pts_cond = x.reshape(B, self.patch_size, -1).permute(0, 2, 1)   # [B, hidden, S] as the reference returns it
        return pts_cond, img_cond
