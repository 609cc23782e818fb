"""Drop-in for ``StructuralLosses.nn_distance`` (reference
evaluation/pytorch_structural_losses/StructuralLosses/nn_distance.py:7-41):
``nn_distance(a[b,n,3], b[b,m,3]) -> (dist1[b,n], dist2[b,m])``."""
import torch
from torch.autograd import Function

from StructuralLossesBackend import NNDistance


class NNDistanceFunction(Function):
    @staticmethod
    def forward(ctx, seta, setb):
        dist1, idx1, dist2, idx2 = NNDistance(seta, setb)
        ctx.idx1, ctx.idx2 = idx1, idx2
        ctx.mark_non_differentiable(dist1, dist2)
        return dist1, dist2

    @staticmethod
    def backward(ctx, grad_dist1, grad_dist2):
        raise NotImplementedError("nn_distance backward (training loss) is outside the ldt_b200 sampling/eval path")


nn_distance = NNDistanceFunction.apply
