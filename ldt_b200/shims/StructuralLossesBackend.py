"""Drop-in for the reference's compiled torch extension ``StructuralLossesBackend``
(evaluation/pytorch_structural_losses/pybind/bind.cpp:10-15).  Put ``ldt_b200/shims`` on ``sys.path`` and the
reference's ``from StructuralLossesBackend import NNDistance, ApproxMatch, MatchCost`` resolves here.
"""
import torch as _torch

from ldt_b200 import ops as _ops


def NNDistance(set_d, set_q):
    """-> [dist1, idx1(int32), dist2, idx2(int32)]  (src/structural_loss.cpp:80-99)."""
    return list(_ops.nn_distance_idx(set_d, set_q))


def ApproxMatch(set_d, set_q):
    """-> [match [b,m,n], temp [b,(n+m)*2]]  (src/structural_loss.cpp:14-44).  ``temp`` is the reference kernel's
    scratch; this implementation keeps that state in registers, so it is returned zero-filled for shape compatibility."""
    match = _ops.approx_match(set_d, set_q)
    temp = _torch.zeros((set_d.shape[0], (set_d.shape[1] + set_q.shape[1]) * 2), dtype=_torch.float32, device=set_d.device)
    return [match, temp]


def MatchCost(set_d, set_q, match):
    """-> cost [b]  (src/structural_loss.cpp:46-78)."""
    return _ops.match_cost_from_match(set_d, set_q, match)


def _backward_only(name):
    def fn(*a, **k):
        raise NotImplementedError(f"{name}: training-loss backward kernels are outside the ldt_b200 sampling/eval path")
    return fn


NNDistanceGrad = _backward_only("NNDistanceGrad")
MatchCostGrad = _backward_only("MatchCostGrad")
