"""Drop-in for the reference's compiled torch extension ``StructuralLossesBackend``
(evaluation/pytorch_structural_losses/pybind/bind.cpp:10-15).  Put ``ldt_b200/shims`` on ``sys.path`` and the
reference's ``from StructuralLossesBackend import NNDistance`` resolves here.
"""
from ldt_b200 import ops as _ops


def NNDistance(set_d, set_q):
    """-> [dist1, idx1(int32), dist2, idx2(int32)]  (src/structural_loss.cpp:80-99)."""
    return list(_ops.nn_distance_idx(set_d, set_q))


def _backward_only(name):
    def fn(*a, **k):
        raise NotImplementedError(f"{name}: training-loss backward kernels are outside the ldt_b200 sampling/eval path")
    return fn


NNDistanceGrad = _backward_only("NNDistanceGrad")
MatchCostGrad = _backward_only("MatchCostGrad")


def ApproxMatch(set_d, set_q):
    raise NotImplementedError("ApproxMatch (approximate EMD) is the next row of the hot-path scope (SURVEY.md 8f1)")


def MatchCost(set_d, set_q, match):
    raise NotImplementedError("MatchCost (approximate EMD) is the next row of the hot-path scope (SURVEY.md 8f1)")
