"""Drop-in for ``StructuralLosses.match_cost`` (approximate EMD, reference StructuralLosses/match_cost.py:6-45):
``match_cost(seta[b,n,3], setb[b,m,3]) -> cost[b]``.  Forward only (evaluation); the fused kernel never writes the
dense match matrix the reference materialises between ApproxMatch and MatchCost."""
from ldt_b200 import ops as _ops


def match_cost(seta, setb):
    return _ops.match_cost(seta.contiguous(), setb.contiguous())
