"""Placeholder for ``StructuralLosses.match_cost`` (approximate EMD, reference StructuralLosses/match_cost.py:6-45).
Importing succeeds so ``evaluation_metrics`` binds its CUDA path; calling raises until SURVEY.md row 8f1 is built."""


def match_cost(seta, setb):
    raise NotImplementedError("match_cost (approximate EMD) is the next row of the hot-path scope (SURVEY.md 8f1)")
