"""Generation metrics on the Chamfer matrix -- host mirror of reference evaluation/evaluation_metrics.py.

``_pairwise_CD_`` (:165-198) becomes ONE kernel launch per matrix (``ldt_pairwise_cd``) instead of
2*N*ceil(N/batch) launches plus expand/mean/cat; ``lgan_mmd_cov`` (:234-246) and ``knn`` (:202-231) are
O(N^2) bookkeeping on the finished matrices and stay in torch, with the reference's exact op sequence so
cloud-level argmins agree.  ``batch_size`` is accepted for signature compatibility and ignored.
"""
from __future__ import annotations

import torch

from . import ops


def distChamferCUDA(x, y):
    """(dl, dr) as StructuralLosses.nn_distance returns them (:17-18)."""
    d1, _, d2, _ = ops.nn_distance_idx(x.contiguous(), y.contiguous())
    return d1, d2


def _pairwise_CD_(sample_pcs, ref_pcs, batch_size=None, verbose=True, rows=None):
    """[N_sample, N_ref] matrix, entry (i, j) = CD(sample i, ref j).  ``rows=(begin, end)`` computes a row block
    (used to shard the matrix over GPUs)."""
    a = sample_pcs.contiguous().float()
    b = ref_pcs.contiguous().float()
    begin, end = (0, a.shape[0]) if rows is None else rows
    return ops.pairwise_cd(a, b, begin, end)


def knn(Mxx, Mxy, Myy, k, sqrt=False):
    """1-NN two-sample test (:202-231)."""
    n0 = Mxx.size(0)
    n1 = Myy.size(0)
    label = torch.cat((torch.ones(n0), torch.zeros(n1))).to(Mxx)
    M = torch.cat((torch.cat((Mxx, Mxy), 1), torch.cat((Mxy.transpose(0, 1), Myy), 1)), 0)
    if sqrt:
        M = M.abs().sqrt()
    INFINITY = float("inf")
    val, idx = (M + torch.diag(INFINITY * torch.ones(n0 + n1).to(Mxx))).topk(k, 0, False)
    count = torch.zeros(n0 + n1).to(Mxx)
    for i in range(0, k):
        count = count + label.index_select(0, idx[i])
    pred = torch.ge(count, (float(k) / 2) * torch.ones(n0 + n1).to(Mxx)).float()
    s = {
        "tp": (pred * label).sum(),
        "fp": (pred * (1 - label)).sum(),
        "fn": ((1 - pred) * label).sum(),
        "tn": ((1 - pred) * (1 - label)).sum(),
    }
    s.update({
        "precision": s["tp"] / (s["tp"] + s["fp"] + 1e-10),
        "recall": s["tp"] / (s["tp"] + s["fn"] + 1e-10),
        "acc": torch.eq(label, pred).float().mean(),
    })
    return s


def lgan_mmd_cov(all_dist):
    """MMD / COV from an [N_sample, N_ref] distance matrix (:234-246)."""
    N_sample, N_ref = all_dist.size(0), all_dist.size(1)
    min_val_fromsmp, min_idx = torch.min(all_dist, dim=1)
    min_val, _ = torch.min(all_dist, dim=0)
    mmd = min_val.mean()
    cov = float(min_idx.unique().view(-1).size(0)) / float(N_ref)
    cov = torch.tensor(cov).to(all_dist)
    return {"mmd": mmd, "cov": cov}


def compute_CD_metrics(sample_pcs, ref_pcs, batch_size=None):
    """MMD-CD, COV-CD and 1-NNA-CD (:299-318)."""
    results = {}
    ref_pcs, sample_pcs = ref_pcs.cuda(), sample_pcs.cuda()
    M_rs_cd = _pairwise_CD_(ref_pcs, sample_pcs, batch_size)
    res_cd = lgan_mmd_cov(M_rs_cd.t())
    results.update({"%s-CD" % k: v for k, v in res_cd.items()})
    M_rr_cd = _pairwise_CD_(ref_pcs, ref_pcs, batch_size)
    M_ss_cd = _pairwise_CD_(sample_pcs, sample_pcs, batch_size)
    one_nn_cd_res = knn(M_rr_cd, M_rs_cd, M_ss_cd, 1, sqrt=False)
    results.update({"1-NN-CD-%s" % k: v for k, v in one_nn_cd_res.items() if "acc" in k})
    return results
