"""Generation metrics on pairwise cloud-distance matrices -- host mirror of reference
evaluation/evaluation_metrics.py (same function names, arguments and result keys).

What changes underneath:
* ``_pairwise_CD_`` (:165-198) and ``_pairwise_EMD_CD_`` (:112-162) loop in Python over every row cloud and every
  column batch (2 kernel launches + expand/contiguous/mean/cat per iteration); here a whole matrix is ONE launch of
  ``ldt_pairwise_cd`` / ``ldt_pairwise_emd``.  ``batch_size`` and the ``accelerated_*`` flags are accepted for signature
  compatibility and ignored (there is only the CUDA path).
* ``lgan_mmd_cov`` (:234-246) and ``knn`` (:202-231) are O(N^2) bookkeeping on finished matrices and stay in torch;
  they are written so that every reduction the reference takes (column/row minima, first-occurrence argmin, the 1-NN
  vote) sees identical inputs and therefore returns identical numbers.
"""
from __future__ import annotations

import torch

from . import ops


def distChamferCUDA(x, y):
    """(dl, dr) as StructuralLosses.nn_distance returns them (:17-18)."""
    d1, _, d2, _ = ops.nn_distance_idx(x.contiguous(), y.contiguous())
    return d1, d2


def emd_approx_cuda(sample, ref):
    """Approximate EMD per pair, divided by the point count (:39-45)."""
    B, N, N_ref = sample.size(0), sample.size(1), ref.size(1)
    assert N == N_ref, "Not sure what would EMD do in this case"
    return ops.match_cost(sample.contiguous(), ref.contiguous()) / float(N)


def _rows(rows, n):
    return (0, n) if rows is None else rows


def _same_set(a, b) -> bool:
    """Both arguments are the same cloud set (the M_rr / M_ss calls, :311-312): the matrix is symmetric."""
    return a is b or (a.shape == b.shape and a.dtype == b.dtype and a.device == b.device and a.data_ptr() == b.data_ptr()
                      and a.stride() == b.stride())


def _pairwise_CD_(sample_pcs, ref_pcs, batch_size=None, verbose=True, rows=None):
    """[N_sample, N_ref] matrix, entry (i, j) = CD(sample i, ref j).  ``rows=(begin, end)`` computes a row block
    (used to shard the matrix over GPUs).

    When both arguments are the same set (M_rr, M_ss) and the whole matrix is asked for, only the entries on or above the
    diagonal are evaluated and mirrored: the kernel's (i, j) and (j, i) values are bit-identical (csrc/nn_distance.cu), so
    the result equals the full evaluation the reference performs -- at half the pair evaluations.  The approximate EMD is
    NOT symmetric (ApproxMatch treats its two sets differently, approxmatch.cu:3-182), so ``_pairwise_EMD_CD_`` mirrors
    only its CD half."""
    same = _same_set(sample_pcs, ref_pcs)
    a = sample_pcs.contiguous().float()
    begin, end = _rows(rows, a.shape[0])
    if same and rows is None:
        return ops.mirror_upper(ops.pairwise_cd_upper(a))
    b = a if same else ref_pcs.contiguous().float()
    return ops.pairwise_cd(a, b, begin, end)


def _pairwise_EMD_CD_(sample_pcs, ref_pcs, batch_size=None, accelerated_cd=True, accelerated_emd=True, rows=None):
    """(CD matrix, EMD matrix), both [N_sample, N_ref]."""
    a = sample_pcs.contiguous().float()
    b = ref_pcs.contiguous().float()
    begin, end = _rows(rows, a.shape[0])
    return _pairwise_CD_(sample_pcs, ref_pcs, rows=rows), ops.pairwise_emd(a, b, begin, end)


def knn(Mxx, Mxy, Myy, k, sqrt=False):
    """1-NN two-sample test (:202-231): leave-one-out k-NN classification of the union of both sets."""
    n0, n1 = Mxx.size(0), Myy.size(0)
    n = n0 + n1
    dist = torch.cat([torch.cat([Mxx, Mxy], dim=1), torch.cat([Mxy.t(), Myy], dim=1)], dim=0)
    if sqrt:
        dist = dist.abs().sqrt()
    dist = dist + torch.diag(torch.full((n,), float("inf"), device=dist.device, dtype=dist.dtype))  # exclude self
    nearest = dist.topk(k, dim=0, largest=False).indices                      # [k, n]
    label = (torch.arange(n, device=dist.device) < n0).to(Mxx.dtype)          # 1 = first set
    votes = label[nearest].sum(dim=0)
    pred = (votes >= k / 2.0).to(Mxx.dtype)
    stats = {
        "tp": (pred * label).sum(),
        "fp": (pred * (1 - label)).sum(),
        "fn": ((1 - pred) * label).sum(),
        "tn": ((1 - pred) * (1 - label)).sum(),
    }
    stats["precision"] = stats["tp"] / (stats["tp"] + stats["fp"] + 1e-10)
    stats["recall"] = stats["tp"] / (stats["tp"] + stats["fn"] + 1e-10)
    stats["acc"] = (label == pred).to(Mxx.dtype).mean()
    return stats


def lgan_mmd_cov(all_dist):
    """MMD / COV from an [N_sample, N_ref] distance matrix (:234-246)."""
    n_ref = all_dist.size(1)
    covered = all_dist.argmin(dim=1)               # which reference cloud each sample is closest to
    mmd = all_dist.min(dim=0).values.mean()        # every reference cloud's distance to its closest sample
    cov = torch.tensor(covered.unique().numel() / float(n_ref)).to(all_dist)
    return {"mmd": mmd, "cov": cov}


def _update(results, res, suffix, only_acc=False):
    for k, v in res.items():
        if only_acc and "acc" not in k:
            continue
        results[("1-NN-%s-%s" % (suffix, k)) if only_acc else ("%s-%s" % (k, suffix))] = v


def compute_all_metrics(sample_pcs, ref_pcs, batch_size=None, accelerated_cd=True, accelerated_emd=True):
    """MMD / COV / 1-NNA for both CD and approximate EMD (:249-277)."""
    results = {}
    ref_pcs, sample_pcs = ref_pcs.cuda(), sample_pcs.cuda()
    M_rs_cd, M_rs_emd = _pairwise_EMD_CD_(ref_pcs, sample_pcs, batch_size)
    _update(results, lgan_mmd_cov(M_rs_cd.t()), "CD")
    _update(results, lgan_mmd_cov(M_rs_emd.t()), "EMD")
    M_rr_cd, M_rr_emd = _pairwise_EMD_CD_(ref_pcs, ref_pcs, batch_size)
    M_ss_cd, M_ss_emd = _pairwise_EMD_CD_(sample_pcs, sample_pcs, batch_size)
    _update(results, knn(M_rr_cd, M_rs_cd, M_ss_cd, 1, sqrt=False), "CD", only_acc=True)
    _update(results, knn(M_rr_emd, M_rs_emd, M_ss_emd, 1, sqrt=False), "EMD", only_acc=True)
    return results


def compute_MMD_metrics(sample_pcs, ref_pcs, batch_size=None, accelerated_cd=True, accelerated_emd=True):
    """MMD / COV for CD and EMD only (:280-296)."""
    results = {}
    ref_pcs, sample_pcs = ref_pcs.cuda(), sample_pcs.cuda()
    M_rs_cd, M_rs_emd = _pairwise_EMD_CD_(ref_pcs, sample_pcs, batch_size)
    _update(results, lgan_mmd_cov(M_rs_cd.t()), "CD")
    _update(results, lgan_mmd_cov(M_rs_emd.t()), "EMD")
    return results


def compute_CD_metrics(sample_pcs, ref_pcs, batch_size=None):
    """MMD-CD, COV-CD and 1-NNA-CD (:299-318)."""
    results = {}
    ref_pcs, sample_pcs = ref_pcs.cuda(), sample_pcs.cuda()
    M_rs_cd = _pairwise_CD_(ref_pcs, sample_pcs, batch_size)
    _update(results, lgan_mmd_cov(M_rs_cd.t()), "CD")
    M_rr_cd = _pairwise_CD_(ref_pcs, ref_pcs, batch_size)
    M_ss_cd = _pairwise_CD_(sample_pcs, sample_pcs, batch_size)
    _update(results, knn(M_rr_cd, M_rs_cd, M_ss_cd, 1, sqrt=False), "CD", only_acc=True)
    return results


# ------------------------------------------------------------------------------------------------------------------
# Completion metrics (reference completion_trainer/Latent_SDE_Trainer.py:41-53), on the same NN kernel
# ------------------------------------------------------------------------------------------------------------------
def L2_ChamferEval_1000(array1, array2):
    """1000 * (mean NN distance array1 -> array2 + mean NN distance array2 -> array1) over the whole batch."""
    dist1, dist2 = distChamferCUDA(array1.float(), array2.float())
    return (torch.mean(dist1) + torch.mean(dist2)) * 1000


def F1Score(array1, array2, threshold=0.001):
    """(fscore [B], precision_1 [B], precision_2 [B]) at a squared-distance threshold; 0/0 -> 0 as the reference."""
    dist1, dist2 = distChamferCUDA(array1.float(), array2.float())
    precision_1 = torch.mean((dist1 < threshold).float(), dim=1)
    precision_2 = torch.mean((dist2 < threshold).float(), dim=1)
    fscore = 2 * precision_1 * precision_2 / (precision_1 + precision_2)
    fscore[torch.isnan(fscore)] = 0
    return fscore, precision_1, precision_2
