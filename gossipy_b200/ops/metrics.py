"""Classification / clustering metrics computed from device-side sufficient statistics.

The reference ships every prediction to the host and calls scikit-learn
(``gossipy/model/handler.py:282-334, 375-391, 632-636``).  Here evaluation kernels reduce to a
``C x C`` confusion matrix (or a contingency table) on the device; these functions turn that
tiny table into the same numbers sklearn reports (macro averages over the labels that occur in
``y_true`` or ``y_pred``, ``zero_division=0``).  ``tests/test_metrics.py`` checks parity.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch


def classification_report(cm: np.ndarray) -> Dict[str, float]:
    """accuracy / macro precision / recall / f1 from a confusion matrix ``cm[true, pred]``."""
    cm = np.asarray(cm, dtype=np.float64)
    total = cm.sum()
    tp = np.diag(cm)
    pred_tot, true_tot = cm.sum(axis=0), cm.sum(axis=1)
    present = (pred_tot + true_tot) > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        prec = np.where(pred_tot > 0, tp / pred_tot, 0.0)
        rec = np.where(true_tot > 0, tp / true_tot, 0.0)
        f1 = np.where(prec + rec > 0, 2 * prec * rec / (prec + rec), 0.0)
    k = max(int(present.sum()), 1)
    return {"accuracy": float(tp.sum() / total) if total else 0.0,
            "precision": float(prec[present].sum() / k),
            "recall": float(rec[present].sum() / k),
            "f1_score": float(f1[present].sum() / k)}


@torch.no_grad()
def roc_auc(y_pos: torch.Tensor, scores: torch.Tensor) -> float:
    """Area under the ROC curve with tie handling (Mann-Whitney U with average ranks).

    ``y_pos`` is a boolean mask of the positive class.  Returns 0.5 when only one class is
    present (the reference logs a warning and reports 0.5, ``handler.py:325-331``).
    """
    y_pos = y_pos.reshape(-1).bool()
    scores = scores.reshape(-1).double()
    n_pos = int(y_pos.sum())
    n_neg = y_pos.numel() - n_pos
    if n_pos == 0 or n_neg == 0:
        return 0.5
    order = torch.argsort(scores)
    s_sorted = scores[order]
    _, inv, counts = torch.unique_consecutive(s_sorted, return_inverse=True, return_counts=True)
    ends = torch.cumsum(counts, 0).double()
    avg_rank = ends - (counts.double() - 1) / 2.0  # 1-based average rank of each tie group
    ranks = avg_rank[inv]
    rank_sum_pos = ranks[y_pos[order]].sum()
    u = rank_sum_pos - n_pos * (n_pos + 1) / 2.0
    return float(u / (n_pos * n_neg))


def nmi_from_contingency(ct: np.ndarray) -> float:
    """Normalised mutual information (arithmetic-mean normalisation, sklearn's default)."""
    ct = np.asarray(ct, dtype=np.float64)
    n = ct.sum()
    if n == 0:
        return 1.0
    pi, pj = ct.sum(axis=1), ct.sum(axis=0)
    if (pi > 0).sum() == 1 and (pj > 0).sum() == 1:
        return 1.0
    nz = ct > 0
    outer = np.outer(pi, pj)
    mi = float((ct[nz] / n * (np.log(ct[nz] * n) - np.log(outer[nz]))).sum())
    mi = max(mi, 0.0)

    def ent(p):
        p = p[p > 0] / n
        return float(-(p * np.log(p)).sum())
    hu, hv = ent(pi), ent(pj)
    denom = (hu + hv) / 2.0
    return float(mi / denom) if denom > 0 else 1.0
