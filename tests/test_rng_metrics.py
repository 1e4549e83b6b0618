"""Keyed permutation and metric parity with scikit-learn."""
import numpy as np
import pytest
import torch

from gossipy_b200.engine import rng
from gossipy_b200.ops import metrics, torch_ref


@pytest.mark.parametrize("n", [1, 2, 3, 7, 64, 100, 517, 7500])
def test_perm_is_permutation_and_matches_scalar(n):
    key = rng.mix64(12345 + n)
    p = torch_ref.perm_indices(n, key)
    assert sorted(p.tolist()) == list(range(n))
    for i in (0, n // 2, n - 1):
        assert rng.feistel_perm(i, n, key) == p[i]
    if n > 50:
        assert not np.array_equal(p, np.arange(n))
        assert not np.array_equal(p, torch_ref.perm_indices(n, key + 1))


def test_classification_report_matches_sklearn():
    from sklearn.metrics import accuracy_score, f1_score, precision_score, recall_score
    rs = np.random.RandomState(0)
    for c, absent in ((2, False), (5, False), (10, True)):
        y = rs.randint(0, c, 400)
        p = rs.randint(0, c, 400)
        if absent:
            y[y == 3] = 4
            p[p == 3] = 4          # class 3 occurs nowhere -> excluded from the macro mean
            p[p == 7] = 1          # class 7 never predicted -> precision 0 (zero_division)
        cm = torch_ref.confusion_matrix(torch.tensor(y), torch.tensor(p), c).numpy()
        r = metrics.classification_report(cm)
        assert r["accuracy"] == pytest.approx(accuracy_score(y, p))
        for k, fn in (("precision", precision_score), ("recall", recall_score), ("f1_score", f1_score)):
            assert r[k] == pytest.approx(fn(y, p, zero_division=0, average="macro")), k


def test_auc_with_ties_matches_sklearn():
    from sklearn.metrics import roc_auc_score
    rs = np.random.RandomState(1)
    y = rs.randint(0, 2, 500)
    s = np.round(rs.randn(500) + y, 1)   # many ties
    assert metrics.roc_auc(torch.tensor(y == 1), torch.tensor(s)) == pytest.approx(roc_auc_score(y, s))
    assert metrics.roc_auc(torch.ones(5, dtype=torch.bool), torch.arange(5.)) == 0.5


def test_nmi_matches_sklearn():
    from sklearn.metrics.cluster import normalized_mutual_info_score
    rs = np.random.RandomState(2)
    a, b = rs.randint(0, 3, 300), rs.randint(0, 4, 300)
    b[:100] = a[:100]
    ct = np.zeros((3, 4))
    np.add.at(ct, (a, b), 1)
    assert metrics.nmi_from_contingency(ct) == pytest.approx(normalized_mutual_info_score(a, b))
