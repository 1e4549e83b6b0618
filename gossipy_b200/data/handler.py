"""Concrete data handlers (behavioural reference: ``gossipy/data/handler.py:25-245``)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from . import DataHandler

__all__ = ["ClassificationDataHandler", "ClusteringDataHandler", "RegressionDataHandler",
           "RecSysDataHandler"]

ArrayLike = Union[np.ndarray, torch.Tensor]


def _take(X: ArrayLike, idx: Any) -> ArrayLike:
    if isinstance(X, torch.Tensor):
        if isinstance(idx, np.ndarray):
            idx = torch.from_numpy(np.ascontiguousarray(idx)).long()
        elif isinstance(idx, (list, tuple, range)):
            idx = torch.as_tensor(list(idx), dtype=torch.long)
        return X[idx]
    return X[idx]


class ClassificationDataHandler(DataHandler):
    """Feature/label container with a train and an (optional) evaluation split.

    When no evaluation data is given and ``test_size > 0`` the data is split with a private
    generator seeded by ``seed`` (tensors and arrays alike).
    """

    def __init__(self, X: ArrayLike, y: ArrayLike, X_te: Optional[ArrayLike] = None,
                 y_te: Optional[ArrayLike] = None, test_size: float = 0.2, seed: int = 42) -> None:
        assert 0 <= test_size < 1
        assert isinstance(X, (torch.Tensor, np.ndarray))
        if test_size > 0 and (X_te is None or y_te is None):
            n = X.shape[0]
            n_te = int(round(n * test_size))
            from .. import GlobalSettings
            if GlobalSettings().reference_compat and not isinstance(X, torch.Tensor):
                # arrays: the reference delegates to scikit-learn (data/handler.py:69-72)
                from sklearn.model_selection import train_test_split
                self.Xtr, self.Xte, self.ytr, self.yte = train_test_split(X, y, test_size=test_size, random_state=seed,
                                                                          shuffle=True)
            else:
                if GlobalSettings().reference_compat:
                    # tensors: the reference re-seeds torch's GLOBAL stream and cuts a randperm (data/handler.py:60-67)
                    torch.manual_seed(seed)
                    order = torch.randperm(n).numpy()
                else:
                    order = np.random.default_rng(seed).permutation(n)
                tr_ids, te_ids = order[:n - n_te], order[n - n_te:]
                self.Xtr, self.ytr = _take(X, tr_ids), _take(y, tr_ids)
                self.Xte, self.yte = _take(X, te_ids), _take(y, te_ids)
        else:
            self.Xtr, self.ytr = X, y
            self.Xte, self.yte = X_te, y_te
        ytr = self.ytr.detach().cpu().numpy() if isinstance(self.ytr, torch.Tensor) else self.ytr
        self.n_classes = len(np.unique(ytr))

    def __getitem__(self, idx: Any) -> Tuple[ArrayLike, Any]:
        return _take(self.Xtr, idx), _take(self.ytr, idx)

    def at(self, idx: Any, eval_set: bool = False) -> Any:
        if not eval_set:
            return self[idx]
        if isinstance(idx, (list, tuple, np.ndarray)) and len(idx) == 0:
            return None
        return _take(self.Xte, idx), _take(self.yte, idx)

    def size(self, dim: int = 0) -> int:
        return self.Xtr.shape[dim]

    def get_train_set(self) -> Tuple[Any, Any]:
        return self.Xtr, self.ytr

    def get_eval_set(self) -> Tuple[Any, Any]:
        return self.Xte, self.yte

    def eval_size(self) -> int:
        return self.Xte.shape[0] if self.Xte is not None else 0

    def __repr__(self) -> str:
        return str(self)

    def __str__(self) -> str:
        return "%s(size_tr=%d, size_te=%d, n_feats=%d, n_classes=%d)" % (
            self.__class__.__name__, self.size(), self.eval_size(), self.size(1), self.n_classes)


class ClusteringDataHandler(ClassificationDataHandler):
    """Evaluation happens on the training data itself.

    FIX(B18): the reference passes ``0`` as ``X_te`` positionally (``data/handler.py:153``) so
    the default 80/20 split still happens; here all data stays in the training set.
    """

    def __init__(self, X: ArrayLike, y: ArrayLike) -> None:
        super().__init__(X, y, test_size=0.0)

    def get_eval_set(self) -> Tuple[Any, Any]:
        return self.get_train_set()

    def eval_size(self) -> int:
        return self.size()

    def __str__(self) -> str:
        return "%s(size=%d)" % (self.__class__.__name__, self.size())


class RegressionDataHandler(ClassificationDataHandler):
    """Same container with real-valued targets.  FIX(B19): ``at`` returns its value."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.n_classes = 0


class RecSysDataHandler(DataHandler):
    """Per-user rating lists with a per-user train/test cut (ref ``data/handler.py:181-245``)."""

    def __init__(self, ratings: Dict[int, List[Tuple[int, float]]], n_users: int, n_items: int,
                 test_size: float = 0.2, seed: int = 42) -> None:
        self.n_users, self.n_items = n_users, n_items
        from .. import GlobalSettings
        compat = GlobalSettings().reference_compat
        if compat:
            np.random.seed(seed)        # the reference re-seeds NumPy's GLOBAL stream here (data/handler.py:211)
        rng = np.random.default_rng(seed)
        self.ratings: Dict[int, np.ndarray] = {}
        self.test_id: List[int] = []
        for u in range(len(ratings)):
            arr = np.asarray(ratings[u], dtype=np.float64).reshape(-1, 2)
            self.test_id.append(max(1, int(len(arr) * (1 - test_size))))
            self.ratings[u] = np.random.permutation(arr) if compat else arr[rng.permutation(len(arr))]

    def __getitem__(self, idx: int) -> np.ndarray:
        return self.ratings[idx][:self.test_id[idx]]

    def at(self, idx: int, eval_set: bool = False) -> np.ndarray:
        return self.ratings[idx][self.test_id[idx]:] if eval_set else self[idx]

    def size(self, dim: int = 0) -> int:
        return self.n_users

    def get_train_set(self) -> Dict[int, np.ndarray]:
        return {u: self[u] for u in range(self.n_users)}

    def get_eval_set(self) -> Dict[int, np.ndarray]:
        return {u: self.at(u, True) for u in range(self.n_users)}

    def eval_size(self) -> int:
        return 0

    def __str__(self) -> str:
        n_rat = sum(len(self.ratings[u]) for u in range(self.n_users))
        return "%s(n_users=%d, n_items=%d, n_ratings=%d)" % (
            self.__class__.__name__, self.n_users, self.n_items, n_rat)
