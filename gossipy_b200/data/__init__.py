"""Data layer: handlers, sample->client assignment strategies, dispatchers, loaders.

Behavioural reference: ``gossipy/data/__init__.py`` (cited per item).  The dispatcher hands
every node ``((X_train, y_train), (X_test, y_test) | None)`` exactly like the reference; the
device engine uploads each node's shard to HBM once (``engine.shards``) and the fused training
kernels index it directly with a keyed permutation, so no shuffled copy is ever materialised.
There is no network on the target machines: :mod:`gossipy_b200.data.synthetic` generates data of
the named shapes (spambase / MNIST / CIFAR-10 / MovieLens) and the loaders fall back to it.
"""
from __future__ import annotations

import os
import shutil
from abc import ABC, abstractmethod
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .. import LOG
from ..utils import download_and_untar, download_and_unzip

__all__ = ["DataHandler", "AssignmentHandler", "DataDispatcher", "RecSysDataDispatcher",
           "load_classification_dataset", "load_recsys_dataset", "get_CIFAR10",
           "get_FashionMNIST", "get_FEMNIST"]

UCI_BASE_URL = "https://archive.ics.uci.edu/ml/machine-learning-databases/"
UCI_URL_AND_CLASS = {
    "spambase": (UCI_BASE_URL + "spambase/spambase.data", 57),
    "sonar": (UCI_BASE_URL + "undocumented/connectionist-bench/sonar/sonar.all-data", 60),
    "ionosphere": (UCI_BASE_URL + "ionosphere/ionosphere.data", 34),
    "abalone": (UCI_BASE_URL + "abalone/abalone.data", 0),
    "banknote": (UCI_BASE_URL + "00267/data_banknote_authentication.txt", 4),
}


class DataHandler(ABC):
    """Train/eval container interface (ref ``data/__init__.py:55-161``)."""

    @abstractmethod
    def __getitem__(self, idx: Union[int, List[int]]) -> Any: ...

    @abstractmethod
    def at(self, idx: Union[int, List[int]], eval_set: bool = False) -> Any: ...

    @abstractmethod
    def size(self, dim: int = 0) -> int: ...

    @abstractmethod
    def get_eval_set(self) -> Tuple[Any, Any]: ...

    @abstractmethod
    def get_train_set(self) -> Tuple[Any, Any]: ...

    @abstractmethod
    def eval_size(self) -> int: ...


def _labels_np(y: Any) -> np.ndarray:
    return y.detach().cpu().numpy() if isinstance(y, torch.Tensor) else np.asarray(y)


class _GlobalNumpyStream:
    """The legacy global NumPy stream behind the few ``Generator`` methods the strategies use."""

    permutation = staticmethod(np.random.permutation)
    power = staticmethod(np.random.power)
    shuffle = staticmethod(np.random.shuffle)
    choice = staticmethod(np.random.choice)
    dirichlet = staticmethod(np.random.dirichlet)

    @staticmethod
    def integers(low: int, high: int) -> int:
        return int(np.random.randint(low, high))


class AssignmentHandler:
    """Strategies mapping samples to ``n`` clients (ref ``data/__init__.py:164-373``).

    Every strategy returns a list of ``n`` index arrays.  Draws come from a private
    ``numpy`` Generator seeded with ``seed`` (the reference reseeds the *global* numpy/torch
    RNGs as a side effect, which silently couples data assignment to the simulation's RNG);
    with ``GlobalSettings().reference_compat`` the reference's draws are issued in the reference's
    order on the global streams instead, which reproduces its index sets for a given seed.
    """

    def __init__(self, seed: int) -> None:
        from .. import GlobalSettings
        self.compat = bool(GlobalSettings().reference_compat)
        if self.compat:
            # the reference re-seeds the GLOBAL torch / NumPy streams here and draws from the global NumPy stream
            # (``data/__init__.py:166-168``); under ``reference_compat`` the strategies below issue the same draws in the
            # same order, so a given seed yields the reference's index sets
            torch.manual_seed(seed)
            np.random.seed(seed)
            self.rng: Any = _GlobalNumpyStream()
        else:
            self.rng = np.random.default_rng(seed)

    # -- IID ------------------------------------------------------------------------
    def uniform(self, y: Any, n: int) -> List[np.ndarray]:
        """Equal-size random shards; the remainder ``len(y) % n`` is dropped."""
        per = len(y) // n
        order = self.rng.permutation(len(y))
        return [order[i * per:(i + 1) * per] for i in range(n)]

    # -- quantity skew --------------------------------------------------------------
    def _power_law_owner(self, count: int, n: int, alpha: float) -> np.ndarray:
        return np.minimum((self.rng.power(alpha, count) * n).astype(int), n - 1)

    def quantity_skew(self, y: Any, n: int, min_quantity: int = 2,
                      alpha: float = 4.) -> List[np.ndarray]:
        """Power-law shard sizes, each client gets at least ``min_quantity`` samples."""
        m = len(y)
        assert min_quantity > 0, "min_quantity must be >= 1"
        assert min_quantity * n <= m, "# of instances must be > than min_quantity*n"
        owner = np.concatenate([self._power_law_owner(m - min_quantity * n, n, alpha),
                                np.repeat(np.arange(n), min_quantity)])
        self.rng.shuffle(owner)
        return [np.flatnonzero(owner == i) for i in range(n)]

    def classwise_quantity_skew(self, y: Any, n: int, min_quantity: int = 2,
                                alpha: float = 4.) -> List[np.ndarray]:
        """Power-law skew applied independently inside every class."""
        yy = _labels_np(y)
        assert min_quantity > 0 and min_quantity * n <= len(yy)
        out: List[List[int]] = [[] for _ in range(n)]
        classes = [np.flatnonzero(yy == c) for c in np.unique(yy)]
        assert all(len(ids) >= n for ids in classes), "Under represented class!"
        skew = [self._power_law_owner(len(ids) - n, n, alpha) for ids in classes]     # all power-law draws first ...
        for ids, sk in zip(classes, skew):
            owner = np.concatenate([sk, np.arange(n)])
            self.rng.shuffle(owner)                                                   # ... then one shuffle per class
            for i in range(n):
                out[i].extend(ids[owner == i].tolist())
        return [np.array(o, dtype=int) for o in out]

    # -- label skew -----------------------------------------------------------------
    def label_quantity_skew(self, y: Any, n: int, class_per_client: int = 2) -> List[np.ndarray]:
        """Each client sees exactly ``class_per_client`` classes; all classes are covered."""
        yy = _labels_np(y)
        labels = np.unique(yy)
        L = len(labels)
        assert 0 < class_per_client <= L, "class_per_client must be > 0 and <= #classes"
        assert class_per_client * n >= L, "class_per_client * n must be >= #classes"
        picks = [self.rng.choice(L, class_per_client, replace=False) for _ in range(n)]
        while True:
            covered = set(int(c) for p in picks for c in p)
            missing = [c for c in range(L) if c not in covered]
            if not missing:
                break
            for c in missing:
                u = int(self.rng.integers(0, n))
                if self.compat:      # the reference overwrites a random slot (may evict the only holder of another class)
                    picks[u][int(self.rng.integers(0, class_per_client))] = c
                    continue
                # only overwrite a slot whose class some OTHER client holds too: covering c must not uncover another
                # class (class_per_client * n >= L guarantees such a slot exists somewhere; retry with another client)
                holders = np.bincount(np.concatenate(picks), minlength=L)
                free = [s for s in range(class_per_client) if holders[picks[u][s]] > 1]
                if free:
                    picks[u][int(self.rng.choice(free))] = c
        owner = np.zeros(len(yy), dtype=int)
        for ci, lbl in enumerate(labels):
            users = [u for u in range(n) if ci in picks[u]]
            ids = np.flatnonzero(yy == lbl)
            owner[ids] = self.rng.choice(users, len(ids))
        return [np.flatnonzero(owner == i) for i in range(n)]

    def label_dirichlet_skew(self, y: Any, n: int, beta: float = .1) -> List[np.ndarray]:
        """Per-class client proportions ~ Dirichlet(beta); every client gets >=1 per class."""
        assert beta > 0, "beta must be > 0"
        yy = _labels_np(y)
        owner = np.zeros(len(yy), dtype=int)
        classes = np.unique(yy)
        props = [self.rng.dirichlet([beta] * n) for _ in classes]         # all proportions first (one draw per class)
        for c, p in zip(classes, props):
            ids = np.flatnonzero(yy == c)
            self.rng.shuffle(ids)
            self.rng.shuffle(p)                                           # (exchangeable; the reference does it too)
            if len(ids) > n:
                owner[ids[n:]] = self.rng.choice(n, size=len(ids) - n, p=p)
            owner[ids[:n]] = np.arange(n)[:len(ids[:n])]
        return [np.flatnonzero(owner == i) for i in range(n)]

    def label_pathological_skew(self, y: Any, n: int,
                                shards_per_client: int = 2) -> List[np.ndarray]:
        """McMahan et al.: sort by label, cut into ``n*shards_per_client`` shards, deal them."""
        yy = _labels_np(y)
        order = np.argsort(yy) if self.compat else np.argsort(yy, kind="stable")    # (ties: the reference's default sort)
        n_shards = int(shards_per_client * n)
        shard = int(np.ceil(len(yy) / n_shards))
        deal = self.rng.permutation(n_shards)
        out = []
        for i in range(n):
            mine = deal[i * shards_per_client:(i + 1) * shards_per_client]
            out.append(np.sort(np.concatenate([order[s * shard:min((s + 1) * shard, len(yy))]
                                               for s in mine])))
        return out


class DataDispatcher:
    """Splits a :class:`DataHandler` over ``n`` clients (ref ``data/__init__.py:376-510``)."""

    def __init__(self, data_handler: DataHandler, n: int = 0, eval_on_user: bool = True,
                 auto_assign: bool = True) -> None:
        assert data_handler.size() >= n
        if n <= 1:
            n = data_handler.size()  # one sample per client
        self.data_handler = data_handler
        self.n = n
        self.eval_on_user = eval_on_user
        self.tr_assignments: Optional[List[Any]] = None
        self.te_assignments: Optional[List[Any]] = None
        if auto_assign:
            self.assign()

    def set_assignments(self, tr_assignments: List[Any],
                        te_assignments: Optional[List[Any]]) -> None:
        assert len(tr_assignments) == self.n
        assert not te_assignments or len(te_assignments) == self.n
        self.tr_assignments = tr_assignments
        self.te_assignments = te_assignments if te_assignments else [[] for _ in range(self.n)]

    def assign(self, seed: Optional[int] = 42) -> None:
        ah = AssignmentHandler(seed)
        self.tr_assignments = ah.uniform(self.data_handler.ytr, self.n)
        if self.eval_on_user:
            self.te_assignments = ah.uniform(self.data_handler.yte, self.n)
        else:
            self.te_assignments = [[] for _ in range(self.n)]

    def __getitem__(self, idx: int) -> Any:
        assert 0 <= idx < self.n, "Index %d out of range." % idx
        return (self.data_handler.at(self.tr_assignments[idx]),
                self.data_handler.at(self.te_assignments[idx], True))

    def size(self) -> int:
        return self.n

    def get_eval_set(self) -> Tuple[Any, Any]:
        return self.data_handler.get_eval_set()

    def has_test(self) -> bool:
        return self.data_handler.eval_size() > 0

    def __repr__(self) -> str:
        return str(self)

    def __str__(self) -> str:
        return "DataDispatcher(handler=%s, n=%d, eval_on_user=%s)" % (
            self.data_handler, self.n, self.eval_on_user)


class RecSysDataDispatcher(DataDispatcher):
    """One user per client (ref ``data/__init__.py:513-558``)."""

    def __init__(self, data_handler: "RecSysDataHandler") -> None:  # noqa: F821
        self.data_handler = data_handler
        self.n = data_handler.n_users
        self.eval_on_user = True
        self.assignments: List[int] = list(range(self.n))

    def assign(self, seed: int = 42) -> None:
        """User -> client permutation.  Default: a private generator (no side effects).  Under
        ``reference_compat``: the reference's statement, which re-seeds torch's GLOBAL stream
        (``data/__init__.py:534-536``)."""
        from .. import GlobalSettings
        if GlobalSettings().reference_compat:
            import torch
            torch.manual_seed(seed)
            self.assignments = torch.randperm(self.data_handler.size()).tolist()
            return
        self.assignments = np.random.default_rng(seed).permutation(self.n).tolist()

    def __getitem__(self, idx: int) -> Any:
        assert 0 <= idx < self.n, "Index %d out of range." % idx
        u = self.assignments[idx]
        return self.data_handler.at(u), self.data_handler.at(u, True)

    def get_eval_set(self) -> None:
        return None

    def has_test(self) -> bool:
        return False

    def __str__(self) -> str:
        return "RecSysDataDispatcher(handler=%s, eval_on_user=%s)" % (self.data_handler,
                                                                     self.eval_on_user)


# --------------------------------------------------------------------------------------
# loaders
# --------------------------------------------------------------------------------------
def _standardize(X: np.ndarray) -> np.ndarray:
    mu, sd = X.mean(axis=0), X.std(axis=0)
    sd[sd == 0] = 1.0
    return (X - mu) / sd


def _fallback_allowed(flag: bool) -> bool:
    """Synthetic stand-ins are opt-in: per call (``synthetic_fallback=True``) or for a whole process through the
    environment (``GOSSIPY_SYNTHETIC_FALLBACK=1``, what ``python -m gossipy_b200.compat --synthetic`` sets to run
    unmodified reference scripts on a machine without network)."""
    return bool(flag) or os.environ.get("GOSSIPY_SYNTHETIC_FALLBACK", "") == "1"


def load_classification_dataset(name_or_path: str, normalize: bool = True,
                                as_tensor: bool = True, synthetic_fallback: bool = False):
    """Load a classification data set by name or svmlight path (ref ``data/__init__.py:561-624``).

    sklearn's bundled sets (iris, breast, digits, wine) load offline.  The UCI / reuters sets
    need a download; when that fails the error is raised, like in the reference.  Only with an explicit
    ``synthetic_fallback=True`` is synthetic data of the same shape returned instead (with a warning), so
    that scripts still run on air-gapped GPU boxes -- results on real and synthetic data cannot be mixed up
    silently (the examples opt in).
    """
    from . import synthetic
    X = y = None
    if name_or_path in {"iris", "breast", "digits", "wine"}:
        from sklearn import datasets
        ds = {"iris": datasets.load_iris, "breast": datasets.load_breast_cancer,
              "digits": datasets.load_digits, "wine": datasets.load_wine}[name_or_path]()
        X, y = ds.data, ds.target
    elif name_or_path in UCI_URL_AND_CLASS or name_or_path == "reuters":
        try:
            X, y = _download_named(name_or_path)
        except Exception as exc:
            if not _fallback_allowed(synthetic_fallback):
                raise
            LOG.warning("'%s' cannot be downloaded (%s): using synthetic data of the same shape"
                        % (name_or_path, type(exc).__name__))
            X, y = synthetic.classification_like(name_or_path, as_tensor=False)
    elif name_or_path.startswith("synthetic:"):
        X, y = synthetic.classification_like(name_or_path.split(":", 1)[1], as_tensor=False)
    else:
        from sklearn.datasets import load_svmlight_file
        X, y = load_svmlight_file(name_or_path)
        X = X.toarray()
    X = np.asarray(X, dtype="float64")
    if normalize:
        X = _standardize(X)
    if as_tensor:
        return torch.tensor(X).float(), torch.tensor(np.asarray(y)).long()
    return X, np.asarray(y)


def _download_named(name: str):
    import pandas as pd
    from sklearn.datasets import load_svmlight_file
    from sklearn.preprocessing import LabelEncoder
    if name == "reuters":
        folder = download_and_untar("http://download.joachims.org/svm_light/examples/example1.tar.gz")[0]
        X_tr, y_tr = load_svmlight_file(folder + "/train.dat")
        X_te, y_te = load_svmlight_file(folder + "/test.dat")
        X_te = np.pad(X_te.toarray(), [(0, 0), (0, X_tr.shape[1] - X_te.shape[1])])
        X = np.vstack([X_tr.toarray(), X_te])
        y = LabelEncoder().fit_transform(np.concatenate([y_tr, y_te]))
        shutil.rmtree(folder)
        return X, y
    url, label_id = UCI_URL_AND_CLASS[name]
    data = pd.read_csv(url, header=None).to_numpy()
    y = LabelEncoder().fit_transform(data[:, label_id])
    X = np.delete(data, [label_id], axis=1).astype("float64")
    return X, y


def load_recsys_dataset(name: str, path: str = ".", synthetic_fallback: bool = False):
    """MovieLens ratings as ``{user: [(item, rating), ...]}`` (ref ``data/__init__.py:628-681``)."""
    from . import synthetic
    if name.startswith("synthetic:"):
        return synthetic.ratings_like(name.split(":", 1)[1])
    if name not in {"ml-100k", "ml-1m", "ml-10m", "ml-20m"}:
        raise ValueError("Unknown dataset %s." % name)
    try:
        folder = download_and_unzip("https://files.grouplens.org/datasets/movielens/%s.zip" % name)[0]
    except Exception as exc:
        if not _fallback_allowed(synthetic_fallback):
            raise
        LOG.warning("'%s' cannot be downloaded (%s): using synthetic ratings of the same shape"
                    % (name, type(exc).__name__))
        return synthetic.ratings_like(name)
    filename, sep = {"ml-100k": ("u.data", "\t"), "ml-20m": ("ratings.csv", ",")}.get(
        name, ("ratings.dat", "::"))
    ratings: Dict[int, List[Tuple[int, float]]] = {}
    umap: Dict[int, int] = {}
    imap: Dict[int, int] = {}
    with open(os.path.join(path, folder, filename), "r") as f:
        for line in f:
            parts = line.strip().split(sep)[:3]
            try:
                u, i, r = int(parts[0]), int(parts[1]), float(parts[2])
            except ValueError:
                continue  # header line of ml-20m
            uu = umap.setdefault(u, len(umap))
            ii = imap.setdefault(i, len(imap))
            ratings.setdefault(uu, []).append((ii, r))
    shutil.rmtree(folder)
    return ratings, len(umap), len(imap)


def _torchvision_pair(cls_name: str, path: str):
    import torchvision
    cls = getattr(torchvision.datasets, cls_name)
    try:
        return cls(root=path, train=True, download=False), cls(root=path, train=False, download=False)
    except Exception:
        return cls(root=path, train=True, download=True), cls(root=path, train=False, download=True)


def get_CIFAR10(path: str = "./data", as_tensor: bool = True, synthetic_fallback: bool = False):
    """CIFAR-10 as ``((Xtr, ytr), (Xte, yte))``, images NCHW in [0,1] (ref ``:684-722``)."""
    from . import synthetic
    try:
        tr, te = _torchvision_pair("CIFAR10", path)
    except Exception as exc:
        if not _fallback_allowed(synthetic_fallback):
            raise
        LOG.warning("CIFAR-10 unavailable (%s): synthetic CIFAR-shape data" % type(exc).__name__)
        return synthetic.images_like("cifar10", as_tensor=as_tensor)
    if as_tensor:
        return ((torch.tensor(tr.data).float().permute(0, 3, 1, 2) / 255., torch.tensor(tr.targets)),
                (torch.tensor(te.data).float().permute(0, 3, 1, 2) / 255., torch.tensor(te.targets)))
    return (tr.data, tr.targets), (te.data, te.targets)


def get_FashionMNIST(path: str = "./data", as_tensor: bool = True, synthetic_fallback: bool = False):
    """Fashion-MNIST as ``((Xtr, ytr), (Xte, yte))`` in [0,1] (ref ``:725-762``)."""
    from . import synthetic
    try:
        tr, te = _torchvision_pair("FashionMNIST", path)
    except Exception as exc:
        if not _fallback_allowed(synthetic_fallback):
            raise
        LOG.warning("FashionMNIST unavailable (%s): synthetic MNIST-shape data" % type(exc).__name__)
        return synthetic.images_like("fashionmnist", as_tensor=as_tensor)
    if as_tensor:
        return (tr.data / 255., tr.targets), (te.data / 255., te.targets)
    return (tr.data.numpy() / 255., tr.targets.numpy()), (te.data.numpy() / 255., te.targets.numpy())


def get_FEMNIST(path: str = "./data"):
    """FEMNIST with its natural per-writer split (ref ``:765-778``).

    FIX(B24): the reference never advances the running offsets, so every client receives the
    same index range; here client ``i`` gets its own contiguous range.
    """
    url = "https://raw.githubusercontent.com/tao-shen/FEMNIST_pytorch/master/femnist.tar.gz"
    te_name, tr_name = download_and_untar(url, path)
    Xtr, ytr, ids_tr = torch.load(os.path.join(path, tr_name))
    Xte, yte, ids_te = torch.load(os.path.join(path, te_name))
    tr_assignment, te_assignment, s_tr, s_te = [], [], 0, 0
    for ntr, nte in zip(ids_tr, ids_te):
        tr_assignment.append(list(range(s_tr, s_tr + ntr)))
        te_assignment.append(list(range(s_te, s_te + nte)))
        s_tr, s_te = s_tr + ntr, s_te + nte
    return (Xtr, ytr, tr_assignment), (Xte, yte, te_assignment)
