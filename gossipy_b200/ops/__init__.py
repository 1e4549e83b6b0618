"""Engine ops: one entry point per logical op of SURVEY §2.13, dispatched by device.

* tensors on ``cuda`` -> hand-written sm_100a kernels from the in-tree extension
  ``gossipy_b200._C`` (built by ``__graft_entry__.build()``); a missing extension on a GPU box is
  a hard error -- there is no silent eager fallback;
* tensors on ``cpu`` -> the fp32 PyTorch reference implementations in :mod:`.torch_ref`
  (also the oracle the kernels are tested against).

``launch_count`` counts native kernel launches (reported by ``bench.py`` as ``gpu_launches``).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import metrics, torch_ref
from .native import native, native_available, require_native

launch_count = 0
FORCE_TORCH = False  # tests flip this to run the oracle on GPU tensors


def _use_native(t: torch.Tensor) -> bool:
    if not t.is_cuda or FORCE_TORCH:
        return False
    require_native()
    return True


def _count(n: int = 1) -> None:
    global launch_count
    launch_count += n


def merge_pair(dst, src, w_dst: float, w_src: float, lo: int = 0, hi: Optional[int] = None) -> None:
    if _use_native(dst):
        hi_ = dst.numel() if hi is None else hi
        native().merge_pair(dst, src, float(w_dst), float(w_src), int(lo), int(hi_))
        _count()
    else:
        torch_ref.merge_pair(dst, src, w_dst, w_src, lo, hi)


def merge_segments(dst, src, segments, w_dst: float, w_src: float) -> None:
    if _use_native(dst):
        native().merge_segments(dst, src, segments, float(w_dst), float(w_src))
        _count()
    else:
        torch_ref.merge_segments(dst, src, segments.cpu(), w_dst, w_src)


def merge_indexed(dst, src, index, w_dst: float, w_src: float) -> None:
    if _use_native(dst):
        native().merge_indexed(dst, src, index, float(w_dst), float(w_src))
        _count()
    else:
        torch_ref.merge_indexed(dst, src, index, w_dst, w_src)


def merge_kway(dst, srcs: Sequence[torch.Tensor], weights: Sequence[float]) -> None:
    if _use_native(dst):
        native().merge_kway(dst, list(srcs), [float(w) for w in weights])
        _count()
    else:
        torch_ref.merge_kway(dst, srcs, weights)


def snapshot(dst, src) -> None:
    if _use_native(dst):
        native().merge_pair(dst, src, 0.0, 1.0, 0, dst.numel())
        _count()
    else:
        torch_ref.snapshot(dst, src)


def sgd_step(p, g, n, lr, weight_decay=0.0, momentum=0.0, buf=None, dampening=0.0,
             nesterov=False, first=False, scale=None) -> None:
    if _use_native(p):
        native().sgd_step(p, g, int(n), float(lr), float(weight_decay), float(momentum), buf,
                          float(dampening), bool(nesterov), bool(first), scale)
        _count()
    else:
        torch_ref.sgd_step(p, g, n, lr, weight_decay, momentum, buf, dampening, nesterov, first,
                           scale)


def adam_step(p, g, n, m, v, step, lr, beta1, beta2, eps, weight_decay=0.0,
              decoupled=False) -> None:
    if _use_native(p):
        native().adam_step(p, g, int(n), m, v, int(step), float(lr), float(beta1), float(beta2),
                           float(eps), float(weight_decay), bool(decoupled))
        _count()
    else:
        torch_ref.adam_step(p, g, n, m, v, step, lr, beta1, beta2, eps, weight_decay, decoupled)


def mlp1_train(row, X, y, dims, batch_size, local_epochs, lr, weight_decay, key,
               elem_scale_ages=None, impl: Optional[str] = None) -> int:
    """Fused local update of a 1-hidden-layer ReLU MLP; returns the number of SGD steps."""
    if _use_native(row):
        n = native().mlp1_train(row, X, y, tuple(int(d) for d in dims), int(batch_size),
                                int(local_epochs), float(lr), float(weight_decay), int(key),
                                None if elem_scale_ages is None else elem_scale_ages[0],
                                None if elem_scale_ages is None else elem_scale_ages[1],
                                impl or "")
        _count()
        return n
    return torch_ref.mlp1_train(row, X, y, dims, batch_size, local_epochs, lr, weight_decay, key,
                                elem_scale_ages)


def mlp1_eval(row, X, y, dims, n_classes: int, X_lp=None) -> torch.Tensor:
    """Confusion matrix ``[C,C]`` (int32/int64 tensor on the row's device) of the MLP on (X, y)."""
    if _use_native(row):
        cm = native().mlp1_eval(row, X, y, tuple(int(d) for d in dims), int(n_classes), X_lp)
        _count()
        return cm
    pred = torch_ref.mlp1_logits(row, X, dims).argmax(dim=1)
    return torch_ref.confusion_matrix(y, pred, n_classes)


def logreg_train(row, X, y, dims, batch_size, local_epochs, lr, weight_decay, key,
                 elem_scale_ages=None) -> int:
    if _use_native(row):
        n = native().logreg_train(row, X, y, tuple(int(d) for d in dims), int(batch_size),
                                  int(local_epochs), float(lr), float(weight_decay), int(key),
                                  None if elem_scale_ages is None else elem_scale_ages[0],
                                  None if elem_scale_ages is None else elem_scale_ages[1])
        _count()
        return n
    return torch_ref.logreg_train(row, X, y, dims, batch_size, local_epochs, lr, weight_decay, key,
                                  elem_scale_ages)


def logreg_scores(row, X, dims) -> torch.Tensor:
    if _use_native(row):
        out = native().logreg_scores(row, X, tuple(int(d) for d in dims))
        _count()
        return out
    return torch_ref.logreg_scores(row, X, dims)


def adaline_update(w, X, y, lr) -> None:
    if _use_native(w):
        native().linear_seq_update(w, X, y, 0, float(lr), 0)
        _count()
    else:
        torch_ref.adaline_update(w, X, y, lr)


def pegasos_update(w, X, y, lam, n_updates) -> int:
    if _use_native(w):
        native().linear_seq_update(w, X, y, 1, float(lam), int(n_updates))
        _count()
        return int(n_updates) + int(X.shape[0])
    return torch_ref.pegasos_update(w, X, y, lam, n_updates)


def kmeans_update(C, X, alpha) -> None:
    if _use_native(C):
        native().kmeans_update(C, X, float(alpha))
        _count()
    else:
        torch_ref.kmeans_update(C, X, alpha)


def kmeans_assign(C, X) -> torch.Tensor:
    if _use_native(C):
        out = native().kmeans_assign(C, X)
        _count()
        return out
    return torch_ref.kmeans_assign(C, X)


def mf_update(X, b, Y, c, ratings, reg, lr) -> int:
    if _use_native(Y):
        native().mf_update(X, b, Y, c, ratings, float(reg), float(lr))
        _count()
        return int(ratings.shape[0])
    return torch_ref.mf_update(X, b, Y, c, ratings, reg, lr)
