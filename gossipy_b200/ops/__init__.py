"""Engine ops: one entry point per logical op of SURVEY §2.13, dispatched by device.

* tensors on ``cuda`` -> hand-written sm_100a kernels from the in-tree extension
  ``gossipy_b200._C`` (built by ``__graft_entry__.build()``); a missing extension on a GPU box is
  a hard error -- there is no silent eager fallback;
* tensors on ``cpu`` -> the fp32 PyTorch reference implementations in :mod:`.torch_ref`
  (also the oracle the kernels are tested against).

Cross-rank reads.  Every op that reads a *source* row accepts ``sync``: a :class:`RowSync`
describing the handshake with the rank that owns the row (see ``parallel/runtime.py``).  On CUDA
the handshake is fused into the kernel (spin on the owner's ``ready`` flag before the first peer
load, one ``red.release.sys`` on its ``done`` counter after the last); on CPU (shared-memory
arenas, ``gloo`` plumbing runs) it is a host spin on the same flags.

``launch_count`` counts native kernel launches (reported by ``bench.py`` as ``gpu_launches``).
"""
from __future__ import annotations

import time
from typing import Any, List, Optional, Sequence, Tuple

import torch

from . import metrics, torch_ref
from .native import native, native_available, require_native

launch_count = 0
FORCE_TORCH = False  # tests flip this to run the oracle on GPU tensors


class RowSync:
    """Handshake with the owner of a source row: wait ``ready >= gen``, then ``done += 1``.

    ``ready`` / ``done`` are device addresses (CUDA) or ``(int32 numpy array, index)`` pairs
    (CPU shared memory)."""

    __slots__ = ("ready", "gen", "done")

    def __init__(self, ready: Any, gen: int, done: Any) -> None:
        self.ready, self.gen, self.done = ready, int(gen), done

    def as_tuple(self) -> Tuple[int, int, int]:
        return (int(self.ready), self.gen, int(self.done))

    # CPU side -----------------------------------------------------------------------------
    def host_wait(self, timeout: float = 120.0) -> None:
        arr, i = self.ready
        t0 = time.monotonic()
        while int(arr[i]) - self.gen < 0:
            if time.monotonic() - t0 > timeout:
                raise TimeoutError("peer row was never published (gen %d, flag %d)" % (self.gen, int(arr[i])))
            time.sleep(0)

    def host_done(self) -> None:
        host_flag_add(self.done, 1)


def host_flag_add(flag: Any, value: int) -> None:
    """Increment a shared-memory counter (CPU transport).  Every counter has exactly ONE writing process:
    the ``done`` word of a row has one slot per reader rank (``engine/arena.py``), so a plain
    read-modify-write is race free."""
    arr, i = flag
    arr[i] += value


def _use_native(t: torch.Tensor) -> bool:
    if not t.is_cuda or FORCE_TORCH:
        return False
    require_native()
    return True


def _count(n: int = 1) -> None:
    global launch_count
    launch_count += n


def _st(sync: Optional[RowSync]):
    return None if sync is None else sync.as_tuple()


def merge_pair(dst, src, w_dst: float, w_src: float, lo: int = 0, hi: Optional[int] = None,
               sync: Optional[RowSync] = None) -> None:
    if _use_native(dst):
        hi_ = dst.numel() if hi is None else hi
        native().merge_pair(dst, src, float(w_dst), float(w_src), int(lo), int(hi_), _st(sync))
        _count()
    else:
        if sync is not None:
            sync.host_wait()
        torch_ref.merge_pair(dst, src, w_dst, w_src, lo, hi)
        if sync is not None:
            sync.host_done()


def merge_segments(dst, src, segments, w_dst: float, w_src: float,
                   sync: Optional[RowSync] = None) -> None:
    if _use_native(dst):
        native().merge_segments(dst, src, segments, float(w_dst), float(w_src), _st(sync))
        _count()
    else:
        if sync is not None:
            sync.host_wait()
        torch_ref.merge_segments(dst, src, segments.cpu(), w_dst, w_src)
        if sync is not None:
            sync.host_done()


def merge_indexed(dst, src, index, w_dst: float, w_src: float,
                  sync: Optional[RowSync] = None) -> None:
    if _use_native(dst):
        native().merge_indexed(dst, src, index, float(w_dst), float(w_src), _st(sync))
        _count(2)
    else:
        if sync is not None:
            sync.host_wait()
        torch_ref.merge_indexed(dst, src, index, w_dst, w_src)
        if sync is not None:
            sync.host_done()


def merge_kway(dst, srcs: Sequence[torch.Tensor], weights: Sequence[float],
               syncs: Optional[Sequence[Optional[RowSync]]] = None) -> None:
    if _use_native(dst):
        st = None
        if syncs is not None and any(s is not None for s in syncs):
            st = [(0, 0, 0) if s is None else s.as_tuple() for s in syncs]
        native().merge_kway(dst, list(srcs), [float(w) for w in weights], st)
        _count(max(1, (len(srcs) + 31) // 32))
    else:
        for s in (syncs or ()):
            if s is not None:
                s.host_wait()
        torch_ref.merge_kway(dst, srcs, weights)
        for s in (syncs or ()):
            if s is not None:
                s.host_done()


def snapshot(dst, src, sync: Optional[RowSync] = None) -> None:
    merge_pair(dst, src, 0.0, 1.0, 0, dst.numel(), sync)


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


def keyed_perm(n: int, key: int, device: Any) -> torch.Tensor:
    """The keyed permutation of ``range(n)`` (``torch_ref.perm_indices``) as an int64 tensor on ``device``: written by a
    kernel on a GPU (no host-built vector, no H2D copy -- a pageable copy would block the host until the node's stream
    has drained, i.e. serialise the nodes)."""
    device = torch.device(device)
    if device.type == "cuda" and native_available():
        k = int(key) & ((1 << 64) - 1)
        out = native().keyed_perm(int(n), k - (1 << 64) if k >= (1 << 63) else k, device)     # two's complement for int64
        _count()
        return out
    return torch.from_numpy(torch_ref.perm_indices(n, key)).to(device)


def keyed_randint(k: int, n: int, key: int, device: Any) -> torch.Tensor:
    """``k`` keyed draws from ``range(n)`` with replacement (the coordinate sample of ``SamplingTMH``), on ``device``;
    the same values on every device and in the C++ executor."""
    device = torch.device(device)
    if device.type == "cuda" and native_available():
        kk = int(key) & ((1 << 64) - 1)
        out = native().keyed_randint(int(k), int(n), kk - (1 << 64) if kk >= (1 << 63) else kk, device)
        _count()
        return out
    return torch.from_numpy(torch_ref.keyed_randint(int(k), int(n), int(key))).to(device)


MergeFrom = Tuple[torch.Tensor, float, float, Optional[RowSync]]   # (peer row, w_self, w_peer, sync)
TRAIN_IMPL = ""   # process-wide choice of the fused MLP training kernel, see set_train_impl()
TRAIN_IMPLS = ("", "auto", "tc8", "cluster", "tc8-tf32", "tc3")


def set_train_impl(name: str = "") -> None:
    """Choose the fused MLP training kernel for this process (Python handlers and the C++ executor).

    ``""`` / ``"auto"`` (default): fp32-equivalent -- ``tc8`` (error-compensated 3xTF32 tcgen05 kernel on an 8-CTA
    cluster) with the fp32 CUDA-core ``cluster`` kernel as fall-back.  The plain-tf32 kernels (``tc8-tf32``, ``tc3``) truncate operands to 10 mantissa bits -- below the reference's fp32 -- and run only when
    asked for by name or through ``GlobalSettings().allow_tf32 = True``."""
    global TRAIN_IMPL
    if name not in TRAIN_IMPLS:
        raise ValueError("unknown training kernel %r (one of %s)" % (name, ", ".join(repr(t) for t in TRAIN_IMPLS)))
    TRAIN_IMPL = "" if name == "auto" else name
    from .native import native_available, _try_import
    if native_available():
        _try_import().set_train_impl(TRAIN_IMPL)


def train_dtype() -> str:
    """Human-readable arithmetic of the fused MLP training path in use (bench / reports)."""
    if TRAIN_IMPL in ("tc8-tf32", "tc3"):
        return "tf32 (operands truncated to 10 mantissa bits, fp32 accumulate and master weights)"
    if TRAIN_IMPL == "cluster":
        return "fp32"
    return "fp32-equivalent (3xTF32 error-compensated tcgen05 products, fp32 accumulate; second layer exact fp32)"


def _cpu_premerge(row, merge_from: Optional[MergeFrom]) -> None:
    if merge_from is not None:
        peer, ws, wp, sync = merge_from
        if sync is not None:
            sync.host_wait()
        torch_ref.merge_pair(row, peer, ws, wp, 0, min(row.numel(), peer.numel()))
        if sync is not None:
            sync.host_done()


def mlp1_momentum_supported(dims, batch_size: int, n_samples: int) -> bool:
    """Envelope of the fused momentum-SGD kernel (tcgen05, 8-CTA cluster; momentum buffer of W1 in a TMEM tile)."""
    d_in, d_h, d_out = (int(v) for v in dims)
    bs = n_samples if not batch_size else min(int(batch_size), n_samples)
    return d_in % 4 == 0 and 32 <= d_in <= 896 and d_h <= 128 and d_out <= 10 and bs <= 32


def mlp1_train(row, X, y, dims, batch_size, local_epochs, lr, weight_decay, key,
               elem_scale_ages=None, impl: Optional[str] = None,
               merge_from: Optional[MergeFrom] = None, momentum=None) -> int:
    """Fused local update of a 1-hidden-layer ReLU MLP; returns the number of SGD steps.

    ``merge_from = (peer_row, w_self, w_peer, sync)`` fuses the preceding merge into the same
    launch (MERGE_UPDATE): training starts from ``w_self*row + w_peer*peer_row``.
    ``momentum = (mu, dampening, nesterov, buffer_row, first)``: torch.optim.SGD's momentum inside the kernel."""
    if _use_native(row):
        peer, ws, wp, sync = merge_from if merge_from is not None else (None, 1.0, 0.0, None)
        mu, damp, nest, buf, first = momentum if momentum is not None else (0.0, 0.0, False, None, False)
        n = native().mlp1_train(row, X, y, tuple(int(d) for d in dims), int(batch_size),
                                int(local_epochs), float(lr), float(weight_decay), int(key),
                                None if elem_scale_ages is None else elem_scale_ages[0],
                                None if elem_scale_ages is None else elem_scale_ages[1],
                                ("" if momentum is not None and (impl or TRAIN_IMPL) != "tc8" else impl or TRAIN_IMPL),
                                peer, float(ws), float(wp), _st(sync),
                                float(mu), float(damp), bool(nest), buf, bool(first))
        _count()
        return n
    _cpu_premerge(row, merge_from)
    return torch_ref.mlp1_train(row, X, y, dims, batch_size, local_epochs, lr, weight_decay, key,
                                elem_scale_ages, momentum)


EVAL_IMPL = ""          # "", "tc" or "simt": evaluation kernel choice ("" = tensor cores when the shape fits)


def set_eval_tf32(on: bool) -> None:
    from .native import native_available, _try_import
    if native_available():
        _try_import().set_eval_tf32(bool(on))
_PRETILED: dict = {}     # id(X) -> (X, pre-tiled copy): the test set is tiled once for the tcgen05 kernel


def _pretiled(X: torch.Tensor) -> torch.Tensor:
    hit = _PRETILED.get(id(X))
    if hit is None or hit[0] is not X:
        out = native().mlp1_eval_pretile(X)
        torch.cuda.current_stream(X.device).synchronize()     # other (node) streams will read it
        hit = _PRETILED[id(X)] = (X, out)
        _count()
    return hit[1]


def mlp1_eval(row, X, y, dims, n_classes: int, X_lp=None, want_scores: bool = False):
    """Confusion matrix ``[C,C]`` (int32/int64 tensor on the row's device) of the MLP on (X, y); with
    ``want_scores`` also the class-1 logit of every sample (for the AUC of 2-output networks) as a second value."""
    if _use_native(row):
        d = tuple(int(v) for v in dims)
        if (X_lp is None and EVAL_IMPL != "simt" and d[0] % 4 == 0 and d[1] <= 128 and d[2] <= 10 and n_classes <= 16
                and X.shape[0] >= 512):
            X_lp = _pretiled(X)
        cm, sc = native().mlp1_eval(row, X, y, d, int(n_classes), X_lp, bool(want_scores))
        _count(2 if X_lp is not None else 1)
        return (cm, sc) if want_scores else cm
    logits = torch_ref.mlp1_logits(row, X, dims)
    cm = torch_ref.confusion_matrix(y, logits.argmax(dim=1), n_classes)
    return (cm, logits[:, 1 if logits.shape[1] > 1 else 0]) if want_scores else cm


def logreg_train(row, X, y, dims, batch_size, local_epochs, lr, weight_decay, key,
                 elem_scale_ages=None, merge_from: Optional[MergeFrom] = None) -> int:
    if _use_native(row):
        peer, ws, wp, sync = merge_from if merge_from is not None else (None, 1.0, 0.0, None)
        n = native().logreg_train(row, X, y, tuple(int(d) for d in dims), int(batch_size),
                                  int(local_epochs), float(lr), float(weight_decay), int(key),
                                  None if elem_scale_ages is None else elem_scale_ages[0],
                                  None if elem_scale_ages is None else elem_scale_ages[1],
                                  peer, float(ws), float(wp), _st(sync))
        _count()
        return n
    _cpu_premerge(row, merge_from)
    return torch_ref.logreg_train(row, X, y, dims, batch_size, local_epochs, lr, weight_decay, key,
                                  elem_scale_ages)


def kmeans_match_merge(C_row, P_row, k: int, dim: int, w_own: float = .5, w_peer: float = .5, sync=None) -> torch.Tensor:
    """Optimal-matching merge of two centroid sets in place (k <= 8 on the GPU: exhaustive search in one kernel);
    returns the assignment ``perm`` (peer centroid matched to each own centroid)."""
    if _use_native(C_row) and k <= 8:
        perm = native().kmeans_match_merge(C_row, P_row, int(k), int(dim), float(w_own), float(w_peer), _st(sync))
        _count()
        return perm
    if sync is not None:
        sync.host_wait()
    own, theirs = C_row[:k * dim].view(k, dim), P_row[:k * dim].view(k, dim)
    from scipy.optimize import linear_sum_assignment
    cols = linear_sum_assignment(torch.cdist(own, theirs).cpu().numpy())[1]
    perm = torch.as_tensor(cols, device=C_row.device)
    own.copy_(w_own * own + w_peer * theirs[perm])
    if sync is not None:
        sync.host_done()
    return perm


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
        _count(2)
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
