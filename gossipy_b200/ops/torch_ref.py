"""Plain PyTorch (fp32) implementations of every engine op.

Two roles: (1) the CPU backend of the framework (tests, ``gloo`` plumbing runs), and (2) the
numerics oracle every sm_100a kernel is tested against (``tests/test_kernels_gpu.py``).  Each
function documents the reference line whose arithmetic it reproduces.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from ..engine import rng as _rng

_U64 = np.uint64


# --------------------------------------------------------------------------------------
# keyed permutation (mirrors gb_perm in csrc/common.cuh)
# --------------------------------------------------------------------------------------
def _mix64_np(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x + _U64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> _U64(27))) * _U64(0x94D049BB133111EB)
        return x ^ (x >> _U64(31))


def perm_indices(n: int, key: int) -> np.ndarray:
    """The whole keyed permutation of ``range(n)`` as an int64 array (vectorised)."""
    bits = max(2, (n - 1).bit_length())
    bits += bits & 1
    half = _U64(bits // 2)
    hmask = _U64((1 << (bits // 2)) - 1)
    key = _U64(key & ((1 << 64) - 1))
    x = np.arange(n, dtype=np.uint64)
    pending = np.ones(n, dtype=bool)
    out = np.empty(n, dtype=np.uint64)
    cur = x.copy()
    while pending.any():
        v = cur[pending]
        l, r = v >> half, v & hmask
        for rnd in range(4):
            f = _mix64_np(key ^ _U64(rnd << 56) ^ r) & hmask
            l, r = r, l ^ f
        v = (l << half) | r
        cur[pending] = v
        done_now = v < _U64(n)
        idx = np.flatnonzero(pending)
        out[idx[done_now]] = v[done_now]
        pending[idx[done_now]] = False
    return out.astype(np.int64)


def keyed_randint(k: int, n: int, key: int) -> np.ndarray:
    """``k`` keyed draws from ``range(n)`` with replacement: ``mix64(key ^ i) % n`` (mirrors ``keyed_randint_kernel``)."""
    key = _U64(key & ((1 << 64) - 1))
    return (_mix64_np(key ^ np.arange(k, dtype=np.uint64)) % _U64(n)).astype(np.int64)


# --------------------------------------------------------------------------------------
# merges on flat rows
# --------------------------------------------------------------------------------------
@torch.no_grad()
def merge_pair(dst: torch.Tensor, src: torch.Tensor, w_dst: float, w_src: float,
               lo: int = 0, hi: Optional[int] = None) -> None:
    """``dst[lo:hi] = w_dst*dst + w_src*src`` -- ref ``handler.py:260-280`` (w=.5/.5),
    ``:695-715`` (age weights / adopt), ``:122-125`` (copy, w_dst=0)."""
    hi = dst.numel() if hi is None else hi
    d, s = dst[lo:hi], src[lo:hi]
    if w_dst == 0.0:
        d.copy_(s).mul_(w_src) if w_src != 1.0 else d.copy_(s)
    else:
        d.mul_(w_dst).add_(s, alpha=w_src)


@torch.no_grad()
def merge_segments(dst: torch.Tensor, src: torch.Tensor, segments: torch.Tensor,
                   w_dst: float, w_src: float) -> None:
    """Weighted merge restricted to strided blocks ``segments`` = int64 ``[S,4]`` rows
    ``(start, n_runs, run_len, run_stride)``: element ``(r, c)`` of a block sits at
    ``start + r*run_stride + c``.  A column-major partition of a row-major ``[out,in]`` weight is
    a handful of such blocks (ref ``sampling.py:201-234``)."""
    for start, n_runs, run_len, stride in segments.tolist():
        d = torch.as_strided(dst, (n_runs, run_len), (stride, 1), dst.storage_offset() + start)
        s = torch.as_strided(src, (n_runs, run_len), (stride, 1), src.storage_offset() + start)
        d.mul_(w_dst).add_(s, alpha=w_src)


@torch.no_grad()
def merge_indexed(dst: torch.Tensor, src: torch.Tensor, index: torch.Tensor,
                  w_dst: float, w_src: float) -> None:
    """``dst[index] = w_dst*dst[index] + w_src*src[index]`` (duplicates are benign: every
    duplicate writes the same value) -- ref ``sampling.py:76-107``."""
    vals = dst[index] * w_dst + src[index] * w_src
    dst[index] = vals


@torch.no_grad()
def merge_kway(dst: torch.Tensor, srcs: Sequence[torch.Tensor], weights: Sequence[float]) -> None:
    """``dst = w0*dst + sum_i w_{i+1} src_i`` -- ref ``handler.py:666-688`` and the list form of
    ``:260-280`` (uniform ``1/(k+1)``)."""
    dst.mul_(float(weights[0]))
    for s, w in zip(srcs, weights[1:]):
        dst.add_(s, alpha=float(w))


@torch.no_grad()
def snapshot(dst: torch.Tensor, src: torch.Tensor) -> None:
    dst.copy_(src)


# --------------------------------------------------------------------------------------
# flat optimizers (generic nn.Module path)
# --------------------------------------------------------------------------------------
@torch.no_grad()
def sgd_step(p: torch.Tensor, g: torch.Tensor, n: int, lr: float, weight_decay: float = 0.0,
             momentum: float = 0.0, buf: Optional[torch.Tensor] = None, dampening: float = 0.0,
             nesterov: bool = False, first: bool = False,
             scale: Optional[torch.Tensor] = None) -> None:
    """torch.optim.SGD semantics over the first ``n`` elements of flat ``p`` / ``g``.

    ``scale`` (optional, per element) multiplies the raw gradient first: used for the
    per-partition ``1/age`` scaling of ``PartitionedTMH`` (ref ``handler.py:514-520``)."""
    pp, gg = p[:n], g[:n]
    d = gg * scale[:n] if scale is not None else gg.clone()
    if weight_decay:
        d.add_(pp, alpha=weight_decay)
    if momentum:
        b = buf[:n]
        if first:
            b.copy_(d)
        else:
            b.mul_(momentum).add_(d, alpha=1 - dampening)
        d = d.add(b, alpha=momentum) if nesterov else b
    pp.add_(d, alpha=-lr)


@torch.no_grad()
def adam_step(p: torch.Tensor, g: torch.Tensor, n: int, m: torch.Tensor, v: torch.Tensor,
              step: int, lr: float, beta1: float, beta2: float, eps: float,
              weight_decay: float = 0.0, decoupled: bool = False) -> None:
    """torch.optim.Adam / AdamW semantics over flat vectors (``step`` is 1-based)."""
    pp, gg, mm, vv = p[:n], g[:n], m[:n], v[:n]
    if weight_decay and decoupled:
        pp.mul_(1 - lr * weight_decay)
    d = gg.add(pp, alpha=weight_decay) if (weight_decay and not decoupled) else gg
    mm.mul_(beta1).add_(d, alpha=1 - beta1)
    vv.mul_(beta2).addcmul_(d, d, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    denom = (vv.sqrt() / (bc2 ** 0.5)).add_(eps)
    pp.addcdiv_(mm, denom, value=-lr / bc1)


# --------------------------------------------------------------------------------------
# fused training: 1-hidden-layer ReLU MLP and logistic regression, SGD, cross-entropy
# --------------------------------------------------------------------------------------
def _batches(n: int, batch_size: int, local_epochs: int, key: int):
    """Yield index arrays exactly as ``TorchModelHandler._update`` slices them
    (ref ``handler.py:235-248``) but with the engine's keyed permutation."""
    bs = n if not batch_size else batch_size
    if local_epochs > 0:
        for e in range(local_epochs):
            perm = perm_indices(n, _rng.mix64(key ^ e))
            for i in range(0, n, bs):
                yield perm[i:i + bs]
    else:
        perm = perm_indices(n, _rng.mix64(key))
        yield perm[:bs]


def n_steps(n: int, batch_size: int, local_epochs: int) -> int:
    bs = n if not batch_size else batch_size
    return local_epochs * ((n + bs - 1) // bs) if local_epochs > 0 else 1


def mlp1_unpack(row: torch.Tensor, dims: Tuple[int, int, int]):
    d_in, d_h, d_out = dims
    o = 0
    W1 = row[o:o + d_h * d_in].view(d_h, d_in); o += d_h * d_in
    b1 = row[o:o + d_h]; o += d_h
    W2 = row[o:o + d_out * d_h].view(d_out, d_h); o += d_out * d_h
    b2 = row[o:o + d_out]
    return W1, b1, W2, b2


@torch.no_grad()
def mlp1_train(row: torch.Tensor, X: torch.Tensor, y: torch.Tensor, dims: Tuple[int, int, int],
               batch_size: int, local_epochs: int, lr: float, weight_decay: float, key: int,
               elem_scale_ages: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
               momentum: Optional[Tuple[float, float, bool, torch.Tensor, bool]] = None) -> int:
    """One ``_update`` of a ``Linear-ReLU-Linear`` net with mean cross-entropy and plain SGD
    (ref ``handler.py:235-258``), done with explicit fp32 forward/backward on the flat row.

    ``elem_scale_ages = (part_id[int64, P], ages[int64, n_parts])``: PartitionedTMH semantics --
    before every step all ages are incremented and the gradient of element e is divided by
    ``ages[part_id[e]]`` (ref ``handler.py:503-520``).  ``momentum = (mu, dampening, nesterov, buffer row,
    first)``: torch.optim.SGD's momentum update on the flat buffer (``first``: the buffer holds no state
    yet).  Returns the number of SGD steps.
    """
    W1, b1, W2, b2 = mlp1_unpack(row, dims)
    n = X.shape[0]
    P = W1.numel() + b1.numel() + W2.numel() + b2.numel()
    steps = 0
    ages = None
    if elem_scale_ages is not None:
        part_id, ages = elem_scale_ages
        ages = ages.clone().double()
    for idx in _batches(n, batch_size, local_epochs, key):
        idx_t = torch.from_numpy(idx).to(X.device)
        x, t = X[idx_t], y[idx_t]
        B = x.shape[0]
        z1 = torch.addmm(b1, x, W1.t())
        h = torch.relu(z1)
        z2 = torch.addmm(b2, h, W2.t())
        p = torch.softmax(z2, dim=1)
        p[torch.arange(B, device=X.device), t] -= 1.0
        dz2 = p / B
        gW2 = dz2.t() @ h
        gb2 = dz2.sum(0)
        dz1 = (dz2 @ W2) * (z1 > 0)
        gW1 = dz1.t() @ x
        gb1 = dz1.sum(0)
        g = torch.cat([gW1.reshape(-1), gb1, gW2.reshape(-1), gb2])
        if ages is not None:
            ages += 1
            g = g / ages[part_id].to(g.dtype)
        pr = row[:P]
        if momentum is not None:
            mu, damp, nesterov, buf, first = momentum
            d = g + weight_decay * pr
            b = buf[:P]
            if first and steps == 0:
                b.copy_(d)
            else:
                b.mul_(mu).add_(d, alpha=1 - damp)
            pr.add_(d.add(b, alpha=mu) if nesterov else b, alpha=-lr)
        else:
            pr.add_(g + weight_decay * pr, alpha=-lr)
        steps += 1
    return steps


@torch.no_grad()
def mlp1_logits(row: torch.Tensor, X: torch.Tensor, dims: Tuple[int, int, int]) -> torch.Tensor:
    W1, b1, W2, b2 = mlp1_unpack(row, dims)
    return torch.addmm(b2, torch.relu(torch.addmm(b1, X, W1.t())), W2.t())


def logreg_unpack(row: torch.Tensor, dims: Tuple[int, int]):
    d_in, d_out = dims
    W = row[:d_out * d_in].view(d_out, d_in)
    b = row[d_out * d_in:d_out * d_in + d_out]
    return W, b


@torch.no_grad()
def logreg_train(row: torch.Tensor, X: torch.Tensor, y: torch.Tensor, dims: Tuple[int, int],
                 batch_size: int, local_epochs: int, lr: float, weight_decay: float, key: int,
                 elem_scale_ages: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> int:
    """``sigmoid(Linear)`` fed to mean cross-entropy, SGD (ref ``nn.py:147-174`` + scripts)."""
    W, b = logreg_unpack(row, dims)
    n = X.shape[0]
    P = W.numel() + b.numel()
    steps = 0
    ages = None
    if elem_scale_ages is not None:
        part_id, ages = elem_scale_ages
        ages = ages.clone().double()
    for idx in _batches(n, batch_size, local_epochs, key):
        idx_t = torch.from_numpy(idx).to(X.device)
        x, t = X[idx_t], y[idx_t]
        B = x.shape[0]
        s = torch.sigmoid(torch.addmm(b, x, W.t()))
        p = torch.softmax(s, dim=1)
        p[torch.arange(B, device=X.device), t] -= 1.0
        dz = (p / B) * s * (1 - s)
        g = torch.cat([(dz.t() @ x).reshape(-1), dz.sum(0)])
        if ages is not None:
            ages += 1
            g = g / ages[part_id].to(g.dtype)
        pr = row[:P]
        pr.add_(g + weight_decay * pr, alpha=-lr)
        steps += 1
    return steps


@torch.no_grad()
def logreg_scores(row: torch.Tensor, X: torch.Tensor, dims: Tuple[int, int]) -> torch.Tensor:
    W, b = logreg_unpack(row, dims)
    return torch.sigmoid(torch.addmm(b, X, W.t()))


# --------------------------------------------------------------------------------------
# sequential linear learners
# --------------------------------------------------------------------------------------
@torch.no_grad()
def adaline_update(w: torch.Tensor, X: torch.Tensor, y: torch.Tensor, lr: float) -> None:
    """Per-sample ``w += lr*(y_i - w.x_i)*x_i`` in data order (ref ``handler.py:364-368``)."""
    for i in range(X.shape[0]):
        w.add_(X[i], alpha=float(lr * (y[i] - torch.dot(w, X[i]))))


@torch.no_grad()
def pegasos_update(w: torch.Tensor, X: torch.Tensor, y: torch.Tensor, lam: float,
                   n_updates: int) -> int:
    """Pegasos steps in data order starting at age ``n_updates`` (ref ``handler.py:416-423``):
    ``t=++age; eta=1/(t*lam); yhat=w.x; w*=(1-eta*lam); if yhat*y<1: w+=eta*y*x``."""
    t = n_updates
    for i in range(X.shape[0]):
        t += 1
        eta = 1.0 / (t * lam)
        yhat = float(torch.dot(w, X[i]))
        w.mul_(1.0 - eta * lam)
        if yhat * float(y[i]) - 1 < 0:
            w.add_(X[i], alpha=eta * float(y[i]))
    return t


# --------------------------------------------------------------------------------------
# k-means and matrix factorisation
# --------------------------------------------------------------------------------------
@torch.no_grad()
def kmeans_assign(C: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
    return torch.cdist(X, C, p=2).argmin(dim=1)


@torch.no_grad()
def kmeans_update(C: torch.Tensor, X: torch.Tensor, alpha: float) -> None:
    """``idx=argmin dist; C[idx] = (1-a) C[idx] + a x`` with all assignments computed against the
    *incoming* centroids and, for samples sharing a centroid, the last one winning -- i.e. exactly
    the reference's batched indexing statement (``handler.py:604-615``)."""
    idx = kmeans_assign(C, X)
    # `C[idx] = C[idx] * (1 - alpha) + alpha * X` leaves the winner among duplicate indices to the
    # indexing kernel (thread-count dependent on CPU); "last sample wins" is made explicit here
    n, k = X.shape[0], C.shape[0]
    winner = torch.full((k,), -1, dtype=torch.long, device=C.device)
    winner.scatter_reduce_(0, idx, torch.arange(n, device=C.device), "amax")
    sel = winner >= 0
    C[sel] = C[sel] * (1 - alpha) + alpha * X[winner[sel]]


@torch.no_grad()
def mf_update(X: torch.Tensor, b: torch.Tensor, Y: torch.Tensor, c: torch.Tensor,
              ratings: torch.Tensor, reg: float, lr: float) -> int:
    """Sequential rank-k SGD over a user's ratings (ref ``handler.py:550-560``).  ``X`` [k],
    ``b`` [1], ``Y`` [n_items,k], ``c`` [n_items]; ``ratings`` [m,2] = (item, rating)."""
    decay = 1.0 - reg * lr
    if X.device.type == "cpu" and all(t.dtype == torch.float32 and t.is_contiguous() for t in (X, b, Y, c)):
        # the same recurrence on NumPy views of the tensors' memory (fp32 arithmetic, scalars rounded to fp32 like torch
        # does): ~6x less interpreter overhead than one torch op per term -- this loop is the CPU cost of a recsys run
        import numpy as np
        Xn, bn, Yn, cn = X.numpy(), b.numpy(), Y.numpy(), c.numpy()
        f32 = np.float32
        dec = f32(decay)
        for item, r in ratings.tolist():
            i = int(item)
            yi = Yn[i]
            err = float(f32(r) - np.dot(Xn, yi) - bn[0] - cn[i])
            step = f32(lr * err)
            yi *= dec
            yi += step * Xn
            Xn *= dec
            Xn += step * yi
            bn[0] += step
            cn[i] += step
        return int(ratings.shape[0])
    for item, r in ratings.tolist():
        i = int(item)
        err = float(r - torch.dot(X, Y[i]) - b[0] - c[i])
        Y[i] = decay * Y[i] + lr * err * X
        X.copy_(decay * X + lr * err * Y[i])
        b.add_(lr * err)
        c[i] += lr * err
    return int(ratings.shape[0])


# --------------------------------------------------------------------------------------
# evaluation helpers
# --------------------------------------------------------------------------------------
@torch.no_grad()
def confusion_matrix(y_true: torch.Tensor, y_pred: torch.Tensor, n_classes: int) -> torch.Tensor:
    idx = y_true.long() * n_classes + y_pred.long()
    return torch.bincount(idx, minlength=n_classes * n_classes).view(n_classes, n_classes)
