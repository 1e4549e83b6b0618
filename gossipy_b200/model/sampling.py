"""Parameter sub-sampling and partitioning (Hegedűs et al. 2021).

Behavioural reference: ``gossipy/model/sampling.py:27-234``.  The reference expresses a sample /
partition as per-tensor tuples of index tensors and merges with advanced indexing.  Here the
primary representation is *flat*: positions inside the model's parameter row
(:class:`~gossipy_b200.engine.flat.FlatLayout`), because that is what the fused merge kernels
consume -- a partition becomes a few strided blocks (only those bytes cross NVLink), a sample
becomes one int32 index vector generated on the device.  The reference's dict-of-index-tuples
view is still available (``partitions`` / ``sample``) for API parity.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import LongTensor

from .. import LOG
from . import TorchModel

__all__ = ["TorchModelSampling", "TorchModelPartition"]

SampleDict = Dict[int, Optional[Tuple[LongTensor, ...]]]


def _param_shapes(net: torch.nn.Module) -> List[Tuple[int, ...]]:
    return [tuple(p.shape) for p in net.parameters()]


def _row_major_strides(shape: Sequence[int]) -> List[int]:
    strides, acc = [], 1
    for s in reversed(shape):
        strides.append(acc)
        acc *= s
    return list(reversed(strides))


def sample_dict_to_flat(sample: SampleDict, net: torch.nn.Module) -> torch.Tensor:
    """Convert the reference's per-tensor index tuples into flat row positions."""
    out, off = [], 0
    for ti, shape in enumerate(_param_shapes(net)):
        ids = sample.get(ti)
        if ids is not None:
            strides = _row_major_strides(shape)
            pos = torch.zeros_like(ids[0], dtype=torch.long)
            for d, ix in enumerate(ids):
                pos = pos + ix.long() * strides[d]
            out.append(pos + off)
        off += int(np.prod(shape)) if shape else 1
    return torch.cat(out) if out else torch.zeros(0, dtype=torch.long)


class TorchModelSampling:
    """Random coordinate subsets of a model (ref ``sampling.py:27-107``)."""

    @classmethod
    def sample_size(cls, size: float, n_params: int) -> int:
        assert 0 < size <= 1, "size must be in the range (0, 1]."
        return max(1, int(round(size * n_params)))

    @classmethod
    def sample_flat(cls, size: float, n_params: int, device=None,
                    generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """``round(size*P)`` positions uniform over the P parameters, with replacement.

        Distributionally identical to the reference (tensor chosen proportionally to its numel,
        then an independent uniform index per dimension) but drawn as ONE device-side randint.
        """
        k = cls.sample_size(size, n_params)
        return torch.randint(0, n_params, (k,), device=device, generator=generator,
                             dtype=torch.int64)

    @classmethod
    def sample_reference(cls, size: float, net: TorchModel) -> SampleDict:
        """The reference's draw, call for call, on the NumPy global stream (ref ``sampling.py:56-72``): a tensor per
        sampled coordinate (probability ~ numel), then an independent index per dimension.  Used under
        ``GlobalSettings().reference_compat`` so that a simulation consumes the host RNG exactly like the
        reference (differential tests); the default path draws one device-side ``randint`` instead."""
        from collections import Counter
        shapes = _param_shapes(net)
        probs = np.array([int(np.prod(sh)) if sh else 1 for sh in shapes], dtype="float")
        total = int(probs.sum())
        probs /= sum(probs)
        k = max(1, int(round(size * total)))
        counter = dict(Counter(list(np.random.choice(len(shapes), size=k, p=probs))))
        out: SampleDict = {i: None for i in range(len(shapes))}
        for i, c in counter.items():
            out[int(i)] = tuple(LongTensor(list(np.random.choice(sdim, size=c))) for sdim in shapes[int(i)])
        return out

    @classmethod
    def sample(cls, size: float, net: TorchModel) -> SampleDict:
        """The reference's representation: ``{tensor_idx: tuple(index per dim) | None}``."""
        if size >= 0.9:
            LOG.warning("You are using a high sample size (=%.2f) which can impact the "
                        "performance without much advantage in terms of saved bandwith." % size)
        shapes = _param_shapes(net)
        numels = np.array([int(np.prod(s)) if s else 1 for s in shapes], dtype=np.int64)
        flat = cls.sample_flat(size, int(numels.sum())).numpy()
        bounds = np.concatenate([[0], np.cumsum(numels)])
        owner = np.searchsorted(bounds, flat, side="right") - 1
        out: SampleDict = {i: None for i in range(len(shapes))}
        for ti in np.unique(owner):
            local = flat[owner == ti] - bounds[ti]
            idx = np.unravel_index(local, shapes[ti]) if shapes[ti] else (local,)
            out[int(ti)] = tuple(LongTensor(a) for a in idx)
        return out

    @classmethod
    def merge(cls, sample: SampleDict, net1: TorchModel, net2: TorchModel,
              reduce: str = "mean") -> None:
        """``p1[idx] = (p1[idx] + p2[idx]) / (2 if mean else 1)`` on the sampled coordinates."""
        assert str(net1) == str(net2), "net1 and net2 must have the same architecture."
        assert reduce in {"mean", "sum"}, "reduce must be either 'sum' or 'mean'."
        p1, p2 = list(net1.parameters()), list(net2.parameters())
        assert len(p1) == len(sample), "The provided sample is incompatible with the network."
        w = 0.5 if reduce == "mean" else 1.0
        with torch.no_grad():
            for i, ids in sample.items():
                if ids is not None:
                    p1[i][ids] = (p1[i][ids] + p2[i][ids]) * w


class TorchModelPartition:
    """Deterministic split of a model's parameters into ``n_parts`` near-equal parts.

    The walk order is the reference's (``sampling.py:144-198``): tensors in ``parameters()``
    order, inside a tensor dimension 0 varies fastest (column-major).  Part ``p`` covers walk
    positions ``[bounds[p], bounds[p+1])`` with the first ``P mod n`` parts one element larger.
    Unlike the reference, tensors of any rank are accepted (conv weights included).
    """

    def __init__(self, net_proto: TorchModel, n_parts: int) -> None:
        self.str_arch = str(net_proto)
        self._shapes = _param_shapes(net_proto)
        numels = [int(np.prod(s)) if s else 1 for s in self._shapes]
        self.n_params = int(sum(numels))
        self.n_parts = min(int(n_parts), self.n_params)
        mu, rem = divmod(self.n_params, self.n_parts)
        sizes = [mu + (1 if p < rem else 0) for p in range(self.n_parts)]
        self.bounds = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        # walk position -> flat (row-major, parameters() order) position
        walk_to_flat = np.empty(self.n_params, dtype=np.int64)
        off = 0
        for shape, n in zip(self._shapes, numels):
            if len(shape) <= 1:
                walk_to_flat[off:off + n] = off + np.arange(n)
            else:
                # flat offsets laid out so that iterating dim0-fastest visits them in walk order
                rm = np.arange(n, dtype=np.int64).reshape(shape)
                walk_to_flat[off:off + n] = off + rm.transpose(*reversed(range(len(shape)))).reshape(-1)
            off += n
        self._walk_to_flat = walk_to_flat
        self._tensor_bounds = np.concatenate([[0], np.cumsum(numels)]).astype(np.int64)
        part_id = np.empty(self.n_params, dtype=np.int64)
        for p in range(self.n_parts):
            part_id[walk_to_flat[self.bounds[p]:self.bounds[p + 1]]] = p
        self.part_id = torch.from_numpy(part_id)  #: part of every flat parameter position
        self._partitions: Optional[Dict[int, SampleDict]] = None
        self._segments: Dict[int, torch.Tensor] = {}

    # -- flat views ----------------------------------------------------------------------
    def flat_index(self, id_part: int) -> torch.Tensor:
        p = id_part % self.n_parts
        return torch.from_numpy(np.sort(self._walk_to_flat[self.bounds[p]:self.bounds[p + 1]]))

    def segments(self, id_part: int) -> torch.Tensor:
        """Strided blocks ``[S,4] = (start, n_runs, run_len, run_stride)`` covering part ``p``.

        For a 2-D weight ``[out,in]`` a run of full columns ``c0..c1`` is one block of ``out``
        contiguous runs of ``c1-c0+1`` floats -- coalesced, and the only bytes fetched from the
        peer.  Partial columns / higher-rank tensors degrade to single-column blocks.
        """
        p = id_part % self.n_parts
        seg = self._segments.get(p)
        if seg is not None:
            return seg
        blocks: List[Tuple[int, int, int, int]] = []
        a, b = int(self.bounds[p]), int(self.bounds[p + 1])
        for ti, shape in enumerate(self._shapes):
            t0, t1 = int(self._tensor_bounds[ti]), int(self._tensor_bounds[ti + 1])
            lo, hi = max(a, t0), min(b, t1)
            if lo >= hi:
                continue
            lo, hi = lo - t0, hi - t0  # walk range inside this tensor
            if len(shape) <= 1:
                blocks.append((t0 + lo, 1, hi - lo, 1))
                continue
            s0 = shape[0]
            row_stride = int(np.prod(shape[1:]))
            two_d = len(shape) == 2

            def col_offset(q: int) -> int:
                # q enumerates trailing multi-indices with dim1 fastest
                rem, o = q, 0
                strides = _row_major_strides(shape)
                for d in range(1, len(shape)):
                    o += (rem % shape[d]) * strides[d]
                    rem //= shape[d]
                return o
            c_lo, r_lo = divmod(lo, s0)
            c_hi, r_hi = divmod(hi, s0)  # exclusive end: column c_hi, rows < r_hi
            if c_lo == c_hi:
                blocks.append((t0 + r_lo * row_stride + col_offset(c_lo), r_hi - r_lo, 1, row_stride))
                continue
            if r_lo:
                blocks.append((t0 + r_lo * row_stride + col_offset(c_lo), s0 - r_lo, 1, row_stride))
                c_lo += 1
            if c_hi > c_lo:
                if two_d:
                    blocks.append((t0 + c_lo, s0, c_hi - c_lo, row_stride))
                else:
                    for q in range(c_lo, c_hi):
                        blocks.append((t0 + col_offset(q), s0, 1, row_stride))
            if r_hi:
                blocks.append((t0 + col_offset(c_hi), r_hi, 1, row_stride))
        seg = torch.tensor(blocks, dtype=torch.int64).reshape(-1, 4)
        self._segments[p] = seg
        return seg

    # -- the reference's representation ----------------------------------------------------
    @property
    def partitions(self) -> Dict[int, SampleDict]:
        if self._partitions is None:
            parts: Dict[int, SampleDict] = {}
            for p in range(self.n_parts):
                flat = self._walk_to_flat[self.bounds[p]:self.bounds[p + 1]]
                owner = np.searchsorted(self._tensor_bounds, flat, side="right") - 1
                d: SampleDict = {i: None for i in range(len(self._shapes))}
                for ti in np.unique(owner):
                    local = flat[owner == ti] - self._tensor_bounds[ti]
                    shape = self._shapes[ti]
                    idx = np.unravel_index(local, shape) if shape else (local,)
                    d[int(ti)] = tuple(LongTensor(np.ascontiguousarray(a)) for a in idx)
                parts[p] = d
            self._partitions = parts
        return self._partitions

    def merge(self, id_part: int, net1: TorchModel, net2: TorchModel,
              weights: Optional[Tuple[int, int]] = None) -> None:
        """Weighted average of part ``id_part`` of ``net1`` with ``net2`` (ref ``:201-234``)."""
        assert str(net1) == self.str_arch, "net1 is not compatible."
        assert str(net2) == self.str_arch, "net2 is not compatible."
        w1, w2 = self.mixing_weights(weights)
        part = self.partitions[id_part % self.n_parts]
        p1, p2 = list(net1.parameters()), list(net2.parameters())
        with torch.no_grad():
            for i, ids in part.items():
                if ids is not None:
                    p1[i][ids] = w1 * p1[i][ids] + w2 * p2[i][ids]

    @staticmethod
    def mixing_weights(weights: Optional[Tuple[int, int]]) -> Tuple[float, float]:
        w = weights if (weights is not None and tuple(weights) != (0, 0)) else (1, 1)
        tot = float(w[0] + w[1])
        return float(w[0]) / tot, float(w[1]) / tot
