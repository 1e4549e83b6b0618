"""Model handlers: local learning, merging and evaluation of gossip models.

Behavioural reference: ``gossipy/model/handler.py`` (cited per class).  Architectural
difference: a handler does not own a graph of Python objects that is deep-copied on every send;
it owns ONE flat fp32 *row* in an HBM arena (``engine.arena``).  Consequently

* ``_merge``      = one fused weighted-merge kernel over rows (the peer row is read in place,
                    over NVLink when it lives on another GPU) -- ``ops.merge_*``;
* ``_update``     = one fused local-epoch kernel for the kernel families (``mlp1``, ``logreg``,
                    ``linear``) or autograd + one flat optimizer kernel for arbitrary modules;
* ``caching``     = one D2D snapshot copy into an arena row (skipped when an identical
                    snapshot is already in flight);
* ``evaluate``    = a device-side confusion matrix, only ``C*C`` integers reach the host.

All device work of a handler is enqueued on the CUDA stream of the gossip node that owns it;
snapshots carry events, so exchanges between disjoint node pairs overlap on the GPU.
"""
from __future__ import annotations

import copy
from abc import ABC, abstractmethod
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F

from .. import CACHE, LOG, CacheKey, GlobalSettings, Sizeable
from .. import ops
from ..core import CreateModelMode
from ..engine import arena as _arena
from ..engine import rng as _rng
from ..engine.flat import FlatLayout
from ..parallel import runtime as _prt
from ..ops import metrics as _metrics
from . import TorchModel
from .nn import AdaLine
from .sampling import TorchModelPartition, TorchModelSampling, sample_dict_to_flat

__all__ = ["ModelHandler", "TorchModelHandler", "AdaLineHandler", "PegasosHandler",
           "SamplingTMH", "PartitionedTMH", "MFModelHandler", "KMeansHandler", "WeightedTMH",
           "LimitedMergeTMH", "LimitedMergeMixin", "PendingEval"]


# --------------------------------------------------------------------------------------
# deferred evaluation results
# --------------------------------------------------------------------------------------
class PendingEval:
    """Evaluation whose sufficient statistics are still on the device.

    ``result()`` performs the (tiny) device->host read and returns the metric dict.  The
    simulators enqueue the evaluation of every node first and resolve afterwards, so a round has
    one host synchronisation instead of one per node.
    """

    def __init__(self, finish: Callable[[], Dict[str, float]], event: Any = None) -> None:
        self._finish = finish
        self._value: Optional[Dict[str, float]] = None
        self._event = event   # recorded on the stream that produced the statistics

    def result(self) -> Dict[str, float]:
        if self._value is None:
            if self._event is not None:
                self._event.synchronize()
            self._value = self._finish()
            self._finish = None
        return self._value


def _mark(device: Any):
    """Event on the *current* stream of ``device`` (None on CPU)."""
    device = torch.device(device)
    if device.type != "cuda":
        return None
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(device))
    return ev


def _resolved(d: Dict[str, float]) -> PendingEval:
    p = PendingEval(lambda: d)
    return p


class _NotMineEval(PendingEval):
    """Evaluation computed by another rank (merged later by ``runtime.share_metrics``)."""

    def __init__(self) -> None:
        super().__init__(lambda: None)

    def result(self) -> None:   # type: ignore[override]
        return None


_NOT_MINE = _NotMineEval()
FUSE_MERGE_UPDATE = True                # fold the pairwise merge into the local-update kernel


# --------------------------------------------------------------------------------------
# base class
# --------------------------------------------------------------------------------------
class ModelHandler(Sizeable, ABC):
    """Interface of the learn / merge / evaluate object of a node (ref ``handler.py:58-182``)."""

    def __init__(self, create_model_mode: CreateModelMode = CreateModelMode.MERGE_UPDATE,
                 *args, **kwargs) -> None:
        self.model: Any = None
        self.mode = create_model_mode
        self.n_updates: Any = 0
        self.owner: int = -1      # id of the node that owns this handler (stream selection)
        self._version = 0         # bumped on every change of the model's values

    # -- to implement --------------------------------------------------------------------
    @abstractmethod
    def init(self, *args, **kwargs) -> None: ...

    @abstractmethod
    def _update(self, data: Any, *args, **kwargs) -> None: ...

    @abstractmethod
    def _merge(self, other_model_handler: "ModelHandler", *args, **kwargs) -> None: ...

    @abstractmethod
    def evaluate(self, *args, **kwargs) -> Any: ...

    # -- behaviour shared by all handlers ----------------------------------------------
    def _adopt(self, other: "ModelHandler") -> None:
        """Replace this model's values by ``other``'s (reference: ``deepcopy(recv.model)``)."""
        self.model = copy.deepcopy(other.model)
        self._version += 1

    def _scratch_copy(self, other: "ModelHandler") -> "ModelHandler":
        """A private, trainable copy of a received model (for ``UPDATE_MERGE``)."""
        return other.copy()

    def __call__(self, recv_model: Any, data: Any, *args, **kwargs) -> None:
        """Combine a received model with the local one according to ``self.mode``
        (ref ``handler.py:117-136``)."""
        mode = self.mode
        compat = GlobalSettings().reference_compat
        if mode == CreateModelMode.UPDATE:
            # train the received model on local data and adopt it: adopting first and training
            # the local row is the same computation without mutating the in-flight snapshot
            self._adopt(recv_model)
            self.n_updates = copy.copy(recv_model.n_updates)
            if compat:
                # B13 mimicked: the reference trains the received handler with the optimizer that travelled with it,
                # which is detached if the SENDER had adopted a model before; afterwards this node's own optimizer
                # is detached from the adopted model (it is what this node's later snapshots ship)
                self._opt_detached = bool(recv_model.__dict__.get("_opt_detached", False))
                self._update(data)
                self._opt_detached = True
            else:
                self._update(data)
        elif mode == CreateModelMode.MERGE_UPDATE:
            if not (args or kwargs) and self._merge_update_fused(recv_model, data):
                return
            self._merge(recv_model, *args, **kwargs)
            self._update(data)
        elif mode == CreateModelMode.UPDATE_MERGE:
            self._update(data)
            tmp = self._scratch_copy(recv_model)
            tmp._update(data)
            self._merge(tmp, *args, **kwargs)
            _dispose(tmp)
        elif mode == CreateModelMode.PASS:
            self._adopt(recv_model)
            if compat:
                self._opt_detached = True      # B13: ``self.model = deepcopy(...)`` leaves the optimizer on the old tensors
        else:
            raise ValueError("Unknown create model mode %s" % str(mode))

    def _merge_update_fused(self, recv_model: Any, data: Any) -> bool:
        """Hook: do merge + local update as ONE device launch; False = not available."""
        return False

    def evaluate_async(self, *args, **kwargs) -> PendingEval:
        return _resolved(self.evaluate(*args, **kwargs))

    def copy(self) -> Any:
        return copy.deepcopy(self)

    def get_size(self) -> int:
        return self.model.get_size() if self.model is not None else 0

    def _age_key(self) -> Any:
        return self.n_updates

    def caching(self, owner: int) -> CacheKey:
        """Put a snapshot of this model "on the wire" and return its key.

        The key is ``(owner, age, version)``; when the same version is already in flight only a
        reference is added (no copy) -- e.g. an all-to-all node pushing to 20 neighbours takes
        one snapshot, not 20.
        """
        key = CacheKey(owner, self._age_key(), self._version)
        if key in CACHE:
            CACHE.push(key, None)
        else:
            CACHE.push(key, self._snapshot())
        return key

    def _snapshot(self) -> "ModelHandler":
        return self.copy()

    def release(self) -> None:
        """Return device resources (called when a snapshot leaves the cache)."""

    def __eq__(self, other: Any) -> bool:
        return isinstance(other, self.__class__) and _state_eq(self.__dict__, other.__dict__)

    def __ne__(self, other: Any) -> bool:
        return not self.__eq__(other)

    __hash__ = object.__hash__

    def __repr__(self) -> str:
        return str(self)

    def __str__(self) -> str:
        return "%s(model=%s_%s, mode=%s)" % (self.__class__.__name__, str(self.model),
                                             self.n_updates, self.mode)


def _dispose(h: Any) -> None:
    rel = getattr(h, "release", None)
    if callable(rel):
        rel()


def _state_eq(a: Dict[str, Any], b: Dict[str, Any]) -> bool:
    skip = {"_row", "_module", "_proto", "_grad_row", "_opt_rows", "_torch_opt", "owner", "_size_cache", "_graphs",
            "_version", "_is_snapshot", "_bound_to", "_grad_bound", "_update_counter",
            "_part_id_dev", "_seg_dev", "layout"}
    for k in a.keys() | b.keys():
        if k in skip:
            continue
        va, vb = a.get(k), b.get(k)
        if isinstance(va, np.ndarray) or isinstance(vb, np.ndarray):
            if not np.array_equal(va, vb):
                return False
        elif isinstance(va, torch.Tensor) or isinstance(vb, torch.Tensor):
            if not (isinstance(va, torch.Tensor) and isinstance(vb, torch.Tensor)
                    and torch.equal(va.cpu(), vb.cpu())):
                return False
        elif isinstance(va, torch.nn.Module) and isinstance(vb, torch.nn.Module):
            if str(va) != str(vb):
                return False
        elif callable(va) and callable(vb):
            continue
        elif va != vb:
            return False
    return True


# --------------------------------------------------------------------------------------
# device-resident data: every distinct host tensor is uploaded once and shared by all handlers
# --------------------------------------------------------------------------------------
_DEVICE_CACHE: Dict[Tuple[int, str], Tuple[Any, Any]] = {}


def _move(t: Any, dev: torch.device) -> Any:
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t))
    if not isinstance(t, torch.Tensor):
        return t
    if t.dtype == torch.float64:
        t = t.float()
    if t.device == dev:
        return t
    key = (id(t), str(dev))
    hit = _DEVICE_CACHE.get(key)
    if hit is not None and hit[0] is t:
        return hit[1]
    moved = t.to(dev).contiguous()
    if dev.type == "cuda":
        # the upload ran on whatever stream was current; later users sit on other (node) streams
        torch.cuda.current_stream(dev).synchronize()
    _DEVICE_CACHE[key] = (t, moved)  # holding `t` keeps id() unique
    return moved


def to_device_cached(data: Any, dev: torch.device) -> Any:
    if data is None:
        return None
    if isinstance(data, (tuple, list)):
        return tuple(_move(t, dev) for t in data)
    return _move(data, dev)


def clear_device_cache() -> None:
    _DEVICE_CACHE.clear()
    _PINNED.clear()
    _STREAMED.clear()
    _COPY_STREAMS.clear()


_PINNED: Dict[int, Tuple[Any, torch.Tensor]] = {}
_STREAMED: Dict[Tuple[int, str], Dict[str, Any]] = {}     # double buffers of streamed inputs
_COPY_STREAMS: Dict[str, Any] = {}


def refresh_device_copy(data: Any, dev: torch.device) -> int:
    """Fresh inputs for the coming round, host (pinned) -> device (``stream_inputs`` mode of the
    simulators).  Double buffered: the copy for round r+1 runs on a dedicated copy stream while
    round r computes; calling this at the start of a round (on the node's stream) makes the buffer
    filled during the previous round current -- the node's stream only waits for that copy's event --
    and starts the next prefetch into the buffer the previous round has finished reading.
    Returns the number of bytes copied per call."""
    if data is None or dev.type != "cuda":
        return 0
    total = 0
    node_stream = torch.cuda.current_stream(dev)
    cs = _COPY_STREAMS.get(str(dev))
    if cs is None:
        cs = _COPY_STREAMS[str(dev)] = torch.cuda.Stream(device=dev)
    for t in (data if isinstance(data, (tuple, list)) else (data,)):
        if not isinstance(t, torch.Tensor) or t.device == dev:
            continue
        key = (id(t), str(dev))
        hit = _DEVICE_CACHE.get(key)
        if hit is None or hit[0] is not t:
            _move(t, dev)
            hit = _DEVICE_CACHE[key]
        pin = _PINNED.get(id(t))
        if pin is None or pin[0] is not t:
            src = t.float() if t.dtype == torch.float64 else t
            pin = _PINNED[id(t)] = (t, src.contiguous().pin_memory())
        st = _STREAMED.get(key)
        if st is None:      # first round: upload in line, allocate the second buffer
            st = _STREAMED[key] = {"bufs": [hit[1], torch.empty_like(hit[1])], "cur": 0, "event": None}
            hit[1].copy_(pin[1], non_blocking=True)
        else:               # the prefetched buffer becomes the resident copy
            st["cur"] ^= 1
            node_stream.wait_event(st["event"])
            _DEVICE_CACHE[key] = (t, st["bufs"][st["cur"]])
        other = st["bufs"][st["cur"] ^ 1]
        free = torch.cuda.Event()
        free.record(node_stream)            # the previous round's readers of `other` sit before this point
        cs.wait_event(free)
        with torch.cuda.stream(cs):
            other.copy_(pin[1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        st["event"] = ev
        total += pin[1].numel() * pin[1].element_size()
    return total


def join_copy_streams(dev: torch.device) -> None:
    """Make the current stream wait for every in-flight input prefetch (used before stopping a timer)."""
    dev = torch.device(dev)
    cs = _COPY_STREAMS.get(str(dev))
    if cs is not None and dev.type == "cuda":
        ev = torch.cuda.Event()
        ev.record(cs)
        torch.cuda.current_stream(dev).wait_event(ev)


# --------------------------------------------------------------------------------------
# row-backed handlers
# --------------------------------------------------------------------------------------
class RowHandler(ModelHandler):
    """A handler whose model values live in one arena row (see module docstring)."""

    def __init__(self, create_model_mode: CreateModelMode) -> None:
        super().__init__(create_model_mode)
        self._row: Optional[_arena.Row] = None
        self._row_numel = 0
        self._is_snapshot = False

    # -- device / stream plumbing -------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return GlobalSettings().get_device()

    def _stream(self):
        return _arena.stream_for(self.device, self.owner)

    # -- multi-rank execution: every rank replays the bookkeeping, the owner's rank does the work --
    def _rank(self) -> int:
        return _prt.rank_of(self.owner) if _prt.active() else 0

    def _mine(self) -> bool:
        return (not _prt.active()) or _prt.rank_of(self.owner) == _prt.rank()

    def _ensure_row(self) -> _arena.Row:
        dev = self.device
        if self._row is None:
            self._row = _arena.arena_for(dev, self._row_numel, self._rank()).alloc()
            self._claim_row()
            self._on_new_row(None)
        elif self._row.tensor is not None and self._row.tensor.device.type != dev.type:      # (ghost rows of other ranks carry no tensor)
            old = self._row
            self._row = _arena.arena_for(dev, self._row_numel, self._rank()).alloc()
            self._claim_row()
            if self._mine():
                self._row.tensor.copy_(old.tensor)
            old.release()
            self._on_new_row(old)
        return self._row

    def _claim_row(self) -> None:
        """A recycled arena row becomes this handler's live row: order my stream after every
        earlier user of the row (local streams through events, other ranks through ``done``)."""
        row = self._row
        if not self._mine() or self.device.type != "cuda" and not _prt.active():
            return
        s = self._stream()
        with _arena.on_stream(s):
            _arena.before_write(row, s if s is not None else _arena.current(self.device))
            _arena.wait_remote_readers(row)

    def _on_new_row(self, old: Optional[_arena.Row]) -> None:
        """Hook: (re)bind views after the row moved."""

    @property
    def row(self) -> torch.Tensor:
        return self._ensure_row().tensor

    def _to_device(self, data: Any) -> Any:
        """Device-resident copy of a node's shard / the eval set, uploaded once."""
        return to_device_cached(data, self.device)

    def refresh_inputs(self, data: Any) -> int:
        """Re-upload this node's inputs from pinned host memory on the node's stream."""
        if not self._mine():
            return 0
        with _arena.on_stream(self._stream()):
            return refresh_device_copy(data, self.device)

    # -- snapshot / copy ------------------------------------------------------------------
    def _clone_shell(self) -> "RowHandler":
        new = object.__new__(self.__class__)
        new.__dict__.update(self.__dict__)
        # the refcount stamped by CacheItem belongs to the in-flight snapshot, never to its clones
        # (a clone that inherited it would refuse to release its row: unbounded arena growth)
        new.__dict__.pop("_cache_refs", None)
        new.__dict__.pop("_graphs", None)       # captured steps are bound to MY row's addresses
        new._row = None
        new.n_updates = copy.copy(self.n_updates)
        return new

    def _snapshot(self) -> "RowHandler":
        """Light clone: new arena row filled by a D2D copy on the owner's stream (and, with several
        ranks, published to the other GPUs through the row's ``ready`` flag)."""
        src = self._ensure_row()
        new = self._clone_shell()
        new._is_snapshot = True
        new._after_clone(self, light=True)
        dst = _arena.arena_for(self.device, self._row_numel, self._rank()).alloc()
        mine = self._mine()
        if mine:
            s = self._stream()
            with _arena.on_stream(s):
                cur = s if s is not None else _arena.current(self.device)
                _arena.before_write(dst, cur)
                _arena.wait_remote_readers(dst)
                ops.snapshot(dst.tensor, src.tensor)
                _arena.after_write(dst, cur, shared=True)
                _arena.publish(dst, True)
        else:
            _arena.publish(dst, False)
        new._row = dst
        return new

    def copy(self) -> "RowHandler":
        new = self._clone_shell()
        new._is_snapshot = False
        new._after_clone(self, light=False)
        if self._row is not None:
            dst = new._ensure_row()
            if self._mine():
                s = self._stream()
                with _arena.on_stream(s):
                    ops.snapshot(dst.tensor, self._row.tensor)
        return new

    def __deepcopy__(self, memo) -> "RowHandler":
        return self.copy()

    def _after_clone(self, src: "RowHandler", light: bool) -> None:
        """Hook: fix up per-instance members of a fresh clone."""

    def release(self) -> None:
        if self.__dict__.get("_cache_refs", 0) > 0:
            return  # still referenced by other in-flight messages
        if self._row is not None:
            self._row.release()
            self._row = None

    # -- generic row operations ----------------------------------------------------------
    def _pull(self, others: Any, fn: Callable) -> None:
        """Read the rows of ``others`` (one handler or a list) into an op on MY row.

        ``fn(srcs, syncs)`` receives the source tensors (peer-mapped when a row lives on another
        GPU) and the cross-rank handshakes (``None`` for local rows) and runs on this handler's
        stream.  The read bookkeeping is replayed on every rank; only the owner of ``self``
        executes ``fn``."""
        single = not isinstance(others, (list, tuple))
        hs = [others] if single else list(others)
        rows = [o._ensure_row() for o in hs]
        my_rank = self._rank()
        if _prt.active() and _prt.transport() == "nccl":
            self._pull_nccl(rows, fn, single)
            return
        syncs = [_arena.read_sync(r, my_rank) for r in rows]
        if not self._mine():
            return
        s = self._stream()
        with _arena.on_stream(s):
            cur = s if s is not None else _arena.current(self.device)
            local = [r for r in rows if (not _prt.active()) or r.rank == my_rank]
            for r in local:
                _arena.before_read(r, cur)
            srcs = [r.tensor for r in rows]
            fn(srcs[0] if single else srcs, syncs[0] if single else syncs)
            for r in local:
                _arena.after_read(r, cur)

    def _pull_nccl(self, rows: List[Any], fn: Callable, single: bool) -> None:
        """``transport="nccl"`` (the NCCL-only baseline of the same engine): a row that lives on
        another rank is SENT by its owner (``ncclSend``, ordered after the row's last writer) and RECEIVED by
        the reader into a staging tensor, on which the same merge / training kernel then runs -- no peer
        memory, no flags.  Every rank replays the same event sequence, so sends and receives pair up."""
        import torch.distributed as dist
        me, dst = _prt.rank(), self._rank()
        mine = self._mine()
        cuda = self.device.type == "cuda"
        cur = _arena.current(self.device) if cuda else None
        srcs: List[Any] = []
        staged = False
        for r in rows:
            if r.rank == dst:
                srcs.append(r.tensor)
                continue
            if me == r.rank:                    # my row: ship it (other ranks that own neither end do nothing)
                _arena.before_read(r, cur)
                dist.send(r.tensor, dst=dst)
                _arena.after_read(r, cur)
            if mine:
                stage = torch.empty(r.tensor.shape if r.tensor is not None else (self._row_numel,), dtype=torch.float32,
                                    device=self.device)
                dist.recv(stage, src=r.rank)
                srcs.append(stage)
                staged = True
        if not mine:
            return
        s = self._stream()
        if cuda and staged and s is not None:   # the node's stream consumes what arrived on the communication stream
            ev = torch.cuda.Event()
            ev.record(cur)
            s.wait_event(ev)
        with _arena.on_stream(s):
            c = s if s is not None else cur
            local = [r for r in rows if r.rank == dst]
            for r in local:
                _arena.before_read(r, c)
            none = [None] * len(rows)
            fn(srcs[0] if single else srcs, None if single else none)
            for r in local:
                _arena.after_read(r, c)
            if cuda and staged and s is not None:
                for t in srcs:
                    if t is not None:
                        t.record_stream(s)

    def _adopt(self, other: "RowHandler") -> None:
        self._pull(other, lambda src, sync: ops.merge_pair(self.row, src, 0.0, 1.0, sync=sync))
        self._version += 1

    def _scratch_copy(self, other: "RowHandler") -> "RowHandler":
        tmp = other._clone_shell()
        tmp._is_snapshot = True
        tmp._after_clone(other, light=True)
        tmp.owner = self.owner
        tmp._ensure_row()
        self._pull(other, lambda src, sync: ops.snapshot(tmp._row.tensor, src, sync))
        return tmp

    def _weighted_merge(self, other: "RowHandler", w_self: float, w_other: float,
                        lo: int = 0, hi: Optional[int] = None) -> None:
        self._pull(other, lambda src, sync: ops.merge_pair(self.row, src, w_self, w_other, lo, hi, sync))
        self._version += 1

    # -- (de)serialisation: rows travel as CPU tensors ------------------------------------
    def __getstate__(self) -> Dict[str, Any]:
        st = dict(self.__dict__)
        st["_row"] = None if self._row is None else self._row.tensor.detach().cpu().clone()
        st.pop("_module", None)
        st.pop("_torch_opt", None)
        st.pop("_graphs", None)
        st["_grad_row"] = None
        opt = st.get("_opt_rows")
        if opt:
            st["_opt_rows"] = {k: v.detach().cpu().clone() for k, v in opt.items()}
        return st

    def __setstate__(self, st: Dict[str, Any]) -> None:
        saved = st.pop("_row", None)
        self.__dict__.update(st)
        self._row = None
        self._restore_members()
        if saved is not None:
            row = self._ensure_row()            # (several ranks: every rank replays the allocation ...)
            if self._mine():                    # ... but only the owner restores the values: a peer's copy may be older
                row.tensor.copy_(saved)
        opt = self.__dict__.get("_opt_rows")
        if opt:
            self._opt_rows = {k: v.to(self.device) for k, v in opt.items()}

    def _restore_members(self) -> None:
        """Hook: rebuild members dropped by ``__getstate__``."""


# --------------------------------------------------------------------------------------
# TorchModelHandler
# --------------------------------------------------------------------------------------
def _is_plain_ce(criterion: Any) -> bool:
    if criterion is F.cross_entropy:
        return True
    if isinstance(criterion, torch.nn.CrossEntropyLoss):
        return (criterion.weight is None and criterion.reduction == "mean"
                and getattr(criterion, "label_smoothing", 0.0) == 0.0
                and criterion.ignore_index == -100)
    return False


class _GraphStep:
    """One captured forward + backward (``TorchModelHandler._graph_fwd_bwd``): static inputs + the graph."""

    __slots__ = ("graph", "x", "y", "seen", "failed", "error")

    def __init__(self) -> None:
        self.graph = None
        self.x = self.y = None
        self.seen = 0
        self.failed = False
        self.error = ""

    def fill(self, x: torch.Tensor, y: torch.Tensor, idx: Optional[torch.Tensor]) -> None:
        if idx is not None:
            torch.index_select(x, 0, idx, out=self.x)
            torch.index_select(y, 0, idx, out=self.y)
        else:
            self.x.copy_(x)
            self.y.copy_(y)


_GRAPH_POOLS: Dict[Tuple[int, int], Any] = {}


def _graph_pool(device: torch.device):
    """The graph memory pool of the current stream of ``device`` (shared by the handlers that run on that stream).

    The pool is owned by a ``torch.cuda.MemPool`` object that lives as long as the process: a bare pool handle dies with
    the last graph captured into it, and capturing into a dead handle trips an internal assertion of the caching allocator
    (met when handlers -- and their graphs -- of an earlier simulation had been collected).  ``None`` (a private pool per
    graph) when the API is missing."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ent = _GRAPH_POOLS.get(key)
    if ent is None:
        mem_pool = getattr(torch.cuda, "MemPool", None)
        ent = (None, None)
        if mem_pool is not None:
            try:
                with torch.cuda.device(device):
                    owner = mem_pool()
                ent = (owner.id, owner)
            except Exception:              # pragma: no cover - depends on the torch build
                ent = (None, None)
        _GRAPH_POOLS[key] = ent
    return ent[0]


class TorchModelHandler(RowHandler):
    """Mini-batch gradient learner for any :class:`TorchModel` (ref ``handler.py:185-334``).

    Parameters as in the reference: ``net, optimizer (class), optimizer_params, criterion,
    local_epochs=1, batch_size=32, create_model_mode=MERGE_UPDATE, copy_model=True``.

    Execution paths chosen once at construction:

    ``fused``    the net advertises a kernel family (``TorchMLP`` with one ReLU hidden layer,
                 ``LogisticRegression``), the optimizer is momentum-free SGD and the criterion is
                 mean cross-entropy -> a whole local epoch (shuffle, forward, backward, SGD) is
                 ONE persistent kernel launch, weights never leave the SM.
    ``generic``  everything else -> the module's parameters are views of the row, autograd runs
                 the step, and SGD(+momentum/nesterov) / Adam / AdamW are a single flat kernel
                 over the row (other optimizer classes fall back to the torch optimizer).
    """

    def __init__(self, net: TorchModel, optimizer: Any, optimizer_params: Dict[str, Any],
                 criterion: Callable, local_epochs: int = 1, batch_size: int = 32,
                 create_model_mode: CreateModelMode = CreateModelMode.MERGE_UPDATE,
                 copy_model: bool = True) -> None:
        super().__init__(create_model_mode)
        assert (batch_size == 0 and local_epochs > 0) or (batch_size > 0)
        # `_proto` is an unbound template (architecture + initial values) shared by all clones;
        # `_module` is this handler's live module whose tensors are views of the row (lazy).
        self._proto = copy.deepcopy(net)
        self._module = None if copy_model else net
        cl = GlobalSettings().channels_last
        # "auto": on a GPU, one rank.  With several ranks cuDNN's NHWC BatchNorm kernels are off limits: they are persistent
        # (cooperative: all CTAs must be co-resident), so one of them queues behind the kernels that spin on a peer's
        # flag while the snapshot the peer is waiting for queues behind it -- on both GPUs at once (measured: config 5 on
        # 2 GPUs timed out in the bounded waits with channels-last rows, not with NCHW rows)
        self.layout = FlatLayout(self._proto, channels_last=self._ROW_CHANNELS_LAST_OK and (
            cl is True or (cl == "auto" and GlobalSettings().is_cuda() and not _prt.active())))
        self._row_numel = self.layout.padded
        self._size_cache = int(self._proto.get_size())
        self.optimizer_cls = optimizer
        self.optimizer_params = dict(optimizer_params)
        self.criterion = criterion
        self.local_epochs = local_epochs
        self.batch_size = batch_size
        self._family = self._proto.fused_family() if hasattr(self._proto, "fused_family") else None
        self._opt_kind = self._classify_optimizer()
        self._fused = (self._family is not None and self._family[0] in ("mlp1", "logreg")
                       and self._opt_kind == "sgd_plain" and _is_plain_ce(criterion))
        # torch.optim.SGD with momentum: fused as well for the MLP family inside the tensor-core kernel's envelope
        # (momentum buffer of W1 in a TMEM tile); checked per update against the shard's shape
        self._fused_momentum = (self._family is not None and self._family[0] == "mlp1"
                                and self._opt_kind == "sgd_momentum" and _is_plain_ce(criterion)
                                and type(self)._elem_scale is TorchModelHandler._elem_scale)
        self._grad_row: Optional[torch.Tensor] = None
        self._opt_rows: Dict[str, torch.Tensor] = {}
        self._opt_steps = 0
        self._torch_opt = None
        self._update_counter = 0
        self._bound_to = None
        self._int_state: Optional[Dict[str, torch.Tensor]] = None

    # -- construction helpers -------------------------------------------------------------
    def _classify_optimizer(self) -> str:
        p = self.optimizer_params
        if self.optimizer_cls is torch.optim.SGD:
            extra = set(p) - {"lr", "weight_decay", "momentum", "dampening", "nesterov"}
            if extra:
                return "torch"
            if not p.get("momentum", 0) and not p.get("nesterov", False):
                return "sgd_plain"
            return "sgd_momentum"
        if self.optimizer_cls in (torch.optim.Adam, torch.optim.AdamW):
            extra = set(p) - {"lr", "betas", "eps", "weight_decay"}
            return "torch" if extra else ("adamw" if self.optimizer_cls is torch.optim.AdamW
                                          else "adam")
        return "torch"

    @property
    def model(self) -> TorchModel:
        """The live ``nn.Module`` whose parameters are views of this handler's row."""
        if self._module is None:
            self._module = copy.deepcopy(self._proto)
            self._bound_to = None
            if self._int_state:
                bufs = dict(self._module.named_buffers())
                for n, v in self._int_state.items():
                    bufs[n].copy_(v)
        self._bind()
        return self._module

    @model.setter
    def model(self, value: Any) -> None:
        if value is None:
            return
        row = self._ensure_row()
        if value.__dict__.get("_parameters") is not None:
            dev = row.tensor.device
            if any(p.device != dev for p in value.parameters()):
                value = copy.deepcopy(value).to(dev)
        self.layout.gather(value, row.tensor)
        self._module = value
        self._bound_to = None
        self._version += 1

    def _bind(self) -> None:
        row = self._ensure_row()
        if self._bound_to is not row.tensor and self._module is not None:
            if self.layout.entries:
                dev = row.tensor.device
                if next(self._module.parameters()).device != dev:
                    self._module.to(dev)
            self.layout.bind(self._module, row.tensor, None)
            self._bound_to = row.tensor
            self._grad_bound = False

    def _on_new_row(self, old) -> None:
        if old is None and not self._is_snapshot and self._mine():
            # first materialisation of a live handler: start from the template's values
            src = self._module if self._module is not None else self._proto
            self.layout.gather(src, self._row.tensor)
        self._bound_to = None
        self._grad_row = None

    def _after_clone(self, src: "TorchModelHandler", light: bool) -> None:
        self._grad_row = None
        self._torch_opt = None
        self._bound_to = None
        self._module = None  # rebuilt lazily from the shared template when someone asks for it
        self._int_state = src._int_values() if self.layout.int_buffers else None
        need_state = (not light) or self.mode in (CreateModelMode.UPDATE,
                                                  CreateModelMode.UPDATE_MERGE)
        self._opt_rows = ({k: v.clone() for k, v in src._opt_rows.items()} if need_state else {})
        if not light and src._row is None and src._module is not None:
            # copying a never-materialised handler built with copy_model=False
            self._proto = copy.deepcopy(src._module)

    def _restore_members(self) -> None:
        self._torch_opt = None
        self._bound_to = None
        self._module = None

    def __getstate__(self) -> Dict[str, Any]:
        st = super().__getstate__()
        st["_module"] = None
        st["_bound_to"] = None
        return st

    # -- API --------------------------------------------------------------------------------
    def init(self) -> None:
        """Initialise the weights.  The random stream is keyed by (base seed, owner), so a node's
        initial model does not depend on how nodes are placed on GPUs."""
        self._version += 1
        if not self._mine():
            self._ensure_row()
            return
        mod = self.model
        dev = self.device
        with _arena.on_stream(self._stream()):
            if GlobalSettings().reference_compat:   # global torch stream, like the reference (placement dependent)
                mod.init_weights()
                return
            with torch.random.fork_rng(devices=[dev] if dev.type == "cuda" else []):
                torch.manual_seed(_rng.derive(0x1217, self.owner if self.owner >= 0 else 0))
                mod.init_weights()

    def get_size(self) -> int:
        size = self.__dict__.get("_size_cache")
        if size is None:            # every message asks for it; the architecture never changes
            size = self.__dict__["_size_cache"] = int(self._proto.get_size())
        return size

    # -- local learning -------------------------------------------------------------------
    def _next_key(self) -> int:
        self._update_counter += 1
        return _rng.derive(0x5EED, self.owner if self.owner >= 0 else 0, self._update_counter,
                           int(np.sum(self.n_updates)))

    def _elem_scale(self) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
        return None

    def _count_steps(self, steps: int) -> None:
        self.n_updates += steps

    def _n_local_steps(self, data: Any) -> int:
        return ops.torch_ref.n_steps(int(data[0].shape[0]), self.batch_size, self.local_epochs)

    def _update(self, data: Tuple[torch.Tensor, torch.Tensor], merge_from: Any = None) -> None:
        self._version += 1
        if self.__dict__.get("_opt_detached") and GlobalSettings().reference_compat and merge_from is None:
            # B13 mimicked (differential tests only): the reference's optimizer still points at parameters that were
            # replaced by an adoption, so its step changes nothing; the age still advances (handler.py:250-258)
            self._next_key()
            steps = self._n_local_steps(data)
            if self._fused:
                self._count_steps(steps)
            else:
                for _ in range(steps):
                    self._pre_step()
                self._count_steps(steps)
            return
        if not self._mine():       # another rank trains this node: replay the bookkeeping only
            self._ensure_row()
            self._next_key()
            if self._fused:
                self._count_steps(self._n_local_steps(data))
            else:
                for _ in range(self._n_local_steps(data)):
                    self._pre_step()
                self._count_steps(self._n_local_steps(data))
            return
        x, y = self._to_device(data)
        s = self._stream()
        with _arena.on_stream(s):
            if self._fused and not (GlobalSettings().reference_compat and merge_from is None and self.batch_size):
                steps = self._update_fused(x, y, merge_from)
            elif (self.__dict__.get("_fused_momentum") and not GlobalSettings().reference_compat and not self.layout.int_buffers
                  and ops.mlp1_momentum_supported(self._family[1], self.batch_size, int(x.shape[0]))):
                steps = self._update_fused(x, y, merge_from, momentum=True)
            else:       # (compat + mini-batches: the reference's own shuffles, which only the autograd path can follow)
                steps = self._update_generic(x, y)
        self._count_steps(steps)

    def _update_fused(self, x: torch.Tensor, y: torch.Tensor, merge_from: Any = None, momentum: bool = False) -> int:
        fam, dims = self._family
        p = self.optimizer_params
        lr = float(p.get("lr", 1e-3))
        wd = float(p.get("weight_decay", 0.0))
        fn = ops.mlp1_train if fam == "mlp1" else ops.logreg_train
        if x.dim() > 2:
            x = x.reshape(x.shape[0], -1)
        if momentum:          # torch.optim.SGD(momentum=...): the buffer row is the one the flat optimizer kernel uses
            buf = self._opt_rows.get("momentum")
            first = buf is None or bool(self.__dict__.pop("_mom_pending", False))   # (a buffer the C++ executor created but never used)
            if buf is None:
                buf = self._opt_rows["momentum"] = torch.zeros_like(self.row)
            mom = (float(p.get("momentum", 0.0)), float(p.get("dampening", 0.0)), bool(p.get("nesterov", False)), buf, first)
            return fn(self.row, x, y, dims, self.batch_size, self.local_epochs, lr, wd, self._next_key(), None,
                      merge_from=merge_from, momentum=mom)
        return fn(self.row, x, y, dims, self.batch_size, self.local_epochs, lr, wd,
                  self._next_key(), self._elem_scale(), merge_from=merge_from)

    # -- fused MERGE_UPDATE: the merge rides on the training kernel's weight load ------------
    def _fused_merge_weights(self, other: Any) -> Optional[Tuple[float, float]]:
        """``(w_self, w_other)`` of a pairwise merge that may be folded into the local-update
        kernel, or ``None`` when this handler's merge is not a plain weighted pair merge."""
        if type(self)._merge is not TorchModelHandler._merge:
            return None
        return (0.5, 0.5) if isinstance(other, TorchModelHandler) else None

    def _merge_update_fused(self, recv_model: Any, data: Any) -> bool:
        if not (self._fused and FUSE_MERGE_UPDATE) or self.layout.int_buffers:
            return False
        if GlobalSettings().reference_compat and (self.__dict__.get("_opt_detached") or self.batch_size):
            # B13 mimicked: merge, then an update whose optimizer step is lost (see _update); mini-batches: merge, then
            # the autograd path with the reference's own shuffles (the fused kernels shuffle with the keyed permutation)
            return False
        w = self._fused_merge_weights(recv_model)
        if w is None:
            return False
        new_age = max(self.n_updates, recv_model.n_updates)

        def run(src, sync):
            self.n_updates = new_age
            self._update(data, merge_from=(src, w[0], w[1], sync))
        if self._mine():
            self._pull(recv_model, run)
        else:
            self._pull(recv_model, run)     # bookkeeping of the read ...
            self.n_updates = new_age
            self._update(data)              # ... and of the update
        return True

    def _ensure_grad(self) -> torch.Tensor:
        mod = self.model
        if self._grad_row is None or self._grad_row.device != self.row.device:
            self._grad_row = torch.zeros_like(self.row)
            self._grad_bound = False
        if not getattr(self, "_grad_bound", False):
            self.layout.bind(mod, self.row, self._grad_row)
            self._grad_bound = True
        return self._grad_row

    def _update_generic(self, x: torch.Tensor, y: torch.Tensor) -> int:
        mod = self.model
        nhwc = False
        if self.layout.channels_last and x.dim() == 4 and not GlobalSettings().reference_compat:
            # the shard as [N, H, W, C] (one pass per update): a gathered mini-batch viewed back as [B, C, H, W] is a
            # channels-last tensor, so no layer has to convert layouts
            x, nhwc = x.permute(0, 2, 3, 1).contiguous(), True
        n = x.size(0)
        bs = n if not self.batch_size else self.batch_size
        gen_key = self._next_key()
        steps = 0
        if GlobalSettings().reference_compat:
            # the reference's shuffles, call for call on torch's global stream (handler.py:238-247): every epoch permutes
            # the ALREADY permuted arrays
            if self.local_epochs > 0:
                for _ in range(self.local_epochs):
                    perm = torch.randperm(n).to(x.device)
                    x, y = x[perm], y[perm]
                    for i in range(0, n, bs):
                        self._local_step(mod, x[i:i + bs], y[i:i + bs])
                        steps += 1
            else:
                perm = torch.randperm(n).to(x.device)
                self._local_step(mod, x[perm][:bs], y[perm][:bs])
                steps = 1
            return steps
        if self.local_epochs > 0:
            for e in range(self.local_epochs):
                perm = ops.keyed_perm(n, _rng.mix64(gen_key ^ e), x.device)
                for i in range(0, n, bs):
                    self._local_step(mod, x, y, perm[i:i + bs], nhwc)
                    steps += 1
        else:
            perm = ops.keyed_perm(n, _rng.mix64(gen_key), x.device)
            self._local_step(mod, x, y, perm[:bs], nhwc)
            steps = 1
        return steps

    #: whole-model merges only: handlers that address parameters by flat index (sampling, partitions) keep the plain order
    _ROW_CHANNELS_LAST_OK = True

    def _local_step(self, mod: TorchModel, x: torch.Tensor, y: torch.Tensor,
                    idx: Optional[torch.Tensor] = None, nhwc: bool = False) -> None:
        """forward / loss / backward / optimizer step (ref ``handler.py:250-258``) on the mini-batch ``x[idx], y[idx]``
        (``idx = None``: all of ``x, y``).  On a GPU the forward + backward is replayed from a CUDA graph."""
        mod.train()
        g = self._ensure_grad()
        if not self._graph_fwd_bwd(mod, g, x, y, idx, nhwc):
            if idx is not None:
                x, y = x[idx], y[idx]
            g.zero_()
            loss = self.criterion(mod(x.permute(0, 3, 1, 2) if nhwc else x), y)
            loss.backward()
        self._pre_step()
        self._apply_optimizer(g)

    #: eager steps of a (handler, batch shape) before its graph is captured: cuDNN / cuBLAS handles, workspaces and
    #: autotuned algorithms must exist before a capture starts
    _GRAPH_WARMUP = 2

    def _graph_fwd_bwd(self, mod: TorchModel, g: torch.Tensor, x: torch.Tensor, y: torch.Tensor,
                       idx: Optional[torch.Tensor], nhwc: bool = False) -> bool:
        """Zero the gradient row, forward, loss, backward of one mini-batch as ONE graph launch.

        A generic model's step is a few hundred small kernels (ResNet-20: ~600) whose launches from Python, not
        their execution, bound the update on a B200.  The parameters are views of the handler's arena row and the
        gradients views of its gradient row, so every address in the step is fixed: after ``_GRAPH_WARMUP`` eager
        steps the step is captured per (row, gradient row, module, batch shape) with the mini-batch gathered into
        static input buffers (one ``index_select`` per tensor), and replayed afterwards.  The optimizer stays outside
        (one fused launch, host-side step counters / hooks keep working).  Graphs of the handlers of one stream share
        a memory pool -- replays on a stream are ordered and nothing allocated under capture outlives the step.
        Returns ``False`` when the step has to run eagerly."""
        if not (g.is_cuda and GlobalSettings().cuda_graphs) or torch.cuda.is_current_stream_capturing():
            return False
        if _prt.active():
            # One rank only.  With several ranks a node's stream may hold a kernel that waits on another GPU's flag; the
            # replayed steps keep the GPU full of cuDNN kernels, some of them persistent (all CTAs co-resident), which then
            # queue behind the waiting kernels while the snapshot the peer needs queues behind them -- on both GPUs at once.
            # Measured on 2 GPUs (BASELINE config 5): replayed steps run into the bounded waits, eager steps (the GPU is
            # mostly idle between the host's launches) do not.  A capture would also synchronise the device.
            return False
        nb = int(idx.numel()) if idx is not None else int(x.size(0))
        key = (self.row.data_ptr(), g.data_ptr(), id(mod), nb, tuple(x.shape[1:]), x.dtype, tuple(y.shape[1:]),
               y.dtype, x.device.index, nhwc)
        cache = self.__dict__.get("_graphs")
        if cache is None:
            cache = self.__dict__["_graphs"] = {}
        ent = cache.get(key)
        if ent is None:
            if len(cache) >= 4:          # the row moved / the batch size changed: drop the stale captures
                cache.clear()
            ent = cache[key] = _GraphStep()
        if ent.failed:
            return False
        if ent.graph is None:
            ent.seen += 1
            if ent.seen <= self._GRAPH_WARMUP:
                return False

            ent.x = torch.empty((nb,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
            ent.y = torch.empty((nb,) + tuple(y.shape[1:]), dtype=y.dtype, device=y.device)
            ent.fill(x, y, idx)
            graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(graph, pool=_graph_pool(x.device), capture_error_mode="thread_local"):
                    g.zero_()
                    loss = self.criterion(mod(ent.x.permute(0, 3, 1, 2) if nhwc else ent.x), ent.y)
                    loss.backward()
                    del loss
            except Exception as err:            # an op that cannot be captured (host sync, CPU tensors, ...)
                ent.failed = True
                ent.error = "%s: %s" % (type(err).__name__, err)
                ent.x = ent.y = None
                LOG.warning("CUDA-graph capture of a %s step failed, this batch shape runs eagerly: %s"
                            % (type(mod).__name__, ent.error))
                return False
            ent.graph = graph
        else:
            ent.fill(x, y, idx)
        ent.graph.replay()
        return True

    def _pre_step(self) -> None:
        """Hook between backward and the optimizer step (gradient adjustment)."""

    def _apply_optimizer(self, g: torch.Tensor) -> None:
        p = self.optimizer_params
        n = self.layout.n_params
        row = self.row
        kind = self._opt_kind
        if kind in ("sgd_plain", "sgd_momentum"):
            mom = float(p.get("momentum", 0.0))
            buf = None
            first = False
            if mom:
                buf = self._opt_rows.get("momentum")
                if buf is None:
                    buf = self._opt_rows["momentum"] = torch.zeros_like(row)
                    first = True
            ops.sgd_step(row, g, n, float(p.get("lr", 1e-3)), float(p.get("weight_decay", 0.0)),
                         mom, buf, float(p.get("dampening", 0.0)), bool(p.get("nesterov", False)),
                         first, self._grad_scale())
        elif kind in ("adam", "adamw"):
            m = self._opt_rows.get("exp_avg")
            if m is None:
                m = self._opt_rows["exp_avg"] = torch.zeros_like(row)
                self._opt_rows["exp_avg_sq"] = torch.zeros_like(row)
                self._opt_steps = 0
            self._opt_steps += 1
            b1, b2 = p.get("betas", (0.9, 0.999))
            default_wd = 0.01 if kind == "adamw" else 0.0
            ops.adam_step(row, g, n, m, self._opt_rows["exp_avg_sq"], self._opt_steps,
                          float(p.get("lr", 1e-3)), float(b1), float(b2), float(p.get("eps", 1e-8)),
                          float(p.get("weight_decay", default_wd)), kind == "adamw")
        else:
            if self._torch_opt is None:
                self._torch_opt = self.optimizer_cls(self.model.parameters(), **p)
            self._torch_opt.step()

    def _grad_scale(self) -> Optional[torch.Tensor]:
        return None

    # -- merging -----------------------------------------------------------------------------
    def _merge(self, other_model_handler: Union["TorchModelHandler", Iterable["TorchModelHandler"]]
               ) -> None:
        """Uniform average with one or several models; age = max (ref ``handler.py:260-280``)."""
        if isinstance(other_model_handler, TorchModelHandler):
            self._weighted_merge(other_model_handler, 0.5, 0.5)
            self.n_updates = max(self.n_updates, other_model_handler.n_updates)
        else:
            others = list(other_model_handler)
            k = len(others) + 1
            self._kway_merge(others, [1.0 / k] * k)
            self.n_updates = max([self.n_updates] + [o.n_updates for o in others])
        self._merge_int_buffers(other_model_handler)

    def _kway_merge(self, others: Sequence["TorchModelHandler"], weights: Sequence[float]) -> None:
        self._pull(list(others), lambda srcs, syncs: ops.merge_kway(self.row, srcs, weights, syncs))
        self._version += 1

    def _merge_int_buffers(self, other: Any) -> None:
        """Integer buffers (BN ``num_batches_tracked``) are combined with ``max``."""
        if not self.layout.int_buffers:
            return
        if _prt.active():
            # several ranks: the counters follow the same recurrences as the model age (max on merge,
            # += steps on update, both start at 0), which every rank already replicates -- derive them
            if self._mine():
                for name, buf in self.model.named_buffers():
                    if name in self.layout.int_buffers:
                        buf.fill_(int(np.max(self.n_updates)))
            return
        others = [other] if isinstance(other, TorchModelHandler) else list(other)
        mine = dict(self.model.named_buffers())
        for o in others:
            theirs = o._int_values()
            for name in self.layout.int_buffers:
                mine[name].copy_(torch.maximum(mine[name], theirs[name].to(mine[name].device)))

    def _int_values(self) -> Dict[str, torch.Tensor]:
        """Current integer-buffer values (live module if there is one, else the carried copy)."""
        if not self.layout.int_buffers:
            return {}
        if self._module is not None:
            bufs = dict(self._module.named_buffers())
            return {n: bufs[n].detach().clone() for n in self.layout.int_buffers}
        if self._int_state is not None:
            return self._int_state
        bufs = dict(self._proto.named_buffers())
        return {n: bufs[n].detach().clone() for n in self.layout.int_buffers}

    # -- evaluation -----------------------------------------------------------------------------
    def _forward_scores(self, x: torch.Tensor) -> torch.Tensor:
        if self.layout.channels_last and x.dim() == 4:
            x = x.contiguous(memory_format=torch.channels_last)
        mod = self._module
        if mod is not None:
            self._bind()
            mod.eval()
            with torch.no_grad():
                return mod(x)
        proto = self._proto
        if self.layout.entries and next(proto.parameters()).device != self.row.device:
            proto.to(self.row.device)
        views = self.layout.views(self.row)
        was_training = proto.training
        proto.eval()
        try:
            with torch.no_grad():
                return torch.func.functional_call(proto, views, (x,))
        finally:
            proto.train(was_training)

    def evaluate_async(self, data: Tuple[torch.Tensor, torch.Tensor]) -> PendingEval:
        """accuracy / macro precision / recall / F1 (+AUC for 2 outputs) -- ref ``:282-334``."""
        if not self._mine():
            return _NOT_MINE
        x, y = self._to_device(data)
        if y.dim() > 1:
            y = torch.argmax(y, dim=-1)
        s = self._stream()
        with _arena.on_stream(s):
            _arena.before_read(self._ensure_row(), s if s is not None else _arena.current(self.device))
            fam = self._family
            auc_in = None
            if fam is not None and fam[0] == "mlp1":
                n_out = fam[1][2]
                xx = x.reshape(x.shape[0], -1) if x.dim() > 2 else x
                if n_out == 2:       # the evaluation kernel also writes the class-1 logit (AUC), no eager forward pass
                    cm, auc_in = ops.mlp1_eval(self.row, xx, y, fam[1], n_out, want_scores=True)
                else:
                    cm = ops.mlp1_eval(self.row, xx, y, fam[1], n_out)
            else:
                if fam is not None and fam[0] == "logreg":
                    scores = ops.logreg_scores(self.row, x, fam[1])
                else:
                    scores = self._forward_scores(x)
                n_out = scores.shape[1]
                cm = ops.torch_ref.confusion_matrix(y, scores.argmax(dim=-1), n_out)
                if n_out == 2:
                    auc_in = scores[:, 1]
            auc_t = None
            if auc_in is not None:
                auc_t = (y == 1, auc_in.detach())
            done = _mark(self.device)

        def finish() -> Dict[str, float]:
            res = _metrics.classification_report(cm.cpu().numpy())
            if auc_t is not None:
                pos, sc = auc_t
                if int(pos.sum()) in (0, pos.numel()):
                    LOG.warning("# of classes != 2. AUC is set to 0.5.")
                    res["auc"] = 0.5
                else:
                    res["auc"] = _metrics.roc_auc(pos, sc)
            return res
        return PendingEval(finish, done)

    def evaluate(self, data: Tuple[torch.Tensor, torch.Tensor]) -> Dict[str, float]:
        return self.evaluate_async(data).result()

    def __str__(self) -> str:
        return "%s(model=%s_%s, mode=%s)" % (self.__class__.__name__, str(self._proto),
                                             self.n_updates, self.mode)


# --------------------------------------------------------------------------------------
# AdaLine / Pegasos
# --------------------------------------------------------------------------------------
class AdaLineHandler(RowHandler):
    """Widrow-Hoff learner on a bare weight vector (ref ``handler.py:337-391``)."""

    def __init__(self, net: AdaLine, learning_rate: float,
                 create_model_mode: CreateModelMode = CreateModelMode.UPDATE,
                 copy_model: bool = True) -> None:
        super().__init__(create_model_mode)
        self._module = copy.deepcopy(net) if copy_model else net
        self.dim = self._module.input_dim
        self.learning_rate = learning_rate
        self._row_numel = max(32, (self.dim + 31) // 32 * 32)

    @property
    def model(self) -> AdaLine:
        mod = self._module
        w = self.row[:self.dim]
        if mod.model.data_ptr() != w.data_ptr():
            mod.model.data = w
        return mod

    @model.setter
    def model(self, value: Any) -> None:
        if value is None:
            return
        self._module = value
        if self._row is not None:
            self.row[:self.dim].copy_(value.model.detach().to(self.row.device))
            self._version += 1

    def _on_new_row(self, old) -> None:
        if old is None and not self._is_snapshot and self._mine():
            self._row.tensor.zero_()
            self._row.tensor[:self.dim].copy_(self._module.model.detach().to(self._row.tensor.device))

    def _after_clone(self, src: "AdaLineHandler", light: bool) -> None:
        # a materialised source's values arrive through the row copy; otherwise keep its net
        self._module = AdaLine(src.dim) if src._row is not None else copy.deepcopy(src._module)

    def __getstate__(self) -> Dict[str, Any]:
        st = super().__getstate__()
        st["_module"] = AdaLine(self.dim)
        return st

    def init(self) -> None:
        self._version += 1
        if not self._mine():
            self._ensure_row()
            return
        self.model.init_weights()

    def get_size(self) -> int:
        return self.dim

    def _w(self) -> torch.Tensor:
        return self.row[:self.dim]

    def _update(self, data: Tuple[torch.Tensor, torch.Tensor]) -> None:
        self.n_updates += len(data[1])
        self._version += 1
        if not self._mine():
            self._ensure_row()
            return
        x, y = self._to_device(data)
        with _arena.on_stream(self._stream()):
            ops.adaline_update(self._w(), x, y.to(torch.float32), self.learning_rate)

    def _merge(self, other_model_handler: "AdaLineHandler") -> None:
        self._weighted_merge(other_model_handler, 0.5, 0.5)
        self.n_updates = max(self.n_updates, other_model_handler.n_updates)

    def evaluate_async(self, data: Tuple[torch.Tensor, torch.Tensor]) -> PendingEval:
        if not self._mine():
            return _NOT_MINE
        x, y = self._to_device(data)
        with _arena.on_stream(self._stream()):
            scores = x.float() @ self._w()
            pred_pos = scores >= 0
            true_pos = y > 0
            cm = ops.torch_ref.confusion_matrix(true_pos.long(), pred_pos.long(), 2)
            sc = scores.detach()
            done = _mark(self.device)

        def finish() -> Dict[str, float]:
            res = _metrics.classification_report(cm.cpu().numpy())
            res["auc"] = _metrics.roc_auc(true_pos, sc)
            return res
        return PendingEval(finish, done)

    def evaluate(self, data: Tuple[torch.Tensor, torch.Tensor]) -> Dict[str, float]:
        return self.evaluate_async(data).result()


class PegasosHandler(AdaLineHandler):
    """Pegasos SVM steps, one per local sample (ref ``handler.py:394-423``)."""

    def _update(self, data: Tuple[torch.Tensor, torch.Tensor]) -> None:
        self._version += 1
        if not self._mine():
            self._ensure_row()
            self.n_updates = int(self.n_updates) + len(data[1])
            return
        x, y = self._to_device(data)
        with _arena.on_stream(self._stream()):
            self.n_updates = ops.pegasos_update(self._w(), x, y.to(torch.float32),
                                                self.learning_rate, int(self.n_updates))


# --------------------------------------------------------------------------------------
# sampled / partitioned / weighted / limited merges
# --------------------------------------------------------------------------------------
class SamplingTMH(TorchModelHandler):
    _ROW_CHANNELS_LAST_OK = False
    """Merge only a random subset of coordinates (ref ``handler.py:426-452``)."""

    def __init__(self, sample_size: float, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.sample_size = sample_size

    def draw_sample(self) -> torch.Tensor:
        """Flat positions to merge, generated on the handler's device."""
        key = _rng.derive(0x5A3F, max(self.owner, 0), self._next_key() & 0xFFFFFFFF)
        k = TorchModelSampling.sample_size(self.sample_size, self.layout.n_params)
        # counter-based draws (uniform with replacement, like ``sample_flat``): the same sample on every device, and
        # reproducible by the C++ executor
        return ops.keyed_randint(k, self.layout.n_params, key, self.device)

    def _merge(self, other_model_handler: "SamplingTMH", sample: Any) -> None:
        if isinstance(sample, dict):
            sample = sample_dict_to_flat(sample, self._proto)
        self._pull(other_model_handler, lambda src, sync: ops.merge_indexed(
            self.row, src, sample.to(self.device), 0.5, 0.5, sync))
        self._version += 1

    def __call__(self, recv_model: Any, data: Any, sample: Any) -> None:
        if self.mode == CreateModelMode.PASS:
            raise ValueError("Mode PASS not allowed for sampled models.")
        if self.mode == CreateModelMode.UPDATE:
            tmp = self._scratch_copy(recv_model)
            tmp._update(data)
            self._merge(tmp, sample)
            _dispose(tmp)
        else:
            super().__call__(recv_model, data, sample)


class PartitionedTMH(TorchModelHandler):
    _ROW_CHANNELS_LAST_OK = False
    """Per-partition ages and merges (ref ``handler.py:455-525``).

    ``n_updates`` is an int array with one age per partition.  Every local step increments all
    ages and divides the gradient of partition ``p`` by its age (a per-partition ``1/t``
    learning-rate decay) -- folded into the fused kernels' SGD epilogue / the flat optimizer.
    """

    def __init__(self, net: TorchModel, tm_partition: TorchModelPartition, optimizer: Any,
                 optimizer_params: Dict[str, Any], criterion: Callable, local_epochs: int = 1,
                 batch_size: int = 32,
                 create_model_mode: CreateModelMode = CreateModelMode.MERGE_UPDATE,
                 copy_model: bool = True) -> None:
        super().__init__(net, optimizer, optimizer_params, criterion, local_epochs, batch_size,
                         create_model_mode, copy_model)
        self.tm_partition = tm_partition
        self.n_updates = np.zeros(tm_partition.n_parts, dtype=int)
        self._part_id_dev: Optional[torch.Tensor] = None
        self._seg_dev: Dict[int, torch.Tensor] = {}

    def _part_ids(self) -> torch.Tensor:
        if self._part_id_dev is None or self._part_id_dev.device != self.row.device:
            self._part_id_dev = self.tm_partition.part_id.to(self.row.device)
        return self._part_id_dev

    def _elem_scale(self) -> Tuple[torch.Tensor, torch.Tensor]:
        ages = torch.as_tensor(np.asarray(self.n_updates, dtype=np.int64))
        dev = self.row.device
        if dev.type == "cuda" and ages.numel() > 16:
            # through pinned memory, asynchronously on the node's stream: a pageable H2D copy would block the host
            # until everything queued on that stream has finished (the caching host allocator keeps the staging
            # buffer alive until the copy has run)
            ages = ages.pin_memory().to(dev, non_blocking=True)
        # (<= 16 partitions: the ages stay on the host and travel by value in the kernel's launch parameters)
        return (self._part_ids(), ages)

    def _count_steps(self, steps: int) -> None:
        if self._fused:
            self.n_updates = self.n_updates + steps  # generic path counts per step

    def _pre_step(self) -> None:
        self.n_updates = self.n_updates + 1

    def _grad_scale(self) -> torch.Tensor:
        ages = torch.as_tensor(self.n_updates, dtype=torch.float32)
        if self.row.is_cuda:        # through pinned memory: a pageable upload would block the host until this stream drained
            ages = ages.pin_memory().to(self.row.device, non_blocking=True)
        inv = 1.0 / ages
        scale = torch.ones_like(self.row)
        scale[:self.layout.n_params] = inv[self._part_ids()]
        return scale

    def __call__(self, recv_model: Any, data: Any, id_part: int) -> None:
        if self.mode == CreateModelMode.PASS:
            raise ValueError("Mode PASS not allowed for partitioned models.")
        if self.mode == CreateModelMode.UPDATE:
            tmp = self._scratch_copy(recv_model)
            tmp._update(data)
            self._merge(tmp, id_part)
            _dispose(tmp)
        else:
            super().__call__(recv_model, data, id_part)

    def _merge(self, other_model_handler: "PartitionedTMH", id_part: int) -> None:
        pid = id_part % self.tm_partition.n_parts
        a, b = int(self.n_updates[pid]), int(other_model_handler.n_updates[pid])
        w1, w2 = TorchModelPartition.mixing_weights((a, b))
        def run(src, sync):
            seg = self._seg_dev.get(pid)
            if seg is None or seg.device != self.row.device:
                seg = self._seg_dev[pid] = self.tm_partition.segments(pid).to(self.row.device)
            ops.merge_segments(self.row, src, seg, w1, w2, sync)
        self._pull(other_model_handler, run)
        self.n_updates[pid] = max(a, b)
        self._version += 1

    def _age_key(self) -> str:
        return str(self.n_updates)

    def _after_clone(self, src: "PartitionedTMH", light: bool) -> None:
        super()._after_clone(src, light)
        self.n_updates = np.array(src.n_updates, copy=True)

    def __getstate__(self) -> Dict[str, Any]:
        st = super().__getstate__()
        st["_part_id_dev"] = None
        st["_seg_dev"] = {}
        return st


class WeightedTMH(TorchModelHandler):
    """Weighted neighbourhood averaging for decentralised SGD (ref ``handler.py:642-688``)."""

    def __call__(self, recv_model: Any, data: Any, weights: Iterable[float]) -> None:
        if self.mode == CreateModelMode.UPDATE:
            super().__call__(recv_model, data)
        elif self.mode == CreateModelMode.MERGE_UPDATE:
            self._merge(recv_model, weights)
            self._update(data)
        elif self.mode == CreateModelMode.UPDATE_MERGE:
            self._update(data)
            recvs = list(recv_model) if not isinstance(recv_model, TorchModelHandler) else [recv_model]
            tmps = [self._scratch_copy(r) for r in recvs]
            for t in tmps:
                t._update(data)
            self._merge(tmps if not isinstance(recv_model, TorchModelHandler) else tmps[0], weights)
            for t in tmps:
                _dispose(t)
        else:
            raise ValueError("Invalid create model mode %s for WeightedTMH." % str(self.mode))

    def _merge(self, other_model_handler: Union[TorchModelHandler, Iterable[TorchModelHandler]],
               weights: Iterable[float]) -> None:
        w = [float(x) for x in weights]
        others = ([other_model_handler] if isinstance(other_model_handler, TorchModelHandler)
                  else list(other_model_handler))
        self._kway_merge(others, w[:len(others) + 1])
        self.n_updates = max([self.n_updates] + [o.n_updates for o in others])
        self._merge_int_buffers(others)


class LimitedMergeMixin:
    """Age-limited merge of Danner et al. 2023 (ref ``handler.py:690-715``)."""

    def __init__(self, age_diff_threshold: int = 1) -> None:
        self.L = age_diff_threshold

    def _merge(self, other_model_handler: Any) -> None:
        if not isinstance(other_model_handler, TorchModelHandler):
            raise ValueError("Invalid type for other_model_handler: %s" % type(other_model_handler))
        a, b = self.n_updates, other_model_handler.n_updates
        if a > b + self.L:
            pass  # own model is much older: keep it
        elif b > a + self.L:
            self._adopt(other_model_handler)
        else:
            tot = a + b
            if tot == 0:
                self._weighted_merge(other_model_handler, 0.5, 0.5)
            else:
                self._weighted_merge(other_model_handler, a / tot, b / tot)
        self.n_updates = max(a, b)


class LimitedMergeTMH(LimitedMergeMixin, TorchModelHandler):
    def _fused_merge_weights(self, other: Any) -> Optional[Tuple[float, float]]:
        if not isinstance(other, TorchModelHandler):
            return None
        a, b = self.n_updates, other.n_updates
        if a > b + self.L:
            return (1.0, 0.0)
        if b > a + self.L:
            return (0.0, 1.0)
        tot = a + b
        return (0.5, 0.5) if tot == 0 else (a / tot, b / tot)

    def __init__(self, net: TorchModel, optimizer: Any, optimizer_params: Dict[str, Any],
                 criterion: Callable, local_epochs: int = 1, batch_size: int = 32,
                 create_model_mode: CreateModelMode = CreateModelMode.MERGE_UPDATE,
                 age_diff_threshold: int = 1, copy_model: bool = True) -> None:
        TorchModelHandler.__init__(self, net, optimizer, optimizer_params, criterion,
                                   local_epochs, batch_size, create_model_mode, copy_model)
        LimitedMergeMixin.__init__(self, age_diff_threshold)


# --------------------------------------------------------------------------------------
# matrix factorisation
# --------------------------------------------------------------------------------------
class MFModelHandler(RowHandler):
    """Rank-k matrix factorisation of one user's ratings (ref ``handler.py:528-576``).

    Row layout ``[Y (n_items*k) | c (n_items) | X (k) | b (1)]``: the shared item factors are a
    contiguous prefix, so the merge (item side only) is one ranged merge kernel.
    """

    def __init__(self, dim: int, n_items: int, lam_reg: float = 0.1, learning_rate: float = 0.001,
                 create_model_mode: CreateModelMode = CreateModelMode.UPDATE) -> None:
        super().__init__(create_model_mode)
        self.reg, self.k, self.lr, self.n_items = lam_reg, dim, learning_rate, n_items
        self.n_updates = 1
        n = n_items * dim + n_items + dim + 1
        self._n_shared = n_items * dim + n_items
        self._row_numel = (n + 31) // 32 * 32

    def _parts(self):
        r, k, m = self.row, self.k, self.n_items
        Y = r[:m * k].view(m, k)
        c = r[m * k:m * k + m]
        X = r[m * k + m:m * k + m + k]
        b = r[m * k + m + k:m * k + m + k + 1]
        return X, b, Y, c

    @property
    def model(self):
        if self._row is None:
            return None
        X, b, Y, c = self._parts()
        return ((X.cpu().numpy().reshape(1, -1), float(b[0])), (Y.cpu().numpy(), c.cpu().numpy()))

    @model.setter
    def model(self, value: Any) -> None:
        if value is None:
            return
        (Xn, bn), (Yn, cn) = value
        X, b, Y, c = self._parts()
        X.copy_(torch.as_tensor(np.asarray(Xn).reshape(-1), dtype=torch.float32))
        b.fill_(float(bn))
        Y.copy_(torch.as_tensor(np.asarray(Yn), dtype=torch.float32))
        c.copy_(torch.as_tensor(np.asarray(cn), dtype=torch.float32))
        self._version += 1

    def init(self, r_min: int = 1, r_max: int = 5) -> None:
        mul = float(np.sqrt((r_max - r_min) / self.k))
        self._version += 1
        if not self._mine():
            self._ensure_row()
            return
        X, b, Y, c = self._parts()
        if GlobalSettings().reference_compat:       # the reference's draws on the NumPy stream (handler.py:543-545)
            mul64 = np.sqrt((r_max - r_min) / self.k)
            X.copy_(torch.as_tensor(np.random.rand(1, self.k) * mul64, dtype=torch.float32).reshape(-1))
            Y.copy_(torch.as_tensor(np.random.rand(self.n_items, self.k) * mul64, dtype=torch.float32))
        else:
            gen = torch.Generator().manual_seed(_rng.derive(0x3F, max(self.owner, 0)))
            X.copy_(torch.rand(self.k, generator=gen) * mul)
            Y.copy_(torch.rand(self.n_items, self.k, generator=gen) * mul)
        b.fill_(r_min / 2.0)
        c.fill_(r_min / 2.0)

    def _update(self, data: Any) -> None:
        if not self._mine():
            self._ensure_row()
            self.n_updates += len(data)
            self._version += 1
            return
        ratings = self._to_device(data)
        if isinstance(ratings, (list, tuple)):
            ratings = torch.as_tensor(np.asarray(ratings), dtype=torch.float32, device=self.device)
        ratings = ratings.to(torch.float32).reshape(-1, 2)
        X, b, Y, c = self._parts()
        with _arena.on_stream(self._stream()):
            self.n_updates += ops.mf_update(X, b, Y, c, ratings, self.reg, self.lr)
        self._version += 1

    def _merge(self, other_model_handler: "MFModelHandler") -> None:
        """Item factors only: ``Y = (Y n + Y' n') / (2 (n + n'))`` -- the extra 1/2 is the
        reference's (``handler.py:566-567``) and is kept by default because it shapes
        the published-style curves."""
        a, b = self.n_updates, other_model_handler.n_updates
        den = 2.0 * (a + b)
        self._weighted_merge(other_model_handler, a / den, b / den, 0, self._n_shared)

    def evaluate_async(self, ratings: Any) -> PendingEval:
        """RMSE on the node's ratings (ref ``handler.py:569-573``); the scalar stays on the device until ``result()``."""
        if not self._mine():
            return _NOT_MINE
        r = self._to_device(ratings)
        if not isinstance(r, torch.Tensor):
            r = torch.as_tensor(np.asarray(r), dtype=torch.float32, device=self.device)
        r = r.to(torch.float32).reshape(-1, 2)
        X, b, Y, c = self._parts()
        s = self._stream()
        with _arena.on_stream(s):
            _arena.before_read(self._ensure_row(), s if s is not None else _arena.current(self.device))
            idx = r[:, 0].long()
            pred = Y[idx] @ X + b + c[idx]
            mse = torch.mean((r[:, 1] - pred) ** 2)
            ev = _mark(self.device)
        return PendingEval(lambda: {"rmse": float(torch.sqrt(mse))}, ev)

    def evaluate(self, ratings: Any) -> Dict[str, float]:
        return self.evaluate_async(ratings).result()

    def get_size(self) -> int:
        return self.k * (self.n_items + 1)


# --------------------------------------------------------------------------------------
# k-means
# --------------------------------------------------------------------------------------
class KMeansHandler(RowHandler):
    """Online k-means with centroid gossip (ref ``handler.py:579-639``)."""

    def __init__(self, k: int, dim: int, alpha: float = 0.1, matching: str = "naive",
                 create_model_mode: CreateModelMode = CreateModelMode.UPDATE) -> None:
        assert matching in {"naive", "hungarian"}, "Invalid matching method."
        super().__init__(create_model_mode)
        self.k, self.dim, self.matching, self.alpha = k, dim, matching, alpha
        self._row_numel = max(32, (k * dim + 31) // 32 * 32)

    @property
    def model(self) -> Optional[torch.Tensor]:
        if self._row is None:
            return None
        return self.row[:self.k * self.dim].view(self.k, self.dim)

    @model.setter
    def model(self, value: Any) -> None:
        if value is None:
            return
        self.row[:self.k * self.dim].copy_(torch.as_tensor(value, dtype=torch.float32).reshape(-1))
        self._version += 1

    def init(self) -> None:
        if not self._mine():
            self._ensure_row()
            self._version += 1
            return
        if GlobalSettings().reference_compat:       # the reference draws from the global torch stream (handler.py:595)
            self.model = torch.rand(size=(self.k, self.dim))
            return
        gen = torch.Generator().manual_seed(_rng.derive(0x4B, max(self.owner, 0)))
        self.model = torch.rand(self.k, self.dim, generator=gen)

    def _update(self, data: Tuple[torch.Tensor, Any]) -> None:
        if not self._mine():
            self._ensure_row()
            self.n_updates += 1
            self._version += 1
            return
        x, _ = self._to_device(data)
        with _arena.on_stream(self._stream()):
            ops.kmeans_update(self.model, x.float().reshape(-1, self.dim), self.alpha)
        self.n_updates += 1
        self._version += 1

    def _merge(self, other_model_handler: "KMeansHandler") -> None:
        """Average centroids, optionally after optimal (Hungarian) matching.

        FIX(B16): the reference indexes with the *row* assignment (an identity), so its
        "hungarian" equals "naive"; here the peer's centroids are permuted by the column
        assignment.  With ``reference_compat`` the identity behaviour is reproduced."""
        if self.matching == "naive" or GlobalSettings().reference_compat:
            self._weighted_merge(other_model_handler, 0.5, 0.5)
            return
        def run(src, sync):
            # k <= 8 on a GPU: exhaustive optimal matching + merge in ONE kernel (cross-rank handshake included);
            # otherwise cdist -> scipy.linear_sum_assignment on the host
            ops.kmeans_match_merge(self.row, src, self.k, self.dim, 0.5, 0.5, sync)
        self._pull(other_model_handler, run)
        self._version += 1

    def evaluate_async(self, data: Tuple[torch.Tensor, torch.Tensor]) -> PendingEval:
        """NMI of the cluster assignment (ref ``handler.py:632-636``): the contingency table is built on the device and
        read when the result is asked for."""
        if not self._mine():
            return _NOT_MINE
        X, y = self._to_device(data)
        yl = y.long().reshape(-1)
        y_host = data[1]
        if isinstance(y_host, torch.Tensor) and y_host.is_cuda:
            n_true = int(yl.max()) + 1 if yl.numel() else 1       # labels only exist on the device: one read
        else:                                                     # number of classes from the host copy: no device read
            n_true = int(np.asarray(y_host).max()) + 1 if yl.numel() else 1
        s = self._stream()
        with _arena.on_stream(s):
            _arena.before_read(self._ensure_row(), s if s is not None else _arena.current(self.device))
            pred = ops.kmeans_assign(self.model, X.float().reshape(-1, self.dim))
            ct = torch.bincount(yl * self.k + pred.long(), minlength=n_true * self.k).view(n_true, self.k)
            ev = _mark(self.device)
        return PendingEval(lambda: {"nmi": _metrics.nmi_from_contingency(ct.cpu().numpy())}, ev)

    def evaluate(self, data: Tuple[torch.Tensor, torch.Tensor]) -> Dict[str, float]:
        return self.evaluate_async(data).result()

    def get_size(self) -> int:
        return self.k * self.dim
