"""Parameter-row arenas and per-node execution streams.

This replaces the reference's ``CACHE`` of deep-copied handlers (``gossipy/__init__.py:283-380``,
``gossipy/model/handler.py:160-176``): a model "on the wire" is one row of a pre-allocated HBM
arena, written by a D2D snapshot kernel on the sender's stream and later *pulled* by the
receiver's fused merge kernel (over NVLink when the row lives on another GPU).

Concurrency model (single process): every gossip node owns one CUDA stream; all writes to the
node's live row are ordered on it.  Snapshot rows are the only objects shared across streams;
they carry a ``ready`` event (recorded by the writer) and ``consumed`` events (recorded by
readers) so that disjoint node pairs overlap on the GPU while conflicting accesses serialise
without any host synchronisation.  On CPU all of this degenerates to plain tensors.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import Dict, List, Optional

import torch


class Row:
    """One parameter vector (a 1-D fp32 tensor view) plus its cross-stream / cross-rank dependencies.

    ``tensor`` is ``None`` for a row of another rank under the ``sendrecv`` transport and a
    peer-mapped view (CUDA IPC) under ``p2p``.  ``gen`` / ``remote_reads`` are replicated
    bookkeeping for the flag protocol (see :mod:`gossipy_b200.parallel.runtime`).
    """

    __slots__ = ("arena", "index", "tensor", "ready", "consumed", "stream", "rank", "gen",
                 "remote_reads", "flag_ready", "flag_done")

    def __init__(self, arena: "RowArena", index: int, tensor: Optional[torch.Tensor]) -> None:
        self.arena = arena
        self.index = index
        self.tensor = tensor
        self.ready = None        # event: last write finished
        self.consumed: List = []  # events: reads since the last write
        self.stream = None       # stream of the last writer
        self.rank = arena.rank
        self.gen = 0             # number of (shared) writes so far
        self.remote_reads = 0    # cumulative reads by other ranks
        self.flag_ready = 0      # device address of the owner's `ready` flag (p2p transport)
        self.flag_done = 0       # device address of the owner's `done` counter

    def release(self) -> None:
        self.arena.free(self)


class RowArena:
    """Pool of equally sized fp32 rows on one device, grown chunk-wise and reused FIFO.

    ``ghost=True`` mirrors the bookkeeping of another rank's arena without owning memory
    (``sendrecv`` transport); every rank replays the same alloc/free sequence on every mirror, so
    row indices agree everywhere without any communication.
    """

    def __init__(self, device: torch.device, row_numel: int, chunk_rows: int = 16,
                 rank: int = 0, ghost: bool = False) -> None:
        self.device = torch.device(device)
        self.row_numel = int(row_numel)
        self.chunk_rows = chunk_rows
        self.rank = rank
        self.ghost = ghost
        self._chunks: List[torch.Tensor] = []
        self._free: List[Row] = []
        self._n_rows = 0
        self.live = 0
        self.high_water = 0

    def _grow(self) -> None:
        chunk = None if self.ghost else torch.zeros(self.chunk_rows, self.row_numel,
                                                    dtype=torch.float32, device=self.device)
        if chunk is not None:
            self._chunks.append(chunk)
        for i in range(self.chunk_rows):
            self._free.append(Row(self, self._n_rows + i, None if chunk is None else chunk[i]))
        self._n_rows += self.chunk_rows
        self.chunk_rows = min(self.chunk_rows * 2, 1024)

    def alloc(self) -> Row:
        if not self._free:
            self._grow()
        row = self._free.pop(0)
        self.live += 1
        self.high_water = max(self.high_water, self.live)
        return row

    def free(self, row: Row) -> None:
        if row.index < 0:
            return
        self.live -= 1
        self._free.append(row)  # FIFO reuse: pending readers have usually long finished

    def nbytes(self) -> int:
        return self._n_rows * self.row_numel * 4


class SymmetricArenas:
    """Identical IPC-shared arenas on every rank (``p2p`` transport).

    Layout of each rank's allocation: ``capacity`` rows of ``row_numel`` floats followed by two
    uint32 flags per row (``ready`` generation, ``done`` read counter).  Creation is collective.
    """

    def __init__(self, device: torch.device, row_numel: int) -> None:
        import torch.distributed as dist
        from ..ops.native import native
        from ..parallel import runtime as prt
        nat = native()
        self.device = torch.device(device)
        self.row_numel = int(row_numel)
        row_bytes = self.row_numel * 4
        self.capacity = int(max(16, min(512, (1 << 31) // row_bytes)))
        self.flags_off = self.capacity * row_bytes
        total = self.flags_off + self.capacity * 8
        torch.cuda.set_device(self.device)
        self.base = nat.ipc_alloc(total)
        handles: List = [None] * prt.world()
        dist.all_gather_object(handles, nat.ipc_get_handle(self.base))
        self.bases = [self.base if r == prt.rank() else nat.ipc_open_handle(handles[r])
                      for r in range(prt.world())]
        dist.barrier()
        self.mirrors: List[RowArena] = []
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        for r in range(prt.world()):
            mirror = RowArena(self.device, row_numel, rank=r, ghost=True)
            mirror._n_rows = self.capacity
            for i in range(self.capacity):
                t = nat.tensor_from_ptr(self.bases[r] + i * row_bytes, [self.row_numel], dev_index, False)
                row = Row(mirror, i, t)
                row.flag_ready = self.bases[r] + self.flags_off + 8 * i
                row.flag_done = self.bases[r] + self.flags_off + 8 * i + 4
                mirror._free.append(row)
            mirror._grow = _no_growth  # type: ignore[assignment]
            self.mirrors.append(mirror)


def _no_growth() -> None:
    raise RuntimeError("symmetric arena exhausted: too many models in flight for the p2p transport")


_ARENAS: Dict[tuple, RowArena] = {}
_SYMMETRIC: Dict[tuple, SymmetricArenas] = {}


def arena_for(device: torch.device, row_numel: int, rank: Optional[int] = None) -> RowArena:
    """The arena (mirror) holding rows of ``row_numel`` floats that live on ``rank``."""
    from ..parallel import runtime as prt
    device = torch.device(device)
    if rank is None or not prt.active():
        rank = prt.rank() if prt.active() else 0
    key = (device.type, device.index, int(row_numel), rank)
    arena = _ARENAS.get(key)
    if arena is not None:
        return arena
    if prt.active() and prt.transport() == "p2p" and device.type == "cuda":
        skey = (device.index, int(row_numel))
        sym = _SYMMETRIC.get(skey)
        if sym is None:
            sym = _SYMMETRIC[skey] = SymmetricArenas(device, row_numel)
        for r, mirror in enumerate(sym.mirrors):
            _ARENAS[(device.type, device.index, int(row_numel), r)] = mirror
        return _ARENAS[key]
    ghost = prt.active() and rank != prt.rank()
    arena = _ARENAS[key] = RowArena(device, row_numel, rank=rank, ghost=ghost)
    return arena


def reset_arenas() -> None:
    _ARENAS.clear()
    _SYMMETRIC.clear()
    _STREAMS.clear()


# --------------------------------------------------------------------------------------
# streams
# --------------------------------------------------------------------------------------
_STREAMS: Dict[tuple, "torch.cuda.Stream"] = {}
_MAX_STREAMS = 32
multi_stream = True  # module switch: False serialises everything on the current stream


def stream_for(device: torch.device, owner: int):
    """The CUDA stream of gossip node ``owner`` on ``device`` (``None`` on CPU)."""
    device = torch.device(device)
    if device.type != "cuda" or not multi_stream or owner is None or owner < 0:
        return None
    key = (device.index, owner % _MAX_STREAMS)
    s = _STREAMS.get(key)
    if s is None:
        s = _STREAMS[key] = torch.cuda.Stream(device=device)
    return s


@contextmanager
def on_stream(stream):
    if stream is None:
        yield
    else:
        with torch.cuda.stream(stream):
            yield


def current(device: torch.device):
    device = torch.device(device)
    return torch.cuda.current_stream(device) if device.type == "cuda" else None


def before_read(row: Row, stream) -> None:
    """Make ``stream`` wait until ``row``'s content is complete."""
    if stream is None or row.ready is None or row.stream is stream:
        return
    stream.wait_event(row.ready)


def after_read(row: Row, stream) -> None:
    if stream is None:
        return
    ev = torch.cuda.Event()
    ev.record(stream)
    row.consumed.append(ev)


def before_write(row: Row, stream) -> None:
    """Make ``stream`` wait for the previous writer and all readers of ``row``."""
    if stream is None:
        return
    if row.ready is not None and row.stream is not stream:
        stream.wait_event(row.ready)
    for ev in row.consumed:
        stream.wait_event(ev)
    row.consumed = []


def after_write(row: Row, stream, shared: bool) -> None:
    row.stream = stream
    if stream is not None and shared:
        ev = torch.cuda.Event()
        ev.record(stream)
        row.ready = ev
    else:
        row.ready = None


def sync_all_streams(device: torch.device) -> None:
    """Join every node stream into the current stream (used before reading results)."""
    device = torch.device(device)
    if device.type != "cuda":
        return
    cur = torch.cuda.current_stream(device)
    for (dev_idx, _), s in _STREAMS.items():
        if dev_idx == device.index and s is not cur:
            ev = torch.cuda.Event()
            ev.record(s)
            cur.wait_event(ev)


def fork_from_current(device: torch.device) -> None:
    """Make every node stream wait for the work already enqueued on the current stream."""
    device = torch.device(device)
    if device.type != "cuda":
        return
    cur = torch.cuda.current_stream(device)
    ev = torch.cuda.Event()
    ev.record(cur)
    for (dev_idx, _), s in _STREAMS.items():
        if dev_idx == device.index and s is not cur:
            s.wait_event(ev)
