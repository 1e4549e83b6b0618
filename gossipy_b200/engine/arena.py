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

    For a row of another rank ``tensor`` is a peer-mapped view (CUDA IPC on GPUs, POSIX shared
    memory on CPU).  ``gen`` / ``remote_reads`` are replicated bookkeeping for the flag protocol
    (see :mod:`gossipy_b200.parallel.runtime`).
    """

    __slots__ = ("arena", "index", "tensor", "ready", "consumed", "stream", "rank", "gen",
                 "remote_reads", "flag_ready", "flag_done", "_acked", "in_use")

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
        self._acked = 0          # remote reads already waited for (owner side)
        self.in_use = False      # set by alloc, cleared by free (double-free guard)

    def release(self) -> None:
        self.arena.free(self)


class RowArena:
    """Pool of equally sized fp32 rows on one device, grown chunk-wise and reused FIFO.

    ``ghost=True`` mirrors the bookkeeping of another rank's arena (its rows are views of that
    rank's shared memory); every rank replays the same alloc/free sequence on every mirror, so
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
        row.in_use = True
        self.live += 1
        self.high_water = max(self.high_water, self.live)
        return row

    def free(self, row: Row) -> None:
        if row.index < 0:
            return
        if not row.in_use:
            # a second free would put the row on the free list twice and hand it to two handlers
            raise RuntimeError("arena row %d freed twice" % row.index)
        row.in_use = False
        self.live -= 1
        self._free.append(row)  # FIFO reuse: pending readers have usually long finished

    def nbytes(self) -> int:
        return self._n_rows * self.row_numel * 4


class SymmetricArenas:
    """Identical shared arenas on every rank (``p2p`` transport).  Creation is collective.

    Layout of each rank's allocation: ``capacity`` rows of ``row_numel`` floats followed by the
    per-row flags: ``ready`` (uint32 generation published by the owner) and ``done`` (reads
    acknowledged by other ranks).

    * CUDA: ``cudaMalloc`` + CUDA IPC; every peer's arena is mapped into this process, so a peer
      row is a device pointer whose loads travel over NVLink / NVSwitch.  ``done`` is one counter
      (readers use ``red.release.sys.add``).
    * CPU (gloo plumbing runs): POSIX shared memory; ``done`` has one slot per reader rank (a
      slot has a single writer, so plain stores suffice).
    """

    def __init__(self, device: torch.device, row_numel: int) -> None:
        from ..parallel import runtime as prt
        self.device = torch.device(device)
        self.row_numel = int(row_numel)
        self.world = prt.world()
        row_bytes = self.row_numel * 4
        cap = prt.arena_capacity()
        if cap is None:
            cap = int(max(64, min(8192, (256 << 20) // row_bytes)))
        self.capacity = cap               # rows per segment
        self.cuda = self.device.type == "cuda"
        self.flag_words = 2 if self.cuda else 1 + self.world
        self.n_segments = 0
        self._shm: List = []
        self.mirrors: List[RowArena] = []
        for r in range(self.world):
            mirror = RowArena(self.device, row_numel, rank=r, ghost=True)
            mirror._grow = self.grow  # type: ignore[assignment]
            self.mirrors.append(mirror)
        self.grow()

    def grow(self) -> None:
        """Add one segment of ``capacity`` rows (+ flags) on EVERY rank.  Collective -- and safe to call from the
        allocation path: all ranks replay the same alloc / free sequence on every mirror, so they all run out of rows
        of a mirror at the same point of the replicated bookkeeping.  Rows keep their addresses (earlier segments are
        never moved), so captured pointers, in-flight reads and the executor's slot tables stay valid."""
        import torch.distributed as dist
        from ..parallel import runtime as prt
        row_bytes = self.row_numel * 4
        flags_off = self.capacity * row_bytes
        total = flags_off + self.capacity * 4 * self.flag_words
        seg = self.n_segments
        first = seg * self.capacity
        if self.cuda:
            from ..ops.native import native
            nat = native()
            torch.cuda.set_device(self.device)
            base = nat.ipc_alloc(total)
            handles: List = [None] * self.world
            dist.all_gather_object(handles, nat.ipc_get_handle(base))
            bases = [base if r == prt.rank() else nat.ipc_open_handle(handles[r]) for r in range(self.world)]
            dist.barrier()
            dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
            for r, mirror in enumerate(self.mirrors):
                for i in range(self.capacity):
                    t = nat.tensor_from_ptr(bases[r] + i * row_bytes, [self.row_numel], dev_index, False)
                    row = Row(mirror, first + i, t)
                    row.flag_ready = bases[r] + flags_off + 8 * i
                    row.flag_done = bases[r] + flags_off + 8 * i + 4
                    mirror._free.append(row)
                mirror._n_rows += self.capacity
        else:
            import numpy as np
            from multiprocessing import shared_memory
            tag = prt.session_tag()
            names = ["gb200_%s_%d_%d_%d" % (tag, self.row_numel, seg, r) for r in range(self.world)]
            mine = shared_memory.SharedMemory(name=names[prt.rank()], create=True, size=total)
            _unlink_at_exit()
            np.frombuffer(mine.buf, dtype=np.uint8)[:] = 0
            dist.barrier()
            shms = [mine if r == prt.rank() else attach_shm(names[r]) for r in range(self.world)]
            dist.barrier()
            self._shm.append(shms)
            _SHM_KEEPALIVE.extend(shms)
            for r, mirror in enumerate(self.mirrors):
                buf = shms[r].buf
                data = torch.frombuffer(buf, dtype=torch.float32, count=self.capacity * self.row_numel)
                flags = np.frombuffer(buf, dtype=np.int32, offset=flags_off,
                                      count=self.capacity * self.flag_words)
                for i in range(self.capacity):
                    row = Row(mirror, first + i, data[i * self.row_numel:(i + 1) * self.row_numel])
                    row.flag_ready = (flags, i * self.flag_words)
                    row.flag_done = (flags, i * self.flag_words + 1)     # + reader rank
                    mirror._free.append(row)
                mirror._n_rows += self.capacity
        self.n_segments += 1

    def close(self) -> None:
        from ..parallel import runtime as prt
        if self.cuda:
            return
        for shms in self._shm:
            for r, shm in enumerate(shms):
                if r == prt.rank():
                    try:
                        shm.unlink()
                    except Exception:
                        pass
                try:
                    shm.close()
                except Exception:      # tensors created with torch.frombuffer may still reference the mapping
                    pass


_SHM_KEEPALIVE: List = []


_EXIT_HOOK = [False]


def _unlink_at_exit() -> None:
    """A script that never calls ``parallel.runtime.shutdown`` must not leave its segments in /dev/shm."""
    if not _EXIT_HOOK[0]:
        import atexit
        _EXIT_HOOK[0] = True
        atexit.register(lambda: _quiet(reset_arenas))


def _quiet(fn) -> None:
    try:
        fn()
    except Exception:
        pass


def attach_shm(name: str):
    """Map a peer's POSIX shared-memory segment.  Python (< 3.13) registers every attached segment with this process's
    resource tracker, which then unlinks it -- or warns about a "leak" -- at exit although the creator owns it."""
    from multiprocessing import resource_tracker, shared_memory
    shm = shared_memory.SharedMemory(name=name)
    try:
        resource_tracker.unregister(shm._name, "shared_memory")
    except Exception:
        pass
    return shm


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
    if prt.active() and prt.transport() == "p2p":
        skey = (device.type, device.index, int(row_numel))
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
    from . import bank as _bank
    _bank.close_all()
    for sym in _SYMMETRIC.values():
        sym.close()
    _ARENAS.clear()
    _SYMMETRIC.clear()
    _STREAMS.clear()


# --------------------------------------------------------------------------------------
# cross-rank protocol (replicated bookkeeping; see parallel/runtime.py)
# --------------------------------------------------------------------------------------
def read_sync(row: Row, reader_rank: int):
    """Account one read of ``row`` by ``reader_rank`` (called identically on EVERY rank) and return
    the handshake the reader must perform -- ``None`` for same-rank reads or on other ranks."""
    from ..ops import RowSync
    from ..parallel import runtime as prt
    if not prt.active() or row.rank == reader_rank:
        return None
    row.remote_reads += 1
    if reader_rank != prt.rank():
        return None
    done = row.flag_done
    if isinstance(done, tuple):           # CPU: one done slot per reader rank
        done = (done[0], done[1] + reader_rank)
    return RowSync(row.flag_ready, row.gen, done)


def publish(row: Row, mine: bool) -> None:
    """After a write to a shared row on the CURRENT stream: bump the generation everywhere and,
    on the owner, publish it to the readers of other ranks."""
    from ..parallel import runtime as prt
    if not prt.active():
        return
    row.gen += 1
    if not mine or not row.flag_ready:
        return
    if isinstance(row.flag_ready, tuple):
        arr, i = row.flag_ready
        arr[i] = row.gen
    else:
        from ..ops.native import native
        native().flag_signal(row.flag_ready, row.gen)


def wait_remote_readers(row: Row) -> None:
    """Owner only, before re-writing a shared row on the CURRENT stream: wait until every read by
    another rank has been acknowledged."""
    from ..parallel import runtime as prt
    if not prt.active() or row.remote_reads == 0 or not row.flag_done:
        return
    if row.remote_reads == getattr(row, "_acked", 0):
        return
    if isinstance(row.flag_done, tuple):
        import time
        arr, i = row.flag_done
        t0 = time.monotonic()
        while int(arr[i:i + prt.world()].sum()) < row.remote_reads:
            if time.monotonic() - t0 > 120:
                raise TimeoutError("remote readers never acknowledged row %d" % row.index)
            time.sleep(0)
    else:
        from ..ops.native import native
        native().flag_wait(row.flag_done, row.remote_reads)
    row._acked = row.remote_reads


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
