"""One-process-per-GPU runtime.

Execution model (world_size > 1): every rank runs the SAME cheap host-side simulation (identical
seeds -> identical event schedule, ages, token balances, arena allocations) but executes the device
work only of the gossip nodes it owns (block placement ``node -> rank``).  A model "sent" to a node
on another GPU is never copied by the sender: its snapshot row stays in the sender's HBM arena and
the receiver's fused merge kernel *pulls* it.

Transport (``p2p``) for a row that lives on another rank: every rank allocates an identical
("symmetric") arena and maps all peers' arenas -- ``cudaMalloc`` + CUDA IPC on GPUs (handles exchanged
once with ``torch.distributed``; a peer row is then a device pointer and the merge / training kernels'
loads travel over NVLink/NVSwitch), POSIX shared memory on CPU (gloo plumbing runs).  Cross-rank
ordering uses two words per row inside the arenas: the producer publishes ``ready = generation``
(``st.release.sys`` after the snapshot kernel), the consumer's kernel spins on it
(``ld.acquire.sys``) before its first peer load and afterwards bumps the producer's ``done``
counter so the row can be recycled -- no host synchronisation, no NCCL on the data path.

Because the schedule is replicated, matching sends/receives, flag generations and arena row indices
are derived independently and identically on every rank -- there is no control traffic at all.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional

import torch

from .. import GlobalSettings

_state: Dict[str, Any] = {"rank": 0, "world": 1, "n_nodes": None, "transport": "none",
                          "placement": None, "arena_capacity": None, "tag": "0"}


def init(rank: Optional[int] = None, world: Optional[int] = None,
         transport: Optional[str] = None, arena_capacity: Optional[int] = None) -> None:
    """Activate multi-rank execution.  ``torch.distributed`` must already be initialised."""
    import torch.distributed as dist
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    _state["rank"], _state["world"] = int(rank), int(world)
    GlobalSettings().set_topology(int(rank), int(world))
    if transport is None:
        transport = os.environ.get("GOSSIPY_B200_TRANSPORT", "")
    if not transport:
        transport = "p2p"
    transport = {"nccl-baseline": "nccl"}.get(transport, transport)
    if transport not in ("p2p", "nccl", "loopback"):
        raise ValueError("transport must be p2p, nccl (nccl-baseline) or loopback, not %r" % (transport,))
    if transport == "loopback":
        # no inter-rank transport: this process hosts ALL nodes on its own device (every rank of a job then runs the
        # whole simulation by itself) -- the single-device reference point for the other two transports
        rank, world = 0, 1
        _state["rank"], _state["world"] = 0, 1
        GlobalSettings().set_topology(0, 1)
    _state["transport"] = transport if world > 1 else ("loopback" if transport == "loopback" else "none")
    if arena_capacity is None and os.environ.get("GOSSIPY_B200_ARENA_ROWS"):
        arena_capacity = int(os.environ["GOSSIPY_B200_ARENA_ROWS"])     # rows per segment of the symmetric arenas
    _state["arena_capacity"] = arena_capacity
    if world > 1:
        assert dist.is_initialized(), "initialise torch.distributed before parallel.runtime.init"
        _state["inits"] = _state.get("inits", 0) + 1
        tag = ["%s_%d_%d" % (os.environ.get("MASTER_PORT", "0"), os.getpid(), _state["inits"])]
        dist.broadcast_object_list(tag, src=0)
        _state["tag"] = tag[0]


def shutdown() -> None:
    from ..engine import arena
    arena.reset_arenas()
    _state.update(rank=0, world=1, n_nodes=None, transport="none", placement=None, arena_capacity=None)
    GlobalSettings().set_topology(0, 1)


def arena_capacity() -> Optional[int]:
    return _state["arena_capacity"]


def session_tag() -> str:
    return _state["tag"]


def active() -> bool:
    return _state["world"] > 1


def rank() -> int:
    return _state["rank"]


def world() -> int:
    return _state["world"]


def transport() -> str:
    return _state["transport"]


class Placement:
    """Which rank (GPU) hosts which gossip node.

    ``Placement.block(n, world)`` keeps neighbouring ids together (the default: node ``i`` on rank ``i * world // n``),
    ``Placement.round_robin(n, world)`` deals them out, ``Placement.explicit([...])`` takes one rank per node and
    ``Placement.by_load(weights, world)`` balances per-node costs (e.g. shard sizes) greedily, heaviest first.
    Every rank must install the same placement before ``init_nodes``::

        runtime.set_num_nodes(n, Placement.by_load([len(disp[i][0][0]) for i in range(n)], world))
    """

    def __init__(self, ranks: List[int], world: Optional[int] = None) -> None:
        self.ranks = [int(r) for r in ranks]
        self.world = int(world) if world is not None else (max(self.ranks) + 1 if self.ranks else 1)
        if any(r < 0 or r >= self.world for r in self.ranks):
            raise ValueError("placement names a rank outside 0..%d" % (self.world - 1))

    @classmethod
    def block(cls, n: int, world: int) -> "Placement":
        return cls([min(world - 1, i * world // n) for i in range(n)], world)

    @classmethod
    def round_robin(cls, n: int, world: int) -> "Placement":
        return cls([i % world for i in range(n)], world)

    @classmethod
    def explicit(cls, ranks: List[int], world: Optional[int] = None) -> "Placement":
        return cls(ranks, world)

    @classmethod
    def by_load(cls, weights: List[float], world: int) -> "Placement":
        load = [0.0] * world
        ranks = [0] * len(weights)
        for i in sorted(range(len(weights)), key=lambda k: (-float(weights[k]), k)):
            r = min(range(world), key=lambda q: (load[q], q))
            ranks[i] = r
            load[r] += float(weights[i])
        return cls(ranks, world)

    def rank_of(self, node_id: int) -> int:
        return self.ranks[node_id]

    def nodes_of(self, rank: int) -> List[int]:
        return [i for i, r in enumerate(self.ranks) if r == rank]

    def __len__(self) -> int:
        return len(self.ranks)

    def __eq__(self, other: Any) -> bool:
        return isinstance(other, Placement) and self.ranks == other.ranks and self.world == other.world

    def __repr__(self) -> str:
        return "Placement(%s, world=%d)" % (self.ranks, self.world)


def set_num_nodes(n: int, placement: Any = None) -> None:
    """Fix the node -> rank placement: a :class:`Placement`, one rank per node, or ``None`` = keep an installed
    placement of ``n`` nodes, else block placement (what ``init_nodes`` calls)."""
    if isinstance(placement, Placement):
        placement = placement.ranks
    if placement is None:
        prev = _state["placement"]
        placement = prev if prev is not None and len(prev) == int(n) else None
    else:
        placement = [int(r) for r in placement]
        if len(placement) != int(n):
            raise ValueError("placement has %d entries for %d nodes" % (len(placement), int(n)))
        if any(r < 0 or r >= max(1, _state["world"]) for r in placement):
            raise ValueError("placement names a rank outside the job")
    _state["n_nodes"] = int(n)
    _state["placement"] = placement


def placement() -> Optional[Placement]:
    """The installed placement (``None`` when single-process or before ``set_num_nodes``)."""
    n, w = _state["n_nodes"], _state["world"]
    if w == 1 or not n:
        return None
    return Placement([rank_of(i) for i in range(n)], w)


def rank_of(node_id: int) -> int:
    """Rank that owns gossip node ``node_id`` (0 when single-process)."""
    w = _state["world"]
    if w == 1 or node_id is None or node_id < 0:
        return 0
    pl = _state["placement"]
    if pl is not None:
        return pl[node_id]
    n = _state["n_nodes"]
    if not n:
        return node_id % w
    return min(w - 1, node_id * w // n)


def is_mine(node_id: int) -> bool:
    return _state["world"] == 1 or rank_of(node_id) == _state["rank"]


# --------------------------------------------------------------------------------------
# collectives on tiny host-visible results
# --------------------------------------------------------------------------------------
METRIC_KEYS = ("accuracy", "precision", "recall", "f1_score", "auc", "rmse", "nmi")


def share_metrics(local: List[Optional[Dict[str, float]]]) -> List[Dict[str, float]]:
    """``local[i]`` is the metric dict of evaluation ``i`` on the rank that computed it and ``None``
    elsewhere; returns the complete list on every rank (one small all-reduce)."""
    if not active():
        return [d for d in local]  # type: ignore[misc]
    import torch.distributed as dist
    k = len(METRIC_KEYS)
    buf = torch.zeros(len(local), 2 * k, dtype=torch.float64)
    for i, d in enumerate(local):
        if d is not None:
            for j, name in enumerate(METRIC_KEYS):
                if name in d:
                    buf[i, j] = float(d[name])
                    buf[i, k + j] = 1.0
    dev = GlobalSettings().get_device()
    if dist.get_backend() == "nccl":
        buf = buf.to(dev)
    dist.all_reduce(buf)
    buf = buf.cpu()
    out = []
    for i in range(len(local)):
        out.append({name: float(buf[i, j]) for j, name in enumerate(METRIC_KEYS) if buf[i, k + j] > 0})
    return out


def share_ints(values: Optional[List[int]], src_rank: int, n: int) -> List[int]:
    """Broadcast ``n`` small integers decided on ``src_rank`` (a data-dependent choice such as PENS'
    top-m senders) so that the replicated bookkeeping stays identical everywhere.  Every rank calls it
    at the same point of the (replicated) event sequence; ``values`` is ignored off ``src_rank``."""
    if not active():
        return list(values or [])
    import torch.distributed as dist
    buf = torch.zeros(n, dtype=torch.int64)
    if rank() == src_rank:
        assert values is not None and len(values) == n
        buf.copy_(torch.as_tensor(list(values), dtype=torch.int64))
    if dist.get_backend() == "nccl":
        buf = buf.to(GlobalSettings().get_device())
    dist.broadcast(buf, src=src_rank)
    return [int(v) for v in buf.cpu().tolist()]


def barrier() -> None:
    if active():
        import torch.distributed as dist
        dist.barrier()
