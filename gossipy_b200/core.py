"""Core vocabulary: protocol enums, messages, delay models, topologies, mixing weights.

Behavioural reference: ``gossipy/core.py`` (cited per item).  Deliberate deviations (intended
behaviour instead of a reference bug) are marked ``FIX(Bn)`` with the SURVEY appendix-B id and
can be switched back with ``GlobalSettings().reference_compat = True`` where that is cheap.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from enum import Enum
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from . import GlobalSettings, Sizeable, atoms_of

try:  # scipy is optional at import time
    from scipy.sparse import issparse as _issparse
except Exception:  # pragma: no cover
    def _issparse(_x):
        return False

__all__ = ["CreateModelMode", "AntiEntropyProtocol", "MessageType", "Message", "Delay",
           "ConstantDelay", "UniformDelay", "LinearDelay", "P2PNetwork", "StaticP2PNetwork",
           "MixingMatrix", "UniformMixing", "MetropolisHastingsMixing"]


class CreateModelMode(Enum):
    """How a receiver combines a received model with its own (ref ``core.py:31-44``)."""
    UPDATE = 1         #: train the received model locally and adopt it
    MERGE_UPDATE = 2   #: merge into the local model, then train
    UPDATE_MERGE = 3   #: train both, then merge
    PASS = 4           #: adopt the received model untouched


class AntiEntropyProtocol(Enum):
    """Gossip exchange pattern (ref ``core.py:47-58``)."""
    PUSH = 1
    PULL = 2
    PUSH_PULL = 3


class MessageType(Enum):
    """Kind of a message on the wire (ref ``core.py:61-75``)."""
    PUSH = 1
    PULL = 2
    REPLY = 3
    PUSH_PULL = 4


_CARRIES_MODEL = (MessageType.PUSH, MessageType.REPLY, MessageType.PUSH_PULL)
_WANTS_REPLY = (MessageType.PULL, MessageType.PUSH_PULL)


class Message(Sizeable):
    """A message between two nodes (ref ``core.py:78-152``).

    ``value`` is a tuple whose first element is usually the :class:`~gossipy_b200.CacheKey` of
    the model snapshot; the payload itself never moves until the receiver's merge kernel pulls
    it (over NVLink when sender and receiver live on different GPUs).
    """

    __slots__ = ("timestamp", "sender", "receiver", "type", "value")

    def __init__(self, timestamp: int, sender: int, receiver: int, type: MessageType,
                 value: Optional[Tuple[Any, ...]]) -> None:
        self.timestamp = timestamp
        self.sender = sender
        self.receiver = receiver
        self.type = type
        self.value = value

    def carries_model(self) -> bool:
        return self.type in _CARRIES_MODEL

    def wants_reply(self) -> bool:
        return self.type in _WANTS_REPLY

    def get_size(self) -> int:
        """Size in atoms: ``None`` -> 1, tuple -> sum of element sizes (min 1)."""
        v = self.value
        if v is None:
            return 1
        if isinstance(v, (tuple, list)):
            return max(sum(atoms_of(el) for el in v), 1)
        return atoms_of(v)

    def __repr__(self) -> str:
        body = "ACK" if self.value is None else str(self.value)
        return "T%d [%d -> %d] {%s}: %s" % (self.timestamp, self.sender, self.receiver,
                                            self.type.name, body)


# --------------------------------------------------------------------------------------
# delays (ref core.py:155-307)
# --------------------------------------------------------------------------------------
class Delay(ABC):
    """Maps a message to its delivery delay in simulation ticks."""

    @abstractmethod
    def get(self, msg: Message) -> int:
        ...

    def __repr__(self) -> str:
        return str(self)


class ConstantDelay(Delay):
    def __init__(self, delay: int = 0) -> None:
        assert delay >= 0, "Delay must be non-negative!"
        self._delay = int(delay)

    def get(self, msg: Message) -> int:
        return self._delay

    def __str__(self) -> str:
        return "ConstantDelay(%d)" % self._delay


class UniformDelay(Delay):
    """Uniform integer delay in ``[min_delay, max_delay]`` (both inclusive)."""

    def __init__(self, min_delay: int, max_delay: int) -> None:
        assert 0 <= min_delay <= max_delay, \
            "The minimum delay must be non-negative and less than or equal to the maximum delay!"
        self._min_delay, self._max_delay = int(min_delay), int(max_delay)

    def get(self, msg: Message) -> int:
        return int(np.random.randint(self._min_delay, self._max_delay + 1))

    def __str__(self) -> str:
        return "UniformDelay(%d, %d)" % (self._min_delay, self._max_delay)


class LinearDelay(Delay):
    """Delay proportional to the message size: ``int(timexunit * size) + overhead``."""

    def __init__(self, timexunit: float, overhead: int) -> None:
        assert timexunit >= 0 and overhead >= 0
        self._timexunit, self._overhead = float(timexunit), int(overhead)

    def get(self, msg: Message) -> int:
        return int(self._timexunit * msg.get_size()) + self._overhead

    def __str__(self) -> str:
        return "LinearDelay(time_x_unit=%g, overhead=%d)" % (self._timexunit, self._overhead)


# --------------------------------------------------------------------------------------
# topology (ref core.py:311-389)
# --------------------------------------------------------------------------------------
class P2PNetwork(ABC):
    """Peer lists of a (possibly implicit clique) overlay network.

    ``topology`` may be a dense 0/1 ``ndarray``, a scipy sparse matrix, a ``networkx`` graph or
    ``None`` (clique).  Neighbour lists are also exported in CSR form (:meth:`as_csr`) for the
    C++ scheduler.
    """

    def __init__(self, num_nodes: int, topology: Any = None) -> None:
        if topology is None:
            assert num_nodes > 0, "The number of nodes must be positive!"
        elif hasattr(topology, "shape"):
            # FIX(B2): the reference's check is a no-op expression (core.py:328-329)
            assert num_nodes == topology.shape[0], \
                "The number of nodes must match the number of rows of the topology!"
        self._num_nodes = int(num_nodes)
        self._clique = topology is None     # implicit clique: the C++ scheduler indexes it without a peer table
        self._topology: Dict[int, List[int]] = {}
        if topology is None:
            for i in range(num_nodes):
                self._topology[i] = [j for j in range(num_nodes) if j != i]
        elif isinstance(topology, np.ndarray):
            for i in range(num_nodes):
                self._topology[i] = [int(j) for j in np.flatnonzero(topology[i] > 0)]
        elif _issparse(topology):
            csr = topology.tocsr()
            for i in range(num_nodes):
                self._topology[i] = [int(j) for j in csr.indices[csr.indptr[i]:csr.indptr[i + 1]]]
        elif hasattr(topology, "neighbors"):  # networkx graph
            for i in range(num_nodes):
                self._topology[i] = sorted(int(j) for j in topology.neighbors(i))
        else:
            raise TypeError("Unsupported topology type %s" % type(topology))

    def size(self, node: Optional[int] = None) -> int:
        """#nodes when called without argument, otherwise the degree of ``node``.

        FIX(B1): the reference tests ``if node:`` so node 0 reports ``num_nodes``
        (``core.py:346-349``); with ``reference_compat`` that behaviour is reproduced.
        """
        if node is None or (node == 0 and GlobalSettings().reference_compat):
            return self._num_nodes
        peers = self._topology[node]
        return len(peers) if peers else self._num_nodes - 1

    @abstractmethod
    def get_peers(self, node_id: int) -> List[int]:
        ...

    def as_csr(self) -> Tuple[np.ndarray, np.ndarray]:
        indptr = np.zeros(self._num_nodes + 1, dtype=np.int64)
        for i in range(self._num_nodes):
            indptr[i + 1] = indptr[i] + len(self._topology[i])
        indices = np.fromiter((j for i in range(self._num_nodes) for j in self._topology[i]),
                              dtype=np.int64, count=int(indptr[-1]))
        return indptr, indices

    def __str__(self) -> str:
        return "%s(n=%d)" % (self.__class__.__name__, self._num_nodes)


class StaticP2PNetwork(P2PNetwork):
    """A network whose peer lists never change (ref ``core.py:364-389``)."""

    def get_peers(self, node_id: int) -> List[int]:
        assert 0 <= node_id < self._num_nodes
        return self._topology[node_id]


# --------------------------------------------------------------------------------------
# mixing weights for neighbourhood averaging (ref core.py:392-453)
# --------------------------------------------------------------------------------------
class MixingMatrix:
    """Row ``i`` = ``[w_self, w_peer_0, w_peer_1, ...]`` in ``get_peers(i)`` order."""

    def __init__(self, p2p_net: P2PNetwork) -> None:
        self.p2p_net = p2p_net

    def get(self, node_id: int) -> np.ndarray:
        raise NotImplementedError

    def __getitem__(self, node_id: int) -> np.ndarray:
        return self.get(node_id)

    def __str__(self) -> str:
        return "%s(%s)" % (self.__class__.__name__, self.p2p_net)


class UniformMixing(MixingMatrix):
    def get(self, node_id: int) -> np.ndarray:
        k = self.p2p_net.size(node_id) + 1
        return np.full(k, 1.0 / k)


class MetropolisHastingsMixing(MixingMatrix):
    """Metropolis-Hastings weights.

    By default mimics the reference (``core.py:451-453``: self weight ``1/deg``, rows need not
    sum to one -- SURVEY B3) because that defines the published algorithm's numbers; pass
    ``normalized=True`` for the textbook ``w_ii = 1 - sum_j w_ij``.
    """

    def __init__(self, p2p_net: P2PNetwork, normalized: bool = False) -> None:
        super().__init__(p2p_net)
        self.normalized = normalized

    def get(self, node_id: int) -> np.ndarray:
        deg = self.p2p_net.size(node_id)
        peers = self.p2p_net.get_peers(node_id)
        w_peers = [1.0 / (min(self.p2p_net.size(k), deg) + 1) for k in peers]
        w_self = (1.0 - sum(w_peers)) if self.normalized else 1.0 / deg
        return np.array([w_self] + w_peers)
