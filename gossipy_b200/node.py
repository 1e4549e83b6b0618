"""Gossip nodes: who sends what to whom, and what a receiver does with it.

Behavioural reference: ``gossipy/node.py`` (cited per class).  Nodes are thin protocol objects;
all numerical work is delegated to their model handler, whose device work runs on the node's
own CUDA stream (``engine.arena.stream_for``).  The shared send/receive skeleton lives in
:class:`GossipNode`; subclasses only override the small hooks ``_payload_extras`` (what rides
along with the model key) and ``_consume`` (what to do with a received model).
"""
from __future__ import annotations

import random
from typing import Any, Dict, Iterable, List, Optional, Tuple, Union

import numpy as np
from numpy import ndarray
from torch import Tensor

from . import CACHE, LOG, GlobalSettings
from .core import AntiEntropyProtocol, CreateModelMode, Message, MessageType, P2PNetwork
from .data import DataDispatcher
from .model.handler import ModelHandler, PartitionedTMH, SamplingTMH, WeightedTMH
from .utils import choice_not_n

__all__ = ["GossipNode", "PassThroughNode", "CacheNeighNode", "SamplingBasedNode",
           "PartitioningBasedNode", "PENSNode", "All2AllGossipNode"]

NodeData = Union[Tuple[Tensor, Optional[Tensor]], Tuple[ndarray, Optional[ndarray]]]

_SEND_TYPE = {AntiEntropyProtocol.PUSH: MessageType.PUSH,
              AntiEntropyProtocol.PULL: MessageType.PULL,
              AntiEntropyProtocol.PUSH_PULL: MessageType.PUSH_PULL}


class GossipNode:
    """A vanilla gossip node (ref ``node.py:34-286``).

    ``sync=True``: fires once per round at a random offset ``delta`` in ``[0, round_len)``;
    ``sync=False``: fires every ``delta ~ N(round_len, round_len/10)`` ticks (guarded to be
    at least 1) and therefore also at ``t=0``.
    """

    def __init__(self, idx: int, data: NodeData, round_len: int, model_handler: ModelHandler,
                 p2p_net: P2PNetwork, sync: bool = True) -> None:
        self.idx = idx
        self.data = data
        self.round_len = round_len
        self.model_handler = model_handler
        self.model_handler.owner = idx
        self.sync = sync
        if sync:
            self.delta = int(np.random.randint(0, round_len))
        else:
            self.delta = max(1, int(np.random.normal(round_len, round_len / 10)))
        self.p2p_net = p2p_net

    # -- life cycle -----------------------------------------------------------------------
    def init_model(self, local_train: bool = True, *args, **kwargs) -> None:
        """Initialise the model and (by default) run one local update."""
        self.model_handler.owner = self.idx
        self.model_handler.init()
        if local_train:
            self.model_handler._update(self.data[0])

    def get_peer(self) -> Optional[int]:
        """A uniformly random neighbour, ``None`` (with a warning) when there is none."""
        peers = self.p2p_net.get_peers(self.idx)
        if not peers:
            LOG.warning("Node %d has no peers." % self.idx)
            return None
        return random.choice(peers)

    def timed_out(self, t: int) -> bool:
        if self.sync:
            return (t % self.round_len) == self.delta
        return (t % self.delta) == 0

    # -- sending -----------------------------------------------------------------------------
    def _payload_extras(self) -> Tuple[Any, ...]:
        """Scalars that travel with the model key (degree, sample size, partition id, ...)."""
        return ()

    def _before_snapshot(self) -> None:
        """Hook run right before the model is snapshotted for sending."""

    def _model_message(self, t: int, peer: int, mtype: MessageType) -> Message:
        self._before_snapshot()
        key = self.model_handler.caching(self.idx)
        return Message(t, self.idx, peer, mtype, (key,) + tuple(self._payload_extras()))

    def send(self, t: int, peer: int, protocol: AntiEntropyProtocol) -> Message:
        """Build the message for ``peer``; PUSH / PUSH_PULL snapshot the model *now*."""
        mtype = _SEND_TYPE.get(protocol)
        if mtype is None:
            raise ValueError("Unknown protocol %s." % protocol)
        if mtype == MessageType.PULL:
            return Message(t, self.idx, peer, MessageType.PULL, None)
        return self._model_message(t, peer, mtype)

    # -- receiving -----------------------------------------------------------------------------
    def _consume(self, msg: Message, recv_model: ModelHandler, extras: Tuple[Any, ...]) -> None:
        """Use a received model: default = the handler's merge/update rule."""
        self.model_handler(recv_model, self.data[0])
        _release(recv_model)

    def receive(self, t: int, msg: Message) -> Optional[Message]:
        """Process ``msg``; returns a REPLY for PULL / PUSH_PULL requests.

        The reply snapshot is taken *after* the local merge/update (ref ``node.py:171-204``).
        """
        if msg.carries_model():
            key = msg.value[0]
            self._on_model(msg, key, tuple(msg.value[1:]))
        if msg.wants_reply():
            return self._model_message(t, msg.sender, MessageType.REPLY)
        return None

    def _on_model(self, msg: Message, key: Any, extras: Tuple[Any, ...]) -> None:
        recv_model = CACHE.pop(key)
        self._consume(msg, recv_model, extras)

    # -- evaluation ------------------------------------------------------------------------------
    def evaluate(self, ext_data: Optional[Any] = None) -> Dict[str, float]:
        return self.model_handler.evaluate(self.data[1] if ext_data is None else ext_data)

    def evaluate_async(self, ext_data: Optional[Any] = None):
        return self.model_handler.evaluate_async(self.data[1] if ext_data is None else ext_data)

    def has_test(self) -> bool:
        if isinstance(self.data, tuple):
            return self.data[1] is not None
        return True

    def __repr__(self) -> str:
        return str(self)

    def __str__(self) -> str:
        return "%s #%d (Δ=%d)" % (self.__class__.__name__, self.idx, self.delta)

    @classmethod
    def generate(cls, data_dispatcher: DataDispatcher, p2p_net: P2PNetwork,
                 model_proto: ModelHandler, round_len: int, sync: bool,
                 **kwargs) -> Dict[int, "GossipNode"]:
        """One node per network vertex, each with a copy of ``model_proto`` and its data shard."""
        return {idx: cls(idx=idx, data=data_dispatcher[idx], round_len=round_len,
                         model_handler=model_proto.copy(), p2p_net=p2p_net, sync=sync, **kwargs)
                for idx in range(p2p_net.size())}


def _release(snapshot: Any) -> None:
    rel = getattr(snapshot, "release", None)
    if callable(rel):
        rel()


class PassThroughNode(GossipNode):
    """Giaretta & Girdzijauskas 2019: degree-aware pass-through (ref ``node.py:289-392``).

    The sender's degree rides along; the receiver merges with probability
    ``min(1, deg_sender/deg_self)`` and otherwise adopts the model untouched (mode ``PASS``).
    """

    def __init__(self, idx: int, data: NodeData, round_len: int, model_handler: ModelHandler,
                 p2p_net: P2PNetwork, sync: bool = True) -> None:
        super().__init__(idx, data, round_len, model_handler, p2p_net, sync)
        self.n_neighs = p2p_net.size(idx)  # true degree also for node 0 (FIX B1)

    def _payload_extras(self) -> Tuple[Any, ...]:
        return (self.n_neighs,)

    def _accepts(self, deg: int) -> bool:
        """Merge with probability ``min(1, deg_sender / deg_self)``.  Under the native engine the draw is keyed by
        (node, number of draws so far) and compared in integer arithmetic, so that the C++ executor and the banked engine
        reproduce it (``u < 2**43``; accept iff ``u * deg_self < deg_sender * 2**43``)."""
        if getattr(self, "_keyed_draws", False):
            from .engine import rng as _rng
            k = int(getattr(self, "_pt_draws", 0))
            self._pt_draws = k + 1
            u = _rng.derive(0x9A55, self.idx, k) >> 20
            return u * int(self.n_neighs) < (int(deg) << 43)
        return np.random.rand() < min(1, deg / self.n_neighs)

    def _consume(self, msg: Message, recv_model: ModelHandler, extras: Tuple[Any, ...]) -> None:
        deg = extras[0]
        if self._accepts(deg):
            self.model_handler(recv_model, self.data[0])
        else:
            prev = self.model_handler.mode
            self.model_handler.mode = CreateModelMode.PASS
            try:
                self.model_handler(recv_model, self.data[0])
            finally:
                self.model_handler.mode = prev
        _release(recv_model)


class CacheNeighNode(GossipNode):
    """Giaretta 2019: one cache slot per neighbour (ref ``node.py:395-496``).

    Received models are only stored (newest per sender); at send time one cached model is
    consumed (merge-update) before the snapshot.  FIX(B11): the reference's
    ``random.choice(set(...))`` raises on Python >= 3.11.
    """

    def __init__(self, idx: int, data: NodeData, round_len: int, model_handler: ModelHandler,
                 p2p_net: P2PNetwork, sync: bool = True) -> None:
        super().__init__(idx, data, round_len, model_handler, p2p_net, sync)
        self.local_cache: Dict[int, Any] = {}

    def _before_snapshot(self) -> None:
        pass

    def send(self, t: int, peer: int, protocol: AntiEntropyProtocol) -> Message:
        if protocol in (AntiEntropyProtocol.PUSH, AntiEntropyProtocol.PUSH_PULL) and self.local_cache:
            keys = sorted(self.local_cache.keys())
            if getattr(self, "_keyed_draws", False):
                # native engine: keyed by (node, number of choices so far) so that the banked engine reproduces it
                from .engine import rng as _rng
                n_draws = int(getattr(self, "_cn_draws", 0))
                self._cn_draws = n_draws + 1
                k = keys[_rng.derive(0x9A59, self.idx, n_draws) % len(keys)]
            else:
                k = random.choice(keys)
            cached = CACHE.pop(self.local_cache.pop(k))
            self.model_handler(cached, self.data[0])
            _release(cached)
        return super().send(t, peer, protocol)

    def _on_model(self, msg: Message, key: Any, extras: Tuple[Any, ...]) -> None:
        old = self.local_cache.get(msg.sender)
        if old is not None:
            CACHE.drop(old)
        self.local_cache[msg.sender] = key


class SamplingBasedNode(GossipNode):
    """Hegedűs 2021 sub-sampling: the *receiver* draws the coordinate sample
    (ref ``node.py:499-562``).  Size accounting stays "full model + 1" like the reference; the
    bytes that actually cross NVLink are only the sampled coordinates."""

    def _payload_extras(self) -> Tuple[Any, ...]:
        return (self.model_handler.sample_size,)

    def _consume(self, msg: Message, recv_model: ModelHandler, extras: Tuple[Any, ...]) -> None:
        sample_size = extras[0]
        handler: SamplingTMH = self.model_handler
        if GlobalSettings().reference_compat:      # the reference's own draw on the NumPy stream (differential tests)
            from .model.sampling import TorchModelSampling
            sample = TorchModelSampling.sample_reference(sample_size, handler._proto)
        elif sample_size == handler.sample_size:
            sample = handler.draw_sample()
        else:
            from .model.sampling import TorchModelSampling
            sample = TorchModelSampling.sample_flat(sample_size, handler.layout.n_params,
                                                    device=handler.device)
        handler(recv_model, self.data[0], sample)
        _release(recv_model)


class PartitioningBasedNode(GossipNode):
    """Hegedűs 2021 partitioned models: the *sender* picks the partition id
    (ref ``node.py:566-659``); replies draw a fresh one."""

    def _payload_extras(self) -> Tuple[Any, ...]:
        n_parts = self.model_handler.tm_partition.n_parts
        if getattr(self, "_keyed_draws", False):
            # native engine: a counter-based draw (node id, number of model messages sent so far) that the C++ executor
            # reproduces (csrc/exec/executor.cpp::snapshot) -- the host NumPy stream is not consumed
            from .engine import rng as _rng
            k = int(getattr(self, "_model_msgs", 0))
            self._model_msgs = k + 1
            return (int(_rng.derive(0x9A57, self.idx, k) % n_parts),)
        return (int(np.random.randint(0, n_parts)),)

    def _model_message(self, t: int, peer: int, mtype: MessageType) -> Message:
        extras = self._payload_extras()  # the reference draws the pid before snapshotting
        key = self.model_handler.caching(self.idx)
        return Message(t, self.idx, peer, mtype, (key,) + extras)

    def _consume(self, msg: Message, recv_model: ModelHandler, extras: Tuple[Any, ...]) -> None:
        self.model_handler(recv_model, self.data[0], extras[0])
        _release(recv_model)


class PENSNode(GossipNode):
    """Onoszko 2021 performance-based neighbour selection (ref ``node.py:663-785``).

    Step 1 (first ``step1_rounds`` rounds): received models are scored on the local *training*
    data; once ``n_sampled`` distinct senders are cached the ``m_top`` best are merged (k-way) and
    their senders credited.  Step 2: gossip only with peers selected more often than chance.
    """

    def __init__(self, idx: int, data: NodeData, round_len: int, model_handler: ModelHandler,
                 p2p_net: P2PNetwork, n_sampled: int = 10, m_top: int = 2,
                 step1_rounds: int = 200, sync: bool = True) -> None:
        super().__init__(idx, data, round_len, model_handler, p2p_net, sync)
        assert self.model_handler.mode == CreateModelMode.MERGE_UPDATE, \
            "PENSNode can only be used with MERGE_UPDATE mode."
        self.cache: Dict[int, Tuple[Any, float]] = {}
        self.n_sampled = n_sampled
        self.m_top = m_top
        known = p2p_net.get_peers(self.idx)
        if not known:
            known = [i for i in range(p2p_net.size()) if i != idx]
        self.neigh_counter = {i: 0 for i in known}
        self.selected = {i: 0 for i in known}
        self.step1_rounds = step1_rounds
        self.step = 1
        self.best_nodes: Optional[List[int]] = None

    def __getstate__(self) -> Dict[str, Any]:
        st = dict(self.__dict__)       # checkpoints carry the scores, not the device-side pending reads
        st["cache"] = {s: (k, v.result() if hasattr(v, "result") else v) for s, (k, v) in self.cache.items()}
        return st

    def _select_neighbors(self) -> None:
        thr = self.m_top / self.n_sampled
        self.best_nodes = [i for i, cnt in self.neigh_counter.items() if cnt > self.selected[i] * thr]

    def timed_out(self, t: int) -> bool:
        if self.step == 1 and (t // self.round_len) >= self.step1_rounds:
            self.step = 2
            self._select_neighbors()
        return super().timed_out(t)

    def get_peer(self) -> Optional[int]:
        if self.step == 1 or not self.best_nodes:
            peer = super().get_peer()
            if peer is not None and self.step == 1:
                self.selected[peer] += 1
            return peer
        return random.choice(self.best_nodes)

    def send(self, t: int, peer: int, protocol: AntiEntropyProtocol) -> Message:
        if protocol != AntiEntropyProtocol.PUSH:
            LOG.warning("PENSNode only supports PUSH protocol.")
        return self._model_message(t, peer, MessageType.PUSH)

    def receive(self, t: int, msg: Message) -> Optional[Message]:
        if msg.type != MessageType.PUSH:
            LOG.warning("PENSNode only supports PUSH protocol.")
        key = msg.value[0]
        from .parallel import runtime as _prt
        if _prt.active() and self.step == 1:
            return self._receive_step1_multirank(msg, _prt)
        if self.step != 1:
            recv = CACHE.pop(key)
            self.model_handler(recv, self.data[0])
            _release(recv)
            return None
        # the candidate is scored on the device right away (its kernels queue behind the sender's snapshot), but the
        # score is only READ when the selection needs it: one host synchronisation per n_sampled models instead of one
        # per message (the reference evaluates synchronously in every receive, node.py:764-785)
        cand = CACHE[key]
        score = cand.evaluate_async(self.data[0]) if hasattr(cand, "evaluate_async") else cand.evaluate(self.data[0])
        stale = self.cache.get(msg.sender)
        if stale is not None:
            CACHE.drop(stale[0])
        self.cache[msg.sender] = (key, score)  # newest model per sender
        if len(self.cache) >= self.n_sampled:
            self.cache = {s: (k, -float((v.result() if hasattr(v, "result") else v)["accuracy"]))
                          for s, (k, v) in self.cache.items()}
            top = sorted(self.cache, key=lambda s: self.cache[s][1])[:self.m_top]
            models = [CACHE.pop(self.cache[s][0]) for s in top]
            self.model_handler(models, self.data[0])
            for m in models:
                _release(m)
            for s, (k, _) in self.cache.items():
                if s not in top:
                    CACHE.drop(k)
            self.cache = {}
            for s in top:
                self.neigh_counter[s] += 1
        return None


    def _receive_step1_multirank(self, msg: Message, _prt: Any) -> None:
        """Step 1 with the nodes spread over several ranks.

        The received snapshot (possibly in another GPU's arena) is pulled ONCE into a scratch row of
        this node -- over NVLink, on the node's stream, with the usual ready/done handshake -- and
        scored there; the later k-way merge reads the local scratch rows.  Which senders make the
        top-m is only known on the owner (it depends on device results): the owner broadcasts the m
        sender ids, so ages, selection counters and the arena bookkeeping stay identical on every
        rank (the peer choice of step 2 is part of the replicated schedule)."""
        snap = CACHE.pop(msg.value[0])
        local = self.model_handler._scratch_copy(snap)
        _release(snap)
        pending = local.evaluate_async(self.data[0])     # read at selection time only (see receive)
        stale = self.cache.get(msg.sender)
        if stale is not None:
            stale[0].release()
        self.cache[msg.sender] = (local, pending)    # newest model per sender
        if len(self.cache) < self.n_sampled:
            return None
        owner = _prt.rank_of(self.idx)
        top = None
        if _prt.rank() == owner:
            def neg_acc(s):
                res = self.cache[s][1]
                res = res.result() if hasattr(res, "result") else res
                return -float(res["accuracy"]) if res is not None else 0.0
            scores = {s: neg_acc(s) for s in self.cache}
            top = sorted(self.cache, key=lambda s: scores[s])[:self.m_top]
        top = _prt.share_ints(top, owner, min(self.m_top, len(self.cache)))
        self.model_handler([self.cache[s][0] for s in top], self.data[0])
        for s in self.cache:                          # canonical release order: same free lists everywhere
            self.cache[s][0].release()
        self.cache = {}
        for s in top:
            self.neigh_counter[s] += 1
        return None


class All2AllGossipNode(GossipNode):
    """Koloskova 2020 decentralised SGD node (ref ``node.py:789-870``).

    Receives only store the newest model per sender.  On timeout the cached models are merged
    with the mixing weights and a local update follows; then the node pushes to *all* peers.
    FIX(B17): weights are matched to senders by peer id and renormalised over the models that
    actually arrived (the reference pairs them by arrival order); ``reference_compat`` restores
    the arrival-order pairing.
    """

    def __init__(self, idx: int, data: NodeData, round_len: int, model_handler: WeightedTMH,
                 p2p_net: P2PNetwork, sync: bool = True) -> None:
        super().__init__(idx, data, round_len, model_handler, p2p_net, sync)
        self.local_cache: Dict[int, Any] = {}

    def timed_out(self, t: int, weights: Iterable[float]) -> bool:  # type: ignore[override]
        tout = super().timed_out(t)
        if tout:
            self.on_timeout(weights)
        return tout

    def on_timeout(self, weights: Iterable[float]) -> None:
        """Merge the cached neighbourhood with the mixing weights and train (the part of the
        reference's ``timed_out`` that has side effects; also called by the native scheduler)."""
        if self.local_cache:
            senders = list(self.local_cache.keys())
            models = [CACHE.pop(self.local_cache[s]) for s in senders]
            w = np.asarray(weights, dtype=float)
            if GlobalSettings().reference_compat:
                use = w[:len(models) + 1]
            else:
                peers = self.p2p_net.get_peers(self.idx)
                pos = {p: i + 1 for i, p in enumerate(peers)}
                use = np.array([w[0]] + [w[pos[s]] if s in pos and pos[s] < len(w) else 0.0
                                         for s in senders])
                if len(models) < len(peers):
                    total = 0.0
                    for u in use.tolist():           # sequential sum: the C++ executor reproduces it bit for bit
                        total += u
                    if total > 0:
                        use = use / total
            self.model_handler(models, self.data[0], use)
            for m in models:
                _release(m)
            self.local_cache = {}

    def get_peers(self) -> List[int]:
        return self.p2p_net.get_peers(self.idx)

    def send(self, t: int, peer: int, protocol: AntiEntropyProtocol) -> Message:
        if protocol != AntiEntropyProtocol.PUSH:
            raise ValueError("All2AllNode only supports PUSH protocol.")
        return super().send(t, peer, protocol)

    def receive(self, t: int, msg: Message) -> Optional[Message]:
        if msg.type == MessageType.PUSH:
            old = self.local_cache.get(msg.sender)
            if old is not None:
                CACHE.drop(old)
            self.local_cache[msg.sender] = msg.value[0]
        return None
