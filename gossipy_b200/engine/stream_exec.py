"""Glue of the native executor (``csrc/exec/executor.cpp``): a round's event list is enqueued from C++.

Eligible simulations (``eligible`` returns ``None``): a :class:`GossipSimulator`, :class:`TokenizedGossipSimulator`
or (asynchronous) :class:`All2AllGossipSimulator` whose nodes are all of ONE class -- :class:`GossipNode`,
:class:`PassThroughNode`, :class:`CacheNeighNode`, :class:`SamplingBasedNode` + :class:`SamplingTMH`,
:class:`PartitioningBasedNode` + :class:`PartitionedTMH`, :class:`All2AllGossipNode` + :class:`WeightedTMH` -- with
handlers on the fused kernel path (1-hidden-layer ReLU MLP or logistic regression, SGD with or without momentum, mean
cross-entropy), any :class:`CreateModelMode` the node class allows, identical hyper-parameters on all nodes.
A :class:`PENSNode` population joins once every node has left its selection phase (step 2 = plain deliveries).
Everything else (PENS step 1, generic autograd models) keeps the per-event Python executor on the same native
schedule, linear learners the bank (``engine/bank.py``).

Several ranks: every rank drives its own executor over the same event list (replicated books), launches
the work of its own nodes only, and the snapshot slots are rows of the symmetric arenas
(``engine/arena.py``), so a reader on another GPU hands the slot's ``ready`` / ``done`` words to the
training kernel exactly like the Python executor does.

Python owns all memory -- the handlers' arena rows, ONE tensor of snapshot slots, the nodes' torch streams
(so evaluation, which stays in Python, is ordered after the native launches on the same streams) -- and
mirrors ages / update counters back into the handlers after every ``start``.  On CPU the two launches are
Python callbacks into :mod:`gossipy_b200.ops`, which makes the native bookkeeping testable without a GPU.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from .. import ops
from ..core import AntiEntropyProtocol, CreateModelMode
from . import arena as _arena
from . import rng as _rng


def eligible(sim: Any) -> Optional[str]:
    """``None`` when the native executor can run ``sim``, else the reason it cannot."""
    from ..model import handler as H
    from ..node import GossipNode
    from ..parallel import runtime as prt
    from .. import GlobalSettings
    if GlobalSettings().reference_compat:
        return "reference_compat (bug-for-bug behaviours live in the Python handlers)"
    a2a = type(sim).__name__ == "All2AllGossipSimulator"
    if type(sim).__name__ not in ("GossipSimulator", "TokenizedGossipSimulator") and not a2a:
        return "simulator variant"
    if prt.active() and prt.transport() != "p2p":
        return "several ranks without shared arenas"
    if not H.FUSE_MERGE_UPDATE:
        return "fused merge+update disabled"
    ids = sorted(sim.nodes)
    if ids != list(range(len(ids))):
        return "node ids must be 0..N-1"
    from ..node import (All2AllGossipNode, CacheNeighNode, PartitioningBasedNode, PassThroughNode, PENSNode,
                        SamplingBasedNode)
    if a2a:
        if sim.protocol != AntiEntropyProtocol.PUSH:
            return "all-to-all: protocol %s" % sim.protocol.name
        if getattr(sim, "_W", None) is None:
            return "all-to-all: no mixing matrix"
    ref = None
    cls0 = type(sim.nodes[ids[0]])
    for i in ids:
        node = sim.nodes[i]
        h = node.model_handler
        partitioned = type(node) is PartitioningBasedNode and type(h) is H.PartitionedTMH
        sampled = type(node) is SamplingBasedNode and type(h) is H.SamplingTMH
        weighted = a2a and type(node) is All2AllGossipNode and type(h) is H.WeightedTMH
        if type(node) is not cls0:
            return "mixed node classes"
        if a2a != weighted:
            return "all-to-all needs All2AllGossipNode + WeightedTMH"
        if weighted and (h.mode != CreateModelMode.MERGE_UPDATE or node.local_cache and "_stream_exec" not in sim.__dict__):
            return "all-to-all: mode %s / caches filled by another executor" % h.mode.name
        # a PENSNode that has left its selection phase consumes a delivery like a plain node (ref node.py:752-757);
        # its restricted peer choice lives in the scheduler
        pens2 = type(node) is PENSNode and node.step == 2 and type(h) is H.TorchModelHandler
        if pens2 and h.device.type == "cuda" and os.environ.get("GOSSIPY_EXEC_PENS_STEP2", "") != "1":
            # the hand-over in the middle of a run (executor built, messages on the wire moved into its slots) was written
            # after the last GPU session: exact on CPU with 1 - 2 ranks, not yet run on a GPU -> opt-in there
            return "PENS step 2 on the C++ executor is opt-in on CUDA (GOSSIPY_EXEC_PENS_STEP2=1) until validated on a GPU"
        if (type(node) not in (GossipNode, PassThroughNode, CacheNeighNode) and not partitioned and not sampled
                and not weighted and not pens2):
            return "node class %s" % type(node).__name__
        if type(node) in (PassThroughNode, CacheNeighNode) and not getattr(node, "_keyed_draws", False):
            return "node-side draws from the host stream"
        if type(node) is CacheNeighNode and node.local_cache and "_stream_exec" not in sim.__dict__:
            return "neighbour caches filled by another executor"
        if type(h) not in (H.TorchModelHandler, H.LimitedMergeTMH) and not partitioned and not sampled and not weighted:
            return "handler class %s" % type(h).__name__
        if sampled and h.mode not in (CreateModelMode.MERGE_UPDATE, CreateModelMode.UPDATE):
            return "sampled models: mode %s" % h.mode.name
        if (sampled and h.mode == CreateModelMode.UPDATE and h.device.type == "cuda"
                and os.environ.get("GOSSIPY_EXEC_PART_UPDATE", "") != "1"):
            return "sampled UPDATE on CUDA is opt-in (GOSSIPY_EXEC_PART_UPDATE=1) until it has been validated on a GPU"
        momentum = bool(h.__dict__.get("_fused_momentum")) and not h._fused
        if momentum:            # torch.optim.SGD with momentum inside the tensor-core kernel: plain nodes, MERGE_UPDATE
            n_i = int(node.data[0][0].shape[0]) if isinstance(node.data[0], (tuple, list)) else 0
            if (partitioned or sampled or weighted or h.mode != CreateModelMode.MERGE_UPDATE
                    or not ops.mlp1_momentum_supported(h._family[1], h.batch_size, n_i)):
                return "momentum-SGD outside the fused envelope"
        if not (h._fused or momentum) or h.layout.int_buffers:
            return "handler is not on the fused kernel path"
        if partitioned:
            if h.mode not in (CreateModelMode.MERGE_UPDATE, CreateModelMode.UPDATE):
                return "partitioned models: mode %s" % h.mode.name
            if (h.mode == CreateModelMode.UPDATE and h.device.type == "cuda"
                    and os.environ.get("GOSSIPY_EXEC_PART_UPDATE", "") != "1"):
                # the launch sequence (copy, train the copy with ages by value, segment merge) was written after the last
                # GPU session: bit-identical to the per-event executor on CPU, not yet run on a GPU -> opt-in there
                return "partitioned UPDATE on CUDA is opt-in (GOSSIPY_EXEC_PART_UPDATE=1) until it has been validated on a GPU"
            if h.tm_partition.n_parts > 16:
                return "more than 16 partitions"
        elif h.mode not in (CreateModelMode.MERGE_UPDATE, CreateModelMode.UPDATE, CreateModelMode.UPDATE_MERGE,
                            CreateModelMode.PASS):
            return "mode %s" % h.mode.name
        elif not isinstance(h.n_updates, (int, np.integer)):
            return "vector-valued model age"
        sig = (h._family, h.batch_size, h.local_epochs, float(h.optimizer_params.get("lr", 1e-3)),
               float(h.optimizer_params.get("weight_decay", 0.0)), h._row_numel, type(h), h.mode,
               getattr(h, "L", None), h.tm_partition.n_parts if partitioned else 0, getattr(h, "sample_size", None),
               (float(h.optimizer_params.get("momentum", 0.0)), float(h.optimizer_params.get("dampening", 0.0)),
                bool(h.optimizer_params.get("nesterov", False))) if momentum else None)
        if ref is None:
            ref = sig
        elif sig != ref:
            return "nodes differ in architecture or hyper-parameters"
        x = node.data[0][0] if isinstance(node.data[0], (tuple, list)) else None
        if x is None or int(x.shape[0]) == 0:
            return "node without training data"
    return None


class StreamExec:
    def __init__(self, sim: Any) -> None:
        from ..ops.native import _try_import
        self.sim = sim
        self.C = _try_import()
        ids = sorted(sim.nodes)
        h0 = sim.nodes[ids[0]].model_handler
        self.device = h0.device
        self.cuda = self.device.type == "cuda"
        fam, dims = h0._family
        self.family = fam
        self.dims = tuple(int(d) for d in dims)
        self.bs, self.epochs = int(h0.batch_size), int(h0.local_epochs)
        self.lr = float(h0.optimizer_params.get("lr", 1e-3))
        self.wd = float(h0.optimizer_params.get("weight_decay", 0.0))
        self.row_numel = int(h0._row_numel)
        if fam == "mlp1":
            IN, Hd, OUT = self.dims
        else:
            IN, OUT = self.dims[0], self.dims[-1]
            Hd = 0
        limited = int(h0.L) if hasattr(h0, "L") else -1          # LimitedMergeTMH: age-limited merge weights
        self.ex = self.C.StreamExecutor(len(ids), 0 if fam == "mlp1" else 1, IN, Hd, OUT, self.bs, self.epochs,
                                        self.lr, self.wd, _rng.base_seed(), self.cuda, int(h0.mode.value), limited)
        self.n_parts = 0
        if type(h0).__name__ == "PartitionedTMH":       # per-partition ages, segment merges, 1/age gradient scaling (K3)
            part = h0.tm_partition
            self.n_parts = int(part.n_parts)
            self._part_id = part.part_id.to(self.device)
            self._segs = [part.segments(p).to(self.device).contiguous() for p in range(self.n_parts)]
            if self.cuda:
                self.ex.set_partition(self.n_parts, self._part_id.data_ptr(), [t.data_ptr() for t in self._segs],
                                      [int(t.shape[0]) for t in self._segs])
            else:
                self.ex.set_partition(self.n_parts, 0, [0] * self.n_parts, [int(t.shape[0]) for t in self._segs])
                self.ex.set_partition_callbacks(self._cb_merge_part, self._cb_train_part)
        from ..parallel import runtime as prt
        self.multi = prt.active()
        self.rank = prt.rank() if self.multi else 0
        self._data: Dict[int, Any] = {}
        if self.multi:
            self.owner = [int(prt.rank_of(i)) for i in ids]
            self.ex.set_ranks(self.rank, prt.world(), self.owner)
            # a rank's slots are rows of ITS symmetric arena (mapped everywhere); every rank performs the same
            # allocations on every mirror, so (rank, slot) means the same row in all processes
            per_rank = max(self.owner.count(r) for r in range(prt.world()))
            self.pool_rows = [[] for _ in range(prt.world())]
            self._add_pool_rows(max(8, 6 * per_rank))
            self.slots = None
        else:
            self.owner = [0] * len(ids)
            self.slots = torch.zeros(max(32, 6 * len(ids)), self.row_numel, dtype=torch.float32, device=self.device)
            self._publish_slots()
        if not self.cuda:
            self.ex.set_callbacks(self._cb_snapshot, self._cb_train, self._cb_adopt)
        from ..node import PassThroughNode
        self.passthrough = type(sim.nodes[ids[0]]) is PassThroughNode
        self.sample_k = 0
        if type(h0).__name__ == "SamplingTMH":          # the receiver merges k keyed coordinates (with replacement), then trains
            from ..model.sampling import TorchModelSampling
            self.n_params = int(h0.layout.n_params)
            self.sample_k = int(TorchModelSampling.sample_size(h0.sample_size, self.n_params))
            self.ex.set_sampling(self.sample_k, self.n_params)
            mine = [i for i in ids if self.owner[i] == self.rank]
            self._samp_idx = torch.zeros(max(1, len(mine)), self.sample_k, dtype=torch.int64, device=self.device)
            self._samp_val = torch.zeros(max(1, len(mine)), self.sample_k, dtype=torch.float32, device=self.device)
            for k_, i in enumerate(mine):
                self.ex.set_node_sample_buffers(i, self._samp_idx[k_].data_ptr(), self._samp_val[k_].data_ptr())
            if not self.cuda:
                self.ex.set_sample_merge_callback(self._cb_sample_merge)
        from ..node import CacheNeighNode
        self.cacheneigh = type(sim.nodes[ids[0]]) is CacheNeighNode
        self.momentum = None
        if h0.__dict__.get("_fused_momentum") and not h0._fused:
            p0 = h0.optimizer_params
            self.momentum = (float(p0.get("momentum", 0.0)), float(p0.get("dampening", 0.0)), bool(p0.get("nesterov", False)))
            self.ex.set_momentum(*self.momentum)
            if not self.cuda:
                self.ex.set_merge_pair_callback(self._cb_merge_pair)
        self.a2a = type(sim).__name__ == "All2AllGossipSimulator"
        if self.a2a:                    # cached neighbourhood + k-way merge on timeout; the pushes of a timeout share a snapshot
            self.ex.set_all2all(True)
            if not self.cuda:
                self.ex.set_kway_callback(self._cb_kway)
        self._scratch = None
        if h0.mode == CreateModelMode.UPDATE_MERGE or ((self.n_parts or self.sample_k) and h0.mode == CreateModelMode.UPDATE):
            # one private row per node of this rank for the copy of a received model that is trained
            mine = [i for i in ids if self.owner[i] == self.rank]
            self._scratch = torch.zeros(max(1, len(mine)), self.row_numel, dtype=torch.float32, device=self.device)
            self._scratch_of = {i: k for k, i in enumerate(mine)}
            for i, k in self._scratch_of.items():
                self.ex.set_node_scratch(i, self._scratch[k].data_ptr())
            if not self.cuda:
                self.ex.set_update_merge_callback(self._cb_update_merge)
                self.ex.set_partition_update_callback(self._cb_update_part)
                self.ex.set_sample_update_callback(self._cb_sample_update)
        self.bind_nodes()

    def _add_pool_rows(self, k: int) -> None:
        """``k`` more snapshot slots per rank (replicated: every rank allocates the same rows on every mirror; the
        symmetric arenas add a segment collectively when they run out)."""
        for r, rows in enumerate(self.pool_rows):
            for _ in range(k):
                row = _arena.arena_for(self.device, self.row_numel, r).alloc()
                rows.append(row)
                ready = row.flag_ready if self.cuda else 0
                done = row.flag_done if self.cuda else 0
                self.ex.add_slot(r, row.tensor.data_ptr(), int(ready), int(done), self.row_numel, int(row.gen),
                                 int(row.remote_reads), int(getattr(row, "_acked", 0)))

    # -- state shared with the handlers ---------------------------------------------------------------
    def _publish_slots(self) -> None:
        self.ex.set_slots(self.slots.data_ptr(), int(self.slots.shape[0]), int(self.slots.stride(0)), self.row_numel)

    def _node_data(self, i: int):
        h = self.sim.nodes[i].model_handler
        with _arena.on_stream(h._stream()):
            x, y = h._to_device(self.sim.nodes[i].data[0])
        if x.dim() > 2:
            x = x.reshape(x.shape[0], -1)
        if y.dim() > 1:
            y = torch.argmax(y, dim=-1)
        x, y = x.contiguous(), y.contiguous()
        if y.dtype != torch.int64:
            y = y.long()
        self._data[i] = (x, y)                  # keeps the tensors alive while C++ holds their addresses
        return x, y

    def _mine(self, i: int) -> bool:
        return self.owner[i] == self.rank

    def bind_nodes(self) -> None:
        """(Re)read rows, data, ages, counters and streams from the handlers (start of every ``start``)."""
        for i, node in self.sim.nodes.items():
            h = node.model_handler
            age = int(np.sum(h.n_updates))
            if not self._mine(i):      # another rank runs this node: only its sample count and counters matter here
                self.ex.set_node(i, 0, 0, 0, int(node.data[0][0].shape[0]), age, int(h._update_counter), 0)
            else:
                row = h.row
                x, y = self._node_data(i)
                s = h._stream()
                self.ex.set_node(i, row.data_ptr(), x.data_ptr(), y.data_ptr(), int(x.shape[0]), age,
                                 int(h._update_counter), int(s.cuda_stream) if s is not None else 0)
            if self.n_parts:
                self.ex.set_node_ages(i, [int(a) for a in h.n_updates], int(getattr(node, "_model_msgs", 0)))
        if self.a2a:
            for i, node in self.sim.nodes.items():
                self.ex.set_node_mixing(i, [int(p) for p in node.p2p_net.get_peers(i)],
                                        [float(w) for w in np.asarray(self.sim._W[i], dtype=float)])
        if self.momentum is not None:
            self._mom_first = {}
            for i, node in self.sim.nodes.items():
                if not self._mine(i):
                    continue
                h = node.model_handler
                buf = h._opt_rows.get("momentum")
                first = buf is None or bool(h.__dict__.get("_mom_pending"))
                if buf is None:
                    buf = h._opt_rows["momentum"] = torch.zeros_like(h.row)
                self.ex.set_node_momentum(i, buf.data_ptr(), first)
                self._mom_first[i] = first          # (CPU: the callback trains, so the flag is kept here)
        if self.cacheneigh:
            self.ex.set_cache_neigh([int(getattr(self.sim.nodes[i], "_cn_draws", 0)) for i in sorted(self.sim.nodes)])
        if self.passthrough:
            ids = sorted(self.sim.nodes)
            self.ex.set_passthrough([int(self.sim.nodes[i].n_neighs) for i in ids],
                                    [int(getattr(self.sim.nodes[i], "_pt_draws", 0)) for i in ids])

    def refresh_data(self) -> None:
        """Streamed inputs: the resident buffers alternate every round."""
        for i in self.sim.nodes:
            if self._mine(i):
                x, y = self._node_data(i)
                self.ex.set_node_data(i, x.data_ptr(), y.data_ptr(), int(x.shape[0]))

    def sync_back(self) -> None:
        ages, counters = self.ex.ages(), self.ex.counters()
        if self.n_parts:
            ages_v, msgs = self.ex.ages_v(), self.ex.model_msgs()
            for i, node in self.sim.nodes.items():
                h = node.model_handler
                new = np.asarray(ages_v[i], dtype=h.n_updates.dtype)
                if not np.array_equal(new, h.n_updates) or h._update_counter != counters[i]:
                    h._version += 1
                h.n_updates = new
                h._update_counter = int(counters[i])
                node._model_msgs = int(msgs[i])
            return
        if self.momentum is not None:       # a buffer the executor created but never used still has no state
            for i, first in enumerate(self.ex.mom_first()):
                if self._mine(i):
                    self.sim.nodes[i].model_handler._mom_pending = bool(first if self.cuda else self._mom_first.get(i, False))
        draws = self.ex.pt_draws() if self.passthrough else None
        if self.cacheneigh:
            for i, c in enumerate(self.ex.cn_draws()):
                self.sim.nodes[i]._cn_draws = int(c)
        for i, node in self.sim.nodes.items():
            h = node.model_handler
            if int(h.n_updates) != ages[i] or h._update_counter != counters[i]:
                h._version += 1
            h.n_updates = int(ages[i])
            h._update_counter = int(counters[i])
            if draws is not None:
                node._pt_draws = int(draws[i])

    # -- CPU callbacks (the same handshakes the kernels perform on a GPU, on the shared-memory flags) ----------
    def _slot(self, rank: int, slot: int, gen: int):
        """(tensor, cross-rank handshake or None) of a snapshot slot as seen by THIS rank."""
        if rank < 0:                    # elided snapshot: the sender's live row (same rank by construction)
            return self.sim.nodes[-1 - rank].model_handler.row, None
        if not self.multi:
            return self.slots[slot], None
        row = self.pool_rows[rank][slot]
        if rank == self.rank:
            return row.tensor, None
        done = row.flag_done
        return row.tensor, ops.RowSync(row.flag_ready, gen, (done[0], done[1] + self.rank))

    def _cb_snapshot(self, node: int, rank: int, slot: int, gen: int, remote_reads: int) -> None:
        src = self.sim.nodes[node].model_handler.row
        if not self.multi:
            self.slots[slot].copy_(src)
            return
        row = self.pool_rows[rank][slot]
        row.remote_reads = remote_reads
        _arena.wait_remote_readers(row)             # every reader on another rank has acknowledged the slot's last life
        row.tensor.copy_(src)
        row.gen = gen
        arr, i = row.flag_ready
        arr[i] = gen                                # publish

    def _cb_adopt(self, node: int, rank: int, slot: int, gen: int) -> None:
        src, sync = self._slot(rank, slot, gen)
        ops.merge_pair(self.sim.nodes[node].model_handler.row, src, 0.0, 1.0, sync=sync)

    def _cb_merge_pair(self, node: int, rank: int, slot: int, w_self: float, w_peer: float, gen: int) -> None:
        src, sync = self._slot(rank, slot, gen)
        ops.merge_pair(self.sim.nodes[node].model_handler.row, src, float(w_self), float(w_peer), sync=sync)

    def _cb_train(self, node: int, rank: int, slot: int, key: int, w_self: float, w_peer: float, gen: int) -> None:
        h = self.sim.nodes[node].model_handler
        x, y = self._data[node]
        fn = ops.mlp1_train if self.family == "mlp1" else ops.logreg_train
        merge = None
        if slot >= 0:
            src, sync = self._slot(rank, slot, gen)
            merge = (src, float(w_self), float(w_peer), sync)
        if self.momentum is not None:
            first = bool(self._mom_first.get(node, False))
            self._mom_first[node] = False
            fn(h.row, x, y, self.dims, self.bs, self.epochs, self.lr, self.wd, int(key), None, merge_from=merge,
               momentum=self.momentum + (h._opt_rows["momentum"], first))
            return
        fn(h.row, x, y, self.dims, self.bs, self.epochs, self.lr, self.wd, int(key), None, merge_from=merge)

    def _cb_update_merge(self, node: int, rank: int, slot: int, key_own: int, key_tmp: int, w_self: float, w_peer: float,
                         gen: int) -> None:
        h = self.sim.nodes[node].model_handler
        x, y = self._data[node]
        fn = ops.mlp1_train if self.family == "mlp1" else ops.logreg_train
        fn(h.row, x, y, self.dims, self.bs, self.epochs, self.lr, self.wd, int(key_own), None)
        tmp = self._scratch[self._scratch_of[node]]
        src, sync = self._slot(rank, slot, gen)
        ops.merge_pair(tmp, src, 0.0, 1.0, sync=sync)
        fn(tmp, x, y, self.dims, self.bs, self.epochs, self.lr, self.wd, int(key_tmp), None)
        if w_peer != 0.0:
            ops.merge_pair(h.row, tmp, float(w_self), float(w_peer))

    def _cb_kway(self, node: int, srcs: List[List[int]], weights: List[float]) -> None:
        rows, syncs = [], []
        for rank, slot, gen in srcs:
            t, sync = self._slot(int(rank), int(slot), int(gen))
            rows.append(t)
            syncs.append(sync)
        ops.merge_kway(self.sim.nodes[node].model_handler.row, rows, [float(w) for w in weights],
                       syncs if any(s is not None for s in syncs) else None)

    def _cb_sample_merge(self, node: int, rank: int, slot: int, key: int, gen: int) -> None:
        src, sync = self._slot(rank, slot, gen)
        idx = ops.keyed_randint(self.sample_k, self.n_params, int(key), self.device)
        ops.merge_indexed(self.sim.nodes[node].model_handler.row, src, idx, 0.5, 0.5, sync)

    def _cb_sample_update(self, node: int, rank: int, slot: int, gen: int, key_s: int, key_tmp: int) -> None:
        """Sampled UPDATE: train a private copy of the received model, merge its sampled coordinates into the own model."""
        h = self.sim.nodes[node].model_handler
        x, y = self._data[node]
        tmp = self._scratch[self._scratch_of[node]]
        src, sync = self._slot(rank, slot, gen)
        ops.merge_pair(tmp, src, 0.0, 1.0, sync=sync)
        fn = ops.mlp1_train if self.family == "mlp1" else ops.logreg_train
        fn(tmp, x, y, self.dims, self.bs, self.epochs, self.lr, self.wd, int(key_tmp), None)
        idx = ops.keyed_randint(self.sample_k, self.n_params, int(key_s), self.device)
        ops.merge_indexed(h.row, tmp, idx, 0.5, 0.5, None)

    def _cb_merge_part(self, node: int, rank: int, slot: int, pid: int, w1: float, w2: float, gen: int) -> None:
        src, sync = self._slot(rank, slot, gen)
        ops.merge_segments(self.sim.nodes[node].model_handler.row, src, self._segs[pid], float(w1), float(w2), sync)

    def _cb_update_part(self, node: int, rank: int, slot: int, gen: int, key: int, ages: List[int], pid: int, w1: float,
                        w2: float) -> None:
        """Partitioned UPDATE: train a private copy of the received model (its ages scale the gradient), merge its
        partition ``pid`` into the own model."""
        h = self.sim.nodes[node].model_handler
        x, y = self._data[node]
        tmp = self._scratch[self._scratch_of[node]]
        src, sync = self._slot(rank, slot, gen)
        ops.merge_pair(tmp, src, 0.0, 1.0, sync=sync)
        fn = ops.mlp1_train if self.family == "mlp1" else ops.logreg_train
        fn(tmp, x, y, self.dims, self.bs, self.epochs, self.lr, self.wd, int(key),
           (self._part_id, torch.as_tensor(ages, dtype=torch.int64)))
        ops.merge_segments(h.row, tmp, self._segs[pid], float(w1), float(w2), None)

    def _cb_train_part(self, node: int, key: int, ages: List[int]) -> None:
        h = self.sim.nodes[node].model_handler
        x, y = self._data[node]
        fn = ops.mlp1_train if self.family == "mlp1" else ops.logreg_train
        fn(h.row, x, y, self.dims, self.bs, self.epochs, self.lr, self.wd, int(key),
           (self._part_id, torch.as_tensor(ages, dtype=torch.int64)))

    # -- one round -------------------------------------------------------------------------------------------
    def run_round(self, events: np.ndarray) -> List[int]:
        before = self.ex.launches
        if self.cuda:
            with torch.cuda.device(self.device):
                evals = self._run(events)
            ops._count(int(self.ex.launches - before))
            return evals
        return self._run(events)

    def _run(self, events: np.ndarray) -> List[int]:
        evals = list(self.ex.run(events, 0))
        while self.ex.resume_at >= 0:           # out of snapshot slots: grow the pool, continue where it stopped
            self._grow()
            evals += list(self.ex.run(events, int(self.ex.resume_at)))
        return evals

    def _grow(self) -> None:
        if self.multi:          # replicated books: every rank runs out of the same pool at the same event
            self._add_pool_rows(len(self.pool_rows[0]))
            return
        if self.cuda:
            torch.cuda.synchronize(self.device)
        bigger = torch.zeros(2 * int(self.slots.shape[0]), self.row_numel, dtype=torch.float32, device=self.device)
        bigger[:self.slots.shape[0]].copy_(self.slots)
        if self.cuda:
            torch.cuda.synchronize(self.device)
        self.slots = bigger
        self._publish_slots()

    # -- checkpointing -----------------------------------------------------------------------------------------
    def _slot_rows(self, where: List[Any]) -> torch.Tensor:
        """The contents of the slots ``[(rank, slot), ...]`` as one CPU tensor.  Several ranks: every owner contributes its
        slots (one small all-gather), so the checkpoint of every rank is complete."""
        if not where:
            return torch.zeros(0, self.row_numel)
        if not self.multi:
            idx = torch.as_tensor([s for _, s in where], dtype=torch.int64, device=self.device)
            return self.slots[idx].cpu()
        import torch.distributed as dist
        if self.cuda:
            torch.cuda.synchronize(self.device)
        local = {(int(r), int(s)): self.pool_rows[int(r)][int(s)].tensor.detach().cpu().clone()
                 for r, s in where if int(r) == self.rank}
        gathered: List[Any] = [None] * len(self.pool_rows)
        dist.all_gather_object(gathered, local)
        merged = {k: v for part in gathered for k, v in part.items()}
        return torch.stack([merged[(int(r), int(s))] for r, s in where])

    def drop_inflight(self) -> int:
        """Release the slots of the messages that are still on the wire (a fresh ``start`` forgets them; models cached by
        cache-neighbour / all-to-all nodes are node state and stay).  Replicated: every rank drops the same ids."""
        ids = [int(r[0]) for r in self.ex.inflight()]
        if ids:
            ev = np.zeros((len(ids), 6), dtype=np.int32)
            ev[:, 0] = int(self.C.EV_DROP)
            ev[:, 4] = ids
            self.ex.run(ev, 0)
        return len(ids)

    def export_inflight(self) -> Dict[str, Any]:
        rows = self.ex.inflight()
        out = {"ids": [int(r[0]) for r in rows], "ranks": [int(r[1]) for r in rows], "ages": [int(r[3]) for r in rows],
               "extra": [list(map(int, r[4:])) for r in rows], "rows": self._slot_rows([(r[1], r[2]) for r in rows])}
        if self.a2a or self.cacheneigh:  # models waiting in the neighbour caches are state as well
            ent = self.ex.caches()
            out["cache"] = {"nodes": [int(e[0]) for e in ent], "senders": [int(e[1]) for e in ent], "ranks": [int(e[2]) for e in ent],
                            "ages": [int(e[4]) for e in ent], "rows": self._slot_rows([(e[2], e[3]) for e in ent])}
        return out

    def _restore_slots(self, ranks: List[int], rows: torch.Tensor) -> List[int]:
        """Free slots (one per entry, in the pool of the entry's rank -- the same choice on every rank) filled with ``rows``
        by their owners."""
        n = len(ranks)
        if not self.multi:
            while int(self.slots.shape[0]) < n or self.ex.free_slots < n:
                self._grow()
            taken = {int(r[2]) for r in self.ex.inflight()} | {int(e[3]) for e in self.ex.caches()}
            free = [s for s in range(int(self.slots.shape[0])) if s not in taken][:n]
            self.slots[torch.as_tensor(free, dtype=torch.int64, device=self.device)] = rows.to(self.device)
            return free
        taken = {(int(r[1]), int(r[2])) for r in self.ex.inflight()} | {(int(e[2]), int(e[3])) for e in self.ex.caches()}
        chosen: List[int] = []
        for i, rk in enumerate(ranks):
            while True:
                free = [s for s in range(len(self.pool_rows[rk])) if (rk, s) not in taken]
                if free:
                    break
                self._add_pool_rows(len(self.pool_rows[0]))
            taken.add((rk, free[0]))
            chosen.append(free[0])
            if rk == self.rank:
                self.pool_rows[rk][free[0]].tensor.copy_(rows[i].to(self.device))
        return chosen

    def _after_restore(self) -> None:
        if self.cuda:
            torch.cuda.synchronize(self.device)     # the node streams read these slots without a writer event
        if self.multi:
            from ..parallel import runtime as prt
            prt.barrier()                           # ... and no rank reads a peer's slot before its owner has filled it

    def import_inflight(self, st: Dict[str, Any]) -> None:
        cache = st.get("cache")
        if cache is not None and len(cache["nodes"]):
            ranks = [int(r) for r in cache.get("ranks", [0] * len(cache["nodes"]))]
            free = self._restore_slots(ranks, cache["rows"])
            self.ex.import_cache([[int(nd), int(sd), int(s), int(a), int(rk)] for nd, sd, s, a, rk in
                                  zip(cache["nodes"], cache["senders"], free, cache["ages"], ranks)])
        n = len(st["ids"])
        if n:
            ranks = [int(r) for r in st.get("ranks", [0] * n)]
            free = self._restore_slots(ranks, st["rows"])
            extra = st.get("extra") or [[] for _ in range(n)]
            self.ex.import_inflight([[int(m), int(rk), int(s), int(a)] + list(e)
                                     for m, rk, s, a, e in zip(st["ids"], ranks, free, st["ages"], extra)])
        self._after_restore()
