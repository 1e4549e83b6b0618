#!/usr/bin/env python
"""Node-count scaling: the reference's main_ormandi_2013 set-up at full size (one node per training
sample: 4 141 Pegasos nodes, clique, PUSH, UniformDelay(0,10), online .2, drop .1, 10 % evaluated
per round) -- per-event execution vs the banked engine (many nodes per launch) vs the unmodified reference.

    python benchmarks/many_nodes.py [--nodes 4141] [--rounds 20] [--impl banked|events|python|reference]
"""
import argparse
import json
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def data(n_nodes):
    g = torch.Generator().manual_seed(0)
    n = n_nodes + 460
    X = torch.randn(n, 57, generator=g)
    y = torch.sign(X @ torch.randn(57, generator=g) + 0.3 * torch.randn(n, generator=g))
    return X[:n_nodes], y[:n_nodes], X[n_nodes:], y[n_nodes:]


def ours(a):
    import gossipy_b200 as g
    from gossipy_b200.core import AntiEntropyProtocol, CreateModelMode, StaticP2PNetwork, UniformDelay
    from gossipy_b200.data import DataDispatcher
    from gossipy_b200.data.handler import ClassificationDataHandler
    from gossipy_b200.model.handler import PegasosHandler
    from gossipy_b200.model.nn import AdaLine
    from gossipy_b200.node import CacheNeighNode, GossipNode, PassThroughNode
    from gossipy_b200.simul import GossipSimulator, SimulationReport
    g.LOG.setLevel(50)
    dev = a.device
    g.GlobalSettings().set_device(dev)
    g.set_seed(98765)
    Xtr, ytr, Xte, yte = data(a.nodes)
    disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), eval_on_user=False, auto_assign=True)
    proto = PegasosHandler(AdaLine(57), .01, CreateModelMode.MERGE_UPDATE)
    node_cls = {"gossip": GossipNode, "passthrough": PassThroughNode, "cacheneigh": CacheNeighNode}[a.node]
    topo = StaticP2PNetwork(disp.size())
    if a.node == "passthrough":             # main_giaretta_2019: Barabasi-Albert graph (unequal degrees), m = 10
        rng = np.random.default_rng(7)
        n, m_ = disp.size(), 10
        A = np.zeros((n, n), dtype=np.int8)
        deg = np.zeros(n)
        for i in range(m_ + 1):
            for j in range(i):
                A[i, j] = A[j, i] = 1
        deg[:m_ + 1] = m_
        for i in range(m_ + 1, n):
            tgt = rng.choice(i, size=m_, replace=False, p=deg[:i] / deg[:i].sum())
            A[i, tgt] = A[tgt, i] = 1
            deg[tgt] += 1
            deg[i] = m_
        topo = StaticP2PNetwork(n, A)
    nodes = node_cls.generate(disp, topo, proto, 100, False)
    sim = GossipSimulator(nodes, disp, 100, AntiEntropyProtocol.PUSH, delay=UniformDelay(0, 10), online_prob=.2,
                          drop_prob=.1, sampling_eval=.1)
    sim.progress = False
    sim.engine = "python" if a.impl == "python" else "native"
    sim.batched = a.impl == "banked"
    rep = SimulationReport(); sim.add_receiver(rep)
    sim.init_nodes(seed=42)
    sim.start(2)                                   # warm-up
    if dev.startswith("cuda"):
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim.start(a.rounds, resume=True)
    if dev.startswith("cuda"):
        torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    from gossipy_b200 import ops
    ev = rep.get_evaluation(False)
    return {"impl": a.impl, "node_class": a.node, "banked": "_bank" in sim.__dict__, "device": dev, "nodes": disp.size(), "rounds": a.rounds, "rounds_per_s": a.rounds / sec,
            "ms_per_round": sec / a.rounds * 1e3, "last_eval": {k: round(v, 4) for k, v in ev[-1][1].items()},
            "sent": rep._sent_messages, "failed": rep._failed_messages, "native_launches": ops.launch_count}


def reference(a):
    for name in ("matplotlib", "matplotlib.pyplot", "pyparsing"):
        try:
            __import__(name)
        except Exception:
            m = types.ModuleType(name)
            if name == "pyparsing":
                m.ParseSyntaxException = Exception
            sys.modules[name] = m
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    import gossipy
    import gossipy.model.handler as H
    import gossipy.simul as S
    from gossipy import set_seed
    from gossipy.core import AntiEntropyProtocol, CreateModelMode, StaticP2PNetwork, UniformDelay
    from gossipy.data import DataDispatcher
    from gossipy.data.handler import ClassificationDataHandler
    from gossipy.model.handler import PegasosHandler
    from gossipy.model.nn import AdaLine
    from gossipy.node import GossipNode
    from gossipy.simul import GossipSimulator, SimulationReport
    _auc = H.roc_auc_score
    H.roc_auc_score = lambda *x, **k: np.float64(_auc(*x, **k))    # environment shim (new scikit-learn returns float)

    class _It:
        def __init__(self, it): self.it = it
        def __iter__(self): return iter(self.it)
        def close(self): pass
    S.track = lambda it, description="": _It(it)
    gossipy.LOG.setLevel(50)
    set_seed(98765)
    Xtr, ytr, Xte, yte = data(a.nodes)
    disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), eval_on_user=False, auto_assign=True)
    proto = PegasosHandler(net=AdaLine(57), learning_rate=.01, create_model_mode=CreateModelMode.MERGE_UPDATE)
    nodes = GossipNode.generate(data_dispatcher=disp, p2p_net=StaticP2PNetwork(disp.size(), None), model_proto=proto,
                                round_len=100, sync=False)
    sim = GossipSimulator(nodes=nodes, data_dispatcher=disp, delta=100, protocol=AntiEntropyProtocol.PUSH,
                          delay=UniformDelay(0, 10), online_prob=.2, drop_prob=.1, sampling_eval=.1)
    rep = SimulationReport(); sim.add_receiver(rep)
    sim.init_nodes(seed=42)
    t0 = time.perf_counter()
    sim.start(n_rounds=a.rounds)
    sec = time.perf_counter() - t0
    ev = rep.get_evaluation(False)
    return {"impl": "reference", "device": "cpu (the reference's default)", "nodes": disp.size(), "rounds": a.rounds,
            "rounds_per_s": a.rounds / sec, "ms_per_round": sec / a.rounds * 1e3,
            "last_eval": {k: round(float(v), 4) for k, v in ev[-1][1].items()}, "sent": rep._sent_messages,
            "failed": rep._failed_messages}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=4141)
    ap.add_argument("--rounds", type=int, default=20)
    ap.add_argument("--impl", default="banked", choices=["banked", "events", "python", "reference"])
    ap.add_argument("--device", default="cuda:0" if torch.cuda.is_available() else "cpu")
    ap.add_argument("--node", default="gossip", choices=["gossip", "passthrough", "cacheneigh"],
                    help="node class (the three variants of the reference's main_giaretta_2019.py)")
    a = ap.parse_args()
    print(json.dumps(reference(a) if a.impl == "reference" else ours(a)))
