#!/usr/bin/env python
"""Test-accuracy-vs-round of the headline experiment: this framework (native engine + fused fp32-equivalent kernels)
next to the UNMODIFIED reference (baseline/_ref, stock path, device = cuda), several seeds each.

BASELINE.json's metric is "rounds/sec AND test-acc-vs-round"; the two implementations draw their gossip schedules
and mini-batch orders from different random streams, so single runs differ -- what must agree is the distribution.
For every round the script reports mean and standard deviation over the seeds of both arms and checks that our mean
lies inside the reference's band:  |mean_ours - mean_ref| <= 3 * sqrt(var_ours/S + var_ref/S) + 0.01.

    python benchmarks/acc_band.py --seeds 5 --rounds 50 [--small] > profiles/.../acc_band.json

``--small`` uses 8 x 600 training samples and 2 000 test samples (the GPU test-suite variant, ~20 s).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def data(small: bool, seed: int):
    Xtr, ytr, Xte, yte = bench.make_data(seed)
    if small:
        return Xtr[:4800], ytr[:4800], Xte[:2000], yte[:2000]
    return Xtr, ytr, Xte, yte


def run_ours(seed: int, rounds: int, small: bool, train_impl: str = ""):
    import gossipy_b200 as g
    from gossipy_b200 import CACHE, ops
    from gossipy_b200.core import AntiEntropyProtocol, CreateModelMode, StaticP2PNetwork
    from gossipy_b200.data import AssignmentHandler, DataDispatcher
    from gossipy_b200.data.handler import ClassificationDataHandler
    from gossipy_b200.model.handler import TorchModelHandler
    from gossipy_b200.model.nn import TorchMLP
    from gossipy_b200.node import GossipNode
    from gossipy_b200.simul import GossipSimulator, SimulationReport
    g.GlobalSettings().set_device("cuda:0" if torch.cuda.is_available() else "cpu")
    ops.set_train_impl(train_impl)
    g.set_seed(1000 + seed)
    Xtr, ytr, Xte, yte = data(small, 0)
    disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=8, eval_on_user=False, auto_assign=False)
    disp.set_assignments(bench.make_split(ytr), None)      # identical shards in both arms
    proto = TorchModelHandler(net=TorchMLP(784, 10, (100,)), optimizer=torch.optim.SGD, optimizer_params={"lr": .1},
                              criterion=torch.nn.CrossEntropyLoss(), local_epochs=1, batch_size=32,
                              create_model_mode=CreateModelMode.MERGE_UPDATE)
    nodes = GossipNode.generate(disp, StaticP2PNetwork(8), proto, round_len=100, sync=True)
    sim = GossipSimulator(nodes, disp, 100, AntiEntropyProtocol.PUSH_PULL)
    sim.progress = False
    sim.engine = "native"
    rep = SimulationReport()
    sim.add_receiver(rep)
    sim.init_nodes(seed=4242 + seed)
    sim.start(rounds)
    acc = [float(e["accuracy"]) for _, e in rep.get_evaluation(False)]
    CACHE.clear()
    return acc


def run_reference(seed: int, rounds: int, small: bool):
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
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
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    import gossipy
    from gossipy import GlobalSettings, set_seed
    from gossipy.core import AntiEntropyProtocol, CreateModelMode, StaticP2PNetwork
    from gossipy.data import AssignmentHandler, DataDispatcher
    from gossipy.data.handler import ClassificationDataHandler
    from gossipy.model.handler import TorchModelHandler
    from gossipy.model.nn import TorchMLP
    from gossipy.node import GossipNode
    from gossipy.simul import GossipSimulator, SimulationReport
    import gossipy.simul as S

    class _It:
        def __init__(self, it): self.it = it
        def __iter__(self): return iter(self.it)
        def close(self): pass
    S.track = lambda it, description="": _It(it)
    gossipy.LOG.setLevel(50)
    GlobalSettings().set_device("cuda" if torch.cuda.is_available() else "cpu")
    set_seed(1000 + seed)
    Xtr, ytr, Xte, yte = data(small, 0)
    disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=8, eval_on_user=False, auto_assign=False)
    disp.set_assignments(bench.make_split(ytr), None)      # identical shards in both arms
    proto = TorchModelHandler(net=TorchMLP(784, 10, (100,)), optimizer=torch.optim.SGD, optimizer_params={"lr": .1},
                              criterion=torch.nn.CrossEntropyLoss(), local_epochs=1, batch_size=32,
                              create_model_mode=CreateModelMode.MERGE_UPDATE)
    nodes = GossipNode.generate(data_dispatcher=disp, p2p_net=StaticP2PNetwork(8), model_proto=proto, round_len=100, sync=True)
    sim = GossipSimulator(nodes=nodes, data_dispatcher=disp, delta=100, protocol=AntiEntropyProtocol.PUSH_PULL)
    rep = SimulationReport()
    sim.add_receiver(rep)
    sim.init_nodes(seed=4242 + seed)
    sim.start(n_rounds=rounds)
    acc = [float(e["accuracy"]) for _, e in rep.get_evaluation(False)]
    gossipy.CACHE.clear()
    return acc


def band(ours, ref):
    a, b = np.asarray(ours), np.asarray(ref)          # [seeds, rounds]
    S = a.shape[0]
    n = min(a.shape[1], b.shape[1])
    a, b = a[:, :n], b[:, :n]
    ma, mb = a.mean(0), b.mean(0)
    tol = 3.0 * np.sqrt(a.var(0, ddof=1) / S + b.var(0, ddof=1) / S) + 0.01
    gap = np.abs(ma - mb)
    return {"rounds": n, "seeds": S, "ours_mean": [round(float(v), 4) for v in ma], "ours_std": [round(float(v), 4) for v in a.std(0, ddof=1)],
            "ref_mean": [round(float(v), 4) for v in mb], "ref_std": [round(float(v), 4) for v in b.std(0, ddof=1)],
            "max_gap": float(gap.max()), "max_gap_over_tolerance": float((gap / tol).max()), "inside_band": bool((gap <= tol).all())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=5)
    ap.add_argument("--rounds", type=int, default=50)
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--train-impl", default="")
    a = ap.parse_args()
    ours = [run_ours(s, a.rounds, a.small, a.train_impl) for s in range(a.seeds)]
    ref = [run_reference(s, a.rounds, a.small) for s in range(a.seeds)]
    out = band(ours, ref)
    out.update({"experiment": "8-node MLP 784-100-10 push-pull, %s" % ("4 800 / 2 000 samples" if a.small else "60 000 / 10 000 samples"),
                "ours": "engine=native, default (fp32-equivalent) kernels" if not a.train_impl else a.train_impl,
                "reference": "unmodified gossipy (baseline/_ref), device=cuda", "ours_curves": ours, "ref_curves": ref})
    print(json.dumps(out))
    return 0 if out["inside_band"] else 1


if __name__ == "__main__":
    sys.exit(main())
