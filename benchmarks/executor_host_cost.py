#!/usr/bin/env python
"""Host cost of the C++ executor's bookkeeping per event (no GPU needed).

The executor is built in its CPU mode and every launch callback is replaced by a no-op, so what is timed is what the
host does per event on any device: slot allocation, ages / counters / keys, the cross-rank generation bookkeeping, the
elision look-ahead -- plus one pybind call per launch where a GPU run has a kernel launch instead.

    python benchmarks/executor_host_cost.py [--nodes 64] [--rounds 300] [--protocol PUSH_PULL]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=64)
    ap.add_argument("--rounds", type=int, default=300)
    ap.add_argument("--protocol", default="PUSH_PULL", choices=["PUSH", "PULL", "PUSH_PULL"])
    a = ap.parse_args()
    import gossipy_b200 as g
    from gossipy_b200.core import AntiEntropyProtocol, CreateModelMode, StaticP2PNetwork
    from gossipy_b200.data import DataDispatcher, synthetic
    from gossipy_b200.data.handler import ClassificationDataHandler
    from gossipy_b200.engine.stream_exec import StreamExec, eligible
    from gossipy_b200.model.handler import TorchModelHandler
    from gossipy_b200.model.nn import LogisticRegression
    from gossipy_b200.node import GossipNode
    from gossipy_b200.simul import GossipSimulator
    g.LOG.setLevel(50)
    g.GlobalSettings().set_device("cpu")
    g.set_seed(1)                      # (the nodes' timeout offsets come from the host stream)
    n = a.nodes
    (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(20 * n, 50)
    disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=n, eval_on_user=False)
    proto = TorchModelHandler(LogisticRegression(57, 2), torch.optim.SGD, {"lr": .1}, torch.nn.CrossEntropyLoss(), batch_size=16,
                              create_model_mode=CreateModelMode.MERGE_UPDATE)
    nodes = GossipNode.generate(disp, StaticP2PNetwork(n), proto, 100, True)
    sim = GossipSimulator(nodes, disp, 100, getattr(AntiEntropyProtocol, a.protocol))
    sim.progress = False
    sim.engine = "native"
    sim.init_nodes(seed=1)
    assert eligible(sim) is None
    sch = sim._make_scheduler()
    sx = StreamExec(sim)
    noop = lambda *args: None          # noqa: E731
    sx.ex.set_callbacks(noop, noop, noop)
    events = [sch.run(1) for _ in range(a.rounds)]
    n_events = sum(int(e.shape[0]) for e in events)
    launches0 = int(sx.ex.launches)
    t_sched0 = time.perf_counter()
    sch2 = sim._make_scheduler()
    for _ in range(a.rounds):
        sch2.run(1)
    t_sched = time.perf_counter() - t_sched0
    t0 = time.perf_counter()
    for ev in events:
        sx._run(ev)
    dt = time.perf_counter() - t0
    launches = int(sx.ex.launches) - launches0
    print(json.dumps({"nodes": n, "protocol": a.protocol, "rounds": a.rounds, "events": n_events, "launch_callbacks": launches,
                      "executor_us_per_event": round(dt / n_events * 1e6, 3), "executor_us_per_round": round(dt / a.rounds * 1e6, 1),
                      "scheduler_us_per_round": round(t_sched / a.rounds * 1e6, 1),
                      "note": "CPU mode, launches replaced by no-op pybind callbacks"}))


if __name__ == "__main__":
    main()
