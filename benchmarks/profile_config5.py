#!/usr/bin/env python
"""Where does a round of BASELINE config 5 (ResNet-20, 8 nodes, token account) go?  torch.profiler over a few rounds:
wall time per round, summed device time per kernel family, host-side top entries.

    python benchmarks/profile_config5.py [--rounds 3] [--warmup 8] > gpurun_out/profile_config5.txt
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=8)
    a = ap.parse_args()
    import gossipy_b200 as g
    from gossipy_b200.core import AntiEntropyProtocol as AEP, StaticP2PNetwork
    from gossipy_b200.data import DataDispatcher, synthetic
    from gossipy_b200.data.handler import ClassificationDataHandler
    from gossipy_b200.flow_control import RandomizedTokenAccount
    from gossipy_b200.model.handler import TorchModelHandler
    from gossipy_b200.models import ResNet20
    from gossipy_b200.node import GossipNode
    from gossipy_b200.simul import TokenizedGossipSimulator
    g.GlobalSettings().set_device("cuda:0")
    g.LOG.setLevel(50)
    g.set_seed(98765)
    (Xtr, ytr), (Xte, yte) = synthetic.images_like("cifar10", n_train=8000, n_test=1000)
    disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=8, eval_on_user=False)
    proto = TorchModelHandler(ResNet20(10), torch.optim.SGD, {"lr": .05, "momentum": .9, "weight_decay": 1e-4},
                              torch.nn.functional.cross_entropy, batch_size=64, local_epochs=1)
    nodes = GossipNode.generate(disp, StaticP2PNetwork(8), proto, 100, True)
    sim = TokenizedGossipSimulator(nodes, disp, RandomizedTokenAccount(C=4, A=2), lambda a_, b_, m: 1, 100, AEP.PUSH, sampling_eval=0.)
    sim.native_utility = 1
    sim.progress = False
    sim.engine = "native"
    sim.init_nodes(seed=42)
    sim.start(a.warmup)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim.start(a.rounds, resume=True)
    torch.cuda.synchronize()
    print("wall ms per round (no profiler): %.1f" % ((time.perf_counter() - t0) / a.rounds * 1e3))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        t0 = time.perf_counter()
        sim.start(a.rounds, resume=True)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    print("wall ms per round (under the profiler): %.1f" % (wall / a.rounds * 1e3))
    ka = prof.key_averages()
    print(ka.table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
    print(ka.table(sort_by="self_cpu_time_total", row_limit=18, max_name_column_width=60))


if __name__ == "__main__":
    main()
