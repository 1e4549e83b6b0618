#!/usr/bin/env python
"""The five benchmark configurations named in BASELINE.json, on this framework.

    python benchmarks/baseline_configs.py --config 1..5 [--rounds R] [--scale S]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/baseline_configs.py --config 4

1  main_ormandi_2013   Pegasos linear model, 8 nodes (the CPU world_size=2 plumbing config)
2  main_hegedus_2021   2-layer MLP push-pull, 8 nodes = 8 GPUs, MNIST-shape non-IID  (== bench.py)
3  main_all2all        8-node all-to-all averaging, NVLS multicast all-reduce
4  main_giaretta_2019  partitioned-model merge, 64 nodes over 8 GPUs (8 nodes per GPU)
5  main_onoszko_2021   ResNet-20 on CIFAR-shape data, 8 nodes, TokenAccount flow control
6  weak scaling        config 2 with 8 nodes PER GPU (8 x world nodes, 7 500 samples each): aggregate node-rounds/s
7  main_all2all as the reference runs it: ASYNCHRONOUS all-to-all (nodes fire at their own ticks, cache their neighbours'
   models, merge them with the mixing weights on timeout) -- the C++ executor's all-to-all mode; `--no-executor` = per-event

Rank 0 prints one JSON line: rounds/s (device time, max over ranks), metric curve tail, message counters.
``--scale`` shrinks the data sets (1.0 = the sizes named above)."""
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
    ap.add_argument("--config", type=int, required=True)
    ap.add_argument("--rounds", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    ap.add_argument("--engine", default="native")
    ap.add_argument("--no-executor", action="store_true", help="per-event Python executor instead of the C++ one")
    ap.add_argument("--metrics-every", type=int, default=1,
                    help="several ranks: exchange the evaluation results every k rounds (sim.metrics_sync_every)")
    a = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    dev = a.device
    if dev.startswith("cuda"):
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dev = "cuda:%d" % torch.cuda.current_device()
    import gossipy_b200 as g
    from gossipy_b200.core import (AntiEntropyProtocol as AEP, CreateModelMode as CMM, StaticP2PNetwork, UniformDelay,
                                   UniformMixing)
    from gossipy_b200.data import AssignmentHandler, DataDispatcher, synthetic
    from gossipy_b200.data.handler import ClassificationDataHandler
    from gossipy_b200.flow_control import RandomizedTokenAccount
    from gossipy_b200.model.handler import PartitionedTMH, PegasosHandler, TorchModelHandler, WeightedTMH
    from gossipy_b200.model.nn import AdaLine, TorchMLP
    from gossipy_b200.model.sampling import TorchModelPartition
    from gossipy_b200.node import All2AllGossipNode, GossipNode, PartitioningBasedNode
    from gossipy_b200.simul import (All2AllGossipSimulator, GossipSimulator, SimulationReport,
                                    TokenizedGossipSimulator)
    g.LOG.setLevel(50)
    g.GlobalSettings().set_device(dev)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if dev.startswith("cuda") else "gloo")
        from gossipy_b200.parallel import runtime as prt
        prt.init(rank, world)
    g.set_seed(98765)
    sc = a.scale
    start_extra, start_kw, desc = (), {}, ""
    if a.config == 1:
        (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(int(4141 * sc), int(460 * sc))
        disp = DataDispatcher(ClassificationDataHandler(Xtr, 2 * ytr - 1, Xte, 2 * yte - 1), n=8, eval_on_user=False)
        proto = PegasosHandler(AdaLine(57), .01, CMM.MERGE_UPDATE)
        nodes = GossipNode.generate(disp, StaticP2PNetwork(8), proto, 100, False)
        sim = GossipSimulator(nodes, disp, 100, AEP.PUSH, drop_prob=.1, online_prob=.2, delay=UniformDelay(0, 10),
                              sampling_eval=.1)
        desc = "Pegasos, 8 nodes, PUSH, churn"
    elif a.config in (2, 3, 7):
        (Xtr, ytr), (Xte, yte) = synthetic.mnist_like(int(60000 * sc), int(10000 * sc))
        disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=8, eval_on_user=False, auto_assign=False)
        disp.set_assignments(AssignmentHandler(42).label_pathological_skew(ytr, 8, 2), None)
        net = StaticP2PNetwork(8)
        if a.config == 2:
            proto = TorchModelHandler(TorchMLP(784, 10, (100,)), torch.optim.SGD, {"lr": .1},
                                      torch.nn.CrossEntropyLoss(), batch_size=32)
            nodes = GossipNode.generate(disp, net, proto, 100, True)
            sim = GossipSimulator(nodes, disp, 100, AEP.PUSH_PULL)
            desc = "MLP 784-100-10 push-pull, 8 nodes, non-IID"
        else:
            proto = WeightedTMH(TorchMLP(784, 10, (100,)), torch.optim.SGD, {"lr": .1},
                                torch.nn.CrossEntropyLoss(), batch_size=32)
            nodes = All2AllGossipNode.generate(disp, net, proto, 100, True)
            sim = All2AllGossipSimulator(nodes, disp, 100, AEP.PUSH)
            if a.config == 3:
                start_extra, start_kw = (UniformMixing(net),), {"synchronous": True}
                desc = "MLP all-to-all averaging, 8 nodes, one-shot all-reduce per round"
            else:
                start_extra = (UniformMixing(net),)
                desc = "MLP asynchronous all-to-all (reference semantics), 8 nodes, %s" % (
                    "per-event Python executor" if a.no_executor else "C++ executor")
    elif a.config == 4:
        n_nodes = 64
        (Xtr, ytr), (Xte, yte) = synthetic.mnist_like(int(60000 * sc), int(10000 * sc))
        disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=n_nodes, eval_on_user=False)
        import numpy as np
        rng = np.random.default_rng(7)
        A = np.zeros((n_nodes, n_nodes), dtype=int)
        for i in range(n_nodes):                      # 10-regular circulant + a few random chords
            for k in range(1, 6):
                A[i, (i + k) % n_nodes] = A[(i + k) % n_nodes, i] = 1
        net = StaticP2PNetwork(n_nodes, A)
        mlp = TorchMLP(784, 10, (100,))
        proto = PartitionedTMH(mlp, TorchModelPartition(mlp, 4), torch.optim.SGD, {"lr": .5, "weight_decay": .001},
                               torch.nn.CrossEntropyLoss(), batch_size=32, create_model_mode=CMM.MERGE_UPDATE)
        nodes = PartitioningBasedNode.generate(disp, net, proto, 100, True)
        sim = GossipSimulator(nodes, disp, 100, AEP.PUSH, delay=UniformDelay(0, 10), sampling_eval=.25)
        desc = "partitioned MLP merge (4 parts), 64 nodes, PUSH"
    elif a.config == 5:
        from gossipy_b200.models import ResNet20
        (Xtr, ytr), (Xte, yte) = synthetic.images_like("cifar10", n_train=int(8000 * sc), n_test=int(1000 * sc))
        disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=8, eval_on_user=False)
        proto = TorchModelHandler(ResNet20(10), torch.optim.SGD, {"lr": .05, "momentum": .9, "weight_decay": 1e-4},
                                  torch.nn.functional.cross_entropy, batch_size=64, local_epochs=1)
        nodes = GossipNode.generate(disp, StaticP2PNetwork(8), proto, 100, True)
        sim = TokenizedGossipSimulator(nodes, disp, RandomizedTokenAccount(C=4, A=2), lambda a_, b_, m: 1, 100,
                                       AEP.PUSH, sampling_eval=0.)
        sim.native_utility = 1
        desc = "ResNet-20, CIFAR-shape, 8 nodes, RandomizedTokenAccount(4,2)"
    elif a.config == 6:
        n_nodes = 8 * world
        (Xtr, ytr), (Xte, yte) = synthetic.mnist_like(int(7500 * n_nodes * sc), int(10000 * sc))
        disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=n_nodes, eval_on_user=False, auto_assign=False)
        disp.set_assignments(AssignmentHandler(42).label_pathological_skew(ytr, n_nodes, 2), None)
        proto = TorchModelHandler(TorchMLP(784, 10, (100,)), torch.optim.SGD, {"lr": .1},
                                  torch.nn.CrossEntropyLoss(), batch_size=32)
        nodes = GossipNode.generate(disp, StaticP2PNetwork(n_nodes), proto, 100, True)
        sim = GossipSimulator(nodes, disp, 100, AEP.PUSH_PULL)
        desc = "weak scaling: MLP 784-100-10 push-pull, %d nodes (8 per GPU), non-IID" % n_nodes
    else:
        raise SystemExit("config must be 1..6")
    sim.progress = False
    sim.engine = a.engine
    if a.no_executor:
        sim.native_executor = False
    sim.metrics_sync_every = a.metrics_every
    rep = SimulationReport()
    sim.add_receiver(rep)
    sim.init_nodes(seed=42)

    def run(rounds, resume):
        if dev.startswith("cuda"):
            torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        if start_kw.get("synchronous"):
            sim.start(*start_extra, rounds, **start_kw)
        else:
            sim.start(*start_extra, rounds, resume=resume)
        if dev.startswith("cuda"):
            torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        return time.perf_counter() - t0
    run(a.warmup, False)
    sec = run(a.rounds, True)
    if world > 1:
        t = torch.tensor([sec], dtype=torch.float64, device=dev if dev.startswith("cuda") else "cpu")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        sec = float(t.item())
    ev = rep.get_evaluation(False)
    if rank == 0:
        from gossipy_b200 import ops
        print(json.dumps({"config": a.config, "what": desc, "n_gpus": world, "rounds": a.rounds,
                          "rounds_per_s": a.rounds / sec, "ms_per_round": sec / a.rounds * 1e3,
                          "node_rounds_per_s": a.rounds / sec * len(sim.nodes),
                          "timing": "host wall clock around synchronised device work, max over ranks",
                          "last_eval": {k: round(v, 4) for k, v in (ev[-1][1].items() if ev else [])},
                          "first_eval": {k: round(v, 4) for k, v in (ev[0][1].items() if ev else [])},
                          "sent": rep._sent_messages, "failed": rep._failed_messages, "size": rep._total_size,
                          "native_launches_rank0": ops.launch_count, "scale": sc, "engine": a.engine}))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
