"""Worker of tests/test_multirank.py: runs small simulations, single- or multi-rank, and prints one
JSON line with per-round metrics, message counters and a checksum of every node's model."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def build(kind, device):
    import gossipy_b200 as g
    from gossipy_b200.core import AntiEntropyProtocol, CreateModelMode, StaticP2PNetwork, UniformDelay
    from gossipy_b200.data import DataDispatcher, synthetic
    from gossipy_b200.data.handler import ClassificationDataHandler
    from gossipy_b200.model.handler import (LimitedMergeTMH, PartitionedTMH, PegasosHandler,
                                            TorchModelHandler, WeightedTMH)
    from gossipy_b200.model.nn import AdaLine, LogisticRegression, TorchMLP
    from gossipy_b200.model.sampling import TorchModelPartition
    from gossipy_b200.node import All2AllGossipNode, GossipNode, PartitioningBasedNode
    from gossipy_b200.simul import All2AllGossipSimulator, GossipSimulator, SimulationReport
    from gossipy_b200.core import UniformMixing
    g.GlobalSettings().set_device(device)
    g.set_seed(11)
    start_args = ()
    if kind == "pegasos":     # BASELINE config 1: main_ormandi_2013 shape, 8 nodes
        (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(800, 200)
        ytr, yte = 2 * ytr - 1, 2 * yte - 1
        disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=8, eval_on_user=False)
        proto = PegasosHandler(AdaLine(57), 0.01, CreateModelMode.MERGE_UPDATE)
        nodes = GossipNode.generate(disp, StaticP2PNetwork(8), proto, 20, False)
        sim = GossipSimulator(nodes, disp, 20, AntiEntropyProtocol.PUSH, drop_prob=.1, online_prob=.8,
                              delay=UniformDelay(0, 3), sampling_eval=.5)
    elif kind in ("bank_pegasos", "bank_adaline_pushpull", "bank_passthrough", "bank_cacheneigh"):
        # the banked engine (engine/bank.py): one node per few samples, many nodes per launch; several ranks push
        # snapshots into the receiver rank's slot bank
        from gossipy_b200.model.handler import AdaLineHandler
        (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(90, 120)
        ytr, yte = 2 * ytr - 1, 2 * yte - 1
        n = 45 if kind == "bank_pegasos" else 30
        disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=n, eval_on_user=False)
        if kind in ("bank_pegasos", "bank_passthrough", "bank_cacheneigh"):
            proto = PegasosHandler(AdaLine(57), 0.01, CreateModelMode.MERGE_UPDATE)
            kws, prt_ = dict(drop_prob=.1, online_prob=.8, delay=UniformDelay(0, 3), sampling_eval=.3), AntiEntropyProtocol.PUSH
        else:
            proto = AdaLineHandler(AdaLine(57), 0.01, CreateModelMode.UPDATE_MERGE)
            kws, prt_ = dict(delay=UniformDelay(0, 2)), AntiEntropyProtocol.PUSH_PULL
        if kind == "bank_passthrough":       # degree-aware pass-through (ring + hub: unequal degrees)
            from gossipy_b200.node import PassThroughNode
            A = np.zeros((n, n), dtype=int)
            for i in range(n):
                A[i, (i + 1) % n] = A[(i + 1) % n, i] = 1
                if i % 3 == 0 and i:
                    A[i, 0] = A[0, i] = 1
            nodes = PassThroughNode.generate(disp, StaticP2PNetwork(n, A), proto, 10, True)
        elif kind == "bank_cacheneigh":      # one cache slot per neighbour, consumed at send time
            from gossipy_b200.node import CacheNeighNode
            nodes = CacheNeighNode.generate(disp, StaticP2PNetwork(n), proto, 10, True)
        else:
            nodes = GossipNode.generate(disp, StaticP2PNetwork(n), proto, 10, kind == "bank_pegasos")
        sim = GossipSimulator(nodes, disp, 10, prt_, **kws)
        sim.engine = "native"
        sim.batched = True
    elif kind == "cnn_pushpull":
        # a generic (autograd) model: conv + BatchNorm (one rank on a GPU: steps replayed from CUDA graphs, channels-last rows;
        # several ranks: eager steps, plain rows)
        from gossipy_b200.model.nn import TorchModel

        class SmallCNN(TorchModel):
            def __init__(self):
                super().__init__()
                self.c1 = torch.nn.Conv2d(1, 4, 3, padding=1)
                self.bn = torch.nn.BatchNorm2d(4)
                self.fc = torch.nn.Linear(4 * 4 * 4, 3)

            def forward(self, x):
                x = torch.nn.functional.max_pool2d(torch.relu(self.bn(self.c1(x))), 2)
                return self.fc(x.flatten(1))

            def init_weights(self):
                pass

            def __str__(self):
                return "SmallCNN"
        gen = torch.Generator().manual_seed(3)
        yall = torch.randint(0, 3, (420,), generator=gen)
        proto_img = torch.randn(3, 1, 8, 8, generator=gen)
        Xall = torch.sigmoid(1.5 * proto_img[yall] + torch.randn(420, 1, 8, 8, generator=gen))
        disp = DataDispatcher(ClassificationDataHandler(Xall[:320], yall[:320], Xall[320:], yall[320:]), n=4, eval_on_user=False)
        torch.manual_seed(5)
        proto = TorchModelHandler(SmallCNN(), torch.optim.SGD, {"lr": .05, "momentum": .9}, torch.nn.CrossEntropyLoss(),
                                  batch_size=16)
        nodes = GossipNode.generate(disp, StaticP2PNetwork(4), proto, 10, True)
        sim = GossipSimulator(nodes, disp, 10, AntiEntropyProtocol.PUSH_PULL)
        sim.engine = "native"
    elif kind == "mlp_pushpull":
        (Xtr, ytr), (Xte, yte) = synthetic.mnist_like(640, 200)
        disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=4, eval_on_user=False)
        proto = TorchModelHandler(TorchMLP(784, 10, (100,)), torch.optim.SGD, {"lr": .1},
                                  torch.nn.CrossEntropyLoss(), batch_size=32)
        nodes = GossipNode.generate(disp, StaticP2PNetwork(4), proto, 10, True)
        sim = GossipSimulator(nodes, disp, 10, AntiEntropyProtocol.PUSH_PULL)
    elif kind == "limited_pull":
        (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(600, 200)
        disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=6, eval_on_user=True)
        proto = LimitedMergeTMH(LogisticRegression(57, 2), torch.optim.SGD, {"lr": 1., "weight_decay": .001},
                                torch.nn.CrossEntropyLoss(), batch_size=16, age_diff_threshold=1)
        nodes = GossipNode.generate(disp, StaticP2PNetwork(6), proto, 10, True)
        sim = GossipSimulator(nodes, disp, 10, AntiEntropyProtocol.PULL, delay=UniformDelay(0, 2))
    elif kind == "partitioned":
        (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(600, 200)
        disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=6, eval_on_user=False)
        net = LogisticRegression(57, 2)
        proto = PartitionedTMH(net, TorchModelPartition(net, 4), torch.optim.SGD, {"lr": 1., "weight_decay": .001},
                               torch.nn.CrossEntropyLoss(), batch_size=16, create_model_mode=CreateModelMode.UPDATE)
        nodes = PartitioningBasedNode.generate(disp, StaticP2PNetwork(6), proto, 10, True)
        sim = GossipSimulator(nodes, disp, 10, AntiEntropyProtocol.PUSH)
    elif kind == "all2all":
        (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(600, 200)
        disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=4, eval_on_user=False)
        net = StaticP2PNetwork(4)
        proto = WeightedTMH(LogisticRegression(57, 2), torch.optim.SGD, {"lr": .1, "weight_decay": .01},
                            torch.nn.CrossEntropyLoss(), batch_size=16)
        nodes = All2AllGossipNode.generate(disp, net, proto, 10, True)
        sim = All2AllGossipSimulator(nodes, disp, 10, AntiEntropyProtocol.PUSH)
        start_args = (UniformMixing(net),)
    elif kind == "all2all_sync":      # synchronous D-PSGD rounds: one all-reduce (NVLS / P2P / gloo) per round
        (Xtr, ytr), (Xte, yte) = synthetic.mnist_like(640, 200)
        disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=4, eval_on_user=False)
        net = StaticP2PNetwork(4)
        proto = WeightedTMH(TorchMLP(784, 10, (100,)), torch.optim.SGD, {"lr": .1},
                            torch.nn.CrossEntropyLoss(), batch_size=32)
        nodes = All2AllGossipNode.generate(disp, net, proto, 10, True)
        sim = All2AllGossipSimulator(nodes, disp, 10, AntiEntropyProtocol.PUSH)
        start_args = (UniformMixing(net),)
        sim._mr_kwargs = {"synchronous": True}
    elif kind == "x_all2all":
        # asynchronous all-to-all (cached neighbourhood, k-way merge on timeout) through the C++ executor; snapshots
        # shared by the pushes of a timeout, read by several ranks
        (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(600, 200)
        disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=6, eval_on_user=False)
        net = StaticP2PNetwork(6)
        proto = WeightedTMH(LogisticRegression(57, 2), torch.optim.SGD, {"lr": .1, "weight_decay": .001},
                            torch.nn.CrossEntropyLoss(), batch_size=16, create_model_mode=CreateModelMode.MERGE_UPDATE)
        nodes = All2AllGossipNode.generate(disp, net, proto, 10, False)
        sim = All2AllGossipSimulator(nodes, disp, 10, AntiEntropyProtocol.PUSH, drop_prob=.1, delay=UniformDelay(0, 3))
        sim.engine = "native"
        sim.native_executor = True
        start_args = (UniformMixing(net),)
    elif kind in ("x_part_mlp", "x_part_logreg", "x_part_update"):
        # partitioned models through the C++ executor: keyed partition draws, ages per partition, segment merges
        if kind == "x_part_mlp":
            (Xtr, ytr), (Xte, yte) = synthetic.mnist_like(640, 200)
            n, bs, net, prt_ = 4, 32, TorchMLP(784, 10, (100,)), AntiEntropyProtocol.PUSH_PULL
        else:
            (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(600, 200)
            n, bs, net, prt_ = 6, 16, LogisticRegression(57, 2), AntiEntropyProtocol.PUSH
        disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=n, eval_on_user=False)
        proto = PartitionedTMH(net, TorchModelPartition(net, 4), torch.optim.SGD, {"lr": .5, "weight_decay": .001},
                               torch.nn.CrossEntropyLoss(), batch_size=bs,
                               create_model_mode=CreateModelMode.UPDATE if kind == "x_part_update" else CreateModelMode.MERGE_UPDATE)
        nodes = PartitioningBasedNode.generate(disp, StaticP2PNetwork(n), proto, 10, True)
        sim = GossipSimulator(nodes, disp, 10, prt_, delay=UniformDelay(0, 2))
        sim.engine = "native"
        sim.native_executor = True
    elif kind in ("x_mlp_pushpull", "x_limited_push", "x_update_pull", "x_update_merge", "x_passthrough", "x_sampled", "x_sampled_update",
                  "x_cacheneigh", "x_momentum"):
        # native engine + the C++ executor (csrc/exec): one executor per rank over the same event list
        if kind == "x_mlp_pushpull":
            (Xtr, ytr), (Xte, yte) = synthetic.mnist_like(640, 200)
            n, bs, net, cls, kwh, proto_, kws = 4, 32, TorchMLP(784, 10, (100,)), TorchModelHandler, {}, AntiEntropyProtocol.PUSH_PULL, {}
        elif kind == "x_limited_push":
            (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(600, 200)
            n, bs, net, cls, kwh = 6, 16, LogisticRegression(57, 2), LimitedMergeTMH, {"age_diff_threshold": 2}
            proto_, kws = AntiEntropyProtocol.PUSH, dict(drop_prob=.1, online_prob=.8, delay=UniformDelay(0, 2), sampling_eval=.5)
        elif kind in ("x_sampled", "x_sampled_update"):
            from gossipy_b200.model.handler import SamplingTMH
            from gossipy_b200.node import SamplingBasedNode
            (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(600, 200)
            n, bs, net, cls, kwh = 6, 16, LogisticRegression(57, 2), (lambda *a, **k: SamplingTMH(.3, *a, **k)), {}
            if kind == "x_sampled_update":
                kwh = {"create_model_mode": CreateModelMode.UPDATE}
            proto_, kws = AntiEntropyProtocol.PUSH_PULL, dict(delay=UniformDelay(0, 2))
        elif kind == "x_momentum":
            (Xtr, ytr), (Xte, yte) = synthetic.mnist_like(480, 200)
            n, bs, net, cls, kwh = 5, 32, TorchMLP(784, 10, (100,)), TorchModelHandler, {}
            proto_, kws = AntiEntropyProtocol.PUSH_PULL, dict(delay=UniformDelay(0, 2))
        elif kind == "x_cacheneigh":
            (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(600, 200)
            n, bs, net, cls, kwh = 6, 16, LogisticRegression(57, 2), TorchModelHandler, {}
            proto_, kws = AntiEntropyProtocol.PUSH_PULL, dict(delay=UniformDelay(0, 2), drop_prob=.1)
        elif kind == "x_passthrough":
            (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(600, 200)
            n, bs, net, cls, kwh = 6, 16, LogisticRegression(57, 2), TorchModelHandler, {}
            proto_, kws = AntiEntropyProtocol.PUSH_PULL, dict(delay=UniformDelay(0, 2))
        elif kind == "x_update_merge":
            (Xtr, ytr), (Xte, yte) = synthetic.mnist_like(480, 200)
            n, bs, net, cls = 5, 32, TorchMLP(784, 10, (100,)), TorchModelHandler
            kwh, proto_, kws = {"create_model_mode": CreateModelMode.UPDATE_MERGE}, AntiEntropyProtocol.PUSH_PULL, dict(delay=UniformDelay(0, 2))
        else:
            (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(500, 200)
            n, bs, net, cls = 5, 16, LogisticRegression(57, 2), TorchModelHandler
            kwh, proto_, kws = {"create_model_mode": CreateModelMode.UPDATE}, AntiEntropyProtocol.PULL, dict(delay=UniformDelay(0, 3))
        disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=n, eval_on_user=False)
        opt_kw = {"lr": .05, "momentum": .9} if kind == "x_momentum" else {"lr": .1, "weight_decay": .001}
        proto = cls(net, torch.optim.SGD, opt_kw, torch.nn.CrossEntropyLoss(), batch_size=bs, **kwh)
        if kind == "x_passthrough":
            from gossipy_b200.node import PassThroughNode
            A = np.zeros((n, n), dtype=int)
            for i in range(n):
                A[i, (i + 1) % n] = A[(i + 1) % n, i] = 1
                if i > 1:
                    A[i, 0] = A[0, i] = 1
            nodes = PassThroughNode.generate(disp, StaticP2PNetwork(n, A), proto, 10, True)
        elif kind in ("x_sampled", "x_sampled_update"):
            nodes = SamplingBasedNode.generate(disp, StaticP2PNetwork(n), proto, 10, True)
        elif kind == "x_cacheneigh":
            from gossipy_b200.node import CacheNeighNode
            nodes = CacheNeighNode.generate(disp, StaticP2PNetwork(n), proto, 10, True)
        else:
            nodes = GossipNode.generate(disp, StaticP2PNetwork(n), proto, 10, kind != "x_limited_push")
        sim = GossipSimulator(nodes, disp, 10, proto_, **kws)
        sim.engine = "native"
        sim.native_executor = True
    elif kind in ("pens", "pens_native"):              # performance-based neighbour selection: data-dependent top-m, then step 2
        from gossipy_b200.node import PENSNode
        (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(600, 200)
        disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=6, eval_on_user=False)
        proto = TorchModelHandler(LogisticRegression(57, 2), torch.optim.SGD, {"lr": .5},
                                  torch.nn.CrossEntropyLoss(), batch_size=16,
                                  create_model_mode=CreateModelMode.MERGE_UPDATE)
        nodes = PENSNode.generate(disp, StaticP2PNetwork(6), proto, 10, True, n_sampled=2, m_top=1, step1_rounds=5)
        sim = GossipSimulator(nodes, disp, 10, AntiEntropyProtocol.PUSH)
        if kind == "pens_native":      # C++ control plane: step switch and restricted peer lists in the scheduler
            sim.engine = "native"
    else:
        raise ValueError(kind)
    sim.progress = False
    rep = SimulationReport()
    sim.add_receiver(rep)
    return sim, rep, start_args


def run(kind, device, rounds):
    import gossipy_b200 as g
    from gossipy_b200.parallel import runtime as prt
    sim, rep, start_args = build(kind, device)
    if os.environ.get("MR_PLACEMENT") and prt.active():     # a non-default node -> rank map, installed before init_nodes
        n, w = sim.n_nodes, prt.world()
        pl = {"round_robin": prt.Placement.round_robin(n, w),
              "by_load": prt.Placement.by_load([1 + (i * 7) % 5 for i in range(n)], w)}[os.environ["MR_PLACEMENT"]]
        prt.set_num_nodes(n, pl)
    sim.init_nodes(seed=5)
    if os.environ.get("MR_PLACEMENT") and prt.active():
        assert prt.placement() == pl, (prt.placement(), pl)
    if os.environ.get("MR_METRICS_EVERY"):       # exchange the evaluation results every k rounds instead of every round
        sim.metrics_sync_every = int(os.environ["MR_METRICS_EVERY"])
    if os.environ.get("MR_CHECKPOINT"):
        # interrupted run: half of the rounds, checkpoint (every rank writes its own file), reload, resume
        import tempfile
        first = max(1, rounds // 2)
        sim.start(*start_args, first, **getattr(sim, "_mr_kwargs", {}))
        path = os.path.join(tempfile.gettempdir(), "gb200_mr_ckpt_%d_%d.pkl" % (os.getpid(), prt.rank() if prt.active() else 0))
        sim.save(path)
        g.CACHE.clear()
        sim = type(sim).load(path)
        os.remove(path)
        rep = [r for r in sim._receivers if type(r).__name__ == "SimulationReport"][0]
        sim.start(*start_args, rounds - first, resume=True)
    else:
        sim.start(*start_args, rounds, **getattr(sim, "_mr_kwargs", {}))
    if device.startswith("cuda"):
        torch.cuda.synchronize()
    sums = {}
    for i, node in sim.nodes.items():
        h = node.model_handler
        if h._mine():
            r = h.row.detach().double().cpu()
            sums[i] = [float(r.sum()), float((r * r).sum())]
    if prt.active():
        import torch.distributed as dist
        allsums = [None] * prt.world()
        dist.all_gather_object(allsums, sums)
        sums = {k: v for d in allsums for k, v in d.items()}
    ages = {i: np.asarray(n.model_handler.n_updates).tolist() for i, n in sim.nodes.items()}
    return {"glob": rep.get_evaluation(False), "loc": rep.get_evaluation(True), "sent": rep._sent_messages,
            "failed": rep._failed_messages, "size": rep._total_size,
            "sums": {str(k): sums[k] for k in sorted(sums)}, "ages": {str(k): ages[k] for k in sorted(ages)},
            "cache_left": len(g.CACHE),
            "best": {str(i): getattr(n, "best_nodes", None) for i, n in sim.nodes.items()},
            "cpp_executor": "_stream_exec" in sim.__dict__, "banked": "_bank" in sim.__dict__}


def main():
    kinds = sys.argv[1].split(",")
    device = sys.argv[2]
    rounds = int(sys.argv[3])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        if device.startswith("cuda"):
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            device = "cuda:%d" % torch.cuda.current_device()
        dist.init_process_group("nccl" if device.startswith("cuda") else "gloo")
    import gossipy_b200 as g
    g.LOG.setLevel(50)
    out = {}
    for kind in kinds:
        from gossipy_b200.parallel import runtime as prt
        g.GlobalSettings().set_device(device)
        if world > 1:
            prt.init(rank, world)
        out[kind] = run(kind, device, rounds)
        g.CACHE.clear()
        if world > 1:
            prt.shutdown()
    if rank == 0:
        print("RESULT " + json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
