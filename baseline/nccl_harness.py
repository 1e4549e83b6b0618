#!/usr/bin/env python
"""The bar: gossip learning with the reference's semantics written the way a
competent multi-GPU PyTorch user would write it WITHOUT this framework's kernels -- one process per GPU,
``ncclSend`` / ``ncclRecv`` of the flat parameter vector into a staging buffer, the merge as torch element-wise ops,
local SGD as eager autograd over ``F.linear`` (cuBLAS) with ``torch.optim.SGD``, evaluation with torch ops on the
device; data and models stay resident in HBM (unlike the reference, which moves them host<->device every step).

Same experiment as ``bench.py`` (BASELINE.json config 2): 8 nodes, MLP 784-100-10, SGD lr 0.1, batch 32, one local
epoch per update, PUSH_PULL + MERGE_UPDATE (reference ``gossipy/simul.py:366-458``, ``model/handler.py:235-280``): in
every round each node, at its own tick, pushes its model to a random peer, which averages it into its own model,
trains one epoch and replies with the result; the initiator does the same with the reply.  Every node is evaluated on
the global test set every round.  The schedule is drawn from a shared seed, so every rank replays it and takes part
only in the events of the nodes it hosts (block placement, node i on rank i*W//8) -- disjoint exchanges overlap.

    python baseline/nccl_harness.py --steps 20 --warmup 3                                         # 1 GPU
    python -m torch.distributed.run --nproc-per-node W --master-addr 127.0.0.1 baseline/nccl_harness.py ...

Prints one JSON line (rank 0): rounds/s (device-timed with CUDA events, max over ranks) and the accuracy curve.
With ``--all2all`` it runs BASELINE config 3 instead: every round all nodes train one epoch and average their models
with ``ncclAllReduce`` (the NCCL form of the synchronous all-to-all round that ``nvls.cu`` does in one kernel).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import BATCH, D_H, D_IN, D_OUT, LR, METRIC, N_NODES, common_config, make_data, make_split  # noqa: E402


class Node:
    """One gossip node: flat fp32 parameter vector + views, its shard, an eager SGD optimizer."""

    def __init__(self, idx: int, X: torch.Tensor, y: torch.Tensor, gen: torch.Generator, dev):
        self.idx = idx
        P = D_H * D_IN + D_H + D_OUT * D_H + D_OUT
        self.flat = torch.zeros(P, device=dev)
        o = 0
        self.W1 = self.flat[o:o + D_H * D_IN].view(D_H, D_IN); o += D_H * D_IN
        self.b1 = self.flat[o:o + D_H]; o += D_H
        self.W2 = self.flat[o:o + D_OUT * D_H].view(D_OUT, D_H); o += D_OUT * D_H
        self.b2 = self.flat[o:o + D_OUT]
        for W, fan_in in ((self.W1, D_IN), (self.W2, D_H)):          # torch.nn.Linear's default init
            W.uniform_(-1.0 / fan_in ** .5, 1.0 / fan_in ** .5, generator=gen)
        self.b1.uniform_(-1.0 / D_IN ** .5, 1.0 / D_IN ** .5, generator=gen)
        self.b2.uniform_(-1.0 / D_H ** .5, 1.0 / D_H ** .5, generator=gen)
        self.params = [p.requires_grad_(True) for p in (self.W1, self.b1, self.W2, self.b2)]
        self.opt = torch.optim.SGD(self.params, lr=LR)
        self.X, self.y = X.to(dev), y.to(dev)
        self.gen = gen

    def forward(self, x):
        return F.linear(F.relu(F.linear(x, self.W1, self.b1)), self.W2, self.b2)

    def local_epoch(self):
        n = self.X.shape[0]
        perm = torch.randperm(n, device=self.X.device, generator=self.gen)
        for i in range(0, n, BATCH):
            idx = perm[i:i + BATCH]
            loss = F.cross_entropy(self.forward(self.X[idx]), self.y[idx])
            self.opt.zero_grad(set_to_none=True)
            loss.backward()
            self.opt.step()

    @torch.no_grad()
    def merge(self, other_flat: torch.Tensor):
        self.flat.detach().mul_(.5).add_(other_flat, alpha=.5)

    @torch.no_grad()
    def accuracy(self, Xte, yte):
        return (self.forward(Xte).argmax(1) == yte).float().mean()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--all2all", action="store_true")
    ap.add_argument("--gpus", type=int, default=None, help="ignored (world size comes from torchrun)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count())
    dev = torch.device("cuda", torch.cuda.current_device())
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    from gossipy_b200.data import AssignmentHandler
    Xtr, ytr, Xte, yte = make_data()
    parts = make_split(ytr)
    owner = [i * world // N_NODES for i in range(N_NODES)]
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    nodes = {i: Node(i, Xtr[parts[i]], ytr[parts[i]], gen, dev) for i in range(N_NODES) if owner[i] == rank}
    Xte_d, yte_d = Xte.to(dev), yte.to(dev)
    stage = torch.empty_like(next(iter(nodes.values())).flat) if nodes else torch.empty(1, device=dev)
    sched = np.random.RandomState(98765)
    offsets = sched.randint(0, 100, N_NODES)                      # each sync node fires at its own tick of the round
    for nd in nodes.values():                                     # init_nodes: one local update each (reference simul.py:341-355)
        nd.local_epoch()

    def deliver(src: int, dst: int):
        """Node ``src``'s model reaches node ``dst``: merge + one local epoch on dst."""
        rs, rd = owner[src], owner[dst]
        if rs == rd:
            if rd == rank:
                nodes[dst].merge(nodes[src].flat.detach())
                nodes[dst].local_epoch()
        elif rank == rs:
            dist.send(nodes[src].flat.detach(), dst=rd)
        elif rank == rd:
            dist.recv(stage, src=rs)
            nodes[dst].merge(stage)
            nodes[dst].local_epoch()

    accs = []

    def one_round():
        if args.all2all:
            for nd in nodes.values():
                nd.local_epoch()
            acc = torch.zeros_like(stage)
            for nd in nodes.values():
                acc.add_(nd.flat.detach())
            if world > 1:
                dist.all_reduce(acc)
            acc.mul_(1.0 / N_NODES)
            for nd in nodes.values():
                with torch.no_grad():
                    nd.flat.detach().copy_(acc)
        else:
            for i in np.argsort(offsets, kind="stable"):
                j = int(sched.choice([k for k in range(N_NODES) if k != i]))
                deliver(int(i), j)          # PUSH: j merges + updates
                deliver(j, int(i))          # ... and replies (PULL half): i merges + updates
        a = torch.zeros(N_NODES, device=dev)
        for i, nd in nodes.items():
            a[i] = nd.accuracy(Xte_d, yte_d)
        accs.append(a)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        one_round()
    sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        one_round()
    ev1.record()
    sync()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    curve = torch.stack(accs)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(curve)
    if rank == 0:
        value = args.steps / (float(ms) / 1e3)
        print(json.dumps({"metric": METRIC if not args.all2all else "synchronous all-to-all rounds/sec (8-node MLP, ncclAllReduce)",
                          "impl": "nccl+cublas harness (eager PyTorch)", "value": value, "unit": "rounds/s", "n_gpus": world,
                          "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": float(ms) / args.steps,
                          "dtype": "fp32", "data": "synthetic", "config": common_config(world),
                          "test_acc_by_round": [round(float(v), 4) for v in curve.mean(1)]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
