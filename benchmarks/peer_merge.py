#!/usr/bin/env python
"""Peer-merge bandwidth: the fused merge kernel pulling a parameter vector out of ANOTHER GPU's HBM
over NVLink (BASELINE.json: "peer-merge GB/s vs 900 GB/s").

    python -m torch.distributed.run --nproc-per-node W --master-addr 127.0.0.1 benchmarks/peer_merge.py

Every rank r merges the row of rank (r+1) % W into its own row (dst = .5 dst + .5 peer): W concurrent
disjoint pairs, the pattern of a gossip round with one node per GPU.  Also measured: adopt (pure pull,
w_dst = 0), the k-way merge of all W-1 peers, and the one-shot all-reduce (NVLS multicast if the
system has it, else P2P pull).  Next to each, the BASELINE path that only calls NCCL for the same op:
ncclSend/ncclRecv of the row into a staging buffer + torch element-wise merge, and ``ncclAllReduce`` + scale.
Device time (CUDA events), max over ranks; bytes over NVLink per rank = 4 * n for the pair merge.  Rows larger than L2 are used for the bandwidth numbers (and an L2 flush
for the small ones)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gossipy_b200.ops.native import native  # noqa: E402

NVLINK_MEASURED, NVLINK_NOMINAL = 770.0, 900.0     # GB/s per direction per GPU (B200_PROFILING.md)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dev = torch.cuda.current_device()
    dist.init_process_group("nccl")
    nat = native()
    nmax = 1 << 28
    base = nat.ipc_alloc(nmax * 4 + 256)
    handles = [None] * world
    dist.all_gather_object(handles, nat.ipc_get_handle(base))
    bases = [base if r == rank else nat.ipc_open_handle(handles[r]) for r in range(world)]
    dist.barrier()
    rows = [nat.tensor_from_ptr(bases[r], [nmax], dev, False) for r in range(world)]
    rows[rank].normal_()
    mine = torch.randn(nmax, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize(); dist.barrier()

    def timed(fn, iters=5, do_flush=True):
        best = []
        for it in range(iters + 2):
            if do_flush:
                flush.zero_()
            torch.cuda.synchronize(); dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            if it >= 2:
                best.append(a.elapsed_time(b))
        t = torch.tensor([sorted(best)[len(best) // 2]], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = []
    peer = rows[(rank + 1) % world]
    for n in (79520, 1 << 20, 1 << 22, 1 << 24, 1 << 26, 1 << 28):
        ms = timed(lambda: nat.merge_pair(mine[:n], peer[:n], .5, .5, 0, n, None), do_flush=n < (1 << 26))
        gbs = 4.0 * n / ms / 1e6
        out.append({"op": "merge_pair(peer)", "floats": n, "us": ms * 1e3, "nvlink_gbs_per_gpu": gbs,
                    "frac_of_measured_770": gbs / NVLINK_MEASURED, "frac_of_nominal_900": gbs / NVLINK_NOMINAL,
                    "concurrent_pairs": world})
        ms = timed(lambda: nat.merge_pair(mine[:n], peer[:n], 0., 1., 0, n, None), do_flush=n < (1 << 26))
        out.append({"op": "adopt(peer)", "floats": n, "us": ms * 1e3, "nvlink_gbs_per_gpu": 4.0 * n / ms / 1e6})
    # baseline: NCCL point-to-point into a staging buffer, then the merge as torch element-wise ops
    try:
        stage = torch.empty(1 << 26, device="cuda")
        nxt, prv = (rank + 1) % world, (rank - 1) % world
        for n in (79520, 1 << 20, 1 << 22, 1 << 24, 1 << 26):
            def nccl_merge(n=n):
                ops_ = [dist.P2POp(dist.isend, mine[:n], prv), dist.P2POp(dist.irecv, stage[:n], nxt)]
                for req in dist.batch_isend_irecv(ops_):
                    req.wait()
                mine[:n].mul_(.5).add_(stage[:n], alpha=.5)
            ms = timed(nccl_merge, do_flush=n < (1 << 26))
            out.append({"op": "baseline: nccl send/recv + torch merge", "floats": n, "us": ms * 1e3,
                        "nvlink_gbs_per_gpu": 4.0 * n / ms / 1e6, "concurrent_pairs": world})
        del stage
    except Exception as exc:  # noqa: BLE001
        out.append({"op": "baseline: nccl send/recv + torch merge", "error": repr(exc)[:300]})
    if world > 2:
        for n in (79520, 1 << 24):
            srcs = [rows[r][:n] for r in range(world) if r != rank]
            w = [1.0 / world] * world
            ms = timed(lambda: nat.merge_kway(mine[:n], srcs, w, None), do_flush=True)
            out.append({"op": "merge_kway(%d peers)" % (world - 1), "floats": n, "us": ms * 1e3,
                        "nvlink_gbs_per_gpu": 4.0 * n * (world - 1) / ms / 1e6})
    # one-shot all-reduce (NVLS multicast when available)
    try:
        from gossipy_b200.parallel import runtime as prt
        from gossipy_b200.parallel.collectives import SymmetricAllReduce
        import gossipy_b200 as g
        g.GlobalSettings().set_device("cuda:%d" % dev)
        prt.init(rank, world)
        for n, force_p2p in ((79520, False), (79520, True), (1 << 24, False), (1 << 24, True)):
            coll = SymmetricAllReduce(n, torch.device("cuda", dev), use_multicast=False if force_p2p else None)
            coll.contribution.normal_()
            res = torch.empty(n, device="cuda")
            ms = timed(lambda: coll.mean_into(res, world), do_flush=True)
            ref = coll.contribution.clone()
            dist.all_reduce(ref)
            err = float((res - ref / world).abs().max())
            out.append({"op": "allreduce_mean(%s)" % coll.kind, "floats": n, "us": ms * 1e3, "max_err_vs_nccl": err,
                        "bytes_received_per_gpu": 4 * n if coll.kind == "nvls" else 4 * n * (world - 1)})
            if force_p2p:      # once per size: the NCCL call the one-shot kernel replaces
                def nccl_mean():
                    res.copy_(coll.contribution)
                    dist.all_reduce(res)
                    res.mul_(1.0 / world)
                ms = timed(nccl_mean, do_flush=True)
                out.append({"op": "baseline: ncclAllReduce + scale", "floats": n, "us": ms * 1e3})
    except Exception as exc:  # noqa: BLE001
        out.append({"op": "allreduce_mean", "error": repr(exc)[:300]})
    if rank == 0:
        for o in out:
            print(json.dumps(o))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
