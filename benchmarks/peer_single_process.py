#!/usr/bin/env python
"""Peer kernels in ONE process with two (or more) visible GPUs -- the form ncu can capture (never wrap a multi-rank
launch in ncu): the fused pair merge pulling its source row out of cuda:1's HBM, the k-way merge over all other GPUs.

    ncu --set full --clock-control none -k regex:merge_ -c 6 -o gpurun_out/ncu_peer python benchmarks/peer_single_process.py
"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gossipy_b200.ops.native import native

n_dev = torch.cuda.device_count()
assert n_dev >= 2, "needs two GPUs"
torch.cuda.set_device(0)
nat = native()
n = int(os.environ.get("PEER_FLOATS", 1 << 26))
dst = torch.randn(n, device="cuda:0")
srcs = [torch.randn(n, device="cuda:%d" % d) for d in range(1, n_dev)]
probe = torch.empty(16, device="cuda:0")
for s in srcs:                      # kernels on cuda:0 dereference the other GPUs' memory directly
    assert torch.cuda.can_device_access_peer(0, s.device.index)
    nat.enable_peer_access(0, s.device.index)
    probe.copy_(s[:16])
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")


def timed(fn, iters=5):
    ts = []
    for it in range(iters + 2):
        flush.zero_(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if it >= 2:
            ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


want = .5 * dst + .5 * srcs[0].to("cuda:0")
ms = timed(lambda: nat.merge_pair(dst, srcs[0], .5, .5, 0, n, None), iters=1)
print(json.dumps({"op": "merge_pair(src on cuda:1), single process", "floats": n, "us": ms * 1e3, "nvlink_gbs": 4.0 * n / ms / 1e6}))
d2 = torch.randn(n, device="cuda:0"); w2 = .5 * d2 + .5 * srcs[0].to("cuda:0")
nat.merge_pair(d2, srcs[0], .5, .5, 0, n, None); torch.cuda.synchronize()
print(json.dumps({"check": "max |err| vs torch", "value": float((d2 - w2).abs().max())}))
for size in (1 << 22, 1 << 24, 1 << 26):
    if size <= n:
        ms = timed(lambda: nat.merge_pair(dst[:size], srcs[0][:size], .5, .5, 0, size, None))
        print(json.dumps({"op": "merge_pair(src on cuda:1)", "floats": size, "us": ms * 1e3, "nvlink_gbs": 4.0 * size / ms / 1e6}))
k = len(srcs)
w = [1.0 / (k + 1)] * (k + 1)
m = min(n, 1 << 24)
ms = timed(lambda: nat.merge_kway(dst[:m], [s[:m] for s in srcs], w, None))
print(json.dumps({"op": "merge_kway(%d peers)" % k, "floats": m, "us": ms * 1e3, "nvlink_gbs_in": 4.0 * m * k / ms / 1e6}))
