#!/usr/bin/env python
"""Micro-benchmarks of the hot ops (CUDA events, warm-up, L2 flush between iterations).

    python benchmarks/micro.py train|eval|merge|overlap|all [--impl cluster|tc8|tc8-tf32|tc3]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gossipy_b200 import ops  # noqa: E402

DIMS = (784, 100, 10)
P = 79510
PEAKS = {"hbm_gbs": 6477.4}
try:
    PEAKS.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
except Exception:
    pass

_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def timeit(fn, iters=10, warm=3, flush=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def problem(n=7500, seed=0):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(n, DIMS[0], generator=g).cuda()
    y = torch.randint(0, DIMS[2], (n,), generator=g).cuda()
    row = torch.zeros(79520, device="cuda")
    row[:P] = (torch.randn(P, generator=g) * 0.03).cuda()
    return X, y, row


def bench_train(impl):
    X, y, row = problem()
    steps = 235
    med, best = timeit(lambda: ops.mlp1_train(row, X, y, DIMS, 32, 1, 0.1, 0.0, 1234, impl=impl), iters=8)
    flops = steps * 32 * (2 * 784 * 100 * 2 + 2 * 100 * 10 * 3)
    print(json.dumps({"op": "mlp1_train", "impl": impl or "auto", "ms_per_update": med, "best_ms": best,
                      "us_per_sgd_step": med * 1e3 / steps, "gflops": flops / med / 1e6}))
    from gossipy_b200.ops.native import native as _nat
    med_s, best_s = timeit(lambda: _nat().mlp1_stage_debug(X, y, 32, 1, 1234), iters=8)
    print(json.dumps({"op": "mlp1_stage (device-side loader, 235 steps)", "us": med_s * 1e3, "best_us": best_s * 1e3}))
    Xs, ys_ = X[:320].contiguous(), y[:320].contiguous()      # 10 SGD steps: what a launch costs besides its steps
    med10, best10 = timeit(lambda: ops.mlp1_train(row, Xs, ys_, DIMS, 32, 1, 0.1, 0.0, 1234, impl=impl), iters=8)
    per_step = (med - med10) / (steps - 10)
    print(json.dumps({"op": "mlp1_train fixed cost", "impl": impl or "auto", "ms_10_steps": med10,
                      "us_per_step_marginal": per_step * 1e3, "fixed_us_per_launch": (med10 - 10 * per_step) * 1e3}))
    if impl == "tc3":   # per-phase cycle counters of the CTA-pair kernel (thread 0 of both CTAs); tc8: benchmarks/check_tc4.py
        from gossipy_b200.ops.native import native
        dbg = native().mlp1_train_tc_debug(row, X, y, DIMS, 32, 1, 0.1, 0.0, 1234, impl)
        raw = dbg[0, :17].tolist()
        raw1 = dbg[1, :17].tolist()          # CTA 1 of the pair (tc3)
        names = ["A+B issue fwd MMA (+wait X tile, +wait upd(s-1), request X^T)", "C wait fwd MMA + tmem_ld",
                 "C exchange (DSMEM + cluster barrier) + relu + hs", "D+E layer-2 fwd, softmax-CE",
                 "F backward + A2 operand", "G+H issue update MMA + W2/b update"]
        if impl == "tc3":     # 17 fine-grained stamps; the six coarse phases are sums of them
            fine = ["A+B issue fwd", "C wait fwd + ld", "C st.async issued", "C peer partials landed", "C relu + h stores",
                    "C fences + barrier", "D issue logits MMAs", "E wait logits", "E tmem_ld", "E softmax + dz stores + fence",
                    "E barrier (+labels)", "F issue dh/gw2 MMAs", "F wait", "F ld + relu mask + A2 stores", "F fences + barrier",
                    "G issue update MMAs", "H second-layer update"]
            print(json.dumps({"op": "mlp1_train_tc3 fine phase cycles per step (thread 0)",
                              "phases": {k: round(v) for k, v in zip(fine, raw)}, "sum": round(sum(raw))}))
            print(json.dumps({"op": "mlp1_train_tc3 fine phase cycles per step (thread 0 of CTA 1)",
                              "phases": {k: round(v) for k, v in zip(fine, raw1)}, "sum": round(sum(raw1))}))
            prof = [raw[0], raw[1], sum(raw[2:6]), sum(raw[6:11]), sum(raw[11:15]), sum(raw[15:17])]
        else:
            prof = raw[:6]
        print(json.dumps({"op": "mlp1_train_%s phase cycles per step (thread 0)" % (impl or "tc2"),
                          "phases": {k: round(v) for k, v in zip(names, prof)}, "sum": round(sum(prof))}))


def bench_eval():
    X, y, row = problem(10000)
    flops = 10000 * 2 * (784 * 100 + 100 * 10)
    for impl in ("simt", "tc", "tc-tf32"):
        ops.EVAL_IMPL = "tc" if impl == "tc-tf32" else impl
        ops.set_eval_tf32(impl == "tc-tf32")       # "tc": fp32-equivalent (3xTF32, the default); "tc-tf32": plain tf32
        med, best = timeit(lambda: ops.mlp1_eval(row, X, y, DIMS, 10), iters=10)
        print(json.dumps({"op": "mlp1_eval", "impl": impl, "ms": med, "best_ms": best, "tflops": flops / med / 1e9,
                          "x_read_gbs": 10000 * 784 * 4 / med / 1e6,
                          "frac_of_measured_hbm": 10000 * 784 * 4 / med / 1e6 / PEAKS["hbm_gbs"]}))
    ops.EVAL_IMPL = ""
    ops.set_eval_tf32(False)


def bench_merge():
    for n in (79520, 1 << 20, 1 << 24, 1 << 26, 1 << 28):
        d = torch.randn(n, device="cuda"); s = torch.randn(n, device="cuda")
        med, best = timeit(lambda: ops.merge_pair(d, s, .5, .5), iters=10, flush=n < (1 << 26))
        print(json.dumps({"op": "merge_pair(local)", "floats": n, "us": med * 1e3,
                          "algorithmic_gbs": 3 * n * 4 / med / 1e6,
                          "frac_of_measured_hbm": 3 * n * 4 / med / 1e6 / PEAKS["hbm_gbs"]}))
    d = torch.randn(1 << 26, device="cuda"); s = torch.randn(1 << 26, device="cuda")
    med, _ = timeit(lambda: ops.snapshot(d, s), iters=10, flush=False)
    print(json.dumps({"op": "snapshot(copy)", "floats": 1 << 26, "us": med * 1e3,
                      "algorithmic_gbs": 2 * (1 << 26) * 4 / med / 1e6}))


def bench_overlap(impl):
    """8 independent updates on 8 streams vs back to back on one stream."""
    probs = [problem(seed=i) for i in range(8)]
    streams = [torch.cuda.Stream() for _ in range(8)]

    def serial():
        for X, y, row in probs:
            ops.mlp1_train(row, X, y, DIMS, 32, 1, 0.1, 0.0, 1, impl=impl)

    def parallel():
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(cur)
        for s, (X, y, row) in zip(streams, probs):
            s.wait_event(ev)
            with torch.cuda.stream(s):
                ops.mlp1_train(row, X, y, DIMS, 32, 1, 0.1, 0.0, 1, impl=impl)
            e2 = torch.cuda.Event(); e2.record(s); cur.wait_event(e2)
    ms_s, _ = timeit(serial, iters=5)
    ms_p, _ = timeit(parallel, iters=5)
    print(json.dumps({"op": "8 updates", "impl": impl or "auto", "serial_ms": ms_s, "8_streams_ms": ms_p,
                      "overlap_speedup": ms_s / ms_p}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="all")
    ap.add_argument("--impl", default="")
    a = ap.parse_args()
    if a.what in ("train", "all"):
        bench_train(a.impl)
    if a.what in ("eval", "all"):
        bench_eval()
    if a.what in ("merge", "all"):
        bench_merge()
    if a.what in ("overlap", "all"):
        bench_overlap(a.impl)
