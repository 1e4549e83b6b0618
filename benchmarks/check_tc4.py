"""GPU bring-up / accuracy / timing of the fp32-equivalent training kernel (mlp1_train_tc4.cu).

    python benchmarks/check_tc4.py [impl ...]      (default: tc8 tc8-tf32 tc3 cluster)

Prints one JSON line per check: first-step activations vs an fp64 oracle, whole updates vs the fp64 oracle next
to the error of the plain fp32 PyTorch oracle, kernel time and phase counters."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gossipy_b200 import ops
from gossipy_b200.engine import rng
from gossipy_b200.ops import torch_ref as ref
from gossipy_b200.ops.native import native

IMPLS = sys.argv[1:] or ["tc8", "tc8-tf32", "tc3", "cluster"]


def problem(n, d_in, d_h, d_out, seed=0):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(n, d_in, generator=g).cuda()
    y = torch.randint(0, d_out, (n,), generator=g).cuda()
    P = d_h * d_in + d_h + d_out * d_h + d_out
    row = torch.zeros((P + 31) // 32 * 32, device="cuda")
    row[:P] = (torch.randn(P, generator=g) * 0.03).cuda()
    return X, y, row


def first_step(impl):
    dims = (784, 100, 10)
    X, y, row = problem(64, *dims)
    W1, b1, W2, b2 = ref.mlp1_unpack(row.double().clone(), dims)
    idx = torch.from_numpy(ref.perm_indices(64, rng.mix64(0x77 ^ 0))).cuda()[:32]
    z1 = X[idx].double() @ W1.t() + b1
    want = torch.relu(z1)
    if impl.startswith("tc4") or impl.startswith("tc8"):          # these dump dz1 of the first step
        pr = torch.softmax(want @ W2.t() + b2, dim=1)
        pr[torch.arange(32), y[idx]] -= 1.0
        want = ((pr / 32) @ W2) * (z1 > 0)
    got = native().mlp1_train_tc_debug(row.clone(), X, y, dims, 32, 1, 0.0, 0.0, 0x77, impl)
    torch.cuda.synchronize()
    err = float((got[:100, :].t().double() - want).abs().max())
    print(json.dumps({"check": "first-step activations vs fp64", "impl": impl, "max_abs_err": err,
                      "max_abs": float(want.abs().max()), "pad_rows_zero": float(got[100:].abs().max()) == 0.0}))


def whole_update(impl, dims, n, bs, ep, wd, lr):
    X, y, row = problem(n, *dims)
    start = row.clone()
    r64 = row.double().clone()
    ref.mlp1_train(r64, X.double(), y, dims, bs, ep, lr, wd, 0xABCDEF)
    r32 = row.clone()
    ref.mlp1_train(r32, X, y, dims, bs, ep, lr, wd, 0xABCDEF)
    got = row.clone()
    steps = ops.mlp1_train(got, X, y, dims, bs, ep, lr, wd, 0xABCDEF, impl=impl)
    torch.cuda.synchronize()
    P = dims[1] * dims[0] + dims[1] + dims[2] * dims[1] + dims[2]
    moved = float((r64[:P] - start[:P].double()).abs().max())
    e_k = float((got[:P].double() - r64[:P]).abs().max())
    e_32 = float((r32[:P].double() - r64[:P]).abs().max())
    rel = float((got[:P].double() - r64[:P]).norm() / r64[:P].norm())
    rel32 = float((r32[:P].double() - r64[:P]).norm() / r64[:P].norm())
    nW = dims[1] * dims[0]
    dW = got[:nW].double() - r64[:nW]
    shrink = float((torch.sign(r64[:nW]) * dW).mean() / r64[:nW].abs().mean())      # < 0: |W| systematically too small
    dmove = r64[:nW] - start[:nW].double()
    along = float((dW * dmove).sum() / (dmove * dmove).sum())                        # error component along the movement
    print(json.dumps({"check": "update vs fp64", "impl": impl, "dims": dims, "n": n, "bs": bs, "ep": ep, "wd": wd,
                      "steps": steps, "moved": moved, "kernel_max_err": e_k, "torch_fp32_max_err": e_32,
                      "kernel_rel_l2": rel, "torch_fp32_rel_l2": rel32, "W1_shrink": shrink, "W1_err_along_move": along}))


def timeit(fn, iters=8, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def timing(impl):
    dims = (784, 100, 10)
    X, y, row = problem(7500, *dims)
    med, best = timeit(lambda: ops.mlp1_train(row, X, y, dims, 32, 1, 0.1, 0.0, 1234, impl=impl))
    Xs, ys_ = X[:320].contiguous(), y[:320].contiguous()
    med10, _ = timeit(lambda: ops.mlp1_train(row, Xs, ys_, dims, 32, 1, 0.1, 0.0, 1234, impl=impl))
    per = (med - med10) / 225
    out = {"check": "timing 235 steps", "impl": impl, "ms_per_update": med, "best_ms": best,
           "us_per_step_marginal": per * 1e3, "fixed_us_per_launch": (med10 - 10 * per) * 1e3}
    if impl.startswith("tc4") or impl.startswith("tc8"):
        dbg = native().mlp1_train_tc_debug(row, X, y, dims, 32, 1, 0.1, 0.0, 1234, impl)
        torch.cuda.synchronize()
        flat = dbg.flatten()
        cn = ["wait fwd", "ld + RS send", "W2 + RS wait", "reduce + h + logits", "bar + softmax + dz2 send", "bar + dh own + AG send",
              "AG wait", "A2 images + fence", "W2 slice + b2", "wait update", "W += G, re-split", "end barrier"]
        mn = ["wait Wlo + issue fwd(c)", "wait fwd + copy X", "wait A2", "issue update", "wait update + W ready", "issue fwd(a,b) next", "wait X^T"]
        out["compute_cycles_cta0"] = {k: round(float(v)) for k, v in zip(cn, flat[0:12])}
        out["compute_sum"] = round(float(flat[0:12].sum()))
        out["issuer_cycles_cta0"] = {k: round(float(v)) for k, v in zip(mn, flat[32:39])}
        out["warp4_cycles_cta0"] = [round(float(v)) for v in flat[16:30]]
        out["compute_cycles_cta1"] = [round(float(v)) for v in flat[64:76]]
    print(json.dumps(out))


def timing_momentum():
    """The fused momentum-SGD form of tc8 (velocity tile in TMEM)."""
    dims = (784, 100, 10)
    X, y, row = problem(7500, *dims)
    buf = torch.zeros_like(row)
    med, best = timeit(lambda: ops.mlp1_train(row, X, y, dims, 32, 1, 0.01, 0.0, 1234, momentum=(0.9, 0.0, False, buf, False)))
    print(json.dumps({"check": "timing 235 steps", "impl": "tc8 + fused momentum", "ms_per_update": med, "best_ms": best,
                      "us_per_step": med / 235 * 1e3}))


if IMPLS == ["ncuonly"]:          # a handful of flagship updates for ncu to capture
    X, y, row = problem(7500, 784, 100, 10)
    for _ in range(4):
        ops.mlp1_train(row, X, y, (784, 100, 10), 32, 1, 0.1, 0.0, 1234)
    torch.cuda.synchronize()
    IMPLS = []
for impl in IMPLS:
    try:
        if impl != "cluster":
            first_step(impl)
        for cfg in (((784, 100, 10), 96, 32, 1, 0., .1), ((784, 100, 10), 70, 32, 2, .01, .1), ((64, 16, 4), 200, 16, 1, .001, .1),
                    ((784, 100, 10), 300, 32, 0, 0., .1), ((256, 128, 2), 130, 32, 1, 0., .05), ((784, 100, 10), 960, 32, 1, 0., .1),
                    ((784, 100, 10), 3200, 32, 1, 0., .1), ((784, 100, 10), 7500, 32, 1, 0., .1), ((784, 100, 10), 7500, 32, 1, 0., .01)):
            whole_update(impl, *cfg)
        timing(impl)
    except Exception as e:                                   # keep going: one JSON line per failure
        print(json.dumps({"impl": impl, "error": str(e)[:300]}))
if "tc8" in IMPLS:
    try:
        timing_momentum()
    except Exception as e:
        print(json.dumps({"impl": "tc8 + fused momentum", "error": str(e)[:300]}))
