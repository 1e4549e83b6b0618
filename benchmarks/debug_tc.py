import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gossipy_b200 import ops
from gossipy_b200.ops import torch_ref as ref
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_kernels_gpu import _mlp_problem
dims = (784, 100, 10)
for n, lr in ((32, .1), (96, .1)):
    X, y, row = _mlp_problem(n, *dims)
    want = row.clone(); start = row.clone()
    ref.mlp1_train(want, X, y, dims, 32, 1, lr, 0., 0xABCDEF)
    ops.mlp1_train(row, X, y, dims, 32, 1, lr, 0., 0xABCDEF, impl="tc")
    W1g, b1g, W2g, b2g = ref.mlp1_unpack(row, dims)
    W1w, b1w, W2w, b2w = ref.mlp1_unpack(want, dims)
    W1s = ref.mlp1_unpack(start, dims)[0]
    print("n", n)
    for name, a, b in (("b1", b1g, b1w), ("W2", W2g, W2w), ("b2", b2g, b2w)):
        print("  ", name, "max err", float((a - b).abs().max()), "max |ref|", float(b.abs().max()))
    for lo, hi in ((0, 256), (256, 392), (392, 648), (648, 784)):
        e = (W1g[:, lo:hi] - W1w[:, lo:hi]).abs()
        mv = (W1w[:, lo:hi] - W1s[:, lo:hi]).abs()
        print("   W1 cols [%d,%d): max err %.5f  max move %.5f  ratio of moves got/want %.4f" % (
            lo, hi, float(e.max()), float(mv.max()),
            float(((W1g[:, lo:hi] - W1s[:, lo:hi]) * (W1w[:, lo:hi] - W1s[:, lo:hi])).sum() / (mv ** 2).sum())))
    d = (W1g - W1w).abs()
    j, k = divmod(int(d.argmax()), 784)
    print("   worst at hidden", j, "feature", k, "got", float(W1g[j, k]), "want", float(W1w[j, k]), "start", float(W1s[j, k]))
    rows_bad = (d.max(1).values > 1e-3).nonzero().flatten().tolist()
    cols_bad = (d.max(0).values > 1e-3).nonzero().flatten().tolist()
    print("   #bad rows", len(rows_bad), rows_bad[:20], "#bad cols", len(cols_bad), cols_bad[:20])
