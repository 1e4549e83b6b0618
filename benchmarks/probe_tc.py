import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gossipy_b200.ops.native import native
m = native()
torch.manual_seed(0)
for K, N in ((32, 64), (32, 256), (32, 160), (8, 32), (64, 32)):
    A = torch.randn(128, K, device="cuda"); B = torch.randn(K, N, device="cuda")
    want = A @ B
    for v in (2, 4, 8):
        if v == 8 and K % 32: continue
        got = m.tc_probe(A, B, v)
        torch.cuda.synchronize()
        err = float((got - want).abs().max())
        print("K", K, "N", N, "variant", v, "max err %.4f" % err, "max |want| %.2f" % float(want.abs().max()),
              "nonzero frac %.3f" % float((got != 0).float().mean()))
