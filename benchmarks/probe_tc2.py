import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gossipy_b200.ops.native import native
m = native()
torch.manual_seed(0)
for (M, N, K, asw, bsw) in ((64, 16, 128, True, True), (64, 16, 128, False, False), (128, 32, 16, False, False),
                            (128, 16, 32, False, False), (128, 16, 128, True, True), (64, 32, 32, True, False)):
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda")
    want = A @ B.t()
    got = m.tc_probe2(A, B, asw, bsw)
    torch.cuda.synchronize()
    print("M", M, "N", N, "K", K, "a_sw", asw, "b_sw", bsw, "max err %.4f" % float((got - want).abs().max()),
          "max |want| %.2f" % float(want.abs().max()))
