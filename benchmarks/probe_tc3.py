import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gossipy_b200.ops.native import native
m = native()
A = torch.zeros(128, 8, device="cuda")
# 1 + 2^-11 + 2^-12 : truncation -> 1.0, round-to-nearest -> 1 + 2^-10
A[:, 0] = 1.0 + 2.0 ** -11 + 2.0 ** -12
A[:, 1] = 1.0 + 2.0 ** -11            # tie: RN-even -> 1.0, RN-away -> 1 + 2^-10
A[:, 2] = 1.0 + 2.0 ** -10 + 2.0 ** -11   # tie, odd: RN-even -> 1 + 2^-9
A[:, 3] = -(1.0 + 2.0 ** -11 + 2.0 ** -12)
A[:, 4] = 1.0 + 2.0 ** -10
A[:, 5] = 3.0e-39                      # fp32 subnormal
A[:, 6] = 1.0 + 2.0 ** -23
A[:, 7] = 1.9999999
D, T = m.tc_probe3(A, 200, 192)
torch.cuda.synchronize()
print("A   ", [float.hex(float(x)) for x in A[5]])
print("tf32", [float.hex(float(x)) for x in D[5, :8]])
print("cycles per pass over 192 cols/lane (96 per thread), 8 warps: ld16 only %.0f, ld16+st16 %.0f, ld32+2xst16 %.0f" % tuple(T[:3].tolist()))
# SS-mode operand from shared memory
B = torch.zeros(8, 32, device="cuda"); B[0, :] = 1.0
A2 = torch.zeros(128, 8, device="cuda"); A2[:, 0] = A[:, 0]
print("SS A narrow:", float.hex(float(m.tc_probe(A2, B, 2)[0, 0])))
B2 = torch.zeros(8, 32, device="cuda"); B2[0, :] = A[0, 0]
A3 = torch.zeros(128, 8, device="cuda"); A3[:, 0] = 1.0
print("SS B narrow:", float.hex(float(m.tc_probe(A3, B2, 2)[0, 0])))
