"""Event trace of the CTA-pair GEMM's pipeline (build with CST_EXTRA_NVCC_FLAGS=-DTC_PROFILE)."""
import ctypes as C
import math
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chimera_st_b200  # noqa: E402,F401
from chimera_st_b200 import ops, _lib as L  # noqa: E402

lib = L.load()
M, N, K = 24000, 768, 3072
g = torch.Generator().manual_seed(0)
A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16).cuda()
for _ in range(3):
    out = ops.linear(A, W, None, out_dtype=torch.float32)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 4096)()
lib.cst_debug_tc_trace(buf)
ev = [[buf[r * 1024 + i] for i in range(1024)] for r in range(4)]
t0 = min(e[0] for e in ev if e[0])
print("kb   lead_prod  peer_prod  mma_full  mma_commit   (ns since first event)")
for i in list(range(0, 30)) + list(range(100, 130)):
    print("%3d %10d %10d %9d %10d" % (i, ev[0][i] - t0, ev[1][i] - t0, ev[2][i] - t0, ev[3][i] - t0))
d = [ev[2][i + 1] - ev[2][i] for i in range(100, 400)]
print("steady-state k-block period: %.1f ns" % (sum(d) / len(d)))
lag = [ev[2][i] - max(ev[0][i], ev[1][i]) for i in range(100, 400)]
print("TMA issue (later producer) -> MMA sees full: %.1f ns avg" % (sum(lag) / len(lag)))
lag2 = [ev[0][i + 6] - ev[3][i] for i in range(100, 400)]
print("MMA commit(kb) -> leader producer passes empty for kb+6: %.1f ns avg" % (sum(lag2) / len(lag2)))
lag3 = [ev[1][i + 6] - ev[3][i] for i in range(100, 400)]
print("MMA commit(kb) -> peer producer passes empty for kb+6: %.1f ns avg" % (sum(lag3) / len(lag3)))
