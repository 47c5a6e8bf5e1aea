"""Phase breakdown of the tcgen05 GEMM epilogue (build with CST_EXTRA_NVCC_FLAGS=-DTC_PROFILE)."""
import ctypes as C
import math
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chimera_st_b200  # noqa: E402,F401
from chimera_st_b200 import ops, _lib as L  # noqa: E402

lib = L.load()
SHAPES = [
    (24000, 2304, 768, L.ACT_NONE, torch.bfloat16, False, "qkv"),
    (24000, 3072, 768, L.ACT_GELU, torch.bfloat16, False, "fc1"),
    (24000, 768, 3072, L.ACT_NONE, torch.float32, True, "fc2 (+residual)"),
    (24000, 768, 768, L.ACT_NONE, torch.float32, True, "out-proj (+residual)"),
]
g = torch.Generator().manual_seed(0)
for M, N, K, act, od, res, note in SHAPES:
    A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16).cuda()
    b = torch.randn(N, generator=g).cuda()
    R = torch.randn(M, N, generator=g).cuda() if res else None
    for _ in range(2):
        ops.linear(A, W, b, act=act, residual=R, out_dtype=od)
    torch.cuda.synchronize()
    buf = (C.c_longlong * 8)()
    lib.cst_debug_tc_epi(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.linear(A, W, b, act=act, residual=R, out_dtype=od); e1.record(); torch.cuda.synchronize()
    lib.cst_debug_tc_epi(buf, 0)
    n = max(buf[3], 1)
    print("%-22s %7.1f us %7.1f TF | per tile (warp 2, block 0, %d tiles): rows %5.0f  wait %6.0f  chunks %6.0f cycles; MMA floor %d" % (
        note, e0.elapsed_time(e1) * 1e3, 2.0 * M * N * K / e0.elapsed_time(e1) / 1e9, n, buf[0] / n, buf[1] / n, buf[2] / n, K // 64 * 512))
    print("      register-path chunk phases per tile: tmem wait %5.0f  store-read wait + sync %5.0f  transpose STS %5.0f  math + residual + stores %6.0f"
          % (buf[4] / n, buf[5] / n, buf[6] / n, buf[7] / n))
