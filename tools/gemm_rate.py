"""Isolated tcgen05 GEMM rates for the encoder's shapes: `python tools/gemm_rate.py` (honours CST_TC_PAIR / CST_TC_BN)."""
import math
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chimera_st_b200  # noqa: E402,F401
from chimera_st_b200 import ops, _lib as L  # noqa: E402

SHAPES = [  # M, N, K, act, out dtype, note
    (24000, 2304, 768, L.ACT_NONE, torch.bfloat16, "qkv"),
    (24000, 3072, 768, L.ACT_GELU, torch.bfloat16, "fc1"),
    (24000, 768, 3072, L.ACT_NONE, torch.float32, "fc2 +res"),
    (24000, 768, 768, L.ACT_NONE, torch.float32, "out-proj +res"),
    (383000, 512, 1536, L.ACT_GELU, torch.bfloat16, "conv2-like (dense A)"),
    (6000, 2304, 768, L.ACT_NONE, torch.bfloat16, "qkv small batch"),
]


def main():
    if os.environ.get('CST_TC_DBG'):
        L.load().cst_debug_tc_flags(int(os.environ['CST_TC_DBG']))
    g = torch.Generator().manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for M, N, K, act, od, note in SHAPES:
        A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
        W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16).cuda()
        b = torch.randn(N, generator=g).cuda()
        R = torch.randn(M, N, generator=g).cuda() if "+res" in note else None
        for _ in range(3):
            out = ops.linear(A, W, b, act=act, residual=R, out_dtype=od)
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = ops.linear(A, W, b, act=act, residual=R, out_dtype=od)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        # spot check against torch on a row sample
        idx = torch.randint(0, M, (64,), generator=g).cuda()
        ref = A[idx].float() @ W.float().T + b
        if act == L.ACT_GELU:
            ref = torch.nn.functional.gelu(ref)
        if R is not None:
            ref = ref + R[idx]
        err = float((out[idx].float() - ref).norm() / ref.norm())
        print(f"{note:24s} M={M} N={N} K={K}: {ms*1e3:8.1f} us  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s  rel_err={err:.2e}", flush=True)


if __name__ == "__main__":
    main()
