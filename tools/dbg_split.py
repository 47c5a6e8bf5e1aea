"""Accuracy of the 3-term fp16 split GEMM on tcgen05 against fp64, as a function of K (is the fp32 accumulation in tensor memory
rounded or truncated?), next to the CUDA-core FFMA GEMM."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import chimera_st_b200  # noqa
from chimera_st_b200 import ops, _lib as L, weights
lib = L.load()
torch.manual_seed(0)
M, N = 512, 768
for K in (64, 256, 768, 3072):
    for bias_mean in (0.0, 1.0):
        x = (torch.randn(M, K) + bias_mean).cuda()
        w = (torch.randn(N, K) * 0.03).cuda()
        ref = x.double() @ w.double().T
        out_f = torch.empty(M, N, device="cuda")
        ops.gemm(x, w, out_f, M, N, K, lda=K, a_rows=M)
        w16, inv = weights.split_w(w.cpu(), K)
        xs = torch.empty(M, 3 * K, dtype=torch.float16, device="cuda")
        L.check(lib.cst_split_f16(x.data_ptr(), K, M, K, xs.data_ptr(), L.stream_ptr()))
        p = L.GemmParams()
        out_t = torch.empty(M, N, device="cuda")
        w16 = w16.cuda()
        p.A, p.W, p.C = xs.data_ptr(), w16.data_ptr(), out_t.data_ptr()
        p.ab_dtype, p.c_dtype = L.F16, L.F32
        p.M, p.N, p.K, p.lda, p.ldc, p.a_rows = M, N, 3 * K, 3 * K, N, M
        p.alpha, p.nb_outer, p.nb_inner = 1.0, 1, 1
        p.rows_per_seg = p.seg_rows_valid = M
        p.out_rows_per_seg = M
        p.acc_scale = inv
        import ctypes as C
        L.check(lib.cst_gemm(C.byref(p), L.stream_ptr()))
        torch.cuda.synchronize()
        ef = float((out_f.double() - ref).norm() / ref.norm())
        et = float((out_t.double() - ref).norm() / ref.norm())
        bias_t = float(((out_t.double() - ref) / ref.abs().clamp_min(1e-9)).median())
        # operand-side error only: exact fp64 product of the split operands
        xr = xs.double(); wr = w16.double() * inv
        Kk = K
        approx = xr[:, :Kk] @ wr[:, :Kk].T + xr[:, Kk:2 * Kk] @ wr[:, Kk:2 * Kk].T + xr[:, 2 * Kk:] @ wr[:, 2 * Kk:].T
        es = float((approx - ref).norm() / ref.norm())
        print("K=%4d mean=%.0f  FFMA %.2e | split on tcgen05 %.2e (signed median rel %.2e) | split operands in fp64 %.2e" % (K, bias_mean, ef, et, bias_t, es))
