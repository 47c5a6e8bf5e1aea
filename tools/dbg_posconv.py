import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chimera_st_b200
from chimera_st_b200 import _lib as L, ops
DEV = "cuda"
B, T, G, Kc = 1, 128, 16, 128 * 64
Tpp = T + 128
g = torch.Generator().manual_seed(1)
x = torch.randn(B * T, 768, generator=g)
bias = torch.zeros(768)
xd = x.to(DEV)
xg = torch.zeros(B * G * Tpp + 8, 64, dtype=torch.bfloat16, device=DEV)
L.check(L.load().cst_posconv_pack(xd.data_ptr(), B, T, T, xg.data_ptr(), L.BF16, Tpp, L.stream_ptr()))
zero = torch.zeros(B * T, 768, device=DEV)
for taps in ([0], [8], [64], [1], [2], [3], [7], [9], [65], list(range(128))):
    w = torch.zeros(G, 48, 128, 64)
    for j in taps:
        w[:, :, j, :48] = torch.randn(G, 48, 48, generator=g) * 0.1
    wd = w.reshape(G, 48, Kc).to(torch.bfloat16).to(DEV)
    ref = torch.zeros(B * T, 768, dtype=torch.float32, device=DEV)
    ops.gemm(xg, wd, ref, T, 48, Kc, lda=64, a_rows=Tpp, bias=bias.to(DEV), residual=zero, act=0, ldc=768,
             nb_outer=B, nb_inner=G, a_bs=(G * Tpp * 64, Tpp * 64), w_bs=48 * Kc, c_bs=(T * 768, 48), bias_bs=48)
    out = torch.zeros(B * T, 768, dtype=torch.float32, device=DEV)
    L.check(L.load().cst_posconv(xg.data_ptr(), wd.data_ptr(), bias.to(DEV).data_ptr(), zero.data_ptr(), out.data_ptr(), B, T, T, Tpp, L.stream_ptr()))
    torch.cuda.synchronize()
    # out has GELU applied; compare gelu(ref)
    refg = torch.nn.functional.gelu(ref)
    err = float((out - refg).norm() / refg.norm())
    rows_bad = ((out - refg).abs().max(1).values > 1e-2).nonzero().flatten()[:10].tolist()
    print("BO=%s taps=%s rel=%.3e bad_rows=%s" % (os.environ.get("CST_PC_BO", "0"), taps if len(taps) < 5 else "all", err, rows_bad))
