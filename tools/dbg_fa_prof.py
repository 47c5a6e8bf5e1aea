import sys, os, ctypes as C, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chimera_st_b200
from chimera_st_b200 import ops, _lib as L
lib = L.load()
B, H, T = 32, 12, 750
g = torch.Generator().manual_seed(0)
q = (torch.randn(B, T, H * 64, generator=g) * 0.5).to(torch.bfloat16).cuda()
k = torch.randn(B, T, H * 64, generator=g).to(torch.bfloat16).cuda()
v = torch.randn(B, T, H * 64, generator=g).to(torch.bfloat16).cuda()
kl = torch.full((B,), T - 1, dtype=torch.int32).cuda()
for _ in range(3): ops.attention(q, k, v, H, kl)
torch.cuda.synchronize()
buf = (C.c_longlong * 16)()
lib.cst_debug_fa_prof(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.attention(q, k, v, H, kl); e1.record(); torch.cuda.synchronize()
lib.cst_debug_fa_prof(buf, 0)
n_cta = B * H * ((T + 127) // 128); tiles = n_cta * ((T - 1 + 127) // 128)
names = ["wait S", "pass1", "exchange", "take_pv", "pass2", "fence+arrive"]
print("kernel %.1f us, %d CTAs, %d tile-steps" % (e0.elapsed_time(e1) * 1e3, n_cta, tiles))
tot = sum(buf[i] for i in range(6))
for i, n in enumerate(names):
    print("%-14s %8.0f cycles per tile (%.0f%%)" % (n, buf[i] / tiles, 100.0 * buf[i] / tot))
print("sum per tile", tot / tiles)
