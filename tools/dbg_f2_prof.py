"""Phase breakdown of attention_tc2_kernel (build with CST_EXTRA_NVCC_FLAGS=-DF2_PROFILE): clock64 deltas of one softmax warp
per query tile and of the MMA issuer of block 0, c2 shape (B=32, H=12, T=750)."""
import sys, os, ctypes as C, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chimera_st_b200  # noqa
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
buf = (C.c_longlong * 32)()
lib.cst_debug_f2_prof(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.attention(q, k, v, H, kl); e1.record(); torch.cuda.synchronize()
lib.cst_debug_f2_prof(buf, 0)
items = B * H * ((T + 255) // 256)
per_cta = (items + 147) // 148
tiles = per_cta * ((T - 1 + 127) // 128)          # key tiles block 0 walked (upper bound: block 0 gets ceil)
print("kernel %.1f us; block 0: ~%d items, ~%d key tiles per query tile" % (e0.elapsed_time(e1) * 1e3, per_cta, tiles))
names = ["wait S", "pass1", "take_pv", "token", "pass2", "publish", "between"]
for wg in (0, 1):
    tot = sum(buf[wg * 8 + i] for i in range(7))
    print("softmax warpgroup %d: total %.0f cycles per key tile" % (wg, tot / tiles))
    for i, n in enumerate(names):
        print("   %-10s %8.0f cycles per tile (%.0f%%)" % (n, buf[wg * 8 + i] / tiles, 100.0 * buf[wg * 8 + i] / max(tot, 1)))
print("MMA issuer: wait P0 %.0f, wait P1 %.0f, issue-after-P1 %.0f, issue-after-P0 %.0f cycles per key tile" %
      tuple(buf[16 + i] / tiles for i in range(4)))
