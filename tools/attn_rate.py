"""Isolated rate of the bf16 self-attention kernel at the path's shapes (CUDA events, L2-cold inputs rotated between launches)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chimera_st_b200  # noqa
from chimera_st_b200 import ops
for B, H, T in ((32, 12, 749), (64, 12, 999), (16, 12, 1499), (48, 12, 468), (64, 8, 188)):
    g = torch.Generator().manual_seed(0)
    sets = []
    for _ in range(3):
        q = (torch.randn(B, T, H * 64, generator=g) * 0.5).to(torch.bfloat16).cuda()
        k = torch.randn(B, T, H * 64, generator=g).to(torch.bfloat16).cuda()
        v = torch.randn(B, T, H * 64, generator=g).to(torch.bfloat16).cuda()
        sets.append((q, k, v))
    kl = torch.full((B,), T, dtype=torch.int32).cuda()
    for q, k, v in sets:
        ops.attention(q, k, v, H, kl)
    torch.cuda.synchronize()
    n = 12
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        q, k, v = sets[i % 3]
        ops.attention(q, k, v, H, kl)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    print("B%d H%d T%d: %.1f us  %.0f TFLOP/s" % (B, H, T, us, 4.0 * B * H * T * T * 64 / us / 1e6))
