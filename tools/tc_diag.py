"""Bring-up diagnostics for the tcgen05 GEMM: each case runs in its own process (a device trap must not
take the others down) and prints an error map against the FFMA fp32 kernel / torch on bf16-rounded inputs."""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [
    # M, N, K, note
    (128, 256, 64, "single tile, single k-block"),
    (128, 256, 256, "single tile, 4 k-blocks (one ring pass)"),
    (128, 256, 1024, "single tile, ring wraps"),
    (256, 512, 512, "2x2 tiles"),
    (1000, 768, 512, "M tail"),
    (128, 128, 512, "BN=128"),
    (128, 64, 512, "BN=64"),
    (128, 48, 512, "BN=48"),
    (20000, 768, 768, "persistent multi-tile per CTA"),
]

CHILD = r'''
import sys, math, torch
sys.path.insert(0, %r)
import chimera_st_b200
from chimera_st_b200 import ops
M, N, K = %d, %d, %d
g = torch.Generator().manual_seed(1)
A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16)
W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16)
ref = A.float().double() @ W.float().double().T
out = ops.linear(A.cuda(), W.cuda(), None, out_dtype=torch.float32)
torch.cuda.synchronize()
o = out.cpu().double()
err = (o - ref).abs()
rel = float((o - ref).norm() / ref.norm())
print("rel_l2=%%.3e max_abs=%%.3e" %% (rel, float(err.max())))
if rel > 1e-3:
    bad = err > 1e-2
    rows = bad.any(1).nonzero().flatten()
    cols = bad.any(0).nonzero().flatten()
    print("bad rows: n=%%d first=%%s" %% (len(rows), rows[:16].tolist()))
    print("bad cols: n=%%d first=%%s" %% (len(cols), cols[:16].tolist()))
    print("out[0,:8] ", o[0, :8].tolist())
    print("ref[0,:8] ", ref[0, :8].tolist())
    print("out[1,:4] ", o[1, :4].tolist(), "ref[1,:4]", ref[1, :4].tolist())
    # does the output match a K-truncated / permuted product?  (swizzle / descriptor hints)
    for kk in (16, 32, 64, 128):
        if kk <= K:
            part = A.float().double()[:, :kk] @ W.float().double()[:, :kk].T
            print("  vs first-%%d-of-K product: rel=%%.3e" %% (kk, float((o - part).norm() / part.norm())))
'''

if __name__ == "__main__":
    for M, N, K, note in CASES:
        code = CHILD % (ROOT, M, N, K)
        try:
            r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
            tail = (r.stdout + r.stderr).strip().splitlines()[-12:]
        except subprocess.TimeoutExpired:
            tail = ["TIMEOUT"]
        print("== tc gemm M=%d N=%d K=%d (%s)" % (M, N, K, note))
        for t in tail:
            print("   ", t)
        sys.stdout.flush()
