"""The tcgen05 GEMM has two kernels (single CTA, CTA pair) and three epilogue paths (register + STG, bulk store, staged
bulk store for fp32 + residual), selected by problem shape and by CST_TC_PAIR / CST_TC_BULK.  They run the same
arithmetic in the same order, so on the same inputs every combination must produce the SAME BITS -- checked here in
subprocesses (the switches are read once per process)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

CHILD = r'''
import sys, math, torch
sys.path.insert(0, %(root)r)
import chimera_st_b200
from chimera_st_b200 import ops, _lib as L
out = {}
g = torch.Generator().manual_seed(7)
# (M, N, K, act, out dtype, residual): M tails on both tile sizes, every epilogue family
cases = [(24001, 768, 768, L.ACT_NONE, torch.float32, True), (20000 + 77, 2304, 768, L.ACT_NONE, torch.bfloat16, False),
         (33000, 512, 1536, L.ACT_GELU, torch.bfloat16, False), (16500, 3072, 768, L.ACT_GELU, torch.bfloat16, False),
         (16400, 768, 3072, L.ACT_NONE, torch.float32, True), (18000, 512, 512, L.ACT_RELU, torch.float32, False)]
for i, (M, N, K, act, od, res) in enumerate(cases):
    A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16).cuda()
    b = torch.randn(N, generator=g).cuda()
    R = torch.randn(M, N, generator=g).cuda() if res else None
    y = ops.linear(A, W, b, act=act, residual=R, out_dtype=od)
    torch.cuda.synchronize()
    out["c%%d" %% i] = y.cpu()
torch.save(out, %(path)r)
'''


def _run(tmp_path, pair, bulk):
    path = str(tmp_path / ("out_p%d_b%d.pt" % (pair, bulk)))
    env = dict(os.environ, CST_TC_PAIR=str(pair), CST_TC_BULK=str(bulk))
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "path": path}], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return torch.load(path)


def test_all_kernel_and_epilogue_variants_are_bit_identical(tmp_path):
    ref = _run(tmp_path, 0, 0)                      # single-CTA kernel, register-path epilogue
    for pair, bulk in ((0, 3), (2, 0), (2, 3), (3, 3)):
        got = _run(tmp_path, pair, bulk)
        for k in ref:
            assert torch.equal(ref[k], got[k]), (pair, bulk, k, float((ref[k].float() - got[k].float()).abs().max()))
