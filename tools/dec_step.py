"""One eager decoding step at a late step index (for ncu): python tools/dec_step.py [--B 64] [--M 64] [--dtype bf16]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import chimera_st_b200  # noqa: E402,F401
from chimera_st_b200 import synth  # noqa: E402
from chimera_st_b200.decoder import B200GreedyDecoder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=64)
ap.add_argument("--M", type=int, default=64)
ap.add_argument("--max-len", type=int, default=200)
ap.add_argument("--dtype", default="bf16")
a = ap.parse_args()
dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
dec = B200GreedyDecoder(synth.make_decoder_state_dict(seed=1), dtype=dt, device="cuda", use_graph=False)
mem = torch.randn(a.M, a.B, 512, generator=torch.Generator().manual_seed(5)).cuda().to(dt)
plan = dec._plan(a.B, a.M, a.max_len, mem.dtype)
plan.begin(mem)
for s in (0, a.max_len - 1):               # a warm-up step, then the profiled one (cache 200 rows long)
    plan.counters[0] = s
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("dec_step_%d" % s)
    plan._step()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("ok", plan.launches_per_step)
