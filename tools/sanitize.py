"""Tiny forward (fp32 and bf16, eager) for compute-sanitizer runs:
   compute-sanitizer --tool memcheck python tools/sanitize.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chimera_st_b200
from chimera_st_b200 import synth
from chimera_st_b200.encoder import build_encoder_from_state_dict
sd = synth.make_state_dict(seed=0)
# [9000, 5000]: T' = 28 frames (memory_attention kernel everywhere); [48000, 30000]: T' = 149 (tcgen05 attention, several
# GEMM tiles, two pos-conv frame tiles, conv0 tail tile)
for lens_ in ([9000, 5000], [48000, 30000]):
    wave, lens = synth.make_waveforms(lens_, seed=3)
    for dtype in (torch.float32, torch.bfloat16):
        if dtype == torch.float32 and lens_[0] > 9000:
            continue
        enc = build_encoder_from_state_dict(sd, dtype=dtype, device="cuda", use_graph=False)
        out = enc(wave.cuda(), lens.cuda())
        torch.cuda.synchronize()
        print(lens_, dtype, float(out.encoder_out.abs().mean()))
