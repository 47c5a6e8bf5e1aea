"""Tiny training steps with the full dropout recipe (fp32 and bf16, eager) for compute-sanitizer runs:
   compute-sanitizer --tool memcheck python tools/sanitize_dropout.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chimera_st_b200  # noqa
from chimera_st_b200 import synth
from chimera_st_b200.train import EncoderTrainStep
sd = synth.make_state_dict(seed=0, interlingua_length=16, dead_heads=False)
for lens_, dtype in (([6000, 4500], torch.float32), ([9100, 4500, 700], torch.bfloat16)):
    wave, lens = synth.make_waveforms(lens_, seed=3)
    step = EncoderTrainStep(sd, len(lens_), wave.shape[1], device="cuda", dtype=dtype, dropout=0.1, w2v_dropout=0.1, w2v_dropout_input=0.1)
    R = torch.randn(16, len(lens_), 512, generator=torch.Generator().manual_seed(1)).cuda()
    mem, G = step.forward_backward(wave.cuda(), lens.cuda(), R)
    torch.cuda.synchronize()
    print(lens_, dtype, float(mem.float().abs().mean()), len(G), len(step._sites))
