"""Tiny training steps (fp32 and bf16, eager: forward, backward, fused Adam, operand refresh) for compute-sanitizer runs:
   compute-sanitizer --tool memcheck python tools/sanitize_train.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chimera_st_b200  # noqa
from chimera_st_b200 import synth, losses
from chimera_st_b200.train import EncoderTrainStep, FusedAdam
sd = synth.make_state_dict(seed=0, interlingua_length=16, dead_heads=False)
for lens_, dtype in (([6000, 4500], torch.float32), ([6000, 4500], torch.bfloat16), ([30000, 21000], torch.bfloat16)):
    wave, lens = synth.make_waveforms(lens_, seed=3)
    step = EncoderTrainStep(sd, len(lens_), wave.shape[1], device="cuda", dtype=dtype)
    mem = step.forward(wave.cuda(), lens.cuda())
    text = torch.randn(16, len(lens_), 512, generator=torch.Generator().manual_seed(1)).cuda()
    _, loss, da, _ = losses.contrastive_loss(mem.contiguous(), text, temp=0.1, grad_scale=1.0)
    G = step.backward(da)
    opt = FusedAdam({k: step.sd[k] for k in G}, lr=1e-5)
    opt.advance()
    opt.step({k: v.contiguous() for k, v in G.items()})
    step.refresh_weights()
    torch.cuda.synchronize()
    print(lens_, dtype, float(loss), len(G))
