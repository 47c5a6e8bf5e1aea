import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import chimera_st_b200  # noqa
from chimera_st_b200 import synth
from chimera_st_b200.train import EncoderTrainStep
from oracle import chimera_oracle as O
torch.set_num_threads(8)
lens = [6000, 4500]
sd = synth.make_state_dict(seed=0, interlingua_length=16, dead_heads=False)
wave, tl = synth.make_waveforms(lens, seed=31)
R = torch.randn(16, len(lens), 512, generator=torch.Generator().manual_seed(1))
sdg = {k: (v.clone().requires_grad_() if v.is_floating_point() else v) for k, v in sd.items()}
mem, _ = O.encoder_forward(sdg, wave, tl)
(mem * R).sum().backward()
step = EncoderTrainStep(sd, len(lens), wave.shape[1], device="cuda", feature_grad_mult=1.0)
m2, G = step.forward_backward(wave, tl, R)
torch.cuda.synchronize()
sd64 = {k: (v.double().clone().requires_grad_() if v.is_floating_point() else v) for k, v in sd.items()}
mem64, _ = O.encoder_forward(sd64, wave.double(), tl)
(mem64 * R.double()).sum().backward()
rows = []
for k, v in G.items():
    t = sd64[k].grad
    r32 = sdg[k].grad.double()
    n = t.norm().clamp_min(1e-30)
    ours = float((v.cpu().reshape(t.shape).double() - t).norm() / n)
    orc = float((r32 - t).norm() / n)
    rows.append((ours, orc, float(t.abs().max()), k))
for e, eo, rmax, k in sorted(rows, reverse=True)[:40]:
    print("ours %.3e  oracle-fp32 %.3e  ref max %.2e  %s" % (e, eo, rmax, k))
big = [r for r in rows if r[2] > 1e-4]
print("n grads", len(rows), "median ours %.3e oracle32 %.3e; max ours %.3e oracle32 %.3e (non-degenerate)" % (
    sorted(r[0] for r in big)[len(big) // 2], sorted(r[1] for r in big)[len(big) // 2], max(r[0] for r in big), max(r[1] for r in big)))
# timing of the C5 shape
import time
B, Lw = 8, 150000
wave, tl = synth.make_waveforms([Lw] * B, seed=3)
R = torch.randn(16, B, 512)
step = EncoderTrainStep(sd, B, Lw, device="cuda")
for _ in range(2):
    step.forward_backward(wave.cuda(), tl.cuda(), R.cuda())
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); step.forward(wave.cuda(), tl.cuda()); e[1].record(); step.backward(R.cuda()); e[2].record(); torch.cuda.synchronize()
print("C5 shape B=8 L=150000 fp32: forward %.1f ms, backward %.1f ms -> %.0f audio-s/s" % (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), B * Lw / 16000 / (e[0].elapsed_time(e[2]) / 1e3)))
