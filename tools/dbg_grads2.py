"""Bisect the backward pass: gradients w.r.t. intermediate activations, ours vs fp64 autograd through the oracle (and the oracle in fp32)."""
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
def run(dt):
    s = {k: (v.to(dt).clone().requires_grad_() if v.is_floating_point() else v) for k, v in sd.items()}
    st = {}
    mem, _ = O.encoder_forward(s, wave.to(dt), tl, stages=st)
    keep = {k: v for k, v in st.items() if torch.is_tensor(v) and v.is_floating_point() and v.requires_grad}
    for v in keep.values():
        v.retain_grad()
    (mem * R.to(dt)).sum().backward()
    return {k: v.detach() for k, v in keep.items()}, {k: v.grad for k, v in keep.items()}
a64, g64 = run(torch.float64)
a32, g32 = run(torch.float32)
step = EncoderTrainStep(sd, len(lens), wave.shape[1], device="cuda", feature_grad_mult=1.0)
m2, G = step.forward_backward(wave, tl, R)
torch.cuda.synchronize()
g, T = step.g, step.T
B = g.B
def rows(t, rps, n):         # [B*rps(+slack), C] -> [B, n, C]
    return t[:B * rps].view(B, rps, -1)[:, :n].cpu().double()
ours_act = {"h_enc": rows(T["h_enc"], g.T2a, g.T2), "w2v_out": rows(T["w2v_out"], g.T6a, g.Tp)}
ours_g = {"h_enc": rows(step.dbg["h_enc"], g.T2a, g.T2), "sub_out": rows(step.dbg["sub_out"], g.T2a, g.T2),
          "w2v_out": rows(step.dbg["w2v_out"], g.T6a, g.Tp), "w2v_in": rows(step.dbg["w2v_in"], g.T6a, g.Tp),
          "proj_masked": rows(step.dbg["proj_masked"], g.T6a, g.Tp),
          "conv_feats": rows(step.dbg["conv_feats"], g.T6a, g.Tp).transpose(1, 2)}
def rl(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
for k in ours_act:
    print("act  %-12s ours %.3e  oracle32 %.3e" % (k, rl(ours_act[k], a64[k]), rl(a32[k].double(), a64[k])))
for k in ours_g:
    if k in g64 and g64[k] is not None:
        print("grad %-12s ours %.3e  oracle32 %.3e  shape %s" % (k, rl(ours_g[k], g64[k]), rl(g32[k].double(), g64[k]), tuple(g64[k].shape)))
