"""Which ingredient of a super-batch changes fp32 bits: run-to-run, the segment-table attention, or sharing launches?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import chimera_st_b200  # noqa: F401
from chimera_st_b200 import synth
from chimera_st_b200.encoder import build_encoder_from_state_dict

sd = synth.make_state_dict(seed=0, interlingua_length=16)
enc = build_encoder_from_state_dict(sd, dtype=torch.float32, device="cuda", use_graph=False)
A = tuple(t.cuda() for t in synth.make_waveforms([16000, 12345, 8000], seed=100))
Bt = tuple(t.cuda() for t in synth.make_waveforms([9000, 7000], seed=101))
C1 = tuple(t.cuda() for t in synth.make_waveforms([24000], seed=102))

def single(d):
    p = enc._plan(*d[0].shape)
    p.load_inputs(*d)
    p.run(upto="w2v")
    torch.cuda.synchronize()
    return p.w2v_out.clone()

def sup(ds):
    p = enc._plan(None, None, groups=[tuple(d[0].shape) for d in ds])
    for k, d in enumerate(ds):
        p.load_inputs(d[0], d[1], group=k)
    p.run(upto="w2v")
    torch.cuda.synchronize()
    return [p.w2v_out[p.r0[k]:p.r0[k] + p.gs[k].B * p.gs[k].T6a].clone() for k in range(len(ds))]

sA, sA2, sB = single(A), single(A), single(Bt)
print("single A twice        ", float((sA - sA2).abs().max()))
print("super [A, C1]: A      ", float((sup([A, C1])[0] - sA).abs().max()))
print("super [C1, A]: A      ", float((sup([C1, A])[1] - sA).abs().max()))
print("super [A, A]: A0, A1  ", [float((x - sA).abs().max()) for x in sup([A, A])])
print("super [A, B]: A, B    ", [float((x - y).abs().max()) for x, y in zip(sup([A, Bt]), (sA, sB))])
os.environ["CST_DBG"] = "1"
