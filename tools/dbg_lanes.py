import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chimera_st_b200
from chimera_st_b200 import synth
from chimera_st_b200.encoder import build_encoder_from_state_dict
from oracle import chimera_oracle as O
sd = synth.make_state_dict(seed=0)
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())
w1, l1 = synth.make_waveforms([24000], seed=102)
st = {}
with torch.no_grad():
    ref, _ = O.encoder_forward(sd, w1, l1, stages=st)
def report(enc, tag):
    out = enc(w1.cuda(), l1.cuda()).encoder_out
    p = enc._plan(1, 24000)
    print(tag, "mem=%.2e conv=%.2e w2v=%.2e h_enc=%.2e" % (rel(out, ref), rel(p.view("conv_feats"), st["conv_feats"]),
          rel(p.view("w2v_out"), st["w2v_out"]), rel(p.view("h_enc"), st["h_enc"].cpu())), "arena bytes", enc._arena.buf.numel(), "plan bytes", p.nbytes)
for graph in (False, True):
    enc = build_encoder_from_state_dict(sd, dtype=torch.float32, device="cuda", use_graph=graph)
    report(enc, "graph=%s fresh        " % graph)
    for lens in ([16000, 12345, 8000], [9000, 7000]):
        w, l = synth.make_waveforms(lens, seed=5)
        enc(w.cuda(), l.cuda())
    report(enc, "graph=%s after others " % graph)
    enc2 = build_encoder_from_state_dict(sd, dtype=torch.float32, device="cuda", use_graph=graph)
    for lens in ([16000, 12345, 8000], [9000, 7000]):
        w, l = synth.make_waveforms(lens, seed=5)
        enc2(w.cuda(), l.cuda())
    report(enc2, "graph=%s others first " % graph)
