"""Bisect where a super-batch plan departs from the same batches run alone (fp32): truncated layer stacks, per-buffer diffs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import chimera_st_b200  # noqa: F401
from chimera_st_b200 import synth
from chimera_st_b200.encoder import build_encoder_from_state_dict

shapes = [[16000, 12345, 8000], [9000, 7000], [24000], [16000, 3000, 9999], [9000, 8999], [12000, 11000, 10000, 500]]
sd = synth.make_state_dict(seed=0, interlingua_length=16)
dtype = torch.float32
enc = build_encoder_from_state_dict(sd, dtype=dtype, device="cuda", use_graph=False)
data = [tuple(t.cuda() for t in synth.make_waveforms(l, seed=100 + i)) for i, l in enumerate(shapes)]
NAMES = ("x", "y", "xa", "qkv", "ctx", "ffn", "w2v_out")
for nl in (2, 3):
    singles = []
    for w, l in data:
        p = enc._plan(*w.shape)
        fullP = p.P
        p.P = dict(fullP, w2v_layers=fullP["w2v_layers"][:nl])
        p.load_inputs(w, l)
        p.run(upto="w2v")
        torch.cuda.synchronize()
        singles.append({n: getattr(p, n).clone() for n in NAMES} | {"valid": p.w2v_valid.clone()})
        p.P = fullP
    sp = enc._plan(None, None, groups=[tuple(w.shape) for w, _ in data])
    fullP = sp.P
    sp.P = dict(fullP, w2v_layers=fullP["w2v_layers"][:nl])
    for k, (w, l) in enumerate(data):
        sp.load_inputs(w, l, group=k)
    sp.run(upto="w2v")
    torch.cuda.synchronize()
    sp.P = fullP
    for k in (0, 1, 2, 5):
        g = sp.gs[k]
        r0, rows = sp.r0[k], g.B * g.T6a
        print("layers", nl, "group", k, "valid", singles[k]["valid"].tolist(), sp.w2v_valid[sp.utt0[k]:sp.utt0[k] + g.B].tolist(), "T6a", g.T6a, "Tp", g.Tp)
        for n in NAMES:
            a = getattr(sp, n)[r0:r0 + rows].float()
            b = singles[k][n][:rows].float()
            d = (a - b).abs().view(g.B, g.T6a, -1).amax(-1)          # [B, T6a]
            bad = [(int(b_), int(t_)) for b_, t_ in (d > 0).nonzero()[:6]]
            cols = (a - b).abs().view(g.B, g.T6a, -1)[bad[0][0], bad[0][1]].nonzero().flatten()[:8].tolist() if bad else []
            print("   %-8s max %.3e  rows differing %d  first %s cols %s" % (n, float(d.max()), int((d > 0).sum()), bad, cols))
