"""Per-stage bf16 (and fp32) error report against the fp32 goldens of the unmodified reference."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import chimera_st_b200  # noqa
from chimera_st_b200 import synth
from chimera_st_b200.encoder import build_encoder_from_state_dict


def rel(a, b):
    a, b = a.double().cpu(), b.double()
    return float((a - b).norm() / b.norm())


for dtype in (torch.float32, torch.bfloat16):
    sd = synth.make_state_dict(seed=0, interlingua_length=16)
    enc = build_encoder_from_state_dict(sd, dtype=dtype, device="cuda", use_graph=False)
    for name in ("tiny", "c1", "c1mix"):
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        wave, lens = synth.make_waveforms(g["src_lengths"].tolist(), seed=int(g["wave_seed"]))
        out = enc(wave.cuda(), lens.cuda())
        plan = enc._plan(*wave.shape)
        if name == "tiny":
            e = dict(conv=rel(plan.view("conv_feats"), torch.from_numpy(g["conv_feats"])),
                     w2v=rel(plan.view("w2v_out"), torch.from_numpy(g["w2v_out"])),
                     h_enc=rel(plan.view("h_enc"), torch.from_numpy(g["h_enc"])))
        else:
            e = dict(conv=rel(plan.view("conv_feats")[:, ::37, ::11], torch.from_numpy(g["conv_feats_s"])),
                     w2v=rel(plan.view("w2v_out")[:, ::11, ::37], torch.from_numpy(g["w2v_out_s"])),
                     h_enc=rel(plan.view("h_enc")[:, ::3, ::17], torch.from_numpy(g["h_enc_s"])))
        e["mem"] = rel(out.encoder_out, torch.from_numpy(g["memories"]))
        print(str(dtype).replace("torch.", ""), name, " ".join("%s=%.3e" % kv for kv in e.items()))
