"""Tiny greedy decode (fp32 FFMA kernels, bf16 mma kernels + bf16 cache) and text-branch forward, eager launches, for
compute-sanitizer:   compute-sanitizer --tool memcheck python tools/sanitize_dec.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chimera_st_b200
from chimera_st_b200 import synth
from chimera_st_b200.decoder import B200GreedyDecoder
from chimera_st_b200.encoder import build_encoder_from_state_dict
dsd = synth.make_decoder_state_dict(seed=1)
# 11 hypotheses (partial 8-row warp groups, partial 16-row MMA tiles), M = 16 memories, 6 steps
mem = torch.randn(16, 11, 512, generator=torch.Generator().manual_seed(5)).cuda()
only = os.environ.get("SANITIZE_ONLY", "")
for dtype in ((torch.bfloat16,) if only == "bf16" else (torch.float32, torch.bfloat16)):
    dec = B200GreedyDecoder(dsd, dtype=dtype, device="cuda", use_graph=False)
    hyp = dec.generate(mem.to(dtype), max_len=5)
    torch.cuda.synchronize()
    print("decode", dtype, [len(h["tokens"]) for h in hyp], dec.last_launches)
sd = synth.make_state_dict(seed=0, text_vocab=synth.VOCAB)
tok = torch.randint(4, synth.VOCAB, (3, 13), generator=torch.Generator().manual_seed(2))
lens = torch.tensor([13, 7, 1])
for dtype in (() if only == "bf16" else (torch.float32, torch.bfloat16)):
    enc = build_encoder_from_state_dict(sd, dtype=dtype, device="cuda", use_graph=False)
    out = enc(tok.cuda(), lens.cuda())
    torch.cuda.synchronize()
    print("text", dtype, float(out.encoder_out.abs().mean()))
