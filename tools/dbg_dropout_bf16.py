"""Which dropout sites move the bf16 whole-gradient error (GPU): none / elementwise only / attention probabilities only / all."""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
import chimera_st_b200  # noqa: F401
from chimera_st_b200 import synth
from chimera_st_b200.train import EncoderTrainStep
from oracle import chimera_oracle as O
from emu import oracle_dropout_hook
from test_gpu_backward import _oracle_grads, _relu_masks_of, rel_l2

torch.set_num_threads(8)
lens = [6000, 4500]
sd = synth.make_state_dict(seed=0, interlingua_length=16, dead_heads=False)
wave, tl = synth.make_waveforms(lens, seed=31)
R = torch.randn(16, 2, 512, generator=torch.Generator().manual_seed(1))
for name, kw in (("none", {}), ("elementwise", dict(dropout=0.1, w2v_dropout=0.1, w2v_dropout_input=0.1, attention_dropout=0, w2v_attention_dropout=0)),
                 ("attention", dict(attention_dropout=0.1, w2v_attention_dropout=0.1)),
                 ("all", dict(dropout=0.1, w2v_dropout=0.1, w2v_dropout_input=0.1))):
    step = EncoderTrainStep(sd, 2, wave.shape[1], device="cuda", feature_grad_mult=1.0, dtype=torch.bfloat16, seed=9, **kw)
    g = step.g
    mem, G = step.forward_backward(wave, tl, R)
    torch.cuda.synchronize()

    def geom(tag):
        if tag.endswith(".prob"):
            return (g.T6a, g.Tp) if tag.startswith("w2v") else (16, g.T2) if tag.startswith("mem") else (g.T2a, g.T2)
        return g.T6a if tag.startswith("w2v") else 16 if tag.startswith("mem") else g.T2a

    def p_of(tag):
        key = (0, tag)
        return 0.1 if key in step._sites else 0.0
    hook, used = oracle_dropout_hook(step, 0, geom, p_of)
    O.DROPOUT_HOOK = hook
    try:
        ref_mem, ref, _ = _oracle_grads(sd, wave, tl, R, _relu_masks_of(step))
    finally:
        O.DROPOUT_HOOK = None
    num = den = 0.0
    worst = {}
    for k, v in G.items():
        if k.endswith("k_proj.bias"):
            continue
        d = (v.cpu().float().reshape(ref[k].shape) - ref[k]).double()
        num += float((d * d).sum()); den += float((ref[k].double() ** 2).sum())
        worst[k] = rel_l2(v.cpu().float().reshape(ref[k].shape), ref[k])
    top = sorted(worst.items(), key=lambda kv: -kv[1])[:5]
    print("%-12s sites %2d  mem %.2e  whole-gradient %.3e  worst %s" % (name, len(used and set(used)), rel_l2(mem.cpu().float(), ref_mem),
                                                                       (num / den) ** 0.5, [(k[-40:], "%.1e" % e) for k, e in top]), flush=True)
