"""Bisect the memory stage's backward: ours vs fp64 / fp32 autograd of a local restatement, same h_enc."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import chimera_st_b200  # noqa
from chimera_st_b200 import synth
from chimera_st_b200.train import EncoderTrainStep
torch.set_num_threads(8)
lens = [6000, 4500]
sd = synth.make_state_dict(seed=0, interlingua_length=16, dead_heads=False)
wave, tl = synth.make_waveforms(lens, seed=31)
R = torch.randn(16, len(lens), 512, generator=torch.Generator().manual_seed(1))
step = EncoderTrainStep(sd, len(lens), wave.shape[1], device="cuda", feature_grad_mult=1.0)
m2, G = step.forward_backward(wave, tl, R)
torch.cuda.synchronize()
g, T = step.g, step.T
B, M = g.B, 16
h_enc0 = T["h_enc"][:B * g.T2a].view(B, g.T2a, 512)[:, :g.T2].cpu()
def ln(x, s, n):
    return F.layer_norm(x, (512,), s[n + ".weight"], s[n + ".bias"], 1e-5)
def run(dt):
    s = {k: v.to(dt) for k, v in sd.items() if v.is_floating_point()}
    h_enc = h_enc0.to(dt).clone().requires_grad_()
    mem = s["interlingua_embedding.weight"].clone().requires_grad_().unsqueeze(0).repeat(B, 1, 1)
    rec = []
    for l in range(3):
        P = f"interlingua_layers.{l}."
        r = {}
        a = ln(mem, s, P + "self_attn_layer_norm"); r["a"] = a
        kv_in = ln(h_enc, s, P + "self_attn_layer_norm"); r["kv_in"] = kv_in
        q = F.linear(a, s[P + "self_attn.q_proj.weight"], s[P + "self_attn.q_proj.bias"]) * 0.125; r["q"] = q
        k = F.linear(kv_in, s[P + "self_attn.k_proj.weight"], s[P + "self_attn.k_proj.bias"]); r["k"] = k
        v = F.linear(kv_in, s[P + "self_attn.v_proj.weight"], s[P + "self_attn.v_proj.bias"]); r["v"] = v
        qh, kh, vh = (t.view(B, -1, 8, 64).transpose(1, 2) for t in (q, k, v))
        ctx = (torch.softmax(qh @ kh.transpose(-1, -2), -1) @ vh).transpose(1, 2).reshape(B, M, 512); r["ctx"] = ctx
        mm = mem + F.linear(ctx, s[P + "self_attn.out_proj.weight"], s[P + "self_attn.out_proj.bias"]); r["mm"] = mm
        b_ = ln(mm, s, P + "final_layer_norm"); r["b"] = b_
        z = F.linear(b_, s[P + "fc1.weight"], s[P + "fc1.bias"]); r["z"] = z
        mem = mm + F.linear(torch.relu(z), s[P + "fc2.weight"], s[P + "fc2.bias"]); r["out"] = mem
        for t in r.values():
            t.retain_grad()
        rec.append(r)
    (mem.transpose(0, 1) * R.to(dt)).sum().backward()
    return rec, h_enc.grad
r64, hg64 = run(torch.float64)
r32, hg32 = run(torch.float32)
rl = lambda a, b: float((a.double().cpu().reshape(b.shape) - b).norm() / b.norm().clamp_min(1e-30))
print("dh_enc ours %.2e t32 %.2e" % (rl(step.dbg["h_enc"][:B * g.T2a].view(B, g.T2a, 512)[:, :g.T2], hg64), rl(hg32, hg64)))
for l in (2, 1, 0):
    d = step.dbgm[l]
    def kvrows(t):
        return t[:B * g.T2a].view(B, g.T2a, -1)[:, :g.T2]
    pairs = [("out", d["d_out"]), ("z(relu-bwd in: dh)", None), ("b", d["db_in"]), ("mm", d["dmm"]), ("ctx", d["dctx"]), ("q", d["dq"]),
             ("k", kvrows(d["dkv"])[..., :512]), ("v", kvrows(d["dkv"])[..., 512:]), ("a", d["da"]), ("kv_in", kvrows(d["dkv_in"]))]
    for name, ours in pairs:
        if ours is None:
            continue
        print("layer %d d%-6s ours %.2e  t32 %.2e   |ref| %.2e" % (l, name, rl(ours, r64[l][name].grad), rl(r32[l][name].grad, r64[l][name].grad),
                                                                  float(r64[l][name].grad.norm())))
    zo = T["mem"][l]["z"].cpu().double().view(B, M, -1)
    flips = int(((zo > 0) != (r64[l]["z"] > 0)).sum())
    print("layer %d relu sign flips vs fp64: %d of %d; fwd z err %.2e" % (l, flips, zo.numel(), rl(zo, r64[l]["z"].detach())))
