"""Decode-step latency of the greedy decoder at the C4 shape (64 hypotheses, M = 64 memories, max_len_b = 200) and the
combined "encode + greedy decode" rate of BASELINE configs[3] on one GPU.

    python tools/dec_rate.py [--B 64] [--M 64] [--max-len 200] [--encode] [--json gpurun_out/dec_rate.json]

Prints: whole-decode ms and us/step (CUDA events around `generate`, graph replay), and the per-kernel-family split of
one eagerly launched step (CUDA events around every C-ABI call, GPU kept busy first so that launch gaps are not timed).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import chimera_st_b200  # noqa: E402,F401
from chimera_st_b200 import synth  # noqa: E402
from chimera_st_b200.decoder import B200GreedyDecoder  # noqa: E402


class Prof:
    def __init__(self, lib):
        self.lib, self.rec = lib, []

    def __getattr__(self, name):
        fn = getattr(self.lib, name)
        if not name.startswith("cst_dec_"):
            return fn

        def timed(*a):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = fn(*a)
            e.record()
            tag = name
            if name == "cst_dec_linear":
                p = a[0]._obj
                tag = "linear N=%d K=%d%s" % (p.N, p.K, " +LN" if p.ln_gamma else "")
            elif name == "cst_dec_attention":
                tag = "attention " + ("self" if a[13] else "memory")
            self.rec.append((tag, s, e))
            return rc
        return timed


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--M", type=int, default=64)
    ap.add_argument("--max-len", type=int, default=200)
    ap.add_argument("--encode", action="store_true")
    ap.add_argument("--lanes", default="1,2,4,8")
    ap.add_argument("--beam", type=int, default=0, help="also time B200BeamDecoder with this beam width (2..8)")
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    out = {"B": a.B, "M": a.M, "max_len": a.max_len}
    dsd = synth.make_decoder_state_dict(seed=1)
    mem32 = torch.randn(a.M, a.B, 512, generator=torch.Generator().manual_seed(5)).cuda()
    for dt in (torch.float32, torch.bfloat16):
        name = str(dt).replace("torch.", "")
        dec = B200GreedyDecoder(dsd, dtype=dt, device="cuda")
        mem = mem32.to(dt)
        by_lanes = {}
        for nl in [int(x) for x in a.lanes.split(",")][::-1]:
            dec.generate(mem, max_len=a.max_len, n_lanes=nl)       # warm-up + graph capture
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                hyp = dec.generate(mem, max_len=a.max_len, n_lanes=nl)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            by_lanes[nl] = round(sorted(ts)[1], 3)
        ms = by_lanes[1] if 1 in by_lanes else sorted(ts)[1]
        steps = dec.last_steps
        plan = dec._plan(a.B, a.M, a.max_len, mem.dtype)
        # instrumented eager step at a late step index (long self-attention cache)
        plan.begin(mem)
        plan.counters[0] = a.max_len - 1
        prof = Prof(plan.lib)
        real, plan.lib = plan.lib, prof
        torch.cuda._sleep(int(2e7))
        plan._step()
        torch.cuda.synchronize()
        plan.lib = real
        fam = {}
        for tag, s, e in prof.rec:
            d = fam.setdefault(tag, [0, 0.0])
            d[0] += 1
            d[1] += s.elapsed_time(e) * 1e3
        out[name] = {"decode_ms_by_lanes": by_lanes, "decode_ms": round(ms, 3), "steps": steps, "us_per_step": round(1e3 * ms / steps, 1),
                     "launches": dec.last_launches, "tokens_per_s": round(sum(len(h["tokens"]) for h in hyp) / (ms * 1e-3)),
                     "eager_step_us_by_kernel": {k: {"launches": v[0], "us": round(v[1], 1)} for k, v in fam.items()},
                     "eager_step_us": round(sum(v[1] for v in fam.values()), 1)}
        print(name, json.dumps(out[name]))
    if a.beam > 1:
        from chimera_st_b200.decoder import B200BeamDecoder
        for dt in (torch.float32, torch.bfloat16):
            bd = B200BeamDecoder(dsd, beam=a.beam, dtype=dt, device="cuda")
            mem = mem32.to(dt)
            bd.generate(mem, max_len=a.max_len)
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                bd.generate(mem, max_len=a.max_len)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            key = "beam%d_%s" % (a.beam, str(dt).replace("torch.", ""))
            out[key] = {"decode_ms": round(sorted(ts)[1], 3), "steps": bd.last_steps,
                        "us_per_step": round(1e3 * sorted(ts)[1] / max(1, bd.last_steps), 1), "rows": a.B * a.beam}
            print(key, json.dumps(out[key]))
    if a.encode:
        from chimera_st_b200.encoder import build_encoder_from_state_dict
        L = 320000
        enc = build_encoder_from_state_dict(synth.make_state_dict(seed=0, interlingua_length=a.M), dtype=torch.bfloat16, device="cuda")
        dec = B200GreedyDecoder(dsd, dtype=torch.bfloat16, device="cuda")
        wave, lens = synth.make_waveforms([L] * a.B, seed=1)
        wave, lens = wave.cuda(), lens.cuda()
        for _ in range(2):
            dec.generate(enc(wave, lens).encoder_out, max_len=a.max_len)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        mem = enc(wave, lens).encoder_out
        e1.record()
        dec.generate(mem, max_len=a.max_len)
        e2.record()
        torch.cuda.synchronize()
        t_enc, t_dec = e0.elapsed_time(e1), e1.elapsed_time(e2)
        out["c4_encode_decode_bf16"] = {"encode_ms": round(t_enc, 2), "decode_ms": round(t_dec, 2),
                                        "audio_s_per_s": round(a.B * L / 16000 / ((t_enc + t_dec) * 1e-3), 1),
                                        "encode_only_audio_s_per_s": round(a.B * L / 16000 / (t_enc * 1e-3), 1)}
        print("c4", json.dumps(out["c4_encode_decode_bf16"]))
    if a.json:
        os.makedirs(os.path.dirname(a.json) or ".", exist_ok=True)
        json.dump(out, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
