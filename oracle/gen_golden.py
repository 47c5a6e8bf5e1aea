"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED
reference encoder (through oracle/make_overlay.py) on seeded synthetic inputs with the
seeded synthetic weights of `chimera_st_b200.synth` (loaded with strict=True, which also
pins the state-dict key layout).  Run in the dev container only:

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden

Golden cases (weights seed 0 unless stated; waveforms `synth.make_waveforms`):
  tiny   : M=16, lens [16000,12345,8000]            all stage tensors (fp32) + fp64 memories
  tiny64 : M=64, same audio                         memories
  c1     : M=16, B=4 x 80000 (BASELINE configs[0])  memories + strided stage samples
  c1mix  : M=16, lens [80000,64000,48123,32000]     memories, masks, lengths
  lengths: exhaustive (len, L) grid of the frame-mask rule and subsampler lengths (INT)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chimera_st_b200  # noqa: E402
from chimera_st_b200 import synth  # noqa: E402
from oracle.ref_model import build_reference_encoder  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def run_reference(enc, wave, lens, want_stages):
    """Reference forward + stage taps via forward hooks / the reference's own switches."""
    st = {}
    hooks = []
    w2v = enc.wav2vec_model
    if want_stages:
        def tap(name, fn):
            def hook(m, i, o):
                st[name] = fn(i, o)          # returns None: a hook's return value would replace the output
            return hook
        hooks.append(w2v.feature_extractor.register_forward_hook(
            tap("conv_feats", lambda i, o: o.detach().clone())))
        hooks.append(w2v.encoder.layers[0].register_forward_hook(
            tap("w2v_l0", lambda i, o: o[0].detach().transpose(0, 1).clone())))
        hooks.append(w2v.encoder.layers[0].register_forward_hook(
            tap("w2v_in", lambda i, o: i[0].detach().transpose(0, 1).clone())))
        hooks.append(enc.subsample.register_forward_hook(
            tap("sub_out_raw", lambda i, o: o[0].detach().transpose(0, 1).clone())))
        hooks.append(enc.subsample.register_forward_hook(
            tap("sub_len", lambda i, o: o[1].detach().clone())))
    with torch.no_grad():
        feat, fmask, flen = enc._get_w2v_feature(wave, lens)
        st["w2v_out"], st["frame_mask"], st["w2v_len"] = feat, fmask, flen
        enc.no_interlingua = True                      # w2v2_transformer_interlingua.py:260-262 -> h_enc
        st["h_enc"] = enc(wave, lens).encoder_out.transpose(0, 1).contiguous()
        enc.no_interlingua = False
        out = enc(wave, lens)
        st["memories"] = out.encoder_out
        st["encoder_padding_mask"] = out.encoder_padding_mask
    for h in hooks:
        h.remove()
    if "sub_out_raw" in st:
        st["sub_out"] = st.pop("sub_out_raw") * enc.embed_scale
    return st


def npify(d):
    return {k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    enc, _ = build_reference_encoder(16)
    sd = synth.make_state_dict(seed=0, interlingua_length=16)
    missing = enc.load_state_dict(sd, strict=True)
    print("strict load ok", missing)
    ck = synth.state_dict_checksum(sd)

    # ---- tiny: every stage -------------------------------------------------------------
    lens = [16000, 12345, 8000]
    wave, tl = synth.make_waveforms(lens, seed=7)
    st = run_reference(enc, wave, tl, True)
    enc64 = enc.double()
    with torch.no_grad():
        mem64 = enc64(wave.double(), tl).encoder_out
    enc.float()
    enc.load_state_dict(sd, strict=True)               # restore exact fp32 weights
    out = npify(st)
    out.update(memories_f64=mem64.numpy(), src_lengths=np.asarray(lens), wave_seed=7, weight_seed=0,
               weight_checksum=ck, interlingua_length=16)
    np.savez_compressed(os.path.join(GOLD, "tiny.npz"), **out)
    print("tiny", {k: getattr(v, "shape", v) for k, v in out.items()})

    # ---- c1 ------------------------------------------------------------------------------
    for name, lens in (("c1", [80000] * 4), ("c1mix", [80000, 64000, 48123, 32000])):
        wave, tl = synth.make_waveforms(lens, seed=1234)
        st = run_reference(enc, wave, tl, True)
        out = dict(memories=st["memories"].numpy(), frame_mask=st["frame_mask"].numpy(),
                   w2v_len=st["w2v_len"].numpy(), sub_len=st["sub_len"].numpy(),
                   encoder_padding_mask=st["encoder_padding_mask"].numpy(),
                   conv_feats_s=st["conv_feats"][:, ::37, ::11].numpy(),
                   w2v_in_s=st["w2v_in"][:, ::11, ::37].numpy(),
                   w2v_out_s=st["w2v_out"][:, ::11, ::37].numpy(),
                   h_enc_s=st["h_enc"][:, ::3, ::17].numpy(),
                   src_lengths=np.asarray(lens), wave_seed=1234, weight_seed=0, weight_checksum=ck,
                   interlingua_length=16)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        print(name, {k: getattr(v, "shape", v) for k, v in out.items()})

    # ---- exhaustive integer grid through the reference's own tensor code ---------------------
    from fairseq.data.data_utils import lengths_to_padding_mask
    rows = []
    for L in list(range(400, 1700, 13)) + [16000, 32000, 48123, 80000, 150000, 240000, 320000, 480000]:
        T = L
        for _, k, s in synth.CONV_LAYERS:
            T = (T - k) // s + 1
        cand = sorted(set([1, 2, L // 3, L // 2, L - 1, L] + list(range(max(1, L - 700), L, 97))
                          + [min(L, x) for x in (319, 320, 321, 639, 640, 641, 4000)]))
        lens_t = torch.tensor(cand + [L])
        pm = lengths_to_padding_mask(lens_t)
        extra = pm.size(1) % T                                   # wav2vec2.py:543-548 verbatim
        if extra > 0:
            pm = pm[:, :-extra]
        pm = pm.view(pm.size(0), T, -1).all(-1)
        valid = (1 - pm.int()).sum(1)
        sub = enc.subsample.get_out_seq_lens_tensor(valid)
        for n, v, s2 in zip(lens_t.tolist(), valid.tolist(), sub.tolist()):
            rows.append((L, n, T, v, s2))
    np.savez_compressed(os.path.join(GOLD, "lengths.npz"), rows=np.asarray(rows, dtype=np.int64))
    print("lengths rows", len(rows))

    # ---- tiny64 ------------------------------------------------------------------------------
    enc, _ = build_reference_encoder(64)
    sd64 = synth.make_state_dict(seed=0, interlingua_length=64)
    enc.load_state_dict(sd64, strict=True)
    wave, tl = synth.make_waveforms([16000, 12345, 8000], seed=7)
    with torch.no_grad():
        m = enc(wave, tl).encoder_out
    np.savez_compressed(os.path.join(GOLD, "tiny64.npz"), memories=m.numpy(),
                        src_lengths=np.asarray([16000, 12345, 8000]), wave_seed=7, weight_seed=0,
                        weight_checksum=synth.state_dict_checksum(sd64), interlingua_length=64)
    print("tiny64", m.shape)


if __name__ == "__main__":
    main()
