"""Pins the oracle restatement (oracle/chimera_oracle.py) against outputs of the UNMODIFIED
reference (tests/golden/*.npz, written by oracle/gen_golden.py in the dev container)."""
import os

import numpy as np
import pytest
import torch

import chimera_st_b200  # noqa: F401
from chimera_st_b200 import synth, lengths
from oracle import chimera_oracle as O
from conftest import rel_l2, rel_max, GOLDEN

# the reference's own fp32 noise floor is ~5e-7 rel-L2; the oracle is
# the same arithmetic in a different op order, so it must sit inside a few of those floors
TOL_L2 = 5e-6


def _load(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    M = int(g["interlingua_length"])
    sd = synth.make_state_dict(seed=int(g["weight_seed"]), interlingua_length=M)
    assert abs(synth.state_dict_checksum(sd) - float(g["weight_checksum"])) < 1e-6 * abs(float(g["weight_checksum"])), \
        "torch RNG stream differs from the one the goldens were generated with"
    wave, lens = synth.make_waveforms(g["src_lengths"].tolist(), seed=int(g["wave_seed"]))
    return g, sd, wave, lens


@pytest.fixture(scope="module")
def tiny():
    g, sd, wave, lens = _load("tiny")
    st = {}
    with torch.no_grad():
        mem, pm = O.encoder_forward(sd, wave, lens, stages=st)
    return g, sd, wave, lens, st, mem, pm


@pytest.mark.parametrize("stage", ["conv_feats", "w2v_in", "w2v_l0", "w2v_out", "sub_out", "h_enc"])
def test_tiny_stage(tiny, stage):
    g, _, _, _, st, _, _ = tiny
    ref = torch.from_numpy(g[stage])
    assert st[stage].shape == ref.shape
    assert rel_l2(st[stage], ref) < TOL_L2, (stage, rel_l2(st[stage], ref))
    assert rel_max(st[stage], ref) < 10 * TOL_L2


def test_tiny_memories_and_masks(tiny):
    g, _, _, _, st, mem, pm = tiny
    assert rel_l2(mem, torch.from_numpy(g["memories"])) < TOL_L2
    assert rel_l2(mem, torch.from_numpy(g["memories_f64"])) < TOL_L2
    assert torch.equal(st["frame_mask"], torch.from_numpy(g["frame_mask"]))
    assert torch.equal(st["w2v_len"], torch.from_numpy(g["w2v_len"]))
    assert torch.equal(st["sub_len"], torch.from_numpy(g["sub_len"]))
    assert torch.equal(pm, torch.from_numpy(g["encoder_padding_mask"]))
    assert pm.dtype == torch.bool and pm.shape == (3, 16) and not pm.any()


def test_literal_memory_form_equals_cross_attention_form(tiny):
    """cat(h_enc, mem) + additive -1e8 mask == M-query cross attention."""
    _, sd, _, _, st, _, _ = tiny
    with torch.no_grad():
        a = O.memory_stage_literal(sd, st["h_enc"])
        b = O.memory_stage(sd, st["h_enc"])
    assert rel_l2(a, b) < 2e-6


def test_tiny_fp64_oracle_matches_fp64_reference(tiny):
    g, sd, wave, lens, *_ = tiny
    with torch.no_grad():
        mem, _ = O.encoder_forward(O.cast_state_dict(sd, torch.float64), wave.double(), lens)
    # not ~1e-15: the reference's "fp64" run still does GroupNorm and the wav2vec2 FFN GELU in
    # fp32 (fp32_group_norm.py:18, gelu.py:25 `.float()`), the fp64 oracle does not
    assert rel_l2(mem, torch.from_numpy(g["memories_f64"])) < 1e-6


def test_tiny64():
    g, sd, wave, lens = _load("tiny64")
    with torch.no_grad():
        mem, pm = O.encoder_forward(sd, wave, lens)
    assert mem.shape == (64, 3, 512) and pm.shape == (3, 64)
    assert rel_l2(mem, torch.from_numpy(g["memories"])) < TOL_L2


@pytest.mark.parametrize("name", ["c1", "c1mix"])
def test_c1(name):
    g, sd, wave, lens = _load(name)
    st = {}
    with torch.no_grad():
        mem, pm = O.encoder_forward(sd, wave, lens, stages=st)
    assert rel_l2(mem, torch.from_numpy(g["memories"])) < TOL_L2
    assert torch.equal(st["frame_mask"], torch.from_numpy(g["frame_mask"]))
    assert torch.equal(st["w2v_len"], torch.from_numpy(g["w2v_len"]))
    assert torch.equal(st["sub_len"], torch.from_numpy(g["sub_len"]))
    assert rel_l2(st["conv_feats"][:, ::37, ::11], torch.from_numpy(g["conv_feats_s"])) < TOL_L2
    assert rel_l2(st["w2v_in"][:, ::11, ::37], torch.from_numpy(g["w2v_in_s"])) < TOL_L2
    assert rel_l2(st["w2v_out"][:, ::11, ::37], torch.from_numpy(g["w2v_out_s"])) < TOL_L2
    assert rel_l2(st["h_enc"][:, ::3, ::17], torch.from_numpy(g["h_enc_s"])) < TOL_L2
    if name == "c1mix":   # the survey's probe: ceil-style frame lengths
        assert st["w2v_len"].tolist() == [249, 200, 150, 100]


def test_integer_rules_exhaustive():
    """Closed-form host formulas AND the oracle's tensor rule vs the reference's own tensor code."""
    rows = np.load(os.path.join(GOLDEN, "lengths.npz"))["rows"]
    by_L = {}
    for L, n, T, v, s2 in rows.tolist():
        by_L.setdefault(L, []).append((n, T, v, s2))
    for L, items in by_L.items():
        T = items[0][1]
        assert lengths.conv_out_lengths(L)[-1] == T == O.conv_out_lengths(L)[-1]
        ns = [n for n, *_ in items]
        assert max(ns) == L
        got = lengths.frame_valid_counts(ns, L)
        assert got == [v for _, _, v, _ in items], L
        assert [lengths.subsampler_len(v) for v in got] == [s for *_, s in items]
        fm = O.frame_padding_mask(torch.tensor(ns), T)
        assert (~fm).sum(1).tolist() == got
        assert O.subsampler_lengths(torch.tensor(got)).tolist() == [s for *_, s in items]


def test_base_encoder_restatement_matches_the_reference_base_class():
    """oracle.base_encoder_forward (S2T_W2V2_TransformerEncoder.forward, w2v2_transformer.py:338-386) against the unmodified
    reference (tests/golden/base_encoder.npz, oracle/gen_golden_base.py)."""
    g = np.load(os.path.join(GOLDEN, "base_encoder.npz"))
    sd = synth.make_state_dict(seed=0, interlingua_length=16)
    for name in ("tiny", "full"):
        wave, lens = synth.make_waveforms(g[name + "_lens"].tolist(), seed=int(g[name + "_seed"]))
        with torch.no_grad():
            out, pad = O.base_encoder_forward(sd, wave, lens)
        assert rel_l2(out, torch.from_numpy(g[name + "_encoder_out"])) < 5e-6
        assert (pad is not None) == bool(g[name + "_has_mask"])
        if pad is not None:
            assert np.array_equal(pad.numpy(), g[name + "_padding_mask"])
