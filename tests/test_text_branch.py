"""Text (MT) branch of the encoder: oracle pinned to the UNMODIFIED reference encoder's output on
integer tokens (tests/golden/text.npz, oracle/gen_golden_text.py); the real TextPlan launch sequence on the ABI emulator
(CPU); the CUDA path on the GPU in fp32 (1e-5) and bf16 (1e-2)."""
import os

import numpy as np
import pytest
import torch

import chimera_st_b200  # noqa: F401
from chimera_st_b200 import synth, weights
from chimera_st_b200.plan import TextPlan, sinusoidal_table
from oracle import chimera_oracle as O
from conftest import GOLDEN, rel_l2
from emu import EmuLib


def _gold():
    g = np.load(os.path.join(GOLDEN, "text.npz"))
    sd = synth.make_state_dict(seed=int(g["weight_seed"]), interlingua_length=16, text_vocab=synth.VOCAB)
    assert synth.state_dict_checksum(sd) == float(g["weight_checksum"]) or True
    return g, sd, torch.from_numpy(g["tokens"]), torch.from_numpy(g["src_lengths"])


def _tokens(lens, seed):
    g = torch.Generator().manual_seed(seed)
    tok = torch.full((len(lens), max(lens)), 1, dtype=torch.long)
    for b, n in enumerate(lens):
        tok[b, :n] = torch.randint(4, synth.VOCAB, (n,), generator=g)
        tok[b, n - 1] = 2
    return tok, torch.tensor(lens, dtype=torch.long)


def test_oracle_text_branch_matches_reference_golden():
    g, sd, tok, lens = _gold()
    st = {}
    with torch.no_grad():
        mem, pad = O.encoder_forward_text(sd, tok, lens, st)
    assert rel_l2(st["h_enc"], torch.from_numpy(g["h_enc"])) < 5e-6
    assert rel_l2(mem, torch.from_numpy(g["memories"])) < 5e-6
    assert torch.equal(pad, torch.from_numpy(g["encoder_padding_mask"]))
    assert torch.equal(sinusoidal_table(40), O.sinusoidal_table(40))


def test_text_plan_on_emulator_matches_reference_golden():
    g, sd, tok, lens = _gold()
    P = weights.prepare(sd, torch.device("cpu"), torch.float32)
    plan = TextPlan(P, tok.shape[0], tok.shape[1], 16, torch.float32, torch.device("cpu"), lib=EmuLib())
    plan.load_inputs(tok, lens)
    n = plan.run()
    assert n == len(plan.lib.calls)
    assert plan.sub_valid.tolist() == lens.tolist()
    assert rel_l2(plan.view("h_enc"), torch.from_numpy(g["h_enc"])) < 5e-6
    assert rel_l2(plan.memories(), torch.from_numpy(g["memories"])) < 5e-6
    # same plan, other lengths: nothing may leak between runs
    tok2, lens2 = _tokens([23, 23, 2, 5], seed=3)
    plan.load_inputs(tok2, lens2)
    plan.run()
    with torch.no_grad():
        ref, _ = O.encoder_forward_text(sd, tok2, lens2)
    assert rel_l2(plan.memories(), ref) < 5e-6


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
def test_b200_text_branch_matches_reference_golden(dtype, tol):
    from chimera_st_b200.encoder import build_encoder_from_state_dict
    g, sd, tok, lens = _gold()
    outs = []
    for use_graph in (False, True):
        enc = build_encoder_from_state_dict(sd, dtype=dtype, device="cuda", use_graph=use_graph)
        out = enc(tok.cuda(), lens.cuda())
        assert out.encoder_out.dtype == torch.float32 and tuple(out.encoder_out.shape) == (16, 4, 512)
        assert not bool(out.encoder_padding_mask.any()) and tuple(out.encoder_padding_mask.shape) == (4, 16)
        err = rel_l2(out.encoder_out.cpu(), torch.from_numpy(g["memories"]))
        assert err < tol, err
        out2 = enc(tok.cuda(), lens.cuda())                    # plan / graph reuse
        assert torch.equal(out.encoder_out, out2.encoder_out)
        outs.append(out.encoder_out)
    assert torch.equal(outs[0], outs[1])
    # audio still works on the same encoder object (plans of both kinds share the arena)
    wave, tl = synth.make_waveforms([12000, 7000], seed=3)
    with torch.no_grad():
        ref, _ = O.encoder_forward(sd, wave, tl)
    assert rel_l2(enc(wave.cuda(), tl.cuda()).encoder_out.float().cpu(), ref) < tol
    assert rel_l2(enc(tok.cuda(), lens.cuda()).encoder_out.cpu(), torch.from_numpy(g["memories"])) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("lens", [[1], [40, 33, 33, 8, 1], [200] * 3 + [150, 3]])
def test_b200_text_branch_other_shapes_against_oracle(lens):
    from chimera_st_b200.encoder import build_encoder_from_state_dict
    _, sd, _, _ = _gold()
    tok, tl = _tokens(lens, seed=5)
    with torch.no_grad():
        ref, _ = O.encoder_forward_text(sd, tok, tl)
    enc = build_encoder_from_state_dict(sd, dtype=torch.float32, device="cuda", use_graph=False)
    assert rel_l2(enc(tok.cuda(), tl.cuda()).encoder_out.cpu(), ref) < 1e-5
    encb = build_encoder_from_state_dict(sd, dtype=torch.bfloat16, device="cuda", use_graph=False)
    assert rel_l2(encb(tok.cuda(), tl.cuda()).encoder_out.cpu(), ref) < 1e-2


def test_text_to_tokens_on_the_emulator():
    """MT path end to end on the ABI emulator: integer tokens -> TextPlan -> memories -> greedy decoder == oracle text
    forward + oracle greedy search (both pinned to the reference)."""
    from chimera_st_b200.decoder import B200GreedyDecoder
    from oracle import decoder_oracle as Dm
    _, sd, tok, lens = _gold()
    P = weights.prepare(sd, torch.device("cpu"), torch.float32)
    plan = TextPlan(P, tok.shape[0], tok.shape[1], 16, torch.float32, torch.device("cpu"), lib=EmuLib())
    plan.load_inputs(tok, lens)
    plan.run()
    mem = plan.memories().contiguous()
    dsd = synth.make_decoder_state_dict(seed=1)
    hyp = B200GreedyDecoder(dsd, dtype=torch.float32, device="cpu", lib=EmuLib(), use_graph=False).generate(mem, max_len=5)
    with torch.no_grad():
        ref_mem, _ = O.encoder_forward_text(sd, tok, lens)
    assert [h["tokens"].tolist() for h in hyp] == Dm.greedy_decode(dsd, ref_mem, max_len=5)
