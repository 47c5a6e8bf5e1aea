"""Host logic of the greedy decoder without a GPU: the real launch sequence of `B200GreedyDecoder` (weight fusion,
cache / memory strides, device-side step protocol, EOS rules, hypothesis assembly) is driven against the host emulator
of the C ABI (tests/emu.py) and must reproduce the oracle's greedy search, which itself is pinned to the reference's
SequenceGenerator (tests/test_greedy_ids.py, tests/golden/greedy.npz)."""
import os

import numpy as np
import pytest
import torch

import chimera_st_b200  # noqa: F401
from chimera_st_b200 import synth
from chimera_st_b200.decoder import B200GreedyDecoder, sinusoidal_positions, EOS
from oracle import decoder_oracle as Dm
from conftest import GOLDEN
from emu import EmuLib


def test_position_table_matches_oracle():
    t = Dm.sinusoidal_table(2 + 40)[2:]
    assert torch.equal(sinusoidal_positions(40), t)


@pytest.mark.parametrize("max_len,min_len", [(6, 1), (3, 4)])
def test_emulated_decoder_reproduces_oracle_greedy(max_len, min_len):
    g = np.load(os.path.join(GOLDEN, "tiny.npz"))
    mem = torch.from_numpy(g["memories"])                      # [16, 3, 512] reference memories
    dsd = synth.make_decoder_state_dict(seed=1)
    dec = B200GreedyDecoder(dsd, dtype=torch.float32, device="cpu", lib=EmuLib(), use_graph=False)
    hyp = dec.generate(mem, max_len=max_len, min_len=min_len)
    ref = Dm.greedy_decode(dsd, mem, max_len=max_len, min_len=min_len)
    assert [h["tokens"].tolist() for h in hyp] == ref
    for h in hyp:
        assert h["tokens"][-1] == EOS and len(h["positional_scores"]) == len(h["tokens"])
        assert abs(h["score"] - float(h["positional_scores"].mean())) < 1e-6
    assert dec.last_launches == 6 + dec.last_steps * 51


def test_golden_reference_tokens_via_emulator():
    """Full-length (max_len_b = 50) decode of the reference's own memories gives the reference generator's tokens."""
    g, gg = np.load(os.path.join(GOLDEN, "tiny.npz")), np.load(os.path.join(GOLDEN, "greedy.npz"))
    dsd = synth.make_decoder_state_dict(seed=int(gg["decoder_seed"]))
    dec = B200GreedyDecoder(dsd, dtype=torch.float32, device="cpu", lib=EmuLib(), use_graph=False)
    hyp = dec.generate(torch.from_numpy(g["memories"]), max_len=int(gg["max_len_b"]))
    gold = gg["tiny_tokens"]
    for b, h in enumerate(hyp):
        assert h["tokens"].tolist() == [x for x in gold[b].tolist() if x >= 0]


def eos_happy_decoder(scale=4.0):
    """Decoder weights whose EOS output row is scaled up: hypotheses end at scattered steps instead of running into max_len."""
    dsd = synth.make_decoder_state_dict(seed=1)
    E = dsd["decoder.embed_tokens.weight"].clone()
    E[EOS] *= scale
    dsd["decoder.embed_tokens.weight"] = E
    dsd["decoder.output_projection.weight"] = E
    return dsd


def test_rows_finish_independently_and_the_loop_stops_early():
    dsd = eos_happy_decoder(8.0)
    mem = torch.randn(16, 6, 512, generator=torch.Generator().manual_seed(21))
    dec = B200GreedyDecoder(dsd, dtype=torch.float32, device="cpu", lib=EmuLib(), use_graph=False)
    hyp = dec.generate(mem, max_len=40, poll=4)
    ref = Dm.greedy_decode(dsd, mem, max_len=40)
    assert [h["tokens"].tolist() for h in hyp] == ref
    lens = [len(r) for r in ref]
    assert min(lens) < 5 and any(5 < n < 41 for n in lens) and max(lens) == 41, lens     # early, middle and forced endings in one batch
    # only the early rows: the loop stops at the first poll after the last EOS instead of running to max_len
    early = [b for b, n in enumerate(lens) if n < 5]
    hyp = dec.generate(mem[:, early].contiguous(), max_len=40, poll=4)
    assert [h["tokens"].tolist() for h in hyp] == [ref[b] for b in early]
    assert dec.last_steps == 4


def test_cpu_device_without_emulator_raises():
    from chimera_st_b200._lib import CstError
    with pytest.raises(CstError):
        B200GreedyDecoder(synth.make_decoder_state_dict(seed=1), device="cpu")
