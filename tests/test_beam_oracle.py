"""Groundwork for a GPU beam search (fairseq-interactive's default is --beam 5): the oracle's per-sentence restatement of
SequenceGenerator._generate / BeamSearch.step / finalize_hypos is pinned to the UNMODIFIED reference generator
(tests/golden/beam.npz, oracle/gen_golden_beam.py: beam 5, EOS row scaled so that hypotheses end at scattered steps)."""
import os

import numpy as np
import pytest
import torch

import chimera_st_b200  # noqa: F401
from chimera_st_b200 import synth
from oracle import decoder_oracle as Dm
from conftest import GOLDEN


@pytest.mark.parametrize("name", ["tiny", "c1mix"])
def test_oracle_beam_search_reproduces_reference_generator(name):
    g = np.load(os.path.join(GOLDEN, "beam.npz"))
    dsd = synth.make_decoder_state_dict(seed=int(g["decoder_seed"]))
    E = dsd["decoder.embed_tokens.weight"].clone()
    E[2] *= float(g["eos_scale"])
    dsd["decoder.embed_tokens.weight"] = dsd["decoder.output_projection.weight"] = E
    mem = torch.from_numpy(g[name + "_memories"])
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    hyps = Dm.beam_search(dsd, mem, beam=int(g["beam"]), max_len=int(g["max_len_b"]))
    toks, ps, sc = g[name + "_tokens"], g[name + "_pos_scores"], g[name + "_scores"]
    lens = set()
    for b, hs in enumerate(hyps):
        assert len(hs) == int((~np.isnan(sc[b])).sum())
        for k, h in enumerate(hs):
            want = [x for x in toks[b, k].tolist() if x >= 0]
            assert h["tokens"].tolist() == want, (name, b, k)
            assert abs(h["score"] - sc[b, k]) < 2e-4
            assert np.abs(h["positional_scores"].numpy() - ps[b, k, :len(want)]).max() < 2e-4
            lens.add(len(want))
    assert len(lens) > 3                          # early, middle and max-length endings are all exercised


@pytest.mark.parametrize("name", ["tiny", "c1mix"])
def test_beam_decoder_launch_sequence_on_the_emulator_reproduces_the_reference(name):
    """B200BeamDecoder (experimental): the real launch sequence -- double-buffered token / score / history tables, cache
    history instead of cache re-ordering, device-side finalisation -- on the host emulator of the C ABI."""
    from chimera_st_b200.decoder import B200BeamDecoder
    from emu import EmuLib
    g = np.load(os.path.join(GOLDEN, "beam.npz"))
    dsd = synth.make_decoder_state_dict(seed=int(g["decoder_seed"]))
    E = dsd["decoder.embed_tokens.weight"].clone()
    E[2] *= float(g["eos_scale"])
    dsd["decoder.embed_tokens.weight"] = dsd["decoder.output_projection.weight"] = E
    mem = torch.from_numpy(g[name + "_memories"])
    dec = B200BeamDecoder(dsd, beam=int(g["beam"]), dtype=torch.float32, device="cpu", lib=EmuLib(), use_graph=False)
    hyps = dec.generate(mem, max_len=int(g["max_len_b"]))
    toks, ps, sc = g[name + "_tokens"], g[name + "_pos_scores"], g[name + "_scores"]
    for b, hs in enumerate(hyps):
        assert len(hs) == int((~np.isnan(sc[b])).sum())
        for k, h in enumerate(hs):
            want = [x for x in toks[b, k].tolist() if x >= 0]
            assert h["tokens"].tolist() == want, (name, b, k)
            assert abs(h["score"] - sc[b, k]) < 2e-4
            assert np.abs(h["positional_scores"].numpy() - ps[b, k, :len(want)]).max() < 2e-4
