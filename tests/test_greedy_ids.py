"""North-star bar "identical greedy-decoded token IDs on a fixed synthetic set", checked downstream of the encoder.

Goldens: the UNMODIFIED reference model + its own SequenceGenerator(beam_size=1) (oracle/gen_golden_greedy.py).
Here the decoder is the ORACLE's (the B200 decoder has its own tests in test_gpu_decoder.py): the oracle's greedy decoder (oracle/decoder_oracle.py, pinned here against
the goldens on CPU) is run on the memories produced by the B200 encoder; a teacher-forced log-probability probe on a
fixed random target is compared as well, because random-init greedy output is nearly constant per utterance."""
import os

import numpy as np
import pytest
import torch

import chimera_st_b200  # noqa: F401
from chimera_st_b200 import synth
from oracle import chimera_oracle as O, decoder_oracle as Dm
from conftest import GOLDEN

CASES = {"tiny": ([16000, 12345, 8000], 7), "c1mix": ([80000, 64000, 48123, 32000], 1234)}


def _gold():
    return np.load(os.path.join(GOLDEN, "greedy.npz"))


def _check(memories, name, g, lp_tol):
    dsd = synth.make_decoder_state_dict(seed=int(g["decoder_seed"]))
    toks, margins = Dm.greedy_decode(dsd, memories, max_len=int(g["max_len_b"]), return_margins=True)
    gold = g[name + "_tokens"]
    for b, t in enumerate(toks):
        assert t == [x for x in gold[b].tolist() if x >= 0], (name, b, margins[b])
    # teacher-forced probe: top-8 log-probs at 12 positions
    prev = torch.from_numpy(g[name + "_tf_prev"])
    ids, lps = torch.from_numpy(g[name + "_tf_top_ids"]), torch.from_numpy(g[name + "_tf_top_lp"])
    worst = 0.0
    with torch.no_grad():
        for t in range(prev.shape[1]):
            lp = torch.log_softmax(Dm.decoder_logits(dsd, prev[:, :t + 1], memories).float(), -1)
            worst = max(worst, float((lp.gather(1, ids[:, t]) - lps[:, t]).abs().max()))
            # the arg-max must agree wherever the reference's own top-1/top-2 gap exceeds the log-prob tolerance
            clear = (lps[:, t, 0] - lps[:, t, 1]) > 2 * lp_tol
            assert torch.equal(lp.argmax(-1)[clear], ids[:, t, 0][clear]), (name, t)
    assert worst < lp_tol, (name, worst)
    return margins, worst


@pytest.mark.parametrize("name", ["tiny", "c1mix"])
def test_oracle_decoder_reproduces_reference_generator(name):
    g = _gold()
    lens, seed = CASES[name]
    wave, tl = synth.make_waveforms(lens, seed=seed)
    with torch.no_grad():
        mem, _ = O.encoder_forward(synth.make_state_dict(seed=0), wave, tl)
    _check(mem, name, g, 2e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,lp_tol", [(torch.float32, 2e-4), (torch.bfloat16, 6e-2)])
@pytest.mark.parametrize("name", ["tiny", "c1mix"])
def test_b200_memories_give_identical_greedy_ids(name, dtype, lp_tol):
    from chimera_st_b200.encoder import build_encoder_from_state_dict
    g = _gold()
    lens, seed = CASES[name]
    wave, tl = synth.make_waveforms(lens, seed=seed)
    enc = build_encoder_from_state_dict(synth.make_state_dict(seed=0), dtype=dtype, device="cuda", use_graph=False)
    mem = enc(wave.cuda(), tl.cuda()).encoder_out.cpu()
    margins, worst = _check(mem, name, g, lp_tol)
    print("greedy ids identical; min top-1/top-2 margins", ["%.3f" % m for m in margins], "max |dlogp|", worst)
