"""TEST INFRASTRUCTURE ONLY -- beam-search goldens (groundwork for a GPU beam search; fairseq-interactive's default is
--beam 5): the UNMODIFIED reference model + SequenceGenerator(beam_size=5, max_len_a=0, max_len_b=30) on the seeded
synthetic weights / inputs; the EOS row of the output embedding is scaled (x3) so that hypotheses end at scattered steps
and the finalisation / candidate bookkeeping is exercised (random-init hypotheses otherwise all run into max_len).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden_beam      ->  tests/golden/beam.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chimera_st_b200  # noqa: E402,F401
from chimera_st_b200 import synth  # noqa: E402
from oracle.gen_golden_greedy import build_reference_model, CASES  # noqa: E402

BEAM, MAX_LEN_B, EOS_SCALE = 5, 30, 3.0


def main():
    model, d = build_reference_model()                          # activates the import overlay
    from fairseq.sequence_generator import SequenceGenerator
    with torch.no_grad():
        model.decoder.embed_tokens.weight[2] *= EOS_SCALE          # tied with output_projection.weight
    assert model.decoder.output_projection.weight.data_ptr() == model.decoder.embed_tokens.weight.data_ptr()
    gen = SequenceGenerator([model], d, beam_size=BEAM, max_len_a=0, max_len_b=MAX_LEN_B)
    out = {"beam": BEAM, "max_len_b": MAX_LEN_B, "eos_scale": EOS_SCALE, "decoder_seed": 1}
    for name, (lens, seed) in CASES.items():
        wave, tl = synth.make_waveforms(lens, seed=seed)
        sample = {"net_input": {"src_tokens": wave, "src_lengths": tl}, "id": torch.arange(len(lens))}
        with torch.no_grad():
            hyp = gen.generate([model], sample)
            mem = model.encoder(wave, tl).encoder_out
        B = len(hyp)
        L = max(len(h["tokens"]) for hs in hyp for h in hs)
        toks = np.full((B, BEAM, L), -1, dtype=np.int64)
        ps = np.zeros((B, BEAM, L), dtype=np.float32)
        sc = np.full((B, BEAM), np.nan, dtype=np.float64)
        for b, hs in enumerate(hyp):
            assert len(hs) <= BEAM
            for k, h in enumerate(hs):
                n = len(h["tokens"])
                toks[b, k, :n] = h["tokens"].numpy()
                ps[b, k, :n] = h["positional_scores"].numpy()
                sc[b, k] = float(h["score"])
        out[name + "_tokens"], out[name + "_pos_scores"], out[name + "_scores"] = toks, ps, sc
        out[name + "_memories"] = mem.numpy()
        print(name, [[len(h["tokens"]) for h in hs] for hs in hyp], sc[:, 0])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "beam.npz"), **out)
    print("wrote beam.npz")


if __name__ == "__main__":
    main()
