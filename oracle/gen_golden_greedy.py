"""TEST INFRASTRUCTURE ONLY -- goldens for the downstream check "identical greedy-decoded token IDs":
runs the UNMODIFIED reference model (encoder + TransformerDecoderScriptable) and its own SequenceGenerator
(beam_size=1, max_len_a=0, max_len_b=50) on the seeded synthetic weights/inputs, plus teacher-forced decoder
log-probabilities on a fixed random target (a sensitive probe: random-init greedy output is nearly constant).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden_greedy      ->  tests/golden/greedy.npz
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chimera_st_b200  # noqa: E402,F401
from chimera_st_b200 import synth  # noqa: E402
from oracle import make_overlay  # noqa: E402
from oracle.ref_model import W2V_CONV_SPEC  # noqa: E402

CASES = {"tiny": ([16000, 12345, 8000], 7), "c1mix": ([80000, 64000, 48123, 32000], 1234)}
MAX_LEN_B = 50


def build_reference_model():
    """-> (unmodified reference model with the seeded synthetic weights, eval mode; target dictionary)"""
    make_overlay.build()
    make_overlay.activate()
    import tempfile
    import fairseq.models  # noqa: F401
    from fairseq.data import Dictionary
    from fairseq.models.wav2vec import wav2vec2 as W
    from fairseq.models.chimera.w2v2_transformer_interlingua import S2TTransformerInterlinguaModelW2V2
    from fairseq.sequence_generator import SequenceGenerator
    torch.set_num_threads(8)
    d = Dictionary.load(os.path.join(make_overlay.ref_root(), "chimera/resources/wmt14-en-de-spm/spm_unigram10000_wave_joint.txt"))
    assert len(d) == synth.VOCAB and (d.bos(), d.pad(), d.eos(), d.unk()) == (0, 1, 2, 3)

    class Task:
        source_dictionary = None
        target_dictionary = d
    w2v_args = argparse.Namespace(conv_feature_layers=W2V_CONV_SPEC, quantize_targets=True, final_dim=256,
                                  encoder_layerdrop=0.05, dropout_input=0.1, dropout_features=0.1, feature_grad_mult=0.1)
    W.base_architecture(w2v_args)
    tmp = tempfile.NamedTemporaryFile(suffix=".pt", delete=False)
    tmp.close()
    torch.save({"args": w2v_args, "model": W.Wav2Vec2Model.build_model(w2v_args, task=None).state_dict()}, tmp.name)
    args = argparse.Namespace(w2v2_model_path=tmp.name, encoder_layers=6, encoder_embed_dim=512, interlingua_length=16,
                              interlingua_layers=3, interlingua_debug_options=[], dropout=0.1,
                              share_decoder_input_output_embed=True, max_source_positions=6000, max_target_positions=1024)
    model = S2TTransformerInterlinguaModelW2V2.build_model(args, Task())
    os.unlink(tmp.name)
    sd = {"encoder." + k: v for k, v in synth.make_state_dict(seed=0, interlingua_length=16).items()}
    sd.update(synth.make_decoder_state_dict(seed=1))
    # the reference pops this key unconditionally when the task has no source dictionary (interlingua:198-202)
    sd["encoder.text_embed_tokens.weight"] = torch.zeros(synth.VOCAB, 512)
    print("load:", model.load_state_dict(sd, strict=True))
    model.eval()
    return model, d


def main():
    model, d = build_reference_model()                          # activates the import overlay
    from fairseq.sequence_generator import SequenceGenerator
    gen = SequenceGenerator([model], d, beam_size=1, max_len_a=0, max_len_b=MAX_LEN_B)
    out = {"max_len_b": MAX_LEN_B, "decoder_seed": 1}
    g = torch.Generator().manual_seed(99)
    for name, (lens, seed) in CASES.items():
        wave, tl = synth.make_waveforms(lens, seed=seed)
        sample = {"net_input": {"src_tokens": wave, "src_lengths": tl}, "id": torch.arange(len(lens))}
        with torch.no_grad():
            hyp = gen.generate([model], sample)
        toks = [h[0]["tokens"].tolist() for h in hyp]
        print(name, [len(t) for t in toks], [t[:6] for t in toks])
        L = max(len(t) for t in toks)
        arr = np.full((len(toks), L), -1, dtype=np.int64)
        for i, t in enumerate(toks):
            arr[i, :len(t)] = t
        out[name + "_tokens"] = arr
        # teacher-forced probe: fixed random target of 12 tokens (ids >= 4), eos-prefixed like the reference's decoder input
        tgt = torch.randint(4, synth.VOCAB, (len(lens), 12), generator=g)
        prev = torch.cat((torch.full((len(lens), 1), 2), tgt[:, :-1]), 1)
        with torch.no_grad():
            dec_out, _ = model(wave, tl, prev)
            lp = torch.log_softmax(dec_out.float(), -1)
        top = lp.topk(8, dim=-1)
        out[name + "_tf_prev"] = prev.numpy()
        out[name + "_tf_top_ids"] = top.indices.numpy()
        out[name + "_tf_top_lp"] = top.values.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "greedy.npz"), **out)
    print("wrote greedy.npz")


if __name__ == "__main__":
    main()
