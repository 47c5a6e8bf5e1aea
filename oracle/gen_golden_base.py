"""TEST INFRASTRUCTURE ONLY -- golden outputs of the reference's NON-memory base encoder
(S2T_W2V2_TransformerEncoder.forward, fairseq/models/chimera/w2v2_transformer.py:338-386): the same wav2vec2 + subsampler +
6 shared layers + LayerNorm as the interlingua encoder's audio branch, but with sinusoidal positions added after the
subsampler and the real key-padding mask returned.  The interlingua encoder IS a subclass, so the base-class forward is run
UNMODIFIED on the reference module that carries the seeded synthetic weights.  Dev container only:

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden_base      -> tests/golden/base_encoder.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chimera_st_b200  # noqa: E402,F401
from chimera_st_b200 import synth  # noqa: E402
from oracle.ref_model import build_reference_encoder  # noqa: E402

CASES = {"tiny": ([16000, 12345, 8000], 11), "full": ([24000, 24000], 12)}


def main():
    torch.set_num_threads(8)
    enc, _ = build_reference_encoder(16)
    enc.load_state_dict(synth.make_state_dict(seed=0, interlingua_length=16), strict=True)
    from fairseq.models.chimera.w2v2_transformer import S2T_W2V2_TransformerEncoder
    out = {}
    for name, (lens, seed) in CASES.items():
        wave, tl = synth.make_waveforms(lens, seed=seed)
        with torch.no_grad():
            eo = S2T_W2V2_TransformerEncoder.forward(enc, wave, tl)
        out[name + "_lens"] = np.asarray(lens)
        out[name + "_seed"] = np.asarray(seed)
        out[name + "_encoder_out"] = eo.encoder_out.numpy()                                   # [T2, B, 512]
        out[name + "_has_mask"] = np.asarray(eo.encoder_padding_mask is not None)
        out[name + "_padding_mask"] = (eo.encoder_padding_mask.numpy() if eo.encoder_padding_mask is not None
                                       else np.zeros((len(lens), eo.encoder_out.shape[0]), dtype=bool))
        print(name, eo.encoder_out.shape, "mask" if eo.encoder_padding_mask is not None else "no mask")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "base_encoder.npz"), **out)


if __name__ == "__main__":
    main()
