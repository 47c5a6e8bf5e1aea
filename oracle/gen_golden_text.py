"""TEST INFRASTRUCTURE ONLY -- goldens for the TEXT (MT) branch of the encoder: the UNMODIFIED reference
S2T_W2V2_TransformerInterlinguaEncoder.forward on integer tokens (w2v2_transformer_interlingua.py:212-217,230-236),
seeded synthetic weights (incl. text_embed_tokens) and tokens.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden_text      ->  tests/golden/text.npz
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

LENS = [23, 17, 9, 1]
TOK_SEED = 11


def make_tokens(lens, seed, vocab=synth.VOCAB, pad=1, eos=2):
    """Length-sorted token batch as the collater builds it: ids in [4, vocab), EOS last, right-padded with pad."""
    g = torch.Generator().manual_seed(seed)
    T = max(lens)
    tok = torch.full((len(lens), T), pad, dtype=torch.long)
    for b, n in enumerate(lens):
        tok[b, :n] = torch.randint(4, vocab, (n,), generator=g)
        tok[b, n - 1] = eos
    return tok, torch.tensor(lens, dtype=torch.long)


def main():
    torch.set_num_threads(8)
    enc, _ = build_reference_encoder(16, with_text_embedding=True)
    sd = synth.make_state_dict(seed=0, interlingua_length=16, text_vocab=synth.VOCAB)
    print("strict load:", enc.load_state_dict(sd, strict=True))
    tok, lens = make_tokens(LENS, TOK_SEED)
    st = {}
    h = enc.layer_norm.register_forward_hook(lambda m, i, o: st.__setitem__("h_enc", o.detach().transpose(0, 1).clone()))
    with torch.no_grad():
        out = enc(tok, lens)
    h.remove()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "text.npz"), tokens=tok.numpy(), src_lengths=lens.numpy(),
                        memories=out.encoder_out.numpy(), h_enc=st["h_enc"].numpy(),
                        encoder_padding_mask=out.encoder_padding_mask.numpy(), token_seed=TOK_SEED, weight_seed=0,
                        weight_checksum=synth.state_dict_checksum(sd))
    print("text", tuple(out.encoder_out.shape), tuple(st["h_enc"].shape))


if __name__ == "__main__":
    main()
