"""TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference Chimera encoder
(S2T_W2V2_TransformerInterlinguaEncoder, fairseq/models/chimera/
w2v2_transformer_interlingua.py:155-312) on CPU through the overlay of
`make_overlay.py`.  Only usable in the dev container (needs /root/reference).
"""
import argparse
import os
import tempfile

import torch

from . import make_overlay

W2V_CONV_SPEC = "[(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512,2,2)] * 2"   # examples/wav2vec/README.md:59-69


def build_reference_encoder(interlingua_length=16, with_text_embedding=False):
    """Return the reference encoder module (eval mode, fp32, random init)."""
    make_overlay.build()
    make_overlay.activate()
    import fairseq.models  # noqa: F401  registers all archs
    from fairseq.models.wav2vec import wav2vec2 as W
    from fairseq.models.chimera.w2v2_transformer_interlingua import (
        S2TTransformerInterlinguaModelW2V2, s2t_transformer_w2v2_interlingua_base)

    w2v_args = argparse.Namespace(
        conv_feature_layers=W2V_CONV_SPEC, quantize_targets=True, final_dim=256,
        encoder_layerdrop=0.05, dropout_input=0.1, dropout_features=0.1,
        feature_grad_mult=0.1)
    W.base_architecture(w2v_args)
    torch.manual_seed(0)
    w2v = W.Wav2Vec2Model.build_model(w2v_args, task=None)
    tmp = tempfile.NamedTemporaryFile(suffix=".pt", delete=False)
    tmp.close()
    torch.save({"args": w2v_args, "model": w2v.state_dict()}, tmp.name)

    args = argparse.Namespace(
        w2v2_model_path=tmp.name, encoder_layers=6, encoder_embed_dim=512,
        interlingua_length=interlingua_length, interlingua_layers=3,
        interlingua_debug_options=[], dropout=0.1,
        share_decoder_input_output_embed=True,
        max_source_positions=6000, max_target_positions=1024)
    s2t_transformer_w2v2_interlingua_base(args)
    embed = None
    if with_text_embedding:
        from fairseq.models.transformer import Embedding
        embed = Embedding(10000, 512, 1)
    try:
        enc = S2TTransformerInterlinguaModelW2V2.build_encoder(args, None, embed)
    finally:
        os.unlink(tmp.name)
    enc.eval()
    return enc, args
