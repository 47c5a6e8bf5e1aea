"""Deterministic synthetic weights and waveforms for the Chimera speech-encoding path.

There is no network for checkpoints/datasets, so parity tests and `bench.py` use
random-init weights of the reference architecture and synthetic 16 kHz audio.
The state-dict KEY LAYOUT is the reference's (verified by loading
the result with strict=True into the reference encoder in `oracle/gen_golden.py`):
`S2T_W2V2_TransformerInterlinguaEncoder.state_dict()`
(fairseq/models/chimera/w2v2_transformer_interlingua.py:160-188,
 fairseq/models/chimera/w2v2_transformer.py:244-316,
 fairseq/models/wav2vec/wav2vec2.py:283-395,685-853).

Scales follow the reference initialisers (kaiming-normal convs wav2vec2.py:708,
init_bert_params N(0,0.02), pos-conv N(0,sqrt(4/(128*768))) wav2vec2.py:781-783,
xavier MHA multihead_attention.py:98-116) but biases / norm affine parameters are
randomised too (the reference starts them at 0/1), so a kernel that drops a bias or
a LayerNorm gain fails parity instead of passing by accident.
"""
from collections import OrderedDict
import math

import torch

# wav2vec2-base feature extractor: examples/wav2vec/README.md:59-69 (pickled in the ckpt args)
CONV_LAYERS = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)] * 2
W2V_DIM, W2V_FFN, W2V_HEADS, W2V_LAYERS = 768, 3072, 12, 12
POS_K, POS_GROUPS = 128, 16
ENC_DIM, ENC_FFN, ENC_HEADS, ENC_LAYERS = 512, 2048, 8, 6
MEM_LAYERS = 3
SUB_MID = 1024
SAMPLE_RATE = 16000


def encoder_param_spec(interlingua_length=16, dead_heads=True, text_vocab=0, modal_embedding=False, non_shared_encoder_layers=0):
    """[(name, shape, kind, scale)] in the reference's state_dict order."""
    s = []
    W = "wav2vec_model."
    if dead_heads:
        s.append((W + "mask_emb", (W2V_DIM,), "normal", 0.3))
    in_d = 1
    for i, (dim, k, _) in enumerate(CONV_LAYERS):
        s.append((W + f"feature_extractor.conv_layers.{i}.0.weight", (dim, in_d, k), "normal",
                  math.sqrt(2.0 / (in_d * k))))
        if i == 0:
            s.append((W + "feature_extractor.conv_layers.0.2.weight", (dim,), "gain", 0.1))
            s.append((W + "feature_extractor.conv_layers.0.2.bias", (dim,), "normal", 0.1))
        in_d = dim
    s.append((W + "post_extract_proj.weight", (W2V_DIM, 512), "normal", 0.0255))
    s.append((W + "post_extract_proj.bias", (W2V_DIM,), "normal", 0.02))
    if dead_heads:
        s.append((W + "quantizer.vars", (1, 640, 128), "normal", 0.29))
        s.append((W + "quantizer.weight_proj.weight", (640, 512), "normal", 1.0))
        s.append((W + "quantizer.weight_proj.bias", (640,), "zeros", 0))
        s.append((W + "project_q.weight", (256, 256), "normal", 0.036))
        s.append((W + "project_q.bias", (256,), "normal", 0.036))
    s.append((W + "encoder.pos_conv.0.bias", (W2V_DIM,), "normal", 0.02))
    s.append((W + "encoder.pos_conv.0.weight_g", (1, 1, POS_K), "posconv_g", 0.1))
    s.append((W + "encoder.pos_conv.0.weight_v", (W2V_DIM, W2V_DIM // POS_GROUPS, POS_K), "normal",
              math.sqrt(4.0 / (POS_K * W2V_DIM))))

    def layer(prefix, d, ffn, wstd, ostd, f1std, f2std, qk_gain):
        # reference registration order: k, v, q, out (multihead_attention.py:70-85)
        for p, g in (("k", qk_gain), ("v", 1.0), ("q", qk_gain)):
            s.append((prefix + f"self_attn.{p}_proj.weight", (d, d), "normal", wstd * g))
            s.append((prefix + f"self_attn.{p}_proj.bias", (d,), "normal", 0.02))
        s.append((prefix + "self_attn.out_proj.weight", (d, d), "normal", ostd))
        s.append((prefix + "self_attn.out_proj.bias", (d,), "normal", 0.02))
        s.append((prefix + "self_attn_layer_norm.weight", (d,), "gain", 0.1))
        s.append((prefix + "self_attn_layer_norm.bias", (d,), "normal", 0.05))
        s.append((prefix + "fc1.weight", (ffn, d), "normal", f1std))
        s.append((prefix + "fc1.bias", (ffn,), "normal", 0.02))
        s.append((prefix + "fc2.weight", (d, ffn), "normal", f2std))
        s.append((prefix + "fc2.bias", (d,), "normal", 0.02))
        s.append((prefix + "final_layer_norm.weight", (d,), "gain", 0.1))
        s.append((prefix + "final_layer_norm.bias", (d,), "normal", 0.05))

    for i in range(W2V_LAYERS):
        layer(W + f"encoder.layers.{i}.", W2V_DIM, W2V_FFN, 0.02, 0.02, 0.02, 0.02, 3.0)
    s.append((W + "encoder.layer_norm.weight", (W2V_DIM,), "gain", 0.1))
    s.append((W + "encoder.layer_norm.bias", (W2V_DIM,), "normal", 0.05))
    s.append((W + "layer_norm.weight", (512,), "gain", 0.1))
    s.append((W + "layer_norm.bias", (512,), "normal", 0.05))
    if dead_heads:
        s.append((W + "final_proj.weight", (256, W2V_DIM), "normal", 0.02))
        s.append((W + "final_proj.bias", (256,), "normal", 0.02))
    s.append(("subsample.conv_layers.0.weight", (SUB_MID, W2V_DIM, 5), "normal", 0.0093))
    s.append(("subsample.conv_layers.0.bias", (SUB_MID,), "normal", 0.0094))
    s.append(("subsample.conv_layers.1.weight", (2 * ENC_DIM, SUB_MID // 2, 5), "normal", 0.0114))
    s.append(("subsample.conv_layers.1.bias", (2 * ENC_DIM,), "normal", 0.0115))
    s.append(("embed_positions._float_tensor", (1,), "zeros", 0))
    for i in range(ENC_LAYERS):
        layer(f"transformer_layers.{i}.", ENC_DIM, ENC_FFN, 0.03125, 0.0442, 0.0255, 0.01275, 2.0)
    s.append(("layer_norm.weight", (ENC_DIM,), "gain", 0.1))
    s.append(("layer_norm.bias", (ENC_DIM,), "normal", 0.05))
    if text_vocab:
        s.append(("text_embed_tokens.weight", (text_vocab, ENC_DIM), "normal", ENC_DIM ** -0.5))
    s.append(("interlingua_embedding.weight", (interlingua_length, ENC_DIM), "embed0", ENC_DIM ** -0.5))
    for i in range(MEM_LAYERS):
        layer(f"interlingua_layers.{i}.", ENC_DIM, ENC_FFN, 0.03125, 0.0442, 0.0255, 0.01275, 2.0)
    if modal_embedding:                      # 'modal_embedding' in interlingua_debug_options: Embedding(3, 512, padding_idx 2), :179-180
        s.append(("modal_embedding.weight", (3, ENC_DIM), "normal", ENC_DIM ** -0.5))
    for i in range(non_shared_encoder_layers):      # :183-187
        layer(f"audio_exclusive_layers.{i}.", ENC_DIM, ENC_FFN, 0.03125, 0.0442, 0.0255, 0.01275, 2.0)
    return s


def make_state_dict(seed=0, interlingua_length=16, dead_heads=True, text_vocab=0, prefix="", modal_embedding=False,
                    non_shared_encoder_layers=0):
    """Seeded fp32 CPU state dict with the reference encoder's keys/shapes."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    pending_g = None
    for name, shape, kind, scale in encoder_param_spec(interlingua_length, dead_heads, text_vocab, modal_embedding,
                                                       non_shared_encoder_layers):
        if kind == "normal":
            t = torch.randn(shape, generator=g) * scale
        elif kind == "gain":
            t = 1.0 + scale * torch.randn(shape, generator=g)
        elif kind == "zeros":
            t = torch.zeros(shape)
        elif kind == "embed0":        # Embedding(padding_idx=0): row 0 starts at zero (transformer.py:906-910)
            t = torch.randn(shape, generator=g) * scale
            t[0].zero_()
        elif kind == "posconv_g":     # weight_norm(dim=2) init g=||v||_(0,1); perturbed so g != ||v||
            t = 1.0 + scale * torch.randn(shape, generator=g)
            pending_g = prefix + name
        else:
            raise ValueError(kind)
        sd[prefix + name] = t
        if name.endswith("pos_conv.0.weight_v") and pending_g is not None:
            sd[pending_g] = sd[pending_g] * t.norm(dim=(0, 1), keepdim=True)
    return sd


def state_dict_checksum(sd):
    """Order-independent fp64 fingerprint: detects RNG drift between torch builds."""
    tot = 0.0
    for k in sorted(sd):
        v = sd[k].double()
        tot += float(v.sum()) + 0.5 * float((v * v).sum())
    return tot


def make_waveforms(lengths, seed=1234, pad_to=None):
    """x = clamp(0.1*randn, -1, 1), zero tail beyond each length (SURVEY.md §8(d);
    zero padding as `_collate_frames`, fairseq/data/audio/speech_to_text_dataset.py:207-225).
    Returns (wave [B,L] fp32, lengths [B] int64); rows keep the order given."""
    lengths = [int(x) for x in lengths]
    L = int(pad_to) if pad_to else max(lengths)
    g = torch.Generator().manual_seed(seed)
    x = (0.1 * torch.randn(len(lengths), L, generator=g)).clamp_(-1.0, 1.0)
    for b, n in enumerate(lengths):
        x[b, n:] = 0.0
    return x, torch.tensor(lengths, dtype=torch.int64)


# ---- decoder: seeded weights with the reference's keys, for the greedy-decoding path (decoder.py),
# the "identical greedy-decoded token IDs" parity tests (BASELINE.json north_star) and `bench.py --decode`.
DEC_LAYERS, VOCAB = 6, 10000


def decoder_param_spec(vocab=VOCAB):
    s = [("embed_tokens.weight", (vocab, ENC_DIM), "embed_pad1", 1.0), ("embed_positions._float_tensor", (1,), "zeros", 0)]
    for i in range(DEC_LAYERS):
        P = f"layers.{i}."
        for blk, gain in (("self_attn", 1.0), ("encoder_attn", 3.0)):
            for p in ("k", "v", "q"):
                s.append((P + f"{blk}.{p}_proj.weight", (ENC_DIM, ENC_DIM), "normal", 0.03125 * (gain if p != "v" else 1.0)))
                s.append((P + f"{blk}.{p}_proj.bias", (ENC_DIM,), "normal", 0.02))
            s.append((P + f"{blk}.out_proj.weight", (ENC_DIM, ENC_DIM), "normal", 0.0442 * gain))
            s.append((P + f"{blk}.out_proj.bias", (ENC_DIM,), "normal", 0.02))
            s.append((P + f"{blk}_layer_norm.weight", (ENC_DIM,), "gain", 0.1))
            s.append((P + f"{blk}_layer_norm.bias", (ENC_DIM,), "normal", 0.05))
        s.append((P + "fc1.weight", (ENC_FFN, ENC_DIM), "normal", 0.0255))
        s.append((P + "fc1.bias", (ENC_FFN,), "normal", 0.02))
        s.append((P + "fc2.weight", (ENC_DIM, ENC_FFN), "normal", 0.01275))
        s.append((P + "fc2.bias", (ENC_DIM,), "normal", 0.02))
        s.append((P + "final_layer_norm.weight", (ENC_DIM,), "gain", 0.1))
        s.append((P + "final_layer_norm.bias", (ENC_DIM,), "normal", 0.05))
    s.append(("layer_norm.weight", (ENC_DIM,), "gain", 0.1))
    s.append(("layer_norm.bias", (ENC_DIM,), "normal", 0.05))
    return s


def make_decoder_state_dict(seed=1, vocab=VOCAB, embed_std=None, prefix="decoder."):
    """Seeded decoder weights with the reference's keys (tied input/output embeddings: `output_projection.weight`
    is the same tensor as `embed_tokens.weight`; `version` buffer as in fairseq)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    embed_std = embed_std if embed_std is not None else ENC_DIM ** -0.5
    for name, shape, kind, scale in decoder_param_spec(vocab):
        if kind == "normal":
            t = torch.randn(shape, generator=g) * scale
        elif kind == "gain":
            t = 1.0 + scale * torch.randn(shape, generator=g)
        elif kind == "zeros":
            t = torch.zeros(shape)
        elif kind == "embed_pad1":     # Embedding(padding_idx=1): N(0, d^-0.5), pad row zero (transformer.py:906-910)
            t = torch.randn(shape, generator=g) * embed_std
            t[1].zero_()
        else:
            raise ValueError(kind)
        sd[prefix + name] = t
    sd[prefix + "output_projection.weight"] = sd[prefix + "embed_tokens.weight"]
    sd[prefix + "version"] = torch.tensor([3.0])
    return sd
