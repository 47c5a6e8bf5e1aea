"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (plain torch tensor arithmetic, fp32 or fp64) of the reference's
speech-encoding hot path: raw 16 kHz waveform -> M shared semantic memories.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import this module; the product path
(`chimera-st_b200/`) never does and fails loudly when its CUDA library is missing.

Pinning: the reference ships no tests or golden vectors for this path (SURVEY.md
§4, §8(c)); the oracle is pinned against outputs of the UNMODIFIED reference run in
the dev container through `oracle/make_overlay.py` -- `oracle/gen_golden.py` wrote
`tests/golden/*.npz`, `tests/test_oracle_golden.py` checks this file against them.

Every function cites the reference lines it follows (paths under /root/reference).
Attention is written out explicitly (the reference delegates to torch's
F.multi_head_attention_forward, fairseq/modules/multihead_attention.py:155-187).
"""
import math

import torch
import torch.nn.functional as F

CONV_LAYERS = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)] * 2
EPS = 1e-5


# ----------------------------------------------------------------------------- integer side
def conv_out_lengths(L):
    """Conv1d without padding: T_i = (T_{i-1} - k)//s + 1 (wav2vec2.py:707, torch Conv1d)."""
    out = []
    for _, k, s in CONV_LAYERS:
        L = (L - k) // s + 1
        out.append(L)
    return out


def lengths_to_padding_mask(lens):
    """fairseq/data/data_utils.py:491-495."""
    bsz, max_len = lens.size(0), int(lens.max())
    mask = torch.arange(max_len, device=lens.device).view(1, max_len).expand(bsz, -1)      # `.to(lengths.device)` in the reference
    return mask >= lens.view(bsz, 1)


def frame_padding_mask(src_lengths, n_frames):
    """Sample-level mask -> frame-level mask by the reference's trim/view/all rule
    (wav2vec2.py:543-548, via w2v2_transformer.py:327).  Returns bool [B, n_frames]."""
    mask = lengths_to_padding_mask(src_lengths)          # [B, L], L = max(len)
    extra = mask.size(1) % n_frames
    if extra > 0:
        mask = mask[:, :-extra]
    return mask.view(mask.size(0), n_frames, -1).all(-1)


def subsampler_lengths(lens, n_layers=2):
    """Conv1dSubsampler.get_out_seq_lens_tensor, s2t_transformer.py:63-67 (via float)."""
    out = lens.clone()
    for _ in range(n_layers):
        out = ((out.float() - 1) / 2 + 1).floor().long()
    return out


# ----------------------------------------------------------------------------- float helpers
def _ln(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], EPS)


def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def mha(sd, prefix, q_in, kv_in, n_heads, key_padding_mask=None, attn_bias=None, drop_tag=None):
    """softmax(((xWq+bq)*d^-0.5)(xWk+bk)^T + mask)(xWv+bv) Wo + bo, batch-first.
    q_in [B,Tq,C], kv_in [B,Tk,C]; key_padding_mask bool [B,Tk] (True = pad -> -inf);
    attn_bias additive [Tq,Tk].  Follows torch multi_head_attention_forward as called
    from multihead_attention.py:165-187 (scaling after bias: SURVEY App. B.8)."""
    B, Tq, C = q_in.shape
    Tk = kv_in.shape[1]
    d = C // n_heads
    q = F.linear(q_in, sd[prefix + "q_proj.weight"], sd[prefix + "q_proj.bias"]) * (d ** -0.5)
    k = F.linear(kv_in, sd[prefix + "k_proj.weight"], sd[prefix + "k_proj.bias"])
    v = F.linear(kv_in, sd[prefix + "v_proj.weight"], sd[prefix + "v_proj.bias"])
    q = q.view(B, Tq, n_heads, d).transpose(1, 2)
    k = k.view(B, Tk, n_heads, d).transpose(1, 2)
    v = v.view(B, Tk, n_heads, d).transpose(1, 2)
    s = q @ k.transpose(-1, -2)                           # [B,H,Tq,Tk]
    if attn_bias is not None:
        s = s + attn_bias.to(s.dtype)
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    if drop_tag is not None and attn_bias is None:                  # attention_dropout: F.dropout on the probabilities [B,H,Tq,Tk]
        p = _dropout(drop_tag, p)
    o = (p @ v).transpose(1, 2).reshape(B, Tq, C)
    return F.linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])


# ----------------------------------------------------------------------------- stages
def conv_feature_extractor(sd, wave, upto=7):
    """ConvFeatureExtractionModel.forward, wav2vec2.py:755-763; block 0 = conv(k10,s5,no bias)
    -> Fp32GroupNorm(512 groups of 1 channel, over the WHOLE padded time axis) -> GELU
    (wav2vec2.py:697-734, fp32_group_norm.py:17-25); blocks 1-6 = conv(s2,no bias) -> GELU.
    wave [B,L] -> [B,512,T']."""
    P = "wav2vec_model.feature_extractor.conv_layers."
    x = wave.unsqueeze(1)
    for i, (_, k, s) in enumerate(CONV_LAYERS[:upto]):
        x = F.conv1d(x, sd[P + f"{i}.0.weight"], None, stride=s)
        if i == 0:
            mu = x.mean(dim=2, keepdim=True)
            var = x.var(dim=2, unbiased=False, keepdim=True)
            x = (x - mu) / torch.sqrt(var + EPS)
            x = x * sd[P + "0.2.weight"].view(1, -1, 1) + sd[P + "0.2.bias"].view(1, -1, 1)
        x = _gelu(x)
    return x


def pos_conv_weight(sd):
    """weight_norm(dim=2): w = g * v / ||v||_(0,1)  (wav2vec2.py:785); accepts the folded form
    `pos_conv.0.weight` left by make_generation_fast_ (fairseq_model.py:175-182)."""
    P = "wav2vec_model.encoder.pos_conv.0."
    if P + "weight" in sd:
        return sd[P + "weight"]
    v, g = sd[P + "weight_v"], sd[P + "weight_g"]
    return v * (g / v.norm(dim=(0, 1), keepdim=True))


DROPOUT_HOOK = None


def _dropout(tag, x):
    """A dropout site of the training-mode forward (FairseqDropout / F.dropout, fairseq/modules/fairseq_dropout.py:16-27).  The oracle is
    the eval-mode restatement: the identity unless a test installs DROPOUT_HOOK(tag, x) -> x * mask / (1 - p) to replay the masks the
    CUDA path drew (tests/test_train_emulated.py, tests/test_gpu_backward.py)."""
    return x if DROPOUT_HOOK is None else DROPOUT_HOOK(tag, x)


def w2v_layer(sd, i, x, pad_mask):
    """TransformerSentenceEncoderLayer.forward post-LN branch, wav2vec2.py:937-957. x [B,T,768].
    Dropout sites: dropout1 after the attention, dropout3 after fc2 (dropout2 = activation_dropout is 0 in wav2vec2 base)."""
    P = f"wav2vec_model.encoder.layers.{i}."
    a = _dropout(f"w2v{i}.attn", mha(sd, P + "self_attn.", x, x, 12, key_padding_mask=pad_mask, drop_tag=f"w2v{i}.prob"))
    x = _ln(x + a, sd, P + "self_attn_layer_norm")
    h = _gelu(F.linear(x, sd[P + "fc1.weight"], sd[P + "fc1.bias"]))
    h = _dropout(f"w2v{i}.ffn", F.linear(h, sd[P + "fc2.weight"], sd[P + "fc2.bias"]))
    return _ln(x + h, sd, P + "final_layer_norm")


def wav2vec_extract_features(sd, wave, src_lengths, stages=None):
    """Wav2Vec2Model.extract_features(source, padding_mask, mask=False), wav2vec2.py:527-586,650-652
    + TransformerEncoder.extract_features wav2vec2.py:818-845, as driven by
    _get_w2v_feature (w2v2_transformer.py:319-336).
    Returns (x [B,T',768], frame_mask bool [B,T'], out_len int64 [B])."""
    W = "wav2vec_model."
    feats = conv_feature_extractor(sd, wave)                       # [B,512,T']
    if stages is not None:
        stages["conv_feats"] = feats
    x = _ln(feats.transpose(1, 2), sd, W + "layer_norm")            # :539-540
    fmask = frame_padding_mask(src_lengths, x.size(1))             # :543-548
    x = F.linear(x, sd[W + "post_extract_proj.weight"], sd[W + "post_extract_proj.bias"])  # :550-551
    x = _dropout("w2v.input", x)                                    # dropout_input :553
    x = x.masked_fill(fmask.unsqueeze(-1), 0.0)                     # :820-821
    if stages is not None:
        stages["proj_masked"] = x
    w = pos_conv_weight(sd)
    pc = F.conv1d(x.transpose(1, 2), w, sd[W + "encoder.pos_conv.0.bias"],
                  padding=128 // 2, groups=16)[:, :, :-1]           # SamePad same_pad.py:10-18
    x = x + _gelu(pc).transpose(1, 2)                               # :823-825
    x = _ln(x, sd, W + "encoder.layer_norm")                        # :827-828 (layer_norm_first=False)
    x = _dropout("w2v.enc", x)                                      # F.dropout :830
    if stages is not None:
        stages["w2v_in"] = x
    for i in range(12):
        x = w2v_layer(sd, i, x, fmask)                              # :835-840
        if stages is not None and i == 0:
            stages["w2v_l0"] = x
    out_len = (~fmask).sum(dim=1)                                   # w2v2_transformer.py:333
    if stages is not None:
        stages["w2v_out"] = x
    return x, fmask, out_len


def conv1d_subsampler(sd, x, lens):
    """Conv1dSubsampler.forward, s2t_transformer.py:69-77. x [B,T',768] -> [B,T2,512] (batch-first here)."""
    y = x.transpose(1, 2)
    for i in range(2):
        y = F.conv1d(y, sd[f"subsample.conv_layers.{i}.weight"], sd[f"subsample.conv_layers.{i}.bias"],
                     stride=2, padding=2)
        a, b = y.chunk(2, dim=1)                                    # F.glu(dim=1)
        y = a * torch.sigmoid(b)
    return y.transpose(1, 2), subsampler_lengths(lens)


def _layer_tag(P):
    kind, idx = P.rstrip(".").rsplit(".", 1)
    return {"transformer_layers": "enc", "interlingua_layers": "mem", "audio_exclusive_layers": "aenc"}[kind] + idx


def encoder_layer(sd, P, x, pad_mask, attn_bias=None):
    """TransformerEncoderLayer.forward pre-LN (normalize_before=True), ReLU,
    fairseq/modules/transformer_layer.py:105-155. x [B,T,512]."""
    tag = _layer_tag(P)
    h = _ln(x, sd, P + "self_attn_layer_norm")
    x = x + _dropout(tag + ".attn", mha(sd, P + "self_attn.", h, h, 8, key_padding_mask=pad_mask, attn_bias=attn_bias,
                                        drop_tag=tag + ".prob"))                                      # dropout_module
    h = _ln(x, sd, P + "final_layer_norm")
    h = _dropout(tag + ".act", torch.relu(F.linear(h, sd[P + "fc1.weight"], sd[P + "fc1.bias"])))     # activation_dropout_module
    return x + _dropout(tag + ".ffn", F.linear(h, sd[P + "fc2.weight"], sd[P + "fc2.bias"]))          # dropout_module


def memory_stage_literal(sd, h_enc):
    """The reference's literal form (w2v2_transformer_interlingua.py:264-298): a full encoder layer
    over cat(h_enc, memory) with an additive mask hiding the memory columns (1 -> -1e8,
    transformer_layer.py:126-127) and an all-False key-padding mask; keep the last M rows."""
    B, T, C = h_enc.shape
    mem = sd["interlingua_embedding.weight"].unsqueeze(0).repeat(B, 1, 1)
    M = mem.shape[1]
    bias = torch.zeros(T + M, T + M, dtype=h_enc.dtype)
    bias[:, T:] = -1e8
    for l in range(3):
        y = encoder_layer(sd, f"interlingua_layers.{l}.", torch.cat((h_enc, mem), 1), None, attn_bias=bias)
        mem = y[:, -M:]
    return mem


def memory_stage(sd, h_enc):
    """Algebraically identical M-query cross-attention form: per layer
    y = m + MHA(LN1(m) -> q, LN1(h_enc) -> k,v over ALL T2 frames, no key-padding mask);
    m = y + FFN(LN2(y)).  Checked against `memory_stage_literal` in tests."""
    B = h_enc.shape[0]
    mem = sd["interlingua_embedding.weight"].unsqueeze(0).repeat(B, 1, 1)
    for l in range(3):
        P = f"interlingua_layers.{l}."
        q_in = _ln(mem, sd, P + "self_attn_layer_norm")
        kv_in = _ln(h_enc, sd, P + "self_attn_layer_norm")
        y = mem + _dropout(f"mem{l}.attn", mha(sd, P + "self_attn.", q_in, kv_in, 8, drop_tag=f"mem{l}.prob"))
        h = _ln(y, sd, P + "final_layer_norm")
        h = _dropout(f"mem{l}.act", torch.relu(F.linear(h, sd[P + "fc1.weight"], sd[P + "fc1.bias"])))
        mem = y + _dropout(f"mem{l}.ffn", F.linear(h, sd[P + "fc2.weight"], sd[P + "fc2.bias"]))
    return mem


def encoder_forward(sd, wave, src_lengths, stages=None, literal_memory=False):
    """S2T_W2V2_TransformerInterlinguaEncoder.forward (audio branch),
    w2v2_transformer_interlingua.py:207-312.  wave [B,L] float, src_lengths [B] int64,
    max(src_lengths) == L.  Returns (encoder_out [M,B,512], encoder_padding_mask zeros [B,M] bool)."""
    assert int(src_lengths.max()) == wave.shape[1], "collater guarantees max(len)==L"
    x, fmask, lens = wav2vec_extract_features(sd, wave, src_lengths, stages)      # :226-227
    x, lens2 = conv1d_subsampler(sd, x, lens)                                     # :228
    x = math.sqrt(512) * x                                                        # :231
    pad = lengths_to_padding_mask(lens2)                                          # :232
    if pad.shape[1] < x.shape[1]:      # mask width is max(lens2) == T2 whenever max(len)==L
        pad = F.pad(pad, (0, x.shape[1] - pad.shape[1]), value=True)
    if stages is not None:
        stages["sub_out"] = x
        stages["frame_mask"], stages["w2v_len"], stages["sub_len"] = fmask, lens, lens2
    x = _dropout("embed", x)                                                      # dropout_module :237
    for i in range(6):
        x = encoder_layer(sd, f"transformer_layers.{i}.", x, pad)                 # :240-242
    h_enc = _ln(x, sd, "layer_norm")                                              # :254-255
    if stages is not None:
        stages["h_enc"] = h_enc
    mem = memory_stage_literal(sd, h_enc) if literal_memory else memory_stage(sd, h_enc)
    out = mem.transpose(0, 1).contiguous()                                        # [M,B,512]
    return out, torch.zeros(wave.shape[0], out.shape[0], dtype=torch.bool)        # :301-312


def base_encoder_forward(sd, wave, src_lengths):
    """S2T_W2V2_TransformerEncoder.forward (the NON-memory base encoder, fairseq/models/chimera/w2v2_transformer.py:338-386):
    wav2vec2 features -> subsampler -> x = sqrt(512) x + embed_positions(padding mask) (:353-357; positions 2, 3, ... on the valid
    frames, the zero padding row elsewhere, as in the text branch) -> 6 shared layers -> LayerNorm.
    Returns (encoder_out [T2,B,512], encoder_padding_mask bool [B,T2] or None when nothing is padded (:366-367))."""
    assert int(src_lengths.max()) == wave.shape[1], "collater guarantees max(len)==L"
    x, fmask, lens = wav2vec_extract_features(sd, wave, src_lengths)
    x, lens2 = conv1d_subsampler(sd, x, lens)
    x = math.sqrt(512) * x
    pad = lengths_to_padding_mask(lens2)
    if pad.shape[1] < x.shape[1]:
        pad = F.pad(pad, (0, x.shape[1] - pad.shape[1]), value=True)
    valid = ~pad
    positions = torch.cumsum(valid.long(), dim=1) * valid.long() + 1
    x = x + sinusoidal_table(x.shape[1] + 2)[positions].to(x.dtype)
    for i in range(6):
        x = encoder_layer(sd, f"transformer_layers.{i}.", x, pad)
    x = _ln(x, sd, "layer_norm")
    return x.transpose(0, 1).contiguous(), (pad if bool(pad.any()) else None)


def sinusoidal_table(n, dim=512, padding_idx=1):
    """SinusoidalPositionalEmbedding.get_embedding, fairseq/modules/sinusoidal_positional_embedding.py:38-59."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float) * -e)
    e = torch.arange(n, dtype=torch.float).unsqueeze(1) * e.unsqueeze(0)
    t = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
    t[padding_idx] = 0
    return t


def encoder_forward_text(sd, src_tokens, src_lengths, stages=None):
    """S2T_W2V2_TransformerInterlinguaEncoder.forward, TEXT branch (w2v2_transformer_interlingua.py:212-217,230-236):
    x = sqrt(512) * text_embed_tokens(tokens) + embed_positions(padding_mask).  The reference hands the *padding mask*
    (bool) to SinusoidalPositionalEmbedding, whose make_positions (fairseq/utils.py:235-245) then computes
    cumsum(mask != padding_idx) * (mask != padding_idx) + padding_idx with padding_idx = 1, i.e. positions 2, 3, ... on
    the valid tokens and the zeroed padding row elsewhere.  src_tokens [B,T] int64 (max(src_lengths) == T)."""
    B, T = src_tokens.shape
    assert int(src_lengths.max()) == T
    pad = lengths_to_padding_mask(src_lengths)
    valid = ~pad
    positions = torch.cumsum(valid.long(), dim=1) * valid.long() + 1
    x = math.sqrt(512) * sd["text_embed_tokens.weight"][src_tokens] + sinusoidal_table(T + 2)[positions]
    if stages is not None:
        stages["text_in"] = x
    x = _dropout("embed", x)                                                      # dropout_module :237
    for i in range(6):
        x = encoder_layer(sd, f"transformer_layers.{i}.", x, pad)
    h_enc = _ln(x, sd, "layer_norm")
    if stages is not None:
        stages["h_enc"] = h_enc
    out = memory_stage(sd, h_enc).transpose(0, 1).contiguous()
    return out, torch.zeros(B, out.shape[0], dtype=torch.bool)


def cast_state_dict(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
