"""One-time weight preparation: reference state-dict layout -> kernel operand layout (device).

Input keys are the reference encoder's, optionally prefixed "encoder.".  Both
weight-norm forms of the pos-conv are accepted: `pos_conv.0.weight_g/_v` (checkpoint) or the folded
`pos_conv.0.weight` left by make_generation_fast_ (fairseq/models/fairseq_model.py:175-182).

Layout rules (all exact re-arrangements, no arithmetic except the folds noted):
  * conv0 (16-bit mode): 3-term fp16 split of the [512,10] kernel for the tensor-core conv0 (see cst_conv0_apply_tc).
  * conv kernels [Cout, Cin, k] -> [Cout, k*Cin] (tap-major) so a window of channels-last frames is one
    contiguous K run (implicit GEMM).
  * q projection and bias are multiplied by head_dim**-0.5 = 1/8 (exact power of two; the reference
    scales q after the bias, multihead_attention.py -> torch MHA), q|k|v concatenated.
  * subsampler convs: output channels interleaved (value_i, gate_i) so GLU pairs are adjacent columns.
  * pos-conv: w = g * v / ||v||_(0,1) (wav2vec2.py:785), split per group to [16, 48, 128*64] with the 48
    input channels of a tap zero-padded to 64 lanes (one 128-byte swizzle row per tap in bf16).
GEMM operands are stored in `act_dtype` (fp32 or bf16); biases / norm parameters stay fp32.  In the 16-bit mode the
conv feature extractor (conv1..6) may use fp16 operands (`conv_dtype`): that stack has no normalisation between its 7
layers, so operand rounding accumulates (bf16: 6.3e-3 rel-L2 at its output, fp16: 0.8e-3) -- same bytes, same tensor
throughput, and fp16 is the half-precision format of the reference's own recipes (`--fp16`).
"""
import os

import torch

from .synth import CONV_LAYERS, W2V_LAYERS, ENC_LAYERS, MEM_LAYERS, POS_GROUPS, POS_K


def _strip(sd):
    if any(k.startswith("encoder.") for k in sd):
        sd = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
    return sd


def _glu_interleave(w):
    half = w.shape[0] // 2
    return torch.stack((w[:half], w[half:]), dim=1).reshape(w.shape)


def prepare(state_dict, device, act_dtype, conv_dtype=None, training=False):
    """conv_dtype: operand dtype of the conv feature extractor GEMMs (conv1..6); defaults to act_dtype.
    training=True (train.py): only the operands the training step uses, built with device-side ops only (no host round trip), so
    that the refresh after an optimizer update can be captured into a CUDA graph."""
    sd = _strip(state_dict)
    conv_dtype = conv_dtype or act_dtype
    f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()   # noqa: E731
    op = lambda t: t.detach().to(device=device, dtype=torch.float32).to(act_dtype).contiguous()   # noqa: E731
    W = "wav2vec_model."
    P = {}
    fe = W + "feature_extractor.conv_layers."
    P["conv0_w"] = f32(sd[fe + "0.0.weight"].reshape(512, 10))
    if act_dtype != torch.float32 and not training:
        # tcgen05 conv0 (cst_conv0_apply_tc): fp16 [512, 64] rows [hi | hi | lo | 0...], hi = fp16(w), lo = fp16(w - hi)
        w0 = P["conv0_w"]
        hi = w0.to(torch.float16)
        lo = (w0 - hi.float()).to(torch.float16)
        w16 = torch.zeros(512, 64, dtype=torch.float16, device=device)
        w16[:, 0:10], w16[:, 10:20], w16[:, 20:30] = hi, hi, lo
        P["conv0_w16"] = w16.contiguous()
    P["gn_g"], P["gn_b"] = f32(sd[fe + "0.2.weight"]), f32(sd[fe + "0.2.bias"])
    for i in range(1, len(CONV_LAYERS)):
        w = sd[fe + f"{i}.0.weight"]                       # [512, 512, k]
        P[f"conv{i}_w"] = w.detach().to(device=device, dtype=torch.float32).permute(0, 2, 1).reshape(w.shape[0], -1).to(conv_dtype).contiguous()
    P["ln_feat_g"], P["ln_feat_b"] = f32(sd[W + "layer_norm.weight"]), f32(sd[W + "layer_norm.bias"])
    P["proj_w"], P["proj_b"] = op(sd[W + "post_extract_proj.weight"]), f32(sd[W + "post_extract_proj.bias"])

    pc = W + "encoder.pos_conv.0."
    if pc + "weight" in sd:
        w = sd[pc + "weight"].float()
    else:
        v, g = sd[pc + "weight_v"].float(), sd[pc + "weight_g"].float()
        w = v * (g / v.norm(dim=(0, 1), keepdim=True))
    cg = w.shape[1]                                        # 48 channels per group
    wg = w.view(POS_GROUPS, cg, cg, POS_K).permute(0, 1, 3, 2)      # [g, co, tap, ci]
    wp = torch.zeros(POS_GROUPS, cg, POS_K, 64, dtype=torch.float32, device=w.device)
    wp[..., :cg] = wg
    P["pos_w"] = op(wp.reshape(POS_GROUPS, cg, POS_K * 64))
    if act_dtype != torch.float32 and not training:
        # cst_posconv_stacked: [group, tap pair, 128 rows, 64 lanes]; rows 0..47 even tap, rows 64..111 odd tap
        w2 = torch.zeros(POS_GROUPS, POS_K // 2, 128, 64, dtype=torch.float32, device=w.device)
        w2[:, :, 0:cg, :] = wp[:, :, 0::2, :].permute(0, 2, 1, 3)
        w2[:, :, 64:64 + cg, :] = wp[:, :, 1::2, :].permute(0, 2, 1, 3)
        P["pos_w2"] = op(w2)
    P["pos_b"] = f32(sd[pc + "bias"])
    P["ln_enc_g"], P["ln_enc_b"] = f32(sd[W + "encoder.layer_norm.weight"]), f32(sd[W + "encoder.layer_norm.bias"])

    def layer(prefix, fused_qkv=True):
        a = prefix + "self_attn."
        d = {}
        wq, bq = sd[a + "q_proj.weight"].float() * 0.125, sd[a + "q_proj.bias"].float() * 0.125
        wk, bk = sd[a + "k_proj.weight"].float(), sd[a + "k_proj.bias"].float()
        wv, bv = sd[a + "v_proj.weight"].float(), sd[a + "v_proj.bias"].float()
        if fused_qkv:
            d["qkv_w"], d["qkv_b"] = op(torch.cat((wq, wk, wv), 0)), f32(torch.cat((bq, bk, bv), 0))
            d["_qkv_raw"] = (torch.cat((wq, wk, wv), 0), torch.cat((bq, bk, bv), 0))       # fp32 masters for fold_ln, dropped below
            d["_fc1_raw"] = (sd[prefix + "fc1.weight"].float(), sd[prefix + "fc1.bias"].float())
        else:
            d["q_w"], d["q_b"] = op(wq), f32(bq)
            d["kv_w"], d["kv_b"] = op(torch.cat((wk, wv), 0)), f32(torch.cat((bk, bv), 0))
        d["o_w"], d["o_b"] = op(sd[a + "out_proj.weight"]), f32(sd[a + "out_proj.bias"])
        d["ln1_g"], d["ln1_b"] = f32(sd[prefix + "self_attn_layer_norm.weight"]), f32(sd[prefix + "self_attn_layer_norm.bias"])
        d["fc1_w"], d["fc1_b"] = op(sd[prefix + "fc1.weight"]), f32(sd[prefix + "fc1.bias"])
        d["fc2_w"], d["fc2_b"] = op(sd[prefix + "fc2.weight"]), f32(sd[prefix + "fc2.bias"])
        d["ln2_g"], d["ln2_b"] = f32(sd[prefix + "final_layer_norm.weight"]), f32(sd[prefix + "final_layer_norm.bias"])
        return d

    def fold_ln(d, name, gamma, beta):
        """LayerNorm folded around a projection for the fused-LayerNorm GEMM epilogue (cst_gemm_params.ln_in_stats):
        LN(y; g, b) W^T + c = rstd * (y (W diag g)^T) - rstd*mean * colsum(W diag g) + (W b + c).  The GEMM multiplies
        the bf16 copy of the UN-normalised rows by W' = W diag g; colsum is taken over the ROUNDED W' (what the tensor
        cores really multiply), so the mean term cancels exactly what the product contains."""
        w, c = (t.double().cpu() for t in d["_" + name + "_raw"])
        g64, b64 = gamma.double().cpu(), beta.double().cpu()
        wf = (w * g64[None, :]).float().to(act_dtype)
        d[name + "L_w"] = wf.to(device).contiguous()
        d[name + "L_cs"] = wf.double().sum(1).float().to(device).contiguous()
        d[name + "L_b"] = (c + w @ b64).float().to(device).contiguous()

    P["w2v_layers"] = [layer(W + f"encoder.layers.{i}.") for i in range(W2V_LAYERS)]
    for i in range(2):
        w = sd[f"subsample.conv_layers.{i}.weight"].float()        # [1024, Cin, 5]
        w = w.permute(0, 2, 1).reshape(w.shape[0], -1)
        P[f"sub{i}_w"] = op(_glu_interleave(w))
        P[f"sub{i}_b"] = f32(_glu_interleave(sd[f"subsample.conv_layers.{i}.bias"].float()))
    P["enc_layers"] = [layer(f"transformer_layers.{i}.") for i in range(ENC_LAYERS)]
    # non_shared_encoder_layers = n (w2v2_transformer_interlingua.py:183-187,239-249): the AUDIO branch runs audio_exclusive_layers[0..n)
    # in place of transformer_layers[0..n); the text branch keeps all shared layers
    n_excl = 0
    while f"audio_exclusive_layers.{n_excl}.fc1.weight" in sd:
        n_excl += 1
    if n_excl:
        P["enc_layers_text"] = P["enc_layers"]
        P["enc_layers"] = [layer(f"audio_exclusive_layers.{i}.") for i in range(n_excl)] + P["enc_layers_text"][n_excl:]
    if act_dtype != torch.float32 and not training:
        # fused-LayerNorm variants (16-bit mode): post-LN wav2vec2 layers -- QKV of layer i consumes LN2 of layer i-1, fc1
        # consumes LN1 of its own layer; pre-LN shared layers -- QKV consumes LN1, fc1 consumes LN2 of their own layer
        for i, d in enumerate(P["w2v_layers"]):
            if i > 0:
                fold_ln(d, "qkv", P["w2v_layers"][i - 1]["ln2_g"], P["w2v_layers"][i - 1]["ln2_b"])
            fold_ln(d, "fc1", d["ln1_g"], d["ln1_b"])
        for i, d in enumerate(P["enc_layers"]):
            if i > 0:
                fold_ln(d, "qkv", d["ln1_g"], d["ln1_b"])
            fold_ln(d, "fc1", d["ln2_g"], d["ln2_b"])
    for d in P["w2v_layers"] + P["enc_layers"] + P.get("enc_layers_text", []):
        d.pop("_qkv_raw", None)
        d.pop("_fc1_raw", None)
    P["ln_out_g"], P["ln_out_b"] = f32(sd["layer_norm.weight"]), f32(sd["layer_norm.bias"])
    P["mem_embed"] = f32(sd["interlingua_embedding.weight"])
    if "modal_embedding.weight" in sd:
        # 'modal_embedding' debug option (w2v2_transformer_interlingua.py:179-182,272-282): row 0 (audio) / row 1 (text) of a 3-row
        # embedding is added to every memory query before the memory stage
        me = sd["modal_embedding.weight"].float()
        P["mem_embed_text"] = f32(sd["interlingua_embedding.weight"].float() + me[1])
        P["mem_embed"] = f32(sd["interlingua_embedding.weight"].float() + me[0])
    if "text_embed_tokens.weight" in sd:                       # text (MT) branch input, interlingua:216
        P["text_embed"] = f32(sd["text_embed_tokens.weight"])
    P["mem_layers"] = [layer(f"interlingua_layers.{i}.", fused_qkv=False) for i in range(MEM_LAYERS)]
    # K/V side of the memory stage: every layer i normalises the SAME h_enc with its own affine LN1 and projects it
    # (w2v2_transformer_interlingua.py:262-274 -> transformer_layer.py:129-139).  LN(x; g, b) W^T + c =
    # xhat (W diag g)^T + (W b + c), so the affine part folds into the projection exactly: one unit LayerNorm of h_enc
    # and ONE [MEM_LAYERS*1024, 512] GEMM replace MEM_LAYERS LayerNorms and MEM_LAYERS small GEMMs.
    ws, bs = [], []
    for i in range(MEM_LAYERS):
        pre = f"interlingua_layers.{i}."
        g_, b_ = sd[pre + "self_attn_layer_norm.weight"].double(), sd[pre + "self_attn_layer_norm.bias"].double()
        a = pre + "self_attn."
        w = torch.cat((sd[a + "k_proj.weight"], sd[a + "v_proj.weight"]), 0).double()
        c = torch.cat((sd[a + "k_proj.bias"], sd[a + "v_proj.bias"]), 0).double()
        ws.append(w * g_[None, :])
        bs.append(c + w @ b_)
    P["mem_kv_w"], P["mem_kv_b"] = op(torch.cat(ws, 0).float()), f32(torch.cat(bs, 0).float())
    P["unit_g"] = torch.ones(sd["layer_norm.weight"].shape[0], dtype=torch.float32, device=device)
    P["unit_b"] = torch.zeros(sd["layer_norm.weight"].shape[0], dtype=torch.float32, device=device)
    if act_dtype == torch.float32 and os.environ.get("CST_F32_TC", "0") == "1":
        P["_s16"] = split_packs(P)          # only the opt-in "fast fp32" mode needs them (plan.py)
    return P


def split_w(w, blk):
    """fp32 [N, K] (K = n_blk * blk) -> (fp16 [N, 3K], 2^-s): per blk-wide block [hi | hi | lo], hi = fp16(w * 2^s),
    lo = fp16(w * 2^s - hi) -- the weight side of the 3-term split (cst_split_f16 packs the activation rows as [hi | lo | hi]).
    The power-of-two scale 2^s (exact) lifts the largest |w| to ~2^14: trained weights are ~1e-2, whose lo parts (~1e-5) would
    otherwise fall into fp16's denormals and lose their mantissa; the GEMM epilogue multiplies the accumulator by 2^-s."""
    import math
    N, K = w.shape[-2], w.shape[-1]
    amax = float(w.abs().max())
    sh = max(0, min(24, int(math.floor(math.log2(16384.0 / amax))))) if amax > 0 else 0
    w3 = (w.reshape(-1, N, K // blk, blk).float() * float(2.0 ** sh))
    hi = w3.to(torch.float16)
    lo = (w3 - hi.float()).to(torch.float16)
    return torch.cat((hi, hi, lo), dim=-1).reshape(*w.shape[:-1], 3 * K).contiguous(), float(2.0 ** -sh)


def split_packs(P):
    """fp32 mode on the tensor cores: {data_ptr of the fp32 GEMM weight: (fp16 split pack, block width)}.  The block width is
    the width of one SOURCE row of the A operand: K for a plain linear, the channel count for the implicit-GEMM convolutions
    (a window spans several source rows), 64 for the packed pos-conv operand."""
    out = {}

    def add(w, blk):
        w16, inv = split_w(w, blk)
        out[w.data_ptr()] = (w16, blk, inv)
    for i in range(1, len(CONV_LAYERS)):
        add(P[f"conv{i}_w"], 512)
    add(P["proj_w"], 512)
    add(P["pos_w"], 64)
    for d in P["w2v_layers"] + P["enc_layers"]:
        for k in ("qkv_w", "o_w", "fc1_w", "fc2_w"):
            add(d[k], d[k].shape[1])
    add(P["sub0_w"], 768)
    add(P["sub1_w"], 512)
    add(P["mem_kv_w"], 512)
    for d in P["mem_layers"]:
        for k in ("q_w", "o_w", "fc1_w", "fc2_w"):
            add(d[k], d[k].shape[1])
    return out
