"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's decoder forward and of greedy search (beam 1), used only to check the
north-star bar "identical greedy-decoded token IDs on a fixed synthetic set": (a) downstream of the encoder (the
same oracle decoder is run on the reference/oracle memories and on the B200 memories) and (b) as the checker of the
B200 greedy decoder (chimera-st_b200/decoder.py).  Pinned against the reference's own `SequenceGenerator`
(tests/golden/greedy.npz, written by oracle/gen_golden.py).

Reference lines restated:
  TransformerDecoder.extract_features_scriptable  fairseq/models/transformer.py:720-828 (+ output_layer :830-838)
  TransformerDecoderLayer.forward (pre-LN)         fairseq/modules/transformer_layer.py:300-412
  SinusoidalPositionalEmbedding                     fairseq/modules/sinusoidal_positional_embedding.py:38-93
  SequenceGenerator._generate, beam_size=1          fairseq/sequence_generator.py:179-540 (pad never selected,
      EOS forbidden before min_len=1, only EOS at step >= max_len, hypothesis ends when EOS is the top candidate)
"""
import math

import torch
import torch.nn.functional as F

from .chimera_oracle import mha, _ln

PAD, EOS, UNK = 1, 2, 3


def sinusoidal_table(n, dim=512, padding_idx=PAD):
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float) * -e)
    e = torch.arange(n, dtype=torch.float).unsqueeze(1) * e.unsqueeze(0)
    t = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
    t[padding_idx] = 0
    return t


def decoder_logits(sd, prev_tokens, memories):
    """prev_tokens [B,T] (no padding), memories [M,B,512] -> logits of the LAST position [B,V]."""
    B, T = prev_tokens.shape
    E = sd["decoder.embed_tokens.weight"]
    pos = sinusoidal_table(PAD + 1 + T)[PAD + 1:PAD + 1 + T]                  # positions start at padding_idx + 1
    x = math.sqrt(512) * E[prev_tokens] + pos.unsqueeze(0)
    mem = memories.transpose(0, 1)                                            # [B,M,512]
    causal = torch.full((T, T), float("-inf")).triu(1)
    for i in range(6):
        P = f"decoder.layers.{i}."
        h = _ln(x, sd, P + "self_attn_layer_norm")
        x = x + mha(sd, P + "self_attn.", h, h, 8, attn_bias=causal)
        h = _ln(x, sd, P + "encoder_attn_layer_norm")
        x = x + mha(sd, P + "encoder_attn.", h, mem, 8)                       # all-False key padding mask
        h = _ln(x, sd, P + "final_layer_norm")
        h = torch.relu(F.linear(h, sd[P + "fc1.weight"], sd[P + "fc1.bias"]))
        x = x + F.linear(h, sd[P + "fc2.weight"], sd[P + "fc2.bias"])
    x = _ln(x[:, -1], sd, "decoder.layer_norm")
    return F.linear(x, sd["decoder.output_projection.weight"])


def greedy_decode(sd, memories, max_len=200, min_len=1, return_margins=False):
    """-> list (per utterance) of token id lists ending with EOS; optionally the smallest top-1/top-2 log-prob
    margin met along each hypothesis (how close any argmax was to flipping)."""
    B = memories.shape[1]
    tokens = torch.full((B, 1), EOS, dtype=torch.long)
    done = [False] * B
    out = [[] for _ in range(B)]
    margins = [float("inf")] * B
    with torch.no_grad():
        for step in range(max_len + 1):
            lp = torch.log_softmax(decoder_logits(sd, tokens, memories).float(), dim=-1)
            lp[:, PAD] = -math.inf
            if step >= max_len:
                lp[:, :EOS] = -math.inf
                lp[:, EOS + 1:] = -math.inf
            elif step < min_len:
                lp[:, EOS] = -math.inf
            top = lp.topk(2, dim=-1)
            nxt = top.indices[:, 0]
            for b in range(B):
                if not done[b]:
                    out[b].append(int(nxt[b]))
                    margins[b] = min(margins[b], float(top.values[b, 0] - top.values[b, 1]))
                    if int(nxt[b]) == EOS:
                        done[b] = True
            if all(done):
                break
            tokens = torch.cat((tokens, nxt.unsqueeze(1)), dim=1)
    return (out, margins) if return_margins else out
