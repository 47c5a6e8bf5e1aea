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


def beam_search(sd, memories, beam=5, max_len=200, min_len=1, len_penalty=1.0):
    """SequenceGenerator._generate + BeamSearch.step + finalize_hypos restated per sentence (sentences never interact;
    the reference's removal of finished sentences from the batch is an optimisation), fairseq/sequence_generator.py:
    179-540,590-712 and fairseq/search.py:109-160.  Default options only (normalize_scores, no unk penalty, no prefix,
    no n-gram blocking, temperature 1).
    -> per sentence a list (best first) of dicts {tokens (ends in EOS), score, positional_scores}."""
    results = []
    V = sd["decoder.embed_tokens.weight"].shape[0]
    K, cand_size = beam, 2 * beam
    with torch.no_grad():
        for b in range(memories.shape[1]):
            mem = memories[:, b:b + 1].expand(-1, K, -1).contiguous()
            tokens = torch.full((K, max_len + 2), PAD, dtype=torch.long)
            tokens[:, 0] = EOS
            scores = torch.zeros(K, max_len + 1)
            ignore = torch.zeros(K, dtype=torch.bool)                          # cands_to_ignore
            finalized = []
            for step in range(max_len + 1):
                lp = torch.log_softmax(decoder_logits(sd, tokens[:, :step + 1], mem).float(), dim=-1)
                lp[lp != lp] = -math.inf
                lp[:, PAD] = -math.inf
                if step >= max_len:
                    lp[:, :EOS] = -math.inf
                    lp[:, EOS + 1:] = -math.inf
                elif step < min_len:
                    lp[:, EOS] = -math.inf
                # BeamSearch.step: first step uses only the first beam (all hypotheses are identical)
                cand = lp[:1] if step == 0 else lp + scores[:, step - 1:step]
                top = torch.topk(cand.reshape(-1), k=min(cand_size, cand.numel() - 1))
                cand_scores, idx = top.values, top.indices
                cand_beams, cand_tok = idx // V, idx.fmod(V)
                eos_mask = cand_tok.eq(EOS) & cand_scores.ne(-math.inf)
                eos_mask[:K][ignore] = False
                # finalize_hypos: EOS among the top `beam` candidates ends that hypothesis
                for i in range(K):
                    if eos_mask[i] and len(finalized) < K:
                        src = int(cand_beams[i])
                        toks = tokens[src, 1:step + 2].clone()
                        toks[step] = EOS
                        pos = scores[src, :step + 1].clone()
                        pos[step] = cand_scores[i]
                        pos[1:] = pos[1:] - pos[:-1]
                        finalized.append({"tokens": toks, "score": float(cand_scores[i]) / (step + 1) ** len_penalty,
                                          "positional_scores": pos})
                if len(finalized) == K or step == max_len:                      # is_finished
                    break
                # the `beam` best candidates that are not EOS (and not ignored) continue
                eos_mask[:K] = ~((~ignore) & (~eos_mask[:K]))
                active_mask = eos_mask.long() * cand_size + torch.arange(cand_size)[:eos_mask.numel()]
                new_ignore, active = torch.topk(active_mask, k=K, largest=False)
                ignore = new_ignore.ge(cand_size)
                src = cand_beams[active]
                tokens[:, :step + 1] = tokens[src, :step + 1]
                tokens[:, step + 1] = cand_tok[active]
                if step > 0:
                    scores[:, :step] = scores[src, :step]
                scores[:, step] = cand_scores[active]
            order = torch.sort(torch.tensor([h["score"] for h in finalized]), descending=True).indices.tolist()
            results.append([finalized[i] for i in order])
    return results
