"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the two training heads that consume the path's outputs.

  contrastive(audio, text, temp)   TripletSTMTContrastiveCriterion.compute_contrastive
                                   (fairseq/criterions/triplet_st_mt_contrastive.py:154-169)
  label_smoothed_nll(lprobs, ...)  label_smoothed_nll_loss
                                   (fairseq/criterions/label_smoothed_cross_entropy.py:13-30)

Pinned to the reference's own functions in tests/test_loss_oracle.py (run wherever the reference tree or the bundle
oracle/_ref/src is available).  The product never imports this file.
"""
import torch
import torch.nn.functional as F


def contrastive(audio, text, temp=0.1, reduce=True):
    """audio / text: [M, B, C] (encoder_out layout).  triplet_st_mt_contrastive.py:154-169 line by line.
    Note the reference hands F.cross_entropy a 3-D input [batch, i, j]: torch takes dim 1 (the AUDIO index i) as the class
    axis, so the softmax runs over the audio memories for every text memory j, target class j."""
    assert audio.shape == text.shape
    a = audio.transpose(0, 1)                                   # :156-157  [batch, seqlen, dim]
    t = text.transpose(0, 1)
    batch_size, seqlen, _ = a.shape
    logits = torch.cosine_similarity(a.float().unsqueeze(2), t.float().unsqueeze(1), dim=-1).type_as(a)   # :159-163
    logits = logits / temp                                      # :164
    target = torch.arange(seqlen)[None].repeat(batch_size, 1).to(logits.device)                            # :165-166
    return F.cross_entropy(logits, target, reduction="sum" if reduce else "none")                          # :167-168


def label_smoothed_nll(lprobs, target, epsilon, ignore_index=None, reduce=True):
    """label_smoothed_cross_entropy.py:13-30."""
    if target.dim() == lprobs.dim() - 1:
        target = target.unsqueeze(-1)
    nll_loss = -lprobs.gather(dim=-1, index=target)
    smooth_loss = -lprobs.sum(dim=-1, keepdim=True)
    if ignore_index is not None:
        pad_mask = target.eq(ignore_index)
        nll_loss = nll_loss.masked_fill(pad_mask, 0.0)
        smooth_loss = smooth_loss.masked_fill(pad_mask, 0.0)
    else:
        nll_loss = nll_loss.squeeze(-1)
        smooth_loss = smooth_loss.squeeze(-1)
    if reduce:
        nll_loss = nll_loss.sum()
        smooth_loss = smooth_loss.sum()
    eps_i = epsilon / lprobs.size(-1)
    return (1.0 - epsilon) * nll_loss + eps_i * smooth_loss, nll_loss
