"""Training heads over the path's outputs, host side of `csrc/loss.cu`.

`contrastive_loss` = TripletSTMTContrastiveCriterion.compute_contrastive (fairseq/criterions/
triplet_st_mt_contrastive.py:154-169) on the [M, B, C] memories of the audio and the text pass;
`label_smoothed_ce` = label_smoothed_nll_loss (fairseq/criterions/label_smoothed_cross_entropy.py:13-30) on decoder
logits.  Both return the per-row terms (`reduce=False`), their deterministic sum (`reduce=True`) and, when `grad_scale` is
given, the gradient of grad_scale * sum with respect to the inputs.  CUDA only (no CPU fallback).
"""
import torch

from . import _lib as L


def contrastive_loss(audio, text, temp=0.1, grad_scale=None):
    """-> (loss_rows [B, M], loss scalar tensor, d_audio, d_text) ; gradients are None unless grad_scale is given."""
    if not (audio.is_cuda and text.is_cuda):
        raise L.CstError("device tensors required (no CPU fallback)")
    if audio.shape != text.shape or audio.dim() != 3 or audio.dtype != text.dtype:
        raise ValueError("audio / text must be [M, B, C] tensors of one dtype")
    M, B, Cd = audio.shape
    audio, text = audio.contiguous(), text.contiguous()
    dev = audio.device
    rows = torch.empty(B, M, dtype=torch.float32, device=dev)
    lse = torch.empty(B * M, dtype=torch.float32, device=dev)
    total = torch.empty(1, dtype=torch.float32, device=dev)
    da = dt = None
    if grad_scale is not None:
        da = torch.empty(M, B, Cd, dtype=torch.float32, device=dev)
        dt = torch.empty(M, B, Cd, dtype=torch.float32, device=dev)
    lib = L.load()
    L.check(lib.cst_contrastive_loss(audio.data_ptr(), text.data_ptr(), L.DT[audio.dtype], M, B, Cd, float(temp), rows.data_ptr(),
                                     lse.data_ptr(), L.ptr(da), L.ptr(dt), float(grad_scale or 0.0), L.stream_ptr()))
    L.check(lib.cst_sum(rows.data_ptr(), B * M, total.data_ptr(), L.stream_ptr()))
    return rows, total[0], da, dt


def label_smoothed_ce(logits, target, eps=0.1, ignore_index=1, grad_scale=None):
    """logits [N, V] f32, target [N] int64 -> dict(loss_rows, nll_rows, loss, nll_loss, dlogits)."""
    if not (logits.is_cuda and target.is_cuda):
        raise L.CstError("device tensors required (no CPU fallback)")
    if logits.dim() != 2 or logits.dtype != torch.float32 or target.shape != (logits.shape[0],):
        raise ValueError("logits must be [N, V] float32 and target [N]")
    logits, target = logits.contiguous(), target.long().contiguous()
    N, V = logits.shape
    dev = logits.device
    loss_rows = torch.empty(N, dtype=torch.float32, device=dev)
    nll_rows = torch.empty(N, dtype=torch.float32, device=dev)
    sums = torch.empty(2, dtype=torch.float32, device=dev)
    d = torch.empty(N, V, dtype=torch.float32, device=dev) if grad_scale is not None else None
    lib = L.load()
    L.check(lib.cst_label_smoothed_ce(logits.data_ptr(), V, target.data_ptr(), N, V, float(eps), int(ignore_index),
                                      loss_rows.data_ptr(), nll_rows.data_ptr(), L.ptr(d), V, float(grad_scale or 0.0),
                                      L.stream_ptr()))
    L.check(lib.cst_sum(loss_rows.data_ptr(), N, sums[0:].data_ptr(), L.stream_ptr()))
    L.check(lib.cst_sum(nll_rows.data_ptr(), N, sums[1:].data_ptr(), L.stream_ptr()))
    return {"loss_rows": loss_rows, "nll_rows": nll_rows, "loss": sums[0], "nll_loss": sums[1], "dlogits": d}
