"""Tensor-level wrappers of the C-ABI entry points (device tensors in, device tensors out).

Used by the parity tests and tools; the execution plan (`plan.py`) calls the ABI directly on its
pre-allocated buffers.  Everything here launches CUDA kernels -- no computation happens in torch.
"""
import ctypes as C

import torch

from . import _lib as L
from .lengths import conv_out_lengths


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.CstError("device tensors required (no CPU fallback)")


def frame_lengths(src_lengths, Lw):
    _cuda(src_lengths)
    B = src_lengths.shape[0]
    Tp = conv_out_lengths(Lw)[-1]
    dev = src_lengths.device
    w2v = torch.empty(B, dtype=torch.int32, device=dev)
    sub = torch.empty(B, dtype=torch.int32, device=dev)
    l64 = torch.empty(B, dtype=torch.int64, device=dev)
    mask = torch.empty(B, Tp, dtype=torch.uint8, device=dev)
    L.check(L.load().cst_frame_lengths(src_lengths.contiguous().data_ptr(), B, Lw, Tp, w2v.data_ptr(), sub.data_ptr(),
                                       l64.data_ptr(), mask.data_ptr(), L.stream_ptr()))
    return w2v, sub, l64, mask.bool()


def conv0_gn_gelu(wave, w, gamma, beta, out_dtype=torch.float32, rows_per_seg=None, tensor_core=False):
    """wave [B,L] f32 -> channels-last [B, rows_per_seg, 512]; frames >= T0 are zero.
    tensor_core: run the convolution on tcgen05 (16-bit outputs only; 3-term fp16 split of x and w)."""
    _cuda(wave, w, gamma, beta)
    B, Lw = wave.shape
    T0 = (Lw - 10) // 5 + 1
    rps = rows_per_seg or T0
    dev = wave.device
    ss = torch.empty(B, 512, 2, dtype=torch.float32, device=dev)
    ws = torch.empty(B * 72, dtype=torch.float64, device=dev)
    out = torch.empty(B, rps, 512, dtype=out_dtype, device=dev)
    lib = L.load()
    w = w.reshape(512, 10).contiguous()
    L.check(lib.cst_conv0_stats(wave.data_ptr(), B, Lw, w.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                ss.data_ptr(), ws.data_ptr(), L.stream_ptr()))
    if tensor_core:
        hi = w.to(torch.float16)
        w16 = torch.zeros(512, 64, dtype=torch.float16, device=dev)
        w16[:, 0:10], w16[:, 10:20], w16[:, 20:30] = hi, hi, (w - hi.float()).to(torch.float16)
        L.check(lib.cst_conv0_apply_tc(wave.data_ptr(), B, Lw, w16.data_ptr(), ss.data_ptr(), out.data_ptr(),
                                       L.DT[out_dtype], rps, L.stream_ptr()))
        return out, ss
    L.check(lib.cst_conv0_apply(wave.data_ptr(), B, Lw, w.data_ptr(), ss.data_ptr(), out.data_ptr(),
                                L.DT[out_dtype], rps, L.stream_ptr()))
    return out, ss


def gemm(A, W, C_, M, N, K, lda, a_rows, bias=None, residual=None, act=L.ACT_NONE, alpha=1.0,
         rows_per_seg=None, seg_rows_valid=None, out_rows_per_seg=None, out_row_off=0, seg_len=None,
         ldc=None, ldr=None, nb_outer=1, nb_inner=1, a_bs=(0, 0), w_bs=0, c_bs=(0, 0), r_bs=None, bias_bs=0,
         segs_per_outer=1, ln_in=None, res_ln=None, c2=None, out_stats=None, ln_dim=0):
    _cuda(A, W, C_, bias, residual, seg_len)
    p = L.GemmParams()
    p.A, p.W, p.bias, p.residual, p.C = A.data_ptr(), W.data_ptr(), L.ptr(bias), L.ptr(residual), C_.data_ptr()
    p.ab_dtype, p.c_dtype = L.DT[A.dtype], L.DT[C_.dtype]
    p.M, p.N, p.K = M, N, K
    p.lda = lda
    p.ldc = ldc if ldc is not None else C_.shape[-1]
    p.ldr = ldr if ldr is not None else (residual.shape[-1] if residual is not None else 0)
    p.a_rows = a_rows
    p.act, p.alpha = act, alpha
    p.nb_outer, p.nb_inner = nb_outer, nb_inner
    p.a_bs_outer, p.a_bs_inner = a_bs
    p.w_bs_inner = w_bs
    p.c_bs_outer, p.c_bs_inner = c_bs
    p.r_bs_outer, p.r_bs_inner = r_bs if r_bs is not None else c_bs
    p.bias_bs_inner = bias_bs
    p.rows_per_seg = rows_per_seg if rows_per_seg is not None else M
    p.seg_rows_valid = seg_rows_valid if seg_rows_valid is not None else p.rows_per_seg
    p.out_rows_per_seg = out_rows_per_seg if out_rows_per_seg is not None else p.rows_per_seg
    p.out_row_off = out_row_off
    p.seg_len = L.ptr(seg_len)
    p.segs_per_outer = segs_per_outer
    if ln_in is not None:            # (stats [rows,8,2] f32, colsum [N] f32, slots)
        p.ln_in_stats, p.ln_colsum, p.ln_in_slots = ln_in[0].data_ptr(), ln_in[1].data_ptr(), ln_in[2]
    if res_ln is not None:           # (stats, slots, gamma, beta)
        p.res_stats, p.res_slots, p.res_gamma, p.res_beta = res_ln[0].data_ptr(), res_ln[1], res_ln[2].data_ptr(), res_ln[3].data_ptr()
    if c2 is not None:
        p.C2, p.c2_dtype, p.ldc2 = c2.data_ptr(), L.DT[c2.dtype], c2.shape[-1]
    if out_stats is not None:
        p.out_stats = out_stats.data_ptr()
    p.ln_dim = ln_dim
    L.check(L.load().cst_gemm(C.byref(p), L.stream_ptr()))
    return C_


def linear(A, W, bias=None, act=L.ACT_NONE, residual=None, alpha=1.0, out_dtype=None):
    """act(A W^T + b) * alpha (+ residual) for dense row-major A [M,K], W [N,K]."""
    M, K = A.shape
    N = W.shape[0]
    n_out = N // 2 if act == L.ACT_GLU else N
    out = torch.empty(M, n_out, dtype=out_dtype or A.dtype, device=A.device)
    return gemm(A, W, out, M, N, K, lda=K, a_rows=M, bias=bias, residual=residual, act=act, alpha=alpha)


def layernorm(x, gamma, beta, lp_dtype=None):
    _cuda(x, gamma, beta)
    rows, Cd = x.shape
    o32 = torch.empty_like(x)
    olp = torch.empty(rows, Cd, dtype=lp_dtype, device=x.device) if lp_dtype is not None else None
    L.check(L.load().cst_layernorm(x.data_ptr(), Cd, gamma.data_ptr(), beta.data_ptr(), o32.data_ptr(), L.ptr(olp),
                                   L.DT[lp_dtype] if lp_dtype is not None else 0, Cd, rows, Cd, rows, rows, rows, 0, 0,
                                   L.stream_ptr()))
    return o32, olp


def attention(q, k, v, n_heads, kv_len=None):
    """q [B,Tq,C], k/v [B,Tk,C] contiguous (C = n_heads*64), q pre-scaled; kv_len int32 [B] or None."""
    _cuda(q, k, v, kv_len)
    B, Tq, Cd = q.shape
    Tk = k.shape[1]
    out = torch.empty_like(q)
    L.check(L.load().cst_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), L.DT[q.dtype],
                                   Cd, Cd, Cd, B, n_heads, Tq, Tq, Tk, Tk, L.ptr(kv_len), L.stream_ptr()))
    return out


# ---- greedy decoding kernels (cst_dec_*) ------------------------------------------------------------------------------
def dec_linear(A, W, bias=None, ln=None, residual=None, act=L.ACT_NONE, outs=None, ldo=None, step_stride=None, step=None):
    """act(LN?(A) W^T + b) (+ residual) -> f32.  outs: list of 1-3 output tensors (N split evenly); default one [M,N]."""
    _cuda(A, W, bias, residual, step)
    M, K = A.shape
    N = W.shape[0]
    if outs is None:
        outs = [torch.empty(M, N, dtype=torch.float32, device=A.device)]
    p = L.DecLinearParams()
    p.A, p.W, p.bias, p.residual = A.data_ptr(), W.data_ptr(), L.ptr(bias), L.ptr(residual)
    p.ln_gamma, p.ln_beta = (ln[0].data_ptr(), ln[1].data_ptr()) if ln is not None else (0, 0)
    p.lda, p.ldr = A.stride(0), (residual.stride(0) if residual is not None else 0)
    for s in range(3):
        p.out[s] = outs[s].data_ptr() if s < len(outs) else 0
        p.out_dtype[s] = L.DT[outs[s].dtype] if s < len(outs) else L.F32
        p.ldo[s] = (ldo[s] if ldo is not None else N // len(outs)) if s < len(outs) else 0
        p.step_stride[s] = step_stride[s] if (step_stride is not None and s < len(outs)) else 0
    p.step = L.ptr(step)
    p.a_dtype, p.w_dtype = L.DT[A.dtype], L.DT[W.dtype]
    p.M, p.N, p.K, p.n_seg, p.act = M, N, K, len(outs), act
    L.check(L.load().cst_dec_linear(C.byref(p), L.stream_ptr()))
    return outs[0] if len(outs) == 1 else outs


def dec_attention(q, k, v, kv_batch_stride, kv_row_stride, n_heads, n_keys, n_keys_max, step=None):
    """q [B, H*64] f32 (pre-scaled); k / v f32 or bf16; key j of row b at k + b*kv_batch_stride + j*kv_row_stride (elements)."""
    _cuda(q, k, v, step)
    B = q.shape[0]
    out = torch.empty_like(q)
    L.check(L.load().cst_dec_attention(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), L.DT[k.dtype], kv_batch_stride, kv_row_stride,
                                       out.data_ptr(), out.stride(0), B, n_heads, n_keys, n_keys_max, L.ptr(step),
                                       L.stream_ptr()))
    return out


def dec_select(logits, tokens, pos_scores, done, out_len, counters, max_len, min_len=1, pad=1, eos=2):
    _cuda(logits, tokens, pos_scores, done, out_len, counters)
    B, V = logits.shape
    L.check(L.load().cst_dec_select(logits.data_ptr(), V, B, tokens.data_ptr(), tokens.stride(0), pos_scores.data_ptr(),
                                    pos_scores.stride(0), done.data_ptr(), out_len.data_ptr(), counters.data_ptr(),
                                    max_len, min_len, pad, eos, L.stream_ptr()))
