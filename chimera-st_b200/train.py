"""Forward + backward of the speech-encoding path (BASELINE configs[4]), first correct version.

`EncoderTrainStep(state_dict, B, L, M).forward_backward(wave, lens, d_memories)` runs the audio branch of
`S2T_W2V2_TransformerInterlinguaEncoder.forward` (fairseq/models/chimera/w2v2_transformer_interlingua.py:207-312) with every
activation the backward pass needs kept in HBM, then the hand-written backward kernels (`csrc/backward.cu`,
`csrc/attention_bwd.cu`, `csrc/conv0_bwd.cu`, dgrad / wgrad through `cst_gemm` on `cst_transpose`-d operands), and returns the
gradient of every encoder parameter under its REFERENCE state-dict name (so that the reference's optimizer / DDP wrapper --
`fairseq/legacy_distributed_data_parallel.py:94-178`, restated in `ddp.py` -- can consume them).

Scope of this version: fp32 arithmetic (CUDA-core GEMM / attention; parity <= 1e-4 against autograd through the
oracle), dropout = 0 and LayerDrop = 0 (the parity configuration of SURVEY §8(d) C5), one padded batch per step, unfused
activations in the forward pass (the pre-activation is kept).  `GradMultiply(feature_grad_mult)` (wav2vec2.py:530-532) is applied
to the feature-extractor gradients.  torch is used for buffers and for re-arranging WEIGHT tensors between the reference layout and
the kernel layout; everything that touches an activation is a kernel behind the C ABI.
"""
import ctypes as C
import math
import os

import torch

from . import _lib as L
from . import weights as _weights
from .plan import Geometry, SLACK
from .synth import W2V_DIM, W2V_FFN, W2V_HEADS, W2V_LAYERS, ENC_DIM, ENC_FFN, ENC_HEADS, ENC_LAYERS, MEM_LAYERS, POS_K, POS_GROUPS

F32 = torch.float32


_CODE = {torch.float32: L.F32, torch.bfloat16: L.BF16, torch.float16: L.F16}
SM_COUNT = 148


class _Ops:
    """Thin launch helpers over the C ABI (device tensors in / out, current stream).  `op` is the dtype of GEMM operands and of the
    pre-activations kept on the tape (fp32: the parity mode; bf16: BASELINE configs[4]); gradients of activations are always fp32."""

    def __init__(self, device, op=F32, lib=None):
        # `lib` is injectable so tests can drive the launch sequence against the host emulator of the C ABI (tests/emu.py); the product
        # never passes it and always loads the CUDA library.
        self.lib = lib if lib is not None else L.load()
        self.dev = device
        self.op = op
        self.opc = _CODE[op]
        self.esz = 4 if op == F32 else 2
        self.tc_attn_bwd = os.environ.get("CST_ATTN_BWD_TC", "1") != "0"        # A/B lever: 0 = FFMA attention backward in the 16-bit mode too
        self.ws = torch.empty(512 * 1024, dtype=F32, device=device)             # cst_colsum scratch (CST_COLSUM_WS_FLOATS)

    def st(self):
        return L.stream_ptr() if self.dev.type == "cuda" else 0

    def new(self, *shape, zero=False, dtype=F32):
        return (torch.zeros if zero else torch.empty)(*shape, dtype=dtype, device=self.dev)

    def gemm(self, A, W, Cout, M, N, K, lda, a_rows, bias=None, residual=None, rows_per_seg=None, seg_len=None,
             nb_outer=1, nb_inner=1, a_bs=(0, 0), w_bs=0, c_bs=(0, 0), bias_bs=0, ldc=None):
        assert A.dtype == W.dtype, (A.dtype, W.dtype)
        p = L.GemmParams()
        p.A, p.W, p.bias, p.residual, p.C = A.data_ptr(), W.data_ptr(), L.ptr(bias), L.ptr(residual), Cout.data_ptr()
        p.ab_dtype, p.c_dtype = _CODE[A.dtype], _CODE[Cout.dtype]
        p.M, p.N, p.K, p.lda = M, N, K, lda
        p.ldc = ldc if ldc is not None else Cout.shape[-1]
        p.ldr = p.ldc if residual is not None else 0
        p.a_rows, p.act, p.alpha = a_rows, L.ACT_NONE, 1.0
        p.nb_outer, p.nb_inner = nb_outer, nb_inner
        p.a_bs_outer, p.a_bs_inner = a_bs
        p.w_bs_inner = w_bs
        p.c_bs_outer, p.c_bs_inner = c_bs
        p.r_bs_outer, p.r_bs_inner = c_bs
        p.bias_bs_inner = bias_bs
        p.rows_per_seg = rows_per_seg if rows_per_seg is not None else M
        p.seg_rows_valid = p.rows_per_seg
        p.out_rows_per_seg = p.rows_per_seg
        p.seg_len = L.ptr(seg_len)
        p.segs_per_outer = 1
        L.check(self.lib.cst_gemm(C.byref(p), self.st()))
        return Cout

    def linear(self, x, W, b=None, residual=None, rows=None, out_dtype=F32):
        rows = rows if rows is not None else x.shape[0]
        N, K = W.shape
        return self.gemm(x, W, self.new(rows, N, dtype=out_dtype), rows, N, K, lda=x.shape[1], a_rows=x.shape[0], bias=b, residual=residual)

    def transpose(self, x, rows, cols, ldx=None, pad=64, chunks=1, copy=False, out_dtype=None):
        """-> x^T [chunks, cols, rows_pad / chunks] in `out_dtype` (default: the operand dtype), zero-filled beyond `rows`
        (+ the un-transposed cast copy [rows, cols] when copy=True)."""
        od = out_dtype or self.op
        per = ((rows + chunks - 1) // chunks + pad - 1) // pad * pad
        rp = per * chunks
        out = self.new(chunks, cols, per, dtype=od)
        cp = self.new(rows, cols, dtype=od) if copy else None
        L.check(self.lib.cst_transpose(x.data_ptr(), _CODE[x.dtype], ldx if ldx is not None else x.shape[1], rows, cols, out.data_ptr(), _CODE[od],
                                       rp, per, L.ptr(cp), cols, self.st()))
        return (out, cp) if copy else out

    def colsum(self, x, rows, cols, ldx=None, scale=1.0):
        out = self.new(cols)
        step = 8192 if rows >= 256 else cols                       # only the two-level reduction of tall inputs uses the scratch buffer
        for c0 in range(0, cols, step):
            n = min(step, cols - c0)
            L.check(self.lib.cst_colsum(x.data_ptr() + c0 * x.element_size(), _CODE[x.dtype], ldx if ldx is not None else x.shape[1], rows, n,
                                        out.data_ptr() + 4 * c0, self.ws.data_ptr(), scale, self.st()))
        return out

    def wgrad(self, dy, x, rows, N, K, ldx=None):
        """dW [N, K] = dy[:rows]^T x[:rows] (fp32), dy fp32 [rows, N], x operand dtype (row pitch ldx: < K for the overlapping windows
        of an implicit-GEMM convolution).  Also returns the operand-dtype copy of dy for the dgrad GEMM.  When the output has few
        tiles and the reduction is long (the conv stack: 24 tiles, > 1e5 rows) the row axis is split into a GEMM batch (split-K) and
        the partial products are added by cst_colsum in a fixed order."""
        tiles = ((N + 127) // 128) * ((K + 255) // 256)
        S = max(1, min(SM_COUNT // tiles, rows // 1024))
        if self.op == F32:
            dyT, dy_op = self.transpose(dy, rows, N, chunks=S), dy
        else:
            dyT, dy_op = self.transpose(dy, rows, N, chunks=S, copy=True)
        xT = self.transpose(x, rows, K, ldx=ldx, chunks=S)
        per = dyT.shape[2]
        if S == 1:
            dW = self.gemm(dyT, xT, self.new(N, K), N, K, per, lda=per, a_rows=N)
        else:
            part = self.gemm(dyT, xT, self.new(S, N * K), N, K, per, lda=per, a_rows=N, nb_inner=S, a_bs=(0, N * per), w_bs=K * per,
                             c_bs=(0, N * K), ldc=K)
            dW = self.colsum(part, S, N * K).view(N, K)
        return dW, dy_op

    def linear_bwd(self, x, W, dy, rows, need_dx=True, dx_residual=None, bias=True, ldx=None, dx_dtype=F32):
        """y = x W^T + b  ->  (dx = dy W (+ dx_residual), dW = dy^T x, db = colsum(dy)); x / W in the operand dtype, dy fp32."""
        N, K = W.shape
        dW, dy_op = self.wgrad(dy, x, rows, N, K, ldx=ldx)
        dx = None
        if need_dx:
            Wt = self.transpose(W, N, K)[0]                        # [K, N] (N is a multiple of 64 on this path)
            dx = self.gemm(dy_op, Wt, self.new(rows, K, dtype=dx_dtype), rows, K, N, lda=dy_op.shape[1], a_rows=dy_op.shape[0],
                           residual=dx_residual)
        db = self.colsum(dy, rows, N) if bias else None
        return dx, dW, db

    def act(self, kind, z, rows, cols_out, alpha=1.0, out_dtype=None, out=None):
        y = out if out is not None else self.new(rows, cols_out, dtype=out_dtype or self.op)
        L.check(self.lib.cst_act_fwd(kind, z.data_ptr(), _CODE[z.dtype], z.shape[1], rows, cols_out, y.data_ptr(), _CODE[y.dtype], y.shape[1],
                                     alpha, self.st()))
        return y

    def act_bwd(self, kind, z, dy, rows, cols_out, alpha=1.0):
        dz = self.new(rows, z.shape[1])
        L.check(self.lib.cst_act_bwd(kind, z.data_ptr(), _CODE[z.dtype], z.shape[1], dy.data_ptr(), _CODE[dy.dtype], dy.shape[1], rows, cols_out,
                                     dz.data_ptr(), L.F32, z.shape[1], alpha, self.st()))
        return dz

    def dropout(self, x, rows, cols, p, seed, site, add=None, out_dtype=F32, lp_dtype=None):
        """-> add + x * keep / (1 - p) (and its `lp_dtype` copy); the mask is a function of (*seed, site, element index): calling this on a
        gradient with the same seed / site is the derivative (csrc/philox.cuh)."""
        out = self.new(rows, cols, dtype=out_dtype)
        lp = self.new(rows, cols, dtype=lp_dtype) if lp_dtype is not None else None
        L.check(self.lib.cst_dropout(x.data_ptr(), _CODE[x.dtype], x.shape[1], L.ptr(add), add.shape[1] if add is not None else 0,
                                     out.data_ptr(), _CODE[out_dtype], cols, L.ptr(lp), _CODE[lp_dtype] if lp is not None else 0, cols,
                                     rows, cols, float(p), seed.data_ptr(), int(site), self.st()))
        return out if lp is None else (out, lp)

    def ln(self, x, g, b, rows, f32=True):
        """-> (fp32 normalised rows or None, operand-dtype copy); in the fp32 mode both are the same tensor."""
        Cd = x.shape[1]
        lowp = self.op != F32
        out = self.new(rows, Cd) if (f32 or not lowp) else None
        lp = self.new(rows, Cd, dtype=self.op) if lowp else None
        L.check(self.lib.cst_layernorm(x.data_ptr(), Cd, g.data_ptr(), b.data_ptr(), L.ptr(out), L.ptr(lp), self.opc if lowp else L.F32, Cd, rows,
                                       Cd, rows, rows, rows, 0, 0, self.st()))
        return out, (lp if lowp else out)

    def ln_bwd(self, x, g, dy, rows, dx=None):
        """-> (dx (accumulated into `dx` when given), dgamma, dbeta)"""
        Cd = x.shape[1]
        acc = dx is not None
        dx = dx if acc else self.new(rows, Cd)
        nblk = (rows + 7) // 8
        part = self.new(nblk, 2 * Cd)
        L.check(self.lib.cst_layernorm_bwd(x.data_ptr(), Cd, g.data_ptr(), dy.data_ptr(), dy.shape[1], dx.data_ptr(), Cd, part.data_ptr(), rows,
                                           Cd, 1 if acc else 0, self.st()))
        gb = self.colsum(part, nblk, 2 * Cd)
        return dx, gb[:Cd], gb[Cd:]

    def _ws(self, nbytes):
        return torch.empty(int(nbytes), dtype=torch.uint8, device=self.dev)

    def attention(self, q, k, v, ldq, ldkv, B, H, n_q, q_rps, n_kv, kv_rps, kv_len, out_rows, out_cols, drop=None):
        """drop = (p, seed tensor, site): dropout of the attention probabilities -- the materialised form on batched tcgen05 GEMMs
        (csrc/attention_bwd_tc.cu; bf16 panels in both modes) instead of the flash kernel."""
        out = self.new(out_rows, out_cols, zero=True, dtype=self.op)
        if drop is not None and drop[0] > 0:
            ws = self._ws(self.lib.cst_attention_dropout_fwd_ws_bytes(B, H, n_q, n_kv))
            L.check(self.lib.cst_attention_dropout_fwd(q, k, v, self.opc, out.data_ptr(), self.opc, ldq, ldkv, out_cols, B, H, n_q, q_rps, n_kv,
                                                       kv_rps, L.ptr(kv_len), float(drop[0]), drop[1].data_ptr(), int(drop[2]),
                                                       ws.data_ptr(), self.st()))
            return out
        L.check(self.lib.cst_attention(q, k, v, out.data_ptr(), self.opc, ldq, ldkv, out_cols, B, H, n_q, q_rps, n_kv, kv_rps, L.ptr(kv_len),
                                       self.st()))
        return out

    def attention_bwd(self, q, k, v, o, do, dq, dk, dv, ldq, ldkv, ldo, B, H, n_q, q_rps, n_kv, kv_rps, kv_len, dtype=None, drop=None):
        """q / k / v / o: forward tensors (operand dtype, row strides ldq / ldkv / ldo); do, dq, dk, dv fp32 with the SAME row strides.
        16-bit mode: batched tcgen05 GEMMs + one softmax-backward kernel (csrc/attention_bwd_tc.cu); fp32: the FFMA parity kernel."""
        if drop is not None and drop[0] > 0:                       # same masks as the forward (regenerated from seed / site)
            ws = self._ws(self.lib.cst_attention_bwd_tc_ws_bytes(B, H, n_q, n_kv))
            L.check(self.lib.cst_attention_bwd_tc_dropout(q, k, v, self.opc, do.data_ptr(), dq, dk, dv, ldq, ldkv, ldo, ldq, ldkv, B, H, n_q,
                                                          q_rps, n_kv, kv_rps, L.ptr(kv_len), float(drop[0]), drop[1].data_ptr(), int(drop[2]),
                                                          ws.data_ptr(), self.st()))
            return
        if self.op == torch.bfloat16 and dtype is None and self.tc_attn_bwd:
            ws = torch.empty(int(self.lib.cst_attention_bwd_tc_ws_bytes(B, H, n_q, n_kv)), dtype=torch.uint8, device=self.dev)
            L.check(self.lib.cst_attention_bwd_tc(q, k, v, do.data_ptr(), dq, dk, dv, ldq, ldkv, ldo, ldq, ldkv, B, H, n_q, q_rps, n_kv, kv_rps,
                                                  L.ptr(kv_len), ws.data_ptr(), self.st()))
            return
        L.check(self.lib.cst_attention_bwd(q, k, v, o.data_ptr(), _CODE[o.dtype] if dtype is None else dtype, do.data_ptr(), dq, dk, dv,
                                           ldq, ldkv, ldo, ldo, ldq, ldkv, B, H, n_q, q_rps, n_kv, kv_rps, L.ptr(kv_len), self.st()))

    def remap(self, src, in_rps, in_off, dst, out_rps, out_off, n_seg, n_rows, Cd, valid, seg_len=None, accumulate=False, scale=1.0):
        L.check(self.lib.cst_rows_remap(src.data_ptr(), src.shape[1], in_rps, in_off, dst.data_ptr(), _CODE[dst.dtype], dst.shape[1], out_rps,
                                        out_off, n_seg, n_rows, Cd, valid, L.ptr(seg_len), 1 if accumulate else 0, scale, self.st()))
        return dst

    def col2im(self, dcol, M, k, stride, Cd, rows_in):
        dx = self.new(rows_in, Cd)
        L.check(self.lib.cst_col2im(dcol.data_ptr(), _CODE[dcol.dtype], M, k, stride, Cd, dx.data_ptr(), rows_in, 0, self.st()))
        return dx


def _attn_layer_grads(G, name, dqkv_w, dqkv_b, D, fused=True):
    """Kernel layout -> reference names for an attention block's in-projection (q rows carry the folded 1/8)."""
    if fused:
        G[name + "q_proj.weight"], G[name + "k_proj.weight"], G[name + "v_proj.weight"] = dqkv_w[:D] * 0.125, dqkv_w[D:2 * D], dqkv_w[2 * D:]
        G[name + "q_proj.bias"], G[name + "k_proj.bias"], G[name + "v_proj.bias"] = dqkv_b[:D] * 0.125, dqkv_b[D:2 * D], dqkv_b[2 * D:]


class EncoderTrainStep:
    def __init__(self, state_dict, B, Lw, M=None, device="cuda", feature_grad_mult=0.1, dtype=F32, lib=None, dropout=0.0,
                 activation_dropout=None, attention_dropout=None, w2v_dropout=0.0, w2v_attention_dropout=None, w2v_dropout_input=0.0, seed=1):
        dev = torch.device(device)
        if dev.type != "cuda" and lib is None:
            raise L.CstError("EncoderTrainStep runs only on a CUDA device (no CPU fallback)")
        if dtype not in (F32, torch.bfloat16):
            raise L.CstError("EncoderTrainStep: dtype must be torch.float32 or torch.bfloat16")
        sd = _weights._strip(state_dict)
        if any(k.startswith(("audio_exclusive_layers.", "modal_embedding.")) for k in sd):
            raise NotImplementedError("EncoderTrainStep: modal_embedding / non_shared_encoder_layers checkpoints are inference-only here")
        # fp32 master parameters under the reference names; tensors that already are fp32 on `dev` are NOT copied: an optimizer updating
        # `self.sd` in place (FusedAdam) then updates the caller's tensors too -- pass clones to keep the originals
        self.sd = {k: v.detach().to(dev, F32) for k, v in sd.items() if v.is_floating_point()}
        self.M = M or self.sd["interlingua_embedding.weight"].shape[0]
        self.g = Geometry(B, Lw, self.M)
        self.dev = dev
        self.op = dtype
        self.o = _Ops(dev, dtype, lib=lib)
        self.fgm = float(feature_grad_mult)
        # 16-bit mode: the normalisation-free conv stack runs its FORWARD GEMMs on fp16 operands like the inference path (DESIGN.md §2:
        # bf16 there costs 6e-3 of the 1e-2 budget); every backward operand (transposed copies of weights, activations, gradients)
        # is bf16 -- gradients need the range, and cst_transpose converts on the fly.
        self.cdt = dtype if dtype == F32 else torch.float16
        self.P = _weights.prepare(self.sd, dev, dtype, conv_dtype=self.cdt, training=True)
        # Dropout of the training recipe (train-en2any-ST.sh:45 --dropout 0.1; wav2vec2 base: dropout 0.1, dropout_input 0.1,
        # activation_dropout 0): `dropout` = the shared / memory layers' dropout_module + the dropout after the embedding scale
        # (w2v2_transformer_interlingua.py:237), `activation_dropout` their activation_dropout_module (defaults to `dropout`,
        # w2v2_transformer.py:460), `w2v_dropout` = wav2vec2's encoder dropout (wav2vec2.py:830) and dropout1 / dropout3 of its layers,
        # `w2v_dropout_input` = wav2vec2.py:553.  0 everywhere = the parity configuration.  Masks are regenerated from (seed, site) in the
        # backward pass, never stored; `seed_dev` lives on the device so CUDA-graph replays draw fresh masks (`next_dropout_seed`).
        # `attention_dropout` (defaults to `dropout`, w2v2_transformer.py:459) / `w2v_attention_dropout` (defaults to `w2v_dropout`; 0.1
        # in wav2vec2 base) drop the attention probabilities: those layers run the materialised attention of csrc/attention_bwd_tc.cu.
        self.pd = {"shared": float(dropout), "act": float(dropout if activation_dropout is None else activation_dropout),
                   "attn": float(dropout if attention_dropout is None else attention_dropout),
                   "w2v": float(w2v_dropout), "w2v_attn": float(w2v_dropout if w2v_attention_dropout is None else w2v_attention_dropout),
                   "w2v_input": float(w2v_dropout_input)}
        self._sites = {}
        self.seed_dev = torch.tensor([int(seed)], dtype=torch.int64).to(dev)

    def next_dropout_seed(self):
        """New masks for the next step: the seed is incremented ON the device, in stream order (graph replays read the device value; a
        host-side counter copied asynchronously could be overwritten before an earlier copy ran when the host runs ahead)."""
        self.seed_dev.add_(1)

    def _site(self, T, tag):
        return self._sites.setdefault((T.get("pass", 0), tag), len(self._sites))

    def _drop(self, T, tag, p, x, rows, cols, **kw):
        return self.o.dropout(x, rows, cols, p, self.seed_dev, self._site(T, tag), **kw)

    def _adrop(self, T, tag, p):
        return (p, self.seed_dev, self._site(T, tag)) if p > 0 else None

    # ------------------------------------------------------------------ forward (activations kept)
    @staticmethod
    def sample_layerdrop(p, rng=None):
        """The reference's LayerDrop draw for one step (TransformerEncoder.extract_features, wav2vec2.py:835-838): one host-side uniform
        number per wav2vec2 layer, the layer runs iff it exceeds `p` (encoder_layerdrop, 0.05 in the base recipe).  -> set of skipped
        layer indices for `forward(..., skip_w2v_layers=...)`."""
        import numpy as np
        draw = (rng or np.random).random
        return frozenset(i for i in range(W2V_LAYERS) if not draw() > p)

    def forward(self, wave, lens, skip_w2v_layers=()):
        """skip_w2v_layers: wav2vec2 layers LayerDrop removed for this step (`sample_layerdrop`): they are not run, their parameters get
        no gradient (absent from the result; the all-reduce counts them as zeros, legacy_distributed_data_parallel.py:141-143)."""
        o, g, P, lib = self.o, self.g, self.P, self.o.lib
        B, st, op, opc, esz = g.B, self.o.st(), self.op, self.o.opc, self.o.esz
        T = {}                                                     # the tape
        self.T = T
        wave = wave.to(self.dev, F32).contiguous()
        src_len = lens.to(self.dev, torch.int64).contiguous()
        T["wave"] = wave
        i32 = dict(dtype=torch.int32, device=self.dev)
        w2v_valid, sub_valid = torch.empty(B, **i32), torch.empty(B, **i32)
        L.check(lib.cst_frame_lengths(src_len.data_ptr(), B, g.L, g.Tp, w2v_valid.data_ptr(), sub_valid.data_ptr(), 0, 0, st))
        T["w2v_valid"], T["sub_valid"] = w2v_valid, sub_valid
        # ---- conv feature extractor
        ss = o.new(B * 512, 2)
        stats_ws = torch.zeros(B * 72, dtype=torch.float64, device=self.dev)
        L.check(lib.cst_conv0_stats(wave.data_ptr(), B, g.L, P["conv0_w"].data_ptr(), P["gn_g"].data_ptr(), P["gn_b"].data_ptr(),
                                    ss.data_ptr(), stats_ws.data_ptr(), st))
        cdt = self.cdt
        c = o.new(B * g.Ta[0] + SLACK, 512, zero=True, dtype=cdt)
        L.check(lib.cst_conv0_apply(wave.data_ptr(), B, g.L, P["conv0_w"].data_ptr(), ss.data_ptr(), c.data_ptr(), _CODE[cdt], g.Ta[0], st))
        T["ss"] = ss
        T["c"], T["cz"] = [c], [None]
        for i in range(1, 7):
            w = P[f"conv{i}_w"]
            rows = B * g.Ta[i]
            z = o.gemm(c, w, o.new(rows, 512, dtype=cdt), rows, 512, w.shape[1], lda=1024, a_rows=(B * g.Ta[i - 1] + SLACK) // 2)
            cn = o.new(rows + SLACK, 512, zero=True, dtype=(cdt if i < 6 else F32))     # the last block feeds a LayerNorm: fp32
            o.act(L.ACT_GELU, z, rows, 512, out=cn)
            T["c"].append(cn)
            T["cz"].append(z)
            c = cn
        R = B * g.T6a
        feat = c
        _, feat_ln = o.ln(feat, P["ln_feat_g"], P["ln_feat_b"], R, f32=False)
        xp = o.gemm(feat_ln, P["proj_w"], o.new(R, W2V_DIM), R, W2V_DIM, 512, lda=512, a_rows=R, bias=P["proj_b"], rows_per_seg=g.T6a,
                    seg_len=w2v_valid)
        if self.pd["w2v_input"] > 0:                               # dropout_input (wav2vec2.py:553); padded rows stay zero
            xp = self._drop(T, "w2v.input", self.pd["w2v_input"], xp, R, W2V_DIM)
        T["feat_ln"], T["xp"] = feat_ln, xp
        # ---- pos-conv (grouped implicit GEMM on the packed operand), GELU, residual
        xg = torch.zeros(B * 16 * g.Tpp + SLACK, 64, dtype=op, device=self.dev)
        L.check(lib.cst_posconv_pack(xp.data_ptr(), B, g.T6a, g.Tp, xg.data_ptr(), opc, g.Tpp, st))
        zpos = o.gemm(xg, P["pos_w"], o.new(R, W2V_DIM), g.T6a, 48, 128 * 64, lda=64, a_rows=g.Tpp, bias=P["pos_b"], nb_outer=B,
                      nb_inner=16, a_bs=(16 * g.Tpp * 64, g.Tpp * 64), w_bs=48 * 128 * 64, c_bs=(g.T6a * W2V_DIM, 48), bias_bs=48, ldc=W2V_DIM)
        y0 = xp.clone()
        gp = o.act(L.ACT_GELU, zpos, R, W2V_DIM, out_dtype=F32)
        o.remap(gp, R, 0, y0, R, 0, 1, R, W2V_DIM, R, accumulate=True)
        T["xg"], T["zpos"], T["y0"] = xg, zpos, y0
        x, x_op = o.ln(y0, P["ln_enc_g"], P["ln_enc_b"], R)
        pw = self.pd["w2v"]
        if pw > 0:                                                 # F.dropout after the encoder LayerNorm (wav2vec2.py:830)
            x = self._drop(T, "w2v.enc", pw, x, R, W2V_DIM, lp_dtype=None if op == F32 else op)
            x, x_op = (x, x) if op == F32 else x
        # ---- 12 post-LN wav2vec2 layers
        D = W2V_DIM
        T["w2v"] = []
        for li, lw in enumerate(P["w2v_layers"]):
            if li in skip_w2v_layers:
                T["w2v"].append(None)
                continue
            t = {"x_op": x_op}
            qkv = o.linear(x_op, lw["qkv_w"], lw["qkv_b"], out_dtype=op)
            qp = qkv.data_ptr()
            ctx = o.attention(qp, qp + esz * D, qp + 2 * esz * D, 3 * D, 3 * D, B, W2V_HEADS, g.T6a, g.T6a, g.Tp, g.T6a, w2v_valid, R, D,
                              drop=self._adrop(T, f"w2v{li}.prob", self.pd["w2v_attn"]))
            if pw > 0:                                             # x = residual + dropout1(attention output)
                y1 = self._drop(T, f"w2v{li}.attn", pw, o.linear(ctx, lw["o_w"], lw["o_b"]), R, D, add=x)
            else:
                y1 = o.linear(ctx, lw["o_w"], lw["o_b"], residual=x)
            x1, x1_op = o.ln(y1, lw["ln1_g"], lw["ln1_b"], R)
            z = o.linear(x1_op, lw["fc1_w"], lw["fc1_b"])           # fp32 pre-activation: h is rounded once, as in the fused epilogue
            h = o.act(L.ACT_GELU, z, R, W2V_FFN)
            if pw > 0:                                             # x = residual + dropout3(fc2 output); activation_dropout is 0 in wav2vec2 base
                y2 = self._drop(T, f"w2v{li}.ffn", pw, o.linear(h, lw["fc2_w"], lw["fc2_b"]), R, D, add=x1)
            else:
                y2 = o.linear(h, lw["fc2_w"], lw["fc2_b"], residual=x1)
            x, x_op = o.ln(y2, lw["ln2_g"], lw["ln2_b"], R)
            t.update(qkv=qkv, ctx=ctx, y1=y1, x1_op=x1_op, z=z, h=h, y2=y2)
            T["w2v"].append(t)
        T["w2v_out"] = x
        # ---- subsampler: zero-padded operands, conv + GLU twice
        sub_in = torch.zeros(B * g.Tin1 + SLACK, D, dtype=op, device=self.dev)
        o.remap(x, g.T6a, 0, sub_in, g.Tin1, 2, B, g.Tp, D, g.Tp)
        w0, w1 = P["sub0_w"], P["sub1_w"]
        zs0 = o.gemm(sub_in, w0, o.new(B * g.T1a, w0.shape[0]), B * g.T1a, w0.shape[0], w0.shape[1], lda=2 * D,
                     a_rows=(B * g.Tin1 + SLACK) // 2, bias=P["sub0_b"])
        s0 = o.act(L.ACT_GLU, zs0, B * g.T1a, ENC_DIM, out_dtype=F32)
        sub_mid = torch.zeros(B * g.Tin2 + SLACK, ENC_DIM, dtype=op, device=self.dev)
        o.remap(s0, g.T1a, 0, sub_mid, g.Tin2, 2, B, g.T1, ENC_DIM, g.T1)
        zs1 = o.gemm(sub_mid, w1, o.new(B * g.T2a, w1.shape[0]), B * g.T2a, w1.shape[0], w1.shape[1], lda=2 * ENC_DIM,
                     a_rows=(B * g.Tin2 + SLACK) // 2, bias=P["sub1_b"])
        R2 = B * g.T2a
        x2 = o.act(L.ACT_GLU, zs1, R2, ENC_DIM, alpha=math.sqrt(ENC_DIM), out_dtype=F32)
        T.update(sub_in=sub_in, zs0=zs0, sub_mid=sub_mid, zs1=zs1)
        return self._shared_memory_fwd(g, T, x2, sub_valid)

    def _shared_memory_fwd(self, g, T, x2, sub_valid):
        """6 pre-LN shared layers -> LayerNorm -> memory stage on the rows of `x2` [B*T2a, 512] fp32 (audio: sub-sampled frames; text:
        embedded tokens): the part of the forward both branches of the encoder run (w2v2_transformer_interlingua.py:238-298)."""
        o, P, lib = self.o, self.P, self.o.lib
        B, st, op, esz = g.B, self.o.st(), self.op, self.o.esz
        R2 = B * g.T2a
        T["sub_valid"] = sub_valid
        ps, pa = self.pd["shared"], self.pd["act"]
        if ps > 0:                                                 # dropout_module after embed_scale (+ positions) (interlingua.py:237)
            x2 = self._drop(T, "embed", ps, x2, R2, ENC_DIM)
        # ---- 6 pre-LN shared layers
        D2 = ENC_DIM
        T["enc"] = []
        for li, lw in enumerate(self._enc_layers(T)):
            t = {"x_in": x2}
            _, a = o.ln(x2, lw["ln1_g"], lw["ln1_b"], R2, f32=False)
            qkv = o.linear(a, lw["qkv_w"], lw["qkv_b"], out_dtype=op)
            qp = qkv.data_ptr()
            ctx = o.attention(qp, qp + esz * D2, qp + 2 * esz * D2, 3 * D2, 3 * D2, B, ENC_HEADS, g.T2a, g.T2a, g.T2, g.T2a, sub_valid, R2, D2,
                              drop=self._adrop(T, f"enc{li}.prob", self.pd["attn"]))
            if ps > 0:
                xm = self._drop(T, f"enc{li}.attn", ps, o.linear(ctx, lw["o_w"], lw["o_b"]), R2, D2, add=x2)
            else:
                xm = o.linear(ctx, lw["o_w"], lw["o_b"], residual=x2)
            _, b_ = o.ln(xm, lw["ln2_g"], lw["ln2_b"], R2, f32=False)
            z = o.linear(b_, lw["fc1_w"], lw["fc1_b"])
            h = o.act(L.ACT_RELU, z, R2, ENC_FFN)
            if pa > 0:                                             # activation_dropout_module (transformer_layer.py); h = the fc2 operand
                h = self._drop(T, f"enc{li}.act", pa, h, R2, ENC_FFN, out_dtype=op)
            if ps > 0:
                x2 = self._drop(T, f"enc{li}.ffn", ps, o.linear(h, lw["fc2_w"], lw["fc2_b"]), R2, D2, add=xm)
            else:
                x2 = o.linear(h, lw["fc2_w"], lw["fc2_b"], residual=xm)
            t.update(a=a, qkv=qkv, ctx=ctx, xm=xm, b=b_, z=z, h=h)
            T["enc"].append(t)
        T["x2_out"] = x2
        h_enc, _ = o.ln(x2, P["ln_out_g"], P["ln_out_b"], R2)
        T["h_enc"] = h_enc
        # ---- memory stage: M learned queries over ALL T2 frames, per-layer affine LN1 of h_enc (un-folded: training form)
        Mq, RM = self.M, B * self.M
        mem = o.new(RM, D2)
        L.check(lib.cst_broadcast_rows(P["mem_embed"].data_ptr(), Mq, D2, B, mem.data_ptr(), st))
        T["mem"] = []
        for li, lw in enumerate(P["mem_layers"]):
            t = {"m_in": mem}
            _, a = o.ln(mem, lw["ln1_g"], lw["ln1_b"], RM, f32=False)
            _, kv_in = o.ln(h_enc, lw["ln1_g"], lw["ln1_b"], R2, f32=False)
            q = o.linear(a, lw["q_w"], lw["q_b"], out_dtype=op)
            kv = o.linear(kv_in, lw["kv_w"], lw["kv_b"], out_dtype=op)
            kp = kv.data_ptr()
            ctx = o.attention(q.data_ptr(), kp, kp + esz * D2, D2, 2 * D2, B, ENC_HEADS, Mq, Mq, g.T2, g.T2a, None, RM, D2,
                              drop=self._adrop(T, f"mem{li}.prob", self.pd["attn"]))
            if ps > 0:                                             # only the M memory rows of cat(h_enc, memories) survive the layer
                mm = self._drop(T, f"mem{li}.attn", ps, o.linear(ctx, lw["o_w"], lw["o_b"]), RM, D2, add=mem)
            else:
                mm = o.linear(ctx, lw["o_w"], lw["o_b"], residual=mem)
            _, b_ = o.ln(mm, lw["ln2_g"], lw["ln2_b"], RM, f32=False)
            z = o.linear(b_, lw["fc1_w"], lw["fc1_b"])
            h = o.act(L.ACT_RELU, z, RM, ENC_FFN)
            if pa > 0:
                h = self._drop(T, f"mem{li}.act", pa, h, RM, ENC_FFN, out_dtype=op)
            if ps > 0:
                mem = self._drop(T, f"mem{li}.ffn", ps, o.linear(h, lw["fc2_w"], lw["fc2_b"]), RM, D2, add=mm)
            else:
                mem = o.linear(h, lw["fc2_w"], lw["fc2_b"], residual=mm)
            t.update(a=a, kv_in=kv_in, q=q, kv=kv, ctx=ctx, mm=mm, b=b_, z=z, h=h)
            T["mem"].append(t)
        self.mem_out = mem
        return mem.view(B, Mq, D2).transpose(0, 1)                 # [M, B, 512] (view)

    # ------------------------------------------------------------------ backward
    def _ffn_bwd(self, G, name, lw, t, dy, rows, act, x_key, dx_residual=None, T=None, tag=None, p_out=0.0, p_act=0.0):
        """y = x_in + drop(fc2(drop_act(act(fc1(LNorm-ed input))))) pieces shared by all three layer types: returns d(fc1 input)
        (+ dx_residual).  t["h"] is the fc2 operand (after the activation dropout)."""
        o = self.o
        if p_out > 0:
            dy = self._drop(T, tag + ".ffn", p_out, dy, rows, dy.shape[1])
        dh, dW2, db2 = o.linear_bwd(t["h"], lw["fc2_w"], dy, rows)
        G[name + "fc2.weight"], G[name + "fc2.bias"] = dW2, db2
        if p_act > 0:
            dh = self._drop(T, tag + ".act", p_act, dh, rows, dh.shape[1])
        dz = o.act_bwd(act, t["z"], dh, rows, t["z"].shape[1])
        dxin, dW1, db1 = o.linear_bwd(t[x_key], lw["fc1_w"], dz, rows, dx_residual=dx_residual)
        G[name + "fc1.weight"], G[name + "fc1.bias"] = dW1, db1
        return dxin

    def backward(self, d_mem, G0=None):
        """d_mem: [M, B, 512] gradient of the loss w.r.t. the memories.  -> {reference parameter name: gradient}.
        G0: gradients to start from (the text pass of the same step, TextTrainPass.backward): shared parameters are summed."""
        G = {}
        for part in self.backward_iter(d_mem, G0):
            G.update(part)
        return G

    # the order in which `backward_iter` hands gradients over (the DDP bucket order, ddp.GradAllReducer)
    SEGMENTS = ("memory stage + shared layers + subsampler", "wav2vec2 layers 11..6", "wav2vec2 layers 5..0 + pos-conv + projection",
                "conv feature extractor")

    def backward_iter(self, d_mem, G0=None):
        """Generator form of `backward`: yields {name: gradient} of each finished SEGMENT in backward order, so that the caller can
        start the gradient all-reduce of a segment (NCCL, its own stream) while the kernels of the next one run -- the overlap
        `LegacyDistributedDataParallel` does not have (legacy_distributed_data_parallel.py:94-178 reduces after the whole backward)."""
        o, g, P, T = self.o, self.g, self.P, self.T
        B, Mq, D2, esz = g.B, self.M, ENC_DIM, self.o.esz
        RM, R2, R = B * Mq, B * g.T2a, B * g.T6a
        G = dict(G0) if G0 else {}
        dx2 = self._shared_memory_bwd(g, T, G, d_mem)
        self.dbg["sub_out"] = dx2
        # ---- subsampler (GLU -> conv dgrad / wgrad on the zero-padded operands)
        D = W2V_DIM
        w0, w1 = P["sub0_w"], P["sub1_w"]
        dzs1 = o.act_bwd(L.ACT_GLU, T["zs1"], dx2, R2, D2, alpha=math.sqrt(D2))
        dcol, dW1, db1 = self._conv_bwd(T["sub_mid"], w1, dzs1, R2, 2 * D2, (B * g.Tin2 + SLACK) // 2)
        G["subsample.conv_layers.1.weight"], G["subsample.conv_layers.1.bias"] = self._glu_conv_w(dW1, D2), _glu_deinterleave(db1)
        dmid = o.col2im(dcol, R2, 5, 2, D2, B * g.Tin2)
        ds0 = o.new(B * g.T1a, D2, zero=True)
        o.remap(dmid, g.Tin2, 2, ds0, g.T1a, 0, B, g.T1, D2, g.T1)
        dzs0 = o.act_bwd(L.ACT_GLU, T["zs0"], ds0, B * g.T1a, D2)
        dcol, dW0, db0 = self._conv_bwd(T["sub_in"], w0, dzs0, B * g.T1a, 2 * D, (B * g.Tin1 + SLACK) // 2)
        G["subsample.conv_layers.0.weight"], G["subsample.conv_layers.0.bias"] = self._glu_conv_w(dW0, D), _glu_deinterleave(db0)
        din = o.col2im(dcol, B * g.T1a, 5, 2, D, B * g.Tin1)
        dx = o.new(R, D, zero=True)
        o.remap(din, g.Tin1, 2, dx, g.T6a, 0, B, g.Tp, D, g.Tp)
        self.dbg["w2v_out"] = dx
        yield G
        G = {}
        # ---- 12 post-LN wav2vec2 layers
        pw = self.pd["w2v"]
        for li in reversed(range(W2V_LAYERS)):
            if li == W2V_LAYERS // 2 - 1:
                yield G
                G = {}
            lw, t, nm = P["w2v_layers"][li], T["w2v"][li], f"wav2vec_model.encoder.layers.{li}."
            if t is None:                                           # dropped by LayerDrop in this step
                continue
            dy2, dg2, dbt2 = o.ln_bwd(t["y2"], lw["ln2_g"], dx, R)
            G[nm + "final_layer_norm.weight"], G[nm + "final_layer_norm.bias"] = dg2, dbt2
            dx1 = self._ffn_bwd(G, nm, lw, t, dy2, R, L.ACT_GELU, "x1_op", dx_residual=dy2, T=T, tag=f"w2v{li}", p_out=pw)  # + residual x1 -> y2
            dy1, dg1, dbt1 = o.ln_bwd(t["y1"], lw["ln1_g"], dx1, R)
            G[nm + "self_attn_layer_norm.weight"], G[nm + "self_attn_layer_norm.bias"] = dg1, dbt1
            dctx, dWo, dbo = o.linear_bwd(t["ctx"], lw["o_w"], self._drop(T, f"w2v{li}.attn", pw, dy1, R, D) if pw > 0 else dy1, R)
            G[nm + "self_attn.out_proj.weight"], G[nm + "self_attn.out_proj.bias"] = dWo, dbo
            dqkv = o.new(R, 3 * D, zero=True)
            qp, dqp = t["qkv"].data_ptr(), dqkv.data_ptr()
            o.attention_bwd(qp, qp + esz * D, qp + 2 * esz * D, t["ctx"], dctx, dqp, dqp + 4 * D, dqp + 8 * D, 3 * D, 3 * D, D, B, W2V_HEADS,
                            g.T6a, g.T6a, g.Tp, g.T6a, T["w2v_valid"], drop=self._adrop(T, f"w2v{li}.prob", self.pd["w2v_attn"]))
            dx, dWqkv, dbqkv = o.linear_bwd(t["x_op"], lw["qkv_w"], dqkv, R, dx_residual=dy1)   # + residual x -> y1
            _attn_layer_grads(G, nm + "self_attn.", dWqkv, dbqkv, D)
        self.dbg["w2v_in"] = dx
        # ---- encoder LayerNorm, pos-conv, masked projection, feature LayerNorm
        if pw > 0:
            dx = self._drop(T, "w2v.enc", pw, dx, R, D)
        dy0, dge, dbe = o.ln_bwd(T["y0"], P["ln_enc_g"], dx, R)
        G["wav2vec_model.encoder.layer_norm.weight"], G["wav2vec_model.encoder.layer_norm.bias"] = dge, dbe
        dzpos = o.act_bwd(L.ACT_GELU, T["zpos"], dy0, R, D)
        dxp = self._posconv_bwd(G, dzpos, dy0)
        # x[padding_mask] = 0: rows t >= valid carry no gradient into the projection
        dxp_m = o.new(R, D, zero=True)
        o.remap(dxp, g.T6a, 0, dxp_m, g.T6a, 0, B, g.T6a, D, g.T6a, seg_len=T["w2v_valid"])
        self.dbg["proj_masked"] = dxp
        if self.pd["w2v_input"] > 0:
            dxp_m = self._drop(T, "w2v.input", self.pd["w2v_input"], dxp_m, R, D)
        dfl, dWp, dbp = o.linear_bwd(T["feat_ln"], P["proj_w"], dxp_m, R)
        G["wav2vec_model.post_extract_proj.weight"], G["wav2vec_model.post_extract_proj.bias"] = dWp, dbp
        dfeat, dgf, dbf = o.ln_bwd(T["c"][6], P["ln_feat_g"], dfl, R)
        G["wav2vec_model.layer_norm.weight"], G["wav2vec_model.layer_norm.bias"] = dgf, dbf
        self.dbg["conv_feats"] = dfeat
        yield G
        G = {}
        # ---- conv feature extractor; GradMultiply(feature_grad_mult) scales everything that flows into it (wav2vec2.py:530-532)
        dc = o.new(R, 512)
        o.remap(dfeat, R, 0, dc, R, 0, 1, R, 512, R, scale=self.fgm)
        fe = "wav2vec_model.feature_extractor.conv_layers."
        for i in range(6, 0, -1):
            w = P[f"conv{i}_w"]
            rows = B * g.Ta[i]
            k = w.shape[1] // 512
            dz = o.act_bwd(L.ACT_GELU, T["cz"][i], dc, rows, 512)
            dcol, dW, _ = self._conv_bwd(T["c"][i - 1], w, dz, rows, 1024, (B * g.Ta[i - 1] + SLACK) // 2, bias=False)
            G[fe + f"{i}.0.weight"] = dW.view(512, k, 512).permute(0, 2, 1).contiguous()
            dc = o.col2im(dcol, rows, k, 2, 512, B * g.Ta[i - 1])
        nch = (g.T[0] + 127) // 128
        ws = o.new(B * nch * 5120 + B * 1024 + 64 * 5120)
        dw0, dgn, dbn = o.new(512, 10), o.new(512), o.new(512)
        L.check(o.lib.cst_conv0_bwd(T["wave"].data_ptr(), B, g.L, P["conv0_w"].data_ptr(), P["gn_g"].data_ptr(), P["gn_b"].data_ptr(),
                                    T["ss"].data_ptr(), dc.data_ptr(), g.Ta[0], dw0.data_ptr(), dgn.data_ptr(), dbn.data_ptr(), ws.data_ptr(),
                                    1.0, o.st()))
        G[fe + "0.0.weight"], G[fe + "0.2.weight"], G[fe + "0.2.bias"] = dw0.view(512, 1, 10), dgn, dbn
        yield G

    def _shared_memory_bwd(self, g, T, G, d_mem):
        """Backward of `_shared_memory_fwd`: fills G with the gradients of the memory stage, the final LayerNorm and the shared layers
        (ADDING to entries that already exist: the audio and the text pass of one training step share these parameters) and returns
        d x2 [B*T2a, 512]."""
        o, P = self.o, self.P
        B, Mq, D2, esz = g.B, self.M, ENC_DIM, self.o.esz
        RM, R2 = B * Mq, B * g.T2a
        G0, G = G, {}
        ps, pa = self.pd["shared"], self.pd["act"]
        dmem = d_mem.to(self.dev, F32).transpose(0, 1).contiguous().view(RM, D2)
        dh_enc = o.new(R2, D2, zero=True)
        # ---- memory stage
        for li in reversed(range(MEM_LAYERS)):
            lw, t, nm = P["mem_layers"][li], T["mem"][li], f"interlingua_layers.{li}."
            db_in = self._ffn_bwd(G, nm, lw, t, dmem, RM, L.ACT_RELU, "b", T=T, tag=f"mem{li}", p_out=ps, p_act=pa)
            dmm, dg2, dbt2 = o.ln_bwd(t["mm"], lw["ln2_g"], db_in, RM, dx=dmem.clone())     # residual path + LN2 path
            G[nm + "final_layer_norm.weight"], G[nm + "final_layer_norm.bias"] = dg2, dbt2
            dctx, dWo, dbo = o.linear_bwd(t["ctx"], lw["o_w"], self._drop(T, f"mem{li}.attn", ps, dmm, RM, D2) if ps > 0 else dmm, RM)
            G[nm + "self_attn.out_proj.weight"], G[nm + "self_attn.out_proj.bias"] = dWo, dbo
            dq = o.new(RM, D2, zero=True)
            dkv = o.new(R2, 2 * D2, zero=True)
            kp, dkp = t["kv"].data_ptr(), dkv.data_ptr()
            o.attention_bwd(t["q"].data_ptr(), kp, kp + esz * D2, t["ctx"], dctx, dq.data_ptr(), dkp, dkp + 4 * D2, D2, 2 * D2, D2, B, ENC_HEADS,
                            Mq, Mq, g.T2, g.T2a, None, drop=self._adrop(T, f"mem{li}.prob", self.pd["attn"]))
            da, dWq, dbq = o.linear_bwd(t["a"], lw["q_w"], dq, RM)
            G[nm + "self_attn.q_proj.weight"], G[nm + "self_attn.q_proj.bias"] = dWq * 0.125, dbq * 0.125
            dkv_in, dWkv, dbkv = o.linear_bwd(t["kv_in"], lw["kv_w"], dkv, R2)
            G[nm + "self_attn.k_proj.weight"], G[nm + "self_attn.v_proj.weight"] = dWkv[:D2], dWkv[D2:]
            G[nm + "self_attn.k_proj.bias"], G[nm + "self_attn.v_proj.bias"] = dbkv[:D2], dbkv[D2:]
            dmem, dg1a, db1a = o.ln_bwd(t["m_in"], lw["ln1_g"], da, RM, dx=dmm)              # + residual
            _, dg1b, db1b = o.ln_bwd(T["h_enc"], lw["ln1_g"], dkv_in, R2, dx=dh_enc)         # the SAME LN1 on the key/value side
            G[nm + "self_attn_layer_norm.weight"], G[nm + "self_attn_layer_norm.bias"] = dg1a + dg1b, db1a + db1b
        G["interlingua_embedding.weight"] = o.colsum(dmem.view(B, Mq * D2), B, Mq * D2).view(Mq, D2)
        self.dbg = {"h_enc": dh_enc}
        # ---- final LayerNorm + shared layers
        dx2, dgo, dbo_ = o.ln_bwd(T["x2_out"], P["ln_out_g"], dh_enc, R2)
        G["layer_norm.weight"], G["layer_norm.bias"] = dgo, dbo_
        for li in reversed(range(ENC_LAYERS)):
            lw, t, nm = self._enc_layers(T)[li], T["enc"][li], self._enc_names(T)[li]
            db_in = self._ffn_bwd(G, nm, lw, t, dx2, R2, L.ACT_RELU, "b", T=T, tag=f"enc{li}", p_out=ps, p_act=pa)
            dxm, dg2, dbt2 = o.ln_bwd(t["xm"], lw["ln2_g"], db_in, R2, dx=dx2)
            G[nm + "final_layer_norm.weight"], G[nm + "final_layer_norm.bias"] = dg2, dbt2
            dctx, dWo, dbo = o.linear_bwd(t["ctx"], lw["o_w"], self._drop(T, f"enc{li}.attn", ps, dxm, R2, D2) if ps > 0 else dxm, R2)
            G[nm + "self_attn.out_proj.weight"], G[nm + "self_attn.out_proj.bias"] = dWo, dbo
            dqkv = o.new(R2, 3 * D2, zero=True)
            qp, dqp = t["qkv"].data_ptr(), dqkv.data_ptr()
            o.attention_bwd(qp, qp + esz * D2, qp + 2 * esz * D2, t["ctx"], dctx, dqp, dqp + 4 * D2, dqp + 8 * D2, 3 * D2, 3 * D2, D2, B,
                            ENC_HEADS, g.T2a, g.T2a, g.T2, g.T2a, T["sub_valid"], drop=self._adrop(T, f"enc{li}.prob", self.pd["attn"]))
            da, dWqkv, dbqkv = o.linear_bwd(t["a"], lw["qkv_w"], dqkv, R2)
            _attn_layer_grads(G, nm + "self_attn.", dWqkv, dbqkv, D2)
            dx2, dg1, db1 = o.ln_bwd(t["x_in"], lw["ln1_g"], da, R2, dx=dxm)
            G[nm + "self_attn_layer_norm.weight"], G[nm + "self_attn_layer_norm.bias"] = dg1, db1
        for k, v in G.items():
            G0[k] = v if k not in G0 else G0[k] + v
        if ps > 0:
            dx2 = self._drop(T, "embed", ps, dx2, R2, D2)
        return dx2

    def _enc_layers(self, T):
        return self.P["enc_layers"]

    def _enc_names(self, T):
        return [f"transformer_layers.{i}." for i in range(ENC_LAYERS)]

    def _conv_bwd(self, x_src, w, dz, rows, lda, a_rows, bias=True):
        """Implicit-GEMM convolution z[m] = window_m(x_src) . w: -> (dcol [rows, K] = dz w (operand dtype), dW [N, K] = dz^T windows, db)."""
        o = self.o
        N, K = w.shape
        dcol, dW, db = o.linear_bwd(x_src, w, dz, rows, bias=bias, ldx=lda, dx_dtype=o.op)
        return dcol, dW, db

    @staticmethod
    def _glu_conv_w(dW, cin):
        """[1024 interleaved (value_i, gate_i), 5*cin] -> reference Conv1d weight [1024, cin, 5]."""
        n = dW.shape[0]
        d = dW.view(n // 2, 2, 5, cin)
        d = torch.cat((d[:, 0], d[:, 1]), 0)                       # de-interleave: values then gates
        return d.permute(0, 2, 1).contiguous()

    def _posconv_bwd(self, G, dz, dy0):
        """y0 = xp + GELU(conv_g(xp) + b): -> d xp; parameter gradients of the weight-normed grouped conv (wav2vec2.py:773-786)."""
        o, g, P = self.o, self.g, self.P
        B, D, R = g.B, W2V_DIM, g.B * g.T6a
        cg, Kc = D // POS_GROUPS, POS_K * 64
        pc = "wav2vec_model.encoder.pos_conv.0."
        G[pc + "bias"] = o.colsum(dz, R, D)
        # weight gradient: per utterance, per group  dWp[g] [48, 128*64] += dz_g^T [48, T] . windows(xg)[T, 128*64]
        dWp = o.new(POS_GROUPS, cg, Kc, zero=True)
        Tpad = (g.T6a + 63) // 64 * 64
        for b in range(B):
            dzT = o.transpose(dz[b * g.T6a:], g.T6a, D)[0]          # [768, Tpad] = [16 x 48, Tpad]
            winT = o.new(POS_GROUPS, Kc, Tpad, dtype=o.op)
            for gi in range(POS_GROUPS):
                src = self.T["xg"][(b * 16 + gi) * g.Tpp:]
                L.check(o.lib.cst_transpose(src.data_ptr(), o.opc, 64, g.T6a, Kc, winT[gi].data_ptr(), o.opc, Tpad, Tpad, 0, 0, o.st()))
            o.gemm(dzT, winT, dWp, cg, Kc, Tpad, lda=Tpad, a_rows=cg, residual=dWp, nb_outer=1, nb_inner=POS_GROUPS,
                   a_bs=(0, cg * Tpad), w_bs=Kc * Tpad, c_bs=(0, cg * Kc), ldc=Kc)
        dw = dWp.view(POS_GROUPS, cg, POS_K, 64)[..., :cg].permute(0, 1, 3, 2).reshape(D, cg, POS_K)     # [co, ci, tap]
        v, gg = self.sd[pc + "weight_v"], self.sd[pc + "weight_g"]
        nrm = v.norm(dim=(0, 1), keepdim=True)
        G[pc + "weight_g"] = (dw * v / nrm).sum(dim=(0, 1), keepdim=True)
        G[pc + "weight_v"] = gg / nrm * (dw - v * (dw * v).sum(dim=(0, 1), keepdim=True) / (nrm * nrm))
        # input gradient: the same grouped implicit GEMM on the packed dz with the flipped, transposed kernel
        w = v * (gg / nrm)                                          # [768, 48, 128]
        wg = w.view(POS_GROUPS, cg, cg, POS_K)                      # [g, co, ci, tap]
        wd = torch.zeros(POS_GROUPS, cg, POS_K, 64, dtype=F32, device=self.dev)       # [g, ci, tap', co]
        wd[..., :cg] = wg.flip(3).permute(0, 2, 3, 1)
        wd = wd.reshape(POS_GROUPS, cg, Kc).to(o.op).contiguous()
        dzg = torch.zeros(B * 16 * g.Tpp + SLACK + 8, 64, dtype=o.op, device=self.dev)
        L.check(o.lib.cst_posconv_pack(dz.data_ptr(), B, g.T6a, g.T6a, dzg.data_ptr(), o.opc, g.Tpp, o.st()))
        dxp = dy0.clone()                                           # residual path
        shifted = dzg[1:]                                           # window of frame s = packed rows s+1 .. s+128
        o.gemm(shifted, wd, dxp, g.T6a, 48, Kc, lda=64, a_rows=g.Tpp - 1, residual=dxp, nb_outer=B, nb_inner=16,
               a_bs=(16 * g.Tpp * 64, g.Tpp * 64), w_bs=48 * Kc, c_bs=(g.T6a * D, 48), ldc=D)
        return dxp

    def forward_backward(self, wave, lens, d_mem):
        mem = self.forward(wave, lens)
        return mem, self.backward(d_mem)

    def refresh_weights(self):
        """After the fp32 master parameters (`self.sd`, reference layout) changed in place: rebuild the kernel-layout operands and copy
        them INTO the existing device tensors (captured CUDA graphs keep their weight pointers) -- the counterpart of the fp32 -> fp16
        parameter sync of the reference's FP16Optimizer (fairseq/optim/fp16_optimizer.py)."""
        new = _weights.prepare(self.sd, self.dev, self.op, conv_dtype=self.cdt, training=True)

        def copy_tree(dst, src):
            if torch.is_tensor(dst):
                dst.copy_(src)
            elif isinstance(dst, dict):
                for k in dst:
                    copy_tree(dst[k], src[k])
            elif isinstance(dst, (list, tuple)):
                for a, b in zip(dst, src):
                    copy_tree(a, b)
        copy_tree(self.P, new)


class TextTrainPass:
    """The text (MT) pass of the training step: integer tokens -> sqrt(512) * embedding + sinusoidal positions -> the SAME shared layers
    and memory stage as the audio pass (w2v2_transformer_interlingua.py:212-217,230-298), with the tape kept for its backward.  It shares
    the kernels, operands and master parameters of an `EncoderTrainStep`; `backward(d_mem, G)` ADDS its gradients of the shared
    parameters to the audio pass's in `G` and adds `text_embed_tokens.weight` -- what the reference's ST + MT + contrastive criterion
    (criterions/triplet_st_mt_contrastive.py) obtains from autograd over both passes."""

    def __init__(self, step, B, T):
        from .plan import TextGeometry, sinusoidal_table
        if "text_embed" not in step.P:
            raise L.CstError("the checkpoint has no text_embed_tokens.weight: the text pass is unavailable")
        self.s = step
        self.g = TextGeometry(B, T, step.M)
        self.pos = sinusoidal_table(T + 2).to(step.dev)
        self.T = None

    def forward(self, tokens, lengths):
        s, g, o = self.s, self.g, self.s.o
        B, Tn = g.B, g.L
        tokens = tokens.to(s.dev, torch.int64).contiguous()
        lengths = lengths.to(s.dev, torch.int64).contiguous()
        E = s.P["text_embed"]
        x2 = o.new(B * Tn, ENC_DIM)
        valid = torch.empty(B, dtype=torch.int32, device=s.dev)
        L.check(o.lib.cst_text_embed(tokens.data_ptr(), lengths.data_ptr(), E.data_ptr(), self.pos.data_ptr(), math.sqrt(ENC_DIM), x2.data_ptr(),
                                     valid.data_ptr(), B, Tn, Tn, ENC_DIM, E.shape[0], o.st()))
        self.T = {"tokens": tokens, "pass": 1}                     # pass 1: its dropout sites draw masks of their own
        return s._shared_memory_fwd(g, self.T, x2, valid)

    def backward(self, d_mem, G=None):
        s, g, o = self.s, self.g, self.s.o
        G = {} if G is None else G
        dx2 = s._shared_memory_bwd(g, self.T, G, d_mem)
        E = s.P["text_embed"]
        dE = o.new(E.shape[0], ENC_DIM, zero=True)
        L.check(o.lib.cst_embed_bwd(self.T["tokens"].data_ptr(), dx2.data_ptr(), math.sqrt(ENC_DIM), dE.data_ptr(), g.B, g.L, g.L, ENC_DIM,
                                    E.shape[0], 1, o.st()))
        k = "text_embed_tokens.weight"
        G[k] = dE if k not in G else G[k] + dE
        return G


class FusedAdam:
    """Adam over the encoder's fp32 master parameters, one fused kernel per tensor (`cst_adam_step`), the arithmetic of
    fairseq/optim/adam.py:157-224 (decoupled weight decay `p -= wd * lr * p`, eps added to sqrt(v), bias corrections in the step size).
    Gradients may be fp32 or the bf16 wire format of the all-reduce.  The per-step scalars (lr, step size, gradient scale) are read from
    a device buffer, so `step` can be captured into a CUDA graph once and replayed every step after `advance()`."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.0, lib=None):
        self.params = params                                       # {reference name: fp32 device tensor}, updated in place
        self.lr, self.betas, self.eps, self.wd = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        dev = next(iter(params.values())).device
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.t = 0
        self.dyn = torch.zeros(4, dtype=F32, device=dev)
        # a ring of pinned staging rows: the asynchronous copy of step t must not see the values the host wrote for step t + 1
        self._dyn_ring = torch.zeros(64, 4, dtype=F32).pin_memory() if dev.type == "cuda" else torch.zeros(64, 4, dtype=F32)
        if dev.type != "cuda" and lib is None:
            raise L.CstError("FusedAdam runs only on CUDA tensors (no CPU fallback)")
        self.dev = dev
        self.lib = lib if lib is not None else L.load()                 # injectable for the host emulator (tests/emu.py), like _Ops

    def advance(self, lr=None, grad_scale=1.0):
        """Next step number: refresh {lr, step_size, grad_scale} on the device (asynchronous copy on the current stream)."""
        self.t += 1
        if lr is not None:
            self.lr = float(lr)
        b1, b2 = self.betas
        row = self._dyn_ring[self.t % self._dyn_ring.shape[0]]
        row[0] = self.lr
        row[1] = self.lr * math.sqrt(1.0 - b2 ** self.t) / (1.0 - b1 ** self.t)
        row[2] = float(grad_scale)
        self.dyn.copy_(row, non_blocking=True)

    def step(self, grads):
        """Launch the updates (reads the scalars written by the last `advance()`); parameters without a gradient are left alone."""
        b1, b2 = self.betas
        st = L.stream_ptr() if self.dev.type == "cuda" else 0
        for name, p in self.params.items():
            g = grads.get(name)
            if g is None:
                continue
            assert g.numel() == p.numel() and g.is_contiguous() and p.is_contiguous(), name
            L.check(self.lib.cst_adam_step(p.data_ptr(), g.data_ptr(), _CODE[g.dtype], self.m[name].data_ptr(), self.v[name].data_ptr(), p.numel(),
                                           self.lr, b1, b2, self.eps, self.wd, 0.0, 1.0, self.dyn.data_ptr(), st))


class InverseSqrtLR:
    """The recipe's learning-rate schedule (--lr-scheduler inverse_sqrt --lr 1e-4 --warmup-updates 4000, train-en2any-ST.sh:48-49;
    fairseq/optim/lr_scheduler/inverse_square_root_schedule.py:49-95): linear from warmup_init_lr (default 0 when there is a warm-up) to
    `lr` over `warmup_updates` updates, then lr * sqrt(warmup_updates / n).  `at(n)` feeds `FusedAdam.advance(lr=...)`, whose device
    scalar block carries it into the captured update graph."""

    def __init__(self, lr=1e-4, warmup_updates=4000, warmup_init_lr=-1.0):
        self.end_lr, self.warmup = float(lr), int(warmup_updates)
        self.init_lr = float(warmup_init_lr) if warmup_init_lr >= 0 else (0.0 if self.warmup > 0 else self.end_lr)
        self.lr_step = (self.end_lr - self.init_lr) / self.warmup if self.warmup > 0 else 0.0
        self.decay_factor = self.end_lr * self.warmup ** 0.5
        self.initial = self.init_lr

    def at(self, num_updates):
        if num_updates < self.warmup:
            return self.init_lr + num_updates * self.lr_step
        return self.decay_factor * num_updates ** -0.5


class GraphedTrainStep:
    """One training step of the path as CUDA graphs: forward + loss in one graph, the backward pass in one graph per segment of
    `EncoderTrainStep.backward_iter`; between the segment replays the finished gradients are handed to the bucketed all-reduce
    (`ddp.GradAllReducer`), whose NCCL kernels run on the communicator's stream underneath the next segment.  Inputs are static
    device buffers (`wave`, `lens`); `loss_fn(memories [M, B, 512]) -> (loss scalar tensor, d_memories[, G0])` is captured with the
    forward pass (G0: gradients of a text pass run inside loss_fn, merged into the first backward segment).  Every activation / gradient lives in the graphs' shared memory pool, so replays allocate nothing."""

    def __init__(self, step, wave, lens, loss_fn, reducer=None, warmup=2, optimizer=None):
        self.step, self.wave, self.lens, self.reducer = step, wave, lens, reducer
        self.optimizer, self.g_update = optimizer, None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                              # eager warm-up (one-time attribute / descriptor setup)
            for _ in range(warmup):
                res = loss_fn(step.forward(wave, lens))
                step.backward(res[1], res[2] if len(res) > 2 else None)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        pool = self.pool = torch.cuda.graph_pool_handle()
        self.graphs, self.seg_grads = [], []
        g0 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g0, pool=pool):
            res = loss_fn(step.forward(wave, lens))
        self.loss, dmem = res[0], res[1]
        self.graphs.append(g0)
        it = step.backward_iter(dmem, res[2] if len(res) > 2 else None)
        while True:
            gi = torch.cuda.CUDAGraph()
            done = False
            with torch.cuda.graph(gi, pool=pool):
                try:
                    part = next(it)
                except StopIteration:
                    done = True
            if done:
                break
            self.graphs.append(gi)
            self.seg_grads.append(part)
        self.grads = {k: v for part in self.seg_grads for k, v in part.items()}
        self.names = [(k, v.numel()) for part in self.seg_grads for k, v in part.items()]
        self.reduced = None

    def run(self):
        """-> loss (device scalar).  `self.grads`: local gradients; `self.reduced`: all-reduced ones when a reducer is attached."""
        if any(p > 0 for p in self.step.pd.values()):
            self.step.next_dropout_seed()                          # the captured kernels read the seed from the device: fresh masks per replay
        self.graphs[0].replay()
        for gi, part in zip(self.graphs[1:], self.seg_grads):
            gi.replay()
            if self.reducer is not None:
                self.reducer.ready(part)
        if self.reducer is not None:
            self.reduced = self.reducer.finish(self.grads)
        if self.optimizer is not None:
            # parameter update + refresh of the kernel-layout operands as one more CUDA graph (captured on the first step, when the
            # reduced gradients exist; a persistent reducer keeps their addresses)
            grads = self.reduced if self.reducer is not None else self.grads
            self.optimizer.advance()
            if self.g_update is None:
                assert self.reducer is None or self.reducer.persistent, "a graphed optimizer needs GradAllReducer(persistent=True)"
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=self.pool):
                    self.optimizer.step(grads)
                    self.step.refresh_weights()
                self.g_update = g
            self.g_update.replay()
        return self.loss


def _glu_deinterleave(db):
    d = db.view(-1, 2)
    return torch.cat((d[:, 0], d[:, 1]), 0).contiguous()
