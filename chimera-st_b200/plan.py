"""Execution plan of the speech-encoding path for one padded batch shape (B utterances x L samples).

A plan owns every activation buffer (device memory sized once; HBM layout in DESIGN.md §3), knows the
frame geometry, and issues the kernel sequence through the C ABI on the current CUDA stream.  The
sequence has no host synchronisation (lengths/masks are evaluated on the device), so a whole forward
can be captured into one CUDA graph and replayed (`use_graph=True`): the reference's ~2000 ATen
launches + 3 host syncs per batch become one graph launch.

Stage map (reference lines in include/chimera_st_b200.h and DESIGN.md):
  lengths -> conv0+GN+GELU -> conv1..6 (implicit GEMM + GELU) -> LN -> proj (+mask) -> pos-conv (grouped
  implicit GEMM + GELU + residual) -> LN -> 12 x [QKV, attention, out-proj+res, LN, fc1+GELU, fc2+res, LN]
  -> subsampler 2 x (implicit GEMM + GLU) -> 6 x pre-LN layers -> LN -> 3 x memory cross-attention layers.
"""
import ctypes as C
import math
import os

import torch

from . import _lib as L
from .lengths import conv_out_lengths
from .synth import W2V_DIM, W2V_FFN, W2V_HEADS, ENC_DIM, ENC_FFN, ENC_HEADS, MEM_LAYERS

SLACK = 8          # zero rows appended to activation buffers read by overlapping conv windows


def _even_up(n):
    return n + (n & 1)


class Geometry:
    def __init__(self, B, Lw, M):
        self.B, self.L, self.M = B, Lw, M
        self.T = conv_out_lengths(Lw)                 # T0..T6 (true conv lengths)
        if self.T[-1] < 1:
            raise ValueError("waveform too short: L=%d" % Lw)
        self.Tp = self.T[-1]                          # T' wav2vec2 frames
        t0a = 64 * ((self.T[0] + 63) // 64)           # rows allocated per utterance at conv level 0
        self.Ta = [t0a >> i for i in range(7)]        # exact halving: level i+1 row m reads rows 2m.. of level i
        self.T6a = self.Ta[6]
        self.Tpp = self.T6a + 128                     # pos-conv operand rows per (utterance, group)
        self.T1 = (self.Tp + 1) // 2                  # subsampler conv 1 / 2 output frames
        self.T2 = (self.T1 + 1) // 2
        self.Tin1 = _even_up(self.Tp + 4)             # zero-padded subsampler inputs (2 leading rows)
        self.T1a = self.Tin1 // 2
        self.Tin2 = _even_up(self.T1 + 4)
        self.T2a = self.Tin2 // 2
        assert self.T6a >= self.Tp and self.T1a >= self.T1 and self.T2a >= self.T2


class Arena:
    """One growable device allocation shared by all plans of an encoder: only one plan runs at a time, so
    every (B, L) shape overlays the same HBM region instead of owning ~1 GB of activations each."""

    def __init__(self, device):
        self.device = device
        self.buf = torch.empty(0, dtype=torch.uint8, device=device)
        self.generation = 0

    def ensure(self, nbytes):
        if nbytes > self.buf.numel():
            self.buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=self.device)
            self.generation += 1          # views handed out before are stale: the owner drops its plans
        return self.generation


class EncoderPlan:
    def __init__(self, params, B, Lw, M, act_dtype=torch.float32, device=None, use_graph=False, lib=None,
                 arena=None, conv_dtype=None):
        # `lib` is injectable so tests can drive the plan against a host emulator of the C ABI
        # (tests/emu.py); the product never passes it and always loads the CUDA library.
        self.lib = lib if lib is not None else L.load()
        self.P = params
        self.g = g = Geometry(B, Lw, M)
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        self.act = act_dtype
        self.act_code = L.DT[act_dtype]
        self.conv_dt = conv_dtype or act_dtype          # operand dtype of the conv feature extractor (fp16 option)
        self.use_graph = use_graph
        self.graph = None
        self.launches = 0
        self.use_resident_posconv = os.environ.get("CST_POSCONV_RESIDENT", "1") != "0"
        self.use_conv0_tc = os.environ.get("CST_CONV0_TC", "1") != "0"
        self.use_stacked_posconv = os.environ.get("CST_POSCONV_STACKED", "1") != "0"
        self.arena = arena if arena is not None else Arena(self.dev)
        f32, i32, i64, u8, f64 = torch.float32, torch.int32, torch.int64, torch.uint8, torch.float64
        R, R2, RM = B * g.T6a, B * g.T2a, B * M
        spec = [
            # ---- inputs / integer side
            ("wave", B, Lw, f32), ("src_len", 1, B, i64), ("w2v_valid", 1, B, i32), ("sub_valid", 1, B, i32),
            ("w2v_len64", 1, B, i64), ("frame_mask", B, g.Tp, u8),
            # ---- conv stack (ping-pong; level i lives in cbuf{i & 1}); SLACK rows absorb the last windows
            ("scale_shift", B * 512, 2, f32), ("stats_ws", 1, B * 72, f64),
            ("cbuf0", B * g.Ta[0] + SLACK, 512, self.conv_dt), ("cbuf1", B * g.Ta[1] + SLACK, 512, self.conv_dt),
            ("feat", R, 512, f32), ("feat_ln", R, 512, act_dtype),
            # ---- wav2vec2 encoder: fp32 residual stream x, pre-LN sums y, GEMM-operand copy xa
            ("x", R, W2V_DIM, f32), ("y", R, W2V_DIM, f32), ("xa", R, W2V_DIM, act_dtype),
            ("xg", B * 16 * g.Tpp + SLACK, 64, act_dtype), ("qkv", R, 3 * W2V_DIM, act_dtype),
            ("ctx", R, W2V_DIM, act_dtype), ("ffn", R, W2V_FFN, act_dtype), ("w2v_out", R, W2V_DIM, f32),
            # ---- subsampler operands (zero-padded: re-zeroed every run)
            ("sub_in", B * g.Tin1 + SLACK, W2V_DIM, act_dtype), ("sub_mid", B * g.Tin2 + SLACK, ENC_DIM, act_dtype),
            # ---- shared encoder
            ("x2", R2, ENC_DIM, f32), ("x2a", R2, ENC_DIM, act_dtype), ("qkv2", R2, 3 * ENC_DIM, act_dtype),
            ("ctx2", R2, ENC_DIM, act_dtype), ("ffn2", R2, ENC_FFN, act_dtype), ("h_enc", R2, ENC_DIM, f32),
            # ---- memory stage
            ("kv_in", R2, ENC_DIM, act_dtype), ("kv", R2, 2 * ENC_DIM * MEM_LAYERS, act_dtype), ("mem", RM, ENC_DIM, f32),
            ("mem_a", RM, ENC_DIM, act_dtype), ("mq", RM, ENC_DIM, act_dtype), ("mctx", RM, ENC_DIM, act_dtype),
            ("mffn", RM, ENC_FFN, act_dtype),
        ]
        off, table = 0, []
        for name, rows, cols, dt in spec:
            nbytes = rows * cols * torch.empty(0, dtype=dt).element_size()
            table.append((name, off, nbytes, rows, cols, dt))
            off += (nbytes + 1023) // 1024 * 1024
        self.nbytes = off
        self.arena_generation = self.arena.ensure(off)
        for name, o, nbytes, rows, cols, dt in table:
            t = self.arena.buf[o:o + nbytes].view(dt)
            setattr(self, name, t.view(rows, cols) if rows > 1 or name in ("wave", "frame_mask") else t.view(cols))
        self.cbuf = [self.cbuf0, self.cbuf1]
        self.arena.buf[:off].zero_()

    # ------------------------------------------------------------------ launch helpers
    def _gemm(self, A, W, C_, M, N, K, lda, a_rows, bias=None, residual=None, act=L.ACT_NONE, alpha=1.0,
              rows_per_seg=None, seg_rows_valid=None, out_rows_per_seg=None, out_row_off=0, seg_len=None,
              ldc=None, nb_outer=1, nb_inner=1, a_bs=(0, 0), w_bs=0, c_bs=(0, 0), bias_bs=0):
        p = L.GemmParams()
        p.A, p.W, p.bias, p.residual, p.C = A.data_ptr(), W.data_ptr(), L.ptr(bias), L.ptr(residual), C_.data_ptr()
        p.ab_dtype, p.c_dtype = L.DT[A.dtype], L.DT[C_.dtype]
        assert A.dtype == W.dtype
        p.M, p.N, p.K = M, N, K
        p.lda = lda
        p.ldc = ldc if ldc is not None else C_.shape[1]
        p.ldr = residual.shape[1] if residual is not None else 0
        p.a_rows = a_rows
        p.act, p.alpha = act, alpha
        p.nb_outer, p.nb_inner = nb_outer, nb_inner
        p.a_bs_outer, p.a_bs_inner = a_bs
        p.w_bs_inner = w_bs
        p.c_bs_outer, p.c_bs_inner = c_bs
        p.r_bs_outer, p.r_bs_inner = c_bs
        p.bias_bs_inner = bias_bs
        p.rows_per_seg = rows_per_seg if rows_per_seg is not None else M
        p.seg_rows_valid = seg_rows_valid if seg_rows_valid is not None else p.rows_per_seg
        p.out_rows_per_seg = out_rows_per_seg if out_rows_per_seg is not None else p.rows_per_seg
        p.out_row_off = out_row_off
        p.seg_len = L.ptr(seg_len)
        p.segs_per_outer = 1
        L.check(self.lib.cst_gemm(C.byref(p), self.st))
        self.launches += 1

    def _linear(self, A, W, b, C_, rows, act=L.ACT_NONE, residual=None, alpha=1.0):
        """C[rows, N] = act(A[rows, K] W^T + b) * alpha (+ residual); dense row-major operands."""
        N, K = W.shape
        self._gemm(A, W, C_, rows, N, K, lda=A.shape[1], a_rows=A.shape[0], bias=b, residual=residual, act=act, alpha=alpha)

    def _ln(self, x, gb, rows, out_f32=None, out_lp=None, rows_per_seg=None, seg_rows_valid=None,
            out_rows_per_seg=None, out_row_off=0, zero_invalid=0):
        Cdim = x.shape[1]
        rps = rows_per_seg if rows_per_seg is not None else rows
        lp_dt = L.DT[out_lp.dtype] if out_lp is not None else L.F32
        L.check(self.lib.cst_layernorm(
            x.data_ptr(), Cdim, gb[0].data_ptr(), gb[1].data_ptr(), L.ptr(out_f32), L.ptr(out_lp), lp_dt, Cdim,
            rows, Cdim, rps, seg_rows_valid if seg_rows_valid is not None else rps,
            out_rows_per_seg if out_rows_per_seg is not None else rps, out_row_off, zero_invalid, self.st))
        self.launches += 1

    def _attn(self, q, k, v, out, ldq, ldkv, H, n_q, q_rps, n_kv, kv_rps, kv_len):
        L.check(self.lib.cst_attention(q, k, v, out.data_ptr(), L.DT[out.dtype], ldq, ldkv, out.shape[1],
                                       self.g.B, H, n_q, q_rps, n_kv, kv_rps, L.ptr(kv_len), self.st))
        self.launches += 1

    # ------------------------------------------------------------------ stages
    def _stage_frontend(self):
        g, P, lib = self.g, self.P, self.lib
        B = g.B
        L.check(lib.cst_frame_lengths(self.src_len.data_ptr(), B, g.L, g.Tp, self.w2v_valid.data_ptr(),
                                      self.sub_valid.data_ptr(), self.w2v_len64.data_ptr(),
                                      self.frame_mask.data_ptr(), self.st))
        L.check(lib.cst_conv0_stats(self.wave.data_ptr(), B, g.L, P["conv0_w"].data_ptr(), P["gn_g"].data_ptr(),
                                    P["gn_b"].data_ptr(), self.scale_shift.data_ptr(), self.stats_ws.data_ptr(), self.st))
        if self.conv_dt != torch.float32 and self.use_conv0_tc:
            L.check(lib.cst_conv0_apply_tc(self.wave.data_ptr(), B, g.L, P["conv0_w16"].data_ptr(), self.scale_shift.data_ptr(),
                                           self.cbuf[0].data_ptr(), L.DT[self.conv_dt], g.Ta[0], self.st))
        else:
            L.check(lib.cst_conv0_apply(self.wave.data_ptr(), B, g.L, P["conv0_w"].data_ptr(), self.scale_shift.data_ptr(),
                                        self.cbuf[0].data_ptr(), L.DT[self.conv_dt], g.Ta[0], self.st))
        self.launches += 4
        for i in range(1, 7):
            src = self.cbuf[(i - 1) & 1]
            dst = self.feat if i == 6 else self.cbuf[i & 1]
            w = P[f"conv{i}_w"]
            rows = B * g.Ta[i]
            # stride-2 conv over channels-last rows: window of output row m = K contiguous elements at 1024*m
            self._gemm(src, w, dst, rows, 512, w.shape[1], lda=1024, a_rows=(B * g.Ta[i - 1] + SLACK) // 2,
                       act=L.ACT_GELU, rows_per_seg=g.Ta[i], ldc=512)
        R = B * g.T6a
        self._ln(self.feat, (P["ln_feat_g"], P["ln_feat_b"]), R, out_lp=self.feat_ln)
        # post_extract_proj + x[padding_mask] = 0 (rows t >= valid_b, incl. the allocation tail t >= T')
        self._gemm(self.feat_ln, P["proj_w"], self.x, R, W2V_DIM, 512, lda=512, a_rows=R, bias=P["proj_b"],
                   rows_per_seg=g.T6a, seg_len=self.w2v_valid)
        L.check(lib.cst_posconv_pack(self.x.data_ptr(), B, g.T6a, g.Tp, self.xg.data_ptr(), self.act_code, g.Tpp, self.st))
        self.launches += 1
        # grouped pos-conv: z = (utterance, group); window of frame t = rows t..t+127 of the packed operand
        if self.act == torch.bfloat16 and self.use_resident_posconv and self.use_stacked_posconv:
            L.check(lib.cst_posconv_stacked(self.xg.data_ptr(), P["pos_w2"].data_ptr(), P["pos_b"].data_ptr(), self.x.data_ptr(),
                                            self.y.data_ptr(), B, g.T6a, g.T6a, g.Tpp, self.st))
            self.launches += 1
        elif self.act == torch.bfloat16 and self.use_resident_posconv:
            L.check(lib.cst_posconv(self.xg.data_ptr(), P["pos_w"].data_ptr(), P["pos_b"].data_ptr(), self.x.data_ptr(),
                                    self.y.data_ptr(), B, g.T6a, g.T6a, g.Tpp, self.st))
            self.launches += 1
        else:
            self._gemm(self.xg, P["pos_w"], self.y, g.T6a, 48, 128 * 64, lda=64, a_rows=g.Tpp, bias=P["pos_b"],
                       residual=self.x, act=L.ACT_GELU, ldc=W2V_DIM, nb_outer=B, nb_inner=16,
                       a_bs=(16 * g.Tpp * 64, g.Tpp * 64), w_bs=48 * 128 * 64, c_bs=(g.T6a * W2V_DIM, 48), bias_bs=48)
        self._ln(self.y, (P["ln_enc_g"], P["ln_enc_b"]), R, out_f32=self.x, out_lp=self.xa)

    def _stage_w2v_layers(self):
        g, P = self.g, self.P
        R = g.B * g.T6a
        D = W2V_DIM
        # zero-padded subsampler operands share the arena with other shapes: clear them every run
        self.sub_in.zero_()
        self.sub_mid.zero_()
        self.launches += 2
        es = self.qkv.element_size()
        for i, lw in enumerate(P["w2v_layers"]):
            self._linear(self.xa, lw["qkv_w"], lw["qkv_b"], self.qkv, R)
            qp = self.qkv.data_ptr()
            # all allocated query rows are computed (rows >= T' are finite filler, never read as keys):
            # no buffer row is ever left stale, so masked keys always meet finite V rows
            self._attn(qp, qp + D * es, qp + 2 * D * es, self.ctx, 3 * D, 3 * D, W2V_HEADS,
                       g.T6a, g.T6a, g.Tp, g.T6a, self.w2v_valid)
            self._linear(self.ctx, lw["o_w"], lw["o_b"], self.y, R, residual=self.x)
            self._ln(self.y, (lw["ln1_g"], lw["ln1_b"]), R, out_f32=self.x, out_lp=self.xa)
            self._linear(self.xa, lw["fc1_w"], lw["fc1_b"], self.ffn, R, act=L.ACT_GELU)
            self._linear(self.ffn, lw["fc2_w"], lw["fc2_b"], self.y, R, residual=self.x)
            if i + 1 < len(P["w2v_layers"]):
                self._ln(self.y, (lw["ln2_g"], lw["ln2_b"]), R, out_f32=self.x, out_lp=self.xa)
            else:
                # last layer: the LN output is (a) the wav2vec2 feature [B,T',768] and (b) the subsampler's
                # zero-padded operand (2 leading zero frames, zeros from frame T' on)
                self._ln(self.y, (lw["ln2_g"], lw["ln2_b"]), R, out_f32=self.w2v_out)
                self._ln(self.y, (lw["ln2_g"], lw["ln2_b"]), R, out_lp=self.sub_in, rows_per_seg=g.T6a,
                         seg_rows_valid=g.Tp, out_rows_per_seg=g.Tin1, out_row_off=2, zero_invalid=0)

    def _stage_subsample(self):
        g, P = self.g, self.P
        B = g.B
        w0, w1 = P["sub0_w"], P["sub1_w"]
        # Conv1d(k5,s2,p2)+GLU twice; operands are zero-padded so the window of frame t1 starts at padded row 2*t1
        self._gemm(self.sub_in, w0, self.sub_mid, B * g.T1a, w0.shape[0], w0.shape[1], lda=2 * W2V_DIM,
                   a_rows=(B * g.Tin1 + SLACK) // 2, bias=P["sub0_b"], act=L.ACT_GLU, rows_per_seg=g.T1a,
                   seg_rows_valid=g.T1, out_rows_per_seg=g.Tin2, out_row_off=2, ldc=ENC_DIM)
        self._gemm(self.sub_mid, w1, self.x2, B * g.T2a, w1.shape[0], w1.shape[1], lda=2 * ENC_DIM,
                   a_rows=(B * g.Tin2 + SLACK) // 2, bias=P["sub1_b"], act=L.ACT_GLU, alpha=math.sqrt(ENC_DIM),
                   rows_per_seg=g.T2a, out_rows_per_seg=g.T2a, ldc=ENC_DIM)

    def _stage_shared_layers(self):
        g, P = self.g, self.P
        R2 = g.B * g.T2a
        D = ENC_DIM
        es = self.qkv2.element_size()
        for lw in P["enc_layers"]:
            self._ln(self.x2, (lw["ln1_g"], lw["ln1_b"]), R2, out_lp=self.x2a)
            self._linear(self.x2a, lw["qkv_w"], lw["qkv_b"], self.qkv2, R2)
            qp = self.qkv2.data_ptr()
            self._attn(qp, qp + D * es, qp + 2 * D * es, self.ctx2, 3 * D, 3 * D, ENC_HEADS,
                       g.T2a, g.T2a, g.T2, g.T2a, self.sub_valid)
            self._linear(self.ctx2, lw["o_w"], lw["o_b"], self.x2, R2, residual=self.x2)
            self._ln(self.x2, (lw["ln2_g"], lw["ln2_b"]), R2, out_lp=self.x2a)
            self._linear(self.x2a, lw["fc1_w"], lw["fc1_b"], self.ffn2, R2, act=L.ACT_RELU)
            self._linear(self.ffn2, lw["fc2_w"], lw["fc2_b"], self.x2, R2, residual=self.x2)
        self._ln(self.x2, (P["ln_out_g"], P["ln_out_b"]), R2, out_f32=self.h_enc)

    def _stage_memory(self):
        g, P = self.g, self.P
        B, M = g.B, g.M
        R2, RM, D = B * g.T2a, B * M, ENC_DIM
        es = self.kv.element_size()
        L.check(self.lib.cst_broadcast_rows(P["mem_embed"].data_ptr(), M, D, B, self.mem.data_ptr(), self.st))
        self.launches += 1
        # K/V side of all memory layers at once: unit LayerNorm of h_enc, then one GEMM whose weights carry each layer's
        # LN1 affine (weights.py) -> kv[:, i*1024 : (i+1)*1024] = [K_i | V_i]
        nl = len(P["mem_layers"])
        self._ln(self.h_enc, (P["unit_g"], P["unit_b"]), R2, out_lp=self.kv_in)
        self._linear(self.kv_in, P["mem_kv_w"], P["mem_kv_b"], self.kv, R2)
        for i, lw in enumerate(P["mem_layers"]):
            ln1 = (lw["ln1_g"], lw["ln1_b"])
            self._ln(self.mem, ln1, RM, out_lp=self.mem_a)              # the layer's pre-LN on the M queries
            self._linear(self.mem_a, lw["q_w"], lw["q_b"], self.mq, RM)
            kp = self.kv.data_ptr() + i * 2 * D * es
            # memories attend ALL T2 frames: the reference passes an all-False key-padding mask here
            self._attn(self.mq.data_ptr(), kp, kp + D * es, self.mctx, D, 2 * D * nl, ENC_HEADS, M, M, g.T2, g.T2a, None)
            self._linear(self.mctx, lw["o_w"], lw["o_b"], self.mem, RM, residual=self.mem)
            self._ln(self.mem, (lw["ln2_g"], lw["ln2_b"]), RM, out_lp=self.mem_a)
            self._linear(self.mem_a, lw["fc1_w"], lw["fc1_b"], self.mffn, RM, act=L.ACT_RELU)
            self._linear(self.mffn, lw["fc2_w"], lw["fc2_b"], self.mem, RM, residual=self.mem)

    def _issue(self, upto="memory"):
        self.st = L.stream_ptr() if self.dev.type == "cuda" else 0
        self.launches = 0
        self._stage_frontend()
        if upto == "frontend":
            return
        self._stage_w2v_layers()
        if upto == "w2v":
            return
        self._stage_subsample()
        self._stage_shared_layers()
        self._stage_memory()

    # ------------------------------------------------------------------ public
    def load_inputs(self, wave, src_lengths):
        """Copy one padded batch into the plan's static input buffers (async on the current stream)."""
        assert tuple(wave.shape) == (self.g.B, self.g.L), (tuple(wave.shape), (self.g.B, self.g.L))
        self.wave.copy_(wave, non_blocking=True)
        self.src_len.copy_(src_lengths, non_blocking=True)

    def run(self, upto="memory", eager=False):
        """Launch the forward pass for the loaded inputs; returns the number of kernel launches issued."""
        if self.arena_generation != self.arena.generation:
            raise RuntimeError("stale plan: the arena was re-allocated after this plan was built")
        if self.use_graph and upto == "memory" and not eager:
            if self.graph is None:
                self._issue()                                   # warm-up: one-time attribute / descriptor setup
                torch.cuda.current_stream().synchronize()
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._issue()
                self.graph_launches = self.launches
            self.graph.replay()
            return self.graph_launches
        self._issue(upto)
        return self.launches

    # views of the results (valid until the next run)
    def memories(self):
        """[M, B, 512] fp32, the reference's `encoder_out` layout (time-major)."""
        return self.mem.view(self.g.B, self.g.M, ENC_DIM).transpose(0, 1)

    def view(self, name):
        g = self.g
        if name == "conv_feats":     # [B, 512, T'] like ConvFeatureExtractionModel's output
            return self.feat.view(g.B, g.T6a, 512)[:, :g.Tp].transpose(1, 2)
        if name == "w2v_in":
            return None
        if name == "w2v_out":
            return self.w2v_out.view(g.B, g.T6a, W2V_DIM)[:, :g.Tp]
        if name == "h_enc":
            return self.h_enc.view(g.B, g.T2a, ENC_DIM)[:, :g.T2]
        if name == "frame_mask":
            return self.frame_mask.bool()
        raise KeyError(name)


class TextGeometry:
    """Shared-stage geometry of a text batch: T tokens per utterance, no subsampling (T2 = T2a = T)."""

    def __init__(self, B, T, M):
        if T < 1:
            raise ValueError("empty token batch")
        self.B, self.L, self.M = B, T, M
        self.T2 = self.T2a = T


def sinusoidal_table(n, dim=ENC_DIM, padding_idx=1):
    """SinusoidalPositionalEmbedding.get_embedding (fairseq/modules/sinusoidal_positional_embedding.py:38-59), computed
    with the reference's own float32 torch expressions."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float) * -e)
    e = torch.arange(n, dtype=torch.float).unsqueeze(1) * e.unsqueeze(0)
    t = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
    t[padding_idx] = 0
    return t.contiguous()


class TextPlan(EncoderPlan):
    """Text (MT) branch of the encoder (w2v2_transformer_interlingua.py:212-217,230-236): `cst_text_embed` feeds the
    SAME shared-layer and memory stages as the audio branch (`_stage_shared_layers`, `_stage_memory`)."""

    def __init__(self, params, B, T, M, act_dtype=torch.float32, device=None, use_graph=False, lib=None, arena=None):
        if "text_embed" not in params:
            raise RuntimeError("the checkpoint has no text_embed_tokens.weight: the text branch is unavailable")
        self.lib = lib if lib is not None else L.load()
        self.P = params
        self.g = g = TextGeometry(B, T, M)
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        self.act = act_dtype
        self.act_code = L.DT[act_dtype]
        self.use_graph = use_graph
        self.graph = None
        self.launches = 0
        self.arena = arena if arena is not None else Arena(self.dev)
        f32, i32, i64 = torch.float32, torch.int32, torch.int64
        R2, RM = B * g.T2a, B * M
        spec = [
            ("tokens", B, T, i64), ("src_len", 1, B, i64), ("sub_valid", 1, B, i32), ("pos_table", T + 2, ENC_DIM, f32),
            ("x2", R2, ENC_DIM, f32), ("x2a", R2, ENC_DIM, act_dtype), ("qkv2", R2, 3 * ENC_DIM, act_dtype),
            ("ctx2", R2, ENC_DIM, act_dtype), ("ffn2", R2, ENC_FFN, act_dtype), ("h_enc", R2, ENC_DIM, f32),
            ("kv_in", R2, ENC_DIM, act_dtype), ("kv", R2, 2 * ENC_DIM * MEM_LAYERS, act_dtype), ("mem", RM, ENC_DIM, f32),
            ("mem_a", RM, ENC_DIM, act_dtype), ("mq", RM, ENC_DIM, act_dtype), ("mctx", RM, ENC_DIM, act_dtype),
            ("mffn", RM, ENC_FFN, act_dtype),
        ]
        off, table = 0, []
        for name, rows, cols, dt in spec:
            nbytes = rows * cols * torch.empty(0, dtype=dt).element_size()
            table.append((name, off, nbytes, rows, cols, dt))
            off += (nbytes + 1023) // 1024 * 1024
        self.nbytes = off
        self.arena_generation = self.arena.ensure(off)
        for name, o, nbytes, rows, cols, dt in table:
            t = self.arena.buf[o:o + nbytes].view(dt)
            setattr(self, name, t.view(cols) if name in ("src_len", "sub_valid") else t.view(rows, cols))
        self.arena.buf[:off].zero_()
        self._pos_host = sinusoidal_table(T + 2)

    def load_inputs(self, tokens, src_lengths):
        assert tuple(tokens.shape) == (self.g.B, self.g.L), (tuple(tokens.shape), (self.g.B, self.g.L))
        self.tokens.copy_(tokens, non_blocking=True)
        self.src_len.copy_(src_lengths, non_blocking=True)
        self.pos_table.copy_(self._pos_host, non_blocking=True)      # the arena is shared with other shapes

    def _issue(self, upto="memory"):
        g, P = self.g, self.P
        self.st = L.stream_ptr() if self.dev.type == "cuda" else 0
        self.launches = 0
        E = P["text_embed"]
        L.check(self.lib.cst_text_embed(self.tokens.data_ptr(), self.src_len.data_ptr(), E.data_ptr(), self.pos_table.data_ptr(),
                                        math.sqrt(ENC_DIM), self.x2.data_ptr(), self.sub_valid.data_ptr(), g.B, g.L, g.T2a,
                                        ENC_DIM, E.shape[0], self.st))
        self.launches += 1
        self._stage_shared_layers()
        self._stage_memory()

    def view(self, name):
        if name == "h_enc":
            return self.h_enc.view(self.g.B, self.g.T2a, ENC_DIM)[:, :self.g.T2]
        raise KeyError(name)
