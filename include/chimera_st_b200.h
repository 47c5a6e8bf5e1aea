/* chimera_st_b200.h -- C ABI of the B200-native (sm_100a) Chimera-ST speech-encoding path.
 *
 * The reference (Glaciohound/Chimera-ST, a fairseq fork) has NO native seam on this path: every
 * stage is a Python module calling torch ops.  Each entry point below therefore names the
 * reference *Python* interface whose arithmetic it replaces (paths under the reference tree);
 * INTEGRATION.md shows the ctypes binding and the fairseq `--user-dir` plugin that calls them.
 *
 * Conventions
 *   - plain pointers + sizes only; all data pointers are DEVICE pointers unless named *_host;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - nothing allocates, nothing synchronises the stream; workspaces are caller-owned;
 *   - every function returns 0 on success, else a CST_ERR_* code; cst_last_error() (thread-local)
 *     gives the message.  There is no CPU fallback anywhere: no device => error.
 *   - activations are row-major [rows, channels] ("channels-last"); a batch of B utterances is
 *     B segments of `rows_per_seg` rows.  dtypes: CST_F32, CST_BF16 (and CST_F16 for cst_gemm / cst_conv0_apply).
 */
#ifndef CHIMERA_ST_B200_H_
#define CHIMERA_ST_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CST_ABI_VERSION 1

enum { CST_OK = 0, CST_ERR_ARG = 1, CST_ERR_CUDA = 2, CST_ERR_UNSUPPORTED = 3 };
enum { CST_F32 = 0, CST_BF16 = 1, CST_F16 = 2 };   /* F16: conv feature extractor operands of the 16-bit mode */
enum { CST_ACT_NONE = 0, CST_ACT_GELU = 1, CST_ACT_RELU = 2, CST_ACT_GLU = 3 };

int cst_abi_version(void);
const char* cst_last_error(void);
/* Fills name (<=255 chars), SM count, cc major/minor of the current device. */
int cst_device_info(char* name, int name_cap, int* sm_count, int* cc_major, int* cc_minor);

/* ---- a4: frame-level padding rule + lengths --------------------------------------------------
 * Replaces: lengths_to_padding_mask (fairseq/data/data_utils.py:491-495), the trim/view/all frame
 * mask of Wav2Vec2Model.forward (fairseq/models/wav2vec/wav2vec2.py:543-548), output_length of
 * _get_w2v_feature (fairseq/models/chimera/w2v2_transformer.py:327-333) and
 * Conv1dSubsampler.get_out_seq_lens_tensor (fairseq/models/speech_to_text/s2t_transformer.py:63-67).
 * src_len [B] int64 (device); L = padded sample width; n_frames = T'.  Precondition: L == max(src_len) (the collater
 * pads to the longest utterance, speech_to_text_dataset.py:218; the reference takes the mask width from max(src_lengths)
 * and is undefined for an over-padded batch).
 * Outputs (device; any may be NULL): w2v_valid [B] int32 = min(T', ceil(len/(L/T'))),
 * sub_valid [B] int32 = subsampled twice, w2v_len64 [B] int64, frame_mask [B*T'] uint8 (1 = padded). */
int cst_frame_lengths(const int64_t* src_len, int B, int L, int n_frames,
                      int32_t* w2v_valid, int32_t* sub_valid, int64_t* w2v_len64, uint8_t* frame_mask,
                      void* stream);

/* ---- f.3 input pipeline: 16-bit PCM on the wire ------------------------------------------------------------------------
 * Replaces: the host-side int16 -> float32 normalisation of get_waveform (fairseq/data/audio/audio_utils.py:33-55, soundfile
 * dtype="float32": sample / 32768) followed by an fp32 host->device copy (fairseq/utils.py move_to_cuda).  The padded batch is
 * sent as int16 (half the PCIe bytes) and scaled on the device: out[i] = in[i] * 2^-15, bit-identical (exact power of two).
 * in / out: n contiguous samples, both 16-byte aligned. */
int cst_wave_i16_to_f32(const int16_t* in, float* out, long long n, void* stream);

/* ---- a1: conv0 (1->512, k10, s5, no bias) + GroupNorm(512 groups, padded-time statistics) + GELU ----
 * Replaces: ConvFeatureExtractionModel block 0 (wav2vec2.py:697-734,755-763; Fp32GroupNorm
 * fp32_group_norm.py:17-25).  Two launches:
 *   stats : per (b,c) mean / rstd over ALL T0 padded frames, derived from the 10x10 lag-correlation
 *           of the waveform (reads the waveform once; no [B,512,T0] intermediate) -> scale_shift
 *           [B,512] float2 {gamma*rstd, beta - mean*gamma*rstd};  stats_ws: B*72 doubles.
 *   apply : recompute conv, normalise, GELU, store channels-last out[b, t, c] for t < T0, zeros for
 *           T0 <= t < rows_per_seg.  wave [B,L] f32; w [512,10] f32. */
int cst_conv0_stats(const float* wave, int B, int L, const float* w, const float* gamma, const float* beta,
                    float* scale_shift, double* stats_ws, void* stream);
int cst_conv0_apply(const float* wave, int B, int L, const float* w, const float* scale_shift,
                    void* out, int out_dtype, int rows_per_seg, void* stream);

/* ---- a1 (16-bit mode): the same stage with the convolution on tcgen05 (3-term fp16 split: fp32-level accuracy)
 * Replaces: as cst_conv0_apply, for out_dtype CST_F16 / CST_BF16.  w16: fp16 [512, 64], row c =
 * [ hi(w[c,0..9]) | hi(w[c,0..9]) | lo(w[c,0..9]) | 0 ... ] with hi = fp16(w), lo = fp16(w - hi); the kernel builds
 * the matching activation rows [ hi(x) | lo(x) | hi(x) | 0 0 ] so that x.w ~= xh.wh + xl.wh + xh.wl.  The epilogue
 * (scale/shift, exact-erf GELU, pack) runs in the accumulator layout and is written by bulk tensor stores. */
int cst_conv0_apply_tc(const float* wave, int B, int L, const void* w16, const float* scale_shift,
                       void* out, int out_dtype, int rows_per_seg, void* stream);

/* ---- GEMM with fused epilogue: the conv stack (implicit GEMM), every projection and FFN ------------
 * Replaces: nn.Conv1d of blocks 1-6 (wav2vec2.py:707,733-734), post_extract_proj (wav2vec2.py:550-551),
 * the grouped pos_conv (wav2vec2.py:773-786,823-825), F.multi_head_attention_forward's in/out
 * projections (fairseq/modules/multihead_attention.py:165-187), fc1/fc2 (wav2vec2.py:948-957,
 * fairseq/modules/transformer_layer.py:148-154) and Conv1dSubsampler's convs+GLU
 * (s2t_transformer.py:69-77).
 *
 *   acc[m,n] = sum_k A[zoff_a + m*lda + k] * W[zoff_w + n*K + k]          (lda < K: overlapping conv windows)
 *   v = act(acc + bias[n]) * alpha + residual[row,n]     (GLU: v = a * sigmoid(g) on column pairs 2i,2i+1)
 *   C[zoff_c + row*ldc + n] = v          row = seg*out_rows_per_seg + t + out_row_off, (seg,t) = divmod(m, rows_per_seg)
 *   rows with t >= seg_rows_valid are not stored; with seg_len != NULL rows with t >= seg_len[seg]
 *   (seg counted across batch z_outer) are stored as 0 (x[padding_mask] = 0, wav2vec2.py:820-821).
 * Batched over z = zo*nb_inner + zi (pos-conv: zo = utterance, zi = group).
 * CST_F32 inputs use an FFMA kernel (true-fp32 accumulate: the 1e-5 parity mode); CST_BF16 inputs use the
 * TMA + tcgen05/TMEM kernel (fp32 accumulate in tensor memory); CST_F16 inputs use the same kernel with fp16 operand
 * formats (the normalisation-free conv stack of the 16-bit mode: 3 more mantissa bits at the same bytes).
 * A and W must share a dtype. */
typedef struct cst_gemm_params {
  const void* A; const void* W; const float* bias; const float* residual; void* C;
  int ab_dtype; int c_dtype;
  int M, N, K;
  long long lda, ldc, ldr;
  long long a_rows;              /* rows addressable from A (per batch z) without leaving the buffer */
  int act; float alpha;
  int nb_outer, nb_inner;
  long long a_bs_outer, a_bs_inner, w_bs_inner, c_bs_outer, c_bs_inner, r_bs_outer, r_bs_inner;
  int bias_bs_inner;
  int rows_per_seg, seg_rows_valid; long long out_rows_per_seg; int out_row_off;
  const int32_t* seg_len; int segs_per_outer;
  /* ---- LayerNorm fused AROUND the GEMM (tensor-core path, identity row mapping, unbatched; all NULL / 0 = plain GEMM).
   * Replaces the separate LayerNorm passes of the post-LN wav2vec2 layers (wav2vec2.py:937-957) and of the pre-LN shared
   * layers (transformer_layer.py:129-155).  Row statistics travel as PARTIAL sums: stats[row*8 + slot] = {sum, sum of
   * squares} of one 128-column slice of an fp32 row (slot = 2 * n-tile + column half), written by the GEMM that produces
   * the row (out_stats) and added up by the GEMMs that consume it; mean / rstd (eps 1e-5) are formed on the fly.
   *   ln_in_stats : the A operand is the bf16 copy of UN-normalised rows; the epilogue applies their LayerNorm after the
   *                 product: v = rstd*(acc - mean*ln_colsum[n]) + bias[n], with gamma folded into W and W.beta into bias
   *                 by the caller (ln_colsum[n] = sum_k W'[n,k]).  Bulk-store epilogue only (no residual).
   *   res_stats   : the residual rows are UN-normalised; the epilogue adds LayerNorm(residual)*res_gamma + res_beta.
 *                 res_slots == 0: res_stats holds one finished {rstd, -mean*rstd} pair per row (cst_layernorm_ab).
   *   C2          : second copy of the output in c2_dtype (the next GEMM's bf16 operand), row stride ldc2.
   *   out_stats   : partial statistics of the fp32 output rows.  res_stats / C2 / out_stats need an fp32 C with a
   *                 residual (which may be C itself: in-place update of the residual stream). */
  const float* ln_in_stats; const float* ln_colsum; int ln_in_slots;
  const float* res_stats; int res_slots; const float* res_gamma; const float* res_beta;
  void* C2; int c2_dtype; long long ldc2;
  float* out_stats;
  int ln_dim;                    /* width of the normalised rows (K for ln_in_stats, N for res_stats / out_stats) */
  int exact_act;                 /* tensor-core path: 1 = erff GELU / exact sigmoid in the epilogue (the fp32 parity mode's split GEMMs) */
  float acc_scale;               /* tensor-core path: the accumulator is multiplied by this before the bias (0 = 1): undoes the
                                    power-of-two scale that keeps the lo parts of split fp16 weights out of the denormals */
} cst_gemm_params;
int cst_gemm(const cst_gemm_params* p, void* stream);

/* fp32 mode on the tensor cores: 3-term fp16 split of a GEMM operand.  x [rows, C] f32 (row stride ldx) -> out fp16 [rows, 3C] =
 * [hi | lo | hi] per row, hi = fp16(x), lo = fp16(x - hi).  With weight rows packed [whi | whi | wlo] per C-wide block, cst_gemm
 * (CST_F16 operands, K' = 3K, fp32 accumulation, exact_act = 1) gives x.w to ~2^-22 relative: the <= 1e-5 parity mode without the
 * CUDA-core FFMA GEMM.  Same reference lines as cst_gemm. */
int cst_split_f16(const float* x, long long ldx, long long rows, int C, void* out, void* stream);

/* ---- LayerNorm over the channel axis (eps 1e-5), one warp per row ----------------------------------
 * Replaces: LayerNorm of wav2vec2.py:539-540, 827-828, 945/957, transformer_layer.py:129-155,
 * w2v2_transformer_interlingua.py:254-255.  x f32 [rows, C] (C in {512,768}); writes out_f32 and/or
 * out_lp (dtype lp_dtype) -- the fp32 residual stream and the GEMM operand copy.  Row remap as in
 * cst_gemm; rows with t >= seg_rows_valid are written as zeros when zero_invalid != 0. */
int cst_layernorm(const float* x, long long ldx, const float* gamma, const float* beta,
                  float* out_f32, void* out_lp, int lp_dtype, long long ldo,
                  int rows, int C, int rows_per_seg, int seg_rows_valid,
                  long long out_rows_per_seg, int out_row_off, int zero_invalid, void* stream);

/* "Light" LayerNorm of the post-LN wav2vec2 layers (16-bit mode): writes only the bf16 operand copy of the normalised rows and
 * ab_out[row] = {rstd, -mean*rstd}; the fp32 normalised rows are re-created by the next residual GEMM's epilogue
 * (cst_gemm_params.res_stats, res_slots == 0) instead of making a round trip through HBM.  Same reference lines as
 * cst_layernorm (wav2vec2.py:945,957). */
int cst_layernorm_ab(const float* x, long long ldx, const float* gamma, const float* beta, void* out_lp, int lp_dtype,
                     long long ldo, float* ab_out, int rows, int C, void* stream);

/* Generic dtype/layout copy x[rows, C] -> y: used for (a) the pos-conv operand layout
 * [B, 16 groups, Tpad, 64] (48 channels + 16 zero lanes, 64 zero frames each side) and (b) broadcasting
 * the M memory embeddings over the batch (w2v2_transformer_interlingua.py:268-269). */
int cst_posconv_pack(const float* x, int B, int rows_per_seg, int n_frames, void* xg, int xg_dtype,
                     int t_pad_rows, void* stream);
int cst_broadcast_rows(const float* src, int rows, int C, int B, float* dst, void* stream);

/* ---- a5 (bf16): grouped pos-conv + bias + GELU + residual with the activation panel resident in shared memory
 * Replaces: `x + GELU(pos_conv(x))` of TransformerEncoder.extract_features (wav2vec2.py:773-786,823-825) for the
 * bf16 mode (the fp32 mode runs the same arithmetic as a batched cst_gemm).  xg: packed bf16 operand of
 * cst_posconv_pack [B,16,t_pad_rows,64]; w: bf16 [16,48,128*64] (tap-major, 64-lane padded); bias [768] f32;
 * resid / out: f32 [B*rows_per_seg, 768]; frames t < n_rows of every utterance are written. */
int cst_posconv(const void* xg, const void* w, const float* bias, const float* resid, float* out,
                int B, int n_rows, int rows_per_seg, int t_pad_rows, void* stream);

/* Same stage, "taps stacked in M" formulation (default for bf16): weights are the A operand, two taps per instruction,
 * frames are N = 256.  w2: bf16 [16 groups, 64 tap pairs, 128 rows, 64 lanes]; rows 0..47 = W[g, co, tap 2p, ci],
 * rows 64..111 = W[g, co, tap 2p+1, ci], other rows and lanes 48..63 zero.  Other arguments as cst_posconv. */
int cst_posconv_stacked(const void* xg, const void* w2, const float* bias, const float* resid, float* out,
                        int B, int n_rows, int rows_per_seg, int t_pad_rows, void* stream);

/* ---- padding-aware attention: softmax(q k^T + keymask) v, head_dim 64 ---------------------------------
 * Replaces: the attention core of F.multi_head_attention_forward as called from
 * multihead_attention.py:165-187 (wav2vec2 layers wav2vec2.py:938-945 with a -inf key-padding mask;
 * shared layers transformer_layer.py:131-137; memory stage w2v2_transformer_interlingua.py:289-298,
 * where kv_len == NULL: memories attend ALL T2 frames incl. padded ones).
 * q is pre-scaled (1/8 is folded into W_q, b_q: exact power of two).  ALL query rows are computed.
 * q [B*q_rows_per_seg, ldq], k/v [B*kv_rows_per_seg, ldkv]; head h at columns h*64..h*64+63.
 * kv_len [B] int32 (device) or NULL => n_kv keys for every utterance. */
int cst_attention(const void* q, const void* k, const void* v, void* out, int dtype,
                  long long ldq, long long ldkv, long long ldo,
                  int B, int H, int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg,
                  const int32_t* kv_len, void* stream);

/* The same attention over a RAGGED set of utterances (several reference batches of different padded widths sharing one
 * row space, so that the row-wise GEMM / LayerNorm launches around it see M >= 24k rows -- DESIGN.md §3a).
 * seg: int32 [B][4] (device, 16-byte aligned) = {first q row, q rows, first kv row, kv rows} of utterance b;
 * max_n_q / max_n_kv = maxima over the table (grid / shared-memory sizing); q_rows_total / kv_rows_total = rows of the
 * q / kv buffers (tensor-map extents).  Everything else as cst_attention (same reference lines). */
int cst_attention_segs(const void* q, const void* k, const void* v, void* out, int dtype,
                       long long ldq, long long ldkv, long long ldo, int B, int H,
                       const int32_t* seg, int max_n_q, int max_n_kv, long long q_rows_total, long long kv_rows_total,
                       const int32_t* kv_len, void* stream);

/* ---- text (MT) input of the shared encoder ------------------------------------------------------
 * Replaces: the text branch of S2T_W2V2_TransformerInterlinguaEncoder.forward (fairseq/models/chimera/
 * w2v2_transformer_interlingua.py:212-217,230-236): x = scale * text_embed_tokens(tokens) + embed_positions(padding mask).
 * tokens int64 [B,T], lengths int64 [B] (device); embed f32 [V,C]; pos_table f32 [T+2, C] = SinusoidalPositionalEmbedding
 * table with row 1 (padding_idx) zero: token t of utterance b gets row t+2 if t < lengths[b], else row 1.
 * x f32 [B*rows_per_seg, C] (rows t >= T of a segment are zero-filled); valid int32 [B] = lengths (may be NULL). */
int cst_text_embed(const int64_t* tokens, const int64_t* lengths, const float* embed, const float* pos_table,
                   float scale, float* x, int32_t* valid, int B, int T, int rows_per_seg, int C, int V, void* stream);
/* Backward of cst_text_embed's gather (training step, text pass): dE[tokens[b,t], :] += scale * dx[b, t, :] for tokens != pad_idx
 * (nn.Embedding(padding_idx): that row gets no gradient); dE f32 [V, C] is accumulated with fp32 atomics (zero it first). */
int cst_embed_bwd(const int64_t* tokens, const float* dx, float scale, float* dE, int B, int T, int rows_per_seg, int C, int V,
                  int pad_idx, void* stream);
/* Sinusoidal positions of the NON-memory base encoder (S2T_W2V2_TransformerEncoder.forward, fairseq/models/chimera/
 * w2v2_transformer.py:353-357): x[b, t, :] += pos_table[t + 2] for t < valid[b] (padded frames get the zero padding row); same table
 * as cst_text_embed, T + 2 rows.  x f32 [B*rows_per_seg, C]. */
int cst_add_positions(float* x, const int32_t* valid, const float* pos_table, int B, int T, int rows_per_seg, int C, void* stream);

/* ==== training heads over the path's outputs =====================================================
 * Contrastive (InfoNCE) loss over the M shared semantic memories of the audio and the text pass.
 * Replaces: TripletSTMTContrastiveCriterion.compute_contrastive (fairseq/criterions/triplet_st_mt_contrastive.py:154-169):
 *   logits[b,i,j] = cosine_similarity(audio[b,i,:], text[b,j,:]) / temp   (cosine in fp32, eps 1e-8)
 *   loss_rows[b*M + j] = cross_entropy over the AUDIO index i of logits[b,:,j], target = j  (the reference hands the 3-D
 *   [b,i,j] tensor to F.cross_entropy, whose class axis is dim 1; reduce=False form -- sum the rows for reduce=True)
 * audio / text: [M, B, C] (encoder_out layout), dtype CST_F32 or CST_BF16; M <= 64, C <= 1024.  lse_ws: B*M floats.
 * d_audio / d_text (both or neither, f32 [M,B,C]): dscale * d(sum of loss_rows)/d(input). */
int cst_contrastive_loss(const void* audio, const void* text, int dtype, int M, int B, int C, float temp,
                         float* loss_rows, float* lse_ws, float* d_audio, float* d_text, float dscale, void* stream);

/* Label-smoothed cross entropy of N target positions over V classes.
 * Replaces: label_smoothed_nll_loss (fairseq/criterions/label_smoothed_cross_entropy.py:13-30) applied to
 * log_softmax(logits): nll_rows[r] = -lprobs[r, target[r]], loss_rows[r] = (1-eps)*nll + eps/V * (-sum_v lprobs[r,v]);
 * rows with target == ignore_index give 0.  dlogits (optional, [N, ldd]) = dscale * d(sum of loss_rows)/d(logits). */
int cst_label_smoothed_ce(const float* logits, long long ld, const int64_t* target, int N, int V, float eps,
                          long long ignore_index, float* loss_rows, float* nll_rows, float* dlogits, long long ldd,
                          float dscale, void* stream);

/* out[0] = sum of n floats, fixed summation order (deterministic): the reduce=True form of the two criteria. */
int cst_sum(const float* x, long long n, float* out, void* stream);

/* ==== backward pass of the path (BASELINE configs[4]) ===========================================
 * The reference differentiates its modules with torch autograd; these are the hand-written derivatives of the forward kernels
 * above (parity: autograd through oracle/chimera_oracle.py).  GEMM-shaped gradients reuse cst_gemm:
 *   dX [M,K] = dY [M,N] W [N,K]      -> cst_gemm(A = dY, W = W^T)        (W^T from cst_transpose, once per step)
 *   dW [N,K] = dY^T [N,M] X [M,K]    -> cst_gemm(A = dY^T, W = X^T)      (both from cst_transpose, M zero-padded to 64)
 * All gradient tensors are fp32. */

/* Tensors named `x_dtype` etc. take CST_F32 / CST_BF16 / CST_F16: the fp32 parity mode keeps the whole tape in fp32, the 16-bit mode
 * keeps GEMM operands and pre-activations in bf16 and every gradient of an activation in fp32.
 *
 * cst_transpose: outT(c, r) = x[r*ldx + c] cast to out_dtype; columns r in [rows, rows_pad) are zero-filled.  ldx may be smaller than
 * cols: the implicit-GEMM window view of a strided convolution's input.  outT is laid out in `chunk`-column slabs, element (c, r) at
 * ((r / chunk)*cols + c)*chunk + r % chunk (chunk <= 0 or == rows_pad: the plain [cols, rows_pad] matrix; smaller chunks make the
 * reduction axis a GEMM batch = split-K for weight gradients).  copy (optional): the un-transposed cast copy, row stride ldcopy. */
int cst_transpose(const void* x, int x_dtype, long long ldx, int rows, int cols, void* outT, int out_dtype, int rows_pad, int chunk,
                  void* copy, long long ldcopy, void* stream);
int cst_cast(const float* x, long long n, void* out, int out_dtype, void* stream);
/* out[c] = scale * sum_r x[r*ldx + c], fixed summation order; ws: CST_COLSUM_WS_FLOATS floats of scratch. */
#define CST_COLSUM_WS_FLOATS (512 * 1024)
int cst_colsum(const void* x, int x_dtype, long long ldx, int rows, int cols, float* out, float* ws, float scale, void* stream);
/* Activations as separate passes (the training forward keeps the pre-activation z): y = act(z) * alpha; GLU reads interleaved
 * (value, gate) column pairs of z [rows, 2*cols_out].  cst_act_bwd: dz from z and dy. */
int cst_act_fwd(int act, const void* z, int z_dtype, long long ldz, int rows, int cols_out, void* y, int y_dtype, long long ldy, float alpha,
                void* stream);
int cst_act_bwd(int act, const void* z, int z_dtype, long long ldz, const void* dy, int dy_dtype, long long ldy, int rows, int cols_out,
                void* dz, int dz_dtype, long long lddz, float alpha, void* stream);
/* LayerNorm backward (eps 1e-5, C in {512, 768}): dx (+)= d LN(x; gamma, beta) / dx . dy; part (optional) receives per-CTA partial
 * sums [cdiv(rows,8)][dgamma | dbeta][C] to be reduced with cst_colsum(ldx = 2C). */
int cst_layernorm_bwd(const float* x, long long ldx, const float* gamma, const float* dy, long long ldy, float* dx, long long lddx,
                      float* part, int rows, int C, int accumulate, void* stream);
/* Attention backward (head_dim 64): recomputes the probabilities tile by tile in fp32.  q / k / v / o: the forward tensors in `dtype`
 * (row strides ldq, ldkv, ldo_fwd); d_o, dq, dk, dv fp32 (row strides ldo, lddq, lddkv).  dk / dv must be zero on entry. */
int cst_attention_bwd(const void* q, const void* k, const void* v, const void* o, int dtype, const float* d_o,
                      float* dq, float* dk, float* dv, long long ldq, long long ldkv, long long ldo_fwd, long long ldo,
                      long long lddq, long long lddkv,
                      int B, int H, int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg, const int32_t* kv_len, void* stream);
/* The same derivative for the 16-bit mode on the tensor cores (csrc/attention_bwd_tc.cu): S = Q K^T and dP = dO V^T as batched
 * tcgen05 GEMMs, one softmax-backward kernel, dQ / dK / dV as three more batched GEMMs.  q / k / v bf16, d_o / dq / dk / dv fp32;
 * rows t < n_q (dq) and t < n_kv (dk, dv) of every utterance are written.  ws: cst_attention_bwd_tc_ws_bytes(...) bytes, 256-byte aligned. */
long long cst_attention_bwd_tc_ws_bytes(int B, int H, int n_q, int n_kv);
int cst_attention_bwd_tc(const void* q, const void* k, const void* v, const float* d_o, float* dq, float* dk, float* dv,
                         long long ldq, long long ldkv, long long ldo, long long lddq, long long lddkv,
                         int B, int H, int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg, const int32_t* kv_len,
                         void* ws, void* stream);
/* Dropout of the attention probabilities (attention_dropout of the training recipe: F.multi_head_attention_forward drops softmax(S)
 * before the product with V; fairseq/modules/multihead_attention.py:155-187).  With p > 0 the training forward runs as
 * S = Q K^T (batched tcgen05 GEMM) -> softmax + dropout kernel -> Pd V (GEMM), and the backward pass regenerates the same masks:
 * keep(b, h, i, j) = Philox word of element ((b*H + h)*up64(n_q) + i)*up64(n_kv) + j under (*seed, site) (csrc/philox.cuh).
 * q / k / v in qkv_dtype (CST_BF16 / CST_F32; the per-head panels are bf16 either way), out in out_dtype; other arguments as
 * cst_attention / cst_attention_bwd_tc.  ws: the *_ws_bytes figure, 256-byte aligned. */
long long cst_attention_dropout_fwd_ws_bytes(int B, int H, int n_q, int n_kv);
int cst_attention_dropout_fwd(const void* q, const void* k, const void* v, int qkv_dtype, void* out, int out_dtype,
                              long long ldq, long long ldkv, long long ldo, int B, int H, int n_q, int q_rows_per_seg, int n_kv,
                              int kv_rows_per_seg, const int32_t* kv_len, float p, const unsigned long long* seed,
                              unsigned int site, void* ws, void* stream);
int cst_attention_bwd_tc_dropout(const void* q, const void* k, const void* v, int qkv_dtype, const float* d_o, float* dq, float* dk,
                                 float* dv, long long ldq, long long ldkv, long long ldo, long long lddq, long long lddkv,
                                 int B, int H, int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg, const int32_t* kv_len,
                                 float p, const unsigned long long* seed, unsigned int site, void* ws, void* stream);
/* Input gradient of an implicit-GEMM strided convolution from its A-operand gradient dcol [M, k*C] (gather form). */
int cst_col2im(const void* dcol, int dcol_dtype, long long M, int k, int stride, int C, float* dx, long long rows_in, int accumulate,
               void* stream);
/* out[seg*out_rps + t + out_off] (+)= scale * in[seg*in_rps + t + in_off] for t < min(seg_valid, seg_len[seg]), else 0 (t < n_rows). */
int cst_rows_remap(const float* in, long long ldi, int in_rps, int in_off, void* out, int out_dtype, long long ldo, int out_rps, int out_off,
                   int n_seg, int n_rows, int C, int seg_valid, const int32_t* seg_len, int accumulate, float scale, void* stream);
/* Dropout of the training step.  Replaces F.dropout as called by FairseqDropout (fairseq/modules/fairseq_dropout.py:16-27) at the
 * elementwise sites of the path (wav2vec2.py:553,830 and the three per-layer dropouts of TransformerSentenceEncoderLayer;
 * w2v2_transformer_interlingua.py:237; transformer_layer.py dropout_module / activation_dropout_module):
 *   out[r, c] = (add ? add[r, c] : 0) + x[r, c] * keep(r, c) / (1 - p),  keep = Philox4x32-10(key = *seed, counter = {(r*cols + c)/4, site})
 *   word (r*cols + c) % 4 >= floor(p * 2^32)   (csrc/philox.cuh; tests/emu.py restates it in numpy, compared bit for bit).
 * Masks are never stored: the derivative is the same call on the gradient with the same *seed and site.  `seed` is a DEVICE pointer (one
 * 64-bit value per step) so that a captured CUDA graph draws fresh masks on every replay.  add: fp32 residual (optional); out2: optional
 * second copy (the GEMM-operand dtype); cols and all row pitches multiples of 4. */
int cst_dropout(const void* x, int x_dtype, long long ldx, const float* add, long long ldadd, void* out, int out_dtype, long long ldo,
                void* out2, int out2_dtype, long long ldo2, int rows, int cols, float p, const unsigned long long* seed,
                unsigned int site, void* stream);
/* Fused Adam update of one tensor.  Replaces fairseq/optim/adam.py:157-224 (fp32 parameters and moments; the gradient is fp32 or the bf16
 * wire format of the all-reduce): m = b1 m + (1-b1) g', v = b2 v + (1-b2) g'^2 with g' = grad_scale * g; p -= weight_decay*lr*p;
 * p -= step_size * m / (sqrt(v) + eps), step_size = lr * sqrt(1 - b2^t) / (1 - b1^t) formed by the caller.  dyn (optional, device
 * float[3] = {lr, step_size, grad_scale}) overrides the by-value arguments, so that one captured CUDA graph serves every step. */
int cst_adam_step(float* p, const void* g, int g_dtype, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                  float weight_decay, float step_size, float grad_scale, const float* dyn, void* stream);
/* conv0 + GroupNorm + GELU backward (recomputes the convolution from the waveform); see csrc/conv0_bwd.cu for the workspace size. */
int cst_conv0_bwd(const float* wave, int B, int L, const float* w, const float* gamma, const float* beta,
                  const float* scale_shift, const float* dout, int rows_per_seg, float* dw, float* dgamma, float* dbeta,
                  float* ws, float grad_scale, void* stream);

/* ==== greedy incremental decoding from the memories (BASELINE configs[3] "encode + greedy decode")
 * All four entry points read the current step from a DEVICE counter, so one captured CUDA graph of a decoding step is
 * replayed for every step without a host round trip.  Activations of the decoder are fp32; weights fp32 or bf16
 * (fp32 accumulation).  head_dim is 64. */

/* x[b,:] = scale * embed[tokens[b, *step], :] + pos_table[*step, :]
 * Replaces: embed_scale * embed_tokens(prev_output_tokens[:, -1:]) + embed_positions(...) of
 * TransformerDecoder.extract_features_scriptable (fairseq/models/transformer.py:761-778);  pos_table row t is the
 * sinusoidal embedding of position padding_idx + 1 + t (fairseq/modules/sinusoidal_positional_embedding.py:71-93).
 * tokens int32 [B, ld_tok]; embed [V, C] (w_dtype); pos_table f32 [>= max steps, C]; x f32 [B, C]. */
int cst_dec_embed(const int32_t* tokens, int ld_tok, const void* embed, int w_dtype, const float* pos_table,
                  float scale, float* x, int B, int C, const int32_t* step, void* stream);

/* out = act(LN?(A) W^T + bias) (+ residual) for M <= a few hundred rows (weight-streaming "skinny" linear).
 * Replaces: the q/k/v/out projections of the incremental MultiheadAttention (fairseq/modules/multihead_attention.py:
 * 189-379), fc1/fc2 and the three pre-LayerNorms of TransformerDecoderLayer.forward (fairseq/modules/
 * transformer_layer.py:300-412), the final layer_norm + output_projection (fairseq/models/transformer.py:816-838).
 *   A [M, K] (a_dtype, row stride lda), W [N, K] (w_dtype), K in {512, 1024, 2048}; ln_gamma/ln_beta (K == 512): LayerNorm
 *   (eps 1e-5) of each A row is applied before the product.  The N outputs are split into n_seg equal segments;
 *   segment s is written to out[s][m*ldo[s] + (*step)*step_stride[s] + col]  (q | K-cache row | V-cache row).
 *   residual [M, N] (row stride ldr, may alias out[0]) only with n_seg == 1.  step may be NULL (= 0). */
typedef struct cst_dec_linear_params {
  const void* A; const void* W; const float* bias; const float* ln_gamma; const float* ln_beta;
  const float* residual; void* out[3];
  long long lda, ldr, ldo[3], step_stride[3];
  const int32_t* step;
  int a_dtype, w_dtype, M, N, K, n_seg, act;
  int out_dtype[3];              /* CST_F32 or CST_BF16 per segment (bf16: the K / V cache rows of the 16-bit mode) */
} cst_dec_linear_params;
int cst_dec_linear(const cst_dec_linear_params* p, void* stream);

/* out[b, h*64:(h+1)*64] = softmax(q[b,h] . K[b, 0..n, h]^T) V[b, 0..n, h];  key row j of hypothesis b starts at
 * k + b*kv_batch_stride + j*kv_row_stride.  n = *step + 1 when step != NULL (self-attention over the cache: the
 * reference's saved_state prev_key/prev_value, multihead_attention.py:249-296), else n_keys (encoder-decoder attention
 * over the M memories with the all-False padding mask, transformer_layer.py:371-392).  q (f32) is pre-scaled; k / v are
 * kv_dtype CST_F32 or CST_BF16 (strides in elements, multiples of 8). */
int cst_dec_attention(const float* q, long long ldq, const void* k, const void* v, int kv_dtype,
                      long long kv_batch_stride, long long kv_row_stride, float* out, long long ldo, int B, int H,
                      int n_keys, int n_keys_max, const int32_t* step, void* stream);
/* Same, with kv_group consecutive decoder rows sharing ONE K / V set (row b reads set b / kv_group): the K beams of a sentence
 * attend the same memories, whose K / V are projected and stored once per sentence instead of once per beam row (the reference
 * repeats encoder_out K times: reorder_encoder_out with new_order = arange(B).repeat_interleave(K), sequence_generator.py:239-243). */
int cst_dec_attention_grouped(const float* q, long long ldq, const void* k, const void* v, int kv_dtype,
                              long long kv_batch_stride, long long kv_row_stride, float* out, long long ldo, int B, int H,
                              int n_keys, int n_keys_max, const int32_t* step, int kv_group, void* stream);

/* One step of SequenceGenerator._generate for beam_size = 1 (fairseq/sequence_generator.py:294-540):
 * lp = log_softmax(logits[b]); lp[pad] = -inf; step >= max_len: only EOS; step < min_len: no EOS; next = argmax.
 * Unfinished rows: tokens[b, step+1] = next, pos_scores[b, step] = lp[next]; next == EOS => done[b] = 1,
 * out_len[b] = step + 1.  counters int32[3]: [0] step (incremented by this call), [1] scratch, [2] rows finished. */
int cst_dec_select(const float* logits, int V, int B, int32_t* tokens, int ld_tok, float* pos_scores, int ld_ps,
                   int32_t* done, int32_t* out_len, int32_t* counters, int max_len, int min_len, int pad, int eos,
                   void* stream);

/* ==== beam search from the memories (checked against the reference's SequenceGenerator(beam_size=5) hypotheses on the host
 * emulator and on the B200) ===============================================================================================
 * Self-attention over the cache with a per-row history table instead of a cache re-order: position j of logical row r lives
 * in physical cache row hist[r*ld_hist + j] for j < *step and in row r for j == *step.
 * Replaces: reorder_incremental_state + attention over saved_state (fairseq/modules/multihead_attention.py:249-296,381-395). */
int cst_dec_attention_beam(const float* q, long long ldq, const void* k, const void* v, int kv_dtype,
                           long long kv_batch_stride, long long kv_row_stride, float* out, long long ldo, int R, int H,
                           int n_keys_max, const int32_t* hist, int ld_hist, const int32_t* step, void* stream);

/* One step of SequenceGenerator._generate for beam_size K <= 8 (fairseq/sequence_generator.py:294-540, finalize_hypos
 * :590-712, BeamSearch.step fairseq/search.py:109-160), one CTA per sentence: lprobs = log_softmax(logits) with the pad /
 * min_len / max_len rules + cumulative scores (step 0: first beam only); top-2K candidates; EOS candidates among the top K
 * are finalised (tokens, per-position scores, score / (step+1)^len_penalty) until K hypotheses exist; the K best remaining
 * candidates continue: tokens / cumulative scores / history rows are permuted from the *_in into the *_out buffers
 * ([B*K, T], T >= max_len + 2).  counters int32[3]: [0] step (incremented), [1] scratch, [2] sentences finished. */
typedef struct cst_dec_beam_params {
  const float* logits; const int32_t* tok_in; int32_t* tok_out; const float* sc_in; float* sc_out;
  const int32_t* hist_in; int32_t* hist_out; int32_t* ignore;
  int32_t* fin_tokens; float* fin_pos; float* fin_score; int32_t* fin_len; int32_t* n_final; int32_t* finished;
  int32_t* counters;
  int B, K, V, T, max_len, min_len, pad, eos;
  float len_penalty;
} cst_dec_beam_params;
int cst_dec_beam_select(const cst_dec_beam_params* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CHIMERA_ST_B200_H_ */
