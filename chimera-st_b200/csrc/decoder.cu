// Greedy incremental decoding of the target tokens from the M memories (BASELINE configs[3]).
//
// One decoding step touches every decoder weight once (27 M parameters, L2-resident after the first step) for at
// most 64 hypotheses: a skinny, weight-streaming problem.  Four kernels, all driven by a DEVICE-side step counter
// so that one captured CUDA graph of a step can be replayed for every step with no host round trip:
//   dec_embed_kernel      x = sqrt(d) * E[token] + sinusoidal position
//   dec_linear_kernel     out = act(LN?(A) W^T + b) (+ residual), up to three output segments (q | K cache | V cache)
//   dec_attention_kernel  one warp per (hypothesis, head): softmax(q K^T) V over the cached keys / the M memories
//   dec_select_kernel     log-softmax, the reference's candidate rules (pad never, EOS only from min_len, only EOS at
//                         max_len), arg-max, token append, finished bookkeeping, step += 1
// fp32 accumulation everywhere (exact fp32 FFMA, the 1e-5 / identical-ID parity mode); weights fp32 or bf16.
//
// Reference arithmetic replaced: TransformerDecoder.extract_features_scriptable (fairseq/models/transformer.py:720-828),
// output_layer (:830-838), TransformerDecoderLayer.forward (fairseq/modules/transformer_layer.py:300-412) with the
// incremental-state MultiheadAttention (fairseq/modules/multihead_attention.py:189-379), SinusoidalPositionalEmbedding
// (fairseq/modules/sinusoidal_positional_embedding.py:71-93) and SequenceGenerator._generate for beam_size = 1
// (fairseq/sequence_generator.py:294-540).
#include "common.cuh"

namespace cst {

constexpr int DL_ROWS = 64;                 // hypotheses per CTA (8 warps x 8 rows)
constexpr int DL_COLS = 8;                  // output features per CTA
constexpr int DL_KC = 512;                  // K chunk resident in shared memory
constexpr int DL_THREADS = 256;
constexpr int DL_SMEM = (DL_ROWS * DL_KC + DL_COLS * DL_KC) * 4;

// exchange-and-add step of the 64 -> 2 transposing warp reduction: afterwards v[0..N/2) hold the sums of the index
// half selected by lane bit `off`
template <int N>
__device__ __forceinline__ void bfly(float (&v)[64], int off, int lane) {
  const bool up = (lane & off) != 0;
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    const float keep = up ? v[i + N / 2] : v[i];
    const float send = up ? v[i] : v[i + N / 2];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
  }
}

struct DecLinearArgs {
  const void* A; long long lda;
  const void* W; const float* bias; const float* ln_g; const float* ln_b;
  const float* residual; long long ldr;
  float* out0; float* out1; float* out2;
  long long ldo0, ldo1, ldo2, ss0, ss1, ss2;
  const int* step;
  int M, N, K, seg_n, act;
};

// Warp w of a CTA owns rows 8w..8w+7 of the CTA's 64-row block and all 8 output columns; a lane owns the K positions
// 4*lane + 128*j.  Per K chunk a lane does 8 + 8 LDS.128 for 256 FMAs; the 64 partial sums per lane are reduced
// across the warp with 62 shuffles (bfly) and land as outputs (row = lane/4, col = 2*(lane%4) + {0,1}).
// The A rows of a warp are private to it (loaded, optionally LayerNorm-ed, and read back by the same warp).
template <typename AT, typename WT>
__global__ void __launch_bounds__(DL_THREADS, 1) dec_linear_kernel(const DecLinearArgs a) {
  extern __shared__ __align__(16) float dl_smem[];
  float* sW = dl_smem + DL_ROWS * DL_KC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * DL_COLS;
  const int m0 = blockIdx.y * DL_ROWS + warp * 8;
  const int rows_here = min(8, a.M - m0);                     // warp-uniform; may be <= 0
  float* myA = dl_smem + warp * 8 * DL_KC;
  const AT* A = reinterpret_cast<const AT*>(a.A);
  const WT* W = reinterpret_cast<const WT*>(a.W);
  const bool ln = a.ln_g != nullptr;
  pdl_launch_dependents();
  pdl_wait();

  float acc[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = 0.f;

  for (int k0 = 0; k0 < a.K; k0 += DL_KC) {
    if (k0) __syncthreads();                                  // previous chunk's sW is still being read
    for (int i = threadIdx.x; i < DL_COLS * DL_KC / 4; i += DL_THREADS) {
      const int c = i / (DL_KC / 4), kk = (i % (DL_KC / 4)) * 4;
      const int n = n0 + c;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < a.N) w = load4(W + (size_t)n * a.K + k0 + kk);
      *reinterpret_cast<float4*>(sW + c * DL_KC + kk) = w;
    }
    for (int r = 0; r < 8; ++r) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows_here) {
        const AT* row = A + (size_t)(m0 + r) * a.lda + k0 + 4 * lane;
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = load4(row + 128 * j);
        if (ln) {                                             // K == 512: the whole row is in this warp's registers
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
          const float mean = warp_sum(s) * (1.0f / DL_KC);
          float q = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
            q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
          }
          const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / DL_KC) + 1e-5f);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 g = load4(a.ln_g + 4 * lane + 128 * j), b = load4(a.ln_b + 4 * lane + 128 * j);
            v[j].x = v[j].x * rstd * g.x + b.x; v[j].y = v[j].y * rstd * g.y + b.y;
            v[j].z = v[j].z * rstd * g.z + b.z; v[j].w = v[j].w * rstd * g.w + b.w;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(myA + r * DL_KC + 4 * lane + 128 * j) = v[j];
    }
    __syncthreads();
    if (rows_here > 0) {
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        float4 w[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) w[c] = *reinterpret_cast<const float4*>(sW + c * DL_KC + 4 * lane + 128 * j);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float4 x = *reinterpret_cast<const float4*>(myA + r * DL_KC + 4 * lane + 128 * j);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float t = acc[r * 8 + c];
            t = fmaf(x.x, w[c].x, t); t = fmaf(x.y, w[c].y, t); t = fmaf(x.z, w[c].z, t); t = fmaf(x.w, w[c].w, t);
            acc[r * 8 + c] = t;
          }
        }
      }
    }
  }
  if (rows_here <= 0) return;
  bfly<64>(acc, 16, lane); bfly<32>(acc, 8, lane); bfly<16>(acc, 4, lane); bfly<8>(acc, 2, lane); bfly<4>(acc, 1, lane);
  const int m = m0 + (lane >> 2);
  if (m >= a.M) return;
  const long long step = a.step ? (long long)*a.step : 0;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int n = n0 + 2 * (lane & 3) + j;
    if (n >= a.N) continue;
    float v = acc[j];
    if (a.bias) v += a.bias[n];
    if (a.act == CST_ACT_RELU) v = fmaxf(v, 0.f);
    const int seg = n / a.seg_n, col = n - seg * a.seg_n;
    if (a.residual) v += a.residual[(size_t)m * a.ldr + n];
    float* o = seg == 0 ? a.out0 : (seg == 1 ? a.out1 : a.out2);
    const long long ldo = seg == 0 ? a.ldo0 : (seg == 1 ? a.ldo1 : a.ldo2);
    const long long ss = seg == 0 ? a.ss0 : (seg == 1 ? a.ss1 : a.ss2);
    o[(size_t)m * ldo + step * ss + col] = v;
  }
}

// x[b,:] = scale * E[tokens[b, step], :] + pos[step, :]
template <typename WT>
__global__ void dec_embed_kernel(const int* __restrict__ tokens, int ld_tok, const WT* __restrict__ E,
                                 const float* __restrict__ pos, float scale, float* __restrict__ x, int C,
                                 const int* __restrict__ step) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, t = *step;
  const int tok = tokens[(size_t)b * ld_tok + t];
  for (int c = 4 * threadIdx.x; c < C; c += 4 * blockDim.x) {
    const float4 e = load4(E + (size_t)tok * C + c), p = load4(pos + (size_t)t * C + c);
    store4(x + (size_t)b * C + c, make_float4(fmaf(scale, e.x, p.x), fmaf(scale, e.y, p.y), fmaf(scale, e.z, p.z), fmaf(scale, e.w, p.w)));
  }
}

// one warp per (hypothesis, head); head_dim 64; q pre-scaled.  n keys = *step + 1 (self-attention over the cache) or
// n_keys (cross-attention over the memories; the reference passes an all-False key-padding mask).
__global__ void __launch_bounds__(128) dec_attention_kernel(const float* __restrict__ q, long long ldq,
                                                            const float* __restrict__ k, const float* __restrict__ v,
                                                            long long kv_bs, long long kv_rs, float* __restrict__ out,
                                                            long long ldo, int B, int H, int n_keys, int n_max,
                                                            const int* __restrict__ step) {
  extern __shared__ __align__(16) float da_smem[];
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wid = blockIdx.x * 4 + warp;
  if (wid >= B * H) return;
  const int b = wid / H, h = wid - b * H;
  float* sq = da_smem + (size_t)warp * (64 + ((n_max + 3) & ~3));   // 16-byte aligned per warp
  float* ss = sq + 64;
  const int n = step ? min(*step + 1, n_max) : n_keys;
  sq[lane] = q[(size_t)b * ldq + h * 64 + lane];
  sq[lane + 32] = q[(size_t)b * ldq + h * 64 + lane + 32];
  __syncwarp();
  const float* kb = k + (size_t)b * kv_bs + h * 64;
  const float* vb = v + (size_t)b * kv_bs + h * 64;
  float mx = -INFINITY;
  for (int key = lane; key < n; key += 32) {
    const float* kr = kb + (size_t)key * kv_rs;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      const float4 kk = load4(kr + 4 * d), qq = *reinterpret_cast<const float4*>(sq + 4 * d);
      s = fmaf(kk.x, qq.x, s); s = fmaf(kk.y, qq.y, s); s = fmaf(kk.z, qq.z, s); s = fmaf(kk.w, qq.w, s);
    }
    ss[key] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int key = lane; key < n; key += 32) {
    const float p = expf(ss[key] - mx);
    ss[key] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  __syncwarp();
  float o0 = 0.f, o1 = 0.f;
  for (int key = 0; key < n; ++key) {
    const float p = ss[key];
    const float2 vv = *reinterpret_cast<const float2*>(vb + (size_t)key * kv_rs + 2 * lane);
    o0 = fmaf(p, vv.x, o0);
    o1 = fmaf(p, vv.y, o1);
  }
  const float inv = 1.0f / sum;
  *reinterpret_cast<float2*>(out + (size_t)b * ldo + h * 64 + 2 * lane) = make_float2(o0 * inv, o1 * inv);
}

// counters: [0] step, [1] CTAs finished (scratch), [2] hypotheses finished
__global__ void __launch_bounds__(256) dec_select_kernel(const float* __restrict__ logits, int V, int* __restrict__ tokens,
                                                         int ld_tok, float* __restrict__ pos_scores, int ld_ps,
                                                         int* __restrict__ done, int* __restrict__ out_len,
                                                         int* counters, int max_len, int min_len, int pad, int eos) {
  __shared__ float s_f[8];
  __shared__ float s_b[8];
  __shared__ int s_i[8];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int step = *reinterpret_cast<volatile int*>(counters);
  const float* row = logits + (size_t)b * V;
  float m_all = -INFINITY, best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = tid; i < V; i += 256) {
    const float x = row[i];
    m_all = fmaxf(m_all, x);
    const bool allowed = (i != pad) && (step >= max_len ? (i == eos) : (step < min_len ? (i != eos) : true));
    if (allowed && x > best) { best = x; bi = i; }             // ascending i per thread: first maximum wins
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m_all = fmaxf(m_all, __shfl_xor_sync(0xffffffffu, m_all, o));
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) { s_f[warp] = m_all; s_b[warp] = best; s_i[warp] = bi; }
  __syncthreads();
  m_all = s_f[0]; best = s_b[0]; bi = s_i[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) {
    m_all = fmaxf(m_all, s_f[w]);
    if (s_b[w] > best || (s_b[w] == best && s_i[w] < bi)) { best = s_b[w]; bi = s_i[w]; }
  }
  __syncthreads();
  float sum = 0.f;
  for (int i = tid; i < V; i += 256) sum += expf(row[i] - m_all);
  sum = warp_sum(sum);
  if (lane == 0) s_f[warp] = sum;
  __syncthreads();
  if (tid == 0) {
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += s_f[w];
    if (bi < 0 || bi >= V) { bi = eos; best = row[eos]; }      // nothing selectable (non-finite logits): stop the hypothesis
    const float lp = best - (m_all + logf(sum));
    if (step > max_len) return;                                // replayed past the last step: nothing to do
    if (!done[b]) {
      tokens[(size_t)b * ld_tok + step + 1] = bi;
      pos_scores[(size_t)b * ld_ps + step] = lp;
      if (bi == eos) { done[b] = 1; out_len[b] = step + 1; atomicAdd(&counters[2], 1); }
    } else {
      tokens[(size_t)b * ld_tok + step + 1] = eos;             // finished rows keep stepping; their output is ignored
    }
    __threadfence();
    if (atomicAdd(&counters[1], 1) == (int)gridDim.x - 1) {   // last CTA of the step: every CTA has read `step`
      counters[1] = 0;
      counters[0] = step + 1;
    }
  }
}

}  // namespace cst

using namespace cst;

extern "C" int cst_dec_embed(const int32_t* tokens, int ld_tok, const void* embed, int w_dtype, const float* pos_table,
                             float scale, float* x, int B, int C, const int32_t* step, void* stream) {
  CST_REQUIRE(tokens && embed && pos_table && x && step, "cst_dec_embed: null pointer");
  CST_REQUIRE(B > 0 && C > 0 && C % 4 == 0, "cst_dec_embed: bad B=%d C=%d", B, C);
  cudaStream_t st = (cudaStream_t)stream;
  if (w_dtype == CST_F32)
    CST_CHECK_CUDA(launch_k(dec_embed_kernel<float>, dim3(B), dim3(128), 0, st, tokens, ld_tok, (const float*)embed, pos_table, scale, x, C, step));
  else if (w_dtype == CST_BF16)
    CST_CHECK_CUDA(launch_k(dec_embed_kernel<__nv_bfloat16>, dim3(B), dim3(128), 0, st, tokens, ld_tok, (const __nv_bfloat16*)embed, pos_table, scale, x, C, step));
  else
    CST_REQUIRE(false, "cst_dec_embed: unsupported dtype %d", w_dtype);
  return CST_OK;
}

static int dec_linear_init() {
  // all four instantiations up front: the first launch of one of them may happen inside a stream capture
  CST_CHECK_CUDA(cudaFuncSetAttribute(dec_linear_kernel<float, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, DL_SMEM));
  CST_CHECK_CUDA(cudaFuncSetAttribute(dec_linear_kernel<float, __nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, DL_SMEM));
  CST_CHECK_CUDA(cudaFuncSetAttribute(dec_linear_kernel<__nv_bfloat16, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, DL_SMEM));
  CST_CHECK_CUDA(cudaFuncSetAttribute(dec_linear_kernel<__nv_bfloat16, __nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, DL_SMEM));
  return CST_OK;
}

template <typename AT, typename WT>
static int launch_dec_linear(const DecLinearArgs& a, cudaStream_t st) {
  static int init_rc = dec_linear_init();
  if (init_rc != CST_OK) return init_rc;
  dim3 grid(cdiv(a.N, DL_COLS), cdiv(a.M, DL_ROWS));
  CST_CHECK_CUDA(launch_k(dec_linear_kernel<AT, WT>, grid, dim3(DL_THREADS), (size_t)DL_SMEM, st, a));
  return CST_OK;
}

extern "C" int cst_dec_linear(const cst_dec_linear_params* p, void* stream) {
  CST_REQUIRE(p && p->A && p->W && p->out[0], "cst_dec_linear: null pointer");
  CST_REQUIRE(p->M > 0 && p->N > 0 && p->K > 0 && p->K % DL_KC == 0, "cst_dec_linear: K=%d must be a multiple of %d", p->K, DL_KC);
  CST_REQUIRE(p->n_seg >= 1 && p->n_seg <= 3 && p->N % p->n_seg == 0, "cst_dec_linear: bad n_seg=%d for N=%d", p->n_seg, p->N);
  CST_REQUIRE(!(p->ln_gamma || p->ln_beta) || (p->ln_gamma && p->ln_beta && p->K == DL_KC), "cst_dec_linear: fused LayerNorm needs K == %d", DL_KC);
  CST_REQUIRE(!p->residual || p->n_seg == 1, "cst_dec_linear: residual only with one output segment");
  CST_REQUIRE(p->act == CST_ACT_NONE || p->act == CST_ACT_RELU, "cst_dec_linear: act %d unsupported", p->act);
  CST_REQUIRE(p->lda % 4 == 0, "cst_dec_linear: lda must be a multiple of 4");
  for (int s = 0; s < p->n_seg; ++s) CST_REQUIRE(p->out[s], "cst_dec_linear: out[%d] is null", s);
  DecLinearArgs a;
  a.A = p->A; a.lda = p->lda; a.W = p->W; a.bias = p->bias; a.ln_g = p->ln_gamma; a.ln_b = p->ln_beta;
  a.residual = p->residual; a.ldr = p->ldr;
  a.out0 = p->out[0]; a.out1 = p->out[1]; a.out2 = p->out[2];
  a.ldo0 = p->ldo[0]; a.ldo1 = p->ldo[1]; a.ldo2 = p->ldo[2];
  a.ss0 = p->step_stride[0]; a.ss1 = p->step_stride[1]; a.ss2 = p->step_stride[2];
  a.step = p->step; a.M = p->M; a.N = p->N; a.K = p->K; a.seg_n = p->N / p->n_seg; a.act = p->act;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->a_dtype == CST_F32 && p->w_dtype == CST_F32) return launch_dec_linear<float, float>(a, st);
  if (p->a_dtype == CST_F32 && p->w_dtype == CST_BF16) return launch_dec_linear<float, __nv_bfloat16>(a, st);
  if (p->a_dtype == CST_BF16 && p->w_dtype == CST_F32) return launch_dec_linear<__nv_bfloat16, float>(a, st);
  if (p->a_dtype == CST_BF16 && p->w_dtype == CST_BF16) return launch_dec_linear<__nv_bfloat16, __nv_bfloat16>(a, st);
  CST_REQUIRE(false, "cst_dec_linear: unsupported dtypes a=%d w=%d", p->a_dtype, p->w_dtype);
  return CST_OK;
}

extern "C" int cst_dec_attention(const float* q, long long ldq, const float* k, const float* v, long long kv_batch_stride,
                                 long long kv_row_stride, float* out, long long ldo, int B, int H, int n_keys,
                                 int n_keys_max, const int32_t* step, void* stream) {
  CST_REQUIRE(q && k && v && out, "cst_dec_attention: null pointer");
  CST_REQUIRE(B > 0 && H > 0 && n_keys_max > 0 && (step || (n_keys > 0 && n_keys <= n_keys_max)), "cst_dec_attention: bad sizes");
  CST_REQUIRE(kv_row_stride % 4 == 0 && kv_batch_stride % 4 == 0 && ldo % 2 == 0, "cst_dec_attention: strides must keep 16-byte rows");
  const size_t smem = 4 * (size_t)(64 + ((n_keys_max + 3) & ~3)) * sizeof(float);
  CST_REQUIRE(smem <= 48 * 1024, "cst_dec_attention: n_keys_max=%d too large", n_keys_max);
  CST_CHECK_CUDA(launch_k(dec_attention_kernel, dim3(cdiv((long long)B * H, 4)), dim3(128), smem, (cudaStream_t)stream, q, ldq, k, v,
                          kv_batch_stride, kv_row_stride, out, ldo, B, H, n_keys, n_keys_max, step));
  return CST_OK;
}

extern "C" int cst_dec_select(const float* logits, int V, int B, int32_t* tokens, int ld_tok, float* pos_scores, int ld_ps,
                              int32_t* done, int32_t* out_len, int32_t* counters, int max_len, int min_len, int pad, int eos,
                              void* stream) {
  CST_REQUIRE(logits && tokens && pos_scores && done && out_len && counters, "cst_dec_select: null pointer");
  CST_REQUIRE(V > 0 && B > 0 && eos >= 0 && eos < V && max_len >= 0 && ld_tok >= max_len + 2 && ld_ps >= max_len + 1,
              "cst_dec_select: bad sizes (V=%d B=%d max_len=%d ld_tok=%d ld_ps=%d)", V, B, max_len, ld_tok, ld_ps);
  CST_CHECK_CUDA(launch_k(dec_select_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, logits, V, tokens, ld_tok, pos_scores,
                          ld_ps, done, out_len, counters, max_len, min_len, pad, eos));
  return CST_OK;
}
