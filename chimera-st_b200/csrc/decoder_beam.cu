// Beam search on the device (decoder.py: B200BeamDecoder).  Host logic and ABI semantics are verified against the oracle's
// beam search -- itself pinned to the reference's SequenceGenerator -- on the host emulator, the kernels against the same
// goldens on the B200 (tests/test_gpu_beam.py).
//
//   dec_attention_beam_kernel  self-attention over the cache with a per-row history table: position j of logical row r
//                              lives in physical cache row hist[r][j] (j < step) or r (j == step), so that re-ordering the
//                              beams moves T small integers per row instead of the K/V cache
//   dec_beam_select_kernel     one CTA per sentence: log-softmax + cumulative scores, candidate rules, top-2K candidates,
//                              finalisation of EOS candidates among the top K, selection of the K continuing beams,
//                              permutation of tokens / scores / history into the next step's buffers
// Reference arithmetic replaced: BeamSearch.step (fairseq/search.py:109-160), SequenceGenerator._generate and
// finalize_hypos (fairseq/sequence_generator.py:294-540, 590-712), reorder_incremental_state of the decoder
// (fairseq/modules/multihead_attention.py:381-395).
#include "common.cuh"

namespace cst {

constexpr int DB_THREADS = 256;
constexpr int DB_MAXK = 8;              // beam width limit (candidates 2K <= 16)

template <typename KVT> __device__ __forceinline__ float2 db_load2(const KVT* p);
template <> __device__ __forceinline__ float2 db_load2<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }
template <> __device__ __forceinline__ float2 db_load2<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

template <typename KVT>
__global__ void __launch_bounds__(DB_THREADS) dec_attention_beam_kernel(const float* __restrict__ q, long long ldq,
                                                                        const KVT* __restrict__ k, const KVT* __restrict__ v,
                                                                        long long kv_bs, long long kv_rs, float* __restrict__ out,
                                                                        long long ldo, int H, int n_max,
                                                                        const int* __restrict__ hist, int ld_hist,
                                                                        const int* __restrict__ step) {
  extern __shared__ __align__(16) float dab_smem[];           // [64 q][8*64 partial o][16 red][n_max scores][n_max rows]
  float* sq = dab_smem;
  float* po = dab_smem + 64;
  float* red = po + 8 * 64;
  float* ss = red + 16;
  int* srow = reinterpret_cast<int*>(ss + n_max);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = blockIdx.x / H, h = blockIdx.x - r * H;
  pdl_launch_dependents();
  pdl_wait();                        // q / the cache rows of this step are written by the preceding dec_linear launch
  const int cur = min(*step, n_max - 1), n = cur + 1;
  if (tid < 64) sq[tid] = q[(size_t)r * ldq + h * 64 + tid];
  for (int j = tid; j < n; j += DB_THREADS) srow[j] = j == cur ? r : hist[(size_t)r * ld_hist + j];
  __syncthreads();
  const int sub = lane & 3;
  float4 qq[4];
#pragma unroll
  for (int d = 0; d < 4; ++d) qq[d] = *reinterpret_cast<const float4*>(sq + 16 * sub + 4 * d);
  float mx = -INFINITY;
  for (int key0 = 0; key0 < n; key0 += 64) {
    const int key = key0 + warp * 8 + (lane >> 2);
    float s = 0.f;
    if (key < n) {
      const KVT* kr = k + (size_t)srow[key] * kv_bs + (size_t)key * kv_rs + h * 64 + 16 * sub;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const float4 kk = load4(kr + 4 * d);
        s = fmaf(kk.x, qq[d].x, s); s = fmaf(kk.y, qq[d].y, s); s = fmaf(kk.z, qq[d].z, s); s = fmaf(kk.w, qq[d].w, s);
      }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (key < n) {
      if (sub == 0) ss[key] = s;
      mx = fmaxf(mx, s);
    }
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  float sum = 0.f;
  for (int key = tid; key < n; key += DB_THREADS) {
    const float p = expf(ss[key] - mx);
    ss[key] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[8 + warp] = sum;
  __syncthreads();
  float o0 = 0.f, o1 = 0.f;
  for (int key = warp; key < n; key += 8) {
    const float p = ss[key];
    const float2 vv = db_load2<KVT>(v + (size_t)srow[key] * kv_bs + (size_t)key * kv_rs + h * 64 + 2 * lane);
    o0 = fmaf(p, vv.x, o0);
    o1 = fmaf(p, vv.y, o1);
  }
  *reinterpret_cast<float2*>(po + warp * 64 + 2 * lane) = make_float2(o0, o1);
  __syncthreads();
  if (tid < 64) {
    float tot = 0.f, o = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { tot += red[8 + w]; o += po[w * 64 + tid]; }
    out[(size_t)r * ldo + h * 64 + tid] = o / tot;
  }
}

struct DecBeamArgs {
  const float* logits; int V, K, T;                  // logits [B*K, V]; T = max_len + 2 columns
  const int* tok_in; int* tok_out;                   // [B*K, T]
  const float* sc_in; float* sc_out;                 // cumulative scores [B*K, T]
  const int* hist_in; int* hist_out;                 // [B*K, T]
  int* ignore;                                       // [B*K] cands_to_ignore
  int* fin_tokens; float* fin_pos; float* fin_score; int* fin_len; int* n_final; int* finished;   // [B,K,T] [B,K,T] [B,K] [B,K] [B] [B]
  int* counters;                                     // [0] step, [1] CTAs done (scratch), [2] sentences finished
  int max_len, min_len, pad, eos;
  float len_penalty;
};

// candidate value of flat index (beam kb, token t): lprob + cumulative score, with the reference's masks
__device__ __forceinline__ float db_cand(const DecBeamArgs& a, const float* row, float lse, float prev, int t, int step) {
  float v = row[t] - lse;
  if (v != v) v = -INFINITY;
  if (t == a.pad) v = -INFINITY;
  if (step >= a.max_len) { if (t != a.eos) v = -INFINITY; }
  else if (step < a.min_len) { if (t == a.eos) v = -INFINITY; }
  return v + prev;
}

__global__ void __launch_bounds__(DB_THREADS) dec_beam_select_kernel(const DecBeamArgs a) {
  __shared__ float s_lse[DB_MAXK], s_prev[DB_MAXK];
  __shared__ float s_redv[DB_THREADS / 32];
  __shared__ int s_redi[DB_THREADS / 32];
  __shared__ float s_cv[2 * DB_MAXK];
  __shared__ int s_ci[2 * DB_MAXK];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.K, V = a.V, T = a.T, C = 2 * K;
  pdl_launch_dependents();
  pdl_wait();                        // the logits come from the preceding dec_linear launch
  const int step = *reinterpret_cast<volatile int*>(a.counters);
  const bool live = step <= a.max_len && !a.finished[b];
  if (live) {
    const int nb = step == 0 ? 1 : K;                          // first step: only the first beam (all rows are identical)
    // log-sum-exp of every beam row (warp kb)
    for (int kb = warp; kb < nb; kb += DB_THREADS / 32) {
      const float* row = a.logits + (size_t)(b * K + kb) * V;
      float m = -INFINITY;
      for (int t = lane; t < V; t += 32) m = fmaxf(m, row[t]);
      m = warp_max(m);
      float s = 0.f;
      for (int t = lane; t < V; t += 32) s += expf(row[t] - m);
      s = warp_sum(s);
      if (lane == 0) {
        s_lse[kb] = m + logf(s);
        s_prev[kb] = step == 0 ? 0.f : a.sc_in[(size_t)(b * K + kb) * T + step - 1];
      }
    }
    __syncthreads();
    // top-C candidates by repeated block-wide arg-max (ties: lowest flat index), excluding the ones already taken
    const int total = nb * V, ncand = min(C, total - 1);
    for (int c = 0; c < ncand; ++c) {
      float best = -INFINITY;
      int bi = 0x7fffffff;
      for (int f = tid; f < total; f += DB_THREADS) {
        bool taken = false;
        for (int p = 0; p < c; ++p) taken |= (s_ci[p] == f);
        if (taken) continue;
        const int kb = f / V, t = f - kb * V;
        const float v = db_cand(a, a.logits + (size_t)(b * K + kb) * V, s_lse[kb], s_prev[kb], t, step);
        if (v > best || (v == best && f < bi)) { best = v; bi = f; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (lane == 0) { s_redv[warp] = best; s_redi[warp] = bi; }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < DB_THREADS / 32; ++w)
          if (s_redv[w] > best || (s_redv[w] == best && s_redi[w] < bi)) { best = s_redv[w]; bi = s_redi[w]; }
        s_cv[c] = best; s_ci[c] = bi;
      }
      __syncthreads();
    }
    if (tid == 0) {
      // ---- bookkeeping of SequenceGenerator._generate for one sentence
      bool eos_mask[2 * DB_MAXK];
      int cbeam[2 * DB_MAXK], ctok[2 * DB_MAXK];
      for (int c = 0; c < ncand; ++c) {
        cbeam[c] = s_ci[c] / V; ctok[c] = s_ci[c] - cbeam[c] * V;
        eos_mask[c] = ctok[c] == a.eos && s_cv[c] != -INFINITY;
        if (c < K && a.ignore[b * K + c]) eos_mask[c] = false;
      }
      int nf = a.n_final[b];
      for (int c = 0; c < K && c < ncand; ++c) {               // finalize_hypos: EOS among the top K candidates
        if (!eos_mask[c] || nf >= K) continue;
        const int src = b * K + cbeam[c];
        int* ft = a.fin_tokens + (size_t)(b * K + nf) * T;
        float* fp = a.fin_pos + (size_t)(b * K + nf) * T;
        float prevc = 0.f;
        for (int j = 0; j <= step; ++j) {
          ft[j] = j == step ? a.eos : a.tok_in[(size_t)src * T + j + 1];
          const float cum = j == step ? s_cv[c] : a.sc_in[(size_t)src * T + j];
          fp[j] = cum - prevc;                                 // cumulative -> per-position scores
          prevc = cum;
        }
        a.fin_len[b * K + nf] = step + 1;
        a.fin_score[b * K + nf] = s_cv[c] / powf((float)(step + 1), a.len_penalty);
        ++nf;
      }
      a.n_final[b] = nf;
      if (nf == K || step == a.max_len) {                      // is_finished
        a.finished[b] = 1;
        atomicAdd(&a.counters[2], 1);
      } else {
        // the K best candidates that are neither EOS nor ignored continue (stable in candidate order)
        int active[DB_MAXK], na = 0;
        bool flag[2 * DB_MAXK];
        for (int c = 0; c < ncand; ++c) flag[c] = c < K ? (a.ignore[b * K + c] != 0 || eos_mask[c]) : eos_mask[c];
        for (int c = 0; c < ncand && na < K; ++c) if (!flag[c]) active[na++] = c;
        int n_clean = na;
        for (int c = 0; c < ncand && na < K; ++c) if (flag[c]) active[na++] = c;
        for (int i = 0; i < K; ++i) {
          const int c = active[i], src = b * K + cbeam[c], dst = b * K + i;
          a.ignore[dst] = i >= n_clean;
          for (int j = 0; j <= step; ++j) {
            a.tok_out[(size_t)dst * T + j] = a.tok_in[(size_t)src * T + j];
            if (j < step) a.sc_out[(size_t)dst * T + j] = a.sc_in[(size_t)src * T + j];
            a.hist_out[(size_t)dst * T + j] = j == step ? src : a.hist_in[(size_t)src * T + j];
          }
          a.tok_out[(size_t)dst * T + step + 1] = ctok[c];
          a.sc_out[(size_t)dst * T + step] = s_cv[c];
        }
      }
    }
  }
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&a.counters[1], 1) == (int)gridDim.x - 1) {  // last CTA of the step: every CTA has read `step`
      a.counters[1] = 0;
      if (step <= a.max_len) a.counters[0] = step + 1;
    }
  }
}

}  // namespace cst

using namespace cst;

extern "C" int cst_dec_attention_beam(const float* q, long long ldq, const void* k, const void* v, int kv_dtype,
                                      long long kv_batch_stride, long long kv_row_stride, float* out, long long ldo, int R, int H,
                                      int n_keys_max, const int32_t* hist, int ld_hist, const int32_t* step, void* stream) {
  CST_REQUIRE(q && k && v && out && hist && step, "cst_dec_attention_beam: null pointer");
  CST_REQUIRE(R > 0 && H > 0 && n_keys_max > 0 && ld_hist >= n_keys_max, "cst_dec_attention_beam: bad sizes");
  CST_REQUIRE(kv_row_stride % 8 == 0 && kv_batch_stride % 8 == 0 && ldo % 2 == 0, "cst_dec_attention_beam: strides must keep 16-byte rows");
  const size_t smem = (size_t)(64 + 8 * 64 + 16 + 2 * n_keys_max) * sizeof(float);
  CST_REQUIRE(smem <= 48 * 1024, "cst_dec_attention_beam: n_keys_max=%d too large", n_keys_max);
  cudaStream_t st = (cudaStream_t)stream;
  if (kv_dtype == CST_F32)
    CST_CHECK_CUDA(launch_k(dec_attention_beam_kernel<float>, dim3(R * H), dim3(DB_THREADS), smem, st, q, ldq, (const float*)k,
                            (const float*)v, kv_batch_stride, kv_row_stride, out, ldo, H, n_keys_max, hist, ld_hist, step));
  else if (kv_dtype == CST_BF16)
    CST_CHECK_CUDA(launch_k(dec_attention_beam_kernel<__nv_bfloat16>, dim3(R * H), dim3(DB_THREADS), smem, st, q, ldq,
                            (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, kv_batch_stride, kv_row_stride, out, ldo, H,
                            n_keys_max, hist, ld_hist, step));
  else
    CST_REQUIRE(false, "cst_dec_attention_beam: kv_dtype %d unsupported", kv_dtype);
  return CST_OK;
}

extern "C" int cst_dec_beam_select(const cst_dec_beam_params* p, void* stream) {
  CST_REQUIRE(p && p->logits && p->tok_in && p->tok_out && p->sc_in && p->sc_out && p->hist_in && p->hist_out && p->ignore &&
              p->fin_tokens && p->fin_pos && p->fin_score && p->fin_len && p->n_final && p->finished && p->counters,
              "cst_dec_beam_select: null pointer");
  CST_REQUIRE(p->B > 0 && p->K >= 1 && p->K <= DB_MAXK && p->V > 2 * p->K && p->T >= p->max_len + 2 && p->max_len >= 0 &&
              p->eos >= 0 && p->eos < p->V, "cst_dec_beam_select: bad sizes (B=%d K=%d V=%d T=%d max_len=%d)", p->B, p->K, p->V, p->T, p->max_len);
  DecBeamArgs a;
  a.logits = p->logits; a.V = p->V; a.K = p->K; a.T = p->T;
  a.tok_in = p->tok_in; a.tok_out = p->tok_out; a.sc_in = p->sc_in; a.sc_out = p->sc_out; a.hist_in = p->hist_in; a.hist_out = p->hist_out;
  a.ignore = p->ignore; a.fin_tokens = p->fin_tokens; a.fin_pos = p->fin_pos; a.fin_score = p->fin_score; a.fin_len = p->fin_len;
  a.n_final = p->n_final; a.finished = p->finished; a.counters = p->counters;
  a.max_len = p->max_len; a.min_len = p->min_len; a.pad = p->pad; a.eos = p->eos; a.len_penalty = p->len_penalty;
  CST_CHECK_CUDA(launch_k(dec_beam_select_kernel, dim3(p->B), dim3(DB_THREADS), 0, (cudaStream_t)stream, a));
  return CST_OK;
}
