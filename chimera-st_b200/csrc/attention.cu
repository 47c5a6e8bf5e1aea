// Padding-aware attention, head_dim 64: O = softmax(Q K^T + keymask) V with an online (flash-style)
// softmax.  All arithmetic in fp32 on the CUDA cores: this is the exact-parity kernel (fp32 mode and
// the tiny M-query memory stage); keys >= kv_len[b] get -inf exactly like the reference's
// key_padding_mask, every query row is computed (padded query rows are live in the reference).
//   tile: 64 queries x 64 keys, 256 threads, 4x4 micro-tiles; Q^T/K^T kept transposed in shared memory so
//   the inner loops are float4 broadcast/stride-1 reads; P overwrites the K^T tile.
#include "common.cuh"

namespace cst {

constexpr int AT_BQ = 64, AT_BK = 64, AT_D = 64, AT_LD = 68;
constexpr int AT_SMEM = (2 * AT_D * AT_LD + AT_BK * AT_D) * 4;

template <typename T>
__global__ void __launch_bounds__(256) attention_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                        const T* __restrict__ v, T* __restrict__ out,
                                                        long long ldq, long long ldkv, long long ldo,
                                                        int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg,
                                                        const int32_t* __restrict__ kv_len, const int4* __restrict__ seg) {
  extern __shared__ float sm[];
  float* Qt = sm;                         // [d][i], pitch AT_LD
  float* Kt = sm + AT_D * AT_LD;          // [d][j]; reused as Pt [j][i]
  float* Vs = sm + 2 * AT_D * AT_LD;      // [j][d], pitch 64
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AT_BQ;
  // ragged batches: seg[b] = {first q row, q rows, first kv row, kv rows} (static plan data)
  long long q_row0 = (long long)b * q_rows_per_seg, kv_row0 = (long long)b * kv_rows_per_seg;
  if (seg != nullptr) { const int4 sg = seg[b]; q_row0 = sg.x; n_q = sg.y; kv_row0 = sg.z; n_kv = sg.w; }
  if (q0 >= n_q) return;
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  int klen = kv_len ? kv_len[b] : n_kv;
  if (klen > n_kv) klen = n_kv;

  const T* qb = q + q_row0 * ldq + h * AT_D;
  const T* kb = k + kv_row0 * ldkv + h * AT_D;
  const T* vb = v + kv_row0 * ldkv + h * AT_D;

  // Q tile -> Qt (transposed).  loader: row = tid & 63, 4 float4 columns per thread
  {
    const int r = tid & 63, c4 = tid >> 6;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int d = (c4 * 4 + i) * 4;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q0 + r < n_q) t = load4(qb + (long long)(q0 + r) * ldq + d);
      Qt[(d + 0) * AT_LD + r] = t.x; Qt[(d + 1) * AT_LD + r] = t.y;
      Qt[(d + 2) * AT_LD + r] = t.z; Qt[(d + 3) * AT_LD + r] = t.w;
    }
  }
  float m_run[4], l_run[4], o[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    m_run[a] = -INFINITY; l_run[a] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) o[a][c] = 0.f;
  }

  for (int k0 = 0; k0 < klen; k0 += AT_BK) {
    __syncthreads();                      // previous PV done (and Qt visible on the first pass)
    {
      const int r = tid & 63, c4 = tid >> 6;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int d = (c4 * 4 + i) * 4;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f), u = t;
        if (k0 + r < klen) {
          t = load4(kb + (long long)(k0 + r) * ldkv + d);
          u = load4(vb + (long long)(k0 + r) * ldkv + d);
        }
        Kt[(d + 0) * AT_LD + r] = t.x; Kt[(d + 1) * AT_LD + r] = t.y;
        Kt[(d + 2) * AT_LD + r] = t.z; Kt[(d + 3) * AT_LD + r] = t.w;
        *reinterpret_cast<float4*>(&Vs[r * AT_D + d]) = u;
      }
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[a][c] = 0.f;
#pragma unroll 8
    for (int d = 0; d < AT_D; ++d) {
      const float4 qa = *reinterpret_cast<const float4*>(&Qt[d * AT_LD + ty * 4]);
      const float4 kc = *reinterpret_cast<const float4*>(&Kt[d * AT_LD + tx * 4]);
      const float qv[4] = {qa.x, qa.y, qa.z, qa.w}, kv[4] = {kc.x, kc.y, kc.z, kc.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) s[a][c] = fmaf(qv[a], kv[c], s[a][c]);
    }
    // mask + online softmax (row statistics shared by the 16 lanes of a row group)
    float scale_o[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (k0 + tx * 4 + c >= klen) s[a][c] = -INFINITY;
        mx = fmaxf(mx, s[a][c]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m_run[a], mx);          // finite: every visited tile has >= 1 valid key
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) { s[a][c] = expf(s[a][c] - m_new); rs += s[a][c]; }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      scale_o[a] = expf(m_run[a] - m_new);              // exp(-inf) = 0 on the first tile
      l_run[a] = l_run[a] * scale_o[a] + rs;
      m_run[a] = m_new;
    }
    __syncthreads();                      // all S reads of Kt done -> overwrite with P^T
#pragma unroll
    for (int c = 0; c < 4; ++c)
      *reinterpret_cast<float4*>(&Kt[(tx * 4 + c) * AT_LD + ty * 4]) = make_float4(s[0][c], s[1][c], s[2][c], s[3][c]);
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) o[a][c] *= scale_o[a];
#pragma unroll 8
    for (int j = 0; j < AT_BK; ++j) {
      const float4 pa = *reinterpret_cast<const float4*>(&Kt[j * AT_LD + ty * 4]);
      const float4 vc = *reinterpret_cast<const float4*>(&Vs[j * AT_D + tx * 4]);
      const float pv[4] = {pa.x, pa.y, pa.z, pa.w}, vv[4] = {vc.x, vc.y, vc.z, vc.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) o[a][c] = fmaf(pv[a], vv[c], o[a][c]);
    }
  }
  T* ob = out + q_row0 * ldo + h * AT_D + tx * 4;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int r = q0 + ty * 4 + a;
    if (r >= n_q) continue;
    const float inv = l_run[a] > 0.f ? 1.0f / l_run[a] : 0.f;
    store4(ob + (long long)r * ldo, make_float4(o[a][0] * inv, o[a][1] * inv, o[a][2] * inv, o[a][3] * inv));
  }
}
// ---- memory cross-attention: M learned queries (M = 16 / 64, w2v2_transformer_interlingua.py:262-274) over ALL key rows ----
// The generic kernel above tiles 64 queries x 64 keys; with 16 queries three quarters of every tile is padding and
// the key loop is a serial chain of small tiles (35-45 us per call, measured).  Here one CTA owns (utterance, head,
// block of 16 queries) and the whole key axis at once:
//   1. thread = key row: its 64-dim K row goes to registers once and meets the 16 queries (shared memory, broadcast)
//      -> 16 scores per thread, written to S[16][n_kv] in shared memory;
//   2. warp = query row (2 rows per warp): max / exp / sum over the key axis with shuffles (fp32, expf);
//   3. thread = (query, 4 output dims): P[m][:] . V[:, d] with V rows read once per CTA through L1.
// Keys beyond kv_len[b] are masked (the memory stage passes no mask: the reference hands its layers an all-False
// key-padding mask there).  Scores live in dynamic shared memory: n_kv <= MA_MAX_KV, else the generic kernel runs.
constexpr int MA_Q = 16, MA_THREADS = 256, MA_MAX_KV = 2048, MA_VCH = 64;

template <typename T>
__global__ void __launch_bounds__(MA_THREADS) memory_attention_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                                      const T* __restrict__ v, T* __restrict__ out,
                                                                      long long ldq, long long ldkv, long long ldo,
                                                                      int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg,
                                                                      const int32_t* __restrict__ kv_len, const int4* __restrict__ seg) {
  extern __shared__ float sm[];
  float* Qs = sm;                          // [16][64]
  float* inv_l = sm + MA_Q * AT_D;         // [16]
  float* Vs = inv_l + MA_Q;                // [64 key rows][64] staging for step 3
  float* S = Vs + MA_VCH * AT_D;           // [16][n_kv_pad]
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * MA_Q;
  long long q_row0 = (long long)b * q_rows_per_seg, kv_row0 = (long long)b * kv_rows_per_seg;
  if (seg != nullptr) { const int4 sg = seg[b]; q_row0 = sg.x; n_q = sg.y; kv_row0 = sg.z; n_kv = sg.w; }   // n_kv <= launch maximum
  if (q0 >= n_q) return;
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ldS = (n_kv + 3) & ~3;
  int klen = kv_len ? kv_len[b] : n_kv;
  if (klen > n_kv) klen = n_kv;
  const T* qb = q + q_row0 * ldq + h * AT_D;
  const T* kb = k + kv_row0 * ldkv + h * AT_D;
  const T* vb = v + kv_row0 * ldkv + h * AT_D;
  {
    const int m = tid >> 4, d = (tid & 15) * 4;                  // 256 threads = 16 rows x 16 float4
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + m < n_q) t = load4(qb + (long long)(q0 + m) * ldq + d);
    *reinterpret_cast<float4*>(&Qs[m * AT_D + d]) = t;
  }
  __syncthreads();
  // 1. scores
  for (int j = tid; j < n_kv; j += MA_THREADS) {
    float sc[MA_Q];
#pragma unroll
    for (int m = 0; m < MA_Q; ++m) sc[m] = 0.f;
    if (j < klen) {
      const T* kr = kb + (long long)j * ldkv;
#pragma unroll
      for (int d0 = 0; d0 < AT_D; d0 += 16) {                    // 16 dims of the key row at a time (register budget)
        float4 kk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) kk[i] = load4(kr + d0 + 4 * i);
#pragma unroll
        for (int m = 0; m < MA_Q; ++m) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 qq = *reinterpret_cast<const float4*>(&Qs[m * AT_D + d0 + 4 * i]);   // same address in every lane
            sc[m] = fmaf(qq.x, kk[i].x, fmaf(qq.y, kk[i].y, fmaf(qq.z, kk[i].z, fmaf(qq.w, kk[i].w, sc[m]))));
          }
        }
      }
    } else {
#pragma unroll
      for (int m = 0; m < MA_Q; ++m) sc[m] = -INFINITY;
    }
#pragma unroll
    for (int m = 0; m < MA_Q; ++m) S[m * ldS + j] = sc[m];
  }
  __syncthreads();
  // 2. softmax over the key axis, two query rows per warp
#pragma unroll
  for (int mm = 0; mm < 2; ++mm) {
    const int m = warp * 2 + mm;
    float* row = S + m * ldS;
    float mx = -INFINITY;
    for (int j = lane; j < n_kv; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < n_kv; j += 32) {
      const float e = (mx == -INFINITY) ? 0.f : expf(row[j] - mx);
      row[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) inv_l[m] = sum > 0.f ? 1.0f / sum : 0.f;
  }
  __syncthreads();
  // 3. P V: thread = (query m, dims d..d+3).  V is staged 64 key rows at a time through shared memory: four independent
  //    row loads per thread in flight (a serial load -> FMA loop over the keys was latency-bound: 15 us of a 29 us call)
  {
    const int m = tid >> 4, d = (tid & 15) * 4;
    const float* row = S + m * ldS;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j0 = 0; j0 < klen; j0 += MA_VCH) {
      float4 vv[MA_VCH / 16];
#pragma unroll
      for (int i = 0; i < MA_VCH / 16; ++i) {
        const int j = j0 + m + 16 * i;
        vv[i] = (j < klen) ? load4(vb + (long long)j * ldkv + d) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();                                            // previous chunk consumed
#pragma unroll
      for (int i = 0; i < MA_VCH / 16; ++i) *reinterpret_cast<float4*>(&Vs[(m + 16 * i) * AT_D + d]) = vv[i];
      __syncthreads();
      const int n = min(MA_VCH, klen - j0);
#pragma unroll 8
      for (int jj = 0; jj < n; ++jj) {
        const float pj = row[j0 + jj];
        const float4 v4 = *reinterpret_cast<const float4*>(&Vs[jj * AT_D + d]);
        o.x = fmaf(pj, v4.x, o.x); o.y = fmaf(pj, v4.y, o.y); o.z = fmaf(pj, v4.z, o.z); o.w = fmaf(pj, v4.w, o.w);
      }
    }
    if (q0 + m < n_q) {
      const float inv = inv_l[m];
      store4(out + (q_row0 + q0 + m) * ldo + h * AT_D + d, make_float4(o.x * inv, o.y * inv, o.z * inv, o.w * inv));
    }
  }
}

template <typename T>
static int launch_memory_attention(const T* q, const T* k, const T* v, T* out, long long ldq, long long ldkv, long long ldo,
                                   int B, int H, int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg,
                                   const int32_t* kv_len, const int32_t* seg, cudaStream_t st) {
  const int smem = (MA_Q * AT_D + MA_Q + MA_VCH * AT_D + MA_Q * ((n_kv + 3) & ~3)) * 4;
  static int attr_bytes = 0;
  if (smem > attr_bytes) {
    CST_CHECK_CUDA(cudaFuncSetAttribute(memory_attention_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (MA_Q * AT_D + MA_Q + MA_VCH * AT_D + MA_Q * MA_MAX_KV) * 4));
    attr_bytes = (MA_Q * AT_D + MA_Q + MA_VCH * AT_D + MA_Q * MA_MAX_KV) * 4;
  }
  dim3 grid(cdiv(n_q, MA_Q), H, B);
  CST_CHECK_CUDA(launch_k(memory_attention_kernel<T>, grid, dim3(MA_THREADS), smem, st, q, k, v, out, ldq, ldkv, ldo, n_q, q_rows_per_seg,
                          n_kv, kv_rows_per_seg, kv_len, reinterpret_cast<const int4*>(seg)));
  return CST_OK;
}

int launch_attention_tc(const void* q, const void* k, const void* v, void* out, long long ldq, long long ldkv, long long ldo,
                        int B, int H, int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg,
                        const int32_t* kv_len, const int32_t* seg, cudaStream_t st);

int launch_attention_tc2(const void* q, const void* k, const void* v, void* out, long long ldq, long long ldkv, long long ldo,
                         int B, int H, int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg,
                         const int32_t* kv_len, const int32_t* seg, cudaStream_t st);

// seg == nullptr: B uniform segments.  seg != nullptr (int32 [B][4] = {first q row, q rows, first kv row, kv rows}):
// n_q / n_kv are the maxima over the table and q_rows_per_seg / kv_rows_per_seg the total rows of the q / kv buffers.
static int attention_dispatch(const void* q, const void* k, const void* v, void* out, int dtype,
                              long long ldq, long long ldkv, long long ldo,
                              int B, int H, int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg,
                              const int32_t* kv_len, const int32_t* seg, cudaStream_t st) {
  const int4* seg4 = reinterpret_cast<const int4*>(seg);
  // bf16 with enough query rows to fill a 128-row MMA tile: tensor-core kernel (attention_tc.cu);
  // fp32 (exact-parity mode) and the M-query memory stage: the CUDA-core kernel above.
  // CST_ATTN_V2=1 selects the second-generation kernel (attention_tc2.cu: P in tensor memory, persistent CTAs, two query
  // tiles per CTA).  Measured on B200 (profiles/SUMMARY_r02.md): both kernels are bound by the per-instruction floor of
  // the small MMAs attention issues at head_dim 64 (tools/micro/mma_attn_rate.cu), and v2 is not ahead yet -> default 0.
  if (dtype == CST_BF16 && n_q > 64) {
    static const int v2 = [] { const char* e = getenv("CST_ATTN_V2"); return e ? atoi(e) : 0; }();
    if (v2) {
      const int rc = launch_attention_tc2(q, k, v, out, ldq, ldkv, ldo, B, H, n_q, q_rows_per_seg, n_kv, kv_rows_per_seg, kv_len, seg, st);
      if (rc != CST_ERR_UNSUPPORTED) return rc;
    }
    return launch_attention_tc(q, k, v, out, ldq, ldkv, ldo, B, H, n_q, q_rows_per_seg, n_kv, kv_rows_per_seg, kv_len, seg, st);
  }
  // few queries over a whole key axis (the memory stage): dedicated kernel; CST_MEMATTN=0 keeps the generic one
  static const int mem_attn = [] { const char* e = getenv("CST_MEMATTN"); return e ? atoi(e) : 1; }();
  if (mem_attn && n_q <= 64 && n_kv <= MA_MAX_KV && (dtype == CST_BF16 || dtype == CST_F32)) {
    if (dtype == CST_F32)
      return launch_memory_attention((const float*)q, (const float*)k, (const float*)v, (float*)out, ldq, ldkv, ldo, B, H, n_q, q_rows_per_seg,
                                     n_kv, kv_rows_per_seg, kv_len, seg, st);
    return launch_memory_attention((const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, (__nv_bfloat16*)out, ldq, ldkv, ldo,
                                   B, H, n_q, q_rows_per_seg, n_kv, kv_rows_per_seg, kv_len, seg, st);
  }
  dim3 grid(cdiv(n_q, AT_BQ), H, B);
  static bool attr[2] = {false, false};
  if (dtype == CST_F32) {
    if (!attr[0]) { CST_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM)); attr[0] = true; }
    CST_CHECK_CUDA(launch_k(attention_kernel<float>, grid, dim3(256), AT_SMEM, st, (const float*)q, (const float*)k, (const float*)v, (float*)out,
                            ldq, ldkv, ldo, n_q, q_rows_per_seg, n_kv, kv_rows_per_seg, kv_len, seg4));
  } else if (dtype == CST_BF16) {
    if (!attr[1]) { CST_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM)); attr[1] = true; }
    CST_CHECK_CUDA(launch_k(attention_kernel<__nv_bfloat16>, grid, dim3(256), AT_SMEM, st, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v,
                            (__nv_bfloat16*)out, ldq, ldkv, ldo, n_q, q_rows_per_seg, n_kv, kv_rows_per_seg, kv_len, seg4));
  } else {
    CST_REQUIRE(false, "cst_attention: bad dtype %d", dtype);
  }
  CST_LAUNCH_CHECK();
  return CST_OK;
}
}  // namespace cst

extern "C" int cst_attention(const void* q, const void* k, const void* v, void* out, int dtype,
                             long long ldq, long long ldkv, long long ldo,
                             int B, int H, int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg,
                             const int32_t* kv_len, void* stream) {
  using namespace cst;
  CST_REQUIRE(q && k && v && out && B > 0 && H > 0 && n_q > 0 && n_kv > 0, "cst_attention: bad args");
  CST_REQUIRE(n_q <= q_rows_per_seg && n_kv <= kv_rows_per_seg, "cst_attention: n_q/n_kv exceed rows per segment");
  CST_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0, "cst_attention: leading dims must be multiples of 8");
  CST_REQUIRE(H <= 65535 && B <= 65535, "cst_attention: grid too large");
  return attention_dispatch(q, k, v, out, dtype, ldq, ldkv, ldo, B, H, n_q, q_rows_per_seg, n_kv, kv_rows_per_seg, kv_len, nullptr,
                            (cudaStream_t)stream);
}

extern "C" int cst_attention_segs(const void* q, const void* k, const void* v, void* out, int dtype,
                                  long long ldq, long long ldkv, long long ldo, int B, int H,
                                  const int32_t* seg, int max_n_q, int max_n_kv, long long q_rows_total, long long kv_rows_total,
                                  const int32_t* kv_len, void* stream) {
  using namespace cst;
  CST_REQUIRE(q && k && v && out && seg && B > 0 && H > 0 && max_n_q > 0 && max_n_kv > 0, "cst_attention_segs: bad args");
  CST_REQUIRE(q_rows_total > 0 && kv_rows_total > 0 && q_rows_total < (1ll << 31) && kv_rows_total < (1ll << 31),
              "cst_attention_segs: bad buffer extents");
  CST_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0, "cst_attention_segs: leading dims must be multiples of 8");
  CST_REQUIRE(((uintptr_t)seg % 16) == 0, "cst_attention_segs: the segment table must be 16-byte aligned");
  CST_REQUIRE(H <= 65535 && B <= 65535, "cst_attention_segs: grid too large");
  return attention_dispatch(q, k, v, out, dtype, ldq, ldkv, ldo, B, H, max_n_q, (int)q_rows_total, max_n_kv, (int)kv_rows_total, kv_len,
                            seg, (cudaStream_t)stream);
}
