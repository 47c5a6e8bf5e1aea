// f.4 (16-bit mode): attention backward on the tensor cores as five batched tcgen05 GEMMs around one softmax-backward kernel.
// Training shapes are short (T' <= 1499 frames), so the T x T score matrices of one layer are a transient of a few hundred MB and
// every step below is HBM- or tensor-bound instead of the CUDA-core FFMA loop of attention_bwd.cu (which stays as the fp32 parity
// kernel).  Arithmetic: the derivative of cst_attention (F.multi_head_attention_forward as called from
// fairseq/modules/multihead_attention.py:155-187; q pre-scaled, keys >= kv_len[b] masked, all query rows live):
//   S = Q K^T, dP = dO V^T                                  (2 GEMMs, K = 64, fp32 out)
//   P = softmax(S + mask), D = rowsum(P o dP), dS = P o (dP - D)      (one kernel; writes dS, dS^T, P^T in bf16)
//   dQ = dS K, dK = dS^T Q, dV = P^T dO                      (3 GEMMs, N = 64, reduction over the padded key / query axis)
// Operands are re-laid per (utterance, head) as dense [T, 64] / [64, T] bf16 panels by head_pack_kernel; the three dense fp32
// results are scattered back into the strided dq / dk / dv buffers by head_unpack_kernel.
//
// Dropout of the attention probabilities (the training recipe's attention_dropout; F.multi_head_attention_forward applies F.dropout
// to softmax(S) before the product with V): with p > 0 the FORWARD runs here too (cst_attention_dropout_fwd: S GEMM, one
// softmax + dropout kernel writing Pd = P o keep / (1 - p), Pd V GEMM) and the backward uses the same masks:
//   dV = Pd^T dO,   dP = (dO V^T) o keep / (1 - p),   D = rowsum(P o dP),   dS = P o (dP - D)
// keep(b, h, i, j) = Philox word of element ((b*H + h)*Tqp + i)*Tkp + j (philox.cuh), regenerated, never stored.
#include "common.cuh"
#include "philox.cuh"

namespace cst {

__device__ __forceinline__ float hb_ld(const void* p, int dt, long long i) {
  return dt == CST_F32 ? reinterpret_cast<const float*>(p)[i] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}

// src [B*rows_per_seg, ld] (dtype sdt), head h at columns h*64 .. h*64+63  ->  dst [B*H][Tp][64] bf16 (rows t >= valid zero) and
// dstT [B*H][64][Tp] (optional).  valid = min(n, len[b]).  grid (Tp/64, H, B), 256 threads.
__global__ void __launch_bounds__(256) head_pack_kernel(const void* __restrict__ src, int sdt, long long ld, int rows_per_seg, int n,
                                                        const int32_t* __restrict__ len, int H, int Tp,
                                                        __nv_bfloat16* __restrict__ dst, __nv_bfloat16* __restrict__ dstT) {
  __shared__ float tile[64][65];
  pdl_launch_dependents();
  pdl_wait();
  const int t0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  int valid = n;
  if (len != nullptr) valid = min(valid, len[b]);
  const long long bh = (long long)b * H + h;
  // 4 consecutive head columns per thread: 8-byte (bf16) / 16-byte (fp32) loads, 8-byte stores (ld and the head offset are multiples of 64)
  for (int e = threadIdx.x; e < 64 * 16; e += 256) {
    const int i = e >> 4, d = (e & 15) * 4;
    const int t = t0 + i;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < valid) {
      const long long off = ((long long)b * rows_per_seg + t) * ld + h * 64 + d;
      v = sdt == CST_F32 ? load4(reinterpret_cast<const float*>(src) + off) : load4(reinterpret_cast<const __nv_bfloat16*>(src) + off);
    }
    tile[i][d] = v.x; tile[i][d + 1] = v.y; tile[i][d + 2] = v.z; tile[i][d + 3] = v.w;
    store4(dst + (bh * Tp + t) * 64 + d, v);
  }
  if (dstT == nullptr) return;
  __syncthreads();
  for (int e = threadIdx.x; e < 64 * 16; e += 256) {               // transposed panel: 4 consecutive rows t of one head column
    const int d = e >> 4, i = (e & 15) * 4;
    store4(dstT + (bh * 64 + d) * Tp + t0 + i, make_float4(tile[i][d], tile[i + 1][d], tile[i + 2][d], tile[i + 3][d]));
  }
}

// dense [B*H][Tp][64] fp32 -> out[(b*rows_per_seg + t)*ld + h*64 + d] for t < n.  grid (cdiv(n, 16), H, B), 256 threads.
__global__ void __launch_bounds__(256) head_unpack_kernel(const float* __restrict__ src, int Tp, int n, int H, float* __restrict__ out,
                                                          long long ld, int rows_per_seg) {
  pdl_launch_dependents();
  pdl_wait();
  const int h = blockIdx.y, b = blockIdx.z;
  const int t = blockIdx.x * 16 + (threadIdx.x >> 4), d = (threadIdx.x & 15) * 4;
  if (t >= n) return;
  const float4 v = load4(src + (((long long)b * H + h) * Tp + t) * 64 + d);
  store4(out + ((long long)b * rows_per_seg + t) * ld + h * 64 + d, v);
}
// the same scatter into a bf16 tensor (the attention output of the 16-bit training forward)
__global__ void __launch_bounds__(256) head_unpack_bf16_kernel(const float* __restrict__ src, int Tp, int n, int H,
                                                               __nv_bfloat16* __restrict__ out, long long ld, int rows_per_seg) {
  pdl_launch_dependents();
  pdl_wait();
  const int h = blockIdx.y, b = blockIdx.z;
  const int t = blockIdx.x * 16 + (threadIdx.x >> 4), d = (threadIdx.x & 15) * 4;
  if (t >= n) return;
  const float4 v = load4(src + (((long long)b * H + h) * Tp + t) * 64 + d);
  store4(out + ((long long)b * rows_per_seg + t) * ld + h * 64 + d, v);
}

// keep / (1 - p) factors of 4 consecutive score elements e .. e+3 (e % 4 == 0); seed == nullptr: no dropout
struct Keep4 { float x, y, z, w; };
__device__ __forceinline__ Keep4 keep4(const unsigned long long* seed, uint32_t site, long long e, uint32_t thresh, float scale) {
  if (seed == nullptr) return Keep4{1.f, 1.f, 1.f, 1.f};
  const Philox4 r = philox4x32_10(*seed, (unsigned long long)e >> 2, site);
  return Keep4{r.x >= thresh ? scale : 0.f, r.y >= thresh ? scale : 0.f, r.z >= thresh ? scale : 0.f, r.w >= thresh ? scale : 0.f};
}

// Forward softmax + dropout: Pd[bh][r][c] = softmax_c(S[bh][r][c]) * keep / (1 - p) in bf16 (columns >= klen and rows >= n_q: 0).
// One CTA = 32 query rows of one (utterance, head); a warp computes max and sum of 4 rows, then every thread forms 8 columns.
__global__ void __launch_bounds__(256) attn_fwd_softmax_dropout_kernel(const float* __restrict__ S, int Tqp, int Tkp, int n_q, int n_kv,
                                                                       const int32_t* __restrict__ kv_len, int H,
                                                                       __nv_bfloat16* __restrict__ Pd,
                                                                       const unsigned long long* __restrict__ seed, uint32_t site,
                                                                       uint32_t thresh, float scale) {
  __shared__ float st_m[32], st_il[32];
  pdl_launch_dependents();
  pdl_wait();
  const int r0 = blockIdx.x * 32;
  const long long bh = blockIdx.y;
  const int b = (int)(bh / H);
  int klen = n_kv;
  if (kv_len != nullptr) klen = min(klen, kv_len[b]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* Sb = S + bh * (long long)Tqp * Tkp;
  for (int rr = 0; rr < 4; ++rr) {
    const int r = warp * 4 + rr;
    const float* srow = Sb + (long long)(r0 + r) * Tkp;
    float mx = -INFINITY;
    for (int c = lane * 4; c < klen; c += 128) {
      const float4 s4 = load4(srow + c);
      mx = fmaxf(mx, s4.x);
      if (c + 1 < klen) mx = fmaxf(mx, s4.y);
      if (c + 2 < klen) mx = fmaxf(mx, s4.z);
      if (c + 3 < klen) mx = fmaxf(mx, s4.w);
    }
    mx = warp_max(mx);
    float l = 0.f;
    for (int c = lane * 4; c < klen; c += 128) {
      const float4 s4 = load4(srow + c);
      l += (expf(s4.x - mx) + (c + 1 < klen ? expf(s4.y - mx) : 0.f)) + ((c + 2 < klen ? expf(s4.z - mx) : 0.f) + (c + 3 < klen ? expf(s4.w - mx) : 0.f));
    }
    l = warp_sum(l);
    if (lane == 0) {
      const bool live = (r0 + r) < n_q && l > 0.f;
      st_m[r] = mx; st_il[r] = live ? 1.0f / l : 0.f;
    }
  }
  __syncthreads();
  __nv_bfloat16* Pb = Pd + bh * (long long)Tqp * Tkp;
  const int r = threadIdx.x >> 3, cc = (threadIdx.x & 7) * 8;
  const float m = st_m[r], il = st_il[r];
  for (int c0 = 0; c0 < Tkp; c0 += 64) {
    const long long e = (bh * Tqp + r0 + r) * (long long)Tkp + c0 + cc;
    const float* srow = Sb + (long long)(r0 + r) * Tkp + c0 + cc;
    const float4 s0 = load4(srow), s1 = load4(srow + 4);
    const Keep4 k0 = keep4(seed, site, e, thresh, scale), k1 = keep4(seed, site, e + 4, thresh, scale);
    const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const float kv[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
    float pv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pv[i] = (c0 + cc + i < klen) ? expf(sv[i] - m) * il * kv[i] : 0.f;
    uint4 u;
    u.x = pack_bf16x2(pv[0], pv[1]); u.y = pack_bf16x2(pv[2], pv[3]); u.z = pack_bf16x2(pv[4], pv[5]); u.w = pack_bf16x2(pv[6], pv[7]);
    *reinterpret_cast<uint4*>(Pb + (long long)(r0 + r) * Tkp + c0 + cc) = u;
  }
}

// One CTA = 32 query rows of one (utterance, head).  Phase A: a warp owns 4 rows and computes max, sum of exponentials and
// D = sum_j P_ij dP_ij (two sweeps over the row; the rows just came out of the GEMMs and are re-read through L1/L2).  Phase B: per
// 64-key chunk all threads form P and dS, store dS row-major and stage both tiles in shared memory for the transposed stores
// (32 consecutive query rows per key = 64-byte segments).
__global__ void __launch_bounds__(256) attn_bwd_softmax_kernel(const float* __restrict__ S, const float* __restrict__ dP, int Tqp, int Tkp,
                                                               int n_q, int n_kv, const int32_t* __restrict__ kv_len, int H,
                                                               __nv_bfloat16* __restrict__ dS, __nv_bfloat16* __restrict__ dST,
                                                               __nv_bfloat16* __restrict__ PT,
                                                               const unsigned long long* __restrict__ seed, uint32_t site, uint32_t thresh,
                                                               float scale) {
  // with dropout (seed != nullptr): `dP` holds d(Pd) = dO V^T; dP = d(Pd) o keep/(1-p); PT receives Pd^T (the dV operand)
  __shared__ float st_m[32], st_il[32], st_d[32];
  __shared__ float tp[32][65], tds[32][65];
  pdl_launch_dependents();
  pdl_wait();
  const int r0 = blockIdx.x * 32;
  const long long bh = blockIdx.y;
  const int b = (int)(bh / H);
  int klen = n_kv;
  if (kv_len != nullptr) klen = min(klen, kv_len[b]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* Sb = S + bh * (long long)Tqp * Tkp;
  const float* Pb = dP + bh * (long long)Tqp * Tkp;
  for (int rr = 0; rr < 4; ++rr) {
    const int r = warp * 4 + rr;
    const float* srow = Sb + (long long)(r0 + r) * Tkp;
    const float* prow = Pb + (long long)(r0 + r) * Tkp;
    // 4 consecutive columns per lane (16-byte loads, one Philox counter per quad); rows are padded to Tkp >= klen (multiple of 64)
    float mx = -INFINITY;
    for (int c = lane * 4; c < klen; c += 128) {
      const float4 s4 = load4(srow + c);
      mx = fmaxf(mx, s4.x);
      if (c + 1 < klen) mx = fmaxf(mx, s4.y);
      if (c + 2 < klen) mx = fmaxf(mx, s4.z);
      if (c + 3 < klen) mx = fmaxf(mx, s4.w);
    }
    mx = warp_max(mx);
    float l = 0.f, dacc = 0.f;
    const long long e_row = (bh * Tqp + r0 + r) * (long long)Tkp;
    for (int c = lane * 4; c < klen; c += 128) {
      const float4 s4 = load4(srow + c), g4 = load4(prow + c);
      const Keep4 k4 = keep4(seed, site, e_row + c, thresh, scale);
      const float e0 = expf(s4.x - mx), e1 = c + 1 < klen ? expf(s4.y - mx) : 0.f, e2 = c + 2 < klen ? expf(s4.z - mx) : 0.f,
                  e3 = c + 3 < klen ? expf(s4.w - mx) : 0.f;
      l += (e0 + e1) + (e2 + e3);
      dacc = fmaf(e0, g4.x * k4.x, dacc); dacc = fmaf(e1, g4.y * k4.y, dacc);
      dacc = fmaf(e2, g4.z * k4.z, dacc); dacc = fmaf(e3, g4.w * k4.w, dacc);
    }
    l = warp_sum(l); dacc = warp_sum(dacc);
    if (lane == 0) {
      const bool live = (r0 + r) < n_q && l > 0.f;
      st_m[r] = mx; st_il[r] = live ? 1.0f / l : 0.f; st_d[r] = live ? dacc / l : 0.f;
    }
  }
  __syncthreads();
  __nv_bfloat16* dSb = dS + bh * (long long)Tqp * Tkp;
  __nv_bfloat16* dSTb = dST + bh * (long long)Tkp * Tqp;
  __nv_bfloat16* PTb = PT + bh * (long long)Tkp * Tqp;
  for (int c0 = 0; c0 < Tkp; c0 += 64) {
    // 32 x 64 tile: thread -> row (tid / 8), 8 consecutive columns
    {
      const int r = threadIdx.x >> 3, cc = (threadIdx.x & 7) * 8;
      const float m = st_m[r], il = st_il[r], dd = st_d[r];
      const float* srow = Sb + (long long)(r0 + r) * Tkp + c0 + cc;
      const float* prow = Pb + (long long)(r0 + r) * Tkp + c0 + cc;
      const float4 s0 = load4(srow), s1 = load4(srow + 4), g0 = load4(prow), g1 = load4(prow + 4);
      const long long e = (bh * Tqp + r0 + r) * (long long)Tkp + c0 + cc;
      const Keep4 k0 = keep4(seed, site, e, thresh, scale), k1 = keep4(seed, site, e + 4, thresh, scale);
      const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
      const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float kv[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
      float pv[8], dv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float p = (c0 + cc + i < klen) ? expf(sv[i] - m) * il : 0.f;
        pv[i] = p;
        dv[i] = p * (gv[i] * kv[i] - dd);
        tp[r][cc + i] = p * kv[i];
        tds[r][cc + i] = dv[i];
      }
      uint4 u;
      u.x = pack_bf16x2(dv[0], dv[1]); u.y = pack_bf16x2(dv[2], dv[3]); u.z = pack_bf16x2(dv[4], dv[5]); u.w = pack_bf16x2(dv[6], dv[7]);
      *reinterpret_cast<uint4*>(dSb + (long long)(r0 + r) * Tkp + c0 + cc) = u;
    }
    __syncthreads();
    {
      const int c = threadIdx.x >> 2, rg = (threadIdx.x & 3) * 8;          // key column, 8 consecutive query rows
      uint4 u, w;
      u.x = pack_bf16x2(tp[rg][c], tp[rg + 1][c]); u.y = pack_bf16x2(tp[rg + 2][c], tp[rg + 3][c]);
      u.z = pack_bf16x2(tp[rg + 4][c], tp[rg + 5][c]); u.w = pack_bf16x2(tp[rg + 6][c], tp[rg + 7][c]);
      w.x = pack_bf16x2(tds[rg][c], tds[rg + 1][c]); w.y = pack_bf16x2(tds[rg + 2][c], tds[rg + 3][c]);
      w.z = pack_bf16x2(tds[rg + 4][c], tds[rg + 5][c]); w.w = pack_bf16x2(tds[rg + 6][c], tds[rg + 7][c]);
      *reinterpret_cast<uint4*>(PTb + (long long)(c0 + c) * Tqp + r0 + rg) = u;
      *reinterpret_cast<uint4*>(dSTb + (long long)(c0 + c) * Tqp + r0 + rg) = w;
    }
    __syncthreads();
  }
}

static inline long long up64(long long x) { return (x + 63) / 64 * 64; }

}  // namespace cst

using namespace cst;

extern "C" long long cst_attention_bwd_tc_ws_bytes(int B, int H, int n_q, int n_kv) {
  const long long BH = (long long)B * H, Tqp = up64(n_q), Tkp = up64(n_kv);
  return BH * (Tqp * 64 * 2 * 4 + Tkp * 64 * 2 * 3          // Qh QhT Gh GhT | Kh KhT Vh
               + Tqp * Tkp * (4 + 4 + 2 + 2 + 2)           // S dP | dS dST PT
               + Tqp * 64 * 4 + Tkp * 64 * 4 * 2) + 4096;  // dQd | dKd dVd
}

static int batched_gemm(const void* A, const void* W, void* Cout, int c_dtype, int M, int N, int K, long long a_rows, int nz, void* stream) {
  cst_gemm_params p;
  memset(&p, 0, sizeof(p));
  p.A = A; p.W = W; p.C = Cout;
  p.ab_dtype = CST_BF16; p.c_dtype = c_dtype;
  p.M = M; p.N = N; p.K = K; p.lda = K; p.ldc = N; p.a_rows = a_rows;
  p.act = CST_ACT_NONE; p.alpha = 1.0f;
  p.nb_outer = 1; p.nb_inner = nz;
  p.a_bs_inner = (long long)M * K; p.w_bs_inner = (long long)N * K; p.c_bs_inner = (long long)M * N;
  p.rows_per_seg = M; p.seg_rows_valid = M; p.out_rows_per_seg = M; p.segs_per_outer = 1;
  return cst_gemm(&p, stream);
}

extern "C" int cst_attention_bwd_tc_dropout(const void* q, const void* k, const void* v, int qkv_dtype, const float* d_o, float* dq, float* dk,
                                            float* dv, long long ldq, long long ldkv, long long ldo, long long lddq, long long lddkv,
                                            int B, int H, int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg, const int32_t* kv_len,
                                            float p, const unsigned long long* seed, unsigned int site, void* ws, void* stream);

// q / k / v: bf16 forward tensors (row strides ldq / ldkv); d_o fp32 (ldo); dq / dk / dv fp32 (lddq / lddkv), rows t < n_q / n_kv of
// every utterance are written.  ws: cst_attention_bwd_tc_ws_bytes(B, H, n_q, n_kv) bytes, 256-byte aligned.
extern "C" int cst_attention_bwd_tc(const void* q, const void* k, const void* v, const float* d_o, float* dq, float* dk, float* dv,
                                    long long ldq, long long ldkv, long long ldo, long long lddq, long long lddkv,
                                    int B, int H, int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg, const int32_t* kv_len,
                                    void* ws, void* stream) {
  return cst_attention_bwd_tc_dropout(q, k, v, CST_BF16, d_o, dq, dk, dv, ldq, ldkv, ldo, lddq, lddkv, B, H, n_q, q_rows_per_seg, n_kv,
                                      kv_rows_per_seg, kv_len, 0.f, nullptr, 0u, ws, stream);
}

// The same with dropout of the attention probabilities (p > 0: masks of (*seed, site), see the header of this file) and q / k / v in
// `qkv_dtype` (CST_BF16 or CST_F32: the panels are bf16 either way).
extern "C" int cst_attention_bwd_tc_dropout(const void* q, const void* k, const void* v, int qkv_dtype, const float* d_o, float* dq, float* dk,
                                            float* dv, long long ldq, long long ldkv, long long ldo, long long lddq, long long lddkv,
                                            int B, int H, int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg, const int32_t* kv_len,
                                            float p, const unsigned long long* seed, unsigned int site, void* ws, void* stream) {
  CST_REQUIRE(q && k && v && d_o && dq && dk && dv && ws && B > 0 && H > 0 && n_q > 0 && n_kv > 0, "cst_attention_bwd_tc: bad args");
  CST_REQUIRE((qkv_dtype == CST_BF16 || qkv_dtype == CST_F32) && p >= 0.f && p < 1.f && (p == 0.f || seed != nullptr),
              "cst_attention_bwd_tc: bad dtype / dropout arguments");
  const unsigned long long* sd = p > 0.f ? seed : nullptr;
  const uint32_t thresh = dropout_threshold(p);
  const float kscale = 1.0f / (1.0f - p);
  CST_REQUIRE(n_q <= q_rows_per_seg && n_kv <= kv_rows_per_seg && ((uintptr_t)ws % 256) == 0 && lddq % 4 == 0 && lddkv % 4 == 0,
              "cst_attention_bwd_tc: bad geometry");
  cudaStream_t st = (cudaStream_t)stream;
  const long long BH = (long long)B * H;
  const int Tqp = (int)up64(n_q), Tkp = (int)up64(n_kv);
  uint8_t* w = reinterpret_cast<uint8_t*>(ws);
  auto take = [&](long long bytes) { uint8_t* p = w; w += (bytes + 255) / 256 * 256; return p; };
  auto* Qh = (__nv_bfloat16*)take(BH * Tqp * 64 * 2); auto* QhT = (__nv_bfloat16*)take(BH * Tqp * 64 * 2);
  auto* Gh = (__nv_bfloat16*)take(BH * Tqp * 64 * 2); auto* GhT = (__nv_bfloat16*)take(BH * Tqp * 64 * 2);
  auto* Kh = (__nv_bfloat16*)take(BH * Tkp * 64 * 2); auto* KhT = (__nv_bfloat16*)take(BH * Tkp * 64 * 2);
  auto* Vh = (__nv_bfloat16*)take(BH * Tkp * 64 * 2);
  auto* S = (float*)take(BH * Tqp * Tkp * 4); auto* dP = (float*)take(BH * Tqp * Tkp * 4);
  auto* dS = (__nv_bfloat16*)take(BH * Tqp * Tkp * 2); auto* dST = (__nv_bfloat16*)take(BH * Tqp * Tkp * 2);
  auto* PT = (__nv_bfloat16*)take(BH * Tqp * Tkp * 2);
  auto* dQd = (float*)take(BH * Tqp * 64 * 4); auto* dKd = (float*)take(BH * Tkp * 64 * 4); auto* dVd = (float*)take(BH * Tkp * 64 * 4);
  CST_REQUIRE((long long)(w - reinterpret_cast<uint8_t*>(ws)) <= cst_attention_bwd_tc_ws_bytes(B, H, n_q, n_kv) + 16 * 256,
              "cst_attention_bwd_tc: workspace accounting");
  CST_CHECK_CUDA(launch_k(head_pack_kernel, dim3(Tqp / 64, H, B), dim3(256), 0, st, q, qkv_dtype, ldq, q_rows_per_seg, n_q,
                          (const int32_t*)nullptr, H, Tqp, Qh, QhT));
  CST_CHECK_CUDA(launch_k(head_pack_kernel, dim3(Tqp / 64, H, B), dim3(256), 0, st, (const void*)d_o, (int)CST_F32, ldo, q_rows_per_seg, n_q,
                          (const int32_t*)nullptr, H, Tqp, Gh, GhT));
  CST_CHECK_CUDA(launch_k(head_pack_kernel, dim3(Tkp / 64, H, B), dim3(256), 0, st, k, qkv_dtype, ldkv, kv_rows_per_seg, n_kv, kv_len, H,
                          Tkp, Kh, KhT));
  CST_CHECK_CUDA(launch_k(head_pack_kernel, dim3(Tkp / 64, H, B), dim3(256), 0, st, v, qkv_dtype, ldkv, kv_rows_per_seg, n_kv, kv_len, H,
                          Tkp, Vh, (__nv_bfloat16*)nullptr));
  int rc = batched_gemm(Qh, Kh, S, CST_F32, Tqp, Tkp, 64, Tqp, (int)BH, stream);
  if (rc) return rc;
  rc = batched_gemm(Gh, Vh, dP, CST_F32, Tqp, Tkp, 64, Tqp, (int)BH, stream);
  if (rc) return rc;
  CST_CHECK_CUDA(launch_k(attn_bwd_softmax_kernel, dim3(Tqp / 32, (unsigned)BH), dim3(256), 0, st, (const float*)S, (const float*)dP, Tqp, Tkp,
                          n_q, n_kv, kv_len, H, dS, dST, PT, sd, (uint32_t)site, thresh, kscale));
  rc = batched_gemm(dS, KhT, dQd, CST_F32, Tqp, 64, Tkp, Tqp, (int)BH, stream);
  if (rc) return rc;
  rc = batched_gemm(dST, QhT, dKd, CST_F32, Tkp, 64, Tqp, Tkp, (int)BH, stream);
  if (rc) return rc;
  rc = batched_gemm(PT, GhT, dVd, CST_F32, Tkp, 64, Tqp, Tkp, (int)BH, stream);
  if (rc) return rc;
  CST_CHECK_CUDA(launch_k(head_unpack_kernel, dim3(cdiv(n_q, 16), H, B), dim3(256), 0, st, (const float*)dQd, Tqp, n_q, H, dq, lddq,
                          q_rows_per_seg));
  CST_CHECK_CUDA(launch_k(head_unpack_kernel, dim3(cdiv(n_kv, 16), H, B), dim3(256), 0, st, (const float*)dKd, Tkp, n_kv, H, dk, lddkv,
                          kv_rows_per_seg));
  CST_CHECK_CUDA(launch_k(head_unpack_kernel, dim3(cdiv(n_kv, 16), H, B), dim3(256), 0, st, (const float*)dVd, Tkp, n_kv, H, dv, lddkv,
                          kv_rows_per_seg));
  return CST_OK;
}

// ---- forward with dropout of the attention probabilities (training step, attention_dropout > 0)
extern "C" long long cst_attention_dropout_fwd_ws_bytes(int B, int H, int n_q, int n_kv) {
  const long long BH = (long long)B * H, Tqp = up64(n_q), Tkp = up64(n_kv);
  return BH * (Tqp * 64 * 2 + Tkp * 64 * 2 * 3             // Qh | Kh Vh VhT
               + Tqp * Tkp * (4 + 2)                       // S | Pd
               + Tqp * 64 * 4) + 4096;                     // Od
}

// out[(b*q_rows_per_seg + t)*ldo + h*64 + d] = sum_j Pd[b,h,t,j] v[b,j,h,d] for t < n_q, Pd = softmax(q k^T + key mask) o keep / (1 - p).
// q / k / v in `qkv_dtype` (CST_BF16 / CST_F32; panels and products bf16 on tcgen05, fp32 accumulation), out in `out_dtype`.
extern "C" int cst_attention_dropout_fwd(const void* q, const void* k, const void* v, int qkv_dtype, void* out, int out_dtype,
                                         long long ldq, long long ldkv, long long ldo, int B, int H, int n_q, int q_rows_per_seg, int n_kv,
                                         int kv_rows_per_seg, const int32_t* kv_len, float p, const unsigned long long* seed,
                                         unsigned int site, void* ws, void* stream) {
  CST_REQUIRE(q && k && v && out && ws && seed && B > 0 && H > 0 && n_q > 0 && n_kv > 0, "cst_attention_dropout_fwd: bad args");
  CST_REQUIRE((qkv_dtype == CST_BF16 || qkv_dtype == CST_F32) && (out_dtype == CST_BF16 || out_dtype == CST_F32) && p >= 0.f && p < 1.f,
              "cst_attention_dropout_fwd: bad dtype / p");
  CST_REQUIRE(n_q <= q_rows_per_seg && n_kv <= kv_rows_per_seg && ((uintptr_t)ws % 256) == 0 && ldo % 4 == 0,
              "cst_attention_dropout_fwd: bad geometry");
  cudaStream_t st = (cudaStream_t)stream;
  const long long BH = (long long)B * H;
  const int Tqp = (int)up64(n_q), Tkp = (int)up64(n_kv);
  uint8_t* w = reinterpret_cast<uint8_t*>(ws);
  auto take = [&](long long bytes) { uint8_t* p_ = w; w += (bytes + 255) / 256 * 256; return p_; };
  auto* Qh = (__nv_bfloat16*)take(BH * Tqp * 64 * 2);
  auto* Kh = (__nv_bfloat16*)take(BH * Tkp * 64 * 2);
  auto* Vh = (__nv_bfloat16*)take(BH * Tkp * 64 * 2); auto* VhT = (__nv_bfloat16*)take(BH * Tkp * 64 * 2);
  auto* S = (float*)take(BH * Tqp * Tkp * 4);
  auto* Pd = (__nv_bfloat16*)take(BH * Tqp * Tkp * 2);
  auto* Od = (float*)take(BH * Tqp * 64 * 4);
  CST_REQUIRE((long long)(w - reinterpret_cast<uint8_t*>(ws)) <= cst_attention_dropout_fwd_ws_bytes(B, H, n_q, n_kv) + 8 * 256,
              "cst_attention_dropout_fwd: workspace accounting");
  CST_CHECK_CUDA(launch_k(head_pack_kernel, dim3(Tqp / 64, H, B), dim3(256), 0, st, q, qkv_dtype, ldq, q_rows_per_seg, n_q,
                          (const int32_t*)nullptr, H, Tqp, Qh, (__nv_bfloat16*)nullptr));
  CST_CHECK_CUDA(launch_k(head_pack_kernel, dim3(Tkp / 64, H, B), dim3(256), 0, st, k, qkv_dtype, ldkv, kv_rows_per_seg, n_kv, kv_len, H, Tkp,
                          Kh, (__nv_bfloat16*)nullptr));
  CST_CHECK_CUDA(launch_k(head_pack_kernel, dim3(Tkp / 64, H, B), dim3(256), 0, st, v, qkv_dtype, ldkv, kv_rows_per_seg, n_kv, kv_len, H, Tkp,
                          Vh, VhT));
  int rc = batched_gemm(Qh, Kh, S, CST_F32, Tqp, Tkp, 64, Tqp, (int)BH, stream);
  if (rc) return rc;
  CST_CHECK_CUDA(launch_k(attn_fwd_softmax_dropout_kernel, dim3(Tqp / 32, (unsigned)BH), dim3(256), 0, st, (const float*)S, Tqp, Tkp, n_q, n_kv,
                          kv_len, H, Pd, p > 0.f ? seed : (const unsigned long long*)nullptr, (uint32_t)site, dropout_threshold(p),
                          1.0f / (1.0f - p)));
  rc = batched_gemm(Pd, VhT, Od, CST_F32, Tqp, 64, Tkp, Tqp, (int)BH, stream);
  if (rc) return rc;
  if (out_dtype == CST_F32)
    CST_CHECK_CUDA(launch_k(head_unpack_kernel, dim3(cdiv(n_q, 16), H, B), dim3(256), 0, st, (const float*)Od, Tqp, n_q, H, (float*)out, ldo,
                            q_rows_per_seg));
  else
    CST_CHECK_CUDA(launch_k(head_unpack_bf16_kernel, dim3(cdiv(n_q, 16), H, B), dim3(256), 0, st, (const float*)Od, Tqp, n_q, H,
                            (__nv_bfloat16*)out, ldo, q_rows_per_seg));
  return CST_OK;
}
