// Epilogue shared by the FFMA (fp32) and tcgen05 (bf16) GEMM kernels: bias, activation, alpha,
// residual, GLU on adjacent column pairs, row remap / masking, f32 or bf16 store.
#pragma once
#include "common.cuh"

namespace cst {

struct GemmDev {          // device copy of cst_gemm_params (pointers already typed by the launcher)
  const void* A; const void* W; const float* bias; const float* residual; void* C;
  int c_dtype;
  int M, N, K;
  long long lda, ldc, ldr, a_limit;     // a_limit: number of addressable A elements per batch z
  int act; float alpha;
  int nb_inner;
  long long a_bs_outer, a_bs_inner, w_bs_inner, c_bs_outer, c_bs_inner, r_bs_outer, r_bs_inner;
  int bias_bs_inner;
  int rows_per_seg, seg_rows_valid; long long out_rows_per_seg; int out_row_off;
  const int32_t* seg_len; int segs_per_outer;
  // LayerNorm fused around the GEMM (tensor-core path; see cst_gemm_params)
  const float2* ln_in_stats; const float* ln_colsum; int ln_in_slots;
  const float2* res_stats; int res_slots; const float* res_gamma; const float* res_beta;
  __nv_bfloat16* C2; long long ldc2;
  float2* out_stats;
  float ln_inv_dim;
  int exact_act;
  float acc_scale;
};

// {rstd, -mean * rstd} of one row from its partial sums (sum, sum of squares per 128-column slice), eps 1e-5
__device__ __forceinline__ float2 ln_ab_from_partials(const float2* stats, long long row, int slots, float inv_dim) {
  if (slots == 0) return __ldg(stats + row);       // already {rstd, -mean*rstd} (cst_layernorm_ab)
  float s = 0.f, q = 0.f;
  const float2* sp = stats + row * 8;
  for (int i = 0; i < slots; ++i) { const float2 t = __ldg(sp + i); s += t.x; q += t.y; }
  const float mean = s * inv_dim;
  const float var = fmaxf(q * inv_dim - mean * mean, 0.f);
  const float rstd = rsqrtf(var + 1e-5f);
  return make_float2(rstd, -mean * rstd);
}

struct RowInfo { long long out_row; bool store; bool zero; };

__device__ __forceinline__ RowInfo row_info(const GemmDev& p, int m, int zo) {
  RowInfo r;
  const int seg = m / p.rows_per_seg, t = m - seg * p.rows_per_seg;
  r.store = (m < p.M) && (t < p.seg_rows_valid);
  r.zero = false;
  if (p.seg_len != nullptr && r.store) r.zero = t >= p.seg_len[zo * p.segs_per_outer + seg];
  r.out_row = (long long)seg * p.out_rows_per_seg + t + p.out_row_off;
  return r;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// 8 consecutive accumulator columns starting at column n (n % 8 == 0, n < N) of one output row.
// c_off / r_off: batch offsets (elements) into C / residual; bias already offset for the batch.
__device__ __forceinline__ void epilogue8(const GemmDev& p, const RowInfo& ri, int n, const float* bias,
                                          long long c_off, long long r_off, const float (&acc)[8]) {
  if (!ri.store || n >= p.N) return;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = acc[j] + (bias ? __ldg(bias + n + j) : 0.f);
  if (p.act == CST_ACT_GLU) {
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = v[2 * j] * sigmoidf_(v[2 * j + 1]) * p.alpha;
    const int nc = n >> 1;
    if (p.residual) {
      const float4 r = load4(p.residual + r_off + ri.out_row * p.ldr + nc);
      o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
    }
    if (ri.zero) { o[0] = o[1] = o[2] = o[3] = 0.f; }
    const long long off = c_off + ri.out_row * p.ldc + nc;
    if (p.c_dtype == CST_BF16) store4((__nv_bfloat16*)p.C + off, make_float4(o[0], o[1], o[2], o[3]));
    else store4((float*)p.C + off, make_float4(o[0], o[1], o[2], o[3]));
    return;
  }
  if (p.act == CST_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
  } else if (p.act == CST_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] *= p.alpha;
  if (p.residual) {
    const float* rp = p.residual + r_off + ri.out_row * p.ldr + n;
    const float4 r0 = load4(rp), r1 = load4(rp + 4);
    v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
    v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
  }
  if (ri.zero) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
  }
  const long long off = c_off + ri.out_row * p.ldc + n;
  if (p.c_dtype == CST_BF16) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>((__nv_bfloat16*)p.C + off) = u;
  } else {
    float* o = (float*)p.C + off;
    store4(o, make_float4(v[0], v[1], v[2], v[3]));
    store4(o + 4, make_float4(v[4], v[5], v[6], v[7]));
  }
}

// ---- fast exact-erf GELU for the tensor-core epilogue: Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7 on erf),
// branch-free: 1 MUFU.RCP + 1 MUFU.EX2 + ~12 FP32 ops (erff() costs ~2x and diverges on |x|).
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = exp2f(-1.4426950408889634f * z * z);
  const float erf_abs = fmaf(-poly, e, 1.0f);                 // erf(|x|/sqrt2)
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), erf_abs, hx);                        // 0.5x(1 + sign(x) erf(|x|/sqrt2))
}

struct RowDesc { long long out_row; int flags; int pad; };    // flags: 1 = store, 2 = store zeros

// 4 consecutive accumulator columns n..n+3 (n % 4 == 0) of one output row, in the row-contiguous domain.
__device__ __forceinline__ void epilogue4(const GemmDev& p, long long out_row, bool zero, int n, const float4 b4,
                                          long long c_off, long long r_off, const float4 acc) {
  float v[4] = {acc.x + b4.x, acc.y + b4.y, acc.z + b4.z, acc.w + b4.w};
  if (p.act == CST_ACT_GLU) {
    float o0 = v[0] * sigmoidf_(v[1]) * p.alpha, o1 = v[2] * sigmoidf_(v[3]) * p.alpha;
    const int nc = n >> 1;
    if (p.residual) {
      const float2 r = *reinterpret_cast<const float2*>(p.residual + r_off + out_row * p.ldr + nc);
      o0 += r.x; o1 += r.y;
    }
    if (zero) { o0 = 0.f; o1 = 0.f; }
    const long long off = c_off + out_row * p.ldc + nc;
    if (p.c_dtype == CST_BF16) *reinterpret_cast<uint32_t*>((__nv_bfloat16*)p.C + off) = pack_bf16x2(o0, o1);
    else *reinterpret_cast<float2*>((float*)p.C + off) = make_float2(o0, o1);
    return;
  }
  if (p.act == CST_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = gelu_fast(v[j]);
  } else if (p.act == CST_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] *= p.alpha;
  if (p.residual) {
    const float4 r = load4(p.residual + r_off + out_row * p.ldr + n);
    v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
  }
  if (zero) { v[0] = v[1] = v[2] = v[3] = 0.f; }
  const long long off = c_off + out_row * p.ldc + n;
  if (p.c_dtype == CST_BF16) store4((__nv_bfloat16*)p.C + off, make_float4(v[0], v[1], v[2], v[3]));
  else store4((float*)p.C + off, make_float4(v[0], v[1], v[2], v[3]));
}

int launch_gemm_f32(const GemmDev& p, int nz, cudaStream_t st);
int launch_gemm_tc(const cst_gemm_params& hp, const GemmDev& p, int nz, cudaStream_t st);

}  // namespace cst
