// LayerNorm (one warp per row, statistics in fp32 registers), pos-conv operand packing, row broadcast.
#include "common.cuh"

namespace cst {

// NV = C/128 float4 chunks per lane.  Two-pass (mean, then centred variance) in registers.
template <int NV, typename LpT>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, long long ldx,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float* __restrict__ out_f32, LpT* __restrict__ out_lp, long long ldo,
                                                        int rows, int rows_per_seg, int seg_rows_valid,
                                                        long long out_rows_per_seg, int out_row_off, int zero_invalid) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int C = NV * 128;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int seg = row / rows_per_seg, t = row - seg * rows_per_seg;
  const bool valid = t < seg_rows_valid;
  if (!valid && !zero_invalid) return;
  const long long orow = (long long)seg * out_rows_per_seg + t + out_row_off;
  float4 v[NV];
  if (valid) {
    const float* xr = x + (long long)row * ldx;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = load4(xr + (lane + 32 * i) * 4);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + 1e-5f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 g = load4(gamma + (lane + 32 * i) * 4), bb = load4(beta + (lane + 32 * i) * 4);
      v[i].x = (v[i].x - mean) * rstd * g.x + bb.x;
      v[i].y = (v[i].y - mean) * rstd * g.y + bb.y;
      v[i].z = (v[i].z - mean) * rstd * g.z + bb.z;
      v[i].w = (v[i].w - mean) * rstd * g.w + bb.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const long long off = orow * ldo + (lane + 32 * i) * 4;
    if (out_f32) store4(out_f32 + off, v[i]);
    if (out_lp) store4(out_lp + off, v[i]);
  }
}

// "Light" LayerNorm for the post-LN wav2vec2 layers (16-bit mode): writes ONLY the 16-bit GEMM operand copy of the normalised
// row and the row's {rstd, -mean*rstd}; the fp32 normalised row is never stored -- the next residual GEMM re-creates it in its
// epilogue from the un-normalised row and these two numbers (cst_gemm_params.res_stats with res_slots == 0).
template <int NV, typename LpT>
__global__ void __launch_bounds__(256) layernorm_ab_kernel(const float* __restrict__ x, long long ldx,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           LpT* __restrict__ out_lp, long long ldo, float2* __restrict__ ab_out, int rows) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int C = NV * 128;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float4 v[NV];
  const float* xr = x + (long long)row * ldx;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = load4(xr + (lane + 32 * i) * 4);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + 1e-5f);
  if (lane == 0) ab_out[row] = make_float2(rstd, -mean * rstd);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = load4(gamma + (lane + 32 * i) * 4), bb = load4(beta + (lane + 32 * i) * 4);
    v[i].x = (v[i].x - mean) * rstd * g.x + bb.x;
    v[i].y = (v[i].y - mean) * rstd * g.y + bb.y;
    v[i].z = (v[i].z - mean) * rstd * g.z + bb.z;
    v[i].w = (v[i].w - mean) * rstd * g.w + bb.w;
    store4(out_lp + (long long)row * ldo + (lane + 32 * i) * 4, v[i]);
  }
}

// xg[b][g][row][64]: frame t at row t+64, lanes 0..47 = x[b, t, g*48 .. g*48+47], rest zero.
template <typename OutT>
__global__ void posconv_pack_kernel(const float* __restrict__ x, int rows_per_seg, int n_frames,
                                    OutT* __restrict__ xg, int t_pad_rows, long long total_chunks) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one 8-lane chunk
  if (idx >= total_chunks) return;
  const int ch = (int)(idx & 7);
  long long r = idx >> 3;
  const int row = (int)(r % t_pad_rows); r /= t_pad_rows;
  const int g = (int)(r & 15);
  const int b = (int)(r >> 4);
  const int t = row - 64;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
  if (ch < 6 && t >= 0 && t < n_frames) {
    const float* p = x + ((long long)b * rows_per_seg + t) * 768 + g * 48 + ch * 8;
    a = load4(p); c = load4(p + 4);
  }
  OutT* o = xg + idx * 8;
  store4(o, a); store4(o + 4, c);
}

__global__ void broadcast_rows_kernel(const float4* __restrict__ src, long long n4, float4* __restrict__ dst, int B) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = src[i];
  for (int b = 0; b < B; ++b) dst[(long long)b * n4 + i] = v;
}
}  // namespace cst

extern "C" int cst_layernorm(const float* x, long long ldx, const float* gamma, const float* beta,
                             float* out_f32, void* out_lp, int lp_dtype, long long ldo,
                             int rows, int C, int rows_per_seg, int seg_rows_valid,
                             long long out_rows_per_seg, int out_row_off, int zero_invalid, void* stream) {
  using namespace cst;
  CST_REQUIRE(x && gamma && beta && (out_f32 || out_lp) && rows > 0 && rows_per_seg > 0,
              "cst_layernorm: bad args rows=%d", rows);
  CST_REQUIRE(C == 512 || C == 768, "cst_layernorm: C=%d unsupported (512 or 768)", C);
  CST_REQUIRE(ldx % 4 == 0 && ldo % 4 == 0, "cst_layernorm: ldx/ldo must be multiples of 4");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(cdiv(rows, 8));
#define CST_LN(NV, T)                                                                                   \
  CST_CHECK_CUDA(launch_k(layernorm_kernel<NV, T>, grid, dim3(256), 0, st, x, ldx, gamma, beta, out_f32, (T*)out_lp, ldo, rows, \
                          rows_per_seg, seg_rows_valid, out_rows_per_seg, out_row_off, zero_invalid))
  if (out_lp && lp_dtype == CST_BF16) { if (C == 512) CST_LN(4, __nv_bfloat16); else CST_LN(6, __nv_bfloat16); }
  else { if (C == 512) CST_LN(4, float); else CST_LN(6, float); }
#undef CST_LN
  return CST_OK;
}

extern "C" int cst_layernorm_ab(const float* x, long long ldx, const float* gamma, const float* beta, void* out_lp, int lp_dtype,
                                long long ldo, float* ab_out, int rows, int C, void* stream) {
  using namespace cst;
  CST_REQUIRE(x && gamma && beta && out_lp && ab_out && rows > 0, "cst_layernorm_ab: bad args rows=%d", rows);
  CST_REQUIRE(C == 512 || C == 768, "cst_layernorm_ab: C=%d unsupported (512 or 768)", C);
  CST_REQUIRE(lp_dtype == CST_BF16, "cst_layernorm_ab: bf16 operand copy only");
  CST_REQUIRE(ldx % 4 == 0 && ldo % 4 == 0, "cst_layernorm_ab: ldx/ldo must be multiples of 4");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(cdiv(rows, 8));
  if (C == 512) CST_CHECK_CUDA(launch_k(layernorm_ab_kernel<4, __nv_bfloat16>, grid, dim3(256), 0, st, x, ldx, gamma, beta, (__nv_bfloat16*)out_lp, ldo, (float2*)ab_out, rows));
  else CST_CHECK_CUDA(launch_k(layernorm_ab_kernel<6, __nv_bfloat16>, grid, dim3(256), 0, st, x, ldx, gamma, beta, (__nv_bfloat16*)out_lp, ldo, (float2*)ab_out, rows));
  return CST_OK;
}

extern "C" int cst_posconv_pack(const float* x, int B, int rows_per_seg, int n_frames, void* xg, int xg_dtype,
                                int t_pad_rows, void* stream) {
  using namespace cst;
  CST_REQUIRE(x && xg && B > 0 && n_frames > 0 && n_frames <= rows_per_seg && t_pad_rows >= n_frames + 128,
              "cst_posconv_pack: bad args B=%d n_frames=%d rows_per_seg=%d t_pad_rows=%d", B, n_frames, rows_per_seg, t_pad_rows);
  const long long chunks = (long long)B * 16 * t_pad_rows * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (xg_dtype == CST_BF16)
    CST_CHECK_CUDA(launch_k(posconv_pack_kernel<__nv_bfloat16>, dim3(cdiv(chunks, 256)), dim3(256), 0, st, x, rows_per_seg, n_frames, (__nv_bfloat16*)xg, t_pad_rows, chunks));
  else
    CST_CHECK_CUDA(launch_k(posconv_pack_kernel<float>, dim3(cdiv(chunks, 256)), dim3(256), 0, st, x, rows_per_seg, n_frames, (float*)xg, t_pad_rows, chunks));
  return CST_OK;
}

extern "C" int cst_broadcast_rows(const float* src, int rows, int C, int B, float* dst, void* stream) {
  using namespace cst;
  CST_REQUIRE(src && dst && rows > 0 && C % 4 == 0 && B > 0, "cst_broadcast_rows: bad args");
  const long long n4 = (long long)rows * C / 4;
  CST_CHECK_CUDA(launch_k(broadcast_rows_kernel, dim3(cdiv(n4, 256)), dim3(256), 0, (cudaStream_t)stream, (const float4*)src, n4, (float4*)dst, B));
  return CST_OK;
}
