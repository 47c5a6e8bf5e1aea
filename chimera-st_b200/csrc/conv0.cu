// a1 / K1: conv0 (1->512, k=10, s=5, no bias) + GroupNorm(512 groups, statistics over the WHOLE
// padded time axis) + exact GELU, channels-last output.  HBM-bound by design: the [B,T0,512]
// activation is written exactly once; the statistics never touch it.
//
//   y[b,c,t] = sum_j w[c,j] x[b,5t+j]  =>  sum_t y = w_c . S_b,   sum_t y^2 = w_c^T R_b w_c
//   with S_b[j] = sum_t x[b,5t+j] and R_b[j,j'] = sum_t x[b,5t+j] x[b,5t+j']  (10 + 55 numbers per
//   utterance, one pass over the 4*L-byte waveform).  Mean/variance (biased, eps 1e-5, as
//   F.group_norm, fp32_group_norm.py:17-25) are finished in fp64 per (b,c).
#include "common.cuh"

namespace cst {

constexpr int C0 = 512, K0 = 10, S0 = 5;
constexpr int NSTAT = 65;         // 10 sums + 55 upper-triangular lag products
constexpr int STAT_STRIDE = 72;

__global__ void __launch_bounds__(256) conv0_stats_kernel(const float* __restrict__ wave, int L, int T0,
                                                          int frames_per_block, double* __restrict__ ws) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const float* x = wave + (size_t)b * L;
  const int t_begin = blockIdx.x * frames_per_block;
  const int t_end = min(T0, t_begin + frames_per_block);
  float acc[NSTAT];
#pragma unroll
  for (int i = 0; i < NSTAT; ++i) acc[i] = 0.f;
  for (int t = t_begin + threadIdx.x; t < t_end; t += blockDim.x) {
    float v[K0];
#pragma unroll
    for (int j = 0; j < K0; ++j) v[j] = __ldg(x + (size_t)t * S0 + j);
    int q = K0;
#pragma unroll
    for (int j = 0; j < K0; ++j) {
      acc[j] += v[j];
#pragma unroll
      for (int jj = j; jj < K0; ++jj) { acc[q] = fmaf(v[j], v[jj], acc[q]); ++q; }
    }
  }
  __shared__ float red[8][NSTAT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < NSTAT; ++i) {
    float s = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < NSTAT) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += (double)red[w][threadIdx.x];
    atomicAdd(ws + (size_t)b * STAT_STRIDE + threadIdx.x, s);
  }
}

__global__ void __launch_bounds__(C0) conv0_finalize_kernel(const double* __restrict__ ws, const float* __restrict__ w,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             int T0, float2* __restrict__ scale_shift) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, c = threadIdx.x;
  __shared__ double st[NSTAT];
  if (c < NSTAT) st[c] = ws[(size_t)b * STAT_STRIDE + c];
  __syncthreads();
  double wc[K0];
#pragma unroll
  for (int j = 0; j < K0; ++j) wc[j] = (double)w[c * K0 + j];
  double sum = 0.0, sq = 0.0;
  int q = K0;
#pragma unroll
  for (int j = 0; j < K0; ++j) {
    sum += wc[j] * st[j];
#pragma unroll
    for (int jj = j; jj < K0; ++jj) { sq += (j == jj ? 1.0 : 2.0) * wc[j] * wc[jj] * st[q]; ++q; }
  }
  const double mean = sum / T0;
  double var = sq / T0 - mean * mean;
  if (var < 0.0) var = 0.0;
  const double rstd = 1.0 / sqrt(var + 1e-5);
  const double sc = (double)gamma[c] * rstd;
  scale_shift[(size_t)b * C0 + c] = make_float2((float)sc, (float)((double)beta[c] - mean * sc));
}

// block: 4 frame lanes x 64 channel octets; FT frames per block.  Two channels share one packed FFMA2 chain
// (w[c][j], w[c+1][j]) * (x[j], x[j]); the bf16 path uses the branch-free packed erf GELU (gelu2), the fp32
// path keeps erff() so the 1e-5 parity mode stays bit-comparable to torch's erf GELU.
template <typename OutT, int FT>
__global__ void __launch_bounds__(256, 2) conv0_apply_kernel(const float* __restrict__ wave, int L, int T0, int rows_per_seg,
                                                          const float* __restrict__ w, const float2* __restrict__ scale_shift,
                                                          OutT* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * FT;
  __shared__ float xs[FT * S0 + K0];
  const float* x = wave + (size_t)b * L;
  for (int i = threadIdx.x; i < FT * S0 + K0; i += 256) {
    const long long g = (long long)t0 * S0 + i;
    xs[i] = (g < L) ? __ldg(x + g) : 0.f;
  }
  const int cq = threadIdx.x & 63, tq = threadIdx.x >> 6;
  const int c0 = cq * 8;
  uint64_t wr[4][K0], sc[4], sh[4];                 // channel pairs (c0+2i, c0+2i+1)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < K0; ++j) wr[i][j] = pk2(__ldg(w + (c0 + 2 * i) * K0 + j), __ldg(w + (c0 + 2 * i + 1) * K0 + j));
    const float2 s0 = scale_shift[(size_t)b * C0 + c0 + 2 * i], s1 = scale_shift[(size_t)b * C0 + c0 + 2 * i + 1];
    sc[i] = pk2(s0.x, s1.x); sh[i] = pk2(s0.y, s1.y);
  }
  __syncthreads();
  OutT* orow = out + ((size_t)b * rows_per_seg) * C0 + c0;
  // two frames (tl, tl+4) per iteration: 8 independent FFMA2 chains keep the FMA pipe fed between dependent ops
#pragma unroll 1
  for (int tl = tq; tl < FT; tl += 8) {
    float y[2][8];
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const int tf = tl + 4 * f;
      uint64_t xv[K0];
#pragma unroll
      for (int j = 0; j < K0; ++j) { const float v = xs[tf * S0 + j]; xv[j] = pk2(v, v); }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint64_t a = fmul2(wr[i][0], xv[0]);
#pragma unroll
        for (int j = 1; j < K0; ++j) a = ffma2(wr[i][j], xv[j], a);
        upk2(ffma2(a, sc[i], sh[i]), y[f][2 * i], y[f][2 * i + 1]);
      }
    }
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (sizeof(OutT) == 2) gelu2(y[f][2 * i], y[f][2 * i + 1]);
        else { y[f][2 * i] = gelu_erf(y[f][2 * i]); y[f][2 * i + 1] = gelu_erf(y[f][2 * i + 1]); }
      }
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const int t = t0 + tl + 4 * f;
      if (t >= rows_per_seg) continue;
      const bool live = t < T0;
      OutT* o = orow + (size_t)t * C0;
      store4(o, live ? make_float4(y[f][0], y[f][1], y[f][2], y[f][3]) : make_float4(0.f, 0.f, 0.f, 0.f));
      store4(o + 4, live ? make_float4(y[f][4], y[f][5], y[f][6], y[f][7]) : make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
}
}  // namespace cst

extern "C" int cst_conv0_stats(const float* wave, int B, int L, const float* w, const float* gamma,
                               const float* beta, float* scale_shift, double* stats_ws, void* stream) {
  using namespace cst;
  CST_REQUIRE(wave && w && gamma && beta && scale_shift && stats_ws && B > 0 && L >= K0,
              "cst_conv0_stats: bad args B=%d L=%d", B, L);
  cudaStream_t st = (cudaStream_t)stream;
  const int T0 = (L - K0) / S0 + 1;
  CST_CHECK_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(double) * STAT_STRIDE * B, st));
  const int fpb = 2048;                    // 8 frames per thread: short fp32 chains, then fp64
  dim3 grid(cdiv(T0, fpb), B);
  CST_CHECK_CUDA(launch_k(conv0_stats_kernel, grid, dim3(256), 0, st, wave, L, T0, fpb, stats_ws));
  CST_CHECK_CUDA(launch_k(conv0_finalize_kernel, dim3(B), dim3(C0), 0, st, (const double*)stats_ws, w, gamma, beta, T0, reinterpret_cast<float2*>(scale_shift)));
  return CST_OK;
}

extern "C" int cst_conv0_apply(const float* wave, int B, int L, const float* w, const float* scale_shift,
                               void* out, int out_dtype, int rows_per_seg, void* stream) {
  using namespace cst;
  const int T0 = (L - K0) / S0 + 1;
  CST_REQUIRE(wave && w && scale_shift && out && B > 0 && L >= K0 && rows_per_seg >= T0,
              "cst_conv0_apply: bad args B=%d L=%d rows_per_seg=%d (T0=%d)", B, L, rows_per_seg, T0);
  constexpr int FT = 256;                  // frames per block: amortises the per-thread weight fetch (80 + 16 registers)
  dim3 grid(cdiv(rows_per_seg, FT), B);
  cudaStream_t st = (cudaStream_t)stream;
  const float2* ss = reinterpret_cast<const float2*>(scale_shift);
  if (out_dtype == CST_F16)
    CST_CHECK_CUDA(launch_k(conv0_apply_kernel<__half, FT>, grid, dim3(256), 0, st, wave, L, T0, rows_per_seg, w, ss, (__half*)out));
  else if (out_dtype == CST_BF16)
    CST_CHECK_CUDA(launch_k(conv0_apply_kernel<__nv_bfloat16, FT>, grid, dim3(256), 0, st, wave, L, T0, rows_per_seg, w, ss, (__nv_bfloat16*)out));
  else if (out_dtype == CST_F32)
    CST_CHECK_CUDA(launch_k(conv0_apply_kernel<float, FT>, grid, dim3(256), 0, st, wave, L, T0, rows_per_seg, w, ss, (float*)out));
  else
    CST_REQUIRE(false, "cst_conv0_apply: bad out_dtype %d", out_dtype);
  CST_LAUNCH_CHECK();
  return CST_OK;
}
