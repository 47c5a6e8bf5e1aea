// f.4: backward of the fused first stage conv0 (1 -> 512, k 10, s 5, no bias) + GroupNorm(512 groups, statistics over the whole
// padded time axis) + GELU -- the derivative of cst_conv0_stats / cst_conv0_apply (forward reference: ConvFeatureExtractionModel
// block 0, fairseq/models/wav2vec/wav2vec2.py:697-734; Fp32GroupNorm fp32_group_norm.py:17-25).  Like the forward pass nothing
// of size [B, 512, T0] is stored: y = conv(x) is recomputed from the waveform (10 FMAs per value).
//   n = y*scale + shift (scale = gamma*rstd, shift = beta - mean*scale),  out = GELU(n)
//   dn = dout * GELU'(n);   per (b, c):  S1 = sum_t dn,  S2 = sum_t dn * yhat,  yhat = (n - beta) / gamma
//   dbeta_c = sum_b S1,  dgamma_c = sum_b S2,  dy = scale * (dn - S1/T0 - yhat * S2/T0),  dw[c, j] = sum_{b,t} dy * x[b, 5t + j]
// Two passes over the frames (sums, then dy / dw), each CTA = 128 frames x 512 channels of one utterance, partial sums per CTA
// reduced by cst_colsum (deterministic).  The input gradient (d waveform) is not needed.
#include "common.cuh"

namespace cst {
constexpr int CB_FT = 128, CB_K = 10, CB_S = 5, CB_C = 512;

__device__ __forceinline__ float gelu_grad(float a) {
  const float cdf = 0.5f * (1.0f + erff(a * 0.70710678118654752440f));
  return cdf + a * 0.3989422804014327f * expf(-0.5f * a * a);
}

// PASS 0: part[(b*n_chunks + chunk)*1024 + {c, 512 + c}] = {S1, S2} partials.   PASS 1: part[(b*n_chunks + chunk)*5120 + c*10 + j] = dw partials
template <int PASS>
__global__ void __launch_bounds__(CB_C) conv0_bwd_kernel(const float* __restrict__ wave, int L, int T0, int rows_per_seg,
                                                         const float* __restrict__ w, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, const float2* __restrict__ scale_shift,
                                                         const float* __restrict__ dout, const float* __restrict__ sums,
                                                         float* __restrict__ part, int n_chunks) {
  __shared__ float xs[CB_FT * CB_S + CB_K];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y, chunk = blockIdx.x, c = threadIdx.x;
  const int t0 = chunk * CB_FT;
  const float* x = wave + (size_t)b * L;
  for (int i = threadIdx.x; i < CB_FT * CB_S + CB_K; i += CB_C) {
    const long long g = (long long)t0 * CB_S + i;
    xs[i] = (g < L) ? x[g] : 0.f;
  }
  float wr[CB_K];
#pragma unroll
  for (int j = 0; j < CB_K; ++j) wr[j] = w[c * CB_K + j];
  const float2 ss = scale_shift[(size_t)b * CB_C + c];
  const float gm = gamma[c], bt = beta[c];
  const float inv_g = fabsf(gm) > 1e-30f ? 1.0f / gm : 0.f;
  float m1 = 0.f, m2 = 0.f;
  if (PASS == 1) { m1 = sums[((size_t)b * 2) * CB_C + c] / (float)T0; m2 = sums[((size_t)b * 2 + 1) * CB_C + c] / (float)T0; }
  __syncthreads();
  float s1 = 0.f, s2 = 0.f, dw[CB_K];
#pragma unroll
  for (int j = 0; j < CB_K; ++j) dw[j] = 0.f;
  const int t_end = min(CB_FT, T0 - t0);
  for (int t = 0; t < t_end; ++t) {
    float y = 0.f;
#pragma unroll
    for (int j = 0; j < CB_K; ++j) y = fmaf(wr[j], xs[t * CB_S + j], y);
    const float n = fmaf(y, ss.x, ss.y);
    const float dn = dout[((size_t)b * rows_per_seg + t0 + t) * CB_C + c] * gelu_grad(n);
    const float yhat = (n - bt) * inv_g;
    if (PASS == 0) {
      s1 += dn; s2 = fmaf(dn, yhat, s2);
    } else {
      const float dy = ss.x * (dn - m1 - yhat * m2);
#pragma unroll
      for (int j = 0; j < CB_K; ++j) dw[j] = fmaf(dy, xs[t * CB_S + j], dw[j]);
    }
  }
  if (PASS == 0) {
    float* o = part + ((size_t)b * n_chunks + chunk) * 2 * CB_C;
    o[c] = s1; o[CB_C + c] = s2;
  } else {
    float* o = part + ((size_t)b * n_chunks + chunk) * CB_C * CB_K + c * CB_K;
#pragma unroll
    for (int j = 0; j < CB_K; ++j) o[j] = dw[j];
  }
}
}  // namespace cst

// dout: f32 [B, rows_per_seg, 512] (rows t >= T0 ignored).  Workspace: ws >= B * n_chunks * 5120 + B * 1024 + 64 * 5120 floats with
// n_chunks = ceil(T0 / 128).  Outputs: dw [512, 10], dgamma [512], dbeta [512], all multiplied by grad_scale (GradMultiply,
// wav2vec2.py:530-532).
extern "C" int cst_conv0_bwd(const float* wave, int B, int L, const float* w, const float* gamma, const float* beta,
                             const float* scale_shift, const float* dout, int rows_per_seg, float* dw, float* dgamma, float* dbeta,
                             float* ws, float grad_scale, void* stream);

extern "C" int cst_conv0_bwd(const float* wave, int B, int L, const float* w, const float* gamma, const float* beta,
                             const float* scale_shift, const float* dout, int rows_per_seg, float* dw, float* dgamma, float* dbeta,
                             float* ws, float grad_scale, void* stream) {
  using namespace cst;
  CST_REQUIRE(wave && w && gamma && beta && scale_shift && dout && dw && dgamma && dbeta && ws && B > 0 && L >= CB_K,
              "cst_conv0_bwd: bad args");
  const int T0 = (L - CB_K) / CB_S + 1;
  CST_REQUIRE(rows_per_seg >= T0, "cst_conv0_bwd: rows_per_seg < T0");
  const int n_chunks = cdiv(T0, CB_FT);
  cudaStream_t st = (cudaStream_t)stream;
  float* part = ws;                                                  // [B][n_chunks][5120]
  float* sums = part + (size_t)B * n_chunks * CB_C * CB_K;           // [B][2][512]
  float* red_ws = sums + (size_t)B * 2 * CB_C;                       // colsum scratch (64 * 5120)
  const float2* ss = reinterpret_cast<const float2*>(scale_shift);
  CST_CHECK_CUDA(launch_k(conv0_bwd_kernel<0>, dim3(n_chunks, B), dim3(CB_C), 0, st, wave, L, T0, rows_per_seg, w, gamma, beta, ss, dout,
                          (const float*)nullptr, part, n_chunks));
  for (int b = 0; b < B; ++b) {
    int rc = cst_colsum(part + (size_t)b * n_chunks * 2 * CB_C, CST_F32, 2 * CB_C, n_chunks, 2 * CB_C, sums + (size_t)b * 2 * CB_C, red_ws, 1.0f, stream);
    if (rc) return rc;
  }
  // dbeta = sum_b S1, dgamma = sum_b S2 (rows of `sums` are [b][S1 | S2])
  int rc = cst_colsum(sums, CST_F32, 2 * CB_C, B, CB_C, dbeta, red_ws, grad_scale, stream);
  if (rc) return rc;
  rc = cst_colsum(sums + CB_C, CST_F32, 2 * CB_C, B, CB_C, dgamma, red_ws, grad_scale, stream);
  if (rc) return rc;
  CST_CHECK_CUDA(launch_k(conv0_bwd_kernel<1>, dim3(n_chunks, B), dim3(CB_C), 0, st, wave, L, T0, rows_per_seg, w, gamma, beta, ss, dout,
                          (const float*)sums, part, n_chunks));
  return cst_colsum(part, CST_F32, CB_C * CB_K, B * n_chunks, CB_C * CB_K, dw, red_ws, grad_scale, stream);
}
