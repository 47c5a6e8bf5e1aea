// f.2: training heads that consume the path's outputs.
//   * contrastive (InfoNCE) loss over the M shared semantic memories of the audio and text passes --
//     TripletSTMTContrastiveCriterion.compute_contrastive (fairseq/criterions/triplet_st_mt_contrastive.py:154-169):
//     logits[b,i,j] = cos(audio[b,i], text[b,j]) / temp (cosine in fp32); the reference passes the 3-D [b,i,j] tensor to
//     F.cross_entropy, whose class axis is dim 1: loss = sum_b sum_j CE over the AUDIO index i of logits[b,:,j], target j;
//   * label-smoothed cross entropy -- label_smoothed_nll_loss (fairseq/criterions/label_smoothed_cross_entropy.py:13-30).
// Both return per-row terms (fixed summation order on the caller's side: deterministic) and, optionally, the gradient with
// respect to their inputs (the start of the backward pass of BASELINE configs[4]).
#include "common.cuh"

namespace cst {

constexpr int CL_THREADS = 256, CL_MAXM = 64, CL_MAXC = 1024;

__device__ __forceinline__ float cl_block_sum(float v, float* red) {           // all threads get the total
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < CL_THREADS / 32; ++w) s += red[w];
  return s;
}

// x / y: [M, B, C] (the encoder_out layout, time-major).  grid (M, B); CTA (r, b) owns row r of "own" and walks the M rows
// of "other".  SWAP = false: own = text row j, other = audio rows i: the softmax over i of column j -> loss_rows[b*M + j],
// lse[b*M + j], d(text).  SWAP = true: own = audio row i, other = text rows j -> d(audio) from the stored per-column lse.
// cos = x.y / max(|x| |y|, 1e-8) (torch.cosine_similarity).
template <typename T, bool SWAP>
__global__ void __launch_bounds__(CL_THREADS) contrastive_kernel(const T* __restrict__ own, const T* __restrict__ other, int M, int B, int C,
                                                                 float inv_temp, float* __restrict__ loss_rows, float* __restrict__ lse,
                                                                 float* __restrict__ d_own, float dscale) {
  __shared__ float xs[CL_MAXC];
  __shared__ float dots[CL_MAXM], onorm[CL_MAXM], coef[CL_MAXM];
  __shared__ float red[CL_THREADS / 32];
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* xr = own + ((size_t)i * B + b) * C;
  float sq = 0.f;
  for (int c = tid; c < C; c += CL_THREADS) { const float v = (float)xr[c]; xs[c] = v; sq += v * v; }
  const float xn = sqrtf(cl_block_sum(sq, red));
  for (int j = warp; j < M; j += CL_THREADS / 32) {               // warp per row of `other`
    const T* yr = other + ((size_t)j * B + b) * C;
    float d = 0.f, n2 = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = (float)yr[c]; d = fmaf(xs[c], v, d); n2 = fmaf(v, v, n2); }
    d = warp_sum(d); n2 = warp_sum(n2);
    if (lane == 0) { dots[j] = d; onorm[j] = sqrtf(n2); }
  }
  __syncthreads();
  // cosines and logits; for SWAP the logit of pair (audio j, text i) still belongs to audio row j's softmax
  if (tid < M) {
    const float denom = fmaxf(xn * onorm[tid], 1e-8f);
    dots[tid] = dots[tid] / denom;                                 // cos(own, other_j)
  }
  __syncthreads();
  if (!SWAP) {
    if (warp == 0) {                                               // log-sum-exp over j of cos * inv_temp
      float mx = -INFINITY;
      for (int j = lane; j < M; j += 32) mx = fmaxf(mx, dots[j] * inv_temp);
      mx = warp_max(mx);
      float s = 0.f;
      for (int j = lane; j < M; j += 32) s += expf(dots[j] * inv_temp - mx);
      s = warp_sum(s);
      const float l = mx + logf(s);
      if (lane == 0) {
        lse[(size_t)b * M + i] = l;
        loss_rows[(size_t)b * M + i] = l - dots[i] * inv_temp;     // -log softmax at the target column i
      }
      for (int j = lane; j < M; j += 32)                           // dL/dlogit_ij * dlogit/dcos
        coef[j] = (expf(dots[j] * inv_temp - l) - (j == i ? 1.f : 0.f)) * inv_temp * dscale;
    }
  } else {
    if (tid < M)                                                   // audio row j = tid, text column i: softmax of row j
      coef[tid] = (expf(dots[tid] * inv_temp - lse[(size_t)b * M + tid]) - (tid == i ? 1.f : 0.f)) * inv_temp * dscale;
  }
  __syncthreads();
  if (d_own == nullptr) return;
  // d own = sum_j coef_j * ( y_j / (|x||y_j|) - cos_j * x / |x|^2 )
  float self = 0.f;
  for (int j = 0; j < M; ++j) self += coef[j] * dots[j];
  const float inv_x2 = 1.0f / fmaxf(xn * xn, 1e-16f);
  for (int c = tid; c < C; c += CL_THREADS) {
    float acc = 0.f;
    for (int j = 0; j < M; ++j) {
      const float denom = fmaxf(xn * onorm[j], 1e-8f);
      acc = fmaf(coef[j] / denom, (float)other[((size_t)j * B + b) * C + c], acc);
    }
    d_own[((size_t)i * B + b) * C + c] = acc - self * inv_x2 * xs[c];
  }
}

// one CTA per target position: row of V logits -> log-softmax statistics, loss terms, optional gradient
__global__ void __launch_bounds__(CL_THREADS) ls_ce_kernel(const float* __restrict__ logits, long long ld, const int64_t* __restrict__ target,
                                                           int V, float eps, long long ignore_index, float* __restrict__ loss_rows,
                                                           float* __restrict__ nll_rows, float* __restrict__ dlogits, long long ldd,
                                                           float dscale) {
  __shared__ float red[CL_THREADS / 32];
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* x = logits + (size_t)r * ld;
  const long long t = target[r];
  const bool ignored = t == ignore_index;
  float mx = -INFINITY;
  for (int v = tid; v < V; v += CL_THREADS) mx = fmaxf(mx, x[v]);
  mx = warp_max(mx);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < CL_THREADS / 32; ++w) mx = fmaxf(mx, red[w]);
  float se = 0.f, sx = 0.f;
  for (int v = tid; v < V; v += CL_THREADS) { se += expf(x[v] - mx); sx += x[v]; }
  se = cl_block_sum(se, red);
  sx = cl_block_sum(sx, red);
  const float lse = mx + logf(se);
  if (tid == 0) {
    const float nll = ignored ? 0.f : lse - x[t];                  // -lprobs[target]
    const float smooth = ignored ? 0.f : (float)V * lse - sx;      // -sum_v lprobs[v]
    nll_rows[r] = nll;
    loss_rows[r] = (1.0f - eps) * nll + (eps / (float)V) * smooth;
  }
  if (dlogits != nullptr) {
    float* d = dlogits + (size_t)r * ldd;
    const float eps_i = eps / (float)V;
    for (int v = tid; v < V; v += CL_THREADS) {
      float g = 0.f;
      if (!ignored) g = (expf(x[v] - lse) - (v == t ? 1.0f - eps : 0.f) - eps_i) * dscale;
      d[v] = g;
    }
  }
}

// deterministic sum of n floats (single CTA, fixed order): the `reduce=True` form of both criteria
__global__ void __launch_bounds__(CL_THREADS) sum_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  __shared__ float red[CL_THREADS / 32];
  pdl_launch_dependents();
  pdl_wait();
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += CL_THREADS) s += (double)x[i];
  const float t = cl_block_sum((float)s, red);
  if (threadIdx.x == 0) *out = t;
}

}  // namespace cst

extern "C" int cst_contrastive_loss(const void* audio, const void* text, int dtype, int M, int B, int C, float temp,
                                    float* loss_rows, float* lse_ws, float* d_audio, float* d_text, float dscale, void* stream) {
  using namespace cst;
  CST_REQUIRE(audio && text && loss_rows && lse_ws && M > 0 && M <= CL_MAXM && B > 0 && B <= 65535 && C > 0 && C <= CL_MAXC && temp > 0.f,
              "cst_contrastive_loss: bad args M=%d B=%d C=%d", M, B, C);
  CST_REQUIRE((d_audio == nullptr) == (d_text == nullptr), "cst_contrastive_loss: pass both gradients or neither");
  cudaStream_t st = (cudaStream_t)stream;
  const float it = 1.0f / temp;
  dim3 grid(M, B);
  if (dtype == CST_F32) {
    CST_CHECK_CUDA(launch_k(contrastive_kernel<float, false>, grid, dim3(CL_THREADS), 0, st, (const float*)text, (const float*)audio, M, B, C, it,
                            loss_rows, lse_ws, d_text, dscale));
    if (d_audio)
      CST_CHECK_CUDA(launch_k(contrastive_kernel<float, true>, grid, dim3(CL_THREADS), 0, st, (const float*)audio, (const float*)text, M, B, C, it,
                              loss_rows, lse_ws, d_audio, dscale));
  } else if (dtype == CST_BF16) {
    CST_CHECK_CUDA(launch_k(contrastive_kernel<__nv_bfloat16, false>, grid, dim3(CL_THREADS), 0, st, (const __nv_bfloat16*)text,
                            (const __nv_bfloat16*)audio, M, B, C, it, loss_rows, lse_ws, d_text, dscale));
    if (d_audio)
      CST_CHECK_CUDA(launch_k(contrastive_kernel<__nv_bfloat16, true>, grid, dim3(CL_THREADS), 0, st, (const __nv_bfloat16*)audio,
                              (const __nv_bfloat16*)text, M, B, C, it, loss_rows, lse_ws, d_audio, dscale));
  } else {
    CST_REQUIRE(false, "cst_contrastive_loss: bad dtype %d", dtype);
  }
  return CST_OK;
}

extern "C" int cst_label_smoothed_ce(const float* logits, long long ld, const int64_t* target, int N, int V, float eps,
                                     long long ignore_index, float* loss_rows, float* nll_rows, float* dlogits, long long ldd,
                                     float dscale, void* stream) {
  using namespace cst;
  CST_REQUIRE(logits && target && loss_rows && nll_rows && N > 0 && V > 0 && ld >= V && (dlogits == nullptr || ldd >= V),
              "cst_label_smoothed_ce: bad args N=%d V=%d", N, V);
  CST_CHECK_CUDA(launch_k(ls_ce_kernel, dim3(N), dim3(CL_THREADS), 0, (cudaStream_t)stream, logits, ld, target, V, eps, ignore_index,
                          loss_rows, nll_rows, dlogits, ldd, dscale));
  return CST_OK;
}

extern "C" int cst_sum(const float* x, long long n, float* out, void* stream) {
  using namespace cst;
  CST_REQUIRE(x && out && n > 0, "cst_sum: bad args");
  CST_CHECK_CUDA(launch_k(sum_kernel, dim3(1), dim3(CL_THREADS), 0, (cudaStream_t)stream, x, n, out));
  return CST_OK;
}
