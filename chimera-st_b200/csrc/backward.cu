// f.4 (BASELINE configs[4]): the pieces of the BACKWARD pass of the speech-encoding path that are not GEMMs.
// The reference has no backward code of its own on this path -- torch autograd differentiates the modules of
// fairseq/models/wav2vec/wav2vec2.py, fairseq/modules/transformer_layer.py, multihead_attention.py,
// fairseq/models/speech_to_text/s2t_transformer.py:31-77 and fairseq/models/chimera/w2v2_transformer_interlingua.py:207-312;
// these kernels are the hand-written derivatives of the forward kernels in this directory (parity: autograd through
// oracle/chimera_oracle.py, tests/test_gpu_backward.py).  GEMM-shaped gradients (dgrad / wgrad) go through cst_gemm on
// transposed copies made by cst_transpose.  All gradients are fp32.
#include <cuda_fp16.h>
#include "common.cuh"
#include "philox.cuh"

namespace cst {


// ---- runtime-typed element access: the tape of the 16-bit training mode keeps GEMM operands / pre-activations in bf16 and
// gradients in fp32; these passes are HBM-bound, the dtype branch is uniform per launch.
__device__ __forceinline__ float ld_any(const void* p, int dt, long long i) {
  if (dt == CST_F32) return reinterpret_cast<const float*>(p)[i];
  if (dt == CST_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return __half2float(reinterpret_cast<const __half*>(p)[i]);
}
__device__ __forceinline__ void st_any(void* p, int dt, long long i, float v) {
  if (dt == CST_F32) reinterpret_cast<float*>(p)[i] = v;
  else if (dt == CST_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
}
__device__ __forceinline__ float4 ld4_any(const void* p, int dt, long long i) {      // i % 4 == 0, row pitches % 4 == 0
  if (dt == CST_F32) return load4(reinterpret_cast<const float*>(p) + i);
  if (dt == CST_BF16) return load4(reinterpret_cast<const __nv_bfloat16*>(p) + i);
  const __half* h = reinterpret_cast<const __half*>(p) + i;
  return make_float4(__half2float(h[0]), __half2float(h[1]), __half2float(h[2]), __half2float(h[3]));
}
__device__ __forceinline__ void st4_any(void* p, int dt, long long i, float4 v) {
  if (dt == CST_F32) store4(reinterpret_cast<float*>(p) + i, v);
  else if (dt == CST_BF16) store4(reinterpret_cast<__nv_bfloat16*>(p) + i, v);
  else store4(reinterpret_cast<__half*>(p) + i, v);
}

// ---- transpose (+ cast): outT[c][r] = x[r * ldx + c]; columns r in [rows, rows_pad) of outT are zero-filled so that the result
// can be the K-major operand of a GEMM whose reduction axis (rows) must be a multiple of 64.  ldx may be smaller than cols
// (overlapping windows: the implicit-GEMM view of a strided convolution's input).  outT is stored in `chunk`-column slabs,
// element (c, r) at ((r / chunk) * cols + c) * chunk + r % chunk: chunk == rows_pad is the plain [cols, rows_pad] matrix, smaller
// chunks lay the reduction axis out as a batch of GEMMs (split-K for the weight gradients of the conv stack, whose reduction
// runs over >1e5 rows while the output has only a few dozen tiles).  `copy` (optional) receives the un-transposed cast copy.
__global__ void __launch_bounds__(256) transpose_kernel(const void* __restrict__ x, int xdt, long long ldx, int rows, int cols,
                                                        void* __restrict__ out, int odt, int rows_pad, int chunk,
                                                        void* __restrict__ copy, long long ldcopy) {
  __shared__ float tile[32][33];
  pdl_launch_dependents();
  pdl_wait();
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int r = r0 + ty + i, c = c0 + tx;
    const bool ok = r < rows && c < cols;
    const float v = ok ? ld_any(x, xdt, (long long)r * ldx + c) : 0.f;
    tile[ty + i][tx] = v;
    if (copy != nullptr && ok) st_any(copy, odt, (long long)r * ldcopy + c, v);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i, r = r0 + tx;
    if (c < cols && r < rows_pad) {
      const int s = r / chunk;
      st_any(out, odt, ((long long)s * cols + c) * chunk + (r - s * chunk), tile[tx][ty + i]);
    }
  }
}
// 64 x 64 tiles, 8 elements (16 bytes of bf16) per access on both sides; needs cols % 8 == 0, ldx % 8 == 0, 16-byte aligned bases.
__device__ __forceinline__ void ld8_any(const void* p, int dt, long long i, float (&v)[8]) {
  if (dt == CST_F32) {
    const float4 a = load4(reinterpret_cast<const float*>(p) + i), b = load4(reinterpret_cast<const float*>(p) + i + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p) + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (dt == CST_BF16) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
        v[2 * j] = __low2float(h); v[2 * j + 1] = __high2float(h);
      } else {
        const __half2 h = *reinterpret_cast<const __half2*>(&w[j]);
        v[2 * j] = __low2float(h); v[2 * j + 1] = __high2float(h);
      }
    }
  }
}
__device__ __forceinline__ void st8_any(void* p, int dt, long long i, const float (&v)[8]) {
  if (dt == CST_F32) {
    store4(reinterpret_cast<float*>(p) + i, make_float4(v[0], v[1], v[2], v[3]));
    store4(reinterpret_cast<float*>(p) + i + 4, make_float4(v[4], v[5], v[6], v[7]));
  } else {
    uint4 u;
    if (dt == CST_BF16) { u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]); }
    else { u.x = pack_f16x2(v[0], v[1]); u.y = pack_f16x2(v[2], v[3]); u.z = pack_f16x2(v[4], v[5]); u.w = pack_f16x2(v[6], v[7]); }
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p) + i) = u;
  }
}
__global__ void __launch_bounds__(256) transpose64_kernel(const void* __restrict__ x, int xdt, long long ldx, int rows, int cols,
                                                          void* __restrict__ out, int odt, int rows_pad, int chunk,
                                                          void* __restrict__ copy, long long ldcopy) {
  __shared__ float tile[64][65];
  pdl_launch_dependents();
  pdl_wait();
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int idx = threadIdx.x + it * 256;
    const int row = idx >> 3, cs = (idx & 7) * 8;
    const int r = r0 + row, c = c0 + cs;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (r < rows && c < cols) {
      ld8_any(x, xdt, (long long)r * ldx + c, v);
      if (copy != nullptr) st8_any(copy, odt, (long long)r * ldcopy + c, v);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) tile[row][cs + j] = v[j];
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int idx = threadIdx.x + it * 256;
    const int cc = idx >> 3, rs = (idx & 7) * 8;
    const int c = c0 + cc, r = r0 + rs;
    if (c < cols && r < rows_pad) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = tile[rs + j][cc];
      const int s = r / chunk;
      st8_any(out, odt, ((long long)s * cols + c) * chunk + (r - s * chunk), v);
    }
  }
}
__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ x, long long n, void* __restrict__ out, int odt) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    st_any(out, odt, i, x[i]);
}

// ---- column sums (bias / LayerNorm-parameter gradients), deterministic: grid.y row slabs write partials, a second launch of
// the same kernel (one slab) adds the partials up.
__global__ void __launch_bounds__(256) colsum_kernel(const void* __restrict__ xv, int xdt, long long ldx, int rows, int cols, int rows_per_slab,
                                                     float* __restrict__ out, long long ldo, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  const int r0 = blockIdx.y * rows_per_slab;
  const int r1 = min(rows, r0 + rows_per_slab);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int r = r0;
  for (; r + 3 < r1; r += 4) {
    s0 += ld_any(xv, xdt, (long long)r * ldx + c); s1 += ld_any(xv, xdt, (long long)(r + 1) * ldx + c);
    s2 += ld_any(xv, xdt, (long long)(r + 2) * ldx + c); s3 += ld_any(xv, xdt, (long long)(r + 3) * ldx + c);
  }
  for (; r < r1; ++r) s0 += ld_any(xv, xdt, (long long)r * ldx + c);
  out[(long long)blockIdx.y * ldo + c] = ((s0 + s1) + (s2 + s3)) * scale;
}

// The same reduction with 16-byte accesses: a thread owns 4 consecutive columns and every 4th row of its slab (4 independent
// accumulators = 4 loads in flight), the 4 row lanes of a CTA (64 column quads x 4) are added through shared memory in a fixed order.
// Measured on the C5 step (ncu launch list, profiles/launches_r02_c5_dropout.csv): the scalar kernel above averaged 11 us per launch
// (156 column sums x 2 levels = 3.4 ms of the 20.5 ms step) for 2 - 8 us of HBM traffic.
__global__ void __launch_bounds__(256) colsum4_kernel(const void* __restrict__ xv, int xdt, long long ldx, int rows, int cols, int rows_per_slab,
                                                      float* __restrict__ out, long long ldo, float scale) {
  __shared__ float4 red[4][64];
  pdl_launch_dependents();
  pdl_wait();
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int c = (blockIdx.x * 64 + tx) * 4;
  const int r0 = blockIdx.y * rows_per_slab;
  const int r1 = min(rows, r0 + rows_per_slab);
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
  if (c < cols) {
    int r = r0 + ty;
    for (; r + 12 < r1; r += 16) {
      const float4 v0 = ld4_any(xv, xdt, (long long)r * ldx + c), v1 = ld4_any(xv, xdt, (long long)(r + 4) * ldx + c);
      const float4 v2 = ld4_any(xv, xdt, (long long)(r + 8) * ldx + c), v3 = ld4_any(xv, xdt, (long long)(r + 12) * ldx + c);
      a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
      a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
      a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
      a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
    }
    for (; r < r1; r += 4) {
      const float4 v0 = ld4_any(xv, xdt, (long long)r * ldx + c);
      a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
    }
  }
  red[ty][tx] = make_float4((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y), (a0.z + a1.z) + (a2.z + a3.z),
                            (a0.w + a1.w) + (a2.w + a3.w));
  __syncthreads();
  if (ty == 0 && c < cols) {
    const float4 p0 = red[0][tx], p1 = red[1][tx], p2 = red[2][tx], p3 = red[3][tx];
    store4(out + (long long)blockIdx.y * ldo + c, make_float4(((p0.x + p1.x) + (p2.x + p3.x)) * scale, ((p0.y + p1.y) + (p2.y + p3.y)) * scale,
                                                              ((p0.z + p1.z) + (p2.z + p3.z)) * scale, ((p0.w + p1.w) + (p2.w + p3.w)) * scale));
  }
}

// ---- activations as separate passes (the training forward keeps the pre-activation z for the backward pass)
//   act 1 GELU (erf), 2 ReLU, 3 GLU on interleaved (value, gate) column pairs: y[:, i] = z[:, 2i] * sigmoid(z[:, 2i+1]) * alpha
__device__ __forceinline__ float gelu_grad_erf(float a) {         // d/dx [x Phi(x)] = Phi(x) + x phi(x)
  const float cdf = 0.5f * (1.0f + erff(a * 0.70710678118654752440f));
  return cdf + a * 0.3989422804014327f * expf(-0.5f * a * a);
}
// 4 output columns per thread (cols_out % 4 == 0, row pitches % 4 == 0): 16-byte accesses on the fp32 side, 8-byte on the 16-bit side
__global__ void __launch_bounds__(256) act_fwd_kernel(int act, const void* __restrict__ z, int zdt, long long ldz, int rows, int cols_out,
                                                      void* __restrict__ y, int ydt, long long ldy, float alpha) {
  pdl_launch_dependents();
  pdl_wait();
  const int c4n = cols_out >> 2;
  const long long total = (long long)rows * c4n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c4n;
    const int c = (int)(i - r * c4n) * 4;
    float4 v;
    if (act == CST_ACT_GLU) {
      const float4 a = ld4_any(z, zdt, r * ldz + 2 * c), b = ld4_any(z, zdt, r * ldz + 2 * c + 4);      // (v0 g0 v1 g1) (v2 g2 v3 g3)
      v = make_float4(a.x / (1.0f + expf(-a.y)), a.z / (1.0f + expf(-a.w)), b.x / (1.0f + expf(-b.y)), b.z / (1.0f + expf(-b.w)));
    } else {
      const float4 a = ld4_any(z, zdt, r * ldz + c);
      if (act == CST_ACT_GELU) v = make_float4(gelu_erf(a.x), gelu_erf(a.y), gelu_erf(a.z), gelu_erf(a.w));
      else if (act == CST_ACT_RELU) v = make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
      else v = a;
    }
    v.x *= alpha; v.y *= alpha; v.z *= alpha; v.w *= alpha;
    st4_any(y, ydt, r * ldy + c, v);
  }
}
__global__ void __launch_bounds__(256) act_bwd_kernel(int act, const void* __restrict__ z, int zdt, long long ldz, const void* __restrict__ dy,
                                                      int dydt, long long ldy, int rows, int cols_out, void* __restrict__ dz, int dzdt,
                                                      long long lddz, float alpha) {
  pdl_launch_dependents();
  pdl_wait();
  const int c4n = cols_out >> 2;
  const long long total = (long long)rows * c4n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c4n;
    const int c = (int)(i - r * c4n) * 4;
    float4 g = ld4_any(dy, dydt, r * ldy + c);
    g.x *= alpha; g.y *= alpha; g.z *= alpha; g.w *= alpha;
    if (act == CST_ACT_GLU) {
      const float4 a = ld4_any(z, zdt, r * ldz + 2 * c), b = ld4_any(z, zdt, r * ldz + 2 * c + 4);
      const float s0 = 1.0f / (1.0f + expf(-a.y)), s1 = 1.0f / (1.0f + expf(-a.w)), s2 = 1.0f / (1.0f + expf(-b.y)), s3 = 1.0f / (1.0f + expf(-b.w));
      st4_any(dz, dzdt, r * lddz + 2 * c, make_float4(g.x * s0, g.x * a.x * s0 * (1.0f - s0), g.y * s1, g.y * a.z * s1 * (1.0f - s1)));
      st4_any(dz, dzdt, r * lddz + 2 * c + 4, make_float4(g.z * s2, g.z * b.x * s2 * (1.0f - s2), g.w * s3, g.w * b.z * s3 * (1.0f - s3)));
    } else {
      float4 o = g;
      if (act == CST_ACT_GELU) {
        const float4 a = ld4_any(z, zdt, r * ldz + c);
        o = make_float4(g.x * gelu_grad_erf(a.x), g.y * gelu_grad_erf(a.y), g.z * gelu_grad_erf(a.z), g.w * gelu_grad_erf(a.w));
      } else if (act == CST_ACT_RELU) {
        const float4 a = ld4_any(z, zdt, r * ldz + c);
        o = make_float4(a.x > 0.f ? g.x : 0.f, a.y > 0.f ? g.y : 0.f, a.z > 0.f ? g.z : 0.f, a.w > 0.f ? g.w : 0.f);
      }
      st4_any(dz, dzdt, r * lddz + c, o);
    }
  }
}

// ---- LayerNorm backward, one warp per row (statistics recomputed from x in fp32 registers, as the forward kernel):
//   xhat = (x - mean) rstd;  g = dy * gamma;  dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat))  (+ dx if accumulate)
// Per-CTA partial sums of dgamma = sum_rows dy * xhat and dbeta = sum_rows dy go to part[blockIdx.x][2][C]; cst_colsum adds them.
template <int NV>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                                                            const float* __restrict__ dy, long long ldy, float* __restrict__ dx,
                                                            long long lddx, float* __restrict__ part, int rows, int accumulate) {
  constexpr int C = NV * 128;
  __shared__ float red[2][C];
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * 8 + warp;
  float4 dg[NV], db[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) dg[i] = db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < rows) {
    float4 v[NV], g[NV];
    const float* xr = x + (long long)row * ldx;
    const float* dr = dy + (long long)row * ldy;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = load4(xr + (lane + 32 * i) * 4);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + 1e-5f);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 d4 = load4(dr + (lane + 32 * i) * 4), gm = load4(gamma + (lane + 32 * i) * 4);
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;          // xhat
      db[i] = d4;
      dg[i] = make_float4(d4.x * v[i].x, d4.y * v[i].y, d4.z * v[i].z, d4.w * v[i].w);
      g[i] = make_float4(d4.x * gm.x, d4.y * gm.y, d4.z * gm.z, d4.w * gm.w);
      sg += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      sgx += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
    }
    const float mg = warp_sum(sg) * (1.0f / C), mgx = warp_sum(sgx) * (1.0f / C);
    float* o = dx + (long long)row * lddx;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 r = make_float4(rstd * (g[i].x - mg - v[i].x * mgx), rstd * (g[i].y - mg - v[i].y * mgx),
                             rstd * (g[i].z - mg - v[i].z * mgx), rstd * (g[i].w - mg - v[i].w * mgx));
      if (accumulate) { const float4 p = load4(o + (lane + 32 * i) * 4); r.x += p.x; r.y += p.y; r.z += p.z; r.w += p.w; }
      store4(o + (lane + 32 * i) * 4, r);
    }
  }
  if (part != nullptr) {                                         // the 8 rows of the CTA, added in a fixed order
    for (int w = 0; w < 8; ++w) {
      if (warp == w) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          float4 a = dg[i], b = db[i];
          if (w > 0) {
            const float4 pa = load4(&red[0][(lane + 32 * i) * 4]), pb = load4(&red[1][(lane + 32 * i) * 4]);
            a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w;
            b.x += pb.x; b.y += pb.y; b.z += pb.z; b.w += pb.w;
          }
          store4(&red[0][(lane + 32 * i) * 4], a);
          store4(&red[1][(lane + 32 * i) * 4], b);
        }
      }
      __syncthreads();
    }
    for (int c = threadIdx.x; c < 2 * C; c += 256) part[(long long)blockIdx.x * 2 * C + c] = red[c / C][c % C];
  }
}

// ---- col2im of a strided convolution's input gradient.  The forward conv is an implicit GEMM over a channels-last matrix
// viewed with row pitch stride*C: output row m reads input rows stride*m .. stride*m + k - 1.  Its A-operand gradient
// dcol [M, k*C] = dY W is scattered back: dx[r, c] = sum over taps t with (r - t) % stride == 0, m = (r - t) / stride in [0, M)
// of dcol[m, t*C + c]   (gather form: deterministic).  Rows of filler / padding carry zero dY, so the flattened form is exact.
__global__ void __launch_bounds__(256) col2im_kernel(const void* __restrict__ dcol, int cdt, long long M, int k, int stride, int C,
                                                     float* __restrict__ dx, long long rows_in, int accumulate) {
  pdl_launch_dependents();
  pdl_wait();
  const int c4n = C >> 2;
  const long long total = rows_in * c4n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c4n;
    const int c = (int)(i - r * c4n) * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < k; ++t) {
      const long long d = r - t;
      if (d < 0) break;
      if (d % stride) continue;
      const long long m = d / stride;
      if (m >= M) continue;
      const float4 v = ld4_any(dcol, cdt, m * (long long)k * C + (long long)t * C + c);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    float* o = dx + r * C + c;
    if (accumulate) { const float4 p = load4(o); s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w; }
    store4(o, s);
  }
}

// ---- dropout (training step; fairseq_dropout.py:16-27 = F.dropout): out = add + x * keep / (1 - p), masks regenerated from
// (seed, site, element index) -- see philox.cuh.  The SAME kernel is the derivative: the gradient of the dropped tensor is
// dropout(dy) with the same seed and site.  Optional `add` (fp32) is the residual the reference adds right after the dropout
// (x = residual + dropout(x), wav2vec2.py TransformerSentenceEncoderLayer / transformer_layer.py:~160-190); optional second output =
// the GEMM-operand copy.  One thread = 4 consecutive columns = one Philox counter.
__global__ void __launch_bounds__(256) dropout_kernel(const void* __restrict__ x, int xdt, long long ldx, const float* __restrict__ add,
                                                      long long ldadd, void* __restrict__ out, int odt, long long ldo,
                                                      void* __restrict__ out2, int o2dt, long long ldo2, int rows, int cols, uint32_t thresh,
                                                      float scale, const unsigned long long* __restrict__ seed_dev, uint32_t site) {
  pdl_launch_dependents();
  pdl_wait();
  const unsigned long long seed = seed_dev[0];
  const int c4n = cols >> 2;
  const long long total = (long long)rows * c4n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c4n;
    const int c = (int)(i - r * c4n) * 4;
    const Philox4 rnd = philox4x32_10(seed, (unsigned long long)i, site);           // group = (r * cols + c) / 4 = i
    float4 v = ld4_any(x, xdt, r * ldx + c);
    v.x = rnd.x >= thresh ? v.x * scale : 0.f;
    v.y = rnd.y >= thresh ? v.y * scale : 0.f;
    v.z = rnd.z >= thresh ? v.z * scale : 0.f;
    v.w = rnd.w >= thresh ? v.w * scale : 0.f;
    if (add != nullptr) { const float4 a = load4(add + r * ldadd + c); v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
    st4_any(out, odt, r * ldo + c, v);
    if (out2 != nullptr) st4_any(out2, o2dt, r * ldo2 + c, v);
  }
}

// ---- row remap with masking: out[seg*out_rps + t + off] (+)= in[seg*in_rps + t + in_off] for t < valid (seg_len[seg] when given,
// else seg_valid); rows t >= valid of the OUTPUT segment range [0, n_rows_out) are zeroed when zero_rest.  The transpose of the
// forward row remaps (zero-padded subsampler operands, masked projection rows).
__global__ void __launch_bounds__(256) rows_remap_kernel(const float* __restrict__ in, long long ldi, int in_rps, int in_off,
                                                         void* __restrict__ out, int odt, long long ldo, int out_rps, int out_off,
                                                         int n_seg, int n_rows, int C, int seg_valid, const int32_t* __restrict__ seg_len,
                                                         int accumulate, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  const int c4n = C >> 2;
  const long long total = (long long)n_seg * n_rows * c4n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long rr = i / c4n;
    const int c = (int)(i - rr * c4n) * 4;
    const int seg = (int)(rr / n_rows), t = (int)(rr - (long long)seg * n_rows);
    const int valid = seg_len ? min(seg_valid, seg_len[seg]) : seg_valid;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < valid) {
      v = load4(in + ((long long)seg * in_rps + t + in_off) * ldi + c);
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    }
    const long long oi = ((long long)seg * out_rps + t + out_off) * ldo + c;
    if (accumulate) { const float4 p = ld4_any(out, odt, oi); v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w; }
    st4_any(out, odt, oi, v);
  }
}

}  // namespace cst

using namespace cst;

static inline unsigned grid_for(long long n, int per = 256) {
  long long b = (n + per - 1) / per;
  if (b < 1) b = 1;
  if (b > 148 * 32) b = 148 * 32;
  return (unsigned)b;
}

extern "C" int cst_transpose(const void* x, int x_dtype, long long ldx, int rows, int cols, void* out, int out_dtype, int rows_pad, int chunk,
                             void* copy, long long ldcopy, void* stream) {
  CST_REQUIRE(x && out && rows > 0 && cols > 0 && rows_pad >= rows, "cst_transpose: bad args rows=%d cols=%d", rows, cols);
  if (chunk <= 0) chunk = rows_pad;
  CST_REQUIRE(rows_pad % chunk == 0, "cst_transpose: rows_pad=%d must be a multiple of chunk=%d", rows_pad, chunk);
  const int xes = x_dtype == CST_F32 ? 4 : 2, oes = out_dtype == CST_F32 ? 4 : 2;
  const bool vec = cols % 8 == 0 && ldx % 8 == 0 && rows_pad % 8 == 0 && chunk % 8 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0 &&
                   (copy == nullptr || (((uintptr_t)copy % 16) == 0 && ldcopy % 8 == 0)) && (ldx * xes) % 16 == 0 && ((long long)chunk * oes) % 16 == 0;
  if (vec) {
    dim3 grid(cdiv(rows_pad, 64), cdiv(cols, 64));
    CST_CHECK_CUDA(launch_k(transpose64_kernel, grid, dim3(256), 0, (cudaStream_t)stream, x, x_dtype, ldx, rows, cols, out, out_dtype, rows_pad,
                            chunk, copy, ldcopy));
    return CST_OK;
  }
  dim3 grid(cdiv(rows_pad, 32), cdiv(cols, 32));
  CST_CHECK_CUDA(launch_k(transpose_kernel, grid, dim3(256), 0, (cudaStream_t)stream, x, x_dtype, ldx, rows, cols, out, out_dtype, rows_pad, chunk,
                          copy, ldcopy));
  return CST_OK;
}

extern "C" int cst_cast(const float* x, long long n, void* out, int out_dtype, void* stream) {
  CST_REQUIRE(x && out && n > 0 && (out_dtype == CST_BF16 || out_dtype == CST_F16), "cst_cast: bad args");
  CST_CHECK_CUDA(launch_k(cast_kernel, dim3(grid_for(n)), dim3(256), 0, (cudaStream_t)stream, x, n, out, out_dtype));
  return CST_OK;
}

// out[c] = scale * sum_r x[r, c]; ws: CST_COLSUM_WS_FLOATS floats of scratch (inputs of >= 256 rows need cols <= that / 64)
extern "C" int cst_colsum(const void* x, int x_dtype, long long ldx, int rows, int cols, float* out, float* ws, float scale, void* stream) {
  CST_REQUIRE(x && out && ws && rows > 0 && cols > 0, "cst_colsum: bad args");
  CST_REQUIRE(rows < 256 || (long long)cols * 64 <= CST_COLSUM_WS_FLOATS, "cst_colsum: cols=%d too wide for the scratch buffer", cols);
  cudaStream_t st = (cudaStream_t)stream;
  // enough row slabs to fill the machine: the kernel is a latency-bound stream of row reads per thread
  int slabs = 1;
  if (rows >= 256) {
    const int want = cdiv(4 * 148, cdiv(cols, 256));                       // ~4 CTAs per SM
    slabs = min(min(want, rows / 32), (int)(CST_COLSUM_WS_FLOATS / cols));
    if (slabs < 1) slabs = 1;
  }
  // 16-byte path: 4 columns per thread (every caller on the training path; the scalar kernel stays for odd shapes / alignments)
  const int esz = x_dtype == CST_F32 ? 4 : 2;
  static const bool use_vec = []() { const char* e = getenv("CST_COLSUM_VEC"); return !(e && e[0] == '0'); }();
  if (use_vec && cols % 4 == 0 && ldx % 4 == 0 && ((uintptr_t)x % (4 * esz)) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)ws % 16) == 0) {
    const int gx = cdiv(cols, 256);
    int sl = 1;
    if (rows >= 256) {
      sl = min(min(cdiv(6 * 148, gx), rows / 32), (int)(CST_COLSUM_WS_FLOATS / cols));
      if (sl < 1) sl = 1;
    }
    const int pr = cdiv(rows, sl);
    sl = cdiv(rows, pr);
    if (sl == 1) {
      CST_CHECK_CUDA(launch_k(colsum4_kernel, dim3(gx, 1), dim3(256), 0, st, x, x_dtype, ldx, rows, cols, rows, out, (long long)cols, scale));
    } else {
      CST_CHECK_CUDA(launch_k(colsum4_kernel, dim3(gx, sl), dim3(256), 0, st, x, x_dtype, ldx, rows, cols, pr, ws, (long long)cols, 1.0f));
      CST_CHECK_CUDA(launch_k(colsum4_kernel, dim3(gx, 1), dim3(256), 0, st, (const void*)ws, (int)CST_F32, (long long)cols, sl, cols, sl, out,
                              (long long)cols, scale));
    }
    return CST_OK;
  }
  const int per = cdiv(rows, slabs);
  slabs = cdiv(rows, per);
  if (slabs == 1) {
    CST_CHECK_CUDA(launch_k(colsum_kernel, dim3(cdiv(cols, 256), 1), dim3(256), 0, st, x, x_dtype, ldx, rows, cols, rows, out, (long long)cols, scale));
  } else {
    CST_CHECK_CUDA(launch_k(colsum_kernel, dim3(cdiv(cols, 256), slabs), dim3(256), 0, st, x, x_dtype, ldx, rows, cols, per, ws, (long long)cols, 1.0f));
    CST_CHECK_CUDA(launch_k(colsum_kernel, dim3(cdiv(cols, 256), 1), dim3(256), 0, st, (const void*)ws, (int)CST_F32, (long long)cols, slabs, cols, slabs,
                            out, (long long)cols, scale));
  }
  return CST_OK;
}

extern "C" int cst_act_fwd(int act, const void* z, int z_dtype, long long ldz, int rows, int cols_out, void* y, int y_dtype, long long ldy,
                           float alpha, void* stream) {
  CST_REQUIRE(z && y && rows > 0 && cols_out > 0 && act >= CST_ACT_NONE && act <= CST_ACT_GLU, "cst_act_fwd: bad args");
  CST_REQUIRE(cols_out % 4 == 0 && ldz % 4 == 0 && ldy % 4 == 0, "cst_act_fwd: cols_out / ldz / ldy must be multiples of 4");
  CST_CHECK_CUDA(launch_k(act_fwd_kernel, dim3(grid_for((long long)rows * cols_out / 4)), dim3(256), 0, (cudaStream_t)stream, act, z, z_dtype, ldz, rows,
                          cols_out, y, y_dtype, ldy, alpha));
  return CST_OK;
}

extern "C" int cst_act_bwd(int act, const void* z, int z_dtype, long long ldz, const void* dy, int dy_dtype, long long ldy, int rows, int cols_out,
                           void* dz, int dz_dtype, long long lddz, float alpha, void* stream) {
  CST_REQUIRE(z && dy && dz && rows > 0 && cols_out > 0 && act >= CST_ACT_NONE && act <= CST_ACT_GLU, "cst_act_bwd: bad args");
  CST_REQUIRE(cols_out % 4 == 0 && ldz % 4 == 0 && ldy % 4 == 0 && lddz % 4 == 0, "cst_act_bwd: cols_out / ld* must be multiples of 4");
  CST_CHECK_CUDA(launch_k(act_bwd_kernel, dim3(grid_for((long long)rows * cols_out / 4)), dim3(256), 0, (cudaStream_t)stream, act, z, z_dtype, ldz, dy,
                          dy_dtype, ldy, rows, cols_out, dz, dz_dtype, lddz, alpha));
  return CST_OK;
}

// part: NULL (no parameter gradients) or cdiv(rows, 8) * 2 * C floats ([block][dgamma | dbeta][C]); reduce with cst_colsum (ldx = 2C).
extern "C" int cst_layernorm_bwd(const float* x, long long ldx, const float* gamma, const float* dy, long long ldy, float* dx, long long lddx,
                                 float* part, int rows, int C, int accumulate, void* stream) {
  CST_REQUIRE(x && gamma && dy && dx && rows > 0 && (C == 512 || C == 768), "cst_layernorm_bwd: bad args rows=%d C=%d", rows, C);
  CST_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0 && lddx % 4 == 0, "cst_layernorm_bwd: leading dims must be multiples of 4");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(cdiv(rows, 8));
  if (C == 512) {
    CST_CHECK_CUDA(launch_k(layernorm_bwd_kernel<4>, grid, dim3(256), 0, st, x, ldx, gamma, dy, ldy, dx, lddx, part, rows, accumulate));
  } else {
    CST_CHECK_CUDA(launch_k(layernorm_bwd_kernel<6>, grid, dim3(256), 0, st, x, ldx, gamma, dy, ldy, dx, lddx, part, rows, accumulate));
  }
  return CST_OK;
}

extern "C" int cst_col2im(const void* dcol, int dcol_dtype, long long M, int k, int stride, int C, float* dx, long long rows_in, int accumulate,
                          void* stream) {
  CST_REQUIRE(dcol && dx && M > 0 && k > 0 && stride > 0 && C % 4 == 0 && rows_in > 0, "cst_col2im: bad args");
  CST_CHECK_CUDA(launch_k(col2im_kernel, dim3(grid_for(rows_in * (C / 4))), dim3(256), 0, (cudaStream_t)stream, dcol, dcol_dtype, M, k, stride, C, dx,
                          rows_in, accumulate));
  return CST_OK;
}

extern "C" int cst_rows_remap(const float* in, long long ldi, int in_rps, int in_off, void* out, int out_dtype, long long ldo, int out_rps,
                              int out_off, int n_seg, int n_rows, int C, int seg_valid, const int32_t* seg_len, int accumulate, float scale,
                              void* stream) {
  CST_REQUIRE(in && out && n_seg > 0 && n_rows > 0 && C % 4 == 0 && ldo % 4 == 0, "cst_rows_remap: bad args");
  CST_CHECK_CUDA(launch_k(rows_remap_kernel, dim3(grid_for((long long)n_seg * n_rows * (C / 4))), dim3(256), 0, (cudaStream_t)stream, in, ldi,
                          in_rps, in_off, out, out_dtype, ldo, out_rps, out_off, n_seg, n_rows, C, seg_valid, seg_len, accumulate, scale));
  return CST_OK;
}

extern "C" int cst_dropout(const void* x, int x_dtype, long long ldx, const float* add, long long ldadd, void* out, int out_dtype, long long ldo,
                           void* out2, int out2_dtype, long long ldo2, int rows, int cols, float p, const unsigned long long* seed,
                           unsigned int site, void* stream) {
  CST_REQUIRE(x && out && seed && rows > 0 && cols > 0 && p >= 0.f && p < 1.f, "cst_dropout: bad args");
  CST_REQUIRE(cols % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 && (!add || ldadd % 4 == 0) && (!out2 || ldo2 % 4 == 0),
              "cst_dropout: cols / row pitches must be multiples of 4");
  CST_CHECK_CUDA(launch_k(dropout_kernel, dim3(grid_for((long long)rows * cols / 4)), dim3(256), 0, (cudaStream_t)stream, x, x_dtype, ldx, add,
                          ldadd, out, out_dtype, ldo, out2, out2_dtype, ldo2, rows, cols, dropout_threshold(p), 1.0f / (1.0f - p), seed,
                          (uint32_t)site));
  return CST_OK;
}

// ---- fused Adam update of one parameter tensor ("optimizer fusion").
// Replaces: fairseq/optim/adam.py:157-224 (the fp32 path FP16Optimizer drives, fairseq/optim/fp16_optimizer.py):
//   g' = g * grad_scale;  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;  denom = sqrt(v) + eps
//   p -= weight_decay * lr * p;  p -= step_size * m / denom,   step_size = lr sqrt(1 - b2^t) / (1 - b1^t) (formed by the caller)
// One pass over p / m / v (fp32) and the all-reduced gradient (fp32 or the bf16 wire format of ddp.GradAllReducer) instead of the ~10
// elementwise torch kernels of the reference's step.
namespace cst {
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const void* __restrict__ g, int gdt, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
                                                   float step_size, float gscale, const float* __restrict__ dyn) {
  pdl_launch_dependents();
  pdl_wait();
  if (dyn != nullptr) { lr = dyn[0]; step_size = dyn[1]; gscale = dyn[2]; }     // per-step values from the device: one CUDA graph for all steps
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = ld_any(g, gdt, i) * gscale;
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    float pi = p[i];
    pi -= wd * lr * pi;
    pi -= step_size * mi / (sqrtf(vi) + eps);
    m[i] = mi; v[i] = vi; p[i] = pi;
  }
}
}  // namespace cst

extern "C" int cst_adam_step(float* p, const void* g, int g_dtype, float* m, float* v, long long n, float lr, float beta1, float beta2,
                             float eps, float weight_decay, float step_size, float grad_scale, const float* dyn, void* stream) {
  CST_REQUIRE(p && g && m && v && n > 0 && (g_dtype == CST_F32 || g_dtype == CST_BF16), "cst_adam_step: bad args");
  CST_CHECK_CUDA(launch_k(cst::adam_kernel, dim3(grid_for(n)), dim3(256), 0, (cudaStream_t)stream, p, g, g_dtype, m, v, n, lr, beta1, beta2, eps,
                          weight_decay, step_size, grad_scale, dyn));
  return CST_OK;
}
