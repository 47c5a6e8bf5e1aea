// Shared helpers for the chimera-st_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include "../../include/chimera_st_b200.h"

namespace cst {

// ---- error plumbing: every extern "C" entry returns 0 or records a message -------------
void set_error(const char* fmt, ...);
#define CST_CHECK_CUDA(call)                                                             \
  do {                                                                                    \
    cudaError_t _e = (call);                                                              \
    if (_e != cudaSuccess) {                                                              \
      cst::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return CST_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)
#define CST_REQUIRE(cond, ...)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      cst::set_error(__VA_ARGS__);                                                        \
      return CST_ERR_ARG;                                                                 \
    }                                                                                     \
  } while (0)
#define CST_LAUNCH_CHECK() CST_CHECK_CUDA(cudaGetLastError())

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL): every kernel of the path is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, calls pdl_launch_dependents() first thing and
// pdl_wait() before its first global-memory access, so the next kernel's CTAs are already scheduled (and
// through their barrier/TMEM/descriptor prologue) when the previous kernel drains: no launch bubble between
// the ~350 kernels of a forward pass.  Measured on B200 (c3 shapes): no gain over plain stream order (16.7 vs
// 17.0 ms/step) because launches are already queued ahead of the GPU, so it is OFF unless CST_PDL=1
// (without the attribute the device instructions are no-ops).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("CST_PDL"); return e && e[0] == '1'; }();
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---- device math ----------------------------------------------------------------------
// exact (erf) GELU, as torch.nn.functional.gelu / nn.GELU (wav2vec2.py:731, gelu.py:25)
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- typed loads/stores (activations are f32 or bf16) ---------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// load 4 consecutive elements as floats (16B-aligned for f32, 8B-aligned for bf16)
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x), b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void store4(__half* p, float4 v) {
  uint2 u; u.x = pack_f16x2(v.x, v.y); u.y = pack_f16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__nv_bfloat16* p, float4 v) {
  uint2 u; u.x = pack_bf16x2(v.x, v.y); u.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}

// ---- packed fp32x2 helpers (Blackwell FFMA2/FMUL2/FADD2: two fp32 lanes per issue slot) ----------------
__device__ __forceinline__ uint64_t pk2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// erf-GELU of two values for the 16-bit paths, one MUFU per value:
//   GELU(x) = max(x, 0) - |x|/2 * erfc(|x| / sqrt 2),     erfc(s / sqrt 2) = 2^q(s)
// q = degree-7 minimax fit of log2 erfc(s / sqrt 2) on [0, 6] (|error| < 5.7e-6 in log2, i.e. 4e-6 RELATIVE in erfc, so the
// negative tail keeps its relative accuracy; beyond s = 6 the polynomial keeps falling and 2^q flushes to 0).  Checked
// against the exact function in fp32 arithmetic on [-8, 8]: max |error| 6.4e-7, max relative error 5.5e-6
// (the Abramowitz-Stegun 7.1.26 form used before needed a reciprocal as well: 2 MUFU per value, and the MUFU pipe is
// what bounds the conv0 / fc1 epilogues).  ~7.5 issue slots per value with packed FFMA2.
__device__ __forceinline__ void gelu2(float& x0, float& x1) {
  const uint64_t s = pk2(fabsf(x0), fabsf(x1));
  uint64_t q = ffma2(s, pk2(-1.889626233e-06f, -1.889626233e-06f), pk2(6.268140101e-05f, 6.268140101e-05f));
  q = ffma2(q, s, pk2(-9.388679333e-04f, -9.388679333e-04f));
  q = ffma2(q, s, pk2(8.539461332e-03f, 8.539461332e-03f));
  q = ffma2(q, s, pk2(-5.402068712e-02f, -5.402068712e-02f));
  q = ffma2(q, s, pk2(-4.584097768e-01f, -4.584097768e-01f));
  q = ffma2(q, s, pk2(-1.151269147e+00f, -1.151269147e+00f));
  q = ffma2(q, s, pk2(5.659667913e-06f, 5.659667913e-06f));
  float q0, q1;
  upk2(q, q0, q1);
  const uint64_t e = pk2(mufu_ex2(q0), mufu_ex2(q1));
  const uint64_t nh = fmul2(s, pk2(-0.5f, -0.5f));               // -|x| / 2
  upk2(ffma2(nh, e, pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f))), x0, x1);
}


}  // namespace cst
