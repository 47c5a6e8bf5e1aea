// MUFU.EX2 issue rate per SM: fp32 vs packed f16x2 / bf16x2 (is a packed exponential two results per MUFU slot?)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/ex2_rate tools/micro/ex2_rate.cu && tools/micro/ex2_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k(uint32_t* out, int iters, long long* cyc) {
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0x3c003800u + threadIdx.x * 8 + i;      // f16x2 (1.0, 0.5)-ish / some fp32 value
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { float f = __uint_as_float(a[i] & 0x3fffffffu); asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(f)); a[i] = __float_as_uint(f); }
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(a[i]) : "r"(a[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(a[i]) : "r"(a[i]));
      if (MODE == 3) {       // fp32 pair -> f16x2 -> packed exp (the candidate softmax inner step)
        float f0 = __uint_as_float(a[i] & 0x3fffffffu), f1 = f0 * 0.5f; uint32_t h;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(f1), "f"(f0));
        asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(a[i]) : "r"(h));
      }
    }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  uint32_t* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  const char* names[4] = {"ex2.approx.ftz.f32", "ex2.approx.f16x2", "ex2.approx.ftz.bf16x2", "cvt.f16x2 + ex2.f16x2"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k<0><<<148, 1024>>>(out, iters, cyc);
      if (mode == 1) k<1><<<148, 1024>>>(out, iters, cyc);
      if (mode == 2) k<2><<<148, 1024>>>(out, iters, cyc);
      if (mode == 3) k<3><<<148, 1024>>>(out, iters, cyc);
      cudaDeviceSynchronize();
    }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double instr = 1024.0 * 8 * iters;        // thread-level instructions per SM
    printf("%-26s %8lld cycles  -> %.2f MUFU instr/clk/SM, %.2f exponentials/clk/SM  (%s)\n", names[mode], c, instr / c,
           instr * (mode == 0 ? 1 : 2) / c, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
