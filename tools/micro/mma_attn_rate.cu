// tcgen05.mma rates of the shapes the attention kernel issues (M128, K16 per instruction, cta_group::1):
//   SS N=128 (S = Q K^T), SS N=64 K-major B, SS N=64 MN-major B (V straight from its TMA tile), TS N=64 (A = P in TMEM),
//   and TS with wider N (two / four heads' worth of V columns per instruction).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I chimera-st_b200/csrc -I include -o tools/micro/mma_attn_rate tools/micro/mma_attn_rate.cu
#include <cstdio>
#include "tc_common.cuh"
using namespace cst;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: SS K-major B; 1: SS MN-major B; 2: TS MN-major B
__global__ void __launch_bounds__(128, 1) k(long long* out, int n_mma, int bn, int mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384, bar = base + 16384 + 65536, slot = bar + 8;
  uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < (16384 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (threadIdx.x == 0) {
    const uint32_t mn = mode >= 1 ? (1u << 16) : 0u;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | mn | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t ad = make_sw128_desc(sA), bd = make_sw128_desc(sB);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      if (mode == 2) mma_ts(tmem, tmem + 256 + 8 * (i & 7), bd + 128 * (i & 7), idesc, i != 0);
      else if (mode == 1) tc_mma_bf16(tmem, ad + 2 * (i & 3), bd + 128 * (i & 7), idesc, i != 0);
      else tc_mma_bf16(tmem, ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, i != 0);
    }
    const long long t1 = clock64();
    tc_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[2 * blockIdx.x] = t1 - t0;
    out[2 * blockIdx.x + 1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main() {
  long long* d; cudaMalloc(&d, 8 * 2 * 256);
  const int smem = 16384 + 65536 + 1024 + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[3] = {"SS  K-major B", "SS MN-major B", "TS MN-major B"};
  for (int mode = 0; mode < 3; ++mode)
    for (int bn : {256, 128, 64, 32}) {
      if (mode >= 1 && bn > 64 && false) continue;
      for (int rep = 0; rep < 2; ++rep) {
        k<<<148, 128, smem>>>(d, 512, bn, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s N=%d error: %s\n", names[mode], bn, cudaGetErrorString(e)); return 1; }
      }
      long long h[2];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%s N=%3d: issue %6lld clk, issue+drain %6lld clk -> %.1f clk per M128 x N x K16 MMA\n", names[mode], bn, h[0], h[1], h[1] / 512.0);
    }
  return 0;
}
