// Micro-benchmark: issue rate of tcgen05.mma M128xN256xK16 (cta_group::1) vs M256xN256xK16 (cta_group::2) with operands
// already resident in shared memory.  nvcc -gencode arch=compute_100a,code=sm_100a -I../../chimera-st_b200/csrc
#include <cstdio>
#include "tc_common.cuh"
using namespace cst;

template <int PAIR>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* out, int n_mma, int bn) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384, bar = base + 16384 + 32768, slot = bar + 8;
  uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(256u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(256u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
    const uint64_t ad = make_sw128_desc(sA), bd = make_sw128_desc(sB);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      if (PAIR) tc_mma_pair_bf16(tmem, ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, i != 0);
      else tc_mma_bf16(tmem, ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, i != 0);
    }
    const long long t1 = clock64();
    if (PAIR) tc_commit_pair(bar, 1); else tc_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[2 * (blockIdx.x >> PAIR) + 0] = t1 - t0;
    out[2 * (blockIdx.x >> PAIR) + 1] = t2 - t0;
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

template <int PAIR>
void run(int grid, int n_mma, int bn) {
  long long* d; cudaMalloc(&d, 8 * 2 * 256); cudaMemset(d, 0, 8 * 2 * 256);
  const int smem = 16384 + 32768 + 1024 + 64;
  cudaFuncSetAttribute(mma_rate_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = PAIR ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, mma_rate_kernel<PAIR>, d, n_mma, bn);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
  }
  long long h[4];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%s grid=%3d N=%3d n_mma=%d: issue %lld clk, issue+drain %lld clk -> %.1f clk/MMA\n", PAIR ? "pair(M256)" : "single(M128)", grid, bn,
         n_mma, h[0], h[1], (double)h[1] / n_mma);
  cudaFree(d);
}

int main() {
  for (int bn : {256, 128}) {
    run<0>(1, 512, bn); run<0>(148, 512, bn);
    run<1>(2, 512, bn); run<1>(148, 512, bn);
  }
  return 0;
}
