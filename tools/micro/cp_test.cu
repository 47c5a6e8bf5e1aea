// Micro-test: does tcgen05.cp.128x256b with a 128B-swizzle descriptor copy a TMA-layout fp32 tile [128 rows x 32 cols]
// (16-byte chunks XOR-swizzled by row & 7) into TMEM as lane = row, column = col?
#include <cstdio>
#include <vector>
#include "tc_common.cuh"
using namespace cst;

__global__ void __launch_bounds__(128, 1) cp_test_kernel(float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar = base + 16384, slot = bar + 8;
  uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {                                           // row r = tid: 32 floats, swizzled 16-byte chunks
    const int r = tid;
    for (int c = 0; c < 32; ++c)
      *reinterpret_cast<float*>(bptr + r * 128 + (((c >> 2) ^ (r & 7)) << 4) + (c & 3) * 4) = (float)(r * 100 + c);
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (tid == 0) {
    for (int k = 0; k < 4; ++k) {
      const uint64_t desc = make_sw128_desc(base + 32 * k);
      asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem + 8 * k), "l"(desc) : "memory");
    }
    tc_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  float v[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int c = 0; c < 32; ++c) out[(warp * 32 + lane) * 32 + c] = v[c];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main() {
  float* d; cudaMalloc(&d, 128 * 32 * 4);
  const int smem = 16384 + 1024 + 64;
  cudaFuncSetAttribute(cp_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cp_test_kernel<<<1, 128, smem>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> h(128 * 32);
  cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r = 0; r < 128; ++r) for (int c = 0; c < 32; ++c) if (h[r * 32 + c] != (float)(r * 100 + c)) ++bad;
  printf("mismatches: %d of 4096\n", bad);
  for (int r : {0, 1, 2, 7, 8, 9, 33, 127}) {
    printf("row %3d:", r);
    for (int c = 0; c < 32; ++c) printf(" %5.0f", h[r * 32 + c]);
    printf("\n");
  }
  return 0;
}
