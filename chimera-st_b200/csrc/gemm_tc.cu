// bf16 GEMM on the 5th-gen tensor cores: TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory
// ring -> tcgen05.mma (cta_group::1, M=128, N=BN, K=16 per instruction) accumulating fp32 in tensor
// memory -> tcgen05.ld epilogue (bias / GELU / ReLU / GLU / alpha / residual / row masking).
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc),
// warps 2-5 = epilogue; two TMEM accumulator stages so tile i's epilogue overlaps tile i+1's MMAs.
//
// Implicit-GEMM convolutions need no im2col and no overlapping tensor maps: the activation is a plain
// row-major [rows, lda] matrix and the k-th 64-wide K block of output row m lives at
// (row m + (64k)/lda, column (64k)%lda) -- for a stride-s conv over channels-last data lda = s*C, so
// the TMA coordinates simply walk into the following rows.
#include <cuda.h>
#include "gemm_common.cuh"

namespace cst {

constexpr int TC_BM = 128, TC_BK = 64;
constexpr int TC_THREADS = 192;

template <int BN> struct TcCfg {
  static constexpr int BN_PAD = (BN <= 64) ? 64 : (BN <= 128 ? 128 : 256);   // TMEM columns per stage
  static constexpr int TMEM_COLS = 2 * BN_PAD;                               // power of two >= 32
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) break;
    if (clock64() - t0 > 4000000000LL) {   // ~2 s: a protocol bug must not hang the GPU box
      printf("cst gemm_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzled operand tile (rows of 64 bf16 = 128 B, 8-row atoms of 1024 B):
// start>>4 | LBO(ignored, 1)<<16 | SBO(1024>>4)<<32 | version 1<<46 | SWIZZLE_128B(2)<<61
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmDev p, int m_tiles, int n_tiles, int total_tiles, int a_wrap) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + STAGES * Cfg::A_BYTES;
  const uint32_t bars = base + STAGES * Cfg::STAGE_BYTES;        // 8-byte mbarriers
  const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES;
  const uint32_t tfull_bar = bars + 16 * STAGES, tempty_bar = tfull_bar + 16;
  const uint32_t tmem_slot = tempty_bar + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = p.K / TC_BK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + 8 * s, 1); mbar_init(tempty_bar + 8 * s, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nb = tile % n_tiles; const int r = tile / n_tiles;
        const int mb = r % m_tiles; const int z = r / m_tiles;
        const int zo = z / p.nb_inner, zi = z - zo * p.nb_inner;
        const long long a_row0 = (zo * p.a_bs_outer + zi * p.a_bs_inner) / p.lda + (long long)mb * TC_BM;
        const int b_row0 = zi * p.N + nb * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty_bar + 8 * s, ph ^ 1);
          mbar_expect_tx(full_bar + 8 * s, Cfg::STAGE_BYTES);
          const int kk = kb * TC_BK;
          int acol = kk, arow_add = 0;
          if (a_wrap) { arow_add = kk / (int)p.lda; acol = kk - arow_add * (int)p.lda; }
          tma_load_2d(sA + s * Cfg::A_BYTES, &tmA, full_bar + 8 * s, acol, (int)(a_row0 + arow_add));
          tma_load_2d(sB + s * Cfg::B_BYTES, &tmB, full_bar + 8 * s, kk, b_row0);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int s = 0; uint32_t ph = 0; int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
        mbar_wait(tempty_bar + 8 * as, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * Cfg::BN_PAD;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar + 8 * s, ph);
          tc_fence_after();
          const uint64_t adesc = make_sw128_desc(sA + s * Cfg::A_BYTES);
          const uint64_t bdesc = make_sw128_desc(sB + s * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)     // +32 B per K=16 step inside the 128 B swizzle row
            tc_mma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          tc_commit(empty_bar + 8 * s);            // frees the smem stage once these MMAs retire
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        tc_commit(tfull_bar + 8 * as);             // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1) =====================
    const int q = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int nb = tile % n_tiles; const int r = tile / n_tiles;
      const int mb = r % m_tiles; const int z = r / m_tiles;
      const int zo = z / p.nb_inner, zi = z - zo * p.nb_inner;
      const float* bias = p.bias ? p.bias + (long long)zi * p.bias_bs_inner : nullptr;
      const long long c_off = zo * p.c_bs_outer + zi * p.c_bs_inner;
      const long long r_off = zo * p.r_bs_outer + zi * p.r_bs_inner;
      const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
      mbar_wait(tfull_bar + 8 * as, aph);
      tc_fence_after();
      const int m = mb * TC_BM + q * 32 + lane;
      const RowInfo ri = row_info(p, m, zo);
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * Cfg::BN_PAD;
#pragma unroll 1
      for (int c = 0; c < BN; c += 16) {
        float acc[16];
        tmem_ld16(t_row + c, acc);
        tmem_ld_wait();
        const int n = nb * BN + c;
        epilogue8(p, ri, n, bias, c_off, r_off, *reinterpret_cast<const float(*)[8]>(&acc[0]));
        epilogue8(p, ri, n + 8, bias, c_off, r_off, *reinterpret_cast<const float(*)[8]>(&acc[8]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * as);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
  }
}

// ---- host side -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static int make_map_2d(CUtensorMap* map, const void* base, long long inner, long long outer, long long pitch_elems,
                       int box_inner, int box_outer) {
  EncodeTiledFn enc = get_encode_fn();
  CST_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CST_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): inner=%lld outer=%lld pitch=%lld box=%dx%d base=%p",
              (int)r, inner, outer, pitch_elems, box_inner, box_outer, base);
  return CST_OK;
}

template <int BN>
static int launch_tc(const cst_gemm_params& hp, const GemmDev& p, int nz, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    CST_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int a_wrap = hp.K > hp.lda ? 1 : 0;
  if (a_wrap) CST_REQUIRE(hp.lda % TC_BK == 0, "cst_gemm(bf16): wrapped K needs lda %% 64 == 0 (lda=%lld)", hp.lda);
  CST_REQUIRE(hp.a_bs_outer % hp.lda == 0 && hp.a_bs_inner % hp.lda == 0, "cst_gemm(bf16): A batch strides must be multiples of lda");
  CST_REQUIRE(nz == 1 || hp.w_bs_inner == (long long)hp.N * hp.K || hp.nb_inner == 1,
              "cst_gemm(bf16): batched W must be densely packed [nb_inner*N, K]");
  CST_REQUIRE(((uintptr_t)hp.A % 16) == 0 && ((uintptr_t)hp.W % 16) == 0, "cst_gemm(bf16): A/W must be 16-byte aligned");
  const long long a_rows_total = ((hp.nb_outer - 1) * hp.a_bs_outer + (hp.nb_inner - 1) * hp.a_bs_inner) / hp.lda + hp.a_rows;
  const long long a_inner = a_wrap ? hp.lda : hp.K;
  CUtensorMap tmA, tmB;
  int rc = make_map_2d(&tmA, hp.A, a_inner, a_rows_total, hp.lda, TC_BK, TC_BM);
  if (rc) return rc;
  rc = make_map_2d(&tmB, hp.W, hp.K, (long long)hp.nb_inner * hp.N, hp.K, TC_BK, BN);
  if (rc) return rc;
  const int m_tiles = cdiv(hp.M, TC_BM), n_tiles = cdiv(hp.N, BN);
  const long long total = (long long)m_tiles * n_tiles * nz;
  CST_REQUIRE(total < (1ll << 31), "cst_gemm(bf16): too many tiles");
  int dev = 0, sms = 0;
  CST_CHECK_CUDA(cudaGetDevice(&dev));
  CST_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = (int)(total < sms ? total : sms);
  gemm_tc_kernel<BN><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmB, p, m_tiles, n_tiles, (int)total, a_wrap);
  CST_LAUNCH_CHECK();
  return CST_OK;
}

int launch_gemm_tc(const cst_gemm_params& hp, const GemmDev& p, int nz, cudaStream_t st) {
  if (hp.N % 256 == 0) return launch_tc<256>(hp, p, nz, st);
  if (hp.N % 128 == 0) return launch_tc<128>(hp, p, nz, st);
  if (hp.N == 48) return launch_tc<48>(hp, p, nz, st);
  if (hp.N % 64 == 0) return launch_tc<64>(hp, p, nz, st);
  CST_REQUIRE(false, "cst_gemm(bf16): N=%d unsupported (multiple of 64, or 48)", hp.N);
  return CST_ERR_ARG;
}

}  // namespace cst
