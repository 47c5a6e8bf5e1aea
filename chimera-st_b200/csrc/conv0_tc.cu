// a1 / K1 for the 16-bit mode: conv0 (1 -> 512, k = 10, s = 5) on tcgen05 + GroupNorm affine + GELU, bulk-stored.
//
// The CUDA-core kernel (conv0.cu) spends 10 of its ~19 issue slots per output on the convolution FMAs.  Here the
// convolution is a [128 frames x 32] x [32 x 512] GEMM per tile, with fp32-level accuracy from a 3-term fp16 split:
//     x = xh + xl,  w = wh + wl  (fp16 each)      A row  = [ xh(10) | xl(10) | xh(10) | 0 0 ]
//     x.w ~= xh.wh + xl.wh + xh.wl                W row  = [ wh(10) | wh(10) | wl(10) | 0 0 ]   (prepared on the host)
// (the dropped xl.wl term is ~2^-22 relative).  K = 32 -> two K=16 MMAs per 256 output channels; the 128 x 512 fp32
// accumulator fills tensor memory as two 256-column stages, so the MMAs of frame tile i+1 / stage h start as soon as
// stage h of tile i is drained.  What is left on the CUDA cores is the part that bounds the kernel: y*scale+shift,
// erf GELU (one MUFU per output, common.cuh gelu2) and the fp16/bf16 pack -- done in the accumulator's layout (lane = frame), staged
// through swizzled shared memory and written by the TMA engine (3-D map [B, rows_per_seg, 512]: the frame axis clips
// the last tile, frames T0 <= t < rows_per_seg are written as zeros, like conv0.cu).
//
// Warps: 0-3 build the A tile (one frame per thread, two stages), 4..4+CT_EPI-1 epilogue (TMEM lane quarter = warp % 4,
// CT_EPI/4 warps per quarter split the 256 columns of a stage), the last warp = MMA issuer + TMEM owner.  Persistent: each CTA
// owns a contiguous range of frame tiles, so the per-utterance scale/shift table (4 KB) is reloaded only at utterance changes.
// CT_EPI = 16 (round 2; 8 in round 1): the epilogue is a latency chain (tcgen05.ld -> FFMA2 -> MUFU GELU -> pack -> STS ->
// async-proxy fence -> bulk store) and with two epilogue warps per scheduler ncu showed issue-active 53 %, XU 43 %, DRAM 39 %:
// 0.52 -> 0.43 ms on the C2 batch (3.07 -> 3.71 TB/s).  Measured and rejected: one 64-column block (128B swizzle) and one bulk
// store per stage instead of two 32-column blocks -- same time (3.69 TB/s), the stores are not what the warps wait for.
#include "tc_common.cuh"

namespace cst {

constexpr int CT_BM = 128, CT_C = 512;
constexpr int CT_W_BYTES = CT_C * 128;              // 512 rows x 64 fp16 (only the first 32 columns are read)
constexpr int CT_A_BYTES = CT_BM * 128;             // per stage
#ifndef CT_EPI
#define CT_EPI 16
#endif
constexpr int CT_CPW = 1024 / CT_EPI;               // accumulator columns per epilogue warp and stage (128 or 64)
constexpr int CT_STG_BYTES = CT_EPI * 4096;         // per-epilogue-warp staging, 2 x 2 KB each
constexpr int CT_SS_BYTES = 4096;                   // 256 channel pairs x {sc0, sc1, sh0, sh1}
constexpr int CT_SMEM = CT_W_BYTES + 2 * CT_A_BYTES + CT_STG_BYTES + CT_SS_BYTES + 256 + 1024;
constexpr int CT_THREADS = (5 + CT_EPI) * 32;
constexpr int CT_MMA_WARP = 4 + CT_EPI;

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <bool OUT_F16>
__global__ void __launch_bounds__(CT_THREADS, 1)
conv0_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO,
                const float* __restrict__ wave, int L, int T0, int tiles_per_utt, int total_tiles,
                const float2* __restrict__ scale_shift) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW = base, sA = sW + CT_W_BYTES, sStg = sA + 2 * CT_A_BYTES, sSS = sStg + CT_STG_BYTES;
  const uint32_t bars = sSS + CT_SS_BYTES;
  const uint32_t w_full = bars, a_full = bars + 8, a_empty = bars + 24, t_full = bars + 40, t_empty = bars + 56;
  const uint32_t tmem_slot = bars + 72;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float4* ss_tab = reinterpret_cast<float4*>(smem_raw + (sSS - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_cta = (total_tiles + gridDim.x - 1) / gridDim.x;
  const int tile_begin = blockIdx.x * per_cta;
  const int tile_end = min(total_tiles, tile_begin + per_cta);
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_full + 8 * s, 4); mbar_init(a_empty + 8 * s, 1);
      mbar_init(t_full + 8 * s, 1); mbar_init(t_empty + 8 * s, CT_EPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  if (warp == CT_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  if (warp < 4) {
    // ===================== A-tile producers: thread r builds frame r of the tile =====================
    const int r = threadIdx.x;
    int i = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++i) {
      const int s = i & 1; const uint32_t ph = (i >> 1) & 1;
      const int b = tile / tiles_per_utt, t0 = (tile - b * tiles_per_utt) * CT_BM;
      const float* x = wave + (size_t)b * L;
      const long long g0 = (long long)(t0 + r) * 5;
      __half hi[10], lo[10];
#pragma unroll
      for (int j = 0; j < 10; ++j) {
        const float v = (g0 + j < L) ? __ldg(x + g0 + j) : 0.f;
        hi[j] = __float2half_rn(v);
        lo[j] = __float2half_rn(v - __half2float(hi[j]));
      }
      __half row[32];
#pragma unroll
      for (int j = 0; j < 10; ++j) { row[j] = hi[j]; row[10 + j] = lo[j]; row[20 + j] = hi[j]; }
      row[30] = __float2half_rn(0.f); row[31] = row[30];
      uint32_t wd[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) wd[j] = (uint32_t)__half_as_ushort(row[2 * j]) | ((uint32_t)__half_as_ushort(row[2 * j + 1]) << 16);
      mbar_wait(a_empty + 8 * s, ph ^ 1);
      const uint32_t dst = sA + s * CT_A_BYTES + r * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) sts128(dst + ((c ^ (r & 7)) << 4), wd[4 * c], wd[4 * c + 1], wd[4 * c + 2], wd[4 * c + 3]);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full + 8 * s);
    }
  } else if (warp == CT_MMA_WARP) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      mbar_expect_tx(w_full, CT_W_BYTES);
      tma_load_2d(sW, &tmW, w_full, 0, 0);
      tma_load_2d(sW + CT_W_BYTES / 2, &tmW, w_full, 0, 256);
      constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(CT_BM >> 4) << 24);   // fp16 x fp16 -> fp32
      mbar_wait(w_full, 0);
      int i = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile, ++i) {
        const int s = i & 1; const uint32_t ph = (i >> 1) & 1;
        mbar_wait(a_full + 8 * s, ph);
        tc_fence_after();
        const uint64_t adesc = make_sw128_desc(sA + s * CT_A_BYTES);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(t_empty + 8 * h, (uint32_t)(i & 1) ^ 1u);
          tc_fence_after();
          const uint64_t bdesc = make_sw128_desc(sW + h * (CT_W_BYTES / 2));
          tc_mma_bf16(tmem_base + h * 256, adesc, bdesc, idesc, 0);
          tc_mma_bf16(tmem_base + h * 256, adesc + 2, bdesc + 2, idesc, 1);
          tc_commit(t_full + 8 * h);
        }
        tc_commit(a_empty + 8 * s);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 4 .. 4+CT_EPI-1) =====================
    const int ew = warp - 4, q = warp & 3, cpart = ew >> 2;
    const int et = threadIdx.x - 128;                              // 0 .. 32*CT_EPI-1
    const uint32_t stage_s = sStg + ew * 4096;
    uint32_t seq = 0;
    int cur_b = -1, i = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++i) {
      const int b = tile / tiles_per_utt, t0 = (tile - b * tiles_per_utt) * CT_BM;
      if (b != cur_b) {                                            // new utterance: reload {scale, shift}, pair-interleaved
        named_bar_sync(1, 32 * CT_EPI);
        if (et < 256) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(scale_shift + (size_t)b * CT_C) + et);   // sc0 sh0 sc1 sh1
          ss_tab[et] = make_float4(v.x, v.z, v.y, v.w);
        }
        named_bar_sync(1, 32 * CT_EPI);
        cur_b = b;
      }
      const bool live = t0 + q * 32 + lane < T0;
      const bool all_live = __all_sync(0xffffffffu, live);       // only an utterance's last tile has dead frames
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        mbar_wait(t_full + 8 * h, (uint32_t)(i & 1));
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + h * 256 + cpart * CT_CPW;
        float acc[2][32];
        tmem_ld32(t_row, acc[0]);
#pragma unroll
        for (int j = 0; j < CT_CPW / 32; ++j) {
          float (&v)[32] = acc[j & 1];
          tmem_ld_wait();
          if (j + 1 < CT_CPW / 32) {
            tmem_ld32(t_row + (j + 1) * 32, acc[(j + 1) & 1]);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty + 8 * h);
          }
          const int c0 = h * 256 + cpart * CT_CPW + j * 32;        // first channel of the chunk
#pragma unroll
          for (int pr = 0; pr < 16; ++pr) {
            const float4 ss = ss_tab[(c0 >> 1) + pr];              // broadcast read
            upk2(ffma2(pk2(v[2 * pr], v[2 * pr + 1]), pk2(ss.x, ss.y), pk2(ss.z, ss.w)), v[2 * pr], v[2 * pr + 1]);
            gelu2(v[2 * pr], v[2 * pr + 1]);
          }
          const uint32_t buf = stage_s + (seq & 1u) * 2048u;
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
          if (!all_live && !live) {                              // warp-divergent only in an utterance's last tile
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = 0.f;
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t wv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
              wv[k] = OUT_F16 ? pack_f16x2(v[8 * c + 2 * k], v[8 * c + 2 * k + 1]) : pack_bf16x2(v[8 * c + 2 * k], v[8 * c + 2 * k + 1]);
            sts128(buf + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4), wv[0], wv[1], wv[2], wv[3]);
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tmO, buf, c0, t0 + q * 32, b);
            bulk_commit();
          }
          ++seq;
        }
      }
    }
    if (lane == 0) bulk_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == CT_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace cst

extern "C" int cst_conv0_apply_tc(const float* wave, int B, int L, const void* w16, const float* scale_shift,
                                  void* out, int out_dtype, int rows_per_seg, void* stream) {
  using namespace cst;
  const int T0 = (L - 10) / 5 + 1;
  CST_REQUIRE(wave && w16 && scale_shift && out && B > 0 && L >= 10 && rows_per_seg >= T0,
              "cst_conv0_apply_tc: bad args B=%d L=%d rows_per_seg=%d (T0=%d)", B, L, rows_per_seg, T0);
  CST_REQUIRE(out_dtype == CST_F16 || out_dtype == CST_BF16, "cst_conv0_apply_tc: out_dtype must be CST_F16 or CST_BF16 (got %d)", out_dtype);
  CST_REQUIRE(((uintptr_t)out % 16) == 0 && ((uintptr_t)w16 % 16) == 0 && ((uintptr_t)scale_shift % 16) == 0,
              "cst_conv0_apply_tc: out / w16 / scale_shift must be 16-byte aligned");
  static bool attr_set = false;
  if (!attr_set) {
    CST_CHECK_CUDA(cudaFuncSetAttribute(conv0_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM));
    CST_CHECK_CUDA(cudaFuncSetAttribute(conv0_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM));
    attr_set = true;
  }
  CUtensorMap tmW, tmO;
  int rc = make_map_2d(&tmW, w16, 64, CT_C, 64, 64, 256);
  if (rc) return rc;
  {
    EncodeTiledFn enc = get_encode_fn();
    CST_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[3] = {(cuuint64_t)CT_C, (cuuint64_t)rows_per_seg, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)CT_C * 2, (cuuint64_t)rows_per_seg * CT_C * 2};
    cuuint32_t box[3] = {32, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CST_REQUIRE(r == CUDA_SUCCESS, "cst_conv0_apply_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  }
  const int tiles_per_utt = cdiv(rows_per_seg, CT_BM);
  const long long total = (long long)tiles_per_utt * B;
  CST_REQUIRE(total < (1ll << 31), "cst_conv0_apply_tc: too many tiles");
  int dev = 0, sms = 0;
  CST_CHECK_CUDA(cudaGetDevice(&dev));
  CST_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = (int)(total < sms ? total : sms);
  const float2* ss = reinterpret_cast<const float2*>(scale_shift);
  if (out_dtype == CST_F16)
    CST_CHECK_CUDA(launch_k(conv0_tc_kernel<true>, dim3(grid), dim3(CT_THREADS), CT_SMEM, (cudaStream_t)stream, tmW, tmO, wave, L, T0,
                            tiles_per_utt, (int)total, ss));
  else
    CST_CHECK_CUDA(launch_k(conv0_tc_kernel<false>, dim3(grid), dim3(CT_THREADS), CT_SMEM, (cudaStream_t)stream, tmW, tmO, wave, L, T0,
                            tiles_per_utt, (int)total, ss));
  return CST_OK;
}
