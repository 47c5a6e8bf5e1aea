// Grouped positional convolution (768 -> 768, k = 128, 16 groups, pad 64, last frame dropped) + bias + GELU
// + residual on tcgen05, with the activation panel RESIDENT in shared memory.
//
// Per CTA: one (utterance b, group g, run of up to 3 consecutive 128-frame tiles).  The packed operand xg[b][g][Tpp][64] (48 channels +
// 16 zero lanes = one 128-byte row per frame, 64 zero frames either side) gives the window of output frame t as
// rows t .. t+127.  The 255 rows a tile needs are TMA-loaded ONCE (32 KB, 128B swizzle); tap j then uses the
// same panel shifted by j rows -- an A descriptor whose start address is simply panel + 128*j.  Measured on
// B200 (tools/dbg_posconv.py): the 128B-swizzle XOR is taken from the ABSOLUTE shared-memory address bits, so a
// row-shifted start needs `base offset` 0; writing the phase (j & 7) into that field gives wrong results for
// every j % 8 != 0.  Only the 6 KB weight slice of each tap streams through an 8-stage TMA ring, so the
// L2 -> SM traffic per tile drops from 128 x 22 KB (generic implicit GEMM) to 32 KB + 128 x 6 KB.
//   D[128 frames, 48 out channels] (fp32, TMEM) = sum_j  X_j[128, 64] * W_gj[48, 64]^T
// Epilogue (thread = frame): y = x + GELU(D + bias)  (wav2vec2.py:823-825), fp32 store.
#include "tc_common.cuh"

namespace cst {

// One CTA = (utterance b, group g, up to PC_R consecutive 128-frame tiles).  The weight slice of a tap (6 KB) is
// streamed ONCE per CTA and used for all its frame tiles (one 48-column accumulator each): 3x less L2 -> SM weight
// traffic (2.4 GB -> 0.8 GB per C2 launch).  Measured effect on the kernel time: none (0.41 -> 0.42 ms) -- the kernel
// is bound by the tcgen05 instruction stream (see the K-step note below), not by the weight stream.
constexpr int PC_BM = 128, PC_TAPS = 128, PC_CG = 48, PC_LANES = 64, PC_WS = 7, PC_R = 3;
constexpr int PC_PANEL_BYTES = (PC_R + 1) * 128 * 128;  // (PC_R + 1) x 128 frames x 128 B
constexpr int PC_W_BYTES = PC_CG * 128;                // 6 KB per tap
constexpr int PC_SMEM = PC_PANEL_BYTES + PC_WS * PC_W_BYTES + 1024 + 256;
constexpr int PC_THREADS = 192;
constexpr int PC_ACC_STRIDE = 64;                      // TMEM columns between the accumulators of consecutive tiles
constexpr int PC_TMEM_COLS = 256;

__global__ void __launch_bounds__(PC_THREADS, 2)
posconv_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const float* __restrict__ bias, const float* __restrict__ resid, float* __restrict__ out,
                  int n_rows, int rows_per_seg, int t_pad_rows, int tiles_per_cta) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sX = base, sW = base + PC_PANEL_BYTES;
  const uint32_t bars = sW + PC_WS * PC_W_BYTES;
  const uint32_t x_full = bars, w_full = bars + 8, w_empty = w_full + 8 * PC_WS, d_full = w_empty + 8 * PC_WS;
  const uint32_t tmem_slot = d_full + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (n_rows + PC_BM - 1) / PC_BM;
  const int tile0 = blockIdx.x * tiles_per_cta;
  const int R = min(tiles_per_cta, n_tiles - tile0);             // frame tiles of this CTA (1..PC_R)
  const int t0 = tile0 * PC_BM;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    mbar_init(x_full, 1);
    for (int s = 0; s < PC_WS; ++s) { mbar_init(w_full + 8 * s, 1); mbar_init(w_empty + 8 * s, 1); }
    mbar_init(d_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)PC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      const int row0 = (b * 16 + g) * t_pad_rows + t0;       // first packed row of the panel
      mbar_expect_tx(x_full, (uint32_t)(R + 1) * 128 * 128);
      for (int r = 0; r <= R; ++r) tma_load_2d(sX + r * 128 * 128, &tmX, x_full, 0, row0 + r * 128);
      for (int j = 0; j < PC_TAPS; ++j) {
        const int s = j % PC_WS, u = j / PC_WS;
        mbar_wait(w_empty + 8 * s, (u & 1) ^ 1);
        mbar_expect_tx(w_full + 8 * s, PC_W_BYTES);
        tma_load_2d(sW + s * PC_W_BYTES, &tmW, w_full + 8 * s, j * PC_LANES, g * PC_CG);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(PC_CG >> 3) << 17) | ((uint32_t)(PC_BM >> 4) << 24);
      mbar_wait(x_full, 0);
      for (int j = 0; j < PC_TAPS; ++j) {
        const int s = j % PC_WS;
        mbar_wait(w_full + 8 * s, (j / PC_WS) & 1);
        tc_fence_after();
        const uint64_t bdesc = make_sw128_desc(sW + s * PC_W_BYTES);
        for (int r = 0; r < R; ++r) {
          // tile r, tap j: panel rows r*128 + j ... (base offset 0: the swizzle follows absolute smem address bits)
          const uint64_t adesc = make_sw128_desc(sX + (r * 128 + j) * 128);
          // 48 real input channels per group = three K=16 steps; lanes 48..63 of every row are zero padding and are
          // skipped.  (Measured: an M128 x N48 x K16 instruction costs ~73 cycles -- the 4 KB A-operand read from
          // shared memory, not N, sets the pace -- so the instruction count is what matters here.)
#pragma unroll
          for (int k = 0; k < PC_CG / 16; ++k)
            tc_mma_bf16(tmem_base + r * PC_ACC_STRIDE, adesc + 2 * k, bdesc + 2 * k, idesc, (j | k) != 0);
        }
        tc_commit(w_empty + 8 * s);
      }
      tc_commit(d_full);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    mbar_wait(d_full, 0);
    tc_fence_after();
    for (int r = 0; r < R; ++r) {
      const int t = t0 + r * PC_BM + q * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + r * PC_ACC_STRIDE;
      const bool ok = t < n_rows;
      const long long off = ((long long)b * rows_per_seg + t) * 768 + g * PC_CG;
#pragma unroll 1
      for (int c = 0; c < PC_CG; c += 16) {
        float acc[16];
        tmem_ld16(trow + c, acc);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + g * PC_CG + c + i));
            const float4 rr = *reinterpret_cast<const float4*>(resid + off + c + i);
            float v0 = acc[i] + bb.x, v1 = acc[i + 1] + bb.y, v2 = acc[i + 2] + bb.z, v3 = acc[i + 3] + bb.w;
            gelu2(v0, v1); gelu2(v2, v3);
            *reinterpret_cast<float4*>(out + off + c + i) = make_float4(v0 + rr.x, v1 + rr.y, v2 + rr.z, v3 + rr.w);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)PC_TMEM_COLS) : "memory");
  }
}

}  // namespace cst

extern "C" int cst_posconv(const void* xg, const void* w, const float* bias, const float* resid, float* out,
                           int B, int n_rows, int rows_per_seg, int t_pad_rows, void* stream) {
  using namespace cst;
  CST_REQUIRE(xg && w && bias && resid && out && B > 0 && n_rows > 0 && n_rows <= rows_per_seg && t_pad_rows >= n_rows + 128,
              "cst_posconv: bad args B=%d n_rows=%d rows_per_seg=%d t_pad_rows=%d", B, n_rows, rows_per_seg, t_pad_rows);
  static bool attr_set = false;
  if (!attr_set) {
    CST_CHECK_CUDA(cudaFuncSetAttribute(posconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_SMEM));
    attr_set = true;
  }
  CUtensorMap tmX, tmW;
  int rc = make_map_2d(&tmX, xg, PC_LANES, (long long)B * 16 * t_pad_rows, PC_LANES, PC_LANES, 128);
  if (rc) return rc;
  rc = make_map_2d(&tmW, w, (long long)PC_TAPS * PC_LANES, 16 * PC_CG, (long long)PC_TAPS * PC_LANES, PC_LANES, PC_CG);
  if (rc) return rc;
  // balanced split of the frame tiles over CTAs of at most PC_R tiles (6 tiles -> 3 + 3, 7 -> 3 + 2 + 2)
  const int n_tiles = cdiv(n_rows, PC_BM);
  const int chunks = cdiv(n_tiles, PC_R);
  const int tiles_per_cta = cdiv(n_tiles, chunks);
  dim3 grid(cdiv(n_tiles, tiles_per_cta), 16, B);
  CST_CHECK_CUDA(launch_k(posconv_tc_kernel, grid, dim3(PC_THREADS), PC_SMEM, (cudaStream_t)stream, tmX, tmW, bias, resid, out,
                          n_rows, rows_per_seg, t_pad_rows, tiles_per_cta));
  return CST_OK;
}
