// Grouped positional convolution, "taps stacked in M" formulation (bf16 mode).
//
// posconv_tc.cu computes D[frames, 48] += X_j[frames, 48] * W_j[48, 48]^T per tap j: 384 tcgen05.mma M128xN48xK16 per 128
// frames, and each costs ~73 cycles (the 4 KB A-operand read from shared memory sets the pace, not N = 48).  Here the
// roles are swapped and two taps share one instruction:
//     A = [ W_2p ; W_2p+1 ]  (rows 0..47 / 64..111 = output channels of the even / odd tap, 64-lane K rows)
//     B = panel rows t0 + 2p ... (one row per frame, the SAME row-shifted resident panel as before), N = 256 frames
//     D_lo[co, n] = sum_p W_2p   x[t0 + n + 2p]        -> even-tap part of output frame t0 + n
//     D_hi[co, n] = sum_p W_2p+1 x[t0 + n + 2p]        -> odd-tap part of output frame t0 + n - 1
//     y[t0 + n, co] = D_lo[co, n] + D_hi[co, n + 1]
// 192 instructions of 128 cycles per 255 frames instead of 768 x 73: 2.2x less tensor-pipe time.  The accumulator is
// channel-major (lane = channel), so the epilogue goes through a shared-memory transpose (the dead panel) to reach the
// frame-major [t, 768] output: lo + hi + bias -> GELU -> + residual (wav2vec2.py:823-825), coalesced 192 B per frame.
#include "tc_common.cuh"

namespace cst {

constexpr int P2_TILE_MAX = 255;                     // output frames per CTA (column n + 1 <= 255 must exist for D_hi)
constexpr int P2_PANEL_ROWS = 384;                   // frames t0 .. t0 + 255 + 126
constexpr int P2_PANEL_BYTES = P2_PANEL_ROWS * 128;
constexpr int P2_W_BYTES = 128 * 128;                // one tap pair: 128 rows x 64 lanes bf16
constexpr int P2_WS = 4;
constexpr int P2_SMEM = P2_PANEL_BYTES + P2_WS * P2_W_BYTES + 256;     // exact: two CTAs per SM (the base must be 1024-aligned)
constexpr int P2_THREADS = 192;
constexpr int P2_PAIRS = 64;
constexpr int P2_SLD = 129;                          // transpose tile pitch (floats)

__global__ void __launch_bounds__(P2_THREADS, 2)
posconv_stacked_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                       const float* __restrict__ bias, const float* __restrict__ resid, float* __restrict__ out,
                       int n_rows, int rows_per_seg, int t_pad_rows, int frames_per_tile) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0) { if (threadIdx.x == 0) printf("cst posconv_stacked: shared memory base not 1024-byte aligned\n"); __trap(); }
  const uint32_t sX = base, sW = base + P2_PANEL_BYTES;
  const uint32_t bars = sW + P2_WS * P2_W_BYTES;
  const uint32_t x_full = bars, w_full = bars + 8, w_empty = w_full + 8 * P2_WS, d_full = w_empty + 8 * P2_WS;
  const uint32_t tmem_slot = d_full + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* S = reinterpret_cast<float*>(smem_raw + (sX - smem_u32(smem_raw)));      // transpose tile, reuses the panel
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.y, b = blockIdx.z;
  const int t0 = blockIdx.x * frames_per_tile;
  const int nf = min(frames_per_tile, n_rows - t0);
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    mbar_init(x_full, 1);
    for (int s = 0; s < P2_WS; ++s) { mbar_init(w_full + 8 * s, 1); mbar_init(w_empty + 8 * s, 1); }
    mbar_init(d_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      const int row0 = (b * 16 + g) * t_pad_rows + t0;
      mbar_expect_tx(x_full, P2_PANEL_BYTES);
      for (int r = 0; r < P2_PANEL_ROWS / 128; ++r) tma_load_2d(sX + r * 128 * 128, &tmX, x_full, 0, row0 + r * 128);
      for (int p = 0; p < P2_PAIRS; ++p) {
        const int s = p % P2_WS, u = p / P2_WS;
        mbar_wait(w_empty + 8 * s, (u & 1) ^ 1);
        mbar_expect_tx(w_full + 8 * s, P2_W_BYTES);
        tma_load_2d(sW + s * P2_W_BYTES, &tmW, w_full + 8 * s, 0, (g * P2_PAIRS + p) * 128);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      mbar_wait(x_full, 0);
      for (int p = 0; p < P2_PAIRS; ++p) {
        const int s = p % P2_WS;
        mbar_wait(w_full + 8 * s, (p / P2_WS) & 1);
        tc_fence_after();
        const uint64_t adesc = make_sw128_desc(sW + s * P2_W_BYTES);
        const uint64_t bdesc = make_sw128_desc(sX + (2 * p) * 128);      // panel shifted by 2p rows (base offset 0)
#pragma unroll
        for (int k = 0; k < 3; ++k)                                      // 48 real input channels = 3 x K16
          tc_mma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (p | k) != 0);
        tc_commit(w_empty + 8 * s);
      }
      tc_commit(d_full);
    }
    __syncwarp();
  } else {
    // ===================== epilogue: 4 warps = 4 TMEM lane quarters =====================
    const int q = warp & 3;                              // lanes q*32 .. q*32+31: q 0,1 = even-tap sums, q 2,3 = odd-tap sums
    const bool hi = q >= 2;
    const int et = threadIdx.x - 64;                     // 0..127
    mbar_wait(d_full, 0);                                // all MMAs retired: the panel is dead, S may overwrite it
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const int slot = q * 32 + lane;                      // S column: 0..47 even-tap channel, 64..111 odd-tap channel
    // this thread's three (frame, 4-channel) output slots per 32-frame chunk; their residual loads run one chunk ahead
    // (a load -> use chain per slot made the epilogue 24 dependent global round trips per CTA)
    int nn_[3], c4_[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { const int idx = et + 128 * i; nn_[i] = idx / 12; c4_[i] = (idx - nn_[i] * 12) * 4; }
    const float* rbase = resid + ((long long)b * rows_per_seg + t0) * 768 + g * 48;
    float* obase = out + ((long long)b * rows_per_seg + t0) * 768 + g * 48;
    float4 bb[3], rr[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      bb[i] = __ldg(reinterpret_cast<const float4*>(bias + g * 48 + c4_[i]));
      rr[i] = (nn_[i] < nf) ? *reinterpret_cast<const float4*>(rbase + (long long)nn_[i] * 768 + c4_[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int c0 = 0; c0 < nf; c0 += 32) {
      float v[32];
      const bool last_hi = hi && (c0 + 33 > 256);        // columns c0+1 .. c0+32 would leave the accumulator
      tmem_ld32(trow + (last_hi ? c0 : c0 + (hi ? 1 : 0)), v);
      tmem_ld_wait();
      if (last_hi) {
#pragma unroll
        for (int i = 0; i < 31; ++i) S[i * P2_SLD + slot] = v[i + 1];
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) S[i * P2_SLD + slot] = v[i];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float4 rn[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {                      // next chunk's residual in flight during this chunk's math
        const int n2 = c0 + 32 + nn_[i];
        rn[i] = (n2 < nf) ? *reinterpret_cast<const float4*>(rbase + (long long)n2 * 768 + c4_[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int n = c0 + nn_[i], c4 = c4_[i];
        if (n < nf) {
          const float* sr = S + nn_[i] * P2_SLD;
          float y0 = sr[c4] + sr[64 + c4] + bb[i].x, y1 = sr[c4 + 1] + sr[65 + c4] + bb[i].y;
          float y2 = sr[c4 + 2] + sr[66 + c4] + bb[i].z, y3 = sr[c4 + 3] + sr[67 + c4] + bb[i].w;
          gelu2(y0, y1); gelu2(y2, y3);
          *reinterpret_cast<float4*>(obase + (long long)n * 768 + c4) = make_float4(y0 + rr[i].x, y1 + rr[i].y, y2 + rr[i].z, y3 + rr[i].w);
        }
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) rr[i] = rn[i];
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

}  // namespace cst

extern "C" int cst_posconv_stacked(const void* xg, const void* w2, const float* bias, const float* resid, float* out,
                                   int B, int n_rows, int rows_per_seg, int t_pad_rows, void* stream) {
  using namespace cst;
  CST_REQUIRE(xg && w2 && bias && resid && out && B > 0 && n_rows > 0 && n_rows <= rows_per_seg && t_pad_rows >= n_rows + 128,
              "cst_posconv_stacked: bad args B=%d n_rows=%d rows_per_seg=%d t_pad_rows=%d", B, n_rows, rows_per_seg, t_pad_rows);
  static bool attr_set = false;
  if (!attr_set) {
    CST_CHECK_CUDA(cudaFuncSetAttribute(posconv_stacked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM));
    attr_set = true;
  }
  CUtensorMap tmX, tmW;
  int rc = make_map_2d(&tmX, xg, 64, (long long)B * 16 * t_pad_rows, 64, 64, 128);
  if (rc) return rc;
  rc = make_map_2d(&tmW, w2, 64, 16ll * P2_PAIRS * 128, 64, 64, 128);
  if (rc) return rc;
  const int tiles = cdiv(n_rows, P2_TILE_MAX);
  const int fpt = cdiv(n_rows, tiles);                   // balanced: 750 frames -> 3 x 250
  dim3 grid(tiles, 16, B);
  CST_CHECK_CUDA(launch_k(posconv_stacked_kernel, grid, dim3(P2_THREADS), P2_SMEM, (cudaStream_t)stream, tmX, tmW, bias, resid, out,
                          n_rows, rows_per_seg, t_pad_rows, fpt));
  return CST_OK;
}
