// bf16 GEMM on the 5th-gen tensor cores: TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory
// ring -> tcgen05.mma (cta_group::1, M=128, N=BN, K=16 per instruction) accumulating fp32 in tensor
// memory -> tcgen05.ld epilogue (bias / GELU / ReLU / GLU / alpha / residual / row masking).
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc),
// warps 2-9 = epilogue (two warps per TMEM lane quarter, each owning half of the tile's columns); two TMEM
// accumulator stages so tile i's epilogue overlaps tile i+1's MMAs.  The epilogue transposes each 32x32
// accumulator block through a per-warp shared-memory patch so that bias/residual loads and the C stores are
// row-contiguous (a warp instruction covers 4 rows x 128 B) instead of one row per lane.
//
// Implicit-GEMM convolutions need no im2col and no overlapping tensor maps: the activation is a plain
// row-major [rows, lda] matrix and the k-th 64-wide K block of output row m lives at
// (row m + (64k)/lda, column (64k)%lda) -- for a stride-s conv over channels-last data lda = s*C, so
// the TMA coordinates simply walk into the following rows.
#include "gemm_common.cuh"
#include "tc_common.cuh"

namespace cst {

constexpr int TC_BM = 128, TC_BK = 64;
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_PATCH_LD = 32;                               // floats; unpadded rows, 16-byte chunks XOR-swizzled by (row & 7)
constexpr int TC_PATCH_BYTES = 32 * TC_PATCH_LD * 4;           // one 32x32 fp32 block per epilogue warp (4 KB)

template <int BN> struct TcCfg {
  static constexpr int BN_PAD = (BN <= 64) ? 64 : (BN <= 128 ? 128 : 256);   // TMEM columns per stage
  static constexpr int TMEM_COLS = 2 * BN_PAD;                               // power of two >= 32
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);   // bytes in flight hide the ~1 us L2->smem latency
  static constexpr int EPI_BYTES = TC_EPI_WARPS * TC_PATCH_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// ---- lean tile epilogue: compile-time output dtype (0 f32, 1 bf16, 2 f16) and residual flag, alpha == 1, no
// per-segment zeroing.  Everything that does not depend on the column chunk (row pointers, store predicates) is
// resolved once per tile by the caller, so the inner loop is TMEM -> smem transpose -> packed math -> 128-bit stores
// with no address arithmetic, dtype branches or mask tests per element.
template <int BN, int ACT, int CD, bool RES>
__device__ __forceinline__ void epi_tile_fast(const GemmDev& p, float* patch, uint32_t t_row, int nb, int lane, int chalf,
                                              const float* bias, uint8_t* const (&cp)[8], const float* const (&rp)[8],
                                              uint32_t st_mask) {
  constexpr int NCH = (BN + 31) / 32, CH_PER = (NCH + 1) / 2;
  constexpr int ES = CD == 0 ? 4 : 2;
  const int lr = lane >> 3, lc = (lane & 7) * 4;
#pragma unroll 1
  for (int ch = chalf * CH_PER; ch < NCH && ch < (chalf + 1) * CH_PER; ++ch) {
    const int c = ch * 32;
    const int n = nb * BN + c + lc;
    const bool col_ok = (c + lc < BN) && (n < p.N);
    float acc[32];
    if (BN - c >= 32) {
      tmem_ld32(t_row + c, acc);
    } else {
      tmem_ld16(t_row + c, acc);
#pragma unroll
      for (int i = 16; i < 32; ++i) acc[i] = 0.f;
    }
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias && col_ok) b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
    float4 res[RES ? 8 : 1];
    if (RES) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_ok && ((st_mask >> i) & 1)) res[i] = *reinterpret_cast<const float4*>(rp[i] + n);
      }
    }
    tmem_ld_wait();
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<float4*>(&patch[lane * TC_PATCH_LD + 4 * (i ^ (lane & 7))]) = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
    __syncwarp();
    if (col_ok) {
      const uint64_t b01 = pk2(b4.x, b4.y), b23 = pk2(b4.z, b4.w);
#pragma unroll
      for (int i0 = 0; i0 < 8; i0 += 4) {
        float v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int rr = (i0 + u) * 4 + lr;
          const float4 a4 = *reinterpret_cast<const float4*>(&patch[rr * TC_PATCH_LD + 4 * ((lc >> 2) ^ (rr & 7))]);
          upk2(fadd2(pk2(a4.x, a4.y), b01), v[u][0], v[u][1]);
          upk2(fadd2(pk2(a4.z, a4.w), b23), v[u][2], v[u][3]);
        }
        if (ACT == CST_ACT_GELU) {
#pragma unroll
          for (int u = 0; u < 4; ++u) { gelu2(v[u][0], v[u][1]); gelu2(v[u][2], v[u][3]); }
        } else if (ACT == CST_ACT_RELU) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            v[u][0] = fmaxf(v[u][0], 0.f); v[u][1] = fmaxf(v[u][1], 0.f);
            v[u][2] = fmaxf(v[u][2], 0.f); v[u][3] = fmaxf(v[u][3], 0.f);
          }
        }
        if (RES) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            upk2(fadd2(pk2(v[u][0], v[u][1]), pk2(res[i0 + u].x, res[i0 + u].y)), v[u][0], v[u][1]);
            upk2(fadd2(pk2(v[u][2], v[u][3]), pk2(res[i0 + u].z, res[i0 + u].w)), v[u][2], v[u][3]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u;
          if ((st_mask >> i) & 1) {
            uint8_t* o = cp[i] + (size_t)n * ES;
            if (CD == 0) *reinterpret_cast<float4*>(o) = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
            else if (CD == 1) *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16x2(v[u][0], v[u][1]), pack_bf16x2(v[u][2], v[u][3]));
            else *reinterpret_cast<uint2*>(o) = make_uint2(pack_f16x2(v[u][0], v[u][1]), pack_f16x2(v[u][2], v[u][3]));
          }
        }
      }
    }
  }
}

template <int BN, int ACT>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmDev p, int m_tiles, int n_tiles, int total_tiles, int a_wrap, int ab_f16) {
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
  uint8_t* epi_base = smem_raw + (bars + 256 - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = p.K / TC_BK;
  pdl_launch_dependents();            // the next kernel may start its prologue; it still waits for our completion

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + 8 * s, 1); mbar_init(tempty_bar + 8 * s, TC_EPI_WARPS); }
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
  pdl_wait();                         // everything above overlapped the previous kernel; global memory from here on

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
      // fp32 accumulate; operand formats bf16 (1) or fp16 (0) in bits 7-9 / 10-12
      const uint32_t fmt = ab_f16 ? 0u : 1u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
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
    // ===================== epilogue (warps 2..9) =====================
    // TMEM lane quarter q = warp % 4 (hardware rule); the two warps of a quarter split the BN columns.
    const int ew = warp - 2;
    const int q = warp & 3;
    const int chalf = ew >> 2;                                 // 0: first half of the columns, 1: second half
    float* patch = reinterpret_cast<float*>(epi_base + ew * TC_PATCH_BYTES);
    constexpr int NCH = (BN + 31) / 32;                        // 32-column chunks in the tile
    constexpr int CH_PER = (NCH + 1) / 2;
    const int lr = lane >> 3, lc = (lane & 7) * 4;             // row-contiguous domain: 4 rows x 8 float4 per pass
    const bool c_16 = p.c_dtype != CST_F32, c_f16 = p.c_dtype == CST_F16;
    const uint64_t alpha2 = pk2(p.alpha, p.alpha);
    const bool fast = (ACT != CST_ACT_GLU) && p.alpha == 1.0f && p.seg_len == nullptr;
    const int esz = c_16 ? 2 : 4;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int nb = tile % n_tiles; const int r = tile / n_tiles;
      const int mb = r % m_tiles; const int z = r / m_tiles;
      const int zo = z / p.nb_inner, zi = z - zo * p.nb_inner;
      const float* bias = p.bias ? p.bias + (long long)zi * p.bias_bs_inner : nullptr;
      const long long c_off = zo * p.c_bs_outer + zi * p.c_bs_inner;
      const long long r_off = zo * p.r_bs_outer + zi * p.r_bs_inner;
      const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
      // this lane's 8 rows are the same for every column chunk of the tile: resolve them once
      long long crow[8], rrow[8];
      uint32_t st_mask = 0, z_mask = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const RowInfo ri = row_info(p, mb * TC_BM + q * 32 + i * 4 + lr, zo);
        crow[i] = c_off + ri.out_row * p.ldc;
        rrow[i] = r_off + ri.out_row * p.ldr;
        st_mask |= (ri.store ? 1u : 0u) << i;
        z_mask |= (ri.zero ? 1u : 0u) << i;
      }
      mbar_wait(tfull_bar + 8 * as, aph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * Cfg::BN_PAD;
      if (fast) {
        uint8_t* cp[8];
        const float* rp[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          cp[i] = reinterpret_cast<uint8_t*>(p.C) + crow[i] * esz;
          rp[i] = p.residual + rrow[i];
        }
        if (p.residual) {
          if (!c_16) epi_tile_fast<BN, ACT, 0, true>(p, patch, t_row, nb, lane, chalf, bias, cp, rp, st_mask);
          else if (!c_f16) epi_tile_fast<BN, ACT, 1, true>(p, patch, t_row, nb, lane, chalf, bias, cp, rp, st_mask);
          else epi_tile_fast<BN, ACT, 2, true>(p, patch, t_row, nb, lane, chalf, bias, cp, rp, st_mask);
        } else {
          if (!c_16) epi_tile_fast<BN, ACT, 0, false>(p, patch, t_row, nb, lane, chalf, bias, cp, rp, st_mask);
          else if (!c_f16) epi_tile_fast<BN, ACT, 1, false>(p, patch, t_row, nb, lane, chalf, bias, cp, rp, st_mask);
          else epi_tile_fast<BN, ACT, 2, false>(p, patch, t_row, nb, lane, chalf, bias, cp, rp, st_mask);
        }
      } else
#pragma unroll 1
      for (int ch = chalf * CH_PER; ch < NCH && ch < (chalf + 1) * CH_PER; ++ch) {
        const int c = ch * 32;
        const int n = nb * BN + c + lc;
        const bool col_ok = (c + lc < BN) && (n < p.N);
        const int nc = (ACT == CST_ACT_GLU) ? (n >> 1) : n;   // output column
        float acc[32];
        if (BN - c >= 32) {
          tmem_ld32(t_row + c, acc);
        } else {                                              // BN = 48: last chunk holds 16 columns
          tmem_ld16(t_row + c, acc);
#pragma unroll
          for (int i = 16; i < 32; ++i) acc[i] = 0.f;
        }
        // independent global loads first (bias, residual rows): their latency overlaps the TMEM load
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias && col_ok) b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
        float4 res[8];
        if (p.residual) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col_ok && ((st_mask >> i) & 1)) {
              if (ACT == CST_ACT_GLU) { const float2 t2 = *reinterpret_cast<const float2*>(p.residual + rrow[i] + nc); res[i].x = t2.x; res[i].y = t2.y; }
              else res[i] = *reinterpret_cast<const float4*>(p.residual + rrow[i] + nc);
            }
          }
        }
        tmem_ld_wait();
        __syncwarp();                                          // previous pass finished reading the patch
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4*>(&patch[lane * TC_PATCH_LD + 4 * (i ^ (lane & 7))]) = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
        __syncwarp();
        if (col_ok) {
          const uint64_t b01 = pk2(b4.x, b4.y), b23 = pk2(b4.z, b4.w);
          // 4 rows per batch: the 16 element chains (bias, activation, alpha, residual) are independent, so the
          // two epilogue warps of an SM sub-partition keep the FMA/MUFU pipes busy instead of waiting on one chain
#pragma unroll
          for (int i0 = 0; i0 < 8; i0 += 4) {
            float v[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int rr = (i0 + u) * 4 + lr;
              const float4 a4 = *reinterpret_cast<const float4*>(&patch[rr * TC_PATCH_LD + 4 * ((lc >> 2) ^ (rr & 7))]);
              upk2(fadd2(pk2(a4.x, a4.y), b01), v[u][0], v[u][1]);
              upk2(fadd2(pk2(a4.z, a4.w), b23), v[u][2], v[u][3]);
            }
            if (ACT == CST_ACT_GLU) {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                v[u][0] = v[u][0] * mufu_rcp(1.0f + mufu_ex2(-1.4426950408889634f * v[u][1])) * p.alpha;
                v[u][1] = v[u][2] * mufu_rcp(1.0f + mufu_ex2(-1.4426950408889634f * v[u][3])) * p.alpha;
                if (p.residual) { v[u][0] += res[i0 + u].x; v[u][1] += res[i0 + u].y; }
              }
            } else {
              if (ACT == CST_ACT_GELU) {
#pragma unroll
                for (int u = 0; u < 4; ++u) { gelu2(v[u][0], v[u][1]); gelu2(v[u][2], v[u][3]); }
              } else if (ACT == CST_ACT_RELU) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  v[u][0] = fmaxf(v[u][0], 0.f); v[u][1] = fmaxf(v[u][1], 0.f);
                  v[u][2] = fmaxf(v[u][2], 0.f); v[u][3] = fmaxf(v[u][3], 0.f);
                }
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                uint64_t o01 = fmul2(pk2(v[u][0], v[u][1]), alpha2), o23 = fmul2(pk2(v[u][2], v[u][3]), alpha2);
                if (p.residual) {
                  o01 = fadd2(o01, pk2(res[i0 + u].x, res[i0 + u].y));
                  o23 = fadd2(o23, pk2(res[i0 + u].z, res[i0 + u].w));
                }
                upk2(o01, v[u][0], v[u][1]); upk2(o23, v[u][2], v[u][3]);
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int i = i0 + u;
              if (!((st_mask >> i) & 1)) continue;
              if ((z_mask >> i) & 1) { v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f; }
              if (ACT == CST_ACT_GLU) {
                if (c_16) *reinterpret_cast<uint32_t*>((__nv_bfloat16*)p.C + crow[i] + nc) = c_f16 ? pack_f16x2(v[u][0], v[u][1]) : pack_bf16x2(v[u][0], v[u][1]);
                else *reinterpret_cast<float2*>((float*)p.C + crow[i] + nc) = make_float2(v[u][0], v[u][1]);
              } else {
                if (c_f16) store4((__half*)p.C + crow[i] + nc, make_float4(v[u][0], v[u][1], v[u][2], v[u][3]));
                else if (c_16) store4((__nv_bfloat16*)p.C + crow[i] + nc, make_float4(v[u][0], v[u][1], v[u][2], v[u][3]));
                else store4((float*)p.C + crow[i] + nc, make_float4(v[u][0], v[u][1], v[u][2], v[u][3]));
              }
            }
          }
        }
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

template <int BN, int ACT>
static int launch_tc_act(const cst_gemm_params& hp, const GemmDev& p, int nz, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    CST_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
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
  CST_CHECK_CUDA(launch_k(gemm_tc_kernel<BN, ACT>, dim3(grid), dim3(TC_THREADS), Cfg::SMEM_BYTES, st, tmA, tmB, p, m_tiles, n_tiles, (int)total, a_wrap, hp.ab_dtype == CST_F16 ? 1 : 0));
  return CST_OK;
}

template <int BN>
static int launch_tc(const cst_gemm_params& hp, const GemmDev& p, int nz, cudaStream_t st) {
  switch (hp.act) {
    case CST_ACT_NONE: return launch_tc_act<BN, CST_ACT_NONE>(hp, p, nz, st);
    case CST_ACT_GELU: return launch_tc_act<BN, CST_ACT_GELU>(hp, p, nz, st);
    case CST_ACT_RELU: return launch_tc_act<BN, CST_ACT_RELU>(hp, p, nz, st);
    default: return launch_tc_act<BN, CST_ACT_GLU>(hp, p, nz, st);
  }
}

int launch_gemm_tc(const cst_gemm_params& hp, const GemmDev& p, int nz, cudaStream_t st) {
  // 128x256 tiles halve the B-operand traffic per flop.  Measured (c3, 3 stream lanes): switching problems with
  // < 1.5 tiles per SM to 128x128 tiles raises the isolated GEMM rate (548 -> 572 TFLOP/s) but LOWERS end-to-end
  // throughput (34.0k -> 33.0k audio-s/s): concurrent lanes already fill the tails, so 256 stays the default
  // (CST_TC_BN=128 forces the small tile for experiments).
  // Tiny problems (shared / memory layers, M <= ~1600 rows, < 0.5 tile per SM) run 2x faster in isolation with
  // 128x64 tiles (shorter serial K chain, 4x more CTAs), but with stream lanes they already overlap other batches'
  // kernels on the idle SMs and the end-to-end rate does not move (34.1k vs 33.8k audio-s/s): not enabled.
  static const int force_bn = [] { const char* e = getenv("CST_TC_BN"); return e ? atoi(e) : 0; }();
  if (hp.N % 256 == 0 && force_bn != 128 && force_bn != 64) return launch_tc<256>(hp, p, nz, st);
  if (force_bn == 64 && hp.N % 64 == 0) return launch_tc<64>(hp, p, nz, st);
  if (hp.N % 128 == 0) return launch_tc<128>(hp, p, nz, st);
  if (hp.N == 48) return launch_tc<48>(hp, p, nz, st);
  if (hp.N % 64 == 0) return launch_tc<64>(hp, p, nz, st);
  CST_REQUIRE(false, "cst_gemm(bf16): N=%d unsupported (multiple of 64, or 48)", hp.N);
  return CST_ERR_ARG;
}

}  // namespace cst
