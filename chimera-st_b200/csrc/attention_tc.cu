// Padding-aware flash attention on the 5th-gen tensor cores (bf16 operands, fp32 softmax/accumulate),
// head_dim 64.  One CTA = 128 query rows of one (utterance, head); keys/values stream in 128-row tiles.
// Resources are cut to HALF an SM (112 KB shared memory, 256 TMEM columns, <=170 registers) so TWO CTAs are
// resident per SM: while one CTA's softmax warps run (MUFU/FMA bound), the other CTA's MMAs use the tensor
// pipe -- the overlap a single CTA would need double-buffered S/P/PV for.
//   warp 0      : TMA producer (Q once; K_j,V_j into a 2-stage ring, 128B-swizzled)
//   warp 1      : tcgen05.mma issuer.  S_j = Q K_j^T (M128 N128 K64, both operands K-major) into TMEM buffer
//                 j&1 (two 128-column buffers); PV_j = P_j V_j (M128 N64 K128; P from shared memory K-major, V as
//                 an MN-major B operand straight from the TMA tile -- no transpose) into the first 64 columns of
//                 the SAME buffer, which is dead once the softmax has turned S_j into P_j.  S_{j+1} is issued as
//                 soon as PV_{j-1} has been consumed, i.e. in the middle of softmax j, so the next S is ready
//                 when the softmax warps come back (256 TMEM columns per CTA, still two CTAs per SM).
//   warps 2..9  : softmax.  Two warps per TMEM lane quarter: a query row is shared by a thread pair, each
//                 owning 64 of the tile's 128 keys and 32 of the 64 output columns (row max exchanged through
//                 shared memory + a 64-thread named barrier; partial row sums are added at the end).  tcgen05.ld
//                 the S half-row, exact key mask (keys >= kv_len[b] get zero probability), running max / sum in
//                 fp32, P = exp2 in bf16 -> shared memory (manual 128B swizzle); the output accumulator stays
//                 in registers: O = O * alpha_j + PV_j.
// Measured and rejected (round 2): delaying the second resident CTA of the first wave by 800 / 1600 / 2400 cycles to de-phase the two
// CTAs of an SM changes nothing (399 -> 401 TFLOP/s at B32 H12 T749, tools/attn_rate.py).
// Every query row of the tile is computed (padded query rows are live in the reference).  V rows of masked
// keys must be finite (the plan guarantees it: no activation row is ever left unwritten).
#include <cuda_fp16.h>
#include "tc_common.cuh"

namespace cst {

#ifndef FA_POLY_EXP
#define FA_POLY_EXP 0      // 1: half of the softmax exponentials on the FMA pipe. Measured: 1.84 -> 1.95 ms (c2), the pass is issue-bound, not MUFU-bound
#endif

constexpr int FA_BQ = 128, FA_BK = 128, FA_D = 64, FA_KS = 2;
constexpr int FA_SM_WARPS = 8;
constexpr int FA_THREADS = 64 + 32 * FA_SM_WARPS;
constexpr int FA_Q_BYTES = FA_BQ * FA_D * 2;            // 16 KB
constexpr int FA_KV_BYTES = FA_BK * FA_D * 2;           // 16 KB each for K and V
constexpr int FA_P_BYTES = FA_BQ * FA_BK * 2;           // 32 KB (two 64-key halves of 16 KB)
constexpr int FA_SMEM = FA_Q_BYTES + 2 * FA_KS * FA_KV_BYTES + FA_P_BYTES + 128 + 512;   // 112.6 KB: two CTAs per SM (limit 113 KB)
constexpr int FA_TMEM_COLS = 256;
constexpr float FA_LOG2E = 1.4426950408889634f;
#ifdef FA_PROFILE
__device__ long long fa_prof[16];
#define FA_T(i) do { if (prof) { const long long now_ = clock64(); atomicAdd((unsigned long long*)&fa_prof[i], (unsigned long long)(now_ - tprev)); tprev = now_; } } while (0)
#else
#define FA_T(i) do { } while (0)
#endif
struct TrueTag { static constexpr bool value = true; };
struct FalseTag { static constexpr bool value = false; };

__global__ void __launch_bounds__(FA_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, long long ldo,
                    int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg, const int32_t* __restrict__ kv_len,
                    const int4* __restrict__ seg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t sQ = base;
  const uint32_t sK = sQ + FA_Q_BYTES;
  const uint32_t sV = sK + FA_KS * FA_KV_BYTES;
  const uint32_t sP = sV + FA_KS * FA_KV_BYTES;
  const uint32_t bars = sP + FA_P_BYTES;
  const uint32_t q_full = bars;
  // K and V have separate rings: a K stage is free as soon as S_j has consumed it (long before PV_j frees
  // the V stage), so the producer can fetch K two tiles ahead although only two stages exist
  const uint32_t k_full = bars + 8, k_empty = k_full + 8 * FA_KS;
  const uint32_t v_full = k_empty + 8 * FA_KS, v_empty = v_full + 8 * FA_KS;
  const uint32_t s_full = v_empty + 8 * FA_KS;                     // [2]: one per S buffer
  const uint32_t p_full = s_full + 16, pv_full = p_full + 8, pv_empty = pv_full + 8;
  const uint32_t tmem_slot = pv_empty + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * FA_BQ;
  // ragged batches: seg[b] = {first q row, q rows, first kv row, kv rows} of utterance b (static plan data, not produced
  // by the preceding kernel); the grid covers the longest utterance and the surplus CTAs of shorter ones leave at once
  int q_row0 = b * q_rows_per_seg, kv_row0 = b * kv_rows_per_seg;
  if (seg != nullptr) {
    const int4 sg = seg[b];
    q_row0 = sg.x; n_q = sg.y; kv_row0 = sg.z; n_kv = sg.w;
  }
  if (q0 >= n_q) return;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    if ((base & 1023u) != 0) { printf("cst attention_tc: shared memory base not 1024-byte aligned\n"); __trap(); }
    mbar_init(q_full, 1);
    for (int s = 0; s < FA_KS; ++s) {
      mbar_init(k_full + 8 * s, 1); mbar_init(k_empty + 8 * s, 1);
      mbar_init(v_full + 8 * s, 1); mbar_init(v_empty + 8 * s, 1);
    }
    mbar_init(s_full, 1); mbar_init(s_full + 8, 1);
    mbar_init(p_full, FA_SM_WARPS); mbar_init(pv_full, 1); mbar_init(pv_empty, FA_SM_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)FA_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tS = tmem_base;                                // S_j (128 cols) then PV_j (64 cols) in buffer j&1 at col (j&1)*128
  pdl_wait();                                                   // prologue above overlaps the previous kernel
  int klen = kv_len ? kv_len[b] : n_kv;
  klen = klen < n_kv ? klen : n_kv;
  const int n_tiles = (klen + FA_BK - 1) / FA_BK;

  if (warp == 0) {
    if (lane == 0 && n_tiles > 0) {
      const int q_row = q_row0 + q0;
      mbar_expect_tx(q_full, FA_Q_BYTES);
      tma_load_2d(sQ, &tmQ, q_full, h * FA_D, q_row);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % FA_KS, u = j / FA_KS;
        const int k_row = kv_row0 + j * FA_BK;
        mbar_wait(k_empty + 8 * st, (u & 1) ^ 1);        // S_{j-2} done
        mbar_expect_tx(k_full + 8 * st, FA_KV_BYTES);
        tma_load_2d(sK + st * FA_KV_BYTES, &tmK, k_full + 8 * st, h * FA_D, k_row);
        mbar_wait(v_empty + 8 * st, (u & 1) ^ 1);        // PV_{j-2} done
        mbar_expect_tx(v_full + 8 * st, FA_KV_BYTES);
        tma_load_2d(sV + st * FA_KV_BYTES, &tmV, v_full + 8 * st, h * FA_D, k_row);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && n_tiles > 0) {
      // instruction descriptors: fp32 accumulate, bf16 x bf16; S: N=128 K-major B; PV: N=64 MN-major B
      constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FA_BK >> 3) << 17) | ((uint32_t)(FA_BQ >> 4) << 24);
      constexpr uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(FA_D >> 3) << 17) | ((uint32_t)(FA_BQ >> 4) << 24);
      const uint64_t qdesc = make_sw128_desc(sQ);
      auto issue_pv = [&](int j) {
        const int st = j % FA_KS;
        mbar_wait(v_full + 8 * st, (j / FA_KS) & 1);
        mbar_wait(p_full, j & 1);                        // P_j published => S_j fully read: its buffer may take PV_j
        tc_fence_after();
        int keys = klen - j * FA_BK; keys = keys < FA_BK ? keys : FA_BK;
        const int k16 = (keys + 15) >> 4;
        const uint32_t vbase = sV + st * FA_KV_BYTES;
        for (int k = 0; k < k16; ++k) {
          const uint64_t pdesc = make_sw128_desc(sP + (k >> 2) * (FA_P_BYTES / 2) + (k & 3) * 32);
          const uint64_t vdesc = make_sw128_mn_desc(vbase + k * 2048);
          tc_mma_bf16(tS + (j & 1) * FA_BK, pdesc, vdesc, idesc_pv, k != 0);
        }
        tc_commit(pv_full);
        tc_commit(v_empty + 8 * st);
      };
      auto issue_s = [&](int i) {
        const int st = i % FA_KS;
        mbar_wait(k_full + 8 * st, (i / FA_KS) & 1);
        if (i >= 2) mbar_wait(pv_empty, i & 1);          // PV_{i-2} (same buffer) consumed: completion index i-2
        tc_fence_after();
        const uint64_t kdesc = make_sw128_desc(sK + st * FA_KV_BYTES);
#pragma unroll
        for (int k = 0; k < FA_D / 16; ++k) tc_mma_bf16(tS + (i & 1) * FA_BK, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
        tc_commit(s_full + 8 * (i & 1));
        tc_commit(k_empty + 8 * st);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_s(j + 1);             // runs while the softmax warps are still on tile j
        issue_pv(j);
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax / output warps =====================
    const int qd = warp & 3;                             // TMEM lane quarter (hardware: warp % 4)
    const int hh = (warp - 2) >> 2;                      // 0: keys 0..63 / out cols 0..31, 1: keys 64..127 / cols 32..63
    const int r = qd * 32 + lane;                        // query row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    // row-max exchange between the two threads of a row: fp16 slots [2 halves][128 rows] (512 B is all the shared
    // memory left under the two-CTAs-per-SM budget).  Both threads use max(fp16(own), fp16(partner)) -- any common
    // reference point near the maximum is a valid softmax shift.  One slot set suffices: a second 64-thread barrier
    // after the read keeps a fast thread from rewriting its slot before the partner has read it.
    __half* xch = reinterpret_cast<__half*>(smem_raw + (bars + 128 - base));
    float* lsum = reinterpret_cast<float*>(smem_raw + (sK - base));         // reused after the last MMA: partial row sums
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
    const uint32_t srow0 = tS + lane_off + hh * 64;
    const uint32_t prow = sP + hh * (FA_P_BYTES / 2) + r * 128;
    const uint32_t pvrow0 = tS + lane_off + hh * 32;
    auto take_pv = [&](int j) {                          // O = O * alpha_j + PV_j  (PV_j is relative to m_j)
      mbar_wait(pv_full, j & 1);
      tc_fence_after();
      float pv[32];
      tmem_ld32(pvrow0 + (j & 1) * FA_BK, pv);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = fmaf(o[i], alpha_prev, pv[i]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pv_empty);
    };
    // one key tile; TAIL is a compile-time flag so that full tiles carry no per-element masking instructions
#ifdef FA_PROFILE
    const bool prof = (warp == 2 && lane == 0);
    long long tprev = clock64();
#endif
    auto tile_step = [&](int j, auto tail_tag) {
      constexpr bool TAIL = decltype(tail_tag)::value;
      const uint32_t srow = srow0 + (j & 1) * FA_BK;
      FA_T(0);                                           // waited for S_j
      const int nvalid = klen - (j * FA_BK + hh * 64);   // valid keys among this thread's 64 (TAIL only; may be <= 0)
      // pass 1: maximum of this thread's 64 scores (two 32-column TMEM loads; registers are capped at 96/thread
      // so that two 320-thread CTAs fit an SM, hence S is re-read in pass 2 instead of kept live)
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        float s[32];
        tmem_ld32(srow + c, s);
        tmem_ld_wait();
        if (TAIL) {
#pragma unroll
          for (int i = 0; i < 32; ++i) { if (c + i < nvalid) mx = fmaxf(mx, s[i]); }
        } else {
          // three-input max (FMNMX3, sm_100): half the instructions of the row-maximum pass
          float mx2 = mx;                                  // two independent chains
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            asm("max.f32 %0, %0, %1, %2;" : "+f"(mx) : "f"(s[i]), "f"(s[i + 1]));
            asm("max.f32 %0, %0, %1, %2;" : "+f"(mx2) : "f"(s[i + 2]), "f"(s[i + 3]));
          }
          mx = fmaxf(mx, mx2);
        }
      }
      FA_T(1);                                           // pass 1
      // row maximum across the thread pair
      const __half mxh = __float2half_rn(mx);
      xch[hh * 128 + r] = mxh;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
      mx = fmaxf(__half2float(mxh), __half2float(xch[(hh ^ 1) * 128 + r]));
      asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");   // partner has read: the slot may be rewritten next tile
      const float m_new = fmaxf(m_run, mx);              // finite: every visited tile holds >= 1 valid key
      const float alpha = mufu_ex2((m_run - m_new) * FA_LOG2E);
      const float mb = m_new * FA_LOG2E;
      FA_T(2);                                           // exchange
      if (j > 0) take_pv(j - 1);                         // also proves the P buffer is free again
      FA_T(3);                                           // PV_{j-1} wait + accumulate
      // pass 2: P = exp2(s*log2e - m*log2e) -> bf16 -> swizzled shared memory (packed FFMA2 + MUFU.EX2)
      const uint64_t l2e = pk2(FA_LOG2E, FA_LOG2E), nmb = pk2(-mb, -mb);
      uint64_t rs2 = pk2(0.f, 0.f);
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        float s[32];
        tmem_ld32(srow + c, s);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 32; g += 8) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            float a0, a1, p0, p1;
            upk2(ffma2(pk2(s[g + i], s[g + i + 1]), l2e, nmb), a0, a1);
            if (FA_POLY_EXP && (i & 2)) {
              // every other pair: 2^a on the FMA / integer pipes (Cody-Waite split + degree-3 minimax on [-0.5, 0.5], relative
              // error 7.7e-5, far below the bf16 rounding of P).  The 16 MUFU lanes per SM are what bounds this pass with
              // 16 softmax warps resident; half of the exponentials leave that pipe.
              const uint64_t ac = pk2(fmaxf(a0, -125.f), fmaxf(a1, -125.f));
              const uint64_t magic = pk2(12582912.f, 12582912.f);                  // 1.5 * 2^23: low mantissa bits = round(a)
              const uint64_t t = fadd2(ac, magic);
              const uint64_t f = ffma2(fadd2(t, pk2(-12582912.f, -12582912.f)), pk2(-1.f, -1.f), ac);   // a - round(a)
              uint64_t q = ffma2(f, pk2(5.508868381e-02f, 5.508868381e-02f), pk2(2.426040515e-01f, 2.426040515e-01f));
              q = ffma2(q, f, pk2(6.932762417e-01f, 6.932762417e-01f));
              q = ffma2(q, f, pk2(9.999289404e-01f, 9.999289404e-01f));
              float q0, q1, t0, t1;
              upk2(q, q0, q1); upk2(t, t0, t1);
              p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
              p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
            } else {
              p0 = mufu_ex2(a0); p1 = mufu_ex2(a1);
            }
            if (TAIL) {
              if (c + g + i >= nvalid) p0 = 0.f;
              if (c + g + i + 1 >= nvalid) p1 = 0.f;
            }
            rs2 = fadd2(rs2, pk2(p0, p1));
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
          const uint32_t addr = prow + (uint32_t)((((c + g) >> 3) ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
        }
      }
      float rs;
      { float r0, r1; upk2(rs2, r0, r1); rs = r0 + r1; }
      FA_T(4);                                           // pass 2
      tc_fence_before();                                 // S fully read
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // P visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      l_run = l_run * alpha + rs;
      m_run = m_new;
      alpha_prev = alpha;
      FA_T(5);                                           // fences + arrive
    };
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full + 8 * (j & 1), (j >> 1) & 1);
      tc_fence_after();
      if ((j + 1) * FA_BK > klen) tile_step(j, TrueTag{});
      else tile_step(j, FalseTag{});
    }
    if (n_tiles > 0) take_pv(n_tiles - 1);
    // combine the pair's partial row sums, normalise, store this thread's 32 output columns
    if (hh == 1) lsum[r] = l_run;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
    if (hh == 0) lsum[r] = l_run + lsum[r];
    asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
    const float l_tot = lsum[r];
    if (q0 + r < n_q) {
      const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;
      __nv_bfloat16* op = out + ((long long)q_row0 + q0 + r) * ldo + h * FA_D + hh * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 u;
        u.x = pack_bf16x2(o[i] * inv, o[i + 1] * inv); u.y = pack_bf16x2(o[i + 2] * inv, o[i + 3] * inv);
        u.z = pack_bf16x2(o[i + 4] * inv, o[i + 5] * inv); u.w = pack_bf16x2(o[i + 6] * inv, o[i + 7] * inv);
        *reinterpret_cast<uint4*>(op + i) = u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)FA_TMEM_COLS) : "memory");
  }
}

#ifdef FA_PROFILE
extern "C" int cst_debug_fa_prof(long long* host16, int reset) {
  if (reset) { long long z[16] = {0}; cudaMemcpyToSymbol(fa_prof, z, sizeof(z)); return 0; }
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host16, fa_prof, sizeof(long long) * 16);
}
#endif

// seg == nullptr: B uniform segments.  seg != nullptr: n_q / n_kv are the maxima over the table, q_rows_per_seg /
// kv_rows_per_seg the TOTAL rows of the q / kv buffers (tensor-map extents).
int launch_attention_tc(const void* q, const void* k, const void* v, void* out, long long ldq, long long ldkv, long long ldo,
                        int B, int H, int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg,
                        const int32_t* kv_len, const int32_t* seg, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    CST_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    attr_set = true;
  }
  CST_REQUIRE(((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0, "cst_attention(bf16): q/k/v must be 16-byte aligned");
  // tensor maps over the full row-major buffers: inner = one head's 64 columns are addressed by coordinate
  CUtensorMap tmQ, tmK, tmV;
  const long long q_total = seg ? q_rows_per_seg : (long long)B * q_rows_per_seg;
  const long long kv_total = seg ? kv_rows_per_seg : (long long)B * kv_rows_per_seg;
  int rc = make_map_2d(&tmQ, q, H * FA_D, q_total, ldq, FA_D, FA_BQ);
  if (rc) return rc;
  rc = make_map_2d(&tmK, k, H * FA_D, kv_total, ldkv, FA_D, FA_BK);
  if (rc) return rc;
  rc = make_map_2d(&tmV, v, H * FA_D, kv_total, ldkv, FA_D, FA_BK);
  if (rc) return rc;
  dim3 grid(cdiv(n_q, FA_BQ), H, B);
  CST_CHECK_CUDA(launch_k(attention_tc_kernel, grid, dim3(FA_THREADS), FA_SMEM, st, tmQ, tmK, tmV, (__nv_bfloat16*)out, ldo, n_q,
                          q_rows_per_seg, n_kv, kv_rows_per_seg, kv_len, reinterpret_cast<const int4*>(seg)));
  return CST_OK;
}

}  // namespace cst
