// Padding-aware flash attention on tcgen05, second generation: P lives in TENSOR MEMORY, persistent CTAs, one softmax
// thread per query row, two query tiles in flight per CTA (bf16 operands, fp32 softmax / accumulate, head_dim 64).
//
// Why (profiles/SUMMARY_r01.md §3c): in attention_tc.cu the softmax warps' serial chain per key tile (wait S -> row
// max -> pair exchange -> PV accumulate -> exp2 -> bf16 P -> shared memory -> fence) left the MUFU pipe -- the real
// bound at head_dim 64: 128x128 exponentials = 1024 MUFU cycles per tile against 512 MMA cycles -- idle half of the time,
// the PV product read P as a shared-memory A operand (A-read bound, 512 cycles where the N = 64 math needs 256), and a CTA
// owned a single 128-query tile, so ~15 % of its life was prologue.
//
//   grid = #SMs, one CTA per SM (512 TMEM columns, ~145 KB shared memory), 320 threads:
//   warp 0      : TMA producer.  Per work item (utterance b, head h, PAIR of 128-row query tiles): Q0, Q1 once, then
//                 K_j, V_j through separate 3-stage rings.  The next item's Q / K are fetched behind the current item's
//                 last key tile (q_empty is committed right after the item's last S MMAs).
//   warp 1      : tcgen05.mma issuer.  S_t = Q_t K_j^T (M128 N128 K64, SS) into TMEM columns [128t, 128t+128);
//                 PV_t = P_t V_j (M128 N64 K128, A operand = P_t FROM TENSOR MEMORY columns [256+64t, ..), V as an MN-major
//                 B operand straight from its TMA tile) into columns [384+64t, ..).  As soon as P_t,j is published the
//                 issuer sends S_t,j+1 FIRST (the S tile is free: P has its own columns) and PV_t,j after it, so a
//                 warpgroup waits one S MMA, not PV + S; the two query tiles run half a period apart, so one warpgroup's
//                 exponentials overlap the other's MMAs and row-max pass.
//   warps 2..5  : softmax warpgroup of query tile 0 (thread = row = TMEM lane), warps 6..9: query tile 1.
//                 pass 1: row maximum of S (4 x tcgen05.ld of 32 columns, the next chunk's load in flight, FMNMX3);
//                 O = O*alpha + PV_{j-1} (registers); pass 2: p = exp2(s*log2e - m*log2e) (FFMA2 + MUFU.EX2), row sum in
//                 fp32, bf16 pairs written to the P columns with tcgen05.st.
// Every query row is computed (padded query rows are live in the reference); keys >= kv_len[b] get exactly zero
// probability.  V rows of masked keys must be finite (plan.py: no activation row is ever left unwritten).
// Replaces: the attention core of F.multi_head_attention_forward as called from
// fairseq/modules/multihead_attention.py:155-187 (wav2vec2 layers wav2vec2.py:938-945, shared layers
// transformer_layer.py:131-137), same contract as attention_tc.cu.
#include <cuda_fp16.h>
#include "tc_common.cuh"

namespace cst {

constexpr int F2_BQ = 128, F2_BK = 128, F2_D = 64, F2_KS = 3;
constexpr int F2_THREADS = 64 + 2 * 128;
constexpr int F2_TILE_BYTES = 128 * F2_D * 2;                   // 16 KB: a Q, K or V tile
constexpr int F2_MAX_B = 4096;                                  // utterances per launch (work-list prefix in shared memory)
constexpr int F2_SMEM = (2 + 2 * F2_KS) * F2_TILE_BYTES + (F2_MAX_B + 32) * 4 + 256 + 1024;
constexpr uint32_t F2_TMEM_COLS = 512;
constexpr uint32_t F2_COL_S = 0, F2_COL_P = 256, F2_COL_PV = 384;
constexpr float F2_LOG2E = 1.4426950408889634f;

#ifdef F2_PROFILE
// phase clocks: [0..7] warp 2 (query tile 0), [8..15] warp 6 (query tile 1) of block 0; [16..23] the MMA issuer of block 0
__device__ long long f2_prof[32];
#define F2_T(i) do { if (prof) { const long long now_ = clock64(); pacc[i] += now_ - tprev; tprev = now_; } } while (0)
#define F2_FLUSH() do { if (prof) { for (int i_ = 0; i_ < 8; ++i_) f2_prof[pbase + i_] = pacc[i_]; } } while (0)
#else
#define F2_T(i) do { } while (0)
#define F2_FLUSH() do { } while (0)
#endif
struct F2True { static constexpr bool value = true; };
struct F2False { static constexpr bool value = false; };

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void tc_mma_ts_bf16(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

struct F2Item { int b, h, q0, two, q_row0, n_q, kv_row0, klen, n_tiles; };

__global__ void __launch_bounds__(F2_THREADS, 1)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, long long ldo,
                     int B, int H, int n_q_u, int q_rows_per_seg, int n_kv_u, int kv_rows_per_seg,
                     const int32_t* __restrict__ kv_len, const int4* __restrict__ seg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base;                                       // [2] tiles
  const uint32_t sK = sQ + 2 * F2_TILE_BYTES;                     // [KS]
  const uint32_t sV = sK + F2_KS * F2_TILE_BYTES;                 // [KS]
  const uint32_t cum_s = sV + F2_KS * F2_TILE_BYTES;              // int [B + 1]: query-tile PAIRS before utterance b
  const uint32_t bars = cum_s + (F2_MAX_B + 32) * 4;
  int* cum = reinterpret_cast<int*>(smem + (cum_s - base));
  const uint32_t q_full = bars, q_empty = bars + 8;
  const uint32_t k_full = bars + 16, k_empty = k_full + 8 * F2_KS;
  const uint32_t v_full = k_empty + 8 * F2_KS, v_empty = v_full + 8 * F2_KS;
  const uint32_t s_full = v_empty + 8 * F2_KS;                    // [2]
  const uint32_t p_full = s_full + 16, pv_full = p_full + 16;     // [2] each
  const uint32_t turn = pv_full + 16;                             // [2]: MUFU-phase ping-pong between the two warpgroups
  const uint32_t tmem_slot = turn + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int s = 0; s < F2_KS; ++s) {
      mbar_init(k_full + 8 * s, 1); mbar_init(k_empty + 8 * s, 1);
      mbar_init(v_full + 8 * s, 1); mbar_init(v_empty + 8 * s, 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(s_full + 8 * t, 1); mbar_init(p_full + 8 * t, 4); mbar_init(pv_full + 8 * t, 1); mbar_init(turn + 8 * t, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(F2_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- work list: pairs of 128-row query tiles per utterance -> exclusive prefix in shared memory (seg is static plan
  // data, not produced by the preceding kernel)
  for (int b = threadIdx.x; b < B; b += F2_THREADS) {
    const int nq = seg ? seg[b].y : n_q_u;
    cum[b + 1] = (nq + 2 * F2_BQ - 1) / (2 * F2_BQ);
  }
  if (threadIdx.x == 0) cum[0] = 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) {                                               // inclusive scan of cum[1..B] by one warp
    int carry = 0;
    for (int i0 = 1; i0 <= B; i0 += 32) {
      const int i = i0 + lane;
      int v = (i <= B) ? cum[i] : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += u; }
      if (i <= B) cum[i] = v + carry;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
  }
  __syncthreads();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int total_items = cum[B] * H;
  pdl_wait();                                                    // prologue above overlaps the previous kernel

  auto decode = [&](int item, F2Item& w) {
    const int pg = item / H;
    w.h = item - pg * H;
    int lo = 0, hi = B;                                          // largest b with cum[b] <= pg
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (cum[mid] <= pg) lo = mid; else hi = mid; }
    w.b = lo;
    w.q0 = (pg - cum[lo]) * 2 * F2_BQ;
    int n_kv;
    if (seg) { const int4 sg = seg[lo]; w.q_row0 = sg.x; w.n_q = sg.y; w.kv_row0 = sg.z; n_kv = sg.w; }
    else { w.q_row0 = lo * q_rows_per_seg; w.n_q = n_q_u; w.kv_row0 = lo * kv_rows_per_seg; n_kv = n_kv_u; }
    int klen = kv_len ? kv_len[lo] : n_kv;
    w.klen = klen < n_kv ? klen : n_kv;
    w.n_tiles = (w.klen + F2_BK - 1) / F2_BK;
    w.two = (w.q0 + F2_BQ < w.n_q) ? 1 : 0;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int kv_i = 0, it = 0;                                      // running K/V tile index (ring position), items processed
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        F2Item w; decode(item, w);
        if (w.n_tiles == 0) continue;
        mbar_wait(q_empty, (it & 1) ^ 1);                        // previous item's S MMAs are done with the Q tiles
        mbar_expect_tx(q_full, (1 + w.two) * F2_TILE_BYTES);
        tma_load_2d(sQ, &tmQ, q_full, w.h * F2_D, w.q_row0 + w.q0);
        if (w.two) tma_load_2d(sQ + F2_TILE_BYTES, &tmQ, q_full, w.h * F2_D, w.q_row0 + w.q0 + F2_BQ);
        for (int j = 0; j < w.n_tiles; ++j, ++kv_i) {
          const int st = kv_i % F2_KS; const uint32_t ph = (kv_i / F2_KS) & 1;
          const int k_row = w.kv_row0 + j * F2_BK;
          mbar_wait(k_empty + 8 * st, ph ^ 1);
          mbar_expect_tx(k_full + 8 * st, F2_TILE_BYTES);
          tma_load_2d(sK + st * F2_TILE_BYTES, &tmK, k_full + 8 * st, w.h * F2_D, k_row);
          mbar_wait(v_empty + 8 * st, ph ^ 1);
          mbar_expect_tx(v_full + 8 * st, F2_TILE_BYTES);
          tma_load_2d(sV + st * F2_TILE_BYTES, &tmV, v_full + 8 * st, w.h * F2_D, k_row);
        }
        ++it;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptors: fp32 accumulate, bf16 x bf16; S: N=128, K-major B; PV: N=64, MN-major B, A from TMEM
      constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(F2_BK >> 3) << 17) | ((uint32_t)(F2_BQ >> 4) << 24);
      constexpr uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(F2_D >> 3) << 17) | ((uint32_t)(F2_BQ >> 4) << 24);
      int kv_i = 0, it = 0;
      uint32_t n_p[2] = {0, 0};                                  // P tiles consumed per query tile (p_full phase)
#ifdef F2_PROFILE
      const bool prof = blockIdx.x == 0;
      const int pbase = 16;
      long long tprev = clock64();
      long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        F2Item w; decode(item, w);
        if (w.n_tiles == 0) continue;
        const int nt = w.n_tiles;
        auto issue_s = [&](int t, int kvi) {                     // S_t = Q_t K^T for ring position kvi
          const int st = kvi % F2_KS;
          const uint64_t qdesc = make_sw128_desc(sQ + t * F2_TILE_BYTES);
          const uint64_t kdesc = make_sw128_desc(sK + st * F2_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < F2_D / 16; ++k)
            tc_mma_bf16(tmem_base + F2_COL_S + t * F2_BK, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
          tc_commit(s_full + 8 * t);
        };
        auto issue_pv = [&](int t, int j, int kvi) {             // PV_t = P_t V_j
          const int st = kvi % F2_KS;
          int keys = w.klen - j * F2_BK; keys = keys < F2_BK ? keys : F2_BK;
          const int k16 = (keys + 15) >> 4;
          const uint32_t vbase = sV + st * F2_TILE_BYTES;
          for (int k = 0; k < k16; ++k)
            tc_mma_ts_bf16(tmem_base + F2_COL_PV + t * F2_D, tmem_base + F2_COL_P + t * (F2_BK / 2) + k * 8,
                           make_sw128_mn_desc(vbase + k * 2048), idesc_pv, k != 0);
          tc_commit(pv_full + 8 * t);
        };
        const int nq_tiles = 1 + w.two;
        mbar_wait(q_full, it & 1);
        mbar_wait(k_full + 8 * (kv_i % F2_KS), (kv_i / F2_KS) & 1);
        tc_fence_after();
        for (int t = 0; t < nq_tiles; ++t) issue_s(t, kv_i);
        tc_commit(k_empty + 8 * (kv_i % F2_KS));
        if (nt == 1) tc_commit(q_empty);
        for (int j = 0; j < nt; ++j, ++kv_i) {
          const bool more = j + 1 < nt;
          const int st = kv_i % F2_KS, stn = (kv_i + 1) % F2_KS;
          for (int t = 0; t < nq_tiles; ++t) {
            // P_t,j published: S_t,j is fully read (its tile may take S_t,j+1) and PV_t,j-1 has been consumed
            F2_T(2 + t);                                         // issue work since the last wait
            mbar_wait(p_full + 8 * t, n_p[t] & 1); ++n_p[t];
            F2_T(t);                                             // waited for P_t
            if (more) {
              if (t == 0) mbar_wait(k_full + 8 * stn, ((kv_i + 1) / F2_KS) & 1);
              tc_fence_after();
              issue_s(t, kv_i + 1);                              // first: the warpgroup's next tile
              if (t + 1 == nq_tiles) {
                tc_commit(k_empty + 8 * stn);
                if (j + 2 == nt) tc_commit(q_empty);             // last S MMAs of the item: the Q tiles may be replaced
              }
            }
            if (t == 0) mbar_wait(v_full + 8 * st, (kv_i / F2_KS) & 1);
            tc_fence_after();
            issue_pv(t, j, kv_i);
          }
          tc_commit(v_empty + 8 * st);
        }
        ++it;
      }
      F2_FLUSH();
    }
    __syncwarp();
  } else {
    // ===================== softmax / output warpgroups =====================
    const int t = (warp - 2) >> 2;                               // query tile of the pair
    const int qd = warp & 3;                                     // TMEM lane quarter (hardware: warp % 4)
    const int r = qd * 32 + lane;                                // query row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + F2_COL_S + t * F2_BK;
    const uint32_t tP = tmem_base + lane_off + F2_COL_P + t * (F2_BK / 2);
    const uint32_t tPV = tmem_base + lane_off + F2_COL_PV + t * F2_D;
    uint32_t n_s = 0, n_pv = 0;                                  // tiles seen (barrier phases)
    uint32_t n_turn = 0;                                         // exponential passes taken under the ping-pong token
#ifdef F2_PROFILE
    const bool prof = blockIdx.x == 0 && lane == 0 && (warp == 2 || warp == 6);
    const int pbase = warp == 2 ? 0 : 8;
    long long tprev = clock64();
    long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      F2Item w; decode(item, w);
      if (t == 1 && !w.two) continue;
      float o[F2_D];
#pragma unroll
      for (int i = 0; i < F2_D; ++i) o[i] = 0.f;
      float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
      auto take_pv = [&]() {                                     // O = O * alpha + PV  (PV is relative to the tile's maximum)
        mbar_wait(pv_full + 8 * t, n_pv & 1); ++n_pv;
        tc_fence_after();
        float pv[F2_D];
        tmem_ld32(tPV, pv);                                      // both halves in flight, one wait
        tmem_ld32(tPV + 32, pv + 32);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < F2_D; ++i) o[i] = fmaf(o[i], alpha_prev, pv[i]);
      };
      auto tile_step = [&](int j, auto tail_tag) {
        constexpr bool TAIL = decltype(tail_tag)::value;
        const int nvalid = w.klen - j * F2_BK;                   // valid keys of this tile (TAIL only)
        // pass 1: row maximum; chunk c+1's tcgen05.ld is in flight while chunk c is reduced
        float mx = -INFINITY, mx2 = -INFINITY;
        float sb[2][32];
        tmem_ld32(tS, sb[0]);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = ci * 32;
          float (&s)[32] = sb[ci & 1];
          tmem_ld_wait();
          if (ci < 3) tmem_ld32(tS + c + 32, sb[(ci + 1) & 1]);
          if (TAIL) {
#pragma unroll
            for (int i = 0; i < 32; ++i) { if (c + i < nvalid) mx = fmaxf(mx, s[i]); }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              asm("max.f32 %0, %0, %1, %2;" : "+f"(mx) : "f"(s[i]), "f"(s[i + 1]));
              asm("max.f32 %0, %0, %1, %2;" : "+f"(mx2) : "f"(s[i + 2]), "f"(s[i + 3]));
            }
          }
        }
        mx = fmaxf(mx, mx2);
        F2_T(1);                                                 // pass 1
        const float m_new = fmaxf(m_run, mx);                    // finite: every visited tile holds >= 1 valid key
        const float alpha = mufu_ex2((m_run - m_new) * F2_LOG2E);
        const float mb = m_new * F2_LOG2E;
        if (j > 0) take_pv();                                    // PV_{j-1}: also proves the PV buffer is free for PV_j
        F2_T(2);                                                 // PV wait + accumulate
        // The exponential pass is what saturates the MUFU pipe (1024 cycles per tile against 512 MMA cycles).  Left alone the
        // two warpgroups fall into lock-step (both publish P at the same time, both then wait for their MMAs: measured 4000
        // cycles per key tile); a ping-pong token makes them alternate, so one warpgroup's exponentials run while the other
        // takes its row maximum, its PV tile and its S-MMA latency.  Items with a single query tile run without the token.
        if (w.two) mbar_wait(turn + 8 * t, (n_turn & 1) ^ (t == 0 ? 1u : 0u));
        F2_T(3);                                                 // token wait
        // pass 2: P = exp2(s*log2e - m*log2e) -> bf16 pairs -> the P columns of tensor memory
        const uint64_t l2e = pk2(F2_LOG2E, F2_LOG2E), nmb = pk2(-mb, -mb);
        uint64_t rs2 = pk2(0.f, 0.f);
        // (16-column chunks, double-buffered: the next chunk's tcgen05.ld is in flight during this chunk's exponentials)
        tmem_ld16(tS, sb[0]);
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
          const int c = ci * 16;
          float (&s)[32] = sb[ci & 1];
          tmem_ld_wait();
          if (ci < 7) tmem_ld16(tS + c + 16, sb[(ci + 1) & 1]);
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float a0, a1;
            upk2(ffma2(pk2(s[i], s[i + 1]), l2e, nmb), a0, a1);
            float p0 = mufu_ex2(a0), p1 = mufu_ex2(a1);
            if (TAIL) {
              if (c + i >= nvalid) p0 = 0.f;
              if (c + i + 1 >= nvalid) p1 = 0.f;
            }
            rs2 = fadd2(rs2, pk2(p0, p1));
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
          tmem_st8(tP + (c >> 1), pk);                           // keys c..c+15 -> 8 packed columns
        }
        F2_T(4);                                                 // pass 2
        tmem_st_wait();
        tc_fence_before();                                       // S fully read, P written
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(p_full + 8 * t);
          if (w.two) mbar_arrive(turn + 8 * (t ^ 1));            // the other warpgroup's turn
        }
        if (w.two) ++n_turn;
        F2_T(5);                                                 // st wait + fences + arrives
        float r0, r1; upk2(rs2, r0, r1);
        l_run = l_run * alpha + (r0 + r1);
        m_run = m_new;
        alpha_prev = alpha;
      };
      for (int j = 0; j < w.n_tiles; ++j) {
        F2_T(6);                                                 // between tiles (item decode, epilogue of the previous item)
        mbar_wait(s_full + 8 * t, n_s & 1); ++n_s;
        tc_fence_after();
        F2_T(0);                                                 // waited for S
        if ((j + 1) * F2_BK > w.klen) tile_step(j, F2True{});
        else tile_step(j, F2False{});
      }
      if (w.n_tiles > 0) take_pv();
      const int row = w.q0 + t * F2_BQ + r;
      if (row < w.n_q) {
        const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
        __nv_bfloat16* op = out + ((long long)w.q_row0 + row) * ldo + w.h * F2_D;
#pragma unroll
        for (int i = 0; i < F2_D; i += 8) {
          uint4 u;
          u.x = pack_bf16x2(o[i] * inv, o[i + 1] * inv); u.y = pack_bf16x2(o[i + 2] * inv, o[i + 3] * inv);
          u.z = pack_bf16x2(o[i + 4] * inv, o[i + 5] * inv); u.w = pack_bf16x2(o[i + 6] * inv, o[i + 7] * inv);
          *reinterpret_cast<uint4*>(op + i) = u;
        }
      }
    }
    F2_FLUSH();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(F2_TMEM_COLS) : "memory");
  }
}

#ifdef F2_PROFILE
}  // namespace cst
extern "C" int cst_debug_f2_prof(long long* host32, int reset) {
  if (reset) { long long z[32] = {0}; cudaMemcpyToSymbol(cst::f2_prof, z, sizeof(z)); return 0; }
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host32, cst::f2_prof, sizeof(long long) * 32);
}
namespace cst {
#endif

// Same contract as launch_attention_tc (attention_tc.cu).  Returns CST_ERR_UNSUPPORTED when the launch does not fit this
// kernel (more than F2_MAX_B utterances) so that the caller keeps the first-generation kernel.
int launch_attention_tc2(const void* q, const void* k, const void* v, void* out, long long ldq, long long ldkv, long long ldo,
                         int B, int H, int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg,
                         const int32_t* kv_len, const int32_t* seg, cudaStream_t st) {
  if (B > F2_MAX_B) return CST_ERR_UNSUPPORTED;
  static bool attr_set = false;
  static int sms = 0;
  if (!attr_set) {
    CST_CHECK_CUDA(cudaFuncSetAttribute(attention_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM));
    int dev = 0;
    CST_CHECK_CUDA(cudaGetDevice(&dev));
    CST_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    attr_set = true;
  }
  CST_REQUIRE(((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0 && ((uintptr_t)out % 16) == 0,
              "cst_attention(bf16): q/k/v/out must be 16-byte aligned");
  const long long q_total = seg ? q_rows_per_seg : (long long)B * q_rows_per_seg;
  const long long kv_total = seg ? kv_rows_per_seg : (long long)B * kv_rows_per_seg;
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_map_2d(&tmQ, q, H * F2_D, q_total, ldq, F2_D, F2_BQ);
  if (rc) return rc;
  rc = make_map_2d(&tmK, k, H * F2_D, kv_total, ldkv, F2_D, F2_BK);
  if (rc) return rc;
  rc = make_map_2d(&tmV, v, H * F2_D, kv_total, ldkv, F2_D, F2_BK);
  if (rc) return rc;
  // upper bound of the work items (exact for uniform segments): never launch more CTAs than there is work
  const long long items_ub = (long long)B * H * ((n_q + 2 * F2_BQ - 1) / (2 * F2_BQ));
  const int grid = (int)(items_ub < sms ? items_ub : sms);
  CST_CHECK_CUDA(launch_k(attention_tc2_kernel, dim3(grid), dim3(F2_THREADS), F2_SMEM, st, tmQ, tmK, tmV, (__nv_bfloat16*)out, ldo,
                          B, H, n_q, q_rows_per_seg, n_kv, kv_rows_per_seg, kv_len, reinterpret_cast<const int4*>(seg)));
  return CST_OK;
}

}  // namespace cst
