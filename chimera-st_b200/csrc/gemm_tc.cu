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
#include <cstring>
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

#ifdef TC_PROFILE
// event trace of pair 0 (globaltimer ns): [0] leader producer passes the empty wait, [1] peer producer passes it,
// [2] MMA thread sees the stage full, [3] MMA thread has issued the stage's MMAs + commit
__device__ unsigned long long tc_trace[4][1024];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// epilogue phases of warp 2 of block 0 (cycles): [0] row bookkeeping, [1] wait for the accumulator, [2] column-chunk loop, [3] tiles
__device__ long long tc_epi_prof[8];
__device__ int tc_dbg_flags;      // timing experiments: 1 skip global stores, 2 skip the read-back/math, 4 skip the patch writes
#define TC_DBG(bit) (tc_dbg_flags & (bit))
#define TC_EPI_T(var) const long long var = clock64()
#define TC_EPI_ADD(slot, a, b) do { if (blockIdx.x == 0 && threadIdx.x == 64) tc_epi_prof[slot] += (b) - (a); } while (0)
#define TC_TRACE(row, idx) do { if (pair == 0 && (idx) < 1024) tc_trace[row][idx] = gtime(); } while (0)
#else
#define TC_TRACE(row, idx) do { } while (0)
#define TC_EPI_T(var) do { } while (0)
#define TC_DBG(bit) false
#define TC_EPI_ADD(slot, a, b) do { } while (0)
#endif

// ---- lean tile epilogue: compile-time output dtype (0 f32, 1 bf16, 2 f16) and residual flag, alpha == 1, no
// per-segment zeroing.  Measured (tools/dbg_tc_epi.py): for K = 768 the epilogue, not the MMA stream, paces the tile
// (8.7k cycles against a 6.1k MMA floor) and it is latency-bound -- each warp walks its four 32x32 chunks serially
// through TMEM load -> smem transpose -> math -> store.  So everything with a long latency is issued one chunk ahead:
// the tcgen05.ld of chunk c+1 goes out as soon as chunk c's registers are parked in the patch, and the bias / residual
// loads of chunk c+1 are in flight while chunk c's math runs.
// BST (fp32 output, identity row mapping): the finished values go back into the patch -- whose XOR layout IS the
// 128-byte TMA swizzle -- and one lane bulk-stores the 32x32 block, instead of eight STG per lane and chunk.
// LNF (fp32 output with residual): LayerNorm fused around the GEMM (cst_gemm_params: res_stats / C2 / out_stats) -- the
// residual rows are normalised on the fly from their partial statistics, a bf16 copy of the output is stored next to the
// fp32 one, and this warp's partial statistics of the output rows (its 128-column slice) are written for the consumers.
// LNS: the fused-LayerNorm epilogue also emits output statistics / the bf16 copy (full fusion); false = only the lazily normalised
// residual of the default "light LayerNorm" mode, which then needs neither the 16 statistics registers nor their arithmetic.
template <int BN, int ACT, int CD, bool RES, bool BST, bool LNF, int VAR, bool LNS, typename WaitF>
__device__ __forceinline__ void epi_tile_fast(const GemmDev& p, float* patch, uint32_t t_row, int nb, int lane, int chalf,
                                              const float* bias, uint8_t* c_base, const float* r_base, const int (&orow)[8],
                                              uint32_t st_mask, WaitF wait_acc, const CUtensorMap* tmC = nullptr, int row0 = 0) {
  static_assert(!BST || CD == 0, "in-place bulk store: fp32 output only");
  static_assert(!LNF || (RES && CD == 0), "fused LayerNorm epilogue: fp32 output with residual");
  constexpr int NCH = (BN + 31) / 32, CH_PER = (NCH + 1) / 2;
  constexpr int ES = CD == 0 ? 4 : 2;
  const int lr = lane >> 3, lc = (lane & 7) * 4;
  const int ch0 = chalf * CH_PER, ch1 = (ch0 + CH_PER < NCH) ? ch0 + CH_PER : NCH;
  auto col_of = [&](int ch) { return nb * BN + ch * 32 + lc; };
  auto col_ok_of = [&](int ch) { return (ch * 32 + lc < BN) && (col_of(ch) < p.N); };
  auto load_bias = [&](int ch) {
    return (bias && col_ok_of(ch)) ? __ldg(reinterpret_cast<const float4*>(bias + col_of(ch))) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  // residual rows i0 .. i0+3 of chunk `ch` (C and residual rows are addressed as base + row * ld: 32-bit row indices keep
  // the per-lane row state at 8 registers instead of 32 for two pointer arrays)
  auto load_res = [&](int ch, int i0, float4 (&r)[RES ? 8 : 1]) {
    if (RES) {
      const bool ok = col_ok_of(ch);
      const int n = col_of(ch);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u;
        r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok && ((st_mask >> i) & 1)) {
          // 256-byte L2 fill granularity: the neighbouring 128 B of the row is this warp's next chunk
          const float* ra = r_base + (long long)orow[i] * p.ldr + n;
          asm volatile("ld.global.L2::256B.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r[i].x), "=f"(r[i].y), "=f"(r[i].z), "=f"(r[i].w) : "l"(ra));
        }
      }
    }
  };
  auto load_acc = [&](int ch, float (&a)[32]) {
    const int c = ch * 32;
    if (BN - c >= 32) {
      tmem_ld32(t_row + c, a);
    } else {                                                  // BN = 48: the last chunk holds 16 columns
      tmem_ld16(t_row + c, a);
#pragma unroll
      for (int i = 16; i < 32; ++i) a[i] = 0.f;
    }
  };
  float acc[32];
  float4 b_cur = load_bias(ch0), b_nxt = b_cur;
  float4 r_cur[RES ? 8 : 1];
  load_res(ch0, 0, r_cur);
  load_res(ch0, 4, r_cur);
  // fused LayerNorm state: per-row {rstd, -mean*rstd} of the lazily normalised residual, gamma / beta of the current and
  // next chunk, running partial statistics of the output rows
  const bool lazy_res = LNF && p.res_stats != nullptr;
  float ra[LNF ? 8 : 1], rb[LNF ? 8 : 1], s1[LNS ? 8 : 1], s2[LNS ? 8 : 1];
  float4 g_cur = make_float4(1.f, 1.f, 1.f, 1.f), t_cur = make_float4(0.f, 0.f, 0.f, 0.f), g_nxt = g_cur, t_nxt = t_cur;
  auto load_gamma = [&](int ch) {
    return (lazy_res && col_ok_of(ch)) ? __ldg(reinterpret_cast<const float4*>(p.res_gamma + col_of(ch))) : make_float4(1.f, 1.f, 1.f, 1.f);
  };
  auto load_beta = [&](int ch) {
    return (lazy_res && col_ok_of(ch)) ? __ldg(reinterpret_cast<const float4*>(p.res_beta + col_of(ch))) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  if (LNF) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      ra[i] = 1.f; rb[i] = 0.f;
      if (LNS) { s1[i] = 0.f; s2[i] = 0.f; }
      if (lazy_res && ((st_mask >> i) & 1)) {
        const float2 ab = ln_ab_from_partials(p.res_stats, orow[i], p.res_slots, p.ln_inv_dim);
        ra[i] = ab.x; rb[i] = ab.y;
      }
    }
    g_cur = load_gamma(ch0); t_cur = load_beta(ch0);
  }
  wait_acc();
  load_acc(ch0, acc);
#pragma unroll 1
  for (int ch = ch0; ch < ch1; ++ch) {
    const int n = col_of(ch);
    const bool col_ok = col_ok_of(ch);
    TC_EPI_T(q0);
    tmem_ld_wait();
    TC_EPI_T(q1);
    if (BST) { if (lane == 0) bulk_wait_read<0>(); }           // the previous chunk's bulk store has read the patch
    __syncwarp();                                              // previous chunk's readers are done with the patch
    TC_EPI_T(q2);
    if (!TC_DBG(4)) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<float4*>(&patch[lane * TC_PATCH_LD + 4 * (i ^ (lane & 7))]) = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
    }
    __syncwarp();
    TC_EPI_T(q3);
    const bool has_next = ch + 1 < ch1;
    if (has_next) {                                            // next chunk: TMEM and bias loads in flight during the math
      load_acc(ch + 1, acc);
      b_nxt = load_bias(ch + 1);
      if (LNF) { g_nxt = load_gamma(ch + 1); t_nxt = load_beta(ch + 1); }
    }
    if (col_ok && !TC_DBG(2)) {
      const uint64_t b01 = pk2(b_cur.x, b_cur.y), b23 = pk2(b_cur.z, b_cur.w);
      const uint64_t sc2 = VAR == 2 ? pk2(p.acc_scale, p.acc_scale) : 0ull;
#pragma unroll
      for (int i0 = 0; i0 < 8; i0 += 4) {
        float v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int rr = (i0 + u) * 4 + lr;
          const float4 a4 = *reinterpret_cast<const float4*>(&patch[rr * TC_PATCH_LD + 4 * ((lc >> 2) ^ (rr & 7))]);
          if constexpr (VAR == 2) {
            upk2(ffma2(pk2(a4.x, a4.y), sc2, b01), v[u][0], v[u][1]);
            upk2(ffma2(pk2(a4.z, a4.w), sc2, b23), v[u][2], v[u][3]);
          } else {
            upk2(fadd2(pk2(a4.x, a4.y), b01), v[u][0], v[u][1]);
            upk2(fadd2(pk2(a4.z, a4.w), b23), v[u][2], v[u][3]);
          }
        }
        if (ACT == CST_ACT_GELU) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (VAR == 2 && p.exact_act) { v[u][0] = gelu_erf(v[u][0]); v[u][1] = gelu_erf(v[u][1]); v[u][2] = gelu_erf(v[u][2]); v[u][3] = gelu_erf(v[u][3]); }
            else { gelu2(v[u][0], v[u][1]); gelu2(v[u][2], v[u][3]); }
          }
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
            float4 r4 = r_cur[i0 + u];
            if (LNF) {                                         // LayerNorm of the residual row, applied on the fly
              const float a = ra[i0 + u], b = rb[i0 + u];
              r4.x = fmaf(fmaf(r4.x, a, b), g_cur.x, t_cur.x); r4.y = fmaf(fmaf(r4.y, a, b), g_cur.y, t_cur.y);
              r4.z = fmaf(fmaf(r4.z, a, b), g_cur.z, t_cur.z); r4.w = fmaf(fmaf(r4.w, a, b), g_cur.w, t_cur.w);
            }
            upk2(fadd2(pk2(v[u][0], v[u][1]), pk2(r4.x, r4.y)), v[u][0], v[u][1]);
            upk2(fadd2(pk2(v[u][2], v[u][3]), pk2(r4.z, r4.w)), v[u][2], v[u][3]);
          }
          if (has_next) load_res(ch + 1, i0, r_cur);           // these four registers are free again: refill for chunk c+1
        }
        if (LNS) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u;
            if ((st_mask >> i) & 1) {
              s1[i] += (v[u][0] + v[u][1]) + (v[u][2] + v[u][3]);
              s2[i] += fmaf(v[u][0], v[u][0], v[u][1] * v[u][1]) + fmaf(v[u][2], v[u][2], v[u][3] * v[u][3]);
              if (p.C2 != nullptr)
                *reinterpret_cast<uint2*>(p.C2 + (long long)orow[i] * p.ldc2 + n) =
                    make_uint2(pack_bf16x2(v[u][0], v[u][1]), pack_bf16x2(v[u][2], v[u][3]));
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u;
          if (BST) {
            const int rr = i * 4 + lr;
            *reinterpret_cast<float4*>(&patch[rr * TC_PATCH_LD + 4 * ((lc >> 2) ^ (rr & 7))]) = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
          } else if (((st_mask >> i) & 1) && !TC_DBG(1)) {
            uint8_t* o = c_base + ((long long)orow[i] * p.ldc + n) * ES;
            if (CD == 0) *reinterpret_cast<float4*>(o) = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
            else if (CD == 1) *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16x2(v[u][0], v[u][1]), pack_bf16x2(v[u][2], v[u][3]));
            else *reinterpret_cast<uint2*>(o) = make_uint2(pack_f16x2(v[u][0], v[u][1]), pack_f16x2(v[u][2], v[u][3]));
          }
        }
      }
    }
    TC_EPI_T(q4);
    TC_EPI_ADD(4, q0, q1); TC_EPI_ADD(5, q1, q2); TC_EPI_ADD(6, q2, q3); TC_EPI_ADD(7, q3, q4);
    if (BST) {                                                 // (all columns of a BN % 64 == 0 tile are valid)
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmC, smem_u32(patch), nb * BN + ch * 32, row0);
        bulk_commit();
      }
    }
    b_cur = b_nxt;
    if (LNF) { g_cur = g_nxt; t_cur = t_nxt; }
  }
  if (LNS) {
    if (p.out_stats != nullptr) {                              // the 8 lanes of a row group hold 4 columns each of the row's chunks
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) {
          s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], off);
          s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], off);
        }
        if ((lane & 7) == 0 && ((st_mask >> i) & 1))
          p.out_stats[(long long)orow[i] * 8 + nb * 2 + chalf] = make_float2(s1[i], s2[i]);
      }
    }
  }
}

// ---- bulk-store tile epilogue: identity row mapping, no residual / GLU / alpha / per-segment zeroing ----------------
// Measured (tools/gemm_rate.py with CST_TC_DBG): in the transpose-and-STG epilogue above the global stores alone
// cost 10-20 % of the GEMM (QKV 1037 -> 1256 TFLOP/s without them) and the smem round trip another few percent.
// Here the math stays in the accumulator's own layout (lane = row, 32 consecutive columns in registers), each lane
// parks its converted row in a swizzled staging buffer and ONE lane hands the 32x32 block to the TMA engine, which
// writes it (clipping rows >= M) while the warp moves on: no STG, no read-back, no per-row address arithmetic.
// Two register buffers alternate so that the tcgen05.ld of chunk c+1 is in flight during chunk c's math, and the
// accumulator stage is released as soon as its last chunk is in registers.
// LNI: the A operand was the bf16 copy of UN-normalised rows; their LayerNorm is applied after the product (lane = row, so
// mean / rstd are per-lane scalars): v = rstd*acc + (-mean*rstd)*colsum[n] + bias[n]  (cst_gemm_params: ln_in_stats).
template <int BN, int ACT, int CD, bool LNI, int VAR, typename WaitF, typename ReleaseF>
__device__ __forceinline__ void epi_tile_bulk(const GemmDev& p, const CUtensorMap* tmC, uint32_t stage_s, uint32_t& seq, uint32_t t_row,
                                              int row0, int nb, int lane, int chalf, const float* bias, WaitF wait_acc, ReleaseF release_acc) {
  static_assert(BN % 64 == 0, "bulk epilogue: whole 32-column chunks per warp half");
  constexpr int CH_PER = BN / 64;                              // 32-column chunks per warp
  constexpr int NBUF = CD == 0 ? 1 : 2;                        // 4 KB per warp: one fp32 block or two 16-bit blocks
  const int ch0 = chalf * CH_PER;
  float acc[2][32];
  uint64_t ln_a2 = 0, ln_b2 = 0;
  if (LNI) {
    float2 ab = make_float2(0.f, 0.f);
    if (row0 + lane < p.M) ab = ln_ab_from_partials(p.ln_in_stats, row0 + lane, p.ln_in_slots, p.ln_inv_dim);
    ln_a2 = pk2(ab.x, ab.x); ln_b2 = pk2(ab.y, ab.y);
  }
  wait_acc();
  tmem_ld32(t_row + ch0 * 32, acc[0]);
#pragma unroll
  for (int j = 0; j < CH_PER; ++j) {
    float (&v)[32] = acc[j & 1];
    const int n0 = nb * BN + (ch0 + j) * 32;
    tmem_ld_wait();
    if (j + 1 < CH_PER) tmem_ld32(t_row + (ch0 + j + 1) * 32, acc[(j + 1) & 1]);
    else release_acc();                                        // whole accumulator stage is in registers now
    if (LNI) {                                                 // rstd*acc + (-mean*rstd)*colsum + bias
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias + n0) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 cs = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + n0) + i);
        upk2(ffma2(pk2(v[4 * i], v[4 * i + 1]), ln_a2, ffma2(ln_b2, pk2(cs.x, cs.y), pk2(bv.x, bv.y))), v[4 * i], v[4 * i + 1]);
        upk2(ffma2(pk2(v[4 * i + 2], v[4 * i + 3]), ln_a2, ffma2(ln_b2, pk2(cs.z, cs.w), pk2(bv.z, bv.w))), v[4 * i + 2], v[4 * i + 3]);
      }
    } else if (bias) {                                         // same address in every lane: one broadcast transaction each
      const uint64_t sc2 = VAR == 2 ? pk2(p.acc_scale, p.acc_scale) : 0ull;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + n0) + i);
        if constexpr (VAR == 2) {
          upk2(ffma2(pk2(v[4 * i], v[4 * i + 1]), sc2, pk2(bv.x, bv.y)), v[4 * i], v[4 * i + 1]);
          upk2(ffma2(pk2(v[4 * i + 2], v[4 * i + 3]), sc2, pk2(bv.z, bv.w)), v[4 * i + 2], v[4 * i + 3]);
        } else {
          upk2(fadd2(pk2(v[4 * i], v[4 * i + 1]), pk2(bv.x, bv.y)), v[4 * i], v[4 * i + 1]);
          upk2(fadd2(pk2(v[4 * i + 2], v[4 * i + 3]), pk2(bv.z, bv.w)), v[4 * i + 2], v[4 * i + 3]);
        }
      }
    } else if (VAR == 2 && p.acc_scale != 1.0f) {
      const uint64_t sc2 = pk2(p.acc_scale, p.acc_scale);
#pragma unroll
      for (int i = 0; i < 32; i += 2) upk2(fmul2(pk2(v[i], v[i + 1]), sc2), v[i], v[i + 1]);
    }
    if (ACT == CST_ACT_GELU) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        if (VAR == 2 && p.exact_act) { v[i] = gelu_erf(v[i]); v[i + 1] = gelu_erf(v[i + 1]); }      // fp32 parity mode (split GEMMs)
        else gelu2(v[i], v[i + 1]);
      }
    } else if (ACT == CST_ACT_RELU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    // staging buffer (seq % NBUF): its previous bulk store must have finished READING shared memory
    const uint32_t buf = stage_s + (NBUF == 2 ? (seq & 1u) * 2048u : 0u);
    if (lane == 0) bulk_wait_read<NBUF - 1>();
    __syncwarp();
    if (CD == 0) {                                             // 128-byte rows, SWIZZLE_128B: 16-byte chunk index ^ (row & 7)
#pragma unroll
      for (int i = 0; i < 8; ++i)
        sts128(buf + lane * 128 + ((i ^ (lane & 7)) << 4), __float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]),
               __float_as_uint(v[4 * i + 2]), __float_as_uint(v[4 * i + 3]));
    } else {                                                   // 64-byte rows, SWIZZLE_64B: chunk index ^ ((row >> 1) & 3)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          w[k] = CD == 1 ? pack_bf16x2(v[8 * i + 2 * k], v[8 * i + 2 * k + 1]) : pack_f16x2(v[8 * i + 2 * k], v[8 * i + 2 * k + 1]);
        sts128(buf + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4), w[0], w[1], w[2], w[3]);
      }
    }
    fence_async_smem();                                        // generic-proxy writes -> visible to the bulk copy
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tmC, buf, n0, row0);
      bulk_commit();
    }
    ++seq;
  }
}

// ---- one accumulator tile: TMEM -> registers -> per-warp smem transpose -> bias / activation / alpha / residual ->
// global.  `row_base` is the GEMM row of TMEM lane 0 of this warp's lane quarter; `wait_acc` blocks until the
// accumulator stage is complete (called after the row bookkeeping so that work overlaps the wait).
// VAR: the rarely used epilogue variants (LayerNorm fused around the GEMM, exact activations and the accumulator scale of the split
// fp32 mode) live in their own kernel instantiations -- compiled into the default kernels they cost registers (272 bytes of spills
// in the BN = 256 / 128 no-activation kernels) and ~15 % of the c3 GEMM time (measured: 113 -> 133 ms per step).
template <int BN, int ACT, int VAR, typename WaitF, typename ReleaseF>
__device__ __forceinline__ void epi_tile(const GemmDev& p, const CUtensorMap* tmC, int bulk, uint32_t& seq, float* patch, uint32_t t_row,
                                         int row_base, int nb, int zo, int zi, int lane, int chalf, WaitF wait_acc, ReleaseF release_acc) {
  if constexpr (BN % 64 == 0 && ACT != CST_ACT_GLU) {
    if (bulk == 1) {
      const uint32_t stage_s = smem_u32(patch);
      if constexpr (VAR == 2) {
        if (p.ln_in_stats != nullptr) {                          // LayerNorm of the input applied after the product (bf16 outputs)
          if constexpr (ACT == CST_ACT_NONE || ACT == CST_ACT_GELU || ACT == CST_ACT_RELU)
            epi_tile_bulk<BN, ACT, 1, true, true>(p, tmC, stage_s, seq, t_row, row_base, nb, lane, chalf, p.bias, wait_acc, release_acc);
          return;
        }
      }
      if (p.c_dtype == CST_F32) epi_tile_bulk<BN, ACT, 0, false, VAR>(p, tmC, stage_s, seq, t_row, row_base, nb, lane, chalf, p.bias, wait_acc, release_acc);
      else if (p.c_dtype == CST_BF16) epi_tile_bulk<BN, ACT, 1, false, VAR>(p, tmC, stage_s, seq, t_row, row_base, nb, lane, chalf, p.bias, wait_acc, release_acc);
      else epi_tile_bulk<BN, ACT, 2, false, VAR>(p, tmC, stage_s, seq, t_row, row_base, nb, lane, chalf, p.bias, wait_acc, release_acc);
      return;
    }
  }
  constexpr int NCH = (BN + 31) / 32;                        // 32-column chunks in the tile
  constexpr int CH_PER = (NCH + 1) / 2;
  TC_EPI_T(pt0);
  const int lr = lane >> 3, lc = (lane & 7) * 4;             // row-contiguous domain: 4 rows x 8 float4 per pass
  const bool c_16 = p.c_dtype != CST_F32, c_f16 = p.c_dtype == CST_F16;
  const uint64_t alpha2 = pk2(p.alpha, p.alpha);
  const bool fast = (ACT != CST_ACT_GLU) && p.alpha == 1.0f && p.seg_len == nullptr;
  const int esz = c_16 ? 2 : 4;
  const float* bias = p.bias ? p.bias + (long long)zi * p.bias_bs_inner : nullptr;
  const long long c_off = zo * p.c_bs_outer + zi * p.c_bs_inner;
  const long long r_off = zo * p.r_bs_outer + zi * p.r_bs_inner;
  // this lane's 8 rows are the same for every column chunk of the tile: resolve them once
  int orow[8];
  uint32_t st_mask = 0, z_mask = 0;
  {
    // rows row_base + lr + 4 i: one division, then step (the per-row form cost ~1800 cycles per tile, measured)
    int m = row_base + lr;
    int seg = m / p.rows_per_seg, t = m - seg * p.rows_per_seg;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool store = (m < p.M) && (t < p.seg_rows_valid);
      bool zero = false;
      if (p.seg_len != nullptr && store) zero = t >= p.seg_len[zo * p.segs_per_outer + seg];
      orow[i] = (int)((long long)seg * p.out_rows_per_seg + t + p.out_row_off);
      st_mask |= (store ? 1u : 0u) << i;
      z_mask |= (zero ? 1u : 0u) << i;
      m += 4; t += 4;
      while (t >= p.rows_per_seg) { t -= p.rows_per_seg; ++seg; }
    }
  }
  TC_EPI_T(pt1);
  TC_EPI_T(pt2);
  if (fast) {
    uint8_t* c_base = reinterpret_cast<uint8_t*>(p.C) + c_off * esz;
    const float* r_base = p.residual + r_off;
    const bool lnf = VAR != 0 && (p.res_stats != nullptr || p.C2 != nullptr || p.out_stats != nullptr);
    if (p.residual) {
      if (!c_16) {
        if constexpr (BN % 64 == 0) {
          if (lnf) {
            if constexpr (VAR == 2 && ACT == CST_ACT_NONE) {   // out-proj / fc2 of the full fusion: statistics + bf16 copy as well
              if (bulk == 2) epi_tile_fast<BN, ACT, 0, true, true, true, 2, true>(p, patch, t_row, nb, lane, chalf, bias, c_base, r_base, orow, st_mask, wait_acc, tmC, row_base);
              else epi_tile_fast<BN, ACT, 0, true, false, true, 2, true>(p, patch, t_row, nb, lane, chalf, bias, c_base, r_base, orow, st_mask, wait_acc);
            } else if constexpr (VAR == 1 && ACT == CST_ACT_NONE) {   // light LayerNorm mode: lazily normalised residual only
              if (bulk == 2) epi_tile_fast<BN, ACT, 0, true, true, true, 1, false>(p, patch, t_row, nb, lane, chalf, bias, c_base, r_base, orow, st_mask, wait_acc, tmC, row_base);
              else epi_tile_fast<BN, ACT, 0, true, false, true, 1, false>(p, patch, t_row, nb, lane, chalf, bias, c_base, r_base, orow, st_mask, wait_acc);
            }
          }
          else if (bulk == 2) epi_tile_fast<BN, ACT, 0, true, true, false, VAR, false>(p, patch, t_row, nb, lane, chalf, bias, c_base, r_base, orow, st_mask, wait_acc, tmC, row_base);
          else epi_tile_fast<BN, ACT, 0, true, false, false, VAR, false>(p, patch, t_row, nb, lane, chalf, bias, c_base, r_base, orow, st_mask, wait_acc);
        } else {
          epi_tile_fast<BN, ACT, 0, true, false, false, VAR, false>(p, patch, t_row, nb, lane, chalf, bias, c_base, r_base, orow, st_mask, wait_acc);
        }
      }
      else if (!c_f16) epi_tile_fast<BN, ACT, 1, true, false, false, VAR, false>(p, patch, t_row, nb, lane, chalf, bias, c_base, r_base, orow, st_mask, wait_acc);
      else epi_tile_fast<BN, ACT, 2, true, false, false, VAR, false>(p, patch, t_row, nb, lane, chalf, bias, c_base, r_base, orow, st_mask, wait_acc);
    } else {
      if (!c_16) epi_tile_fast<BN, ACT, 0, false, false, false, VAR, false>(p, patch, t_row, nb, lane, chalf, bias, c_base, r_base, orow, st_mask, wait_acc);
      else if (!c_f16) epi_tile_fast<BN, ACT, 1, false, false, false, VAR, false>(p, patch, t_row, nb, lane, chalf, bias, c_base, r_base, orow, st_mask, wait_acc);
      else epi_tile_fast<BN, ACT, 2, false, false, false, VAR, false>(p, patch, t_row, nb, lane, chalf, bias, c_base, r_base, orow, st_mask, wait_acc);
    }
  } else {
  wait_acc();
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
          if (ACT == CST_ACT_GLU) { const float2 t2 = *reinterpret_cast<const float2*>(p.residual + r_off + (long long)orow[i] * p.ldr + nc); res[i].x = t2.x; res[i].y = t2.y; }
          else res[i] = *reinterpret_cast<const float4*>(p.residual + r_off + (long long)orow[i] * p.ldr + nc);
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
      const uint64_t sc2 = VAR == 2 ? pk2(p.acc_scale, p.acc_scale) : pk2(1.0f, 1.0f);
      // 4 rows per batch: the 16 element chains (bias, activation, alpha, residual) are independent, so the
      // two epilogue warps of an SM sub-partition keep the FMA/MUFU pipes busy instead of waiting on one chain
#pragma unroll
      for (int i0 = 0; i0 < 8; i0 += 4) {
        float v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int rr = (i0 + u) * 4 + lr;
          const float4 a4 = *reinterpret_cast<const float4*>(&patch[rr * TC_PATCH_LD + 4 * ((lc >> 2) ^ (rr & 7))]);
          upk2(ffma2(pk2(a4.x, a4.y), sc2, b01), v[u][0], v[u][1]);
          upk2(ffma2(pk2(a4.z, a4.w), sc2, b23), v[u][2], v[u][3]);
        }
        if (ACT == CST_ACT_GLU) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (VAR == 2 && p.exact_act) {
              v[u][0] = v[u][0] * sigmoidf_(v[u][1]) * p.alpha;
              v[u][1] = v[u][2] * sigmoidf_(v[u][3]) * p.alpha;
            } else {
              v[u][0] = v[u][0] * mufu_rcp(1.0f + mufu_ex2(-1.4426950408889634f * v[u][1])) * p.alpha;
              v[u][1] = v[u][2] * mufu_rcp(1.0f + mufu_ex2(-1.4426950408889634f * v[u][3])) * p.alpha;
            }
            if (p.residual) { v[u][0] += res[i0 + u].x; v[u][1] += res[i0 + u].y; }
          }
        } else {
          if (ACT == CST_ACT_GELU) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (VAR == 2 && p.exact_act) { v[u][0] = gelu_erf(v[u][0]); v[u][1] = gelu_erf(v[u][1]); v[u][2] = gelu_erf(v[u][2]); v[u][3] = gelu_erf(v[u][3]); }
              else { gelu2(v[u][0], v[u][1]); gelu2(v[u][2], v[u][3]); }
            }
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
            if (c_16) *reinterpret_cast<uint32_t*>((__nv_bfloat16*)p.C + c_off + (long long)orow[i] * p.ldc + nc) = c_f16 ? pack_f16x2(v[u][0], v[u][1]) : pack_bf16x2(v[u][0], v[u][1]);
            else *reinterpret_cast<float2*>((float*)p.C + c_off + (long long)orow[i] * p.ldc + nc) = make_float2(v[u][0], v[u][1]);
          } else {
            if (c_f16) store4((__half*)p.C + c_off + (long long)orow[i] * p.ldc + nc, make_float4(v[u][0], v[u][1], v[u][2], v[u][3]));
            else if (c_16) store4((__nv_bfloat16*)p.C + c_off + (long long)orow[i] * p.ldc + nc, make_float4(v[u][0], v[u][1], v[u][2], v[u][3]));
            else store4((float*)p.C + c_off + (long long)orow[i] * p.ldc + nc, make_float4(v[u][0], v[u][1], v[u][2], v[u][3]));
          }
        }
      }
    }
  }
  }
  release_acc();
  TC_EPI_T(pt3);
  TC_EPI_ADD(0, pt0, pt1); TC_EPI_ADD(1, pt1, pt2); TC_EPI_ADD(2, pt2, pt3); TC_EPI_ADD(3, 0, 1);
}

template <int BN, int ACT, int VAR>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
               const GemmDev p, int m_tiles, int n_tiles, int total_tiles, int a_wrap, int ab_f16, int bulk) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + STAGES * Cfg::A_BYTES;
  const uint32_t epi_s = base + STAGES * Cfg::STAGE_BYTES;       // per-warp 4 KB epilogue patches (1024-aligned: swizzled staging)
  const uint32_t bars = epi_s + Cfg::EPI_BYTES;                  // 8-byte mbarriers
  const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES;
  const uint32_t tfull_bar = bars + 16 * STAGES, tempty_bar = tfull_bar + 16;
  const uint32_t tmem_slot = tempty_bar + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  uint8_t* epi_base = smem_raw + (epi_s - smem_u32(smem_raw));

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
    // Measured and rejected (round 2): letting this warp pull each tile's fp32 residual rows into L2 a k-loop ahead
    // (cp.async.bulk.prefetch.L2, 128 x 1 KB per tile) for the HBM-bound residual GEMMs (out-proj: 372 TFLOP/s, ~0.55 of the HBM
    // bound) made c3 SLOWER: 191.8 - 192.4 -> 196.7 - 196.8 ms per step; the same prefetch issued by the
    // epilogue warps for their slice of the NEXT tile (one lane per row, 512 B) was also slower (187.9 - 189.2 -> 189.4 - 189.9 ms):
    // those rows were written by the kernel before and mostly sit in the 126 MB L2 already.
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
    int it = 0; uint32_t seq = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int nb = tile % n_tiles; const int r = tile / n_tiles;
      const int mb = r % m_tiles; const int z = r / m_tiles;
      const int zo = z / p.nb_inner, zi = z - zo * p.nb_inner;
      const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * Cfg::BN_PAD;
      epi_tile<BN, ACT, VAR>(p, &tmC, bulk, seq, patch, t_row, mb * TC_BM + q * 32, nb, zo, zi, lane, chalf,
                        [&] { mbar_wait(tfull_bar + 8 * as, aph); tc_fence_after(); },
                        [&] { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(tempty_bar + 8 * as); });
    }
    if (bulk && lane == 0) bulk_wait_all<0>();     // staging buffers are read, and the stores complete, before exit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
  }
}

// The bulk-store epilogue applies when output row == GEMM row (no segment remap / padding rows / zeroing), there is no
// residual, GLU or alpha, and the problem is not batched; C rows beyond M are clipped by the tensor map.
// CST_TC_BULK=0 keeps the register-path epilogue (A/B experiments).
// Returns 0 (register-path stores), 1 (bulk epilogue) or 2 (fp32 + residual: register path, in-place staging, bulk store;
// the residual must not alias a different row of C, which the identity mapping guarantees).  CST_TC_BULK: bit 0 enables
// mode 1, bit 1 mode 2 (default 3).
static int bulk_store_ok(const cst_gemm_params& hp, int nz, int bn) {
  static const int enabled = [] { const char* e = getenv("CST_TC_BULK"); return e ? atoi(e) : 3; }();
  const int esz = hp.c_dtype == CST_F32 ? 4 : 2;
  const bool ident = nz == 1 && bn % 64 == 0 && hp.act != CST_ACT_GLU && hp.alpha == 1.0f &&
         hp.seg_len == nullptr && hp.out_row_off == 0 && hp.out_rows_per_seg == hp.rows_per_seg &&
         hp.seg_rows_valid >= hp.rows_per_seg && hp.N % bn == 0 && ((uintptr_t)hp.C % 16) == 0 && (hp.ldc * esz) % 16 == 0;
  if (!ident) return 0;
  if (hp.residual == nullptr) return (enabled & 1) ? 1 : 0;
  return ((enabled & 2) && hp.c_dtype == CST_F32) ? 2 : 0;
}
// fused LayerNorm modes exist only where the plan uses them: identity-mapped, unbatched GEMMs with 256-column tiles
static int check_ln_fused(const cst_gemm_params& hp, int nz, int bn, int bulk) {
  if (hp.ln_in_stats) {
    CST_REQUIRE(bulk == 1 && hp.c_dtype == CST_BF16 && hp.act != CST_ACT_GLU, "cst_gemm: ln_in_stats needs the bulk-store epilogue with a bf16 output (identity "
                "row mapping, N %% %d == 0, no residual / GLU / alpha)", bn);
    CST_REQUIRE(((uintptr_t)hp.ln_colsum % 16) == 0 && ((uintptr_t)hp.ln_in_stats % 8) == 0, "cst_gemm: ln_colsum / ln_in_stats alignment");
  }
  if (hp.res_stats || hp.C2 || hp.out_stats) {
    const bool ident = nz == 1 && bn % 64 == 0 && hp.act == CST_ACT_NONE && hp.alpha == 1.0f && hp.seg_len == nullptr && hp.out_row_off == 0 &&
                       hp.out_rows_per_seg == hp.rows_per_seg && hp.seg_rows_valid >= hp.rows_per_seg && hp.N % bn == 0;
    CST_REQUIRE(ident && hp.N / bn * 2 <= 8, "cst_gemm: res_stats / C2 / out_stats need an identity-mapped, unbatched GEMM without "
                "activation (N %% %d == 0, at most 8 statistics slots)", bn);
  }
  return CST_OK;
}
static int make_c_map(CUtensorMap* tmC, const cst_gemm_params& hp, int bulk) {
  if (!bulk) { memset(tmC, 0, sizeof(*tmC)); return CST_OK; }
  const int esz = hp.c_dtype == CST_F32 ? 4 : 2;
  return make_map_2d(tmC, hp.C, hp.N, hp.M, hp.ldc, 32, 32, esz, esz == 4 ? 128 : 64);
}

// 0: default kernels; 1: only the lazily normalised residual of the light LayerNorm mode (out-proj / fc2); 2: every variant
static inline int needs_variant_kernel(const cst_gemm_params& hp, const GemmDev& p) {
  if (hp.ln_in_stats || hp.C2 || hp.out_stats || hp.exact_act || p.acc_scale != 1.0f) return 2;
  if (hp.res_stats) return hp.act == CST_ACT_NONE ? 1 : 2;
  return 0;
}

template <int BN, int ACT, int VAR>
static int launch_tc_act(const cst_gemm_params& hp, const GemmDev& p, int nz, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    CST_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, ACT, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
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
  const int bulk = bulk_store_ok(hp, nz, BN);
  rc = check_ln_fused(hp, nz, BN, bulk);
  if (rc) return rc;
  CUtensorMap tmC;
  rc = make_c_map(&tmC, hp, bulk);
  if (rc) return rc;
  const int m_tiles = cdiv(hp.M, TC_BM), n_tiles = cdiv(hp.N, BN);
  const long long total = (long long)m_tiles * n_tiles * nz;
  CST_REQUIRE(total < (1ll << 31), "cst_gemm(bf16): too many tiles");
  int dev = 0, sms = 0;
  CST_CHECK_CUDA(cudaGetDevice(&dev));
  CST_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = (int)(total < sms ? total : sms);
  CST_CHECK_CUDA(launch_k(gemm_tc_kernel<BN, ACT, VAR>, dim3(grid), dim3(TC_THREADS), Cfg::SMEM_BYTES, st, tmA, tmB, tmC, p, m_tiles, n_tiles, (int)total, a_wrap, hp.ab_dtype == CST_F16 ? 1 : 0, bulk));
  return CST_OK;
}

// ---- CTA-pair variant: 256x256 output tile per cluster of two CTAs (tcgen05 cta_group::2, M = 256) ------------------
// Measured on B200 (profiles/SUMMARY_r01.md): the single-CTA kernel's tile period follows the L2 -> SM fill rate
// (~55 B/clk/SM delivered against 96 B/clk needed by a 128x256x64 k-block every 512 MMA cycles).  In a pair each
// CTA stages its own 128 A rows and HALF of the 256 B rows (32 KB per k-block instead of 48 KB) and the leader CTA
// issues one M=256 instruction that reads both shared memories and writes both tensor memories.
//   both CTAs : warp 0 = TMA producer (bytes complete on the LEADER's full barrier), warps 2-9 = epilogue of the
//               CTA's own 128 accumulator rows (arrive on the LEADER's tmem-empty barrier)
//   leader    : warp 1 lane 0 = MMA issuer; tcgen05.commit multicasts to the empty / tmem-full barriers of both CTAs
#ifdef TC_PAIR_SUSPEND
#define PAIR_WAIT mbar_wait
#else
#define PAIR_WAIT mbar_wait_poll
#endif
constexpr int TC2_STAGES = 6;
constexpr int TC2_HALF_BYTES = 128 * TC_BK * 2;                 // 16 KB: 128 A rows or 128 W rows of one k-block
constexpr int TC2_STAGE_BYTES = 2 * TC2_HALF_BYTES;
constexpr int TC2_SMEM_BYTES = TC2_STAGES * TC2_STAGE_BYTES + TC_EPI_WARPS * TC_PATCH_BYTES + 1024 + 256;

template <int ACT, int VAR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                    const GemmDev p, int m_tiles, int n_tiles, int total_tiles, int a_wrap, int ab_f16, int bulk) {
  constexpr int BN = 256, STAGES = TC2_STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + STAGES * TC2_HALF_BYTES;
  const uint32_t epi_s = base + STAGES * TC2_STAGE_BYTES;
  const uint32_t bars = epi_s + TC_EPI_WARPS * TC_PATCH_BYTES;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES;
  const uint32_t tfull_bar = bars + 16 * STAGES, tempty_bar = tfull_bar + 16;
  const uint32_t tmem_slot = tempty_bar + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  uint8_t* epi_base = smem_raw + (epi_s - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                      // 0 = leader
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int nkb = p.K / TC_BK;
  const uint32_t lead_bars = mapa_u32(bars, 0);                 // the leader's barrier block in cluster address space
  const uint32_t lead_full = lead_bars, lead_tempty = lead_bars + 16 * STAGES + 16;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + 8 * s, 1); mbar_init(tempty_bar + 8 * s, 2 * TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  cluster_sync_all();                  // both CTAs resident, barriers initialised before any remote arrive
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0; int ev = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs) {
        const int nb = tile % n_tiles; const int r = tile / n_tiles;
        const int mb = r % m_tiles; const int z = r / m_tiles;
        const int zo = z / p.nb_inner, zi = z - zo * p.nb_inner;
        const long long a_row0 = (zo * p.a_bs_outer + zi * p.a_bs_inner) / p.lda + (long long)mb * 256 + rank * 128;
        const int b_row0 = zi * p.N + nb * BN + (int)rank * 128;
        for (int kb = 0; kb < nkb; ++kb) {
          PAIR_WAIT(empty_bar + 8 * s, ph ^ 1);
          TC_TRACE(rank, ev); ++ev;
          // Both CTAs' bytes complete on the leader's barrier; only the leader arrives (count 1).  The peer's bytes may
          // land before the leader's expect_tx -- the transaction count simply goes negative until then.  (A remote
          // release.cluster arrive from the peer, as a count-2 protocol would need, costs ~500 ns per k-block: traced.)
#ifdef TC_PROFILE
          if (a_wrap & 12) {           // timing experiment: skip the A (4) and/or B (8) loads
            const uint32_t nb_loads = ((a_wrap & 4) ? 0 : 1) + ((a_wrap & 8) ? 0 : 1);
            if (rank == 0) mbar_expect_tx(full_bar + 8 * s, 2 * nb_loads * TC2_HALF_BYTES);
            if (!(a_wrap & 4)) tma_load_2d_pair(sA + s * TC2_HALF_BYTES, &tmA, lead_full + 8 * s, 0, (int)rank * 128);
            if (!(a_wrap & 8)) tma_load_2d_pair(sB + s * TC2_HALF_BYTES, &tmB, lead_full + 8 * s, 0, (int)rank * 128);
            if (++s == STAGES) { s = 0; ph ^= 1; }
            continue;
          }
#endif
          if (rank == 0) mbar_expect_tx(full_bar + 8 * s, 2 * TC2_STAGE_BYTES);
          const int kk = kb * TC_BK;
          int acol = kk, arow_add = 0;
          if (a_wrap & 1) { arow_add = kk / (int)p.lda; acol = kk - arow_add * (int)p.lda; }
#ifdef TC_PROFILE
          if (a_wrap & 2) {            // timing experiment: every load hits the same few lines (wrong results)
            tma_load_2d_pair(sA + s * TC2_HALF_BYTES, &tmA, lead_full + 8 * s, 0, (int)rank * 128);
            tma_load_2d_pair(sB + s * TC2_HALF_BYTES, &tmB, lead_full + 8 * s, 0, (int)rank * 128);
          } else
#endif
          {
          tma_load_2d_pair(sA + s * TC2_HALF_BYTES, &tmA, lead_full + 8 * s, acol, (int)(a_row0 + arow_add));
          tma_load_2d_pair(sB + s * TC2_HALF_BYTES, &tmB, lead_full + 8 * s, kk, b_row0);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      const uint32_t fmt = ab_f16 ? 0u : 1u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int s = 0; uint32_t ph = 0; int it = 0; int ev = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
        const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
        PAIR_WAIT(tempty_bar + 8 * as, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int kb = 0; kb < nkb; ++kb) {
          PAIR_WAIT(full_bar + 8 * s, ph);
          TC_TRACE(2, ev);
          tc_fence_after();
          const uint64_t adesc = make_sw128_desc(sA + s * TC2_HALF_BYTES);
          const uint64_t bdesc = make_sw128_desc(sB + s * TC2_HALF_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            tc_mma_pair_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          tc_commit_pair(empty_bar + 8 * s, 3);
          TC_TRACE(3, ev); ++ev;
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        tc_commit_pair(tfull_bar + 8 * as, 3);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..9 of both CTAs; this CTA's 128 rows) =====================
    const int ew = warp - 2;
    const int q = warp & 3;
    const int chalf = ew >> 2;
    float* patch = reinterpret_cast<float*>(epi_base + ew * TC_PATCH_BYTES);
    int it = 0; uint32_t seq = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
      const int nb = tile % n_tiles; const int r = tile / n_tiles;
      const int mb = r % m_tiles; const int z = r / m_tiles;
      const int zo = z / p.nb_inner, zi = z - zo * p.nb_inner;
      const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256;
      epi_tile<256, ACT, VAR>(p, &tmC, bulk, seq, patch, t_row, mb * 256 + (int)rank * 128 + q * 32, nb, zo, zi, lane, chalf,
                        [&] { PAIR_WAIT(tfull_bar + 8 * as, aph); tc_fence_after(); },
                        [&] { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive_cluster(lead_tempty + 8 * as); });
    }
    if (bulk && lane == 0) bulk_wait_all<0>();
  }

  tc_fence_before();
  cluster_sync_all();                  // the leader's MMAs read the peer's shared memory and write its tensor memory
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

#ifdef TC_PROFILE
}  // namespace cst
extern "C" int cst_debug_tc_flags(int flags) { return (int)cudaMemcpyToSymbol(cst::tc_dbg_flags, &flags, sizeof(int)); }
extern "C" int cst_debug_tc_epi(long long* host8, int reset) {
  if (reset) { long long z[8] = {0}; cudaMemcpyToSymbol(cst::tc_epi_prof, z, sizeof(z)); return 0; }
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host8, cst::tc_epi_prof, sizeof(long long) * 8);
}
extern "C" int cst_debug_tc_trace(unsigned long long* host) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host, cst::tc_trace, sizeof(unsigned long long) * 4 * 1024);
}
namespace cst {
#endif

template <int ACT, int VAR>
static int launch_tc_pair_act(const cst_gemm_params& hp, const GemmDev& p, int nz, cudaStream_t st) {
  static bool attr_set = false;
  static int max_pairs = 0;
  if (!attr_set) {
    CST_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<ACT, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_BYTES));
    int dev = 0, sms = 0;
    CST_CHECK_CUDA(cudaGetDevice(&dev));
    CST_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms & ~1); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = TC2_SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_tc_pair_kernel<ACT, VAR>, &cfg) != cudaSuccess || n <= 0) { (void)cudaGetLastError(); n = sms / 2; }
    max_pairs = n < sms / 2 ? n : sms / 2;
    if (const char* e = getenv("CST_TC_PAIR_GRID")) max_pairs = atoi(e) > 0 ? atoi(e) : max_pairs;
    if (getenv("CST_DEBUG")) fprintf(stderr, "cst: gemm_tc_pair occupancy query = %d clusters, using %d (sms=%d)\n", n, max_pairs, sms);
    attr_set = true;
  }
  int a_wrap = hp.K > hp.lda ? 1 : 0;
  if (a_wrap) CST_REQUIRE(hp.lda % TC_BK == 0, "cst_gemm(bf16): wrapped K needs lda %% 64 == 0 (lda=%lld)", hp.lda);
  CST_REQUIRE(hp.a_bs_outer % hp.lda == 0 && hp.a_bs_inner % hp.lda == 0, "cst_gemm(bf16): A batch strides must be multiples of lda");
  CST_REQUIRE(nz == 1 || hp.w_bs_inner == (long long)hp.N * hp.K || hp.nb_inner == 1,
              "cst_gemm(bf16): batched W must be densely packed [nb_inner*N, K]");
  CST_REQUIRE(((uintptr_t)hp.A % 16) == 0 && ((uintptr_t)hp.W % 16) == 0, "cst_gemm(bf16): A/W must be 16-byte aligned");
  const long long a_rows_total = ((hp.nb_outer - 1) * hp.a_bs_outer + (hp.nb_inner - 1) * hp.a_bs_inner) / hp.lda + hp.a_rows;
  const long long a_inner = a_wrap ? hp.lda : hp.K;
  CUtensorMap tmA, tmB;
  int rc = make_map_2d(&tmA, hp.A, a_inner, a_rows_total, hp.lda, TC_BK, 128);
  if (rc) return rc;
  rc = make_map_2d(&tmB, hp.W, hp.K, (long long)hp.nb_inner * hp.N, hp.K, TC_BK, 128);
  if (rc) return rc;
  const int bulk = bulk_store_ok(hp, nz, 256);
  rc = check_ln_fused(hp, nz, 256, bulk);
  if (rc) return rc;
  CUtensorMap tmC;
  rc = make_c_map(&tmC, hp, bulk);
  if (rc) return rc;
  const int m_tiles = cdiv(hp.M, 256), n_tiles = hp.N / 256;
  const long long total = (long long)m_tiles * n_tiles * nz;
  CST_REQUIRE(total < (1ll << 31), "cst_gemm(bf16): too many tiles");
  const int pairs = (int)(total < max_pairs ? total : max_pairs);
#ifdef TC_PROFILE
  if (getenv("CST_TC_FIXED_COORDS")) a_wrap |= 2;
  if (const char* e = getenv("CST_TC_SKIP")) a_wrap |= atoi(e) & 12;
#endif
  CST_CHECK_CUDA(launch_k(gemm_tc_pair_kernel<ACT, VAR>, dim3(2 * pairs), dim3(TC_THREADS), TC2_SMEM_BYTES, st, tmA, tmB, tmC, p, m_tiles, n_tiles,
                          (int)total, a_wrap, hp.ab_dtype == CST_F16 ? 1 : 0, bulk));
  return CST_OK;
}

static int launch_tc_pair(const cst_gemm_params& hp, const GemmDev& p, int nz, cudaStream_t st) {
  const int var = needs_variant_kernel(hp, p);
  if (var == 1) return launch_tc_pair_act<CST_ACT_NONE, 1>(hp, p, nz, st);
  if (var == 2) {
    switch (hp.act) {
      case CST_ACT_NONE: return launch_tc_pair_act<CST_ACT_NONE, 2>(hp, p, nz, st);
      case CST_ACT_GELU: return launch_tc_pair_act<CST_ACT_GELU, 2>(hp, p, nz, st);
      case CST_ACT_RELU: return launch_tc_pair_act<CST_ACT_RELU, 2>(hp, p, nz, st);
      default: return launch_tc_pair_act<CST_ACT_GLU, 2>(hp, p, nz, st);
    }
  }
  switch (hp.act) {
    case CST_ACT_NONE: return launch_tc_pair_act<CST_ACT_NONE, 0>(hp, p, nz, st);
    case CST_ACT_GELU: return launch_tc_pair_act<CST_ACT_GELU, 0>(hp, p, nz, st);
    case CST_ACT_RELU: return launch_tc_pair_act<CST_ACT_RELU, 0>(hp, p, nz, st);
    default: return launch_tc_pair_act<CST_ACT_GLU, 0>(hp, p, nz, st);
  }
}

template <int BN>
static int launch_tc(const cst_gemm_params& hp, const GemmDev& p, int nz, cudaStream_t st) {
  const int var = needs_variant_kernel(hp, p);
  if (var == 1) return launch_tc_act<BN, CST_ACT_NONE, 1>(hp, p, nz, st);
  if (var == 2) {
    switch (hp.act) {
      case CST_ACT_NONE: return launch_tc_act<BN, CST_ACT_NONE, 2>(hp, p, nz, st);
      case CST_ACT_GELU: return launch_tc_act<BN, CST_ACT_GELU, 2>(hp, p, nz, st);
      case CST_ACT_RELU: return launch_tc_act<BN, CST_ACT_RELU, 2>(hp, p, nz, st);
      default: return launch_tc_act<BN, CST_ACT_GLU, 2>(hp, p, nz, st);
    }
  }
  switch (hp.act) {
    case CST_ACT_NONE: return launch_tc_act<BN, CST_ACT_NONE, 0>(hp, p, nz, st);
    case CST_ACT_GELU: return launch_tc_act<BN, CST_ACT_GELU, 0>(hp, p, nz, st);
    case CST_ACT_RELU: return launch_tc_act<BN, CST_ACT_RELU, 0>(hp, p, nz, st);
    default: return launch_tc_act<BN, CST_ACT_GLU, 0>(hp, p, nz, st);
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
  // CTA pairs (256x256 tiles): CST_TC_PAIR = 0 off, 1 when the problem has at least `pair_min` pair tiles, 2 always when
  // the shape allows it, 3 selective (default, below).  Measured on B200: before the bulk-store epilogue the pair kernel
  // was at parity with the single-CTA kernel (the epilogue, not the loads, paced both); with it the pair kernel is ahead
  // on the big-M MMA-paced shapes and the selective policy gains 1.4-1.7 % of the c2 step (3 A/B repetitions).
  static const int pair_mode = [] { const char* e = getenv("CST_TC_PAIR"); return e ? atoi(e) : 3; }();
  static const int pair_min = [] { const char* e = getenv("CST_TC_PAIR_MIN"); return e ? atoi(e) : 74; }();
  if (pair_mode && hp.N % 256 == 0 && force_bn == 0) {
    const long long pair_tiles = (long long)cdiv(hp.M, 256) * (hp.N / 256) * nz;
    // mode 3: only where the pair kernel measured ahead in isolation (tools/gemm_rate.py): big-M problems whose tile is
    // MMA- or load-paced (QKV 1121 -> 1220, fc2 1107 -> 1217, conv 1252 -> 1279 TFLOP/s), not the GELU epilogue-paced
    // K = 768 fc1 (1192 -> 1143)
    const bool selective = hp.M >= 16384 && !(hp.act == CST_ACT_GELU && hp.K <= 768);
    if (pair_mode == 2 || (pair_mode == 1 && pair_tiles >= pair_min) || (pair_mode == 3 && selective && pair_tiles >= pair_min))
      return launch_tc_pair(hp, p, nz, st);
  }
  if (hp.N % 256 == 0 && force_bn != 128 && force_bn != 64) return launch_tc<256>(hp, p, nz, st);
  if (force_bn == 64 && hp.N % 64 == 0) return launch_tc<64>(hp, p, nz, st);
  if (hp.N % 128 == 0) return launch_tc<128>(hp, p, nz, st);
  if (hp.N == 48) return launch_tc<48>(hp, p, nz, st);
  if (hp.N % 64 == 0) return launch_tc<64>(hp, p, nz, st);
  CST_REQUIRE(false, "cst_gemm(bf16): N=%d unsupported (multiple of 64, or 48)", hp.N);
  return CST_ERR_ARG;
}

}  // namespace cst
