// Padding-aware flash attention on the 5th-gen tensor cores (bf16 operands, fp32 softmax/accumulate),
// head_dim 64.  One CTA = 128 query rows of one (utterance, head); keys/values stream in 128-row tiles.
//   warp 0      : TMA producer (Q once; K_j,V_j into a 3-stage ring, 128B-swizzled)
//   warp 1      : tcgen05.mma issuer.  S_j = Q K_j^T (M128 N128 K64, both operands K-major) into one of two
//                 TMEM S buffers; PV_j = P_j V_j (M128 N64 K128; P from shared memory K-major, V as an
//                 MN-major B operand straight from the TMA tile -- no transpose) into one of two PV buffers.
//                 S_{j+1} is issued before PV_j so the tensor pipe works while the softmax of tile j runs.
//   warps 2..5  : softmax.  thread = query row: tcgen05.ld the S row, exact -inf key mask (keys >= kv_len[b]),
//                 running max / sum in fp32, P = exp2 in bf16 -> shared memory (manual 128B swizzle), output
//                 accumulator O kept in registers: O = O * alpha_j + PV_j (no TMEM read-modify-write).
// Every query row of the tile is computed (padded query rows are live in the reference); keys beyond kv_len
// contribute exactly zero probability.  V rows of masked keys must be finite (the plan guarantees it).
#include "tc_common.cuh"

namespace cst {

constexpr int FA_BQ = 128, FA_BK = 128, FA_D = 64, FA_KS = 3;
constexpr int FA_THREADS = 192;
constexpr int FA_Q_BYTES = FA_BQ * FA_D * 2;            // 16 KB
constexpr int FA_KV_BYTES = FA_BK * FA_D * 2;           // 16 KB each for K and V
constexpr int FA_P_BYTES = FA_BQ * FA_BK * 2;           // 32 KB (two 64-key halves of 16 KB)
constexpr int FA_SMEM = FA_Q_BYTES + 2 * FA_KS * FA_KV_BYTES + 2 * FA_P_BYTES + 1024 + 256;
constexpr float FA_LOG2E = 1.4426950408889634f;

__global__ void __launch_bounds__(FA_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, long long ldo,
                    int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg, const int32_t* __restrict__ kv_len,
                    int q_col0, int k_col0, int v_col0) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base;
  const uint32_t sK = sQ + FA_Q_BYTES;
  const uint32_t sV = sK + FA_KS * FA_KV_BYTES;
  const uint32_t sP = sV + FA_KS * FA_KV_BYTES;
  const uint32_t bars = sP + 2 * FA_P_BYTES;
  const uint32_t q_full = bars;
  const uint32_t kv_full = bars + 8, kv_empty = kv_full + 8 * FA_KS;
  const uint32_t s_full = kv_empty + 8 * FA_KS, s_empty = s_full + 16;
  const uint32_t p_full = s_empty + 16, pv_full = p_full + 16, pv_empty = pv_full + 16;
  const uint32_t tmem_slot = pv_empty + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * FA_BQ;
  int klen = kv_len ? kv_len[b] : n_kv;
  klen = klen < n_kv ? klen : n_kv;
  const int n_tiles = (klen + FA_BK - 1) / FA_BK;

  if (warp == 0 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < FA_KS; ++s) { mbar_init(kv_full + 8 * s, 1); mbar_init(kv_empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_full + 8 * s, 1); mbar_init(s_empty + 8 * s, 4);
      mbar_init(p_full + 8 * s, 4); mbar_init(pv_full + 8 * s, 1); mbar_init(pv_empty + 8 * s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tS = tmem_base, tPV = tmem_base + 256;        // S buffers at cols 0,128; PV at 256,320

  if (warp == 0) {
    if (lane == 0 && n_tiles > 0) {
      const int q_row = b * q_rows_per_seg + q0;
      mbar_expect_tx(q_full, FA_Q_BYTES);
      tma_load_2d(sQ, &tmQ, q_full, q_col0 + h * FA_D, q_row);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % FA_KS, u = j / FA_KS;
        mbar_wait(kv_empty + 8 * st, (u & 1) ^ 1);
        mbar_expect_tx(kv_full + 8 * st, 2 * FA_KV_BYTES);
        const int k_row = b * kv_rows_per_seg + j * FA_BK;
        tma_load_2d(sK + st * FA_KV_BYTES, &tmK, kv_full + 8 * st, k_col0 + h * FA_D, k_row);
        tma_load_2d(sV + st * FA_KV_BYTES, &tmV, kv_full + 8 * st, v_col0 + h * FA_D, k_row);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && n_tiles > 0) {
      // instruction descriptors: fp32 accumulate, bf16 x bf16; S: N=128 K-major B; PV: N=64 MN-major B
      constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FA_BK >> 3) << 17) | ((uint32_t)(FA_BQ >> 4) << 24);
      constexpr uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(FA_D >> 3) << 17) | ((uint32_t)(FA_BQ >> 4) << 24);
      const uint64_t qdesc = make_sw128_desc(sQ);
      auto issue_s = [&](int j) {
        const int st = j % FA_KS, sb = j & 1;
        mbar_wait(kv_full + 8 * st, (j / FA_KS) & 1);
        mbar_wait(s_empty + 8 * sb, ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint64_t kdesc = make_sw128_desc(sK + st * FA_KV_BYTES);
#pragma unroll
        for (int k = 0; k < FA_D / 16; ++k) tc_mma_bf16(tS + sb * FA_BK, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
        tc_commit(s_full + 8 * sb);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_s(j + 1);
        const int st = j % FA_KS, pb = j & 1;
        mbar_wait(p_full + 8 * pb, (j >> 1) & 1);
        mbar_wait(pv_empty + 8 * pb, ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        int keys = klen - j * FA_BK; keys = keys < FA_BK ? keys : FA_BK;
        const int k16 = (keys + 15) >> 4;
        const uint32_t pbase = sP + pb * FA_P_BYTES, vbase = sV + st * FA_KV_BYTES;
        for (int k = 0; k < k16; ++k) {
          const uint64_t pdesc = make_sw128_desc(pbase + (k >> 2) * (FA_P_BYTES / 2) + (k & 3) * 32);
          const uint64_t vdesc = make_sw128_mn_desc(vbase + k * 2048);
          tc_mma_bf16(tPV + pb * FA_D, pdesc, vdesc, idesc_pv, k != 0);
        }
        tc_commit(pv_full + 8 * pb);
        tc_commit(kv_empty + 8 * st);
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax / output warps =====================
    const int qd = warp & 3;
    const int r = qd * 32 + lane;                        // query row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    float o[FA_D];
#pragma unroll
    for (int i = 0; i < FA_D; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int sb = j & 1;
      mbar_wait(s_full + 8 * sb, (j >> 1) & 1);
      tc_fence_after();
      const uint32_t srow = tS + lane_off + sb * FA_BK;
      const int kbase = j * FA_BK;
      const bool tail = kbase + FA_BK > klen;
      // pass 1: row maximum
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < FA_BK; c += 32) {
        float s[32];
        tmem_ld32(srow + c, s);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float v = (tail && kbase + c + i >= klen) ? -INFINITY : s[i];
          mx = fmaxf(mx, v);
        }
      }
      const float m_new = fmaxf(m_run, mx);              // finite: every visited tile holds >= 1 valid key
      const float alpha = mufu_ex2((m_run - m_new) * FA_LOG2E);
      const float mb = m_new * FA_LOG2E;
      // pass 2: P = exp2(s*log2e - m*log2e) -> bf16 -> swizzled shared memory (packed FFMA2 + MUFU.EX2)
      float rs = 0.f;
      const uint32_t prow = sP + sb * FA_P_BYTES + r * 128;
#pragma unroll 1
      for (int c = 0; c < FA_BK; c += 32) {
        float s[32];
        tmem_ld32(srow + c, s);
        tmem_ld_wait();
        uint32_t pk[16];
        const uint64_t l2e = pk2(FA_LOG2E, FA_LOG2E), nmb = pk2(-mb, -mb);
        uint64_t rs2 = pk2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float a0, a1;
          upk2(ffma2(pk2(s[i], s[i + 1]), l2e, nmb), a0, a1);
          float p0 = mufu_ex2(a0), p1 = mufu_ex2(a1);
          if (tail) {
            if (kbase + c + i >= klen) p0 = 0.f;
            if (kbase + c + i + 1 >= klen) p1 = 0.f;
          }
          rs2 = fadd2(rs2, pk2(p0, p1));
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        { float r0, r1; upk2(rs2, r0, r1); rs += r0 + r1; }
        const uint32_t half = prow + (c >> 6) * (FA_P_BYTES / 2);
        const int ch0 = (c & 63) >> 3;                   // first 16-byte chunk of this 32-key group
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint32_t addr = half + (uint32_t)(((ch0 + q4) ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pk[4 * q4]), "r"(pk[4 * q4 + 1]),
                       "r"(pk[4 * q4 + 2]), "r"(pk[4 * q4 + 3]) : "memory");
        }
      }
      tc_fence_before();                                 // S buffer fully read
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // P visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) { mbar_arrive(s_empty + 8 * sb); mbar_arrive(p_full + 8 * sb); }
      // consume PV_{j-1} (relative to m_{j-1}): O = O * alpha_{j-1} + PV_{j-1}
      if (j > 0) {
        const int pb = (j - 1) & 1;
        mbar_wait(pv_full + 8 * pb, ((j - 1) >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < FA_D; c += 32) {
          float pv[32];
          tmem_ld32(tPV + lane_off + pb * FA_D + c, pv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c + i] = fmaf(o[c + i], alpha_prev, pv[i]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pv_empty + 8 * pb);
      }
      l_run = l_run * alpha + rs;
      m_run = m_new;
      alpha_prev = alpha;
    }
    if (n_tiles > 0) {
      const int pb = (n_tiles - 1) & 1;
      mbar_wait(pv_full + 8 * pb, ((n_tiles - 1) >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < FA_D; c += 32) {
        float pv[32];
        tmem_ld32(tPV + lane_off + pb * FA_D + c, pv);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c + i] = fmaf(o[c + i], alpha_prev, pv[i]);
      }
    }
    if (q0 + r < n_q) {
      const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
      __nv_bfloat16* op = out + ((long long)b * q_rows_per_seg + q0 + r) * ldo + h * FA_D;
#pragma unroll
      for (int i = 0; i < FA_D; i += 8) {
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int launch_attention_tc(const void* q, const void* k, const void* v, void* out, long long ldq, long long ldkv, long long ldo,
                        int B, int H, int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg,
                        const int32_t* kv_len, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    CST_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    attr_set = true;
  }
  CST_REQUIRE(((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0, "cst_attention(bf16): q/k/v must be 16-byte aligned");
  // tensor maps over the full row-major buffers: inner = one head's 64 columns are addressed by coordinate
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_map_2d(&tmQ, q, H * FA_D, (long long)B * q_rows_per_seg, ldq, FA_D, FA_BQ);
  if (rc) return rc;
  rc = make_map_2d(&tmK, k, H * FA_D, (long long)B * kv_rows_per_seg, ldkv, FA_D, FA_BK);
  if (rc) return rc;
  rc = make_map_2d(&tmV, v, H * FA_D, (long long)B * kv_rows_per_seg, ldkv, FA_D, FA_BK);
  if (rc) return rc;
  dim3 grid(cdiv(n_q, FA_BQ), H, B);
  attention_tc_kernel<<<grid, FA_THREADS, FA_SMEM, st>>>(tmQ, tmK, tmV, (__nv_bfloat16*)out, ldo, n_q, q_rows_per_seg,
                                                       n_kv, kv_rows_per_seg, kv_len, 0, 0, 0);
  CST_LAUNCH_CHECK();
  return CST_OK;
}

}  // namespace cst
