// f.4: backward of softmax attention (head_dim 64), fp32 on the CUDA cores -- the derivative of cst_attention's arithmetic
// (F.multi_head_attention_forward as called from fairseq/modules/multihead_attention.py:155-187; q pre-scaled, keys >= kv_len[b]
// masked with -inf, all query rows live).  Flash-style: nothing of size Tq x Tk is stored by the forward pass; this kernel
// recomputes S = Q K^T per 64 x 64 tile, first for the row statistics (max, sum) and D_i = sum_j P_ij dP_ij, then for
//   P = exp(S - m) / l,  dP = dO V^T,  dS = P o (dP - D),  dQ += dS K,  dK += dS^T Q,  dV += P^T dO.
// One CTA per (64-query tile, head, utterance); dK / dV are accumulated across query tiles with fp32 atomics (pre-zeroed by the
// caller).  First correct version: FFMA tiles in shared memory; the tcgen05 version is future work.
#include "common.cuh"

namespace cst {

constexpr int AB_T = 64, AB_D = 64, AB_LD = 65;
constexpr int AB_SMEM = 6 * AB_T * AB_LD * 4 + 3 * AB_T * 4;

template <int TRANS_A>   // C[4ty+a][4tx+c] = sum_k A(4ty+a, k) * B(4tx+c, k); A(i,k) = TRANS_A ? As[k][i] : As[i][k]; same for B
__device__ __forceinline__ void tile_mm(const float* __restrict__ As, const float* __restrict__ Bs, int ty, int tx, float (&acc)[4][4]) {
#pragma unroll 4
  for (int k = 0; k < 64; ++k) {
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[i] = TRANS_A ? As[k * AB_LD + 4 * ty + i] : As[(4 * ty + i) * AB_LD + k];
      b[i] = TRANS_A ? Bs[k * AB_LD + 4 * tx + i] : Bs[(4 * tx + i) * AB_LD + k];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

__device__ __forceinline__ float ab_ld(const void* p, int dt, long long i) {
  return dt == CST_F32 ? reinterpret_cast<const float*>(p)[i] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}

__global__ void __launch_bounds__(256) attention_bwd_kernel(const void* __restrict__ q, const void* __restrict__ k, const void* __restrict__ v,
                                                            const void* __restrict__ o, int dt, const float* __restrict__ d_o,
                                                            float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv,
                                                            long long ldq, long long ldkv, long long ldo_fwd, long long ldo,
                                                            long long lddq, long long lddkv,
                                                            int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg,
                                                            const int32_t* __restrict__ kv_len) {
  extern __shared__ float sm[];
  float* Qs = sm;                       // [i][d]
  float* Gs = Qs + AB_T * AB_LD;        // dO [i][d]
  float* Ks = Gs + AB_T * AB_LD;        // [j][d]
  float* Vs = Ks + AB_T * AB_LD;        // [j][d]
  float* Ps = Vs + AB_T * AB_LD;        // P  [i][j]
  float* Ss = Ps + AB_T * AB_LD;        // dS [i][j]
  float* row_m = Ss + AB_T * AB_LD;     // [64]
  float* row_l = row_m + AB_T;
  float* row_d = row_l + AB_T;
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AB_T;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  int klen = kv_len ? kv_len[b] : n_kv;
  if (klen > n_kv) klen = n_kv;
  const long long qrow0 = (long long)b * q_rows_per_seg, krow0 = (long long)b * kv_rows_per_seg;
  // Q, dO tiles
  for (int e = tid; e < AB_T * AB_D; e += 256) {
    const int i = e >> 6, d = e & 63;
    const bool ok = q0 + i < n_q;
    Qs[i * AB_LD + d] = ok ? ab_ld(q, dt, (qrow0 + q0 + i) * ldq + h * AB_D + d) : 0.f;
    Gs[i * AB_LD + d] = ok ? d_o[(qrow0 + q0 + i) * ldo + h * AB_D + d] : 0.f;
  }
  if (tid < AB_T) { row_d[tid] = 0.f; row_m[tid] = -INFINITY; row_l[tid] = 0.f; }
  __syncthreads();
  auto load_kv = [&](int k0, bool with_v) {
    for (int e = tid; e < AB_T * AB_D; e += 256) {
      const int j = e >> 6, d = e & 63;
      const bool ok = k0 + j < klen;
      Ks[j * AB_LD + d] = ok ? ab_ld(k, dt, (krow0 + k0 + j) * ldkv + h * AB_D + d) : 0.f;
      if (with_v) Vs[j * AB_LD + d] = ok ? ab_ld(v, dt, (krow0 + k0 + j) * ldkv + h * AB_D + d) : 0.f;
    }
  };
  // ---- pass 1: row maximum, sum of exponentials and D_i = sum_j P_ij dP_ij over all keys.  D is accumulated from the recomputed
  // probabilities (un-normalised, rescaled with the running maximum like l) instead of dO_i . O_i: the stored O of the 16-bit mode
  // is bf16, and dS = P o (dP - D) cancels heavily when the values of a sequence are close to each other.
  (void)o; (void)ldo_fwd;
  for (int k0 = 0; k0 < klen; k0 += AB_T) {
    __syncthreads();
    load_kv(k0, true);
    __syncthreads();
    float s[4][4] = {}, dp[4][4] = {};
    tile_mm<0>(Qs, Ks, ty, tx, s);
    tile_mm<0>(Gs, Vs, ty, tx, dp);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        Ps[(4 * ty + a) * AB_LD + 4 * tx + c] = (k0 + 4 * tx + c < klen) ? s[a][c] : -INFINITY;
        Ss[(4 * ty + a) * AB_LD + 4 * tx + c] = dp[a][c];
      }
    __syncthreads();
    {
      // 4 threads per query row (16 keys each), combined with two butterfly steps inside the quad
      const int row = tid >> 2, part = tid & 3;
      const float* ps = Ps + row * AB_LD + part * 16;
      const float* ds = Ss + row * AB_LD + part * 16;
      float mx = row_m[row];
#pragma unroll
      for (int j = 0; j < 16; ++j) mx = fmaxf(mx, ps[j]);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float l = 0.f, dacc = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float e = expf(ps[j] - mx);
        l += e;
        dacc = fmaf(e, ds[j], dacc);
      }
      l += __shfl_xor_sync(0xffffffffu, l, 1); l += __shfl_xor_sync(0xffffffffu, l, 2);
      dacc += __shfl_xor_sync(0xffffffffu, dacc, 1); dacc += __shfl_xor_sync(0xffffffffu, dacc, 2);
      __syncwarp();                                              // the quad's reads of row_m[row] above are ordered before the update
      if (part == 0) {
        const float resc = expf(row_m[row] - mx);
        row_l[row] = row_l[row] * resc + l;
        row_d[row] = row_d[row] * resc + dacc;
        row_m[row] = mx;
      }
    }
  }
  __syncthreads();
  if (tid < AB_T) row_d[tid] = row_l[tid] > 0.f ? row_d[tid] / row_l[tid] : 0.f;
  __syncthreads();
  // ---- pass 2: gradients
  float dq_acc[4][4] = {};
  for (int k0 = 0; k0 < klen; k0 += AB_T) {
    __syncthreads();
    load_kv(k0, true);
    __syncthreads();
    float s[4][4] = {}, dp[4][4] = {};
    tile_mm<0>(Qs, Ks, ty, tx, s);
    tile_mm<0>(Gs, Vs, ty, tx, dp);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = 4 * ty + a;
      const float m = row_m[i], inv_l = row_l[i] > 0.f ? 1.0f / row_l[i] : 0.f, di = row_d[i];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = 4 * tx + c;
        const float p = (k0 + j < klen) ? expf(s[a][c] - m) * inv_l : 0.f;
        Ps[i * AB_LD + j] = p;
        Ss[i * AB_LD + j] = p * (dp[a][c] - di);
      }
    }
    __syncthreads();
    // dV[j][d] += sum_i P[i][j] dO[i][d];  dK[j][d] += sum_i dS[i][j] Q[i][d]     (ty -> j block, tx -> d block)
    float gv[4][4] = {}, gk[4][4] = {};
#pragma unroll 4
    for (int i = 0; i < AB_T; ++i) {
      float pj[4], sj[4], go[4], qq[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        pj[a] = Ps[i * AB_LD + 4 * ty + a]; sj[a] = Ss[i * AB_LD + 4 * ty + a];
        go[a] = Gs[i * AB_LD + 4 * tx + a]; qq[a] = Qs[i * AB_LD + 4 * tx + a];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) { gv[a][c] = fmaf(pj[a], go[c], gv[a][c]); gk[a][c] = fmaf(sj[a], qq[c], gk[a][c]); }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int j = k0 + 4 * ty + a;
      if (j < klen) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          atomicAdd(dv + (krow0 + j) * lddkv + h * AB_D + 4 * tx + c, gv[a][c]);
          atomicAdd(dk + (krow0 + j) * lddkv + h * AB_D + 4 * tx + c, gk[a][c]);
        }
      }
    }
    // dQ[i][d] += sum_j dS[i][j] K[j][d]                                           (ty -> i block, tx -> d block)
#pragma unroll 4
    for (int j = 0; j < AB_T; ++j) {
      float sj[4], kk[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) { sj[a] = Ss[(4 * ty + a) * AB_LD + j]; kk[a] = Ks[j * AB_LD + 4 * tx + a]; }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) dq_acc[a][c] = fmaf(sj[a], kk[c], dq_acc[a][c]);
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = q0 + 4 * ty + a;
    if (i < n_q)
      *reinterpret_cast<float4*>(dq + (qrow0 + i) * lddq + h * AB_D + 4 * tx) = make_float4(dq_acc[a][0], dq_acc[a][1], dq_acc[a][2], dq_acc[a][3]);
  }
}

}  // namespace cst

// q [B*q_rows_per_seg, ldq], k / v [B*kv_rows_per_seg, ldkv], o [.., ldo_fwd] in `dtype` (CST_F32 / CST_BF16: the tape of the 16-bit mode);
// d_o [.., ldo], dq [.., lddq], dk / dv [.., lddkv] fp32; head h at columns h*64.  dq rows >= n_q of a segment are left untouched;
// dk / dv must be ZERO on entry (they are accumulated atomically).
extern "C" int cst_attention_bwd(const void* q, const void* k, const void* v, const void* o, int dtype, const float* d_o,
                                 float* dq, float* dk, float* dv, long long ldq, long long ldkv, long long ldo_fwd, long long ldo,
                                 long long lddq, long long lddkv,
                                 int B, int H, int n_q, int q_rows_per_seg, int n_kv, int kv_rows_per_seg, const int32_t* kv_len,
                                 void* stream) {
  using namespace cst;
  CST_REQUIRE(q && k && v && o && d_o && dq && dk && dv && B > 0 && H > 0 && n_q > 0 && n_kv > 0, "cst_attention_bwd: bad args");
  CST_REQUIRE(dtype == CST_F32 || dtype == CST_BF16, "cst_attention_bwd: dtype must be CST_F32 or CST_BF16");
  CST_REQUIRE(n_q <= q_rows_per_seg && n_kv <= kv_rows_per_seg && lddq % 4 == 0, "cst_attention_bwd: bad geometry");
  static bool attr = false;
  if (!attr) { CST_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM)); attr = true; }
  dim3 grid(cdiv(n_q, AB_T), H, B);
  CST_CHECK_CUDA(launch_k(attention_bwd_kernel, grid, dim3(256), AB_SMEM, (cudaStream_t)stream, q, k, v, o, dtype, d_o, dq, dk, dv, ldq, ldkv,
                          ldo_fwd, ldo, lddq, lddkv, n_q, q_rows_per_seg, n_kv, kv_rows_per_seg, kv_len));
  return CST_OK;
}
