// Greedy incremental decoding of the target tokens from the M memories (BASELINE configs[3]).
//
// One decoding step touches every decoder weight once (27 M parameters, L2-resident after the first step) for at
// most 64 hypotheses: a skinny, weight-streaming problem.  Four kernels, all driven by a DEVICE-side step counter
// so that one captured CUDA graph of a step can be replayed for every step with no host round trip:
//   dec_embed_kernel      x = sqrt(d) * E[token] + sinusoidal position
//   dec_linear_kernel     out = act(LN?(A) W^T + b) (+ residual), up to three output segments (q | K cache | V cache)
//   dec_attention_kernel  one warp per (hypothesis, head): softmax(q K^T) V over the cached keys / the M memories
//   dec_select_kernel     log-softmax, the reference's candidate rules (pad never, EOS only from min_len, only EOS at
//                         max_len), arg-max, token append, finished bookkeeping, step += 1
// fp32 accumulation everywhere (exact fp32 FFMA, the 1e-5 / identical-ID parity mode); weights fp32 or bf16.
//
// Reference arithmetic replaced: TransformerDecoder.extract_features_scriptable (fairseq/models/transformer.py:720-828),
// output_layer (:830-838), TransformerDecoderLayer.forward (fairseq/modules/transformer_layer.py:300-412) with the
// incremental-state MultiheadAttention (fairseq/modules/multihead_attention.py:189-379), SinusoidalPositionalEmbedding
// (fairseq/modules/sinusoidal_positional_embedding.py:71-93) and SequenceGenerator._generate for beam_size = 1
// (fairseq/sequence_generator.py:294-540).
#include "common.cuh"

namespace cst {

constexpr int DL_COLS = 8;                  // output features per column tile
constexpr int DL_THREADS = 256;             // 8 warps
constexpr int DL_A_BYTES = 8 * 4096 * 4;    // every warp keeps RPW rows x K = 4096 fp32 activations resident (128 KB per CTA)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 lds4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                     __uint_as_float(u.y & 0xffff0000u));
}

// exchange-and-add step of the transposing warp reduction: afterwards v[0..N/2) hold the sums of the index half
// selected by lane bit `off`
template <int N, int NV>
__device__ __forceinline__ void bfly(float (&v)[NV], int off, int lane) {
  const bool up = (lane & off) != 0;
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    const float keep = up ? v[i + N / 2] : v[i];
    const float send = up ? v[i] : v[i + N / 2];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
  }
}
// NV per-lane partial sums -> full sums: NV >= 32: lane holds indices lane*(NV/32) + j in v[j]; NV < 32: index lane / (32/NV) in v[0]
template <int NV>
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[NV], int lane) {
  if constexpr (NV >= 2) bfly<NV, NV>(v, 16, lane);
  if constexpr (NV >= 4) bfly<NV / 2, NV>(v, 8, lane);
  if constexpr (NV >= 8) bfly<NV / 4, NV>(v, 4, lane);
  if constexpr (NV >= 16) bfly<NV / 8, NV>(v, 2, lane);
  if constexpr (NV >= 32) bfly<NV / 16, NV>(v, 1, lane);
  if constexpr (NV < 32) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  if constexpr (NV < 16) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
  if constexpr (NV < 8) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 4);
}

struct DecLinearArgs {
  const void* A; long long lda;
  const void* W; const float* bias; const float* ln_g; const float* ln_b;
  const float* residual; long long ldr;
  void* out0; void* out1; void* out2;
  long long ldo0, ldo1, ldo2, ss0, ss1, ss2;
  const int* step;
  int M, N, seg_n, act;
  int bf16_mask;                 // bit s: output segment s is bf16 (K / V cache rows of the 16-bit mode)
};

__device__ __forceinline__ void dec_store_out(const DecLinearArgs& a, int seg, size_t idx, float v) {
  void* o = seg == 0 ? (void*)a.out0 : (seg == 1 ? (void*)a.out1 : (void*)a.out2);
  if ((a.bf16_mask >> seg) & 1) reinterpret_cast<__nv_bfloat16*>(o)[idx] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(o)[idx] = v;
}

// K = 4096 / RPW is a compile-time constant (512, 1024, 2048).  A CTA owns 8*RPW rows (warp w: rows w*RPW ...), keeps
// them resident in shared memory for its whole life (cp.async, all rows in flight at once; optionally LayerNorm-ed in
// place by the owning warp) and walks over 8-column weight tiles t = blockIdx.x, += gridDim.x (K = 512: the next
// tile's weights are prefetched with cp.async into the other buffer while this one is multiplied).  A lane owns the
// K positions 4*lane + 128*j; per 128-wide K step it does 8 + RPW LDS.128 for 32*RPW FMAs; the RPW*8 partial sums per
// lane are reduced across the warp by a transposing butterfly (62 shuffles for 64 values).
template <typename AT, typename WT, int RPW>
__global__ void __launch_bounds__(DL_THREADS, 1) dec_linear_kernel(const DecLinearArgs a) {
  constexpr int K = 4096 / RPW, NV = RPW * 8, NBUF = (K == 512) ? 2 : 1;
  constexpr int W_UNITS_PER_COL = K * (int)sizeof(WT) / 16, W_ELEMS_PER_UNIT = 16 / (int)sizeof(WT);
  extern __shared__ __align__(16) unsigned char dl_smem[];
  float* sA = reinterpret_cast<float*>(dl_smem);
  WT* sW = reinterpret_cast<WT*>(dl_smem + DL_A_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * 8 * RPW + warp * RPW;
  const int rows_here = min(RPW, a.M - m0);                    // warp-uniform; may be <= 0
  float* myA = sA + warp * RPW * K;
  const AT* A = reinterpret_cast<const AT*>(a.A);
  const WT* W = reinterpret_cast<const WT*>(a.W);
  const int ntiles = (a.N + DL_COLS - 1) / DL_COLS;
  pdl_launch_dependents();

  auto load_w = [&](int tile, int buf) {
    WT* dst = sW + (size_t)buf * DL_COLS * K;
    for (int u = threadIdx.x; u < DL_COLS * W_UNITS_PER_COL; u += DL_THREADS) {
      const int c = u / W_UNITS_PER_COL, e = (u - c * W_UNITS_PER_COL) * W_ELEMS_PER_UNIT;
      const int n = tile * DL_COLS + c;
      if (n < a.N) cp_async16(dst + c * K + e, W + (size_t)n * K + e);
      else *reinterpret_cast<uint4*>(dst + c * K + e) = make_uint4(0u, 0u, 0u, 0u);
    }
  };

  int t = blockIdx.x;
  if (NBUF == 2 && t < ntiles) load_w(t, 0);                   // weights do not depend on the previous kernel
  pdl_wait();
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
#pragma unroll
    for (int j = 0; j < K / 128; ++j) {
      float* dst = myA + r * K + 4 * lane + 128 * j;
      if (r < rows_here) {
        const AT* src = A + (size_t)(m0 + r) * a.lda + 4 * lane + 128 * j;
        if constexpr (sizeof(AT) == 4) cp_async16(dst, src);
        else *reinterpret_cast<float4*>(dst) = load4(src);
      } else {
        *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  cp_async_commit();

  const long long step = a.step ? (long long)*a.step : 0;
  for (int it = 0; t < ntiles; t += gridDim.x, ++it) {
    const int buf = NBUF == 2 ? (it & 1) : 0;
    if (NBUF == 1) load_w(t, 0);
    else if (t + (int)gridDim.x < ntiles) load_w(t + gridDim.x, buf ^ 1);
    cp_async_commit();
    cp_async_wait<NBUF - 1>();
    __syncthreads();
    if constexpr (K == 512) if (it == 0 && a.ln_g != nullptr) {   // K == 512: LayerNorm of this warp's own rows, in place
      // three sweeps over the warp's rows so that the RPW shuffle reductions of a sweep are independent chains
      float mean[RPW], rstd[RPW];
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < K / 128; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(myA + r * K + 4 * lane + 128 * j);
          s += (v.x + v.y) + (v.z + v.w);
        }
        mean[r] = s;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int r = 0; r < RPW; ++r) mean[r] += __shfl_xor_sync(0xffffffffu, mean[r], o);
      }
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        mean[r] *= (1.0f / K);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < K / 128; ++j) {
          float4 v = *reinterpret_cast<const float4*>(myA + r * K + 4 * lane + 128 * j);
          v.x -= mean[r]; v.y -= mean[r]; v.z -= mean[r]; v.w -= mean[r];
          q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
        rstd[r] = q;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int r = 0; r < RPW; ++r) rstd[r] += __shfl_xor_sync(0xffffffffu, rstd[r], o);
      }
#pragma unroll
      for (int r = 0; r < RPW; ++r) rstd[r] = 1.0f / sqrtf(rstd[r] * (1.0f / K) + 1e-5f);
#pragma unroll
      for (int j = 0; j < K / 128; ++j) {
        const float4 g = load4(a.ln_g + 4 * lane + 128 * j), b = load4(a.ln_b + 4 * lane + 128 * j);
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
          float4 v = *reinterpret_cast<const float4*>(myA + r * K + 4 * lane + 128 * j);
          v.x = (v.x - mean[r]) * rstd[r] * g.x + b.x; v.y = (v.y - mean[r]) * rstd[r] * g.y + b.y;
          v.z = (v.z - mean[r]) * rstd[r] * g.z + b.z; v.w = (v.w - mean[r]) * rstd[r] * g.w + b.w;
          *reinterpret_cast<float4*>(myA + r * K + 4 * lane + 128 * j) = v;
        }
      }
      __syncwarp();
    }
    if (rows_here > 0) {
      const WT* wt = sW + (size_t)buf * DL_COLS * K;
      float acc[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i] = 0.f;
#pragma unroll 1
      for (int j = 0; j < K / 128; ++j) {
        float4 w[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) w[c] = lds4(wt + c * K + 4 * lane + 128 * j);
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
          const float4 x = *reinterpret_cast<const float4*>(myA + r * K + 4 * lane + 128 * j);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float v = acc[r * 8 + c];
            v = fmaf(x.x, w[c].x, v); v = fmaf(x.y, w[c].y, v); v = fmaf(x.z, w[c].z, v); v = fmaf(x.w, w[c].w, v);
            acc[r * 8 + c] = v;
          }
        }
      }
      warp_transpose_reduce<NV>(acc, lane);
      constexpr int VPL = NV >= 32 ? NV / 32 : 1;              // outputs per lane
      constexpr int LPV = NV >= 32 ? 1 : 32 / NV;              // lanes holding the same output
      if (lane % LPV == 0) {
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
          const int idx = NV >= 32 ? lane * VPL + j : lane / LPV;
          const int m = m0 + idx / 8, n = t * DL_COLS + (idx & 7);
          if (m < a.M && n < a.N) {
            float v = acc[j];
            if (a.bias) v += a.bias[n];
            if (a.act == CST_ACT_RELU) v = fmaxf(v, 0.f);
            const int seg = n / a.seg_n, col = n - seg * a.seg_n;
            if (a.residual) v += a.residual[(size_t)m * a.ldr + n];
            const long long ldo = seg == 0 ? a.ldo0 : (seg == 1 ? a.ldo1 : a.ldo2);
            const long long ss = seg == 0 ? a.ss0 : (seg == 1 ? a.ss1 : a.ss2);
            dec_store_out(a, seg, (size_t)m * ldo + step * ss + col, v);
          }
        }
      }
    }
    __syncthreads();                                           // the weight buffer just read is the next prefetch target
  }
  cp_async_wait<0>();
}

// ---- bf16 weights: the same linear on the tensor cores (warp-level mma.sync m16n8k16, fp32 accumulate) -----------------
// A decoding step is latency-bound (51 dependent launches, <= 64 rows), so the point of the tensor-core form is not
// FLOP/s but a short CTA life and a small footprint: activations are LayerNorm-ed in registers and kept as bf16
// (64 KB instead of 128 KB), 100 KB of shared memory and <= 128 registers let two CTAs share an SM, so with PDL the next
// kernel's CTAs are resident and have their weight tile in flight while this kernel drains.  tcgen05 needs >= 64
// accumulator rows per instruction and a TMEM round trip per 8-column tile; at M <= 64, N-tile 8 that buys nothing.
// CTA = RB rows (K * RB = 32768: K 512 / 1024 / 2048 -> 64 / 32 / 16 rows) x 8-column weight tiles; the 8 warps split K
// (16 MMAs per warp per tile), partial tiles are summed through shared memory.  Row padding of 8 bf16 makes the row
// stride 4 banks mod 32: every fragment load (lane -> row lane/4, k pair lane%4) is conflict-free.
constexpr int DM_PAD = 8;

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int RB> struct DmCfg {
  static constexpr int K = 32768 / RB, LD = K + DM_PAD, NBUF = (K == 512) ? 2 : 1;
  static constexpr int A_BYTES = RB * LD * 2, W_BYTES = NBUF * DL_COLS * LD * 2, RED_BYTES = 8 * RB * DL_COLS * 4;
  static constexpr int LN_BYTES = (K == 512) ? 2 * K * 4 : 0;          // gamma | beta staged before griddepcontrol.wait
  static constexpr int SMEM = A_BYTES + W_BYTES + RED_BYTES + LN_BYTES;
};

template <typename AT, int RB>
__global__ void __launch_bounds__(DL_THREADS, 2) dec_linear_mma_kernel(const DecLinearArgs a) {
  using Cfg = DmCfg<RB>;
  constexpr int K = Cfg::K, LD = Cfg::LD, NBUF = Cfg::NBUF;
  constexpr int MT = RB / 16, KW = K / 8, KS = KW / 16, RW = RB / 8, V4 = K / 128, HR = RW / 2;
  static_assert(HR * V4 == 16, "half of a warp's rows = 16 float4 per lane");
  extern __shared__ __align__(16) unsigned char dm_smem[];
  __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(dm_smem);
  __nv_bfloat16* sW = reinterpret_cast<__nv_bfloat16*>(dm_smem + Cfg::A_BYTES);
  float* red = reinterpret_cast<float*>(dm_smem + Cfg::A_BYTES + Cfg::W_BYTES);
  float* sln = reinterpret_cast<float*>(dm_smem + Cfg::A_BYTES + Cfg::W_BYTES + Cfg::RED_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  const int row0 = blockIdx.y * RB;
  const AT* A = reinterpret_cast<const AT*>(a.A);
  const __nv_bfloat16* W = reinterpret_cast<const __nv_bfloat16*>(a.W);
  const int ntiles = (a.N + DL_COLS - 1) / DL_COLS;
  pdl_launch_dependents();

  auto load_w = [&](int tile, int buf) {
    __nv_bfloat16* dst = sW + (size_t)buf * DL_COLS * LD;
    for (int u = tid; u < DL_COLS * (K / 8); u += DL_THREADS) {
      const int c = u / (K / 8), e = (u - c * (K / 8)) * 8;
      const int n = tile * DL_COLS + c;
      if (n < a.N) cp_async16(dst + c * LD + e, W + (size_t)n * K + e);
      else *reinterpret_cast<uint4*>(dst + c * LD + e) = make_uint4(0u, 0u, 0u, 0u);
    }
  };

  int t = blockIdx.x;
  if (NBUF == 2 && t < ntiles) load_w(t, 0);                   // weights do not depend on the previous kernel
  if constexpr (K == 512) if (a.ln_g != nullptr) {             // nor do the LayerNorm parameters: 256 threads x 16 bytes =
    const float* src = tid < 128 ? a.ln_g + 4 * tid : a.ln_b + 4 * (tid - 128);   // gamma[512] | beta[512], staged once per CTA
    cp_async16(sln + 4 * tid, src);
  }
  cp_async_commit();
  pdl_wait();
  // activations: this warp's RW rows, two halves of 16 float4 per lane in flight; LayerNorm in registers; bf16 to smem
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float4 v[HR][V4];
#pragma unroll
    for (int r = 0; r < HR; ++r) {
      const int m = row0 + warp * RW + half * HR + r;
#pragma unroll
      for (int j = 0; j < V4; ++j)
        v[r][j] = m < a.M ? load4(A + (size_t)m * a.lda + 4 * lane + 128 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if constexpr (K == 512) if (a.ln_g != nullptr) {
      float mean[HR], rstd[HR];
#pragma unroll
      for (int r = 0; r < HR; ++r) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < V4; ++j) s += (v[r][j].x + v[r][j].y) + (v[r][j].z + v[r][j].w);
        mean[r] = s;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int r = 0; r < HR; ++r) mean[r] += __shfl_xor_sync(0xffffffffu, mean[r], o);
      }
#pragma unroll
      for (int r = 0; r < HR; ++r) {
        mean[r] *= (1.0f / K);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < V4; ++j) {
          v[r][j].x -= mean[r]; v[r][j].y -= mean[r]; v[r][j].z -= mean[r]; v[r][j].w -= mean[r];
          q += (v[r][j].x * v[r][j].x + v[r][j].y * v[r][j].y) + (v[r][j].z * v[r][j].z + v[r][j].w * v[r][j].w);
        }
        rstd[r] = q;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int r = 0; r < HR; ++r) rstd[r] += __shfl_xor_sync(0xffffffffu, rstd[r], o);
      }
      if (half == 0) {                                           // gamma / beta (and weight tile 0) staged by all threads
        cp_async_wait<0>();
        __syncthreads();                                         // CTA-uniform branch: ln_g is a kernel argument
      }
#pragma unroll
      for (int j = 0; j < V4; ++j) {
        const float4 gm = *reinterpret_cast<const float4*>(sln + 4 * lane + 128 * j);
        const float4 bt = *reinterpret_cast<const float4*>(sln + K + 4 * lane + 128 * j);
#pragma unroll
        for (int r = 0; r < HR; ++r) {
          const float rs = 1.0f / sqrtf(rstd[r] * (1.0f / K) + 1e-5f);
          v[r][j].x = v[r][j].x * rs * gm.x + bt.x; v[r][j].y = v[r][j].y * rs * gm.y + bt.y;
          v[r][j].z = v[r][j].z * rs * gm.z + bt.z; v[r][j].w = v[r][j].w * rs * gm.w + bt.w;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < HR; ++r) {
      __nv_bfloat16* dst = sA + (size_t)(warp * RW + half * HR + r) * LD + 4 * lane;
#pragma unroll
      for (int j = 0; j < V4; ++j)
        *reinterpret_cast<uint2*>(dst + 128 * j) = make_uint2(pack_bf16x2(v[r][j].x, v[r][j].y), pack_bf16x2(v[r][j].z, v[r][j].w));
    }
  }

  const long long step = a.step ? (long long)*a.step : 0;
  for (int it = 0; t < ntiles; t += gridDim.x, ++it) {
    const int buf = NBUF == 2 ? (it & 1) : 0;
    if (NBUF == 1) load_w(t, 0);
    else if (t + (int)gridDim.x < ntiles) load_w(t + gridDim.x, buf ^ 1);
    cp_async_commit();
    cp_async_wait<NBUF - 1>();
    __syncthreads();                                           // weight tile + (first iteration) all activation rows visible
    const __nv_bfloat16* wt = sW + (size_t)buf * DL_COLS * LD;
    float c[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) { c[mt][0] = 0.f; c[mt][1] = 0.f; c[mt][2] = 0.f; c[mt][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int k0 = warp * KW + ks * 16 + 2 * tq;
      uint32_t bfr[2];
      bfr[0] = *reinterpret_cast<const uint32_t*>(wt + g * LD + k0);
      bfr[1] = *reinterpret_cast<const uint32_t*>(wt + g * LD + k0 + 8);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const __nv_bfloat16* ap = sA + (size_t)(mt * 16 + g) * LD + k0;
        uint32_t afr[4];
        afr[0] = *reinterpret_cast<const uint32_t*>(ap);
        afr[1] = *reinterpret_cast<const uint32_t*>(ap + 8 * LD);
        afr[2] = *reinterpret_cast<const uint32_t*>(ap + 8);
        afr[3] = *reinterpret_cast<const uint32_t*>(ap + 8 * LD + 8);
        mma_bf16_16816(c[mt], afr, bfr);
      }
    }
    float* myred = red + warp * RB * DL_COLS;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      *reinterpret_cast<float2*>(myred + (mt * 16 + g) * DL_COLS + 2 * tq) = make_float2(c[mt][0], c[mt][1]);
      *reinterpret_cast<float2*>(myred + (mt * 16 + g + 8) * DL_COLS + 2 * tq) = make_float2(c[mt][2], c[mt][3]);
    }
    __syncthreads();
    for (int idx = tid; idx < RB * DL_COLS; idx += DL_THREADS) {
      const int m = row0 + idx / DL_COLS, n = t * DL_COLS + (idx & (DL_COLS - 1));
      if (m < a.M && n < a.N) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w * RB * DL_COLS + idx];
        if (a.bias) v += a.bias[n];
        if (a.act == CST_ACT_RELU) v = fmaxf(v, 0.f);
        const int seg = n / a.seg_n, col = n - seg * a.seg_n;
        if (a.residual) v += a.residual[(size_t)m * a.ldr + n];
        const long long ldo = seg == 0 ? a.ldo0 : (seg == 1 ? a.ldo1 : a.ldo2);
        const long long ss = seg == 0 ? a.ss0 : (seg == 1 ? a.ss1 : a.ss2);
        dec_store_out(a, seg, (size_t)m * ldo + step * ss + col, v);
      }
    }
    __syncthreads();                                           // `red` and the weight buffer just read are reused next
  }
  cp_async_wait<0>();
}

// x[b,:] = scale * E[tokens[b, step], :] + pos[step, :]
template <typename WT>
__global__ void dec_embed_kernel(const int* __restrict__ tokens, int ld_tok, const WT* __restrict__ E,
                                 const float* __restrict__ pos, float scale, float* __restrict__ x, int C,
                                 const int* __restrict__ step) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, t = *step;
  const int tok = tokens[(size_t)b * ld_tok + t];
  for (int c = 4 * threadIdx.x; c < C; c += 4 * blockDim.x) {
    const float4 e = load4(E + (size_t)tok * C + c), p = load4(pos + (size_t)t * C + c);
    store4(x + (size_t)b * C + c, make_float4(fmaf(scale, e.x, p.x), fmaf(scale, e.y, p.y), fmaf(scale, e.z, p.z), fmaf(scale, e.w, p.w)));
  }
}

// one CTA (8 warps) per (hypothesis, head); head_dim 64; q pre-scaled.  n keys = *step + 1 (self-attention over the
// cache) or n_keys (cross-attention over the memories; the reference passes an all-False key-padding mask).
// Scores: 4 lanes per key (16 dims each, 4 independent LDG.128 per lane), 8 keys per warp pass, 64 keys per CTA pass;
// fp32 softmax over the score row in shared memory; PV: warp w accumulates keys w, w+8, ... (lane = 2 output dims,
// coalesced 256-byte V rows), the 8 partial rows are summed through shared memory.
constexpr int DA_THREADS = 256;
__device__ __forceinline__ float2 load2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 load2(const __nv_bfloat16* p) {
  const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
template <typename KVT>
__global__ void __launch_bounds__(DA_THREADS) dec_attention_kernel(const float* __restrict__ q, long long ldq,
                                                                   const KVT* __restrict__ k, const KVT* __restrict__ v,
                                                                   long long kv_bs, long long kv_rs, float* __restrict__ out,
                                                                   long long ldo, int H, int n_keys, int n_max,
                                                                   const int* __restrict__ step, int kv_group) {
  extern __shared__ __align__(16) float da_smem[];            // [64 q][8*64 partial o][16 red][n_max scores]
  float* sq = da_smem;
  float* po = da_smem + 64;
  float* red = po + 8 * 64;
  float* ss = red + 16;
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int n = step ? min(*step + 1, n_max) : n_keys;
  if (tid < 64) sq[tid] = q[(size_t)b * ldq + h * 64 + tid];
  __syncthreads();
  const KVT* kb = k + (size_t)(b / kv_group) * kv_bs + h * 64;      // kv_group rows share one K / V set (the K beams of a sentence)
  const KVT* vb = v + (size_t)(b / kv_group) * kv_bs + h * 64;
  const int sub = lane & 3;                                    // which 16-dim quarter of the head this lane covers
  float4 qq[4];
#pragma unroll
  for (int d = 0; d < 4; ++d) qq[d] = *reinterpret_cast<const float4*>(sq + 16 * sub + 4 * d);
  float mx = -INFINITY;
  for (int key0 = 0; key0 < n; key0 += 64) {
    const int key = key0 + warp * 8 + (lane >> 2);
    float s = 0.f;
    if (key < n) {
      const KVT* kr = kb + (size_t)key * kv_rs + 16 * sub;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const float4 kk = load4(kr + 4 * d);
        s = fmaf(kk.x, qq[d].x, s); s = fmaf(kk.y, qq[d].y, s); s = fmaf(kk.z, qq[d].z, s); s = fmaf(kk.w, qq[d].w, s);
      }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (key < n) {
      if (sub == 0) ss[key] = s;
      mx = fmaxf(mx, s);
    }
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  float sum = 0.f;
  for (int key = tid; key < n; key += DA_THREADS) {
    const float p = expf(ss[key] - mx);
    ss[key] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[8 + warp] = sum;
  __syncthreads();
  float o0 = 0.f, o1 = 0.f;
  for (int key = warp; key < n; key += 8) {
    const float p = ss[key];
    const float2 vv = load2(vb + (size_t)key * kv_rs + 2 * lane);
    o0 = fmaf(p, vv.x, o0);
    o1 = fmaf(p, vv.y, o1);
  }
  *reinterpret_cast<float2*>(po + warp * 64 + 2 * lane) = make_float2(o0, o1);
  __syncthreads();
  if (tid < 64) {
    float tot = 0.f, o = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { tot += red[8 + w]; o += po[w * 64 + tid]; }
    out[(size_t)b * ldo + h * 64 + tid] = o / tot;
  }
}

// counters: [0] step, [1] CTAs finished (scratch), [2] hypotheses finished
constexpr int DS_THREADS = 1024;
__global__ void __launch_bounds__(DS_THREADS) dec_select_kernel(const float* __restrict__ logits, int V, int* __restrict__ tokens,
                                                                int ld_tok, float* __restrict__ pos_scores, int ld_ps,
                                                                int* __restrict__ done, int* __restrict__ out_len,
                                                                int* counters, int max_len, int min_len, int pad, int eos) {
  __shared__ float s_f[32];
  __shared__ float s_b[32];
  __shared__ int s_i[32];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int step = *reinterpret_cast<volatile int*>(counters);
  const float* row = logits + (size_t)b * V;
  const bool only_eos = step >= max_len, no_eos = step < min_len;
  // the row stays in registers for both sweeps: V <= 3 * 4 * 1024 elements take the vector path (V % 4 == 0)
  constexpr int NV4 = 3;
  float4 x[NV4];
  const bool vec = (V % 4 == 0) && (V <= NV4 * 4 * DS_THREADS);
  float m_all = -INFINITY, best = -INFINITY;
  int bi = 0x7fffffff;
  auto consider = [&](float v, int i) {
    m_all = fmaxf(m_all, v);
    const bool allowed = (i != pad) && (only_eos ? (i == eos) : (no_eos ? (i != eos) : true));
    if (allowed && (v > best || (v == best && i < bi))) { best = v; bi = i; }
  };
  if (vec) {
#pragma unroll
    for (int u = 0; u < NV4; ++u) {
      const int i = 4 * (tid + u * DS_THREADS);
      x[u] = i < V ? load4(row + i) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
#pragma unroll
    for (int u = 0; u < NV4; ++u) {
      const int i = 4 * (tid + u * DS_THREADS);
      if (i < V) { consider(x[u].x, i); consider(x[u].y, i + 1); consider(x[u].z, i + 2); consider(x[u].w, i + 3); }
    }
  } else {
    for (int i = tid; i < V; i += DS_THREADS) consider(row[i], i);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m_all = fmaxf(m_all, __shfl_xor_sync(0xffffffffu, m_all, o));
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) { s_f[warp] = m_all; s_b[warp] = best; s_i[warp] = bi; }
  __syncthreads();
  m_all = s_f[0]; best = s_b[0]; bi = s_i[0];
#pragma unroll
  for (int w = 1; w < DS_THREADS / 32; ++w) {
    m_all = fmaxf(m_all, s_f[w]);
    if (s_b[w] > best || (s_b[w] == best && s_i[w] < bi)) { best = s_b[w]; bi = s_i[w]; }
  }
  __syncthreads();
  float sum = 0.f;
  if (vec) {
#pragma unroll
    for (int u = 0; u < NV4; ++u)     // padding lanes hold -inf: expf(-inf) = 0
      sum += (expf(x[u].x - m_all) + expf(x[u].y - m_all)) + (expf(x[u].z - m_all) + expf(x[u].w - m_all));
  } else {
    for (int i = tid; i < V; i += DS_THREADS) sum += expf(row[i] - m_all);
  }
  sum = warp_sum(sum);
  if (lane == 0) s_f[warp] = sum;
  __syncthreads();
  if (tid == 0) {
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < DS_THREADS / 32; ++w) sum += s_f[w];
    if (bi < 0 || bi >= V) { bi = eos; best = row[eos]; }      // nothing selectable (non-finite logits): stop the hypothesis
    const float lp = best - (m_all + logf(sum));
    if (step > max_len) return;                                // replayed past the last step: nothing to do
    if (!done[b]) {
      tokens[(size_t)b * ld_tok + step + 1] = bi;
      pos_scores[(size_t)b * ld_ps + step] = lp;
      if (bi == eos) { done[b] = 1; out_len[b] = step + 1; atomicAdd(&counters[2], 1); }
    } else {
      tokens[(size_t)b * ld_tok + step + 1] = eos;             // finished rows keep stepping; their output is ignored
    }
    __threadfence();
    if (atomicAdd(&counters[1], 1) == (int)gridDim.x - 1) {   // last CTA of the step: every CTA has read `step`
      counters[1] = 0;
      counters[0] = step + 1;
    }
  }
}

// decoder kernels are launched with programmatic dependent launch by default (CST_DEC_PDL=0 turns it off): a step is a
// chain of 51 short dependent kernels, and the linear kernel prefetches its first weight tile before griddepcontrol.wait
inline bool dec_pdl_enabled() {
  static const bool on = [] { const char* e = getenv("CST_DEC_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_dec(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = dec_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

}  // namespace cst

using namespace cst;

extern "C" int cst_dec_embed(const int32_t* tokens, int ld_tok, const void* embed, int w_dtype, const float* pos_table,
                             float scale, float* x, int B, int C, const int32_t* step, void* stream) {
  CST_REQUIRE(tokens && embed && pos_table && x && step, "cst_dec_embed: null pointer");
  CST_REQUIRE(B > 0 && C > 0 && C % 4 == 0, "cst_dec_embed: bad B=%d C=%d", B, C);
  cudaStream_t st = (cudaStream_t)stream;
  if (w_dtype == CST_F32)
    CST_CHECK_CUDA(launch_dec(dec_embed_kernel<float>, dim3(B), dim3(128), 0, st, tokens, ld_tok, (const float*)embed, pos_table, scale, x, C, step));
  else if (w_dtype == CST_BF16)
    CST_CHECK_CUDA(launch_dec(dec_embed_kernel<__nv_bfloat16>, dim3(B), dim3(128), 0, st, tokens, ld_tok, (const __nv_bfloat16*)embed, pos_table, scale, x, C, step));
  else
    CST_REQUIRE(false, "cst_dec_embed: unsupported dtype %d", w_dtype);
  return CST_OK;
}

template <int RPW>
static constexpr int dl_smem_bytes(int wsize) { return DL_A_BYTES + ((4096 / RPW) == 512 ? 2 : 1) * DL_COLS * (4096 / RPW) * wsize; }

template <typename AT, typename WT, int RPW>
static cudaError_t dl_set_attr() {
  return cudaFuncSetAttribute(dec_linear_kernel<AT, WT, RPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, dl_smem_bytes<RPW>((int)sizeof(WT)));
}
template <typename AT, typename WT>
static cudaError_t dl_set_attr_all() {
  cudaError_t e = dl_set_attr<AT, WT, 8>();
  if (e == cudaSuccess) e = dl_set_attr<AT, WT, 4>();
  if (e == cudaSuccess) e = dl_set_attr<AT, WT, 2>();
  return e;
}
static int g_dl_sms = 148;
static int dec_linear_init() {
  // every instantiation up front: the first launch of one of them may happen inside a stream capture
  CST_CHECK_CUDA((dl_set_attr_all<float, float>()));
  CST_CHECK_CUDA((dl_set_attr_all<float, __nv_bfloat16>()));
  CST_CHECK_CUDA((dl_set_attr_all<__nv_bfloat16, float>()));
  CST_CHECK_CUDA((dl_set_attr_all<__nv_bfloat16, __nv_bfloat16>()));
  CST_CHECK_CUDA(cudaFuncSetAttribute(dec_linear_mma_kernel<float, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, DmCfg<64>::SMEM));
  CST_CHECK_CUDA(cudaFuncSetAttribute(dec_linear_mma_kernel<float, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, DmCfg<32>::SMEM));
  CST_CHECK_CUDA(cudaFuncSetAttribute(dec_linear_mma_kernel<float, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, DmCfg<16>::SMEM));
  CST_CHECK_CUDA(cudaFuncSetAttribute(dec_linear_mma_kernel<__nv_bfloat16, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, DmCfg<64>::SMEM));
  CST_CHECK_CUDA(cudaFuncSetAttribute(dec_linear_mma_kernel<__nv_bfloat16, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, DmCfg<32>::SMEM));
  CST_CHECK_CUDA(cudaFuncSetAttribute(dec_linear_mma_kernel<__nv_bfloat16, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, DmCfg<16>::SMEM));
  int dev = 0;
  CST_CHECK_CUDA(cudaGetDevice(&dev));
  CST_CHECK_CUDA(cudaDeviceGetAttribute(&g_dl_sms, cudaDevAttrMultiProcessorCount, dev));
  // CST_DEC_SMS: SM count the K = 512 linears size their persistent grids for (lever for concurrent decode lanes: smaller grids
  // leave room for the kernels of the other lanes)
  if (const char* e = getenv("CST_DEC_SMS")) { if (atoi(e) > 0) g_dl_sms = atoi(e); }
  return CST_OK;
}

// bf16 weights: tensor-core kernel unless CST_DEC_MMA=0 (then the exact-fp32-activation FFMA kernel; read per call so
// that tests can exercise both)
static bool dec_mma_enabled() { const char* e = getenv("CST_DEC_MMA"); return !(e && e[0] == '0'); }

template <typename AT, int RB>
static int launch_dec_linear_mma_rb(const DecLinearArgs& a, cudaStream_t st) {
  const int ntiles = cdiv(a.N, DL_COLS), gy = cdiv(a.M, RB);
  const int gx = (DmCfg<RB>::K == 512) ? min(ntiles, max(1, 2 * g_dl_sms / gy)) : ntiles;     // two CTAs per SM
  CST_CHECK_CUDA(launch_dec(dec_linear_mma_kernel<AT, RB>, dim3(gx, gy), dim3(DL_THREADS), (size_t)DmCfg<RB>::SMEM, st, a));
  return CST_OK;
}
template <typename AT>
static int launch_dec_linear_mma(const DecLinearArgs& a, int K, cudaStream_t st) {
  if (K == 512) return launch_dec_linear_mma_rb<AT, 64>(a, st);
  if (K == 1024) return launch_dec_linear_mma_rb<AT, 32>(a, st);
  return launch_dec_linear_mma_rb<AT, 16>(a, st);
}

template <typename AT, typename WT, int RPW>
static int launch_dec_linear_rpw(const DecLinearArgs& a, cudaStream_t st) {
  const int ntiles = cdiv(a.N, DL_COLS), gy = cdiv(a.M, 8 * RPW);
  // K = 512: persistent over column tiles (double-buffered weights), one CTA per SM; larger K: one tile per CTA
  const int gx = (4096 / RPW == 512) ? min(ntiles, max(1, g_dl_sms / gy)) : ntiles;
  CST_CHECK_CUDA(launch_dec(dec_linear_kernel<AT, WT, RPW>, dim3(gx, gy), dim3(DL_THREADS), (size_t)dl_smem_bytes<RPW>((int)sizeof(WT)), st, a));
  return CST_OK;
}

template <typename AT, typename WT>
static int launch_dec_linear(const DecLinearArgs& a, int K, cudaStream_t st) {
  static int init_rc = dec_linear_init();
  if (init_rc != CST_OK) return init_rc;
  if constexpr (sizeof(WT) == 2) {
    if (dec_mma_enabled()) return launch_dec_linear_mma<AT>(a, K, st);
  }
  if (K == 512) return launch_dec_linear_rpw<AT, WT, 8>(a, st);
  if (K == 1024) return launch_dec_linear_rpw<AT, WT, 4>(a, st);
  return launch_dec_linear_rpw<AT, WT, 2>(a, st);
}

extern "C" int cst_dec_linear(const cst_dec_linear_params* p, void* stream) {
  CST_REQUIRE(p && p->A && p->W && p->out[0], "cst_dec_linear: null pointer");
  CST_REQUIRE(p->M > 0 && p->N > 0 && (p->K == 512 || p->K == 1024 || p->K == 2048), "cst_dec_linear: K=%d must be 512, 1024 or 2048", p->K);
  CST_REQUIRE(p->n_seg >= 1 && p->n_seg <= 3 && p->N % p->n_seg == 0, "cst_dec_linear: bad n_seg=%d for N=%d", p->n_seg, p->N);
  CST_REQUIRE(!(p->ln_gamma || p->ln_beta) || (p->ln_gamma && p->ln_beta && p->K == 512), "cst_dec_linear: fused LayerNorm needs K == 512");
  CST_REQUIRE(!p->residual || p->n_seg == 1, "cst_dec_linear: residual only with one output segment");
  CST_REQUIRE(p->act == CST_ACT_NONE || p->act == CST_ACT_RELU, "cst_dec_linear: act %d unsupported", p->act);
  CST_REQUIRE(p->lda % 4 == 0, "cst_dec_linear: lda must be a multiple of 4");
  for (int s = 0; s < p->n_seg; ++s) CST_REQUIRE(p->out[s], "cst_dec_linear: out[%d] is null", s);
  DecLinearArgs a;
  a.A = p->A; a.lda = p->lda; a.W = p->W; a.bias = p->bias; a.ln_g = p->ln_gamma; a.ln_b = p->ln_beta;
  a.residual = p->residual; a.ldr = p->ldr;
  a.out0 = p->out[0]; a.out1 = p->out[1]; a.out2 = p->out[2];
  a.bf16_mask = 0;
  for (int s = 0; s < p->n_seg; ++s) {
    CST_REQUIRE(p->out_dtype[s] == CST_F32 || p->out_dtype[s] == CST_BF16, "cst_dec_linear: out_dtype[%d]=%d unsupported", s, p->out_dtype[s]);
    if (p->out_dtype[s] == CST_BF16) a.bf16_mask |= 1 << s;
  }
  CST_REQUIRE(!(p->residual && a.bf16_mask), "cst_dec_linear: residual needs an f32 output");
  a.ldo0 = p->ldo[0]; a.ldo1 = p->ldo[1]; a.ldo2 = p->ldo[2];
  a.ss0 = p->step_stride[0]; a.ss1 = p->step_stride[1]; a.ss2 = p->step_stride[2];
  a.step = p->step; a.M = p->M; a.N = p->N; a.seg_n = p->N / p->n_seg; a.act = p->act;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->a_dtype == CST_F32 && p->w_dtype == CST_F32) return launch_dec_linear<float, float>(a, p->K, st);
  if (p->a_dtype == CST_F32 && p->w_dtype == CST_BF16) return launch_dec_linear<float, __nv_bfloat16>(a, p->K, st);
  if (p->a_dtype == CST_BF16 && p->w_dtype == CST_F32) return launch_dec_linear<__nv_bfloat16, float>(a, p->K, st);
  if (p->a_dtype == CST_BF16 && p->w_dtype == CST_BF16) return launch_dec_linear<__nv_bfloat16, __nv_bfloat16>(a, p->K, st);
  CST_REQUIRE(false, "cst_dec_linear: unsupported dtypes a=%d w=%d", p->a_dtype, p->w_dtype);
  return CST_OK;
}

extern "C" int cst_dec_attention(const float* q, long long ldq, const void* k, const void* v, int kv_dtype,
                                 long long kv_batch_stride, long long kv_row_stride, float* out, long long ldo, int B, int H,
                                 int n_keys, int n_keys_max, const int32_t* step, void* stream) {
  return cst_dec_attention_grouped(q, ldq, k, v, kv_dtype, kv_batch_stride, kv_row_stride, out, ldo, B, H, n_keys, n_keys_max, step, 1, stream);
}

extern "C" int cst_dec_attention_grouped(const float* q, long long ldq, const void* k, const void* v, int kv_dtype,
                                         long long kv_batch_stride, long long kv_row_stride, float* out, long long ldo, int B, int H,
                                         int n_keys, int n_keys_max, const int32_t* step, int kv_group, void* stream) {
  CST_REQUIRE(q && k && v && out && kv_group >= 1, "cst_dec_attention: null pointer / bad kv_group");
  CST_REQUIRE(B > 0 && H > 0 && n_keys_max > 0 && (step || (n_keys > 0 && n_keys <= n_keys_max)), "cst_dec_attention: bad sizes");
  CST_REQUIRE(kv_row_stride % 8 == 0 && kv_batch_stride % 8 == 0 && ldo % 2 == 0, "cst_dec_attention: strides must keep 16-byte rows");
  const size_t smem = (size_t)(64 + 8 * 64 + 16 + n_keys_max) * sizeof(float);
  CST_REQUIRE(smem <= 48 * 1024, "cst_dec_attention: n_keys_max=%d too large", n_keys_max);
  if (kv_dtype == CST_F32)
    CST_CHECK_CUDA(launch_dec(dec_attention_kernel<float>, dim3(B * H), dim3(DA_THREADS), smem, (cudaStream_t)stream, q, ldq,
                              (const float*)k, (const float*)v, kv_batch_stride, kv_row_stride, out, ldo, H, n_keys, n_keys_max, step, kv_group));
  else if (kv_dtype == CST_BF16)
    CST_CHECK_CUDA(launch_dec(dec_attention_kernel<__nv_bfloat16>, dim3(B * H), dim3(DA_THREADS), smem, (cudaStream_t)stream, q, ldq,
                              (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, kv_batch_stride, kv_row_stride, out, ldo, H, n_keys,
                              n_keys_max, step, kv_group));
  else
    CST_REQUIRE(false, "cst_dec_attention: kv_dtype %d unsupported", kv_dtype);
  return CST_OK;
}

extern "C" int cst_dec_select(const float* logits, int V, int B, int32_t* tokens, int ld_tok, float* pos_scores, int ld_ps,
                              int32_t* done, int32_t* out_len, int32_t* counters, int max_len, int min_len, int pad, int eos,
                              void* stream) {
  CST_REQUIRE(logits && tokens && pos_scores && done && out_len && counters, "cst_dec_select: null pointer");
  CST_REQUIRE(V > 0 && B > 0 && eos >= 0 && eos < V && max_len >= 0 && ld_tok >= max_len + 2 && ld_ps >= max_len + 1,
              "cst_dec_select: bad sizes (V=%d B=%d max_len=%d ld_tok=%d ld_ps=%d)", V, B, max_len, ld_tok, ld_ps);
  CST_CHECK_CUDA(launch_dec(dec_select_kernel, dim3(B), dim3(DS_THREADS), 0, (cudaStream_t)stream, logits, V, tokens, ld_tok, pos_scores,
                          ld_ps, done, out_len, counters, max_len, min_len, pad, eos));
  return CST_OK;
}
