// fp32 GEMM (FFMA, true-fp32 accumulate) -- the <=1e-5 parity mode of the path.  TF32 tensor cores are
// ~1e-3 and cannot meet that bar, so this mode runs on the CUDA cores by design.
// 128 x BN x 16 tiles, 256 threads, register-prefetched double buffering, shared epilogue.
#include "gemm_common.cuh"

namespace cst {

template <int BN>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const GemmDev p) {
  constexpr int BM = 128, BK = 16, TN = 8;
  constexpr int TX = BN / TN, TY = 256 / TX, TM = BM / TY;
  constexpr int B_LD = (BN * BK / 4) / 256;            // float4 loads per thread for the W tile (2 or 1)
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];
  pdl_launch_dependents();
  pdl_wait();

  const int z = blockIdx.z, zo = z / p.nb_inner, zi = z - zo * p.nb_inner;
  const float* A = (const float*)p.A + zo * p.a_bs_outer + zi * p.a_bs_inner;
  const float* W = (const float*)p.W + zi * p.w_bs_inner;
  const float* bias = p.bias ? p.bias + (long long)zi * p.bias_bs_inner : nullptr;
  const long long c_off = zo * p.c_bs_outer + zi * p.c_bs_inner;
  const long long r_off = zo * p.r_bs_outer + zi * p.r_bs_inner;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;

  const int lrow = tid >> 2, lk4 = (tid & 3) * 4;      // loader coordinates
  float4 ra[2], rb[B_LD];

  auto gload = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = m0 + lrow + h * 64;
      const long long e = (long long)m * p.lda + k0 + lk4;
      ra[h] = (m < p.M && e + 3 < p.a_limit) ? load4(A + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int h = 0; h < B_LD; ++h) {
      const int n = n0 + lrow + h * 64;
      rb[h] = (n < p.N) ? load4(W + (long long)n * p.K + k0 + lk4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      As[buf][lk4 + 0][lrow + h * 64] = ra[h].x; As[buf][lk4 + 1][lrow + h * 64] = ra[h].y;
      As[buf][lk4 + 2][lrow + h * 64] = ra[h].z; As[buf][lk4 + 3][lrow + h * 64] = ra[h].w;
    }
#pragma unroll
    for (int h = 0; h < B_LD; ++h) {
      Bs[buf][lk4 + 0][lrow + h * 64] = rb[h].x; Bs[buf][lk4 + 1][lrow + h * 64] = rb[h].y;
      Bs[buf][lk4 + 2][lrow + h * 64] = rb[h].z; Bs[buf][lk4 + 3][lrow + h * 64] = rb[h].w;
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = p.K / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + i]);
        a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN + j]);
        b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    const RowInfo ri = row_info(p, m, zo);
    epilogue8(p, ri, n0 + tx * TN, bias, c_off, r_off, acc[i]);
  }
}

int launch_gemm_f32(const GemmDev& p, int nz, cudaStream_t st) {
  if (p.N <= 64) {
    dim3 grid(cdiv(p.M, 128), cdiv(p.N, 64), nz);
    CST_CHECK_CUDA(launch_k(gemm_f32_kernel<64>, grid, dim3(256), 0, st, p));
  } else {
    dim3 grid(cdiv(p.M, 128), cdiv(p.N, 128), nz);
    CST_CHECK_CUDA(launch_k(gemm_f32_kernel<128>, grid, dim3(256), 0, st, p));
  }
  return CST_OK;
}

}  // namespace cst

extern "C" int cst_gemm(const cst_gemm_params* hp, void* stream) {
  using namespace cst;
  CST_REQUIRE(hp && hp->A && hp->W && hp->C, "cst_gemm: null pointer");
  CST_REQUIRE(hp->M > 0 && hp->N > 0 && hp->K > 0, "cst_gemm: bad shape M=%d N=%d K=%d", hp->M, hp->N, hp->K);
  CST_REQUIRE(hp->N % 8 == 0, "cst_gemm: N=%d must be a multiple of 8", hp->N);
  CST_REQUIRE(hp->K % 64 == 0, "cst_gemm: K=%d must be a multiple of 64", hp->K);
  CST_REQUIRE(hp->lda % 8 == 0 && hp->ldc % 8 == 0, "cst_gemm: lda/ldc must be multiples of 8");
  CST_REQUIRE(hp->residual == nullptr || hp->ldr % 4 == 0, "cst_gemm: ldr must be a multiple of 4");
  CST_REQUIRE(hp->act >= CST_ACT_NONE && hp->act <= CST_ACT_GLU, "cst_gemm: bad act %d", hp->act);
  CST_REQUIRE(hp->c_dtype == CST_F32 || hp->c_dtype == CST_BF16 || hp->c_dtype == CST_F16, "cst_gemm: bad c_dtype");
  CST_REQUIRE(hp->ab_dtype != CST_F32 || hp->c_dtype != CST_F16, "cst_gemm: the fp32 kernel has no fp16 output");
  CST_REQUIRE(hp->rows_per_seg > 0 && hp->seg_rows_valid > 0, "cst_gemm: rows_per_seg/seg_rows_valid must be > 0");
  CST_REQUIRE(hp->nb_outer >= 1 && hp->nb_inner >= 1, "cst_gemm: nb_outer/nb_inner must be >= 1");
  CST_REQUIRE(hp->a_rows > 0, "cst_gemm: a_rows must be > 0");
  GemmDev p;
  p.A = hp->A; p.W = hp->W; p.bias = hp->bias; p.residual = hp->residual; p.C = hp->C;
  p.c_dtype = hp->c_dtype; p.M = hp->M; p.N = hp->N; p.K = hp->K;
  p.lda = hp->lda; p.ldc = hp->ldc; p.ldr = hp->ldr; p.a_limit = hp->a_rows * hp->lda;
  p.act = hp->act; p.alpha = hp->alpha; p.nb_inner = hp->nb_inner;
  p.a_bs_outer = hp->a_bs_outer; p.a_bs_inner = hp->a_bs_inner; p.w_bs_inner = hp->w_bs_inner;
  p.c_bs_outer = hp->c_bs_outer; p.c_bs_inner = hp->c_bs_inner;
  p.r_bs_outer = hp->r_bs_outer; p.r_bs_inner = hp->r_bs_inner; p.bias_bs_inner = hp->bias_bs_inner;
  p.rows_per_seg = hp->rows_per_seg; p.seg_rows_valid = hp->seg_rows_valid;
  p.out_rows_per_seg = hp->out_rows_per_seg; p.out_row_off = hp->out_row_off;
  p.seg_len = hp->seg_len; p.segs_per_outer = hp->segs_per_outer;
  p.ln_in_stats = reinterpret_cast<const float2*>(hp->ln_in_stats); p.ln_colsum = hp->ln_colsum; p.ln_in_slots = hp->ln_in_slots;
  p.res_stats = reinterpret_cast<const float2*>(hp->res_stats); p.res_slots = hp->res_slots;
  p.res_gamma = hp->res_gamma; p.res_beta = hp->res_beta;
  p.C2 = reinterpret_cast<__nv_bfloat16*>(hp->C2); p.ldc2 = hp->ldc2;
  p.out_stats = reinterpret_cast<float2*>(hp->out_stats);
  p.ln_inv_dim = hp->ln_dim > 0 ? 1.0f / (float)hp->ln_dim : 0.f;
  p.exact_act = hp->exact_act;
  p.acc_scale = hp->acc_scale == 0.f ? 1.0f : hp->acc_scale;
  CST_REQUIRE(p.acc_scale == 1.0f || hp->ab_dtype != CST_F32, "cst_gemm: acc_scale is a tensor-core path parameter");
  const bool ln_fused = hp->ln_in_stats || hp->res_stats || hp->C2 || hp->out_stats;
  if (ln_fused) {
    CST_REQUIRE(hp->ab_dtype == CST_BF16 || hp->ab_dtype == CST_F16, "cst_gemm: fused LayerNorm needs the tensor-core path (16-bit operands)");
    CST_REQUIRE(hp->ln_dim > 0, "cst_gemm: fused LayerNorm needs ln_dim");
    CST_REQUIRE(!hp->ln_in_stats || (hp->ln_colsum && hp->ln_in_slots > 0 && hp->ln_in_slots <= 8 && !hp->residual),
                "cst_gemm: ln_in_stats needs ln_colsum, 1..8 slots and no residual");
    CST_REQUIRE(!hp->res_stats || (hp->res_gamma && hp->res_beta && hp->res_slots >= 0 && hp->res_slots <= 8 && hp->residual),
                "cst_gemm: res_stats needs res_gamma / res_beta, 0..8 slots and a residual");
    CST_REQUIRE(!(hp->C2 || hp->out_stats || hp->res_stats) || (hp->residual && hp->c_dtype == CST_F32),
                "cst_gemm: C2 / out_stats / res_stats need an fp32 C with a residual");
    CST_REQUIRE(!hp->C2 || (hp->c2_dtype == CST_BF16 && hp->ldc2 % 8 == 0), "cst_gemm: C2 must be bf16 with ldc2 %% 8 == 0");
  }
  const int nz = hp->nb_outer * hp->nb_inner;
  cudaStream_t st = (cudaStream_t)stream;
  if (hp->ab_dtype == CST_F32) return launch_gemm_f32(p, nz, st);
  if (hp->ab_dtype == CST_BF16 || hp->ab_dtype == CST_F16) return launch_gemm_tc(*hp, p, nz, st);
  CST_REQUIRE(false, "cst_gemm: bad ab_dtype %d", hp->ab_dtype);
  return CST_ERR_ARG;
}
