// Text (MT) input of the shared encoder: token embedding + sinusoidal positions, written in the shared stage's layout.
// Replaces: the text branch of S2T_W2V2_TransformerInterlinguaEncoder.forward
// (fairseq/models/chimera/w2v2_transformer_interlingua.py:212-217,230-236): x = sqrt(d) * text_embed_tokens(tokens) +
// embed_positions(padding_mask), positions 2, 3, ... on the valid tokens (t < len_b) and the zeroed padding row elsewhere
// (SinusoidalPositionalEmbedding, fairseq/modules/sinusoidal_positional_embedding.py:71-93; make_positions, utils.py:235-245).
#include "common.cuh"

namespace cst {

__global__ void __launch_bounds__(128) text_embed_kernel(const long long* __restrict__ tokens, const long long* __restrict__ lens,
                                                         const float* __restrict__ E, const float* __restrict__ pos, float scale,
                                                         float* __restrict__ x, int* __restrict__ valid, int T, int rows_per_seg,
                                                         int C, int V) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x / rows_per_seg, t = blockIdx.x - b * rows_per_seg;
  const long long len = lens[b];
  if (t == 0 && threadIdx.x == 0 && valid) valid[b] = (int)min((long long)T, max(0ll, len));
  float* dst = x + (size_t)blockIdx.x * C;
  if (t >= T) {                                                // filler rows of the segment
    for (int c = 4 * threadIdx.x; c < C; c += 4 * blockDim.x) store4(dst + c, make_float4(0.f, 0.f, 0.f, 0.f));
    return;
  }
  long long tok = tokens[(size_t)b * T + t];
  tok = min((long long)V - 1, max(0ll, tok));                  // the reference would raise on an out-of-range id
  const int p = t < len ? t + 2 : 1;                           // row 1 (padding_idx) of the table is zero
  for (int c = 4 * threadIdx.x; c < C; c += 4 * blockDim.x) {
    const float4 e = load4(E + (size_t)tok * C + c), q = load4(pos + (size_t)p * C + c);
    store4(dst + c, make_float4(fmaf(scale, e.x, q.x), fmaf(scale, e.y, q.y), fmaf(scale, e.z, q.z), fmaf(scale, e.w, q.w)));
  }
}

// x[b, t] += pos[t < valid[b] ? t + 2 : 1] for t < T: the sinusoidal positions the NON-memory base encoder adds to the sub-sampled
// audio frames (S2T_W2V2_TransformerEncoder.forward, fairseq/models/chimera/w2v2_transformer.py:353-357; same position rule as
// the text branch: make_positions over the padding mask, row 1 = padding_idx is zero).
__global__ void __launch_bounds__(128) add_positions_kernel(float* __restrict__ x, const int* __restrict__ valid,
                                                            const float* __restrict__ pos, int T, int rows_per_seg, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x / T, t = blockIdx.x - b * T;
  if (t >= valid[b]) return;                                   // padded frames get the zero row
  float* dst = x + ((size_t)b * rows_per_seg + t) * C;
  const float* q = pos + (size_t)(t + 2) * C;
  for (int c = 4 * threadIdx.x; c < C; c += 4 * blockDim.x) {
    const float4 a = load4(dst + c), p4 = load4(q + c);
    store4(dst + c, make_float4(a.x + p4.x, a.y + p4.y, a.z + p4.z, a.w + p4.w));
  }
}

// dE[tok, :] += scale * dx[b, t, :] for every token of the batch except the padding index (nn.Embedding(padding_idx) keeps that row's
// gradient at zero): the backward of the gather in text_embed_kernel.  fp32 atomics (a token may occur many times).
__global__ void __launch_bounds__(128) embed_bwd_kernel(const long long* __restrict__ tokens, const float* __restrict__ dx, float scale,
                                                        float* __restrict__ dE, int T, int rows_per_seg, int C, int V, int pad) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x / T, t = blockIdx.x - b * T;
  long long tok = tokens[(size_t)b * T + t];
  tok = min((long long)V - 1, max(0ll, tok));
  if (tok == pad) return;
  const float* src = dx + ((size_t)b * rows_per_seg + t) * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(dE + (size_t)tok * C + c, scale * src[c]);
}

}  // namespace cst

extern "C" int cst_embed_bwd(const int64_t* tokens, const float* dx, float scale, float* dE, int B, int T, int rows_per_seg, int C, int V,
                             int pad_idx, void* stream) {
  CST_REQUIRE(tokens && dx && dE && B > 0 && T > 0 && rows_per_seg >= T && C > 0 && V > 0, "cst_embed_bwd: bad args");
  CST_CHECK_CUDA(cst::launch_k(cst::embed_bwd_kernel, dim3(B * T), dim3(128), 0, (cudaStream_t)stream, (const long long*)tokens, dx, scale, dE, T,
                               rows_per_seg, C, V, pad_idx));
  return CST_OK;
}

extern "C" int cst_add_positions(float* x, const int32_t* valid, const float* pos_table, int B, int T, int rows_per_seg, int C, void* stream) {
  CST_REQUIRE(x && valid && pos_table && B > 0 && T > 0 && rows_per_seg >= T && C > 0 && C % 4 == 0, "cst_add_positions: bad args");
  CST_CHECK_CUDA(cst::launch_k(cst::add_positions_kernel, dim3(B * T), dim3(128), 0, (cudaStream_t)stream, x, (const int*)valid, pos_table, T,
                               rows_per_seg, C));
  return CST_OK;
}

extern "C" int cst_text_embed(const int64_t* tokens, const int64_t* lengths, const float* embed, const float* pos_table,
                              float scale, float* x, int32_t* valid, int B, int T, int rows_per_seg, int C, int V,
                              void* stream) {
  CST_REQUIRE(tokens && lengths && embed && pos_table && x, "cst_text_embed: null pointer");
  CST_REQUIRE(B > 0 && T > 0 && rows_per_seg >= T && C > 0 && C % 4 == 0 && V > 0, "cst_text_embed: bad sizes");
  CST_CHECK_CUDA(cst::launch_k(cst::text_embed_kernel, dim3(B * rows_per_seg), dim3(128), 0, (cudaStream_t)stream,
                               (const long long*)tokens, (const long long*)lengths, embed, pos_table, scale, x, valid, T,
                               rows_per_seg, C, V));
  return CST_OK;
}
