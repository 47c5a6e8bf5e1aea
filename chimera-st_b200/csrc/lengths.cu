// a4: frame-level padding rule and lengths, evaluated on the device (no host sync).
// Reference behaviour: data_utils.py:491-495 + wav2vec2.py:543-548 + w2v2_transformer.py:333
// + s2t_transformer.py:63-67; closed forms derived in chimera-st_b200/lengths.py.
#include "common.cuh"

namespace cst {
__global__ void frame_lengths_kernel(const int64_t* __restrict__ src_len, int L, int n_frames,
                                     int32_t* w2v_valid, int32_t* sub_valid, int64_t* w2v_len64,
                                     uint8_t* frame_mask) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const long long len = src_len[b];
  const int r = L / n_frames;                    // samples per frame after trimming L % T'
  long long v = (len + r - 1) / r;               // frame t is valid iff t*r < len
  if (v > n_frames) v = n_frames;
  if (v < 0) v = 0;
  if (threadIdx.x == 0) {
    if (w2v_valid) w2v_valid[b] = (int32_t)v;
    if (w2v_len64) w2v_len64[b] = v;
    if (sub_valid) { long long s = (v + 1) / 2; s = (s + 1) / 2; sub_valid[b] = (int32_t)s; }
  }
  if (frame_mask)
    for (int t = threadIdx.x; t < n_frames; t += blockDim.x)
      frame_mask[(size_t)b * n_frames + t] = (t >= v) ? 1 : 0;
}
}  // namespace cst

extern "C" int cst_frame_lengths(const int64_t* src_len, int B, int L, int n_frames,
                                 int32_t* w2v_valid, int32_t* sub_valid, int64_t* w2v_len64,
                                 uint8_t* frame_mask, void* stream) {
  CST_REQUIRE(src_len && B > 0 && L > 0 && n_frames > 0 && n_frames <= L,
              "cst_frame_lengths: bad args B=%d L=%d n_frames=%d", B, L, n_frames);
  CST_CHECK_CUDA(cst::launch_k(cst::frame_lengths_kernel, dim3(B), dim3(128), 0, (cudaStream_t)stream, src_len, L, n_frames, w2v_valid,
                               sub_valid, w2v_len64, frame_mask));
  return CST_OK;
}
