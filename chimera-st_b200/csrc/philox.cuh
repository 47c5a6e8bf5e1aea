// Counter-based random bits for the dropout masks of the training step (Philox4x32-10, Salmon et al. SC'11 -- the generator
// behind torch's CUDA dropout; the reference draws its masks with F.dropout, fairseq/modules/fairseq_dropout.py:16-27).  A mask is
// never stored: forward and backward regenerate it from (seed, site, element index), so dropout costs no HBM for the tape.
//   key     = {seed lo, seed hi}            seed: one 64-bit value per training step, read from DEVICE memory (CUDA-graph replays)
//   counter = {group lo, group hi, site, 0} group = element index / 4; the element uses word (index % 4); site = which dropout
//   keep    = word >= floor(p * 2^32)       -> P(keep) = 1 - p;  kept values are scaled by 1 / (1 - p)
// tests/emu.py holds the same function in numpy; tests compare the two bit for bit.
#pragma once
#include <cstdint>

namespace cst {

struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(unsigned long long seed, unsigned long long group, uint32_t site) {
  uint32_t c0 = (uint32_t)group, c1 = (uint32_t)(group >> 32), c2 = site, c3 = 0u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
  const double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}

}  // namespace cst
