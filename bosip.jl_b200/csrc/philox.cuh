// Philox4x32-10 keyed by the GLOBAL candidate index (bit-identical to bosip.jl_b200/synth.py:philox4x32), so that
// candidate i is the same point whichever GPU generates it and results do not depend on the GPU count.
#pragma once
#include <stdint.h>

namespace bosip {

__host__ __device__ inline void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += W0; k1 += W1;
  }
}

// coordinate k of candidate `idx`: 53 random bits -> [0,1)
__host__ __device__ inline double philox_unit(uint64_t seed, uint64_t idx, int k) {
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)(k >> 1), 0u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint64_t lo = c[2 * (k & 1)], hi = c[2 * (k & 1) + 1];
  const uint64_t bits = (hi << 21) | (lo >> 11);
  return (double)bits * (1.0 / 9007199254740992.0);
}

}  // namespace bosip
