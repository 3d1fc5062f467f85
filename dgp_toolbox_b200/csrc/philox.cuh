// Philox-4x32-10 counter RNG + Box-Muller. Draw layout (must match oracle/dgp_oracle.py:philox_normal bit for bit in
// the integer part): key = (seed_lo, seed_hi), counter = (n_global, s, d, layer) -> words r0..r3;
// u1 from (r0,r1), u2 from (r2,r3), z = sqrt(-2 ln u1) cos(2 pi u2).
#pragma once
#include <cstdint>

namespace dgp {

__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)c0 * M0, p1 = (uint64_t)c2 * M1;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ inline double philox_normal(uint64_t seed, uint32_t layer, uint32_t s, uint32_t n_global, uint32_t d) {
  uint32_t r[4];
  philox4x32_10(n_global, s, d, layer, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  double u1 = ((double)(((uint64_t)(r[0] >> 5) << 26) + (uint64_t)(r[1] >> 6)) + 0.5) * (1.0 / 9007199254740992.0);
  double u2 = ((double)(((uint64_t)(r[2] >> 5) << 26) + (uint64_t)(r[3] >> 6)) + 0.5) * (1.0 / 9007199254740992.0);
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

}  // namespace dgp
