// Threefry-2x32 (20 rounds), the block function of JAX's default PRNG
// (jax_default_prng_impl = "threefry2x32", photon_weave/__init__.py:16).
// Shared by the device sampling kernel and the host key-splitting entry point.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PW_HD __host__ __device__ __forceinline__
#else
#define PW_HD inline
#endif

namespace pw {

PW_HD uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

PW_HD void threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* o0,
                        uint32_t* o1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  const int rot[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  x0 += ks[0];
  x1 += ks[1];
  for (int i = 0; i < 5; ++i) {
    for (int j = 0; j < 4; ++j) {
      x0 += x1;
      x1 = rotl32(x1, rot[i & 1][j]) ^ x0;
    }
    x0 += ks[(i + 1) % 3];
    x1 += ks[(i + 2) % 3] + (uint32_t)(i + 1);
  }
  *o0 = x0;
  *o1 = x1;
}

}  // namespace pw
