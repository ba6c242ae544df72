// Counter-based normal draws for the in-kernel noise mode of the sampler (Philox4x32-10, Salmon et
// al. SC'11; Box-Muller on the four outputs).  One call yields the four standard normals of columns
// [4*c4, 4*c4+4) of one latent row for one draw of one sampler call:
//
//   key     = seed (64 bit)
//   counter = (global row, column group c4 | draw << 16, call offset lo, call offset hi)
//
// so a value depends only on (seed, call offset, draw, GLOBAL row, column): a batch sharded over
// ranks (row_offset = first global row of the shard) or scored in chunks draws exactly the numbers
// the unsharded batch would, and no noise tensor exists in HBM (the injected-noise mode of
// aid_sample reads a [T-1, B, L] fp32 tensor: 1.6 GB per call at 65,536 rows).
// draw 0 = z_T (core/diffusion.py:190), draw 1 + i = the randn_like after the score call of loop
// iteration i (:236).  The stream is this library's own: parity with torch's generator is defined
// on the injected-noise mode; this mode is checked statistically and for shard invariance.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace aid {

struct PhiloxState {
  unsigned long long seed;
  unsigned long long offset;   // advanced by the caller once per sampler call
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// (0, 1]: 2^-24 * (top 24 bits + 1), so the logarithm below is finite
__device__ __forceinline__ float philox_unit(uint32_t x) { return ((float)(x >> 8) + 1.0f) * 5.9604644775390625e-8f; }

__device__ __forceinline__ float4 philox_normal4(const PhiloxState s, uint32_t draw, unsigned long long grow,
                                                 uint32_t c4) {
  uint4 ctr;
  ctr.x = (uint32_t)grow;
  ctr.y = (uint32_t)(grow >> 32) ^ (c4 << 8) ^ (draw << 20);
  ctr.z = (uint32_t)s.offset;
  ctr.w = (uint32_t)(s.offset >> 32);
  const uint4 r = philox4x32_10(ctr, make_uint2((uint32_t)s.seed, (uint32_t)(s.seed >> 32)));
  const float ra = sqrtf(-2.0f * __logf(philox_unit(r.x)));
  const float rb = sqrtf(-2.0f * __logf(philox_unit(r.z)));
  float sa, ca, sb, cb;
  __sincosf(6.283185307179586f * philox_unit(r.y), &sa, &ca);
  __sincosf(6.283185307179586f * philox_unit(r.w), &sb, &cb);
  return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}

}  // namespace aid
