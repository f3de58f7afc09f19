// Philox4x32-10 counter RNG + Box-Muller, in registers.
// Noise is keyed on (seed; sample s, dim-OCTET g, GLOBAL image index, stream id), so a result does
// not depend on how images are sharded over GPUs.
// One Philox call (128 random bits) yields EIGHT normals, dims 8g..8g+7: word h (h = 0..3) gives the pair
// (8g+2h, 8g+2h+1) by Box-Muller from two 16-bit uniforms,
//     u1 = (hi16(word) + 0.5) / 65536     radius  sqrt(-2 ln u1)   (|n| <= 4.86)
//     u2 = (lo16(word) + 0.5) / 65536     angle   2 pi u2
// (round 1 used 24-bit uniforms, 4 normals per call: the generator was ~30 % of the fused evaluation's instructions).
// oracle/iwae_oracle.py:philox_normal restates exactly this mapping for the tests.
#pragma once
#include <stdint.h>

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// One uint32 -> two independent N(0,1) (libm-accurate version).
__device__ __forceinline__ float2 box_muller16(uint32_t r) {
  const float u1 = ((float)(r >> 16) + 0.5f) * 1.52587890625e-5f;      // 2^-16, exact
  const float u2 = ((float)(r & 0xFFFFu) + 0.5f) * 1.52587890625e-5f;
  const float rad = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return make_float2(rad * c, rad * s);
}

// Eight normals for dims 8g..8g+7 of (sample s, image img) in stream `stream_id`.
__device__ __forceinline__ void philox_normal8(uint64_t seed, uint32_t stream_id, uint32_t img, uint32_t s, uint32_t g,
                                               float (&o)[8]) {
  const uint4 r = philox4x32_10(make_uint4(s, g, img, stream_id), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const float2 p = box_muller16(w[h]);
    o[2 * h] = p.x;
    o[2 * h + 1] = p.y;
  }
}
