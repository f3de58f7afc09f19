// Philox4x32-10 counter RNG + Box-Muller, in registers.
// Noise is keyed on (seed; sample s, dim-quad q, GLOBAL image index, stream id), so a result does
// not depend on how images are sharded over GPUs.
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

// Two uint32 -> two independent N(0,1).  u1 in (0,1], u2 in [0,1), both exact 24-bit fractions.
__device__ __forceinline__ float2 box_muller(uint32_t r0, uint32_t r1) {
  const float u1 = (float)((r0 >> 8) + 1u) * 5.9604644775390625e-8f;   // 2^-24
  const float u2 = (float)(r1 >> 8) * 5.9604644775390625e-8f;
  const float rad = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return make_float2(rad * c, rad * s);
}

// Four normals for dims 4q..4q+3 of (sample s, image img) in stream `stream_id`.
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint32_t stream_id, uint32_t img,
                                                 uint32_t s, uint32_t q) {
  const uint4 r = philox4x32_10(make_uint4(s, q, img, stream_id),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float2 a = box_muller(r.x, r.y);
  const float2 b = box_muller(r.z, r.w);
  return make_float4(a.x, a.y, b.x, b.y);
}
