// Counter-based normal draws shared by the K-sample kernels (sampler.cu, tc_fc.cu): Philox4x32-10 + Box-Muller.
// Element i of GLOBAL sample index g uses counter (i/4, g, 0x52424E4E, 0) under key = the 64-bit seed; the four
// outputs of one Philox call are the normals of elements 4(i/4) .. 4(i/4)+3.  The numpy restatement used by the tests
// is oracle/oracle.py::philox_standard_normals.
#pragma once
#include <stdint.h>

namespace rbnn {

__device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                              uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

__device__ __forceinline__ float u01(uint32_t x) {  // ((x>>9)+0.5) * 2^-23, exact in fp32, in (0,1)
  return ((float)(x >> 9) + 0.5f) * 1.1920928955078125e-07f;
}

// Box-Muller on the special-function units.  The angle is taken in (-pi, pi) (same distribution as (0, 2 pi)), where
// __sinf / __cosf err by <= 2^-21.4 absolutely.  -log u: __logf (absolute error 2^-21.4 on [0.5, 2], 2 ulp below) loses
// RELATIVE accuracy as u -> 1, where the radius is small, so u > 0.9 takes the series of -log(1 - t), t = 1 - u <= 0.1
// (8 terms, truncation < 1e-9 relative).  A normal moves by < 1e-5 sigma against libm; the restatement test
// (oracle.philox_standard_normals) allows 2e-5.
__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float& z0, float& z1) {
  const float u = u01(xa);
  const float t = 1.f - u;
  float p = fmaf(t, 0.125f, 0.14285714285714285f);
  p = fmaf(t, p, 0.16666666666666666f);
  p = fmaf(t, p, 0.2f);
  p = fmaf(t, p, 0.25f);
  p = fmaf(t, p, 0.3333333333333333f);
  p = fmaf(t, p, 0.5f);
  p = fmaf(t, p, 1.f);
  const float nlog = u > 0.9f ? t * p : -__logf(u);
  const float r = sqrtf(2.f * nlog);
  const float a = fmaf(6.283185307179586f, u01(xb), -3.14159265358979f);        // cos t = -cos a, sin t = -sin a
  z0 = -r * __cosf(a);
  z1 = -r * __sinf(a);
}

// the four standard normals of quad q (elements 4q .. 4q+3) of global sample g
__device__ __forceinline__ void philox_normals4(uint32_t q, uint32_t g, uint32_t k0, uint32_t k1, float (&z)[4]) {
  uint32_t c0 = q, c1 = g, c2 = 0x52424E4Eu, c3 = 0u;
  philox4x32_10(c0, c1, c2, c3, k0, k1);
  box_muller(c0, c1, z[0], z[1]);
  box_muller(c2, c3, z[2], z[3]);
}

}  // namespace rbnn
