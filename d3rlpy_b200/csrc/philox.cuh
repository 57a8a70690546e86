// Philox4x32-10 counter-based noise shared by noise_fill (optim.cu) and the fused update prologue (cql_fused.cu).
#pragma once
#include "common.cuh"

namespace d3b {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }


// Quad q of the noise stream of (seed, epoch): four floats out[4q .. 4q+3]; indices below n_normal are N(0,1)
// (Box-Muller on pairs), the rest U(-1,1).
__device__ __forceinline__ void noise_quad(float* __restrict__ out, long long q, long long n_normal, long long n,
                                           unsigned long long seed, uint32_t epoch) {
  uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32), epoch, 0x5eedu};
  philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  long long base = q << 2;
  float u0 = u01(c[0]), u1 = u01(c[1]), u2 = u01(c[2]), u3 = u01(c[3]);
  float ra = sqrtf(-2.f * logf(u0)), rb = sqrtf(-2.f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.f * u1, &s0, &c0);
  sincospif(2.f * u3, &s1, &c1);
  float nrm[4] = {ra * c0, ra * s0, rb * c1, rb * s1};
  float uni[4] = {2.f * u0 - 1.f, 2.f * u1 - 1.f, 2.f * u2 - 1.f, 2.f * u3 - 1.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    long long i = base + j;
    if (i < n) out[i] = (i < n_normal) ? nrm[j] : uni[j];
  }
}

}  // namespace d3b
