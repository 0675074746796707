// Counter-based dropout masks shared by the fp32 and the bf16 kernels.
//
// The reference draws Bernoulli(1-p) keep-masks from torch's generator inside nn.Dropout
// (nnueehcs/models.py:156-158 with MCDropoutModel.eval, :165-169).  Here every mask bit is a
// pure function of (seed, offset, global pass id, dropout layer, sample, feature), so a K-axis
// shard on any GPU reproduces exactly the bits the single-GPU run would draw, and
// uq_philox_keep_masks() can export the very same bits for an exact replay through the oracle.
//
// One Philox4x32-10 call yields 128 bits = eight 16-bit lanes = the keep decisions of eight
// consecutive features [8g, 8g+8) of one sample:  keep <=> lane < thr16,
// thr16 = round((1-p) * 65536)  (keep probability quantised to 2^-16).
#pragma once
#include <stdint.h>

namespace uq {

struct PhiloxKey {
  uint32_t k0, k1;   // seed lo / hi
  uint32_t off;      // philox_offset (low 32 bits) -> counter word 3
};

__host__ __device__ inline uint32_t dropout_thr16(float p) {
  float keep = 1.0f - p;
  float t = keep * 65536.0f + 0.5f;
  if (t < 0.f) t = 0.f;
  if (t > 65536.f) t = 65536.f;
  return (uint32_t)t;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return make_uint4(c0, c1, c2, c3);
}

// 128 random bits for (pass, layer, sample, feature-group-of-8)
__device__ __forceinline__ uint4 dropout_bits(const PhiloxKey& key, uint32_t pass, uint32_t layer,
                                              uint32_t sample, uint32_t group8) {
  return philox4x32_10(sample, (layer << 24) | group8, pass, key.off, key.k0, key.k1);
}

// 8-bit keep mask (bit i <=> feature 8*group8 + i is kept)
__device__ __forceinline__ uint32_t dropout_keep8(const PhiloxKey& key, uint32_t thr16, uint32_t pass,
                                                  uint32_t layer, uint32_t sample, uint32_t group8) {
  uint4 r = dropout_bits(key, pass, layer, sample, group8);
  uint32_t m = 0;
  m |= ((r.x & 0xFFFFu) < thr16) ? 1u : 0u;
  m |= ((r.x >> 16) < thr16) ? 2u : 0u;
  m |= ((r.y & 0xFFFFu) < thr16) ? 4u : 0u;
  m |= ((r.y >> 16) < thr16) ? 8u : 0u;
  m |= ((r.z & 0xFFFFu) < thr16) ? 16u : 0u;
  m |= ((r.z >> 16) < thr16) ? 32u : 0u;
  m |= ((r.w & 0xFFFFu) < thr16) ? 64u : 0u;
  m |= ((r.w >> 16) < thr16) ? 128u : 0u;
  return m;
}

// single-feature query (fp32 path epilogue; recomputes the group, parity mode only)
__device__ __forceinline__ bool dropout_keep1(const PhiloxKey& key, uint32_t thr16, uint32_t pass,
                                              uint32_t layer, uint32_t sample, uint32_t feature) {
  return (dropout_keep8(key, thr16, pass, layer, sample, feature >> 3) >> (feature & 7u)) & 1u;
}

}  // namespace uq
