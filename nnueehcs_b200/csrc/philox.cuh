// Counter-based dropout masks shared by the fp32 and the bf16 kernels.
//
// The reference draws Bernoulli(1-p) keep-masks from torch's generator inside nn.Dropout
// (nnueehcs/models.py:156-158 with MCDropoutModel.eval, :165-169).  Here every mask bit is a
// pure function of (seed, offset, global pass id, dropout layer, sample, feature), so a K-axis
// shard on any GPU reproduces exactly the bits the single-GPU run would draw, and
// uq_philox_keep_masks() can export the very same bits for an exact replay through the oracle.
//
// Keep decisions are drawn 32 features at a time (dropout_keep32):  keep <=> r < thr16  for a
// uniform 16-bit r per feature, thr16 = round((1-p) * 65536) (keep probability quantised to
// 2^-16), evaluated lazily so that ~2 Philox4x32-10 calls serve 32 features instead of 4:
//   stage 1  one call = four 32-bit words = the top four bits of r for all 32 features at once
//            (bit planes, MSB first); a feature is decided as soon as its bit differs from the
//            threshold's bit: two logic ops per plane for 32 features;
//   stage 2  the features whose top nibble equals the threshold's (1 in 16, ~2 per group) draw
//            their low 12 bits from the next call, 12 bits each, 10 features per call (further
//            calls in the rare case of more than 10).
// Every feature uses bits of its own, so the decisions are independent and exactly
// Bernoulli(thr16 / 65536); the generator cost drops from ~10 to ~5 instructions per feature.
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

// 32-bit keep mask (bit i <=> feature 32 * group32 + i is kept) of one (pass, layer, sample).
// Counter: (sample, layer << 24 | group32 << 4 | call, pass, offset).
__device__ __forceinline__ uint32_t dropout_keep32(const PhiloxKey& key, uint32_t thr16, uint32_t pass,
                                                   uint32_t layer, uint32_t sample,
                                                   uint32_t group32) {
  if (thr16 >= 65536u) return 0xffffffffu;
  const uint32_t c1 = (layer << 24) | (group32 << 4);
  const uint4 r = philox4x32_10(sample, c1, pass, key.off, key.k0, key.k1);
  const uint32_t planes[4] = {r.x, r.y, r.z, r.w};
  uint32_t undecided = 0xffffffffu, keep = 0u;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t t = ((thr16 >> (15 - i)) & 1u) ? 0xffffffffu : 0u;   // threshold bit, replicated
    keep |= undecided & ~planes[i] & t;       // r bit 0 < threshold bit 1: r < thr, keep
    undecided &= ~(planes[i] ^ t);            // still undecided where the bits are equal
  }
  const uint32_t tlow = thr16 & 0xfffu;
  uint32_t call = 1u;
  while (undecided) {
    uint4 s = philox4x32_10(sample, c1 | call, pass, key.off, key.k0, key.k1);
#pragma unroll 1
    for (int j = 0; j < 10 && undecided; ++j) {
      const int e = __ffs(undecided) - 1;
      undecided &= undecided - 1u;
      if ((s.x & 0xfffu) < tlow) keep |= 1u << e;
      s.x = __funnelshift_r(s.x, s.y, 12);
      s.y = __funnelshift_r(s.y, s.z, 12);
      s.z = __funnelshift_r(s.z, s.w, 12);
      s.w >>= 12;
    }
    ++call;    // at most 4 calls (32 undecided features)
  }
  return keep;
}

// single-feature query (recomputes the group of 32: cross-check paths only)
__device__ __forceinline__ bool dropout_keep1(const PhiloxKey& key, uint32_t thr16, uint32_t pass,
                                              uint32_t layer, uint32_t sample, uint32_t feature) {
  return (dropout_keep32(key, thr16, pass, layer, sample, feature >> 5) >> (feature & 31u)) & 1u;
}

}  // namespace uq
