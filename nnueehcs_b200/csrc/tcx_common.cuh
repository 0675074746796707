// Device helpers shared by the fp32-parity split kernels (mlp_tcx.cu: one 64-row tile per CTA,
// hidden width 64 .. 512; mlp_tcx4.cu: four 64-row tile slots per CTA, hidden width 64 / 128):
// power-of-two row scales, the scaled fp16 hi/lo layer-0 operand, and the per-block epilogue math.
#pragma once
#include <cuda_fp16.h>

#include "philox.cuh"
#include "tc_params.cuh"
#include "tc_ptx.cuh"
#include "tc_epilogue.cuh"

namespace uq {
namespace tc {

constexpr int STAT_FLOATS = 4;                 // {1/C, max row norm, max |bias|, C}

// s = 2^e with  bound * s < 2^14  (and >= 2^13 unless clamped); inv = 1 / s.  A zero / denormal /
// non-finite bound gives s = 1.
__device__ __forceinline__ void pow2_scale(float bound, float& s, float& inv) {
  const int ex = (int)((__float_as_uint(bound) >> 23) & 0xffu);   // bound < 2^(ex - 126)
  int se = 267 - ex;                                              // field of 2^(140 - ex)
  se = se < 2 ? 2 : se > 252 ? 252 : se;
  if (ex == 0 || ex == 255) se = 127;
  s = __uint_as_float((uint32_t)se << 23);
  inv = __uint_as_float((uint32_t)(254 - se) << 23);
}

__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// Layer-0 A operand of one sample row: K0 fp16 values [x1 | x2 | x1 | 0 ...] of the row scaled by
// the power of two s_x that puts max |x_i| s_x in [2^13, 2^14); against the packed layer-0 weights
// [g1 | g1 | g2] this is x1 g1 + x2 g1 + x1 g2.  Also leaves 1 / s_x and |x|_2^2 of the row.
static __device__ __noinline__ void build_x_row_x(const TcParams& p, int64_t grow, bool to_stash,
                                                  uint32_t stash_row_addr,
                                                  uint32_t stash_piece_stride, uint32_t a_row,
                                                  int rx, float* xinfo_inv, float* xinfo_n2) {
  const int d = p.d_in;    // <= 21 (three segments inside one 64-column chunk)
  auto addr = [&](int col) -> uint32_t {
    const int piece = col >> 3;
    const uint32_t b = to_stash ? stash_row_addr + (uint32_t)piece * stash_piece_stride
                                : a_row + (uint32_t)((piece ^ rx) << 4);
    return b + (uint32_t)((col & 7) << 1);
  };
  auto st16 = [&](uint32_t a, __half v) {
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"(*reinterpret_cast<const uint16_t*>(&v))
                 : "memory");
  };
  float f[21];
  float amax = 0.f, n2 = 0.f;
#pragma unroll
  for (int i = 0; i < 21; ++i) {
    f[i] = (i < d) ? net_input2(p, grow, 0, i) : 0.f;
    amax = fmaxf(amax, fabsf(f[i]));
    n2 = fmaf(f[i], f[i], n2);
  }
  float s, inv;
  pow2_scale(amax, s, inv);
#pragma unroll
  for (int i = 0; i < 21; ++i) {
    if (i < d) {
      const float t = f[i] * s;
      const __half h1 = __float2half_rn(t);
      const __half h2 = __float2half_rn(t - __half2float(h1));
      st16(addr(i), h1);
      st16(addr(d + i), h2);
      st16(addr(2 * d + i), h1);
    }
  }
  for (int col = 3 * d; col < p.K0; ++col) st16(addr(col), __float2half_rn(0.f));
  *xinfo_inv = inv;
  *xinfo_n2 = n2;
}

// One NC-column block of one accumulator row:  v = acc * rs + bias, dropout, ReLU; then either
// the scaled fp16 hi/lo split into the two A pieces (and the row's running sum of squares) or the
// last-Linear dot product.
template <int H, int DOUT, int NC, bool RELU, bool DROP, bool LAST>
__device__ __forceinline__ void epi_block_x(const float (&acc)[NC], const float* bias_s,
                                            uint32_t keep, float rs, float s_out, uint32_t a1_dst,
                                            uint32_t a2_dst, int piece0, int rx,
                                            const float* __restrict__ wl_s,
                                            const float* __restrict__ wl_g, float (&dot)[DOUT],
                                            float& ss) {
  float v[NC];
#pragma unroll
  for (int j4 = 0; j4 < NC / 4; ++j4) {
    const float4 bv = reinterpret_cast<const float4*>(bias_s)[j4];
    v[j4 * 4 + 0] = fmaf(acc[j4 * 4 + 0], rs, bv.x);
    v[j4 * 4 + 1] = fmaf(acc[j4 * 4 + 1], rs, bv.y);
    v[j4 * 4 + 2] = fmaf(acc[j4 * 4 + 2], rs, bv.z);
    v[j4 * 4 + 3] = fmaf(acc[j4 * 4 + 3], rs, bv.w);
  }
  if (RELU) {
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (DROP) {
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] = ((keep >> j) & 1u) ? v[j] : 0.f;
  }
  if (!LAST) {
    uint32_t p1[NC / 2], p2[NC / 2];
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int j2 = 0; j2 < NC / 2; ++j2) {
      const float a = v[2 * j2], b = v[2 * j2 + 1];
      q0 = fmaf(a, a, q0);
      q1 = fmaf(b, b, q1);
      const float ta = a * s_out, tb = b * s_out;
      const __half2 h = __floats2half2_rn(ta, tb);       // .x (low half, lower address) = ta
      const float2 hf = __half22float2(h);
      const __half2 l = __floats2half2_rn(ta - hf.x, tb - hf.y);
      p1[j2] = h2_bits(h);
      p2[j2] = h2_bits(l);
    }
    ss += q0 + q1;
    epi_store<NC / 2>(p1, a1_dst, piece0, rx);
    epi_store<NC / 2>(p2, a2_dst, piece0, rx);
  } else {
    if (DOUT == 1) {
      float s0 = dot[0], s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int j4 = 0; j4 < NC / 4; ++j4) {
        const float4 wv = reinterpret_cast<const float4*>(wl_s)[j4];
        s0 = fmaf(v[j4 * 4 + 0], wv.x, s0);
        s1 = fmaf(v[j4 * 4 + 1], wv.y, s1);
        s2 = fmaf(v[j4 * 4 + 2], wv.z, s2);
        s3 = fmaf(v[j4 * 4 + 3], wv.w, s3);
      }
      dot[0] = (s0 + s1) + (s2 + s3);
    } else {
#pragma unroll
      for (int o = 0; o < DOUT; ++o) {
        float s = dot[o];
#pragma unroll
        for (int j4 = 0; j4 < NC / 4; ++j4) {
          const float4 wv = __ldg(reinterpret_cast<const float4*>(wl_g + o * H) + j4);
          s = fmaf(v[j4 * 4 + 0], wv.x, s);
          s = fmaf(v[j4 * 4 + 1], wv.y, s);
          s = fmaf(v[j4 * 4 + 2], wv.z, s);
          s = fmaf(v[j4 * 4 + 3], wv.w, s);
        }
        dot[o] = s;
      }
    }
  }
}

}  // namespace tc
}  // namespace uq
