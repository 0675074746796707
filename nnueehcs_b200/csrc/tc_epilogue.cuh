// Epilogue building blocks shared by the CTA-pair kernels (mlp_tc2.cu: 128 rows per CTA,
// mlp_tc3.cu: 64 rows per CTA): network input fetch, dropout keep bits, the per-block
// bias / dropout / ReLU / bf16-rounding math and the swizzled A-operand store.
#pragma once
#include "philox.cuh"
#include "tc_params.cuh"
#include "tc_ptx.cuh"

namespace uq {
namespace tc {

// Network input i of a sample row.  Delta-UQ never reaches this with anchors: W0 [x - a; a] + b0
// = W0a x + (b0 + (W0b - W0a) a), so the kernels see plain x and a per-anchor layer-0 bias.
__device__ __forceinline__ float net_input2(const TcParams& p, int64_t row, int member_global,
                                            int i) {
  (void)member_global;
  if (row >= p.n) return 0.f;
  return __ldg(p.x + row * p.d_x + i);
}

// Layer-0 A operand of one sample row: K0 bf16 values [x_hi | x_lo | x_hi | 0 ...] (split_s
// segments of d_in network inputs each), written as 2-byte shared stores either into the x stash
// (piece-major: piece * stash_piece_stride) or straight into the swizzled chunk-0 row.  All
// global loads of the row are issued before the first is used -- the obvious loop over output
// columns serialises one L1/L2 round trip per element (~4500 cycles per call, measured).
static __device__ __noinline__ void build_x_row(const TcParams& p, int64_t grow, int member_global,
                                                bool to_stash, uint32_t stash_row_addr,
                                                uint32_t stash_piece_stride, uint32_t a_row,
                                                int rx) {
  const int d = p.d_in;
  auto addr = [&](int col) -> uint32_t {
    const int piece = col >> 3;
    const uint32_t b = to_stash ? stash_row_addr + (uint32_t)piece * stash_piece_stride
                                : a_row + (uint32_t)((piece ^ rx) << 4);
    return b + (uint32_t)((col & 7) << 1);
  };
  auto st16 = [&](uint32_t a, __nv_bfloat16 v) {
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"(*reinterpret_cast<const uint16_t*>(&v))
                 : "memory");
  };
#pragma unroll 1
  for (int i0 = 0; i0 < d; i0 += 32) {
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = (i0 + i < d) ? net_input2(p, grow, member_global, i0 + i) : 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (i0 + i < d) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(f[i]);
        st16(addr(i0 + i), hi);
        if (p.split_s > 1) st16(addr(d + i0 + i), __float2bfloat16_rn(f[i] - __bfloat162float(hi)));
        if (p.split_s > 2) st16(addr(2 * d + i0 + i), hi);
      }
    }
  }
  for (int col = p.split_s * d; col < p.K0; ++col) st16(addr(col), __float2bfloat16_rn(0.f));
}

// keep-mask bits of 32 consecutive features of one row
__device__ __forceinline__ uint32_t keep_bits32(const TcParams& p, int drop, int kg, int drop_ord,
                                                int64_t grow, int col0, const uint8_t* mask_layer,
                                                int H) {
  uint32_t keep = 0;
  if (drop == 2) {
    keep = dropout_keep32(p.key, p.thr16, (uint32_t)kg, (uint32_t)drop_ord,
                          (uint32_t)(grow + p.row_base), (uint32_t)(col0 >> 5));
  } else if (grow < p.n) {
    const uint4* mrow = reinterpret_cast<const uint4*>(
        mask_layer + ((size_t)kg * (size_t)p.n + (size_t)grow) * H + col0);
    const uint4 m0 = __ldg(mrow), m1 = __ldg(mrow + 1);
    const uint32_t w[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      keep |= ((w[j] & 0xFFu) ? 1u : 0u) << (4 * j);
      keep |= ((w[j] & 0xFF00u) ? 1u : 0u) << (4 * j + 1);
      keep |= ((w[j] & 0xFF0000u) ? 1u : 0u) << (4 * j + 2);
      keep |= ((w[j] & 0xFF000000u) ? 1u : 0u) << (4 * j + 3);
    }
  }
  return keep;
}

// One NC-column block of one accumulator row: + bias, dropout, then either ReLU + bf16 rounding
// into NC/2 packed words (cvt.rn[.relu].bf16x2) or the last-Linear dot product.
// BIAS_IN_ACC (the bias-in-the-MMA variants): the accumulator already holds acc + bias, bv is not
// read, and the activations are stored WITH their dropout's 1/(1-p) -- in_scale is then the scale of
// THIS layer's dropout, applied to the kept elements (v * (keep ? s : 0): FSEL + FMUL, where the
// epilogue-bias variant spends FFMA + FSEL), not the scale owed by the previous layer.
template <int H, int DOUT, int NC, bool RELU, bool DROP, bool LAST, bool BIAS_IN_ACC = false>
__device__ __forceinline__ void epi_math(const uint32_t (&acc)[NC], const float4 (&bv)[NC / 4],
                                         uint32_t keep, float in_scale, uint32_t* packed,
                                         const float* __restrict__ wl_s,
                                         const float* __restrict__ wl_g, float (&dot)[DOUT]) {
  float v[NC];
  if (BIAS_IN_ACC) {
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] = __uint_as_float(acc[j]);
  }
#pragma unroll
  for (int j4 = 0; j4 < (BIAS_IN_ACC ? 0 : NC / 4); ++j4) {
    // in_scale = 1 / (1 - p) when the previous layer's output went through a dropout (which only
    // zeroes; the rescale rides on this FMA for free), else 1.0 -- fma(acc, 1, b) == acc + b exactly
    v[j4 * 4 + 0] = fmaf(__uint_as_float(acc[j4 * 4 + 0]), in_scale, bv[j4].x);
    v[j4 * 4 + 1] = fmaf(__uint_as_float(acc[j4 * 4 + 1]), in_scale, bv[j4].y);
    v[j4 * 4 + 2] = fmaf(__uint_as_float(acc[j4 * 4 + 2]), in_scale, bv[j4].z);
    v[j4 * 4 + 3] = fmaf(__uint_as_float(acc[j4 * 4 + 3]), in_scale, bv[j4].w);
  }
  if (DROP) {
#pragma unroll
    for (int j = 0; j < NC; ++j)
      v[j] = BIAS_IN_ACC ? v[j] * (((keep >> j) & 1u) ? in_scale : 0.f)
                         : (((keep >> j) & 1u) ? v[j] : 0.f);
  }
  if (!LAST) {
#pragma unroll
    for (int j2 = 0; j2 < NC / 2; ++j2)
      packed[j2] = RELU ? cvt_relu_bf16x2(v[2 * j2], v[2 * j2 + 1]) : cvt_bf16x2(v[2 * j2], v[2 * j2 + 1]);
  } else {
    if (RELU) {
#pragma unroll
      for (int j = 0; j < NC; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (DOUT == 1) {
      // four independent FMA chains (a single one is latency-bound: 4 cycles x 64 per chunk)
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

// NW packed words (NW/4 16-byte pieces starting at piece0) -> this row of a swizzled A chunk
template <int NW>
__device__ __forceinline__ void epi_store(const uint32_t* packed, uint32_t a_dst, int piece0,
                                          int rx) {
#pragma unroll
  for (int pc = 0; pc < NW / 4; ++pc) {
    st_shared_v4(a_dst + (uint32_t)(((piece0 + pc) ^ rx) << 4), packed[pc * 4 + 0],
                 packed[pc * 4 + 1], packed[pc * 4 + 2], packed[pc * 4 + 3]);
  }
}

template <int H, int DOUT, int NC, bool RELU, bool DROP, bool LAST, bool BIAS_IN_ACC = false>
__device__ __forceinline__ void epi_block2(const uint32_t (&acc)[NC], const float4 (&bv)[NC / 4],
                                           uint32_t keep, float in_scale, uint32_t a_dst,
                                           int piece0, int rx, const float* __restrict__ wl_s,
                                           const float* __restrict__ wl_g, float (&dot)[DOUT]) {
  uint32_t packed[NC / 2];
  epi_math<H, DOUT, NC, RELU, DROP, LAST, BIAS_IN_ACC>(acc, bv, keep, in_scale, packed, wl_s, wl_g,
                                                       dot);
  if (!LAST) epi_store<NC / 2>(packed, a_dst, piece0, rx);
}

// One member's output y of one row folded into the row's running state over the member axis:
// Welford (mean, M2), or -- PAGER -- (mean, max |y - Y_k|) in the same two registers.
__device__ __forceinline__ void member_fold(const TcParams& p, int kg, int o, float y, float inv_n,
                                            float& mean, float& m2) {
  const float dlt = y - mean;
  mean += dlt * inv_n;
  if (p.targets) m2 = fmaxf(m2, fabsf(y - __ldg(p.targets + (size_t)kg * p.d_out + o)));
  else m2 = fmaf(dlt, y - mean, m2);
}

// 1 / (1 - p) of a Dropout that sits between the last hidden activation and the final Linear (the
// reference's MC-dropout builder never puts one there, model_builder.py:257-262, but a hand-built
// nn.Sequential may): the final dot product sees masked, unscaled activations
__device__ __forceinline__ float final_dropout_scale(const TcParams& p) {
  return (((p.dropout_mask >> (p.L_mma - 1)) & 1u) && p.drop_mode) ? p.drop_scale : 1.f;
}

// what goes to out1 for one output element
__device__ __forceinline__ float second_output(const TcParams& p, float m2, float n, int64_t idx) {
  if (p.targets) return p.score_floor ? fmaxf(m2, __ldg(p.score_floor + idx)) : m2;
  return (p.output == UQ_OUT_MOMENTS) ? m2 : sqrtf(m2 / (n - 1.f));
}

template <int THREADS>
__device__ __forceinline__ void epi_bar_sync_n() {
  asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
}

}  // namespace tc
}  // namespace uq
