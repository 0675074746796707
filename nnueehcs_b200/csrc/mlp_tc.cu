// bf16 tcgen05 path of the UQ forward: the throughput mode.
//
// One persistent, warp-specialised kernel runs the whole  members x layers  MLP stack of a
// 128-sample tile without the activations ever leaving the SM:
//
//   warp 0      producer   streams the packed weight image (already in UMMA K-major SWIZZLE_128B
//                          smem layout) from L2 into a ring of smem stages with 1-D TMA bulk
//                          copies (cp.async.bulk -> UBLKCP) signalling mbarriers;
//   warp 1      MMA        one lane issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) with
//                          A = activations [128 x 64]-chunks in smem, B = weight stage,
//                          D = [128 x H] fp32 accumulator in TMEM; tcgen05.commit releases
//                          stages and publishes finished layers;
//   warps 2..9  epilogue   tcgen05.ld the accumulator (double-buffered in registers), add the
//                          (BN-folded) bias, ReLU, apply the Philox / injected dropout mask, round
//                          to bf16 and write the next layer's A operand back into the same smem
//                          chunks (in place), chunk by chunk so the next layer's MMAs start while
//                          the tail is still draining.  The last Linear (H -> d_out <= 8) is a
//                          CUDA-core dot product in the same pass, feeding a per-row Welford
//                          (count, mean, M2) across members that lives in registers -- the
//                          [K, N, out] stack of nnueehcs/models.py:103,159 never exists.
//
// Everything shape-dependent is a template parameter (hidden width H, padded output count): the
// single-thread producer / MMA-issue loops must retire a 32 KB weight stage (4 MMAs = 512 tensor
// cycles) in well under 512 issue cycles, which a runtime-generic loop cannot do -- the first
// version of this kernel spent ~2000 cycles per stage on loop overhead alone (profiles/).
//
// Layer 0 (d_in <= 21 inputs) is also an MMA: x is split into bf16 hi + lo parts and the folded
// first-layer weights likewise, laid out as [x_hi | x_lo | x_hi] . [w_hi | w_hi | w_lo] inside
// one K <= 64 chunk, which recovers ~16 mantissa bits for one extra K=16 step.
//
// Replaces: EnsembleModel.forward (models.py:99-108), MCDropoutModel.forward (:147-163) and the
// anchored forward behind DeltaUQMLP.forward (:313-341) for MLPs whose hidden widths are equal.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "philox.cuh"
#include "tc_params.cuh"
#include "tc_ptx.cuh"

namespace uq {

namespace {

using namespace tc;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;  // 320
constexpr int EPI_THREADS = NUM_EPI_WARPS * 32;
constexpr int SMEM_LIMIT = 232448;             // 227 KB opt-in maximum per CTA

// compile-time geometry of one hidden width
template <int H, int DOUT>
struct Geo {
  static_assert(H % 64 == 0 && H >= 64 && H <= 512, "hidden width must be a multiple of 64 <= 512");
  static constexpr int KC = H / CHUNK_K;                 // activation chunks == K-chunks per layer
  static constexpr int NH = (H + 255) / 256;             // accumulator halves (MMA N <= 256)
  static constexpr int NT = H / NH;                      // MMA N
  static_assert(NT % 16 == 0, "MMA N must be a multiple of 16");
  static constexpr int TMEM_COLS = H <= 64 ? 64 : H <= 128 ? 128 : H <= 256 ? 256 : 512;
  static constexpr int STAGE_BYTES = NT * 128;           // [NT x 64] bf16
  static constexpr int A_BYTES = KC * CHUNK_BYTES;
  static constexpr int XCHG_BYTES = 2 * TILE_M * DOUT * 4;
  static constexpr int MISC_BYTES = 1024 /*align slack*/ + 256 /*barriers*/ + XCHG_BYTES;
  // narrow nets leave room for two CTAs per SM (one CTA's epilogue overlaps the other's MMAs)
  static constexpr int BUDGET =
      (H <= 128 ? SMEM_LIMIT / 2 - 1024 : SMEM_LIMIT) - A_BYTES - MISC_BYTES;
  static constexpr int NS_RAW = BUDGET / STAGE_BYTES;
  static constexpr int NSTAGES = NS_RAW > 8 ? 8 : NS_RAW;
  static_assert(NSTAGES >= 2, "not enough shared memory for a weight ring");
  static constexpr int SMEM_BYTES = A_BYTES + NSTAGES * STAGE_BYTES + MISC_BYTES;
  // last chunk overlapping accumulator half nh
  __host__ __device__ static constexpr int hi(int nh) { return ((nh + 1) * NT + CHUNK_K - 1) / CHUNK_K - 1; }
};


// barrier block (byte offsets inside the 256-byte barrier area)
constexpr uint32_t BAR_W_FULL = 0;       // 8 x 8 B
constexpr uint32_t BAR_W_EMPTY = 64;     // 8 x 8 B
constexpr uint32_t BAR_CHUNK = 128;      // 8 x 8 B
constexpr uint32_t BAR_D_FULL = 192;
constexpr uint32_t BAR_X_READY = 200;
constexpr uint32_t BAR_TMEM_PTR = 208;

// input feature i of the network for (sample row, member) -- x, or cat(x - a_k, a_k) for Delta-UQ
__device__ __forceinline__ float net_input(const TcParams& p, int64_t row, int member_global,
                                           int i) {
  if (row >= p.n) return 0.f;
  if (p.mode == UQ_MODE_DELTA_UQ) {
    const int d = p.d_x;
    const float a = __ldg(p.anchors + (int64_t)member_global * d + (i < d ? i : i - d));
    return i < d ? __ldg(p.x + row * d + i) - a : a;
  }
  return __ldg(p.x + row * p.d_x + i);
}

// ---- epilogue math on one 32-column block of one row ---------------------------------------------
// acc: raw accumulator bits; on return v[] holds bias + ReLU + dropout applied, fp32.
template <bool RELU, bool DROP>
__device__ __forceinline__ void epi_activate(const uint32_t (&acc)[32], float (&v)[32],
                                             const float* __restrict__ bias32, uint32_t keep,
                                             float keep_scale) {
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias32) + j4);
    v[j4 * 4 + 0] = __uint_as_float(acc[j4 * 4 + 0]) + bv.x;
    v[j4 * 4 + 1] = __uint_as_float(acc[j4 * 4 + 1]) + bv.y;
    v[j4 * 4 + 2] = __uint_as_float(acc[j4 * 4 + 2]) + bv.z;
    v[j4 * 4 + 3] = __uint_as_float(acc[j4 * 4 + 3]) + bv.w;
  }
  if (RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (DROP) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = ((keep >> j) & 1u) ? v[j] * keep_scale : 0.f;
  }
}

// keep-mask bits of 32 consecutive features of one row
__device__ __forceinline__ uint32_t epi_keep_bits(const TcParams& p, int drop, int kg, int drop_ord,
                                                  int64_t grow, int col0,
                                                  const uint8_t* mask_layer, int H) {
  uint32_t keep = 0;
  if (drop == 2) {
#pragma unroll
    for (int gq = 0; gq < 4; ++gq)
      keep |= dropout_keep8(p.key, p.thr16, (uint32_t)kg, (uint32_t)drop_ord, (uint32_t)grow,
                            (uint32_t)(col0 / 8 + gq))
              << (8 * gq);
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

// one 32-column block: activation, then either the bf16 A-operand write-back or the last-Linear dot
template <int H, int DOUT>
__device__ __forceinline__ void epi_block(const TcParams& p, const uint32_t (&acc)[32],
                                          const float* __restrict__ bias32, bool relu, int drop,
                                          uint32_t keep, bool last, uint32_t a_dst, int piece0,
                                          int rx, const float* __restrict__ wl32,
                                          float (&dot)[DOUT]) {
  float v[32];
  if (relu) {
    if (drop) epi_activate<true, true>(acc, v, bias32, keep, p.drop_scale);
    else epi_activate<true, false>(acc, v, bias32, keep, p.drop_scale);
  } else {
    if (drop) epi_activate<false, true>(acc, v, bias32, keep, p.drop_scale);
    else epi_activate<false, false>(acc, v, bias32, keep, p.drop_scale);
  }
  if (!last) {
#pragma unroll
    for (int pc = 0; pc < 4; ++pc)
      st_shared_v4(a_dst + (uint32_t)(((piece0 + pc) ^ rx) << 4),
                   pack_bf16x2(v[pc * 8 + 0], v[pc * 8 + 1]),
                   pack_bf16x2(v[pc * 8 + 2], v[pc * 8 + 3]),
                   pack_bf16x2(v[pc * 8 + 4], v[pc * 8 + 5]),
                   pack_bf16x2(v[pc * 8 + 6], v[pc * 8 + 7]));
  } else {
#pragma unroll
    for (int o = 0; o < DOUT; ++o) {
      float s = dot[o];
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wl32 + o * H) + j4);
        s = fmaf(v[j4 * 4 + 0], wv.x, s);
        s = fmaf(v[j4 * 4 + 1], wv.y, s);
        s = fmaf(v[j4 * 4 + 2], wv.z, s);
        s = fmaf(v[j4 * 4 + 3], wv.w, s);
      }
      dot[o] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// the fused kernel
// ------------------------------------------------------------------------------------------------
template <int H, int DOUT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
uq_mlp_tc_kernel(const __grid_constant__ TcParams p) {
  using G = Geo<H, DOUT>;
  constexpr int KC = G::KC, NH = G::NH, NT = G::NT, NS = G::NSTAGES;
  constexpr uint32_t STAGE_BYTES = G::STAGE_BYTES;

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_smem = smem;                                  // KC chunks of 16 KB
  uint8_t* w_smem = smem + G::A_BYTES;                     // NS stages
  uint8_t* bar_smem = w_smem + NS * STAGE_BYTES;           // 256 B of mbarriers + tmem pointer
  const uint32_t xchg = smem_u32(bar_smem + 256);          // [2][128][DOUT] dot exchange (floats)
  const uint32_t a_base = smem_u32(a_smem);
  const uint32_t w_base = smem_u32(w_smem);
  const uint32_t bars = smem_u32(bar_smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_units = p.n_tiles * p.splits;
#ifdef UQ_TC_TRACE
  int tr_n = 0;
  auto trace = [&](int role, unsigned kind, unsigned idx) {
    if (p.trace != nullptr && blockIdx.x == 0 && tr_n < TRACE_LEN && (role != 1 || lane == 0)) {
      unsigned long long* t = p.trace + ((size_t)role * TRACE_LEN + tr_n) * 2;
      t[0] = ((unsigned long long)kind << 24) | idx;
      t[1] = (unsigned long long)clock64();
      ++tr_n;
    }
  };
#define UQ_TRACE(role, kind, idx) trace(role, kind, idx)
#else
#define UQ_TRACE(role, kind, idx)
#endif

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(bars + BAR_W_FULL + 8 * s, 1);
      mbar_init(bars + BAR_W_EMPTY + 8 * s, 1);
    }
    for (int c = 0; c < KC; ++c) mbar_init(bars + BAR_CHUNK + 8 * c, 4);
    mbar_init(bars + BAR_D_FULL, 1);
    mbar_init(bars + BAR_X_READY, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(bars + BAR_TMEM_PTR, (uint32_t)G::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(bar_smem + BAR_TMEM_PTR);

  if (warp == 0) {
    // ===================================== producer =============================================
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      unsigned tr_it = 0;
      (void)tr_it;
      const size_t member_bytes = (size_t)p.stages_per_member * STAGE_BYTES;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int split = unit % p.splits;
        const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
        const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
        for (int k = mb; k < me; ++k) {
          const uint8_t* src =
              p.image + (p.shared_weights ? 0 : (size_t)(p.member_begin + k) * member_bytes);
          for (int s = 0; s < p.stages_per_member; ++s) {
            mbar_wait(bars + BAR_W_EMPTY + 8 * slot, phase ^ 1, p.error_flag, 1);
            mbar_arrive_expect_tx(bars + BAR_W_FULL + 8 * slot, STAGE_BYTES);
            bulk_g2s(w_base + slot * STAGE_BYTES, src, STAGE_BYTES, bars + BAR_W_FULL + 8 * slot);
            UQ_TRACE(0, 2, tr_it++);
            src += STAGE_BYTES;
            if (++slot == NS) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ===========================================
    // The whole warp walks the loop convergently (operands stay warp-uniform, so the compiler
    // keeps descriptors in uniform registers); one elected lane issues tcgen05.mma / commit.
    // Barrier probes run one stage ahead: a try_wait costs ~180 cycles even when the phase has
    // already completed, so the probe for stage i+1 is issued before stage i's MMAs and its
    // result is only consumed afterwards.
    {
      constexpr uint32_t idesc = make_idesc_bf16(TILE_M, NT);
      const uint64_t a_desc0 = make_sw128_desc(a_base);
      const uint64_t b_desc0 = make_sw128_desc(w_base);
      const int k0_steps = p.K0 / 16;
      uint32_t slot = 0, phase = 0;
      uint32_t g = 0;      // layer-step counter (d_full / chunk_done phases)
      uint32_t xm = 0;     // member counter (x_ready phase)
      bool w_ready = mbar_try_wait(bars + BAR_W_FULL, 0);
#ifdef UQ_TC_TRACE
      unsigned tr_it = 0;
#endif
      // acquire the current weight stage and probe the next one
      uint32_t nslot = 0, nphase = 0;
      bool w_ready_next = false;
      auto acquire = [&]() {
        if (!w_ready) mbar_wait_slow(bars + BAR_W_FULL + 8 * slot, phase, p.error_flag, 4);
        tc_fence_after();
        nslot = slot + 1;
        nphase = phase;
        if (nslot == NS) { nslot = 0; nphase ^= 1; }
        w_ready_next = mbar_try_wait(bars + BAR_W_FULL + 8 * nslot, nphase);
      };
      auto release = [&]() {
        if (elect_one()) umma_commit(bars + BAR_W_EMPTY + 8 * slot);  // frees the stage on retire
        slot = nslot;
        phase = nphase;
        w_ready = w_ready_next;
      };
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int split = unit % p.splits;
        const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
        const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
        for (int k = mb; k < me; ++k, ++xm) {
          // ---- layer 0: A = split input row in chunk 0, K0 <= 64 ------------------------------
          {
            const uint32_t prev_par = (g - 1) & 1;
            mbar_wait(bars + BAR_X_READY, xm & 1, p.error_flag, 2);
#pragma unroll
            for (int nh = 0; nh < NH; ++nh) {
              if (g != 0) {  // accumulator half nh must have been drained by the previous epilogue
                uint32_t ok = 0;
#pragma unroll
                for (int c = 0; c < KC; ++c)
                  if (c >= (nh == 0 ? 0 : G::hi(nh - 1) + 1) && c <= G::hi(nh))
                    ok |= (mbar_try_wait(bars + BAR_CHUNK + 8 * c, prev_par) ? 1u : 0u) << c;
#pragma unroll
                for (int c = 0; c < KC; ++c)
                  if (c >= (nh == 0 ? 0 : G::hi(nh - 1) + 1) && c <= G::hi(nh))
                    if (!((ok >> c) & 1u))
                      mbar_wait_slow(bars + BAR_CHUNK + 8 * c, prev_par, p.error_flag, 3);
              }
              acquire();
              if (elect_one()) {
                const uint64_t bd = b_desc0 + (uint64_t)((slot * STAGE_BYTES) >> 4);
                for (int ks = 0; ks < k0_steps; ++ks)
                  umma_bf16(tmem_base + nh * NT, a_desc0 + 2 * ks, bd + 2 * ks, idesc,
                            ks > 0 ? 1u : 0u);
              }
              release();
            }
            if (elect_one()) umma_commit(bars + BAR_D_FULL);
            UQ_TRACE(1, 4, g);
            ++g;
          }
          // ---- hidden layers: A = previous activations (in-place chunks), K = H ----------------
          for (int l = 1; l < p.L_mma; ++l) {
            const uint32_t prev_par = (g - 1) & 1;
            // chunks 0..hi(0): accumulator half 0 drained and the first A chunks written
            {
              uint32_t ok = 0;
#pragma unroll
              for (int c = 0; c <= G::hi(0); ++c)
                ok |= (mbar_try_wait(bars + BAR_CHUNK + 8 * c, prev_par) ? 1u : 0u) << c;
#pragma unroll
              for (int c = 0; c <= G::hi(0); ++c)
                if (!((ok >> c) & 1u))
                  mbar_wait_slow(bars + BAR_CHUNK + 8 * c, prev_par, p.error_flag, 3);
            }
            bool c_ready = (G::hi(0) + 1 < KC)
                               ? mbar_try_wait(bars + BAR_CHUNK + 8 * (G::hi(0) + 1), prev_par)
                               : true;
#pragma unroll
            for (int nh = 0; nh < NH; ++nh) {
#pragma unroll
              for (int kc = 0; kc < KC; ++kc) {
                if (nh == 0 && kc > G::hi(0)) {  // A chunk kc (and its accumulator columns)
                  if (!c_ready) mbar_wait_slow(bars + BAR_CHUNK + 8 * kc, prev_par, p.error_flag, 3);
                  if (kc + 1 < KC) c_ready = mbar_try_wait(bars + BAR_CHUNK + 8 * (kc + 1), prev_par);
                }
                UQ_TRACE(1, 1, tr_it);
                acquire();
                UQ_TRACE(1, 2, tr_it);
                if (elect_one()) {
                  const uint64_t ad = a_desc0 + (uint64_t)((kc * CHUNK_BYTES) >> 4);
                  const uint64_t bd = b_desc0 + (uint64_t)((slot * STAGE_BYTES) >> 4);
#pragma unroll
                  for (int ks = 0; ks < CHUNK_K / 16; ++ks)
                    umma_bf16(tmem_base + nh * NT, ad + 2 * ks, bd + 2 * ks, idesc,
                              (kc > 0 || ks > 0) ? 1u : 0u);
                }
                release();
                UQ_TRACE(1, 3, tr_it++);
              }
            }
            if (elect_one()) umma_commit(bars + BAR_D_FULL);  // whole layer accumulated
            UQ_TRACE(1, 4, g);
            ++g;
          }
        }
      }
    }
  } else {
    // ===================================== epilogue =============================================
    const int ew = warp - 2;             // 0..7
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int hf = ew >> 2;              // column-chunk parity handled by this warp
    const int row = q * 32 + lane;       // row of the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t g = 0;                      // layer-step counter
    uint32_t mcount = 0;                 // members processed (exchange buffer parity)
    const uint32_t a_row = a_base + (row >> 3) * 1024 + (row & 7) * 128;  // this row in chunk 0
    const int rx = row & 7;

    // writes the layer-0 A operand (split input row) into chunk 0, pieces [0, K0/8)
    auto write_x = [&](int tile, int member_global) {
      if (hf == 0) {
        const int64_t grow = (int64_t)tile * TILE_M + row;
        const int d = p.d_in;
        int seg = 0, i = 0;
        for (int piece = 0; piece < p.K0 / 8; ++piece) {
          uint32_t w4[4];
#pragma unroll
          for (int h2 = 0; h2 < 4; ++h2) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              float out = 0.f;
              if (seg < p.split_s) {
                const float f = net_input(p, grow, member_global, i);
                const float hi = __bfloat162float(__float2bfloat16_rn(f));
                out = (seg == 1) ? (f - hi) : hi;   // [hi | lo | hi]
              }
              v[e] = out;
              if (++i == d) { i = 0; ++seg; }
            }
            w4[h2] = pack_bf16x2(v[0], v[1]);
          }
          st_shared_v4(a_row + (uint32_t)((piece ^ rx) << 4), w4[0], w4[1], w4[2], w4[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + BAR_X_READY);
      }
    };

    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const int tile = unit / p.splits, split = unit % p.splits;
      const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
      const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
      const int64_t grow = (int64_t)tile * TILE_M + row;

      float wf_n = 0.f, wf_mean[DOUT], wf_m2[DOUT];
#pragma unroll
      for (int o = 0; o < DOUT; ++o) wf_mean[o] = 0.f, wf_m2[o] = 0.f;

      if (unit == (int)blockIdx.x) write_x(tile, p.member_begin + mb);  // very first member

      for (int k = mb; k < me; ++k, ++mcount) {
        const int kg = p.member_begin + k;                 // global member / pass id
        const int wslot = p.shared_weights ? 0 : kg;
        float dot[DOUT];
#pragma unroll
        for (int o = 0; o < DOUT; ++o) dot[o] = 0.f;
        int drop_ord = 0;
        const uint8_t* mask_layer = p.masks;

        for (int l = 0; l < p.L_mma; ++l, ++g) {
          const bool last = (l == p.L_mma - 1);
          const bool relu = (p.relu_mask >> l) & 1u;
          const bool has_drop = (p.dropout_mask >> l) & 1u;
          const int drop = has_drop ? p.drop_mode : 0;
          const float* bias = p.bias[l] + (size_t)wslot * H;
          const float* wl = p.w_last + (size_t)wslot * DOUT * H;

          // one lane polls the layer barrier, the rest of the warp parks on the warp barrier
          if (lane == 0) mbar_wait(bars + BAR_D_FULL, g & 1, p.error_flag, 5);
          __syncwarp();
          tc_fence_after();
          if (warp == 2 && lane == 0) { UQ_TRACE(2, 1, g); }
          if (warp == 6 && lane == 0) { UQ_TRACE(2, 5, g); }

          if (last) {
            // every MMA that reads the A chunks has retired: stage the next member's input row
            int nk = k + 1, ntile = tile;
            bool have_next = true;
            if (nk >= me) {
              const int nunit = unit + gridDim.x;
              have_next = nunit < n_units;
              ntile = nunit / p.splits;
              nk = (int)(((int64_t)p.member_count * (nunit % p.splits)) / p.splits);
            }
            if (have_next) write_x(ntile, p.member_begin + nk);
          }

          // --- drain this warp's chunks; TMEM loads run one 32-column block ahead ---------------
          uint32_t acc0[32], acc1[32];
          if (hf < KC) tmem_ld32(lane_addr + (uint32_t)(hf * CHUNK_K), acc0);
#pragma unroll 1
          for (int c = hf; c < KC; c += 2) {
            const int col0 = c * CHUNK_K;
            const uint32_t a_dst = a_row + (uint32_t)c * CHUNK_BYTES;
            uint32_t keep = 0xffffffffu;
            // ---- block 0 (columns col0 .. col0+31) ----
            tmem_ld_wait();
            tmem_ld32(lane_addr + (uint32_t)(col0 + 32), acc1);
            if (drop) keep = epi_keep_bits(p, drop, kg, drop_ord, grow, col0, mask_layer, H);
            epi_block<H, DOUT>(p, acc0, bias + col0, relu, drop, keep, last, a_dst, 0, rx,
                               wl + col0, dot);
            // ---- block 1 (columns col0+32 .. col0+63) ----
            tmem_ld_wait();
            if (c + 2 < KC) tmem_ld32(lane_addr + (uint32_t)(col0 + 2 * CHUNK_K), acc0);
            if (drop) keep = epi_keep_bits(p, drop, kg, drop_ord, grow, col0 + 32, mask_layer, H);
            epi_block<H, DOUT>(p, acc1, bias + col0 + 32, relu, drop, keep, last, a_dst, 4, rx,
                               wl + col0 + 32, dot);
            // chunk c: accumulator columns drained (+ A chunk rewritten) -> release to the MMA warp
            tc_fence_before();
            if (!last) fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + BAR_CHUNK + 8 * c);
            if (warp == 2 && lane == 0) { UQ_TRACE(2, 2, (g << 4) | c); }
            if (warp == 6 && lane == 0) { UQ_TRACE(2, 6, (g << 4) | c); }
          }
          if (has_drop) {
            if (p.masks) mask_layer += (size_t)p.total_members * (size_t)p.n * (size_t)H;
            ++drop_ord;
          }
        }

        // ---- combine the two column-parity halves of the dot products, then Welford ----------
        const uint32_t xb = xchg + (uint32_t)(((mcount & 1) * TILE_M + row) * DOUT * 4);
        if (KC > 1) {
          if (hf == 1) {
#pragma unroll
            for (int o = 0; o < DOUT; ++o) st_shared_f32(xb + 4 * o, dot[o]);
          }
          epi_bar_sync();
        }
        if (hf == 0) {
          wf_n += 1.f;
          const float inv_n = 1.f / wf_n;
          const float* bl = p.b_last + (size_t)wslot * DOUT;
#pragma unroll
          for (int o = 0; o < DOUT; ++o) {
            float y = dot[o] + (KC > 1 ? ld_shared_f32(xb + 4 * o) : 0.f) + __ldg(bl + o);
            if (p.last_relu) y = fmaxf(y, 0.f);
            const float dlt = y - wf_mean[o];
            wf_mean[o] += dlt * inv_n;
            wf_m2[o] = fmaf(dlt, y - wf_mean[o], wf_m2[o]);
          }
        }
      }

      // ---- tile done: publish (mean, std) / (mean, M2) ------------------------------------------
      if (hf == 0 && grow < p.n) {
#pragma unroll
        for (int o = 0; o < DOUT; ++o) {
          if (o < p.d_out) {
            const int64_t idx = grow * p.d_out + o;
            if (p.splits > 1) {
              p.part_mean[(size_t)split * (size_t)p.n * p.d_out + idx] = wf_mean[o];
              p.part_m2[(size_t)split * (size_t)p.n * p.d_out + idx] = wf_m2[o];
            } else {
              p.out0[idx] = wf_mean[o];
              p.out1[idx] =
                  (p.output == UQ_OUT_MOMENTS) ? wf_m2[o] : sqrtf(wf_m2[o] / (wf_n - 1.f));
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)G::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// weight image packing
// ------------------------------------------------------------------------------------------------
// Writes one member's stages.  Stage order == consumption order of the kernel:
//   layer 0: NH stages (K0 real columns, [w_hi | w_hi | w_lo] split), layer l>=1: NH x KC stages.
__global__ void pack_image_kernel(__nv_bfloat16* __restrict__ image, const float* __restrict__ w,
                                  const float* __restrict__ alpha, int layer, int in, int H,
                                  int n_tile, int NH, int KC, int K0, int split_s,
                                  size_t stage_elems, int stage_base) {
  // one thread per (stage-local row, 16-byte piece)
  const int stages = (layer == 0) ? NH : NH * KC;
  const int64_t total = (int64_t)stages * n_tile * 8;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int piece = (int)(i & 7);
  const int r = (int)((i >> 3) % n_tile);
  const int s = (int)((i >> 3) / n_tile);
  const int nh = (layer == 0) ? s : s / KC;
  const int kc = (layer == 0) ? 0 : s % KC;
  const int n = nh * n_tile + r;                 // output feature
  const float sc = alpha ? alpha[n] : 1.0f;
  __nv_bfloat16 vals[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int col = piece * 8 + e;
    float out = 0.f;
    if (layer == 0) {
      const int seg = col / in, ii = col - seg * in;
      if (seg < split_s && col < K0) {
        const float f = w[(int64_t)n * in + ii] * sc;
        const float hi = __bfloat162float(__float2bfloat16_rn(f));
        out = (seg == 2) ? (f - hi) : hi;        // [hi | hi | lo]
      }
    } else {
      const int kk = kc * 64 + col;
      out = w[(int64_t)n * in + kk] * sc;
    }
    vals[e] = __float2bfloat16_rn(out);
  }
  uint8_t* dst = reinterpret_cast<uint8_t*>(image + (size_t)(stage_base + s) * stage_elems) +
                 sw128_offset(r, piece);
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(vals);
}

__global__ void fold_bias_kernel(const float* __restrict__ bias, const float* __restrict__ alpha,
                                 const float* __restrict__ beta, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float b = bias[i];
  if (alpha) b = b * alpha[i] + beta[i];
  out[i] = b;
}

// padded outputs of the last Linear: zero rows / zero bias beyond d_out
__global__ void pad_last_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                float* __restrict__ w_out, float* __restrict__ b_out, int K,
                                int d_out, int dpad, int H) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)K * dpad * H;
  if (i < total) {
    const int h = (int)(i % H);
    const int o = (int)((i / H) % dpad);
    const int k = (int)(i / ((int64_t)H * dpad));
    w_out[i] = o < d_out ? w[((int64_t)k * d_out + o) * H + h] : 0.f;
  }
  if (i < (int64_t)K * dpad) {
    const int o = (int)(i % dpad), k = (int)(i / dpad);
    b_out[i] = o < d_out ? b[(int64_t)k * d_out + o] : 0.f;
  }
}

int dout_pad(int d_out) { return d_out == 1 ? 1 : MAX_DOUT; }

template <int H, int DOUT>
int launch_tc(const TcParams& p, int64_t units, cudaStream_t st) {
  using G = Geo<H, DOUT>;
  auto kern = uq_mlp_tc_kernel<H, DOUT>;
  UQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NUM_THREADS, G::SMEM_BYTES);
  if (per_sm < 1) per_sm = 1;
  if (per_sm * G::TMEM_COLS > 512) per_sm = 512 / G::TMEM_COLS;  // TMEM columns are per SM
  if (per_sm > 2) per_sm = 2;
  const int grid = (int)(units < (int64_t)sms * per_sm ? units : (int64_t)sms * per_sm);
  kern<<<grid, NUM_THREADS, G::SMEM_BYTES, st>>>(p);
  UQ_LAUNCH_CHECK();
  return UQ_OK;
}

template <int DOUT>
int dispatch_h(int H, const TcParams& p, int64_t units, cudaStream_t st) {
  switch (H) {
    case 64: return launch_tc<64, DOUT>(p, units, st);
    case 128: return launch_tc<128, DOUT>(p, units, st);
    case 192: return launch_tc<192, DOUT>(p, units, st);
    case 256: return launch_tc<256, DOUT>(p, units, st);
    case 320: return launch_tc<320, DOUT>(p, units, st);
    case 384: return launch_tc<384, DOUT>(p, units, st);
    case 448: return launch_tc<448, DOUT>(p, units, st);
    case 512: return launch_tc<512, DOUT>(p, units, st);
  }
  set_error("bf16 kernel: unsupported hidden width %d", H);
  return UQ_ERR_UNSUPPORTED;
}

}  // namespace

void tc_plan(uq_model* m) {
  TcPlan& t = m->tc;
  t.ok = false;
  const int L = m->n_layers;
  if (L < 2) { t.why_not = "needs at least one hidden Linear"; return; }
  const int H = m->layers[0].out;
  for (int l = 0; l < L - 1; ++l) {
    if (m->layers[l].out != H) { t.why_not = "hidden widths differ"; return; }
  }
  if ((H % 64 != 0 || H < 64 || H > 512) && !tc3_supported(H)) {
    t.why_not = "hidden width must be a multiple of 64 in [64, 512], or 768 / 1024 (got " +
                std::to_string(H) + ")";
    return;
  }
  const int NH = (H + 255) / 256;
  const int n_tile = H / NH;
  if (L - 1 > MAX_MMA_LAYERS) { t.why_not = "too many layers"; return; }
  const Layer& last = m->layers[L - 1];
  if (last.out > MAX_DOUT) { t.why_not = "final out_features > 8"; return; }
  if (last.has_bn || last.dropout) { t.why_not = "BatchNorm/Dropout after the final Linear"; return; }
  const int d_in = m->layers[0].in;
  int s = 3;
  while (s > 1 && s * d_in > 64) --s;
  if (s * d_in > 64) { t.why_not = "input width > 64"; return; }
  t.d_in = d_in;
  t.k0 = ((s * d_in + 15) / 16) * 16;
  t.hidden = H;
  t.n_mma_layers = L - 1;
  t.d_out = last.out;
  t.n_tile = n_tile;
  t.stage_bytes = (size_t)n_tile * 128;
  t.stages_per_member = NH + (L - 2) * NH * (H / 64);
  t.ok = true;
}

static int split_factor(const TcPlan& t) {
  int s = 3;
  while (s > 1 && s * t.d_in > 64) --s;
  return s;
}

int tc_pack(uq_model* m, cudaStream_t st) {
  TcPlan& t = m->tc;
  const int K = m->n_members, H = t.hidden, NH = H / t.n_tile, KC = H / 64;
  const size_t stage_elems = t.stage_bytes / sizeof(__nv_bfloat16);
  const size_t member_elems = (size_t)t.stages_per_member * stage_elems;
  void* p = nullptr;
  UQ_CUDA(cudaMalloc(&p, member_elems * K * sizeof(__nv_bfloat16)));
  m->allocations.push_back(p);
  t.image = static_cast<__nv_bfloat16*>(p);
  UQ_CUDA(cudaMemsetAsync(p, 0, member_elems * K * sizeof(__nv_bfloat16), st));
  const int s = split_factor(t);
  for (int k = 0; k < K; ++k) {
    int stage_base = 0;
    for (int l = 0; l < t.n_mma_layers; ++l) {
      const Layer& ly = m->layers[l];
      const int stages = (l == 0) ? NH : NH * KC;
      const int64_t total = (int64_t)stages * t.n_tile * 8;
      pack_image_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
          t.image + (size_t)k * member_elems, ly.w + (size_t)k * ly.out * ly.in,
          ly.has_bn ? ly.alpha + (size_t)k * ly.out : nullptr, l, ly.in, H, t.n_tile, NH, KC,
          t.k0, s, stage_elems, stage_base);
      UQ_LAUNCH_CHECK();
      stage_base += stages;
    }
  }
  for (int l = 0; l < m->n_layers; ++l) {
    const Layer& ly = m->layers[l];
    const int n = ly.out * K;
    fold_bias_kernel<<<(n + 255) / 256, 256, 0, st>>>(ly.bias, ly.has_bn ? ly.alpha : nullptr,
                                                     ly.has_bn ? ly.beta : nullptr,
                                                     ly.bias_folded, n);
    UQ_LAUNCH_CHECK();
  }
  // last Linear, padded to the kernel's compile-time output count
  const Layer& last = m->layers[m->n_layers - 1];
  const int dpad = dout_pad(t.d_out);
  void *wl = nullptr, *bl = nullptr;
  UQ_CUDA(cudaMalloc(&wl, sizeof(float) * (size_t)K * dpad * H));
  m->allocations.push_back(wl);
  UQ_CUDA(cudaMalloc(&bl, sizeof(float) * (size_t)K * dpad));
  m->allocations.push_back(bl);
  const int64_t total = (int64_t)K * dpad * H;
  pad_last_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      last.w, last.bias_folded, static_cast<float*>(wl), static_cast<float*>(bl), K, t.d_out, dpad,
      H);
  UQ_LAUNCH_CHECK();
  t.w_last = static_cast<float*>(wl);
  t.b_last = static_cast<float*>(bl);
  return UQ_OK;
}

static int choose_splits(const uq_model* m, int64_t n, const uq_forward_args* a) {
  const int64_t tiles = (n + TILE_M - 1) / TILE_M;
  int splits = 1;
  if (a->output == UQ_OUT_MOMENTS) return 1;  // a K-shard hands raw moments to the caller
  // few sample tiles but many members/passes: also spread the member axis over the SMs
  while (tiles * splits < 2 * 148 && a->member_count / (splits * 2) >= 4 && splits < 64) splits *= 2;
  return splits;
}

size_t tc_workspace_bytes(const uq_model* m, int64_t n, const uq_forward_args* a) {
  const int splits = choose_splits(m, n, a);
  size_t b = 256;  // error flag
  if (splits > 1) b += 2 * (size_t)splits * (size_t)n * m->d_out * sizeof(float) + 512;
  return b;
}

int tc_forward(const uq_model* m, const float* x, int64_t n, const uq_forward_args* a,
               float* out0, float* out1, void* ws, size_t ws_bytes, cudaStream_t st) {
  const TcPlan& t = m->tc;
  const size_t need = tc_workspace_bytes(m, n, a);
  UQ_REQUIRE(ws != nullptr && ws_bytes >= need, UQ_ERR_WORKSPACE,
             "bf16 forward needs %zu workspace bytes, got %zu", need, ws_bytes);
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.x = x;
  p.n = n;
  p.d_in = m->d_in;
  p.d_x = (a->mode == UQ_MODE_DELTA_UQ) ? m->d_in / 2 : m->d_in;
  p.mode = a->mode;
  p.n_tiles = (int)((n + TILE_M - 1) / TILE_M);
  p.splits = choose_splits(m, n, a);
  p.member_begin = a->member_begin;
  p.member_count = a->member_count;
  p.total_members = a->total_members;
  p.K0 = t.k0;
  p.split_s = split_factor(t);
  p.L_mma = t.n_mma_layers;
  p.d_out = t.d_out;
  p.stages_per_member = t.stages_per_member;
  p.shared_weights = (a->mode != UQ_MODE_ENSEMBLE) ? 1 : 0;
  p.image = reinterpret_cast<const uint8_t*>(t.image);
  const bool mc = (a->mode == UQ_MODE_MC_DROPOUT) && a->dropout_active;
  for (int l = 0; l < t.n_mma_layers; ++l) {
    p.bias[l] = m->layers[l].bias_folded;
    if (m->layers[l].relu) p.relu_mask |= 1u << l;
    if (m->layers[l].dropout) p.dropout_mask |= 1u << l;
  }
  p.w_last = t.w_last;
  p.b_last = t.b_last;
  p.last_relu = m->layers[m->n_layers - 1].relu ? 1 : 0;
  p.drop_mode = !mc ? 0 : (a->masks ? 1 : 2);
  p.drop_scale = 1.0f / (float)(1.0 - a->dropout_p);
  p.thr16 = dropout_thr16((float)a->dropout_p);
  p.key = PhiloxKey{(uint32_t)(a->philox_seed & 0xffffffffu), (uint32_t)(a->philox_seed >> 32),
                    (uint32_t)(a->philox_offset & 0xffffffffu)};
  p.masks = a->masks;
  p.anchors = a->anchors;
  p.out0 = out0;
  p.out1 = out1;
  p.output = a->output;
  char* wsb = static_cast<char*>(ws);
  p.error_flag = reinterpret_cast<unsigned int*>(wsb);
  if (p.splits > 1) {
    const size_t part =
        (((size_t)p.splits * (size_t)n * m->d_out * sizeof(float)) + 255) & ~(size_t)255;
    p.part_mean = reinterpret_cast<float*>(wsb + 256);
    p.part_m2 = reinterpret_cast<float*>(wsb + 256 + part);
  }
  UQ_CUDA(cudaMemsetAsync(p.error_flag, 0, sizeof(unsigned int), st));
#ifdef UQ_TC_TRACE
  const char* trace_env = getenv("UQ_TC_TRACE_FILE");
  unsigned long long* d_trace = nullptr;
  if (trace_env && trace_env[0]) {
    UQ_CUDA(cudaMalloc(&d_trace, TRACE_ROLES * TRACE_LEN * 2 * sizeof(unsigned long long)));
    UQ_CUDA(cudaMemsetAsync(d_trace, 0, TRACE_ROLES * TRACE_LEN * 2 * sizeof(unsigned long long), st));
  }
  p.trace = d_trace;
#endif

  const int64_t units = (int64_t)p.n_tiles * p.splits;
  // CTA-pair kernel (mlp_tc2.cu) by default: measured faster than the one-CTA-per-tile kernel
  // of this file at every width (H = 128: 75 vs 89 ms on deltauq32_binomial_4M).
  // UQ_TC_VARIANT=1|2 overrides (bring-up / A-B measurements).
  int variant = tc2_supported(t.hidden) ? 2 : 1;
  if (const char* v = getenv("UQ_TC_VARIANT")) {
    if (v[0] == '1') variant = 1;
    if (v[0] == '2' && tc2_supported(t.hidden)) variant = 2;
  }
  if (t.hidden > 512) variant = 3;   // wide nets: 64 rows per CTA (mlp_tc3.cu)
  int rc;
  if (variant == 3)
    rc = tc3_launch(p, t.hidden, dout_pad(t.d_out), st);
  else if (variant == 2)
    rc = tc2_launch(p, t.hidden, dout_pad(t.d_out), st);
  else
    rc = (dout_pad(t.d_out) == 1) ? dispatch_h<1>(t.hidden, p, units, st)
                                  : dispatch_h<MAX_DOUT>(t.hidden, p, units, st);
  if (rc != UQ_OK) return rc;
#ifdef UQ_TC_TRACE
  if (d_trace) {  // bring-up aid: dump CTA 0's event timeline as CSV (role, kind, index, clock)
    std::vector<unsigned long long> h(TRACE_ROLES * TRACE_LEN * 2);
    UQ_CUDA(cudaMemcpyAsync(h.data(), d_trace, h.size() * sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost, st));
    UQ_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_trace);
    if (FILE* f = fopen(trace_env, "w")) {
      for (int r = 0; r < TRACE_ROLES; ++r)
        for (int i = 0; i < TRACE_LEN; ++i) {
          const unsigned long long tag = h[((size_t)r * TRACE_LEN + i) * 2];
          if (tag) fprintf(f, "%d,%llu,%llu,%llu\n", r, tag >> 24, tag & 0xFFFFFF,
                           h[((size_t)r * TRACE_LEN + i) * 2 + 1]);
        }
      fclose(f);
    }
  }
#endif
  if (p.splits > 1) {
    double counts[64];
    for (int s = 0; s < p.splits; ++s) {
      const int mb = (int)(((int64_t)p.member_count * s) / p.splits);
      const int me = (int)(((int64_t)p.member_count * (s + 1)) / p.splits);
      counts[s] = (double)(me - mb);
    }
    return moments_merge(p.part_mean, p.part_m2, counts, p.splits, n * m->d_out, out0, out1, st);
  }
  return UQ_OK;
}

}  // namespace uq
