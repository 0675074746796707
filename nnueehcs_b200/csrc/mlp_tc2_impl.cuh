// bf16 tcgen05 path of the UQ forward, CTA-pair variant (cta_group::2): the kernel and its launcher.
// Included by mlp_tc2.cu (epilogue-bias instantiations) and mlp_tc2_bias.cu (bias-in-the-MMA
// instantiations), two translation units so that the ~64 kernels compile in parallel.
//
// Same fused scheme as mlp_tc.cu -- a persistent, warp-specialised kernel that keeps the
// activations of a sample tile on the SM for the whole  members x layers  stack -- but two CTAs
// on the two SMs of a TPC work as one unit:
//
//   * each CTA owns its own 128-sample tile (A operand in its shared memory, [128 x H] fp32
//     accumulator in its tensor memory); one tcgen05.mma.cta_group::2 with M = 256 issued by the
//     leader CTA (cluster rank 0) drives both tensor cores;
//   * the B operand (a weight stage [N x 64] bf16) is split: each CTA streams only N/2 rows of
//     it, so the L2 -> SM weight traffic per flop halves and the same shared-memory budget holds
//     twice as many stages in flight.  The first version of this kernel (one CTA per tile) spent
//     ~760 cycles per 32 KB stage waiting on the weight ring against a 512-cycle MMA floor
//     (profiles/r01_c_*);
//   * cross-CTA signalling: the peer's weight-stage arrivals are relayed to the leader's "full"
//     barriers by the peer's (otherwise idle) MMA warp; epilogue warps of both CTAs arrive on the
//     leader's chunk barriers through mapa'd cluster addresses; tcgen05.commit multicasts
//     "stage free" and "layer accumulated" to both CTAs.
//
// Bias: BIAS = true (the default for d_out 1) lets the tensor core accumulate it -- one bias stage
// per layer and accumulator half in the weight ring, one extra K = 16 MMA of an all-ones A tile
// against it; BIAS = false stages the folded bias once per layer-step in shared memory (prefetched
// one step ahead) and adds it in the epilogue.
// Epilogue (warps 2..9 of each CTA): the accumulator is drained with double-buffered tcgen05.ld,
// [bias +] ReLU + bf16 rounding are one to two instructions per pair ([FFMA, FFMA,]
// cvt.rn.relu.bf16x2), and the result is stored straight into the next layer's swizzled A chunk.
// The last Linear is a CUDA-core dot product feeding the per-row Welford, as in mlp_tc.cu.
//
// Replaces: EnsembleModel.forward (models.py:99-108), MCDropoutModel.forward (:147-163) and the
// anchored forward behind DeltaUQMLP.forward (:313-341) for MLPs whose hidden widths are equal.
#pragma once
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "philox.cuh"
#include "tc_params.cuh"
#include "tc_ptx.cuh"
#include "tc_epilogue.cuh"


namespace uq {

namespace {

using namespace tc;
constexpr int SMEM_LIMIT = 232448;             // 227 KB opt-in maximum per CTA

// NG = epilogue warp groups (4 warps each, one per TMEM lane quarter); group j drains the
// activation chunks c with c % NG == j.
// BIAS: the folded bias is accumulated by the tensor core (one extra K = 16 MMA per layer and
// accumulator half: an all-ones A tile against a bias stage of the ring), not added by the epilogue.
template <int H, int DOUT, int NG, bool BIAS = false>
struct Geo2 {
  static constexpr int EPI_THREADS = NG * 128;
  static constexpr int NUM_THREADS = 64 + EPI_THREADS;
  static_assert(H % 64 == 0 && H >= 64 && H <= 512, "hidden width must be a multiple of 64 <= 512");
  static constexpr int KC = H / CHUNK_K;                 // activation chunks == K-chunks per layer
  static constexpr int NH = (H + 255) / 256;             // accumulator halves (MMA N <= 256)
  static constexpr int NT = H / NH;                      // MMA N of the pair
  static_assert(NT % 16 == 0, "MMA N must be a multiple of 16");
  static constexpr int TMEM_COLS = H <= 64 ? 64 : H <= 128 ? 128 : H <= 256 ? 256 : 512;
  static constexpr int STAGE_BYTES = NT * 128;           // whole stage [NT x 64] bf16 in the image
  static constexpr int HALF_BYTES = STAGE_BYTES / 2;     // what one CTA of the pair loads
  static constexpr int A_BYTES = KC * CHUNK_BYTES;
  static constexpr int WL_OFF = BIAS ? 0 : H;                  // w_last inside a step's aux block
  static constexpr int AUX_FLOATS = WL_OFF + (DOUT == 1 ? H : 0);   // [bias] [+ w_last] of one step
  static constexpr int ONES_BYTES = BIAS ? CONST_TILE_BYTES : 0;
  static constexpr int AUX_BYTES = 2 * AUX_FLOATS * 4;         // double-buffered by step parity
  static constexpr int XCHG_BYTES = 2 * (NG - 1) * TILE_M * DOUT * 4;
  static constexpr int XS_BYTES = TILE_M * 64;   // stash of the next layer-0 A rows (K0 <= 32)
  static constexpr int MISC_BYTES =
      1024 /*align slack*/ + 256 /*barriers*/ + XCHG_BYTES + XS_BYTES;
  static constexpr int BUDGET = SMEM_LIMIT - A_BYTES - AUX_BYTES - MISC_BYTES - ONES_BYTES;
  static constexpr int NS_RAW = BUDGET / HALF_BYTES;
  static constexpr int NSTAGES = NS_RAW > 8 ? 8 : NS_RAW;
  static_assert(NSTAGES >= 2, "not enough shared memory for a weight ring");
  static constexpr int SMEM_BYTES =
      A_BYTES + NSTAGES * HALF_BYTES + ONES_BYTES + AUX_BYTES + MISC_BYTES;
  __host__ __device__ static constexpr int hi(int nh) { return ((nh + 1) * NT + CHUNK_K - 1) / CHUNK_K - 1; }
};

// barrier block (byte offsets inside the 256-byte barrier area)
constexpr uint32_t BAR_W_FULL = 0;       // 8 x 8 B   leader: own bytes + peer relay; peer: own bytes
constexpr uint32_t BAR_W_EMPTY = 64;     // 8 x 8 B   commit multicast from the leader
constexpr uint32_t BAR_CHUNK = 128;      // 8 x 8 B   leader only: 4 warps of each CTA
constexpr uint32_t BAR_D_FULL = 192;     //           commit multicast from the leader
constexpr uint32_t BAR_X_READY = 200;    //           leader only: 4 warps of each CTA
constexpr uint32_t BAR_TMEM_PTR = 208;

struct DrainCtx {
  uint32_t lane_addr;      // TMEM address of this warp's lane quarter, column 0
  uint32_t a_row;          // smem address of this row inside chunk 0
  int rx;                  // row & 7
  int grp;                 // warp group: drains chunks c with c % NG == grp
  int lane;
  uint32_t chunk_bar0;     // cluster address of the leader's chunk barrier 0
  const float* bias_s;     // staged bias of this step (shared)
  const float* wl_s;       // staged w_last (shared, DOUT == 1)
  const float* wl_g;       // w_last of this member (global, DOUT > 1)
  int drop;                // 0 none, 1 injected, 2 philox
  int kg, drop_ord;
  int64_t grow;
  const uint8_t* mask_layer;
  float in_scale;          // 1/(1-p) if this layer's input went through an active dropout, else 1
#ifdef UQ_TC_TRACE
  unsigned long long* tr;  // this thread's trace slots (or nullptr)
  int* tr_n;
  unsigned g;
#endif
};

#ifdef UQ_TC_TRACE
__device__ __forceinline__ void drain_trace(const DrainCtx& cx, unsigned kind, unsigned c) {
  if (cx.tr != nullptr && *cx.tr_n < TRACE_LEN) {
    unsigned long long* t = cx.tr + (size_t)(*cx.tr_n) * 2;
    t[0] = ((unsigned long long)kind << 24) | (cx.g << 4) | c;
    t[1] = (unsigned long long)clock64();
    ++*cx.tr_n;
  }
}
#define UQ_DTRACE(kind, c) drain_trace(cx, kind, c)
#else
#define UQ_DTRACE(kind, c)
#endif

// Keep-mask words of this warp's blocks of one layer-step: word 2 i + b = block b of the warp's
// i-th chunk.  They depend on (pass, layer, row, feature) only, not on the activations, so the
// epilogue computes them BEFORE it waits for the layer's MMAs -- Philox4x32-10 costs ~11
// instructions per element, 5600 issue cycles per layer-step at H = 512, which used to sit between
// "layer accumulated" and "first chunk released" (49 % of peak with dropout, 74 % without).
template <int H, int NG>
struct KeepWords {
  static constexpr int CPW = (H / CHUNK_K + NG - 1) / NG;   // chunks per warp
  uint32_t w[2 * CPW];
};

template <int H, int NG>
__device__ __forceinline__ void compute_keep_words(const TcParams& p, const DrainCtx& cx,
                                                   KeepWords<H, NG>& kw) {
  constexpr int KC = H / CHUNK_K;
#pragma unroll
  for (int i = 0; i < KeepWords<H, NG>::CPW; ++i) {
    const int c = cx.grp + NG * i;
    if (c < KC) {
      kw.w[2 * i] = keep_bits32(p, cx.drop, cx.kg, cx.drop_ord, cx.grow, c * CHUNK_K, cx.mask_layer, H);
      kw.w[2 * i + 1] =
          keep_bits32(p, cx.drop, cx.kg, cx.drop_ord, cx.grow, c * CHUNK_K + 32, cx.mask_layer, H);
    }
  }
}

// Drain this warp's chunks of one layer-step.  NG == 2: 8 epilogue warps, TMEM loads run one
// 32-column block ahead in a second register buffer.  NG == 4: 16 epilogue warps (4 per
// scheduler) hide the tcgen05.ld / LDS latencies by thread-level parallelism instead, within the
// 112-register budget that 18 warps leave.
template <int H, int DOUT, int NG, bool RELU, bool DROP, bool LAST, bool BIAS = false>
__device__ __forceinline__ void drain_step(const TcParams& p, const DrainCtx& cx, int c_begin,
                                           const KeepWords<H, NG>& kw, float (&dot)[DOUT]) {
  constexpr int KC = H / CHUNK_K;
  if (NG == 2) {
    uint32_t acc0[32], acc1[32];
    if (c_begin < KC) tmem_ld32(cx.lane_addr + (uint32_t)(c_begin * CHUNK_K), acc0);
    auto chunk = [&](int c, uint32_t keep0, uint32_t keep1) {
      const int col0 = c * CHUNK_K;
      const uint32_t a_dst = cx.a_row + (uint32_t)c * CHUNK_BYTES;
      float4 bv[8];
      // ---- block 0 (columns col0 .. col0+31) ----
      if (!BIAS) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) bv[j4] = reinterpret_cast<const float4*>(cx.bias_s + col0)[j4];
      }
      UQ_DTRACE(10, c);
      tmem_ld_wait();
      UQ_DTRACE(11, c);
      tmem_ld32(cx.lane_addr + (uint32_t)(col0 + 32), acc1);
      epi_block2<H, DOUT, 32, RELU, DROP, LAST, BIAS>(acc0, bv, keep0, cx.in_scale, a_dst, 0, cx.rx,
                                            cx.wl_s + col0, cx.wl_g + col0, dot);
      UQ_DTRACE(12, c);
      // ---- block 1 (columns col0+32 .. col0+63) ----
      if (!BIAS) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          bv[j4] = reinterpret_cast<const float4*>(cx.bias_s + col0 + 32)[j4];
      }
      tmem_ld_wait();
      if (c + NG < KC) tmem_ld32(cx.lane_addr + (uint32_t)(col0 + NG * CHUNK_K), acc0);
      epi_block2<H, DOUT, 32, RELU, DROP, LAST, BIAS>(acc1, bv, keep1, cx.in_scale, a_dst, 4, cx.rx,
                                            cx.wl_s + col0 + 32, cx.wl_g + col0 + 32, dot);
      UQ_DTRACE(13, c);
      // chunk c: accumulator columns drained (+ A chunk rewritten) -> release to the MMA warp
      tc_fence_before();
      if (!LAST) fence_proxy_async_smem();
      UQ_DTRACE(14, c);
      __syncwarp();
      if (cx.lane == 0) mbar_arrive_cluster(cx.chunk_bar0 + 8 * c);
      UQ_DTRACE(15, c);
    };
    if (DROP) {   // unrolled: the precomputed keep words are indexed statically
#pragma unroll
      for (int i = 0; i < KeepWords<H, NG>::CPW; ++i) {
        const int c = c_begin + NG * i;
        if (c < KC) chunk(c, kw.w[2 * i], kw.w[2 * i + 1]);
      }
    } else {
#pragma unroll 1
      for (int c = c_begin; c < KC; c += NG) chunk(c, 0xffffffffu, 0xffffffffu);
    }
  } else {
#pragma unroll 1
    for (int c = c_begin; c < KC; c += NG) {
      const uint32_t a_dst = cx.a_row + (uint32_t)c * CHUNK_BYTES;
      uint32_t keep32 = 0xffffffffu;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int col0 = c * CHUNK_K + 16 * b;
        uint32_t acc[16];
        tmem_ld16(cx.lane_addr + (uint32_t)col0, acc);
        float4 bv[4];
        if (!BIAS) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4)
            bv[j4] = reinterpret_cast<const float4*>(cx.bias_s + col0)[j4];
        }
        if (DROP && (b & 1) == 0)
          keep32 = keep_bits32(p, cx.drop, cx.kg, cx.drop_ord, cx.grow, col0, cx.mask_layer, H);
        tmem_ld_wait();
        epi_block2<H, DOUT, 16, RELU, DROP, LAST, BIAS>(acc, bv, keep32 >> (16 * (b & 1)),
                                                        cx.in_scale, a_dst, 2 * b, cx.rx,
                                                        cx.wl_s + col0, cx.wl_g + col0, dot);
      }
      tc_fence_before();
      if (!LAST) fence_proxy_async_smem();
      __syncwarp();
      if (cx.lane == 0) mbar_arrive_cluster(cx.chunk_bar0 + 8 * c);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// the fused kernel (one cluster = two CTAs = two sample tiles)
// ------------------------------------------------------------------------------------------------
// MC: the launch has live dropout (MC-dropout passes).  The dropout-free instantiation carries no
// mask code at all -- the headline ensemble kernel keeps its 162 registers and zero spills whatever
// the dropout path needs (keep words live across the layer barrier wait).
template <int H, int DOUT, int NG, bool MC, bool BIAS = false>
__global__ void __launch_bounds__(Geo2<H, DOUT, NG, BIAS>::NUM_THREADS, 1)
uq_mlp_tc2_kernel(const __grid_constant__ TcParams p) {
  static_assert(!BIAS || DOUT == 1, "bias-in-the-MMA variant: d_out 1");
  using G = Geo2<H, DOUT, NG, BIAS>;
  constexpr int EPI_THREADS = G::EPI_THREADS;
  constexpr int KC = G::KC, NH = G::NH, NT = G::NT, NS = G::NSTAGES;
  constexpr uint32_t STAGE_BYTES = G::STAGE_BYTES, HALF_BYTES = G::HALF_BYTES;

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-byte alignment (pointer arithmetic on the shared array keeps the
  // address space visible to the compiler: bias / w_last reads below become LDS)
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* a_smem = smem;                                  // KC chunks of 16 KB
  uint8_t* w_smem = smem + G::A_BYTES;                     // NS half-stages
  uint8_t* ones_smem = w_smem + NS * HALF_BYTES;           // BIAS: constant-1 A tile (bf16)
  float* aux_smem = reinterpret_cast<float*>(ones_smem + G::ONES_BYTES);  // [2][AUX_FLOATS]
  uint8_t* bar_smem = reinterpret_cast<uint8_t*>(aux_smem) + G::AUX_BYTES;
  const uint32_t xchg = smem_u32(bar_smem + 256);  // [2][NG-1][128][DOUT] dot exchange (floats)
  const uint32_t xstash = xchg + G::XCHG_BYTES;    // [K0/8 pieces][128 rows] x 16 B
  const uint32_t a_base = smem_u32(a_smem);
  const uint32_t w_base = smem_u32(w_smem);
  const uint32_t bars = smem_u32(bar_smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int n_units = ((p.n_tiles + 1) >> 1) * p.splits;   // (tile pair, member split)
#ifdef UQ_TC_TRACE
  int tr_n = 0;
  auto trace = [&](int role, unsigned kind, unsigned idx) {
    if (p.trace != nullptr && blockIdx.x < 2 && tr_n < TRACE_LEN && (role != 1 || lane == 0)) {
      unsigned long long* t = p.trace + ((size_t)role * TRACE_LEN + tr_n) * 2;
      t[0] = ((unsigned long long)kind << 24) | idx;
      t[1] = (unsigned long long)clock64();
      ++tr_n;
    }
  };
  unsigned tr_it = 0;
  (void)tr_it;
#define UQ_TRACE(role, kind, idx) trace(role, kind, idx)
#else
#define UQ_TRACE(role, kind, idx)
#endif

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(bars + BAR_W_FULL + 8 * s, leader ? 2 : 1);
      mbar_init(bars + BAR_W_EMPTY + 8 * s, 1);
    }
    for (int c = 0; c < KC; ++c) mbar_init(bars + BAR_CHUNK + 8 * c, 8);
    mbar_init(bars + BAR_D_FULL, 1);
    mbar_init(bars + BAR_X_READY, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(bars + BAR_TMEM_PTR, (uint32_t)G::TMEM_COLS);
  if (BIAS) {   // the all-ones A tile of the bias MMAs (read by the tensor core: async proxy)
    for (int i = threadIdx.x; i < G::ONES_BYTES / 4; i += G::NUM_THREADS)
      reinterpret_cast<uint32_t*>(ones_smem)[i] = 0x3F803F80u;   // bf16 (1.0, 1.0)
    fence_proxy_async_smem();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(bar_smem + BAR_TMEM_PTR);

  if (warp == 0) {
    // ===================================== producer =============================================
    // streams this CTA's half (N/2 rows) of every weight stage
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      const size_t member_bytes = (size_t)p.stages_per_member * STAGE_BYTES;
      for (int unit = cluster_id; unit < n_units; unit += n_clusters) {
        const int split = unit % p.splits;
        const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
        const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
        for (int k = mb; k < me; ++k) {
          const uint8_t* src =
              p.image + (p.shared_weights ? 0 : (size_t)(p.member_begin + k) * member_bytes) +
              rank * HALF_BYTES;
          auto load = [&](const uint8_t* from) {
            mbar_wait(bars + BAR_W_EMPTY + 8 * slot, phase ^ 1, p.error_flag, 1);
            mbar_arrive_expect_tx(bars + BAR_W_FULL + 8 * slot, HALF_BYTES);
            bulk_g2s(w_base + slot * HALF_BYTES, from, HALF_BYTES, bars + BAR_W_FULL + 8 * slot);
            if (leader) { UQ_TRACE(0, 2, tr_it++); }
            if (++slot == NS) { slot = 0; phase ^= 1; }
          };
          if (BIAS) {   // per (layer, accumulator half): its weight stages, then its bias stage
            const uint8_t* bsrc =
                p.bias_image +
                (p.shared_weights ? 0 : (size_t)(p.member_begin + k) * p.L_mma * NH * STAGE_BYTES) +
                rank * HALF_BYTES;
            for (int l = 0; l < p.L_mma; ++l)
              for (int nh = 0; nh < NH; ++nh) {
                for (int s = 0; s < (l == 0 ? 1 : KC); ++s, src += STAGE_BYTES) load(src);
                if (l == 0 && p.bias0_image != nullptr)   // per-anchor layer-0 bias
                  load(p.bias0_image + ((size_t)(p.member_begin + k) * NH + nh) * STAGE_BYTES +
                       rank * HALF_BYTES);
                else
                  load(bsrc);
                bsrc += STAGE_BYTES;
              }
          } else {
            for (int s = 0; s < p.stages_per_member; ++s, src += STAGE_BYTES) load(src);
          }
        }
      }
    }
  } else if (warp == 1 && !leader) {
    // ===================================== peer relay ===========================================
    // tells the leader's MMA warp that this CTA's half of a stage has landed
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      const uint32_t full0 = mapa_shared(bars + BAR_W_FULL, 0);
      for (int unit = cluster_id; unit < n_units; unit += n_clusters) {
        const int split = unit % p.splits;
        const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
        const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
        const int n_stages = (me - mb) * (p.stages_per_member + (BIAS ? p.L_mma * NH : 0));
        for (int s = 0; s < n_stages; ++s) {
          mbar_wait(bars + BAR_W_FULL + 8 * slot, phase, p.error_flag, 6);
          mbar_arrive_cluster(full0 + 8 * slot);
          UQ_TRACE(5, 1, tr_it++);
          if (++slot == NS) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader) ===================================
    // The whole warp walks the loop convergently; one elected lane issues tcgen05.mma / commit.
    // Barrier probes run one stage ahead (a try_wait costs ~180 cycles even when the phase has
    // already completed).
    constexpr uint32_t idesc = make_idesc_bf16(2 * TILE_M, NT);
    const uint64_t a_desc0 = make_sw128_desc(a_base);
    const uint64_t b_desc0 = make_sw128_desc(w_base);
    const uint64_t ones_desc = make_sw128_const_desc(smem_u32(ones_smem));
    const int k0_steps = p.K0 / 16;
    uint32_t slot = 0, phase = 0;
    uint32_t g = 0;      // layer-step counter (d_full / chunk phases)
    uint32_t xm = 0;     // member counter (x_ready phase)
    bool w_ready = mbar_try_wait_cluster(bars + BAR_W_FULL, 0);
    uint32_t nslot = 0, nphase = 0;
    bool w_ready_next = false;
    auto acquire = [&]() {
      if (!w_ready) mbar_wait_cluster_inline(bars + BAR_W_FULL + 8 * slot, phase, p.error_flag, 4);
      tc_fence_after();
      nslot = slot + 1;
      nphase = phase;
      if (nslot == NS) { nslot = 0; nphase ^= 1; }
      w_ready_next = mbar_try_wait_cluster(bars + BAR_W_FULL + 8 * nslot, nphase);
    };
    auto release = [&]() {
      if (elect_one()) umma_commit_pair(bars + BAR_W_EMPTY + 8 * slot, 3);
      slot = nslot;
      phase = nphase;
      w_ready = w_ready_next;
    };
    // accumulator half nh += ones . bias^T: one K = 16 step on the half's bias stage
    auto bias_mma = [&](int nh) {
      acquire();
      if (elect_one())
        umma_bf16_pair(tmem_base + nh * NT, ones_desc,
                       b_desc0 + (uint64_t)((slot * HALF_BYTES) >> 4), idesc, 1u);
      release();
    };
    for (int unit = cluster_id; unit < n_units; unit += n_clusters) {
      const int split = unit % p.splits;
      const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
      const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
      for (int k = mb; k < me; ++k, ++xm) {
        // ---- layer 0: A = split input rows in chunk 0, K0 <= 64 --------------------------------
        {
          const uint32_t prev_par = (g - 1) & 1;
          mbar_wait_cluster(bars + BAR_X_READY, xm & 1, p.error_flag, 2);
#pragma unroll
          for (int nh = 0; nh < NH; ++nh) {
            if (g != 0) {  // accumulator half nh must have been drained by the previous epilogue
              uint32_t ok = 0;
#pragma unroll
              for (int c = 0; c < KC; ++c)
                if (c >= (nh == 0 ? 0 : G::hi(nh - 1) + 1) && c <= G::hi(nh))
                  ok |= (mbar_try_wait_cluster(bars + BAR_CHUNK + 8 * c, prev_par) ? 1u : 0u) << c;
#pragma unroll
              for (int c = 0; c < KC; ++c)
                if (c >= (nh == 0 ? 0 : G::hi(nh - 1) + 1) && c <= G::hi(nh))
                  if (!((ok >> c) & 1u))
                    mbar_wait_cluster_inline(bars + BAR_CHUNK + 8 * c, prev_par, p.error_flag, 3);
            }
            acquire();
            if (elect_one()) {
              const uint64_t bd = b_desc0 + (uint64_t)((slot * HALF_BYTES) >> 4);
              for (int ks = 0; ks < k0_steps; ++ks)
                umma_bf16_pair(tmem_base + nh * NT, a_desc0 + 2 * ks, bd + 2 * ks, idesc,
                               ks > 0 ? 1u : 0u);
            }
            release();
            if (BIAS) bias_mma(nh);
          }
          if (elect_one()) umma_commit_pair(bars + BAR_D_FULL, 3);
          UQ_TRACE(1, 4, g);
          ++g;
        }
        // ---- hidden layers: A = previous activations (in-place chunks), K = H ------------------
        for (int l = 1; l < p.L_mma; ++l) {
          const uint32_t prev_par = (g - 1) & 1;
          {
            uint32_t ok = 0;
#pragma unroll
            for (int c = 0; c <= G::hi(0); ++c)
              ok |= (mbar_try_wait_cluster(bars + BAR_CHUNK + 8 * c, prev_par) ? 1u : 0u) << c;
#pragma unroll
            for (int c = 0; c <= G::hi(0); ++c)
              if (!((ok >> c) & 1u))
                mbar_wait_cluster_inline(bars + BAR_CHUNK + 8 * c, prev_par, p.error_flag, 3);
          }
          if (!w_ready) w_ready = mbar_try_wait_cluster(bars + BAR_W_FULL + 8 * slot, phase);
          UQ_TRACE(1, 5, g);
          bool c_ready = (G::hi(0) + 1 < KC)
                             ? mbar_try_wait_cluster(bars + BAR_CHUNK + 8 * (G::hi(0) + 1), prev_par)
                             : true;
#pragma unroll
          for (int nh = 0; nh < NH; ++nh) {
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) {
              if (nh == 0 && kc > G::hi(0)) {  // A chunk kc (and its accumulator columns)
                if (!c_ready)
                  mbar_wait_cluster_inline(bars + BAR_CHUNK + 8 * kc, prev_par, p.error_flag, 3);
                if (kc + 1 < KC)
                  c_ready = mbar_try_wait_cluster(bars + BAR_CHUNK + 8 * (kc + 1), prev_par);
              }
              UQ_TRACE(1, 1, tr_it);
              acquire();
              UQ_TRACE(1, 2, tr_it);
              if (elect_one()) {
                const uint64_t ad = a_desc0 + (uint64_t)((kc * CHUNK_BYTES) >> 4);
                const uint64_t bd = b_desc0 + (uint64_t)((slot * HALF_BYTES) >> 4);
#pragma unroll
                for (int ks = 0; ks < CHUNK_K / 16; ++ks)
                  umma_bf16_pair(tmem_base + nh * NT, ad + 2 * ks, bd + 2 * ks, idesc,
                                 (kc > 0 || ks > 0) ? 1u : 0u);
              }
              release();
              UQ_TRACE(1, 3, tr_it++);
            }
            if (BIAS) bias_mma(nh);
          }
          if (elect_one()) umma_commit_pair(bars + BAR_D_FULL, 3);  // whole layer accumulated
          UQ_TRACE(1, 4, g);
          ++g;
        }
      }
    }
  } else {
    // ===================================== epilogue =============================================
    const int ew = warp - 2;             // 0 .. 4 NG - 1
    const int et = threadIdx.x - 64;     // 0 .. EPI_THREADS - 1
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int grp = ew >> 2;             // warp group: chunks c with c % NG == grp
    const int row = q * 32 + lane;       // row of the tile == TMEM lane
    const uint32_t a_row = a_base + (row >> 3) * 1024 + (row & 7) * 128;  // this row in chunk 0
    const int rx = row & 7;
    const uint32_t chunk_bar0 = mapa_shared(bars + BAR_CHUNK, 0);
    const uint32_t xready_bar = mapa_shared(bars + BAR_X_READY, 0);
    uint32_t g = 0;                      // layer-step counter
    uint32_t mcount = 0;                 // members processed (exchange buffer parity)
    constexpr int AUX_PER_THREAD = (G::AUX_FLOATS + EPI_THREADS - 1) / EPI_THREADS;

    // Layer-0 A operand (this row's split input [x_hi | x_lo | x_hi], K0 bf16) of one member.
    // Building it needs uncoalesced global reads and a generic packing loop (thousands of cycles),
    // so it is prepared off the critical path in a shared-memory stash -- once per tile, or once
    // per member for Delta-UQ whose input depends on the anchor -- and copied into chunk 0 with
    // K0/8 LDS/STS pairs at the moment the chunk becomes free.
    const bool use_stash = p.K0 <= 32;
    auto build_x = [&](int tile, int member_global, bool to_stash) {
      if (grp == 0)
        build_x_row(p, (int64_t)tile * TILE_M + row, member_global, to_stash,
                    xstash + (uint32_t)(row << 4), (uint32_t)(TILE_M << 4), a_row, rx);
    };
    // chunk 0 <- stash (or built in place when the stash is too small), then signal the MMA warp
    auto publish_x = [&](int tile, int member_global) {
      if (grp == 0) {
        if (use_stash) {
          for (int piece = 0; piece < p.K0 / 8; ++piece) {
            uint32_t a, b, c, d;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
                         : "r"(xstash + (uint32_t)((piece * TILE_M + row) << 4)));
            st_shared_v4(a_row + (uint32_t)((piece ^ rx) << 4), a, b, c, d);
          }
        } else {
          build_x(tile, member_global, false);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(xready_bar);
      }
    };

    // bias (+ w_last when the step is the member's last) of a step, one step ahead in registers
    float aux_pf[AUX_PER_THREAD];
    auto aux_prefetch = [&](int member_global, int l) {
      const int wslot = p.shared_weights ? 0 : member_global;
      const bool last = (l == p.L_mma - 1);
      const float* bias =
          p.bias[l] + (size_t)((l == 0 && p.bias0_per_member) ? member_global : wslot) * H;
      const float* wl = p.w_last + (size_t)wslot * DOUT * H;
#pragma unroll
      for (int j = 0; j < AUX_PER_THREAD; ++j) {
        const int i = et + j * EPI_THREADS;
        float v = 0.f;
        if (!BIAS && i < H) v = __ldg(bias + i);
        else if (DOUT == 1 && last && i >= G::WL_OFF && i < G::WL_OFF + H) v = __ldg(wl + (i - G::WL_OFF));
        aux_pf[j] = v;
      }
    };

    bool first_step = true;
    for (int unit = cluster_id; unit < n_units; unit += n_clusters) {
      const int tile = 2 * (unit / p.splits) + (int)rank, split = unit % p.splits;
      const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
      const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
      const int64_t grow = (int64_t)tile * TILE_M + row;

      float wf_n = 0.f, wf_mean[DOUT], wf_m2[DOUT];
#pragma unroll
      for (int o = 0; o < DOUT; ++o) wf_mean[o] = 0.f, wf_m2[o] = 0.f;

      if (first_step) {  // very first member of this CTA
        if (use_stash) build_x(tile, p.member_begin + mb, true);
        publish_x(tile, p.member_begin + mb);
        aux_prefetch(p.member_begin + mb, 0);
        first_step = false;
      }

      for (int k = mb; k < me; ++k, ++mcount) {
        const int kg = p.member_begin + k;                 // global member / pass id
        const int wslot = p.shared_weights ? 0 : kg;
        float dot[DOUT];
#pragma unroll
        for (int o = 0; o < DOUT; ++o) dot[o] = 0.f;
        int drop_ord = 0;
        const uint8_t* mask_layer = p.masks;

        // coordinates of the member after this one (for x staging and the bias prefetch)
        int nk = k + 1, ntile = tile;
        bool have_next = true;
        if (nk >= me) {
          const int nunit = unit + n_clusters;
          have_next = nunit < n_units;
          ntile = 2 * (nunit / p.splits) + (int)rank;
          nk = (int)(((int64_t)p.member_count * (nunit % p.splits)) / p.splits);
        }

        for (int l = 0; l < p.L_mma; ++l, ++g) {
          const bool last = (l == p.L_mma - 1);
          const bool relu = (p.relu_mask >> l) & 1u;
          const bool has_drop = (p.dropout_mask >> l) & 1u;
          const int drop = (MC && has_drop) ? p.drop_mode : 0;

          if (lane == 0 && (warp == 2 || (warp == 6 && leader))) { UQ_TRACE(leader ? (warp == 2 ? 2 : 3) : 4, 0, g); }
          // ---- publish this step's bias (+ w_last) in smem, prefetch the next step's ------------
          float* aux = aux_smem + (g & 1) * G::AUX_FLOATS;
#pragma unroll
          for (int j = 0; j < AUX_PER_THREAD; ++j) {
            const int i = et + j * EPI_THREADS;
            if (i < G::AUX_FLOATS) aux[i] = aux_pf[j];
          }
          epi_bar_sync_n<EPI_THREADS>();
          if (!last) aux_prefetch(kg, l + 1);
          else if (have_next) aux_prefetch(p.member_begin + nk, 0);
          // refresh the x stash while the last layer's MMAs run (its previous content went into
          // chunk 0 one member ago)
          if (last && have_next && use_stash && (ntile != tile))
            build_x(ntile, p.member_begin + nk, true);

          DrainCtx cx;
          cx.lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
          cx.a_row = a_row;
          cx.rx = rx;
          cx.grp = grp;
          cx.lane = lane;
          cx.chunk_bar0 = chunk_bar0;
          cx.bias_s = aux;
          cx.wl_s = aux + G::WL_OFF;
          cx.wl_g = p.w_last + (size_t)wslot * DOUT * H;
          cx.drop = drop;
          cx.kg = kg;
          cx.drop_ord = drop_ord;
          cx.grow = grow;
          cx.mask_layer = mask_layer;
          // epilogue-bias: 1/(1-p) owed by the previous layer's dropout; bias in the MMA: this layer's
          cx.in_scale = BIAS ? (drop ? p.drop_scale : 1.f)
                             : (l > 0 && ((p.dropout_mask >> (l - 1)) & 1u) && p.drop_mode)
                                   ? p.drop_scale : 1.f;
#ifdef UQ_TC_TRACE
          {
            const int role = (lane == 0 && (warp == 2 || (warp == 6 && leader)))
                                 ? (leader ? (warp == 2 ? 2 : 3) : 4) : -1;
            cx.tr = (role >= 0 && p.trace != nullptr && blockIdx.x < 2)
                        ? p.trace + (size_t)role * TRACE_LEN * 2 : nullptr;
            cx.tr_n = &tr_n;
            cx.g = g;
          }
#endif
#define UQ_STEP_DISPATCH(CALL_TTT, CALL_TFT, CALL_FTT, CALL_FFT, CALL_TTF, CALL_TFF, CALL_FTF, CALL_FFF) \
  if (last) {                                                                                      \
    if (relu) { if (MC && drop) { CALL_TTT; } else { CALL_TFT; } }                                 \
    else { if (MC && drop) { CALL_FTT; } else { CALL_FFT; } }                                      \
  } else {                                                                                         \
    if (relu) { if (MC && drop) { CALL_TTF; } else { CALL_TFF; } }                                 \
    else { if (MC && drop) { CALL_FTF; } else { CALL_FFF; } }                                      \
  }
          // keep masks of this step while the layer's MMAs still run (see KeepWords)
          KeepWords<H, NG> kw;
          if (MC && NG == 2 && drop) compute_keep_words<H, NG>(p, cx, kw);
          // one lane polls the layer barrier, the rest of the warp parks on the warp barrier
          if (lane == 0) mbar_wait(bars + BAR_D_FULL, g & 1, p.error_flag, 5);
          __syncwarp();
          tc_fence_after();
          if (lane == 0 && (warp == 2 || (warp == 6 && leader))) { UQ_TRACE(leader ? (warp == 2 ? 2 : 3) : 4, 1, g); }

          // every MMA that reads the A chunks has retired: stage the next member's input rows
          if (last && have_next) publish_x(ntile, p.member_begin + nk);

          const int c_begin = grp;
          UQ_STEP_DISPATCH((drain_step<H, DOUT, NG, true, MC, true, BIAS>(p, cx, c_begin, kw, dot)),
                           (drain_step<H, DOUT, NG, true, false, true, BIAS>(p, cx, c_begin, kw, dot)),
                           (drain_step<H, DOUT, NG, false, MC, true, BIAS>(p, cx, c_begin, kw, dot)),
                           (drain_step<H, DOUT, NG, false, false, true, BIAS>(p, cx, c_begin, kw, dot)),
                           (drain_step<H, DOUT, NG, true, MC, false, BIAS>(p, cx, c_begin, kw, dot)),
                           (drain_step<H, DOUT, NG, true, false, false, BIAS>(p, cx, c_begin, kw, dot)),
                           (drain_step<H, DOUT, NG, false, MC, false, BIAS>(p, cx, c_begin, kw, dot)),
                           (drain_step<H, DOUT, NG, false, false, false, BIAS>(p, cx, c_begin, kw, dot)))
          if (lane == 0 && (warp == 2 || (warp == 6 && leader))) { UQ_TRACE(leader ? (warp == 2 ? 2 : 3) : 4, 2, g); }
          if (has_drop) {
            if (p.masks) mask_layer += (size_t)p.total_members * (size_t)p.n * (size_t)H;
            ++drop_ord;
          }
        }

        // ---- combine the groups' partial dot products, then Welford ---------------------------
        constexpr int NPART = (KC < NG ? KC : NG) - 1;   // groups other than 0 that hold a part
        const uint32_t xb = xchg + (uint32_t)((mcount & 1) * (NG - 1) * TILE_M * DOUT * 4);
        if (NPART > 0) {
          if (grp >= 1 && grp <= NPART) {
#pragma unroll
            for (int o = 0; o < DOUT; ++o)
              st_shared_f32(xb + (uint32_t)((((grp - 1) * TILE_M + row) * DOUT + o) * 4), dot[o]);
          }
          epi_bar_sync_n<EPI_THREADS>();
        }
        if (grp == 0) {
          wf_n += 1.f;
          const float inv_n = 1.f / wf_n;
          const float* bl = p.b_last + (size_t)wslot * DOUT;
#pragma unroll
          for (int o = 0; o < DOUT; ++o) {
            float y = dot[o];
#pragma unroll
            for (int j = 0; j < NPART; ++j)
              y += ld_shared_f32(xb + (uint32_t)(((j * TILE_M + row) * DOUT + o) * 4));
            y = fmaf(y, BIAS ? 1.f : final_dropout_scale(p), __ldg(bl + o));
            if (p.last_relu) y = fmaxf(y, 0.f);
            member_fold(p, kg, o, y, inv_n, wf_mean[o], wf_m2[o]);
          }
        }
      }

      // ---- tile done: publish (mean, std) / (mean, M2) ------------------------------------------
      if (grp == 0 && grow < p.n) {
#pragma unroll
        for (int o = 0; o < DOUT; ++o) {
          if (o < p.d_out) {
            const int64_t idx = grow * p.d_out + o;
            if (p.splits > 1) {
              p.part_mean[(size_t)split * (size_t)p.n * p.d_out + idx] = wf_mean[o];
              p.part_m2[(size_t)split * (size_t)p.n * p.d_out + idx] = wf_m2[o];
            } else {
              p.out0[idx] = wf_mean[o];
              p.out1[idx] = second_output(p, wf_m2[o], wf_n, idx);
            }
          }
        }
      }
    }
  }

  // both CTAs must be done with each other's shared / tensor memory before either leaves
  __syncwarp();  // the single-lane roles rejoin their warp: cluster barriers are warp-aligned
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, (uint32_t)G::TMEM_COLS);
  }
}

// The bias-in-the-MMA variant is the default where it applies (no live dropout, d_out 1):
// ensemble16x512_1M 14.33 -> 13.58 ms, 0.757 -> 0.798 of the burst peak, same error against the
// oracle (profiles/r02_j_bias_mma_*); see bias_in_mma_enabled() in tc_params.cuh.

template <int H, int DOUT, int NG, bool MC, bool BIAS = false>
int launch_tc2_mc(const TcParams& p, cudaStream_t st) {
  using G = Geo2<H, DOUT, NG, BIAS>;
  auto kern = uq_mlp_tc2_kernel<H, DOUT, NG, MC, BIAS>;
  // per-device launch geometry of this instantiation, queried once (the occupancy query and the
  // attribute call cost tens of microseconds, which shows on millisecond-sized forwards)
  static std::atomic<int> cached_clusters[64];   // zero-initialised; races only repeat the query
  int dev = 0;
  cudaGetDevice(&dev);
  const int64_t units = (int64_t)((p.n_tiles + 1) / 2) * p.splits;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(G::NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = G::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = dev >= 0 && dev < 64 ? cached_clusters[dev].load(std::memory_order_acquire) : 0;
  if (max_clusters == 0) {
    UQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    max_clusters = sms / 2;   // one CTA per SM: the TMEM allocation is per pair
    cfg.gridDim = dim3((unsigned)sms, 1, 1);
    int active = 0;
    if (cudaOccupancyMaxActiveClusters(&active, kern, &cfg) == cudaSuccess && active > 0 &&
        active < max_clusters)
      max_clusters = active;
    (void)cudaGetLastError();
    if (dev >= 0 && dev < 64) cached_clusters[dev].store(max_clusters, std::memory_order_release);
  }
  const int clusters = (int)(units < (int64_t)max_clusters ? units : (int64_t)max_clusters);
  cfg.gridDim = dim3((unsigned)(2 * clusters), 1, 1);
  UQ_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  UQ_LAUNCH_CHECK();
  return UQ_OK;
}

}  // namespace

}  // namespace uq
