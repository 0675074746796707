// fp32-parity mode of the UQ forward ON THE TENSOR CORES: scaled fp16 x 2 split, CTA pairs with
// 64 sample rows per CTA, hidden width 64 .. 512.
//
// north_star asks for the ensemble / Delta-UQ / dropout-off (and injected-mask) paths to match the
// reference's torch CPU arithmetic within 1e-5 relative AND to run as fused tcgen05 GEMMs.  bf16
// (8 significant bits) cannot, TF32 (11) cannot either, and a 3-way bf16 split needs six MMAs per
// K step and 6 bytes of shared memory per activation.  fp16 carries 11 significant bits, so TWO
// pieces  a ~ h1 + h2  (h1 = rn16(a), h2 = rn16(a - h1))  hold 22 bits and
//     a w  ~  h1 g1 + h2 g1 + h1 g2                     (dropped: h2 g2 <= 2^-22 |a w|)
// is THREE tcgen05.mma.kind::f16 per K step into fp32 tensor-memory tiles (per-product error
// <= 3 * 2^-22 = 7e-7, i.e. a few fp32 ulps; main and correction products in separate tiles).  What fp16 lacks is exponent
// range; the kernel supplies it with exact power-of-two scales:
//   * weights: one scale C per (member, layer), max |w C| in [2^13, 2^14)  (tcx_pack);
//   * activations: one scale s per ROW and layer, chosen BEFORE the layer is drained from the
//     Cauchy-Schwarz bound  |z_n| <= max_n |w_n|_2 * |a_prev|_2 + max |b|  on the row's outputs
//     (|a_prev|_2 of the row is a by-product of the previous drain), so that |a s| < 2^14 always:
//     no overflow, and everything within 2^17 of the bound keeps its 22 bits (below that the
//     absolute error is 2^-25 in scaled units, i.e. < 2^-38 of the bound);
//   * the epilogue undoes both with one factor per row:  v = fma(acc, 1 / (s C), bias).
// Everything else is mlp_tc3.cu's scheme: persistent warp-specialised CTA pairs
// (tcgen05.mma.cta_group::2, M = 128: each CTA's [64 x N] slice of D is 128 TMEM lanes x N/2
// columns), split weight stages streamed with cp.async.bulk through an mbarrier ring, in-place
// write-back of the next layer's A operand (two swizzled fp16 pieces), last Linear as a CUDA-core
// dot product on the unrounded fp32 activations, per-row Welford over the members.
//
// Replaces (in UQ_PREC_FP32): EnsembleModel.forward (models.py:99-108), MCDropoutModel.forward
// (:147-163) and the anchored forward behind DeltaUQMLP.forward (:313-341) / PAGERMLP (:396-429).
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"
#include "philox.cuh"
#include "tc_params.cuh"
#include "tc_ptx.cuh"
#include "tc_epilogue.cuh"
#include "tcx_common.cuh"

namespace uq {

namespace {

using namespace tc;
constexpr int SMEM_LIMIT = 232448;             // 227 KB opt-in maximum per CTA
constexpr int ROWS = 64;                       // sample rows per CTA
constexpr int CHUNKX_BYTES = ROWS * 128;       // one activation chunk [64 x 64] fp16
constexpr int NG = 2;                          // epilogue warp groups
constexpr int EPI_THREADS = NG * 128;
constexpr int NUM_THREADS = 64 + EPI_THREADS;

// MMA N of the pair: the largest divisor of H that is a multiple of 64 and <= 256
__host__ __device__ constexpr int x_ntile(int H) {
  return (H % 256 == 0) ? 256 : (H % 192 == 0) ? 192 : (H % 128 == 0) ? 128 : 64;
}

// Two accumulator tiles per N tile.  The tensor core TRUNCATES when it aligns the 16 products of a
// K step to the running sum (measured: with all three products in one tile the outputs of a
// 3 x 512 net sit -3.4e-6 relative below the fp32 reference, 96 steps per layer; FFMA: 1e-8), so
// the two correction products (h2 g1, h1 g2 -- 2^-11 of the main term) accumulate in a tile of
// their own where they lose nothing, the main tile sees K/16 steps instead of 3K/16 (bias
// -1.2e-6), and the epilogue adds the two tiles in fp32.
template <int H, int DOUT>
struct GeoX {
  static_assert(H % 64 == 0 && H >= 64 && H <= 512, "hidden width must be a multiple of 64 <= 512");
  static constexpr int KC = H / CHUNK_K;                 // activation chunks per piece (<= 8)
  static constexpr int NT = x_ntile(H);                  // MMA N of the pair
  static constexpr int NTILES = H / NT;
  static constexpr int CPT = NT / CHUNK_K;               // chunks per N tile
  static constexpr int TCOLS = NT / 2;                   // TMEM columns per N tile
  static constexpr int BPT = TCOLS / 32;                 // 32-column drain blocks per (tile, half)
  static constexpr int BLOCKS_PER_WARP = (NTILES * BPT + NG - 1) / NG;
  static constexpr int ACC_COLS = H / 2;                 // TMEM columns of one accumulator tile set
  static constexpr int NEED_COLS = H;                    // main tiles, then correction tiles
  static constexpr int TMEM_COLS =
      NEED_COLS <= 32 ? 32 : NEED_COLS <= 64 ? 64 : NEED_COLS <= 128 ? 128 : NEED_COLS <= 256 ? 256 : 512;
  static constexpr int STAGE_BYTES = NT * 128;           // whole stage [NT x 64] fp16 in the image
  static constexpr int HALF_BYTES = STAGE_BYTES / 2;
  static constexpr int A_BYTES = 2 * KC * CHUNKX_BYTES;  // h1 chunks, then h2 chunks
  static constexpr int AUX_FLOATS = (DOUT == 1 ? 2 : 1) * H + STAT_FLOATS;
  static constexpr int AUX_BYTES = 2 * AUX_FLOATS * 4;
  static constexpr int XCHG_BYTES = 2 * 3 * ROWS * DOUT * 4;   // [2][3 partial owners][64][DOUT]
  static constexpr int XS_BYTES = ROWS * 64;                   // x stash (K0 <= 32)
  static constexpr int SS_BYTES = 2 * 4 * ROWS * 4;            // [2][4 parts][64] sums of squares
  static constexpr int XI_BYTES = 2 * 2 * ROWS * 4;            // [2][{1/s_x, |x|^2}][64]
  static constexpr int BAR_BYTES = 384;
  static constexpr int MISC_BYTES = 1024 + BAR_BYTES + XCHG_BYTES + XS_BYTES + SS_BYTES + XI_BYTES;
  static constexpr int BUDGET = SMEM_LIMIT - A_BYTES - AUX_BYTES - MISC_BYTES;
  static constexpr int NS_RAW = BUDGET / HALF_BYTES;
  static constexpr int NSTAGES = NS_RAW > 8 ? 8 : NS_RAW;
  static_assert(NSTAGES >= 3, "not enough shared memory for a weight ring");
  static constexpr int SMEM_BYTES = A_BYTES + NSTAGES * HALF_BYTES + AUX_BYTES + MISC_BYTES;
};

// barrier block (byte offsets inside the 384-byte barrier area)
constexpr uint32_t BX_W_FULL = 0;       // 8 x 8 B
constexpr uint32_t BX_W_EMPTY = 64;     // 8 x 8 B
constexpr uint32_t BX_CHUNK = 128;      // 8 x 8 B   leader only: 2 blocks x 2 row halves x 2 CTAs
constexpr uint32_t BX_D_FULL = 256;
constexpr uint32_t BX_X_READY = 264;    //           leader only: 2 warps of each CTA
constexpr uint32_t BX_TMEM_PTR = 272;

// Drain this warp's 32-column blocks of one layer-step: for every N tile j, the blocks b with
// b % NG == grp of column half ch, as 16-column sub-blocks u = 2 i + h.  Main and correction tile
// of a sub-block are loaded together and summed in fp32; the next sub-block's loads are issued
// before this one is processed (single register buffers: the sums free them).
template <int H, int DOUT, bool RELU, bool DROP, bool LAST>
__device__ __forceinline__ void drain_x(const TcParams& p, uint32_t lane_addr, uint32_t a_row,
                                        int rx, int grp, int ch, int lane, uint32_t chunk_bar0,
                                        const float* bias_s, const float* wl_s, const float* wl_g,
                                        const uint32_t (&keepw)[GeoX<H, DOUT>::BLOCKS_PER_WARP],
                                        float rs, float s_out, float (&dot)[DOUT], float& ss) {
  using G = GeoX<H, DOUT>;
  constexpr int NBLK = G::NTILES * G::BPT;             // blocks of this column half, all tiles
  constexpr int TOTAL = G::BLOCKS_PER_WARP;            // upper bound of blocks per warp
  // i-th block of this warp: t = grp + NG i  ->  tile j = t / BPT, block b = t % BPT; only the
  // tail can be missing (t >= NBLK), so "next block exists" is a plain bound check
  auto valid = [&](int i) { return grp + NG * i < NBLK; };
  auto tcol = [&](int i) {
    const int t = grp + NG * i;
    return (uint32_t)((t / G::BPT) * G::TCOLS + 32 * (t % G::BPT));
  };
  auto block_f0 = [&](int i) {
    const int t = grp + NG * i;
    return (t / G::BPT) * G::NT + ch * G::TCOLS + 32 * (t % G::BPT);   // first feature
  };
  uint32_t m[16], c[16];
  auto issue = [&](int u) {
    const uint32_t col = tcol(u >> 1) + 16u * (uint32_t)(u & 1);
    tmem_ld16(lane_addr + col, m);
    tmem_ld16(lane_addr + (uint32_t)G::ACC_COLS + col, c);
  };
  if (valid(0)) issue(0);
#pragma unroll
  for (int u = 0; u < 2 * TOTAL; ++u) {
    const int i = u >> 1, h = u & 1;
    if (!valid(i)) break;
    const int f0 = block_f0(i);
    const int cidx = f0 >> 6;                               // activation chunk
    const int piece0 = ((f0 & 63) >> 3) + 2 * h;
    const uint32_t a1_dst = a_row + (uint32_t)cidx * CHUNKX_BYTES;
    const uint32_t a2_dst = a1_dst + (uint32_t)G::KC * CHUNKX_BYTES;
    const uint32_t keep = DROP ? keepw[i] : 0xffffffffu;
    float a[16];
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 16; ++e) a[e] = __uint_as_float(m[e]) + __uint_as_float(c[e]);
    if (h == 0 || valid(i + 1)) issue(u + 1);
    epi_block_x<H, DOUT, 16, RELU, DROP, LAST>(a, bias_s + f0 + 16 * h, keep >> (16 * h), rs,
                                                s_out, a1_dst, a2_dst, piece0, rx,
                                                wl_s + f0 + 16 * h, wl_g + f0 + 16 * h, dot, ss);
    if (h == 1) {
      // block drained (+ its half of the A chunk rewritten) -> release to the MMA warp
      tc_fence_before();
      if (!LAST) fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(chunk_bar0 + 8 * cidx);
    }
  }
}

// MC: the launch has live dropout; the dropout-free instantiation carries no mask code (registers)
template <int H, int DOUT, bool MC>
__global__ void __launch_bounds__(NUM_THREADS, 1)
uq_mlp_tcx_kernel(const __grid_constant__ TcParams p) {
  using G = GeoX<H, DOUT>;
  constexpr int KC = G::KC, NT = G::NT, NS = G::NSTAGES, NTILES = G::NTILES, CPT = G::CPT;
  constexpr uint32_t STAGE_BYTES = G::STAGE_BYTES, HALF_BYTES = G::HALF_BYTES;
  constexpr int SBASE = G::AUX_FLOATS - STAT_FLOATS;     // layer statistics inside the aux block

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* a_smem = smem;                                  // 2 x KC chunks of 8 KB
  uint8_t* w_smem = smem + G::A_BYTES;                     // NS half-stages
  float* aux_smem = reinterpret_cast<float*>(w_smem + NS * HALF_BYTES);
  uint8_t* bar_smem = reinterpret_cast<uint8_t*>(aux_smem) + G::AUX_BYTES;
  const uint32_t xchg = smem_u32(bar_smem + G::BAR_BYTES);
  const uint32_t xstash = xchg + G::XCHG_BYTES;
  float* ss_smem = reinterpret_cast<float*>(bar_smem + G::BAR_BYTES + G::XCHG_BYTES + G::XS_BYTES);
  float* xi_smem = ss_smem + 2 * 4 * ROWS;
  const uint32_t a_base = smem_u32(a_smem);
  const uint32_t w_base = smem_u32(w_smem);
  const uint32_t bars = smem_u32(bar_smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int n_tiles = (int)((p.n + ROWS - 1) / ROWS);
  const int n_tp = (n_tiles + 1) >> 1;             // tile pairs
  const int n_units = n_tp * p.splits;
  // unit -> (member split, tile pair); split-major keeps one member group's weights L2-resident
  // (see mlp_tc3.cu)
  auto unit_split = [&](int unit) { return p.split_major ? unit / n_tp : unit % p.splits; };
  auto unit_tp = [&](int unit) { return p.split_major ? unit % n_tp : unit / p.splits; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(bars + BX_W_FULL + 8 * s, leader ? 2 : 1);
      mbar_init(bars + BX_W_EMPTY + 8 * s, 1);
    }
    for (int c = 0; c < KC; ++c) mbar_init(bars + BX_CHUNK + 8 * c, 8);
    mbar_init(bars + BX_D_FULL, 1);
    mbar_init(bars + BX_X_READY, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(bars + BX_TMEM_PTR, (uint32_t)G::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(bar_smem + BX_TMEM_PTR);

  if (warp == 0) {
    // ===================================== producer =============================================
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      const size_t member_bytes = (size_t)p.stages_per_member * STAGE_BYTES;
      for (int unit = cluster_id; unit < n_units; unit += n_clusters) {
        const int split = unit_split(unit);
        const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
        const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
        for (int k = mb; k < me; ++k) {
          const uint8_t* src =
              p.image + (p.shared_weights ? 0 : (size_t)(p.member_begin + k) * member_bytes) +
              rank * HALF_BYTES;
          for (int s = 0; s < p.stages_per_member; ++s) {
            mbar_wait(bars + BX_W_EMPTY + 8 * slot, phase ^ 1, p.error_flag, 1);
            mbar_arrive_expect_tx(bars + BX_W_FULL + 8 * slot, HALF_BYTES);
            bulk_g2s(w_base + slot * HALF_BYTES, src, HALF_BYTES, bars + BX_W_FULL + 8 * slot);
            src += STAGE_BYTES;
            if (++slot == NS) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && !leader) {
    // ===================================== peer relay ===========================================
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      const uint32_t full0 = mapa_shared(bars + BX_W_FULL, 0);
      for (int unit = cluster_id; unit < n_units; unit += n_clusters) {
        const int split = unit_split(unit);
        const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
        const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
        const int n_stages = (me - mb) * p.stages_per_member;
        for (int s = 0; s < n_stages; ++s) {
          mbar_wait(bars + BX_W_FULL + 8 * slot, phase, p.error_flag, 6);
          mbar_arrive_cluster(full0 + 8 * slot);
          if (++slot == NS) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader) ===================================
    constexpr uint32_t idesc = make_idesc_f16(2 * ROWS, NT);
    const uint64_t a_desc0 = make_sw128_desc(a_base);
    const uint64_t b_desc0 = make_sw128_desc(w_base);
    // the last K = 16 step of the layer-0 chunk is all zeros (x side AND weight side): one MMA
    // over it with accumulate = 0 clears the correction tile, which layer 0 does not use
    const int k0_steps = p.K0 / 16 - 1;
    uint32_t slot = 0, phase = 0;
    uint32_t g = 0, xm = 0;
    bool w_ready = mbar_try_wait_cluster(bars + BX_W_FULL, 0);
    uint32_t nslot = 0, nphase = 0;
    bool w_ready_next = false;
    auto acquire = [&]() {
      if (!w_ready) mbar_wait_cluster_inline(bars + BX_W_FULL + 8 * slot, phase, p.error_flag, 4);
      tc_fence_after();
      nslot = slot + 1;
      nphase = phase;
      if (nslot == NS) { nslot = 0; nphase ^= 1; }
      w_ready_next = mbar_try_wait_cluster(bars + BX_W_FULL + 8 * nslot, nphase);
    };
    auto release = [&]() {
      if (elect_one()) umma_commit_pair(bars + BX_W_EMPTY + 8 * slot, 3);
      slot = nslot;
      phase = nphase;
      w_ready = w_ready_next;
    };
    // all chunks of N tile nt drained by the previous step's epilogue (its TMEM columns are free)
    auto wait_tile_drained = [&](int nt, uint32_t prev_par) {
      uint32_t ok = 0;
#pragma unroll
      for (int i = 0; i < CPT; ++i)
        ok |= (mbar_try_wait_cluster(bars + BX_CHUNK + 8 * (CPT * nt + i), prev_par) ? 1u : 0u) << i;
#pragma unroll
      for (int i = 0; i < CPT; ++i)
        if (!((ok >> i) & 1u))
          mbar_wait_cluster_inline(bars + BX_CHUNK + 8 * (CPT * nt + i), prev_par, p.error_flag, 3);
    };
    for (int unit = cluster_id; unit < n_units; unit += n_clusters) {
      const int split = unit_split(unit);
      const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
      const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
      for (int k = mb; k < me; ++k, ++xm) {
        // ---- layer 0: A = [x1 | x2 | x1] in chunk 0 of the h1 piece -----------------------------
        {
          const uint32_t prev_par = (g - 1) & 1;
          mbar_wait_cluster(bars + BX_X_READY, xm & 1, p.error_flag, 2);
#pragma unroll 1
          for (int nt = 0; nt < NTILES; ++nt) {
            if (g != 0) wait_tile_drained(nt, prev_par);
            if (!w_ready) w_ready = mbar_try_wait_cluster(bars + BX_W_FULL + 8 * slot, phase);
            acquire();
            if (elect_one()) {
              const uint64_t bd = b_desc0 + (uint64_t)((slot * HALF_BYTES) >> 4);
              for (int ks = 0; ks < k0_steps; ++ks)
                umma_bf16_pair(tmem_base + nt * G::TCOLS, a_desc0 + 2 * ks, bd + 2 * ks, idesc,
                               ks > 0 ? 1u : 0u);
              umma_bf16_pair(tmem_base + (uint32_t)G::ACC_COLS + nt * G::TCOLS,
                             a_desc0 + 2 * k0_steps, bd + 2 * k0_steps, idesc, 0u);
            }
            release();
          }
          if (elect_one()) umma_commit_pair(bars + BX_D_FULL, 3);
          ++g;
        }
        // ---- hidden layers: per K chunk  g1 x (h1, h2), then g2 x h1 -----------------------------
        for (int l = 1; l < p.L_mma; ++l) {
          const uint32_t prev_par = (g - 1) & 1;
          wait_tile_drained(0, prev_par);   // tile 0: TMEM columns free, A chunks 0..CPT-1 written
          if (!w_ready) w_ready = mbar_try_wait_cluster(bars + BX_W_FULL + 8 * slot, phase);
          bool c_ready = (CPT < KC) ? mbar_try_wait_cluster(bars + BX_CHUNK + 8 * CPT, prev_par)
                                    : true;
#pragma unroll 1
          for (int nt = 0; nt < NTILES; ++nt) {
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) {
              if (nt == 0 && kc >= CPT) {  // the first pass over K meets the chunks as they drain
                if (!c_ready)
                  mbar_wait_cluster_inline(bars + BX_CHUNK + 8 * kc, prev_par, p.error_flag, 3);
                if (kc + 1 < KC)
                  c_ready = mbar_try_wait_cluster(bars + BX_CHUNK + 8 * (kc + 1), prev_par);
              }
              const uint64_t ad1 = a_desc0 + (uint64_t)((kc * CHUNKX_BYTES) >> 4);
              const uint64_t ad2 = a_desc0 + (uint64_t)(((KC + kc) * CHUNKX_BYTES) >> 4);
              const uint32_t d_main = tmem_base + nt * G::TCOLS;
              const uint32_t d_corr = d_main + (uint32_t)G::ACC_COLS;
              acquire();   // g1 stage
              if (elect_one()) {
                const uint64_t bd = b_desc0 + (uint64_t)((slot * HALF_BYTES) >> 4);
#pragma unroll
                for (int ks = 0; ks < CHUNK_K / 16; ++ks)
                  umma_bf16_pair(d_main, ad1 + 2 * ks, bd + 2 * ks, idesc,
                                 (kc > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < CHUNK_K / 16; ++ks)
                  umma_bf16_pair(d_corr, ad2 + 2 * ks, bd + 2 * ks, idesc,
                                 (kc > 0 || ks > 0) ? 1u : 0u);
              }
              release();
              acquire();   // g2 stage
              if (elect_one()) {
                const uint64_t bd = b_desc0 + (uint64_t)((slot * HALF_BYTES) >> 4);
#pragma unroll
                for (int ks = 0; ks < CHUNK_K / 16; ++ks)
                  umma_bf16_pair(d_corr, ad1 + 2 * ks, bd + 2 * ks, idesc, 1u);
              }
              release();
            }
          }
          if (elect_one()) umma_commit_pair(bars + BX_D_FULL, 3);
          ++g;
        }
      }
    }
  } else {
    // ===================================== epilogue =============================================
    const int ew = warp - 2;
    const int et = threadIdx.x - 64;
    const int q = warp & 3;              // TMEM lane quarter
    const int grp = ew >> 2;             // drain blocks b with b % 2 == grp
    const int rh = q & 1, ch = q >> 1;   // row half / column half inside an N tile
    const int row = rh * 32 + lane;      // row of this CTA's 64-row tile
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t a_row = a_base + (row >> 3) * 1024 + (row & 7) * 128;
    const int rx = row & 7;
    const bool x_owner = (ch == 0 && grp == 0);   // writes x and owns the row's Welford state
    const int part_id = ch * 2 + grp;             // 0 = owner, 1..3 = partial holders
    const uint32_t chunk_bar0 = mapa_shared(bars + BX_CHUNK, 0);
    const uint32_t xready_bar = mapa_shared(bars + BX_X_READY, 0);
    uint32_t g = 0, mcount = 0, ucount = 0;
    constexpr int AUX_PER_THREAD = (G::AUX_FLOATS + EPI_THREADS - 1) / EPI_THREADS;

    const bool use_stash = p.K0 <= 32;
    // x of tile `tile` -> stash (or chunk 0), and the row's 1/s_x, |x|^2 -> xi_smem[par]
    auto build_x = [&](int tile, bool to_stash, int par) {
      if (x_owner)
        build_x_row_x(p, (int64_t)tile * ROWS + row, to_stash, xstash + (uint32_t)(row << 4),
                      (uint32_t)(ROWS << 4), a_row, rx, xi_smem + (par * 2 + 0) * ROWS + row,
                      xi_smem + (par * 2 + 1) * ROWS + row);
    };
    auto publish_x = [&](int tile, int par) {
      if (x_owner) {
        if (use_stash) {
          for (int piece = 0; piece < p.K0 / 8; ++piece) {
            uint32_t a, b, c, d;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
                         : "r"(xstash + (uint32_t)((piece * ROWS + row) << 4)));
            st_shared_v4(a_row + (uint32_t)((piece ^ rx) << 4), a, b, c, d);
          }
        } else {
          build_x(tile, false, par);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(xready_bar);
      }
    };

    float aux_pf[AUX_PER_THREAD];
    auto aux_prefetch = [&](int member_global, int l) {
      const int wslot = p.shared_weights ? 0 : member_global;
      const bool last = (l == p.L_mma - 1);
      const float* bias =
          p.bias[l] + (size_t)((l == 0 && p.bias0_per_member) ? member_global : wslot) * H;
      const float* wl = p.w_last + (size_t)wslot * DOUT * H;
      const float* st = p.lstats + ((size_t)wslot * p.L_mma + l) * STAT_FLOATS;
#pragma unroll
      for (int j = 0; j < AUX_PER_THREAD; ++j) {
        const int i = et + j * EPI_THREADS;
        float v = 0.f;
        if (i < H) v = __ldg(bias + i);
        else if (DOUT == 1 && i < 2 * H) { if (last) v = __ldg(wl + (i - H)); }
        else if (i >= SBASE && i < SBASE + STAT_FLOATS) {
          v = __ldg(st + (i - SBASE));
          if (i == SBASE + 2 && l == 0 && p.bmax0) v = __ldg(p.bmax0 + member_global);
        }
        aux_pf[j] = v;
      }
    };

    bool first_step = true;
    for (int unit = cluster_id; unit < n_units; unit += n_clusters, ++ucount) {
      const int tile = 2 * unit_tp(unit) + (int)rank, split = unit_split(unit);
      const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
      const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
      const int64_t grow = (int64_t)tile * ROWS + row;
      const int xpar = (int)(ucount & 1);

      float wf_n = 0.f, wf_mean[DOUT], wf_m2[DOUT];
#pragma unroll
      for (int o = 0; o < DOUT; ++o) wf_mean[o] = 0.f, wf_m2[o] = 0.f;

      if (first_step) {
        if (use_stash) build_x(tile, true, xpar);
        publish_x(tile, xpar);
        aux_prefetch(p.member_begin + mb, 0);
        first_step = false;
      }

      for (int k = mb; k < me; ++k, ++mcount) {
        const int kg = p.member_begin + k;
        const int wslot = p.shared_weights ? 0 : kg;
        float dot[DOUT];
#pragma unroll
        for (int o = 0; o < DOUT; ++o) dot[o] = 0.f;
        int drop_ord = 0;
        const uint8_t* mask_layer = p.masks;
        float a_inv = 1.f;     // 1 / (row scale of the A operand feeding the current layer)

        int nk = k + 1, ntile = tile;
        bool have_next = true;
        if (nk >= me) {
          const int nunit = unit + n_clusters;
          have_next = nunit < n_units;
          ntile = 2 * unit_tp(nunit) + (int)rank;
          nk = (int)(((int64_t)p.member_count * unit_split(nunit)) / p.splits);
        }
        const bool next_unit = (k + 1 >= me);   // the next member belongs to the next unit

        for (int l = 0; l < p.L_mma; ++l, ++g) {
          const bool last = (l == p.L_mma - 1);
          const bool relu = (p.relu_mask >> l) & 1u;
          const bool has_drop = (p.dropout_mask >> l) & 1u;
          const int drop = (MC && has_drop) ? p.drop_mode : 0;

          float* aux = aux_smem + (g & 1) * G::AUX_FLOATS;
#pragma unroll
          for (int j = 0; j < AUX_PER_THREAD; ++j) {
            const int i = et + j * EPI_THREADS;
            if (i < G::AUX_FLOATS) aux[i] = aux_pf[j];
          }
          epi_bar_sync_n<EPI_THREADS>();
          if (!last) aux_prefetch(kg, l + 1);
          else if (have_next) aux_prefetch(p.member_begin + nk, 0);
          // refresh the x stash (and the row's x statistics) of the NEXT tile while the last
          // layer's MMAs run; it goes to the other parity slot of xi_smem
          if (last && have_next && next_unit && use_stash) build_x(ntile, true, xpar ^ 1);

          // ---- row scales of this step ------------------------------------------------------
          // in-norm of the row: |x|^2 (layer 0) or the four partial sums of squares the previous
          // drain left; bound on |output| -> power-of-two scale of what this drain stores
          const float in_scale =
              (l > 0 && ((p.dropout_mask >> (l - 1)) & 1u) && p.drop_mode) ? p.drop_scale : 1.f;
          float n2;
          if (l == 0) {
            n2 = xi_smem[(xpar * 2 + 1) * ROWS + row];
            a_inv = xi_smem[(xpar * 2 + 0) * ROWS + row];
          } else {
            const float* sp = ss_smem + ((g - 1) & 1) * 4 * ROWS + row;
            n2 = (sp[0] + sp[ROWS]) + (sp[2 * ROWS] + sp[3 * ROWS]);
          }
          const float rs = a_inv * aux[SBASE + 0] * in_scale;
          float s_out = 1.f, s_inv = 1.f;
          if (!last) {
            const float bound = fmaf(aux[SBASE + 1] * in_scale, sqrtf(n2), aux[SBASE + 2]);
            pow2_scale(bound, s_out, s_inv);
          }

          // keep masks of this step (they do not depend on the activations) while the layer's
          // MMAs still run: word i = the warp's i-th block (see drain_x)
          uint32_t keepw[G::BLOCKS_PER_WARP];
          if (drop) {
#pragma unroll
            for (int i = 0; i < G::BLOCKS_PER_WARP; ++i) {
              const int t = grp + NG * i;
              if (t < G::NTILES * G::BPT)
                keepw[i] = keep_bits32(p, drop, kg, drop_ord, grow,
                                       (t / G::BPT) * G::NT + ch * G::TCOLS + 32 * (t % G::BPT),
                                       mask_layer, H);
            }
          }
          if (lane == 0) mbar_wait(bars + BX_D_FULL, g & 1, p.error_flag, 5);
          __syncwarp();
          tc_fence_after();
          if (last && have_next) publish_x(ntile, next_unit ? (xpar ^ 1) : xpar);

          const float* wl_g = p.w_last + (size_t)wslot * DOUT * H;
          float ss = 0.f;
#define UQ_DRAINX(R, D, L)                                                                        \
  drain_x<H, DOUT, R, (D) && MC, L>(p, lane_addr, a_row, rx, grp, ch, lane, chunk_bar0, aux, aux + H, wl_g, \
                            keepw, rs, s_out, dot, ss)
          if (last) {
            if (relu) { if (drop) UQ_DRAINX(true, true, true); else UQ_DRAINX(true, false, true); }
            else { if (drop) UQ_DRAINX(false, true, true); else UQ_DRAINX(false, false, true); }
          } else {
            if (relu) { if (drop) UQ_DRAINX(true, true, false); else UQ_DRAINX(true, false, false); }
            else { if (drop) UQ_DRAINX(false, true, false); else UQ_DRAINX(false, false, false); }
          }
#undef UQ_DRAINX
          if (!last) {
            ss_smem[((g & 1) * 4 + part_id) * ROWS + row] = ss;
            a_inv = s_inv;
          }
          if (has_drop) {
            if (p.masks) mask_layer += (size_t)p.total_members * (size_t)p.n * (size_t)H;
            ++drop_ord;
          }
        }

        // ---- combine the four partial dot products of a row, then Welford ----------------------
        const uint32_t xb = xchg + (uint32_t)((mcount & 1) * 3 * ROWS * DOUT * 4);
        if (part_id != 0) {
#pragma unroll
          for (int o = 0; o < DOUT; ++o)
            st_shared_f32(xb + (uint32_t)((((part_id - 1) * ROWS + row) * DOUT + o) * 4), dot[o]);
        }
        epi_bar_sync_n<EPI_THREADS>();
        if (x_owner) {
          wf_n += 1.f;
          const float inv_n = 1.f / wf_n;
          const float* bl = p.b_last + (size_t)wslot * DOUT;
          const float out_scale = final_dropout_scale(p);
#pragma unroll
          for (int o = 0; o < DOUT; ++o) {
            float y = dot[o];
#pragma unroll
            for (int j = 0; j < 3; ++j)
              y += ld_shared_f32(xb + (uint32_t)(((j * ROWS + row) * DOUT + o) * 4));
            y = fmaf(y, out_scale, __ldg(bl + o));
            if (p.last_relu) y = fmaxf(y, 0.f);
            member_fold(p, kg, o, y, inv_n, wf_mean[o], wf_m2[o]);
          }
        }
      }

      if (x_owner && grow < p.n) {
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

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, (uint32_t)G::TMEM_COLS);
  }
}

template <int H, int DOUT, bool MC>
int launch_tcx(const TcParams& p, cudaStream_t st) {
  using G = GeoX<H, DOUT>;
  auto kern = uq_mlp_tcx_kernel<H, DOUT, MC>;
  // per-device launch geometry of this instantiation, queried once
  static std::atomic<int> cached_clusters[64];
  int dev = 0;
  cudaGetDevice(&dev);
  const int64_t n_tiles = (p.n + ROWS - 1) / ROWS;
  const int64_t units = ((n_tiles + 1) / 2) * p.splits;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
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

template <int DOUT>
int dispatch_hx(int H, const TcParams& p, cudaStream_t st) {
  const bool mc = p.drop_mode != 0 && p.dropout_mask != 0;
  switch (H) {
    case 64: return mc ? launch_tcx<64, DOUT, true>(p, st) : launch_tcx<64, DOUT, false>(p, st);
    case 128: return mc ? launch_tcx<128, DOUT, true>(p, st) : launch_tcx<128, DOUT, false>(p, st);
    case 192: return mc ? launch_tcx<192, DOUT, true>(p, st) : launch_tcx<192, DOUT, false>(p, st);
    case 256: return mc ? launch_tcx<256, DOUT, true>(p, st) : launch_tcx<256, DOUT, false>(p, st);
    case 320: return mc ? launch_tcx<320, DOUT, true>(p, st) : launch_tcx<320, DOUT, false>(p, st);
    case 384: return mc ? launch_tcx<384, DOUT, true>(p, st) : launch_tcx<384, DOUT, false>(p, st);
    case 448: return mc ? launch_tcx<448, DOUT, true>(p, st) : launch_tcx<448, DOUT, false>(p, st);
    case 512: return mc ? launch_tcx<512, DOUT, true>(p, st) : launch_tcx<512, DOUT, false>(p, st);
  }
  set_error("fp32 split kernel: unsupported hidden width %d", H);
  return UQ_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// weight image of the split mode
// ------------------------------------------------------------------------------------------------
// {1/C, max_n |w_n|_2, max |bias|, C} of one (member, layer): w' = w * alpha (eval-BN folded),
// C = the power of two with max |w' C| in [2^13, 2^14).  One block per (member, layer).
__global__ void layer_stats_kernel(const float* __restrict__ w, const float* __restrict__ alpha,
                                   const float* __restrict__ bias_folded, int out, int in, int ld,
                                   int64_t w_member_stride, int n_layers_stats, int layer,
                                   float* __restrict__ stats) {
  const int k = blockIdx.x;
  w += (size_t)k * w_member_stride;
  if (alpha) alpha += (size_t)k * out;
  bias_folded += (size_t)k * out;
  float amax = 0.f, nmax = 0.f, bmax = 0.f;
  for (int n = threadIdx.x; n < out; n += blockDim.x) {
    const float sc = alpha ? alpha[n] : 1.0f;
    float s2 = 0.f;
    for (int i = 0; i < in; ++i) {
      const float f = w[(int64_t)n * ld + i] * sc;
      amax = fmaxf(amax, fabsf(f));
      s2 = fmaf(f, f, s2);
    }
    nmax = fmaxf(nmax, sqrtf(s2));
    bmax = fmaxf(bmax, fabsf(bias_folded[n]));
  }
  __shared__ float red[3][256];
  red[0][threadIdx.x] = amax;
  red[1][threadIdx.x] = nmax;
  red[2][threadIdx.x] = bmax;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s)
      for (int j = 0; j < 3; ++j)
        red[j][threadIdx.x] = fmaxf(red[j][threadIdx.x], red[j][threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // C = 2^(13 - floor(log2 amax)); a zero / non-finite matrix keeps C = 1
    const int ex = (int)((__float_as_uint(red[0][0]) >> 23) & 0xffu);   // amax in [2^(ex-127), ..)
    int ce = 127 + 13 - (ex - 127);
    ce = ce < 2 ? 2 : ce > 252 ? 252 : ce;
    if (ex == 0 || ex == 255) ce = 127;
    float* o = stats + ((size_t)k * n_layers_stats + layer) * STAT_FLOATS;
    o[0] = __uint_as_float((uint32_t)(254 - ce) << 23);
    o[1] = red[1][0] * 1.0001f;   // the bound must hold with the rounding of the norm itself
    o[2] = red[2][0];
    o[3] = __uint_as_float((uint32_t)ce << 23);
  }
}

// One member's stages, in the order the MMA warp consumes them:
//   layer 0: NTILES stages, K0 real columns [g1 | g1 | g2];
//   layer l >= 1: for every N tile, for every K chunk: the g1 stage, then the g2 stage.
__global__ void pack_image_x_kernel(__half* __restrict__ image, const float* __restrict__ w,
                                    const float* __restrict__ alpha,
                                    const float* __restrict__ stats, int layer, int in, int ld,
                                    int n_tile, int n_tiles, int KC, int K0, size_t stage_elems,
                                    int stage_base) {
  const int stages = (layer == 0) ? n_tiles : n_tiles * KC * 2;
  const int64_t total = (int64_t)stages * n_tile * 8;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int piece = (int)(i & 7);
  const int r = (int)((i >> 3) % n_tile);
  const int s = (int)((i >> 3) / n_tile);
  const int nt = (layer == 0) ? s : s / (2 * KC);
  const int kc = (layer == 0) ? 0 : (s >> 1) % KC;
  const int lo = (layer == 0) ? 0 : (s & 1);
  const int n = nt * n_tile + r;                 // output feature
  const float sc = alpha ? alpha[n] : 1.0f;
  const float C = stats[layer * STAT_FLOATS + 3];
  __half vals[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int col = piece * 8 + e;
    float out = 0.f;
    if (layer == 0) {
      const int seg = col / in, ii = col - seg * in;
      if (seg < 3 && col < K0) {
        const float f = (w[(int64_t)n * ld + ii] * sc) * C;
        const float g1 = __half2float(__float2half_rn(f));
        out = (seg == 2) ? (f - g1) : g1;        // [g1 | g1 | g2]
      }
    } else {
      const float f = (w[(int64_t)n * ld + kc * 64 + col] * sc) * C;
      const float g1 = __half2float(__float2half_rn(f));
      out = lo ? (f - g1) : g1;
    }
    vals[e] = __float2half_rn(out);
  }
  uint8_t* dst = reinterpret_cast<uint8_t*>(image + (size_t)(stage_base + s) * stage_elems) +
                 sw128_offset(r, piece);
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(vals);
}

}  // namespace

bool tcx_supported(int hidden) { return hidden % 64 == 0 && hidden >= 64 && hidden <= 512; }
int tcx_rows_per_unit() { return 2 * ROWS; }

int tcx_launch(const tc::TcParams& p, int hidden, int dout_pad, cudaStream_t st) {
  return dout_pad == 1 ? dispatch_hx<1>(hidden, p, st) : dispatch_hx<tc::MAX_DOUT>(hidden, p, st);
}

void tcx_plan(uq_model* m) {
  TcPlan& t = m->tc;
  t.x_ok = false;
  if (!t.ok) { t.x_why_not = t.why_not; return; }
  if (!tcx_supported(t.hidden)) {
    t.x_why_not = "hidden width " + std::to_string(t.hidden) + " > 512";
    return;
  }
  // layer-0 chunk: three input segments plus one all-zero K = 16 step inside 64 columns
  if (t.d_in > 16) {
    t.x_why_not = "input width > 16";
    return;
  }
  t.x_n_tile = x_ntile(t.hidden);
  t.x_stage_bytes = (size_t)t.x_n_tile * 128;
  const int n_tiles = t.hidden / t.x_n_tile;
  t.x_stages_per_member = n_tiles + (t.n_mma_layers - 1) * n_tiles * (t.hidden / 64) * 2;
  t.x_ok = true;
}

// the [H][dx] column differences of tc_pack's PAGER image are recomputed here (tiny)
__global__ void column_diff_x_kernel(const float* __restrict__ w0, int H, int dx, int diff_at,
                                     float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * dx) return;
  const int h = i / dx, j = i % dx;
  out[i] = w0[(int64_t)h * 2 * dx + (dx - diff_at) + j] - w0[(int64_t)h * 2 * dx + diff_at + j];
}

// needs tc_pack to have run (bias_folded)
int tcx_pack(uq_model* m, cudaStream_t st) {
  TcPlan& t = m->tc;
  const int K = m->n_members, H = t.hidden, L = t.n_mma_layers, KC = H / 64;
  const int n_tiles = H / t.x_n_tile;
  const size_t stage_elems = t.x_stage_bytes / sizeof(__half);
  const size_t member_elems = (size_t)t.x_stages_per_member * stage_elems;
  auto dalloc = [&](void** p, size_t bytes) -> int {
    UQ_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    m->allocations.push_back(*p);
    return UQ_OK;
  };
  void *img = nullptr, *stats = nullptr;
  int rc;
  if ((rc = dalloc(&img, member_elems * K * sizeof(__half)))) return rc;
  if ((rc = dalloc(&stats, sizeof(float) * (size_t)K * L * STAT_FLOATS))) return rc;
  t.x_image = static_cast<uint16_t*>(img);
  t.x_stats = static_cast<float*>(stats);
  UQ_CUDA(cudaMemsetAsync(img, 0, member_elems * K * sizeof(__half), st));
  for (int l = 0; l < L; ++l) {
    const Layer& ly = m->layers[l];
    layer_stats_kernel<<<K, 256, 0, st>>>(ly.w, ly.has_bn ? ly.alpha : nullptr, ly.bias_folded,
                                          ly.out, ly.in, ly.in, (int64_t)ly.out * ly.in, L, l,
                                          t.x_stats);
    UQ_LAUNCH_CHECK();
  }
  const int k0 = ((3 * t.d_in + 15) / 16) * 16;   // real columns; the zero step follows them
  for (int k = 0; k < K; ++k) {
    int stage_base = 0;
    for (int l = 0; l < L; ++l) {
      const Layer& ly = m->layers[l];
      const int stages = (l == 0) ? n_tiles : n_tiles * KC * 2;
      const int64_t total = (int64_t)stages * t.x_n_tile * 8;
      pack_image_x_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
          reinterpret_cast<__half*>(t.x_image) + (size_t)k * member_elems,
          ly.w + (size_t)k * ly.out * ly.in, ly.has_bn ? ly.alpha + (size_t)k * ly.out : nullptr,
          t.x_stats + (size_t)k * L * STAT_FLOATS, l, ly.in, ly.in, t.x_n_tile, n_tiles, KC, k0,
          stage_elems, stage_base);
      UQ_LAUNCH_CHECK();
      stage_base += stages;
    }
  }
  if (t.k0_delta > 0) {
    // Delta-UQ / PAGER variants: same stages and statistics except layer 0
    const Layer& l0 = m->layers[0];
    const int dx = l0.in / 2;
    const int k0d = ((3 * dx + 15) / 16) * 16;
    const int64_t total0 = (int64_t)n_tiles * t.x_n_tile * 8;
    void *pd = nullptr, *pp = nullptr, *sd = nullptr, *sp = nullptr, *diff = nullptr;
    if ((rc = dalloc(&pd, member_elems * sizeof(__half)))) return rc;
    if ((rc = dalloc(&pp, member_elems * sizeof(__half)))) return rc;
    if ((rc = dalloc(&sd, sizeof(float) * (size_t)L * STAT_FLOATS))) return rc;
    if ((rc = dalloc(&sp, sizeof(float) * (size_t)L * STAT_FLOATS))) return rc;
    if ((rc = dalloc(&diff, sizeof(float) * (size_t)H * dx))) return rc;
    t.x_image_delta = static_cast<uint16_t*>(pd);
    t.x_image_pager = static_cast<uint16_t*>(pp);
    t.x_stats_delta = static_cast<float*>(sd);
    t.x_stats_pager = static_cast<float*>(sp);
    for (void* dst : {pd, pp}) {
      UQ_CUDA(cudaMemcpyAsync(dst, img, member_elems * sizeof(__half), cudaMemcpyDeviceToDevice,
                              st));
      UQ_CUDA(cudaMemsetAsync(dst, 0, (size_t)n_tiles * t.x_stage_bytes, st));
    }
    for (void* dst : {sd, sp})
      UQ_CUDA(cudaMemcpyAsync(dst, stats, sizeof(float) * (size_t)L * STAT_FLOATS,
                              cudaMemcpyDeviceToDevice, st));
    // Delta-UQ: the dx columns of W0 that multiply x (the half that sees the difference)
    const int diff_at = m->anchor_first ? dx : 0;
    layer_stats_kernel<<<1, 256, 0, st>>>(l0.w + diff_at, l0.has_bn ? l0.alpha : nullptr,
                                          l0.bias_folded, l0.out, dx, l0.in, 0, L, 0,
                                          t.x_stats_delta);
    UQ_LAUNCH_CHECK();
    pack_image_x_kernel<<<(unsigned)((total0 + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<__half*>(pd), l0.w + diff_at, l0.has_bn ? l0.alpha : nullptr,
        t.x_stats_delta, 0, dx, l0.in, t.x_n_tile, n_tiles, KC, k0d, stage_elems, 0);
    UQ_LAUNCH_CHECK();
    // PAGER: the column differences W0[:, dx:] - W0[:, :dx]
    column_diff_x_kernel<<<(H * dx + 255) / 256, 256, 0, st>>>(l0.w, H, dx, diff_at,
                                                              static_cast<float*>(diff));
    UQ_LAUNCH_CHECK();
    layer_stats_kernel<<<1, 256, 0, st>>>(static_cast<const float*>(diff),
                                          l0.has_bn ? l0.alpha : nullptr, l0.bias_folded, l0.out,
                                          dx, dx, 0, L, 0, t.x_stats_pager);
    UQ_LAUNCH_CHECK();
    pack_image_x_kernel<<<(unsigned)((total0 + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<__half*>(pp), static_cast<const float*>(diff),
        l0.has_bn ? l0.alpha : nullptr, t.x_stats_pager, 0, dx, dx, t.x_n_tile, n_tiles, KC, k0d,
        stage_elems, 0);
    UQ_LAUNCH_CHECK();
  }
  return UQ_OK;
}

}  // namespace uq
