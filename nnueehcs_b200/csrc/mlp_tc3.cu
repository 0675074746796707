// bf16 tcgen05 path of the UQ forward for WIDE nets (hidden width 768 / 1024): CTA pairs with
// 64 sample rows per CTA.
//
// A [128 x H] bf16 activation tile no longer fits next to a weight ring in one SM's shared memory
// once H > 512 (256 KB at H = 1024), and a [128 x H] fp32 accumulator needs H > 512 tensor-memory
// columns.  So here each CTA of the pair owns 64 rows: tcgen05.mma.cta_group::2 with M = 128
// gives every CTA a [64 x N] slice of D which the hardware stores as 128 TMEM lanes x N/2
// columns -- lanes 0..63 hold columns [0, N/2) of rows 0..63, lanes 64..127 hold columns
// [N/2, N) of the same rows (measured with tools/microbench/tmem_layout.cu).  H = 1024 is four
// N = 256 tiles = 512 TMEM columns, and the A operand is sixteen [64 x 64] bf16 chunks = 128 KB.
//
// Everything else follows mlp_tc2.cu: split weight stages (each CTA streams N/2 rows), the peer
// relay for "stage landed", chunk barriers on the leader, the bias accumulated by the tensor core
// from a bias stage (BIAS; d_out > 1: staged in shared memory one step ahead and added by the
// epilogue), in-place bf16 write-back of the next layer's A operand, last Linear as a CUDA-core dot
// product feeding the per-row Welford.  Per flop this shape streams twice the weight bytes of
// the M = 256 kernel (a stage serves 128 rows instead of 256), which makes it L2 / shared-memory
// bound rather than MMA bound; it exists so that BASELINE.json's 8-layer width-1024 MC-dropout
// configuration runs fused at all.
//
// Epilogue warp (quadrant q = warp % 4, group g): rows 32 (q & 1) + lane, column half q >> 1 of
// the N tiles j with j % 2 == g.
//
// Replaces: MCDropoutModel.forward (models.py:147-163), EnsembleModel.forward (:99-108) and the
// anchored forward behind DeltaUQMLP.forward (:313-341) for hidden widths 768 and 1024.
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
constexpr int ROWS = 64;                       // sample rows per CTA
constexpr int CHUNK3_BYTES = ROWS * 128;       // one activation chunk [64 x 64] bf16
constexpr int NG = 2;                          // epilogue warp groups
constexpr int EPI_THREADS = NG * 128;
constexpr int NUM_THREADS = 64 + EPI_THREADS;

// BIAS (d_out 1): the folded bias is accumulated by the tensor core -- per layer and N tile one bias
// stage in the ring and one extra K = 16 MMA of an all-ones A tile against it -- not added by the
// epilogue; with live dropout the activations are then stored with their 1/(1-p) (see epi_math).
template <int H, int DOUT, bool BIAS = false>
struct Geo3 {
  static_assert(H % 256 == 0 && H >= 256 && H <= 1024, "hidden width must be a multiple of 256");
  static constexpr int KC = H / CHUNK_K;                 // activation chunks (<= 16)
  static constexpr int NT = 256;                         // MMA N of the pair
  static constexpr int NTILES = H / NT;                  // N tiles per layer (<= 4)
  static constexpr int CPT = NT / CHUNK_K;               // chunks per N tile (4)
  static constexpr int TCOLS = NT / 2;                   // TMEM columns per N tile
  static constexpr int TMEM_COLS = H / 2 <= 128 ? 128 : H / 2 <= 256 ? 256 : 512;
  static constexpr int STAGE_BYTES = NT * 128;           // whole stage [NT x 64] bf16 in the image
  static constexpr int HALF_BYTES = STAGE_BYTES / 2;
  static constexpr int A_BYTES = KC * CHUNK3_BYTES;
  static constexpr int WL_OFF = BIAS ? 0 : H;            // w_last inside a step's aux block
  static constexpr int AUX_FLOATS = WL_OFF + (DOUT == 1 ? H : 0);
  static constexpr int AUX_BYTES = 2 * AUX_FLOATS * 4;
  // all-ones A tile of 64 rows: 8 row groups 256 B apart (see make_sw128_const_desc)
  static constexpr int ONES_BYTES = BIAS ? 7 * 256 + 1024 : 0;
  static constexpr int XCHG_BYTES = 2 * 3 * ROWS * DOUT * 4;   // [2][3 partial owners][64][DOUT]
  static constexpr int XS_BYTES = ROWS * 64;                   // x stash (K0 <= 32)
  static constexpr int BAR_BYTES = 384;
  static constexpr int MISC_BYTES = 1024 + BAR_BYTES + XCHG_BYTES + XS_BYTES;
  static constexpr int BUDGET = SMEM_LIMIT - A_BYTES - AUX_BYTES - MISC_BYTES - ONES_BYTES;
  static constexpr int NS_RAW = BUDGET / HALF_BYTES;
  static constexpr int NSTAGES = NS_RAW > 8 ? 8 : NS_RAW;
  static_assert(NSTAGES >= 2, "not enough shared memory for a weight ring");
  static constexpr int SMEM_BYTES =
      A_BYTES + NSTAGES * HALF_BYTES + ONES_BYTES + AUX_BYTES + MISC_BYTES;
};

// barrier block (byte offsets inside the 384-byte barrier area)
constexpr uint32_t B3_W_FULL = 0;       // 8 x 8 B
constexpr uint32_t B3_W_EMPTY = 64;     // 8 x 8 B
constexpr uint32_t B3_CHUNK = 128;      // 16 x 8 B  leader only: 2 warps of each CTA
constexpr uint32_t B3_D_FULL = 256;
constexpr uint32_t B3_X_READY = 264;    //           leader only: 2 warps of each CTA
constexpr uint32_t B3_TMEM_PTR = 272;

// Keep-mask words of this warp's blocks of one layer-step, computed BEFORE the warp waits for the
// layer's MMAs (they do not depend on the activations; see mlp_tc2.cu: KeepWords).  Word
// 4 i + 2 e + b = block b of chunk e of the warp's i-th N tile.
template <int H, int DOUT>
struct KeepWords3 {
  static constexpr int TPW = (Geo3<H, DOUT>::NTILES + NG - 1) / NG;   // N tiles per warp
  uint32_t w[4 * TPW];
};

template <int H, int DOUT>
__device__ __forceinline__ void compute_keep_words3(const TcParams& p, int grp, int ch, int drop,
                                                    int kg, int drop_ord, int64_t grow,
                                                    const uint8_t* mask_layer,
                                                    KeepWords3<H, DOUT>& kw) {
  using G = Geo3<H, DOUT>;
#pragma unroll
  for (int i = 0; i < KeepWords3<H, DOUT>::TPW; ++i) {
    const int j = grp + NG * i;
    if (j < G::NTILES) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col0 = (G::CPT * j + 2 * ch + e) * CHUNK_K;
        kw.w[4 * i + 2 * e] = keep_bits32(p, drop, kg, drop_ord, grow, col0, mask_layer, H);
        kw.w[4 * i + 2 * e + 1] = keep_bits32(p, drop, kg, drop_ord, grow, col0 + 32, mask_layer, H);
      }
    }
  }
}

template <int H, int DOUT, bool RELU, bool DROP, bool LAST, bool BIAS>
__device__ __forceinline__ void drain3(const TcParams& p, uint32_t lane_addr, uint32_t a_row, int rx,
                                       int grp, int ch, int lane, uint32_t chunk_bar0,
                                       const float* bias_s, const float* wl_s, const float* wl_g,
                                       const KeepWords3<H, DOUT>& kw, float in_scale,
                                       float (&dot)[DOUT]) {
  using G = Geo3<H, DOUT>;
  uint32_t acc0[32], acc1[32];
  // this warp's chunks, in order: for j = grp, grp + 2, ...: c = CPT j + 2 ch + {0, 1}
  if (grp < G::NTILES) tmem_ld32(lane_addr + (uint32_t)(grp * G::TCOLS), acc0);
  auto chunk = [&](int j, int e, uint32_t keep0, uint32_t keep1) {
    const int c = G::CPT * j + 2 * ch + e;
    const int col0 = c * CHUNK_K;                                  // feature index
    const uint32_t tcol = (uint32_t)(j * G::TCOLS + e * CHUNK_K);  // TMEM column of this chunk
    const uint32_t a_dst = a_row + (uint32_t)c * CHUNK3_BYTES;
    float4 bv[8];
    if (!BIAS) {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) bv[j4] = reinterpret_cast<const float4*>(bias_s + col0)[j4];
    }
    tmem_ld_wait();
    tmem_ld32(lane_addr + tcol + 32, acc1);
    epi_block2<H, DOUT, 32, RELU, DROP, LAST, BIAS>(acc0, bv, keep0, in_scale, a_dst, 0, rx,
                                                    wl_s + col0, wl_g + col0, dot);
    if (!BIAS) {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        bv[j4] = reinterpret_cast<const float4*>(bias_s + col0 + 32)[j4];
    }
    tmem_ld_wait();
    // next chunk of this warp: e == 0 -> same tile, next 64 TMEM columns; else tile j + NG
    if (e == 0) tmem_ld32(lane_addr + tcol + CHUNK_K, acc0);
    else if (j + NG < G::NTILES) tmem_ld32(lane_addr + (uint32_t)((j + NG) * G::TCOLS), acc0);
    epi_block2<H, DOUT, 32, RELU, DROP, LAST, BIAS>(acc1, bv, keep1, in_scale, a_dst, 4, rx,
                                                    wl_s + col0 + 32, wl_g + col0 + 32, dot);
    tc_fence_before();
    if (!LAST) fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(chunk_bar0 + 8 * c);
  };
  if (DROP) {   // unrolled: the precomputed keep words are indexed statically
#pragma unroll
    for (int i = 0; i < KeepWords3<H, DOUT>::TPW; ++i) {
      const int j = grp + NG * i;
      if (j < G::NTILES) {
        chunk(j, 0, kw.w[4 * i], kw.w[4 * i + 1]);
        chunk(j, 1, kw.w[4 * i + 2], kw.w[4 * i + 3]);
      }
    }
  } else {
#pragma unroll 1
    for (int j = grp; j < G::NTILES; j += NG) {
#pragma unroll 1
      for (int e = 0; e < 2; ++e) chunk(j, e, 0xffffffffu, 0xffffffffu);
    }
  }
}

template <int H, int DOUT, bool BIAS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
uq_mlp_tc3_kernel(const __grid_constant__ TcParams p) {
  static_assert(!BIAS || DOUT == 1, "bias-in-the-MMA variant: d_out 1");
  using G = Geo3<H, DOUT, BIAS>;
  constexpr int KC = G::KC, NT = G::NT, NS = G::NSTAGES, NTILES = G::NTILES, CPT = G::CPT;
  constexpr uint32_t STAGE_BYTES = G::STAGE_BYTES, HALF_BYTES = G::HALF_BYTES;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* a_smem = smem;                                  // KC chunks of 8 KB
  uint8_t* w_smem = smem + G::A_BYTES;                     // NS half-stages
  uint8_t* ones_smem = w_smem + NS * HALF_BYTES;           // BIAS: constant-1 A tile (bf16)
  float* aux_smem = reinterpret_cast<float*>(ones_smem + G::ONES_BYTES);
  uint8_t* bar_smem = reinterpret_cast<uint8_t*>(aux_smem) + G::AUX_BYTES;
  const uint32_t xchg = smem_u32(bar_smem + G::BAR_BYTES);
  const uint32_t xstash = xchg + G::XCHG_BYTES;
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
  // unit -> (member split, tile pair).  Tile-major (default): a cluster's consecutive units walk
  // the splits of one tile pair.  Split-major (p.split_major): all clusters work on the tile pairs
  // of split 0 first, then of split 1, ... so that only one split's weights are live in L2 at a
  // time -- 8 members x 14.7 MB of a 7 x 1024 ensemble do not fit the 2 x 63 MB L2 and were
  // re-read from HBM every round (2.86 GB of DRAM reads per launch against 125 MB of weights).
  auto unit_split = [&](int unit) { return p.split_major ? unit / n_tp : unit % p.splits; };
  auto unit_tp = [&](int unit) { return p.split_major ? unit % n_tp : unit / p.splits; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(bars + B3_W_FULL + 8 * s, leader ? 2 : 1);
      mbar_init(bars + B3_W_EMPTY + 8 * s, 1);
    }
    for (int c = 0; c < KC; ++c) mbar_init(bars + B3_CHUNK + 8 * c, 4);
    mbar_init(bars + B3_D_FULL, 1);
    mbar_init(bars + B3_X_READY, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(bars + B3_TMEM_PTR, (uint32_t)G::TMEM_COLS);
  if (BIAS) {   // the all-ones A tile of the bias MMAs (read by the tensor core: async proxy)
    for (int i = threadIdx.x; i < G::ONES_BYTES / 4; i += NUM_THREADS)
      reinterpret_cast<uint32_t*>(ones_smem)[i] = 0x3F803F80u;   // bf16 (1.0, 1.0)
    fence_proxy_async_smem();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(bar_smem + B3_TMEM_PTR);

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
          auto load = [&](const uint8_t* from) {
            mbar_wait(bars + B3_W_EMPTY + 8 * slot, phase ^ 1, p.error_flag, 1);
            mbar_arrive_expect_tx(bars + B3_W_FULL + 8 * slot, HALF_BYTES);
            bulk_g2s(w_base + slot * HALF_BYTES, from, HALF_BYTES, bars + B3_W_FULL + 8 * slot);
            if (++slot == NS) { slot = 0; phase ^= 1; }
          };
          if (BIAS) {   // per (layer, N tile): its weight stages, then its bias stage
            const uint8_t* bsrc =
                p.bias_image +
                (p.shared_weights ? 0 : (size_t)(p.member_begin + k) * p.L_mma * NTILES * STAGE_BYTES) +
                rank * HALF_BYTES;
            for (int l = 0; l < p.L_mma; ++l)
              for (int nt = 0; nt < NTILES; ++nt, bsrc += STAGE_BYTES) {
                for (int s = 0; s < (l == 0 ? 1 : KC); ++s, src += STAGE_BYTES) load(src);
                if (l == 0 && p.bias0_image != nullptr)   // per-anchor layer-0 bias
                  load(p.bias0_image + ((size_t)(p.member_begin + k) * NTILES + nt) * STAGE_BYTES +
                       rank * HALF_BYTES);
                else
                  load(bsrc);
              }
          } else {
            for (int s = 0; s < p.stages_per_member; ++s, src += STAGE_BYTES) load(src);
          }
        }
      }
    }
  } else if (warp == 1 && !leader) {
    // ===================================== peer relay ===========================================
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      const uint32_t full0 = mapa_shared(bars + B3_W_FULL, 0);
      for (int unit = cluster_id; unit < n_units; unit += n_clusters) {
        const int split = unit_split(unit);
        const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
        const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
        const int n_stages = (me - mb) * (p.stages_per_member + (BIAS ? p.L_mma * NTILES : 0));
        for (int s = 0; s < n_stages; ++s) {
          mbar_wait(bars + B3_W_FULL + 8 * slot, phase, p.error_flag, 6);
          mbar_arrive_cluster(full0 + 8 * slot);
          if (++slot == NS) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader) ===================================
    constexpr uint32_t idesc = make_idesc_bf16(2 * ROWS, NT);
    const uint64_t a_desc0 = make_sw128_desc(a_base);
    const uint64_t b_desc0 = make_sw128_desc(w_base);
    const int k0_steps = p.K0 / 16;
    uint32_t slot = 0, phase = 0;
    uint32_t g = 0, xm = 0;
    bool w_ready = mbar_try_wait_cluster(bars + B3_W_FULL, 0);
    uint32_t nslot = 0, nphase = 0;
    bool w_ready_next = false;
    auto acquire = [&]() {
      if (!w_ready) mbar_wait_cluster_inline(bars + B3_W_FULL + 8 * slot, phase, p.error_flag, 4);
      tc_fence_after();
      nslot = slot + 1;
      nphase = phase;
      if (nslot == NS) { nslot = 0; nphase ^= 1; }
      w_ready_next = mbar_try_wait_cluster(bars + B3_W_FULL + 8 * nslot, nphase);
    };
    auto release = [&]() {
      if (elect_one()) umma_commit_pair(bars + B3_W_EMPTY + 8 * slot, 3);
      slot = nslot;
      phase = nphase;
      w_ready = w_ready_next;
    };
    const uint64_t ones_desc = make_sw128_const_desc(smem_u32(ones_smem));
    // N tile nt += ones . bias^T: one K = 16 step on the tile's bias stage
    auto bias_mma = [&](int nt) {
      acquire();
      if (elect_one())
        umma_bf16_pair(tmem_base + nt * G::TCOLS, ones_desc,
                       b_desc0 + (uint64_t)((slot * HALF_BYTES) >> 4), idesc, 1u);
      release();
    };
    // all chunks of N tile nt drained by the previous step's epilogue (its TMEM columns are free)
    auto wait_tile_drained = [&](int nt, uint32_t prev_par) {
      uint32_t ok = 0;
#pragma unroll
      for (int i = 0; i < CPT; ++i)
        ok |= (mbar_try_wait_cluster(bars + B3_CHUNK + 8 * (CPT * nt + i), prev_par) ? 1u : 0u) << i;
#pragma unroll
      for (int i = 0; i < CPT; ++i)
        if (!((ok >> i) & 1u))
          mbar_wait_cluster_inline(bars + B3_CHUNK + 8 * (CPT * nt + i), prev_par, p.error_flag, 3);
    };
    for (int unit = cluster_id; unit < n_units; unit += n_clusters) {
      const int split = unit_split(unit);
      const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
      const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
      for (int k = mb; k < me; ++k, ++xm) {
        // ---- layer 0: A = split input rows in chunk 0 -------------------------------------------
        {
          const uint32_t prev_par = (g - 1) & 1;
          mbar_wait_cluster(bars + B3_X_READY, xm & 1, p.error_flag, 2);
#pragma unroll 1
          for (int nt = 0; nt < NTILES; ++nt) {
            if (g != 0) wait_tile_drained(nt, prev_par);
            if (!w_ready) w_ready = mbar_try_wait_cluster(bars + B3_W_FULL + 8 * slot, phase);
            acquire();
            if (elect_one()) {
              const uint64_t bd = b_desc0 + (uint64_t)((slot * HALF_BYTES) >> 4);
              for (int ks = 0; ks < k0_steps; ++ks)
                umma_bf16_pair(tmem_base + nt * G::TCOLS, a_desc0 + 2 * ks, bd + 2 * ks, idesc,
                               ks > 0 ? 1u : 0u);
            }
            release();
            if (BIAS) bias_mma(nt);
          }
          if (elect_one()) umma_commit_pair(bars + B3_D_FULL, 3);
          ++g;
        }
        // ---- hidden layers ----------------------------------------------------------------------
        for (int l = 1; l < p.L_mma; ++l) {
          const uint32_t prev_par = (g - 1) & 1;
          wait_tile_drained(0, prev_par);   // tile 0: TMEM columns free, A chunks 0..CPT-1 written
          if (!w_ready) w_ready = mbar_try_wait_cluster(bars + B3_W_FULL + 8 * slot, phase);
          bool c_ready = mbar_try_wait_cluster(bars + B3_CHUNK + 8 * CPT, prev_par);
#pragma unroll 1
          for (int nt = 0; nt < NTILES; ++nt) {
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) {
              if (nt == 0 && kc >= CPT) {  // the first pass over K meets the chunks as they drain
                if (!c_ready)
                  mbar_wait_cluster_inline(bars + B3_CHUNK + 8 * kc, prev_par, p.error_flag, 3);
                if (kc + 1 < KC)
                  c_ready = mbar_try_wait_cluster(bars + B3_CHUNK + 8 * (kc + 1), prev_par);
              }
              acquire();
              if (elect_one()) {
                const uint64_t ad = a_desc0 + (uint64_t)((kc * CHUNK3_BYTES) >> 4);
                const uint64_t bd = b_desc0 + (uint64_t)((slot * HALF_BYTES) >> 4);
#pragma unroll
                for (int ks = 0; ks < CHUNK_K / 16; ++ks)
                  umma_bf16_pair(tmem_base + nt * G::TCOLS, ad + 2 * ks, bd + 2 * ks, idesc,
                                 (kc > 0 || ks > 0) ? 1u : 0u);
              }
              release();
            }
            if (BIAS) bias_mma(nt);
          }
          if (elect_one()) umma_commit_pair(bars + B3_D_FULL, 3);
          ++g;
        }
      }
    }
  } else {
    // ===================================== epilogue =============================================
    const int ew = warp - 2;
    const int et = threadIdx.x - 64;
    const int q = warp & 3;              // TMEM lane quarter
    const int grp = ew >> 2;             // N tiles j with j % 2 == grp
    const int rh = q & 1, ch = q >> 1;   // row half / column half inside an N tile
    const int row = rh * 32 + lane;      // row of this CTA's 64-row tile
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t a_row = a_base + (row >> 3) * 1024 + (row & 7) * 128;
    const int rx = row & 7;
    const bool x_owner = (ch == 0 && grp == 0);   // writes x and owns the row's Welford state
    const int part_id = ch * 2 + grp;             // 0 = owner, 1..3 = partial dot holders
    const uint32_t chunk_bar0 = mapa_shared(bars + B3_CHUNK, 0);
    const uint32_t xready_bar = mapa_shared(bars + B3_X_READY, 0);
    uint32_t g = 0, mcount = 0;
    constexpr int AUX_PER_THREAD = (G::AUX_FLOATS + EPI_THREADS - 1) / EPI_THREADS;

    const bool use_stash = p.K0 <= 32;
    auto build_x = [&](int tile, int member_global, bool to_stash) {
      if (x_owner)
        build_x_row(p, (int64_t)tile * ROWS + row, member_global, to_stash,
                    xstash + (uint32_t)(row << 4), (uint32_t)(ROWS << 4), a_row, rx);
    };
    auto publish_x = [&](int tile, int member_global) {
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
          build_x(tile, member_global, false);
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
      const int tile = 2 * unit_tp(unit) + (int)rank, split = unit_split(unit);
      const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
      const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
      const int64_t grow = (int64_t)tile * ROWS + row;

      float wf_n = 0.f, wf_mean[DOUT], wf_m2[DOUT];
#pragma unroll
      for (int o = 0; o < DOUT; ++o) wf_mean[o] = 0.f, wf_m2[o] = 0.f;

      if (first_step) {
        if (use_stash) build_x(tile, p.member_begin + mb, true);
        publish_x(tile, p.member_begin + mb);
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

        int nk = k + 1, ntile = tile;
        bool have_next = true;
        if (nk >= me) {
          const int nunit = unit + n_clusters;
          have_next = nunit < n_units;
          ntile = 2 * unit_tp(nunit) + (int)rank;
          nk = (int)(((int64_t)p.member_count * unit_split(nunit)) / p.splits);
        }

        for (int l = 0; l < p.L_mma; ++l, ++g) {
          const bool last = (l == p.L_mma - 1);
          const bool relu = (p.relu_mask >> l) & 1u;
          const bool has_drop = (p.dropout_mask >> l) & 1u;
          const int drop = has_drop ? p.drop_mode : 0;

          float* aux = aux_smem + (g & 1) * G::AUX_FLOATS;
#pragma unroll
          for (int j = 0; j < AUX_PER_THREAD; ++j) {
            const int i = et + j * EPI_THREADS;
            if (i < G::AUX_FLOATS) aux[i] = aux_pf[j];
          }
          epi_bar_sync_n<EPI_THREADS>();
          if (!last) aux_prefetch(kg, l + 1);
          else if (have_next) aux_prefetch(p.member_begin + nk, 0);
          if (last && have_next && use_stash && (ntile != tile))
            build_x(ntile, p.member_begin + nk, true);

          // keep masks of this step while the layer's MMAs still run
          KeepWords3<H, DOUT> kw;
          if (drop) compute_keep_words3<H, DOUT>(p, grp, ch, drop, kg, drop_ord, grow, mask_layer, kw);
          if (lane == 0) mbar_wait(bars + B3_D_FULL, g & 1, p.error_flag, 5);
          __syncwarp();
          tc_fence_after();
          if (last && have_next) publish_x(ntile, p.member_begin + nk);

          const float* wl_g = p.w_last + (size_t)wslot * DOUT * H;
          // epilogue-bias: 1/(1-p) owed by the previous layer's dropout; bias in the MMA: this layer's
          const float in_scale =
              BIAS ? (drop ? p.drop_scale : 1.f)
                   : (l > 0 && ((p.dropout_mask >> (l - 1)) & 1u) && p.drop_mode) ? p.drop_scale : 1.f;
#define UQ_DRAIN3(R, D, L)                                                                       \
  drain3<H, DOUT, R, D, L, BIAS>(p, lane_addr, a_row, rx, grp, ch, lane, chunk_bar0, aux,        \
                                 aux + G::WL_OFF, wl_g, kw, in_scale, dot)
          if (last) {
            if (relu) { if (drop) UQ_DRAIN3(true, true, true); else UQ_DRAIN3(true, false, true); }
            else { if (drop) UQ_DRAIN3(false, true, true); else UQ_DRAIN3(false, false, true); }
          } else {
            if (relu) { if (drop) UQ_DRAIN3(true, true, false); else UQ_DRAIN3(true, false, false); }
            else { if (drop) UQ_DRAIN3(false, true, false); else UQ_DRAIN3(false, false, false); }
          }
#undef UQ_DRAIN3
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
#pragma unroll
          for (int o = 0; o < DOUT; ++o) {
            float y = dot[o];
#pragma unroll
            for (int j = 0; j < 3; ++j)
              y += ld_shared_f32(xb + (uint32_t)(((j * ROWS + row) * DOUT + o) * 4));
            y = fmaf(y, BIAS ? 1.f : final_dropout_scale(p), __ldg(bl + o));
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

template <int H, int DOUT, bool BIAS>
int launch_tc3(const TcParams& p, cudaStream_t st) {
  using G = Geo3<H, DOUT, BIAS>;
  auto kern = uq_mlp_tc3_kernel<H, DOUT, BIAS>;
  // per-device launch geometry of this instantiation, queried once (the occupancy query and the
  // attribute call cost tens of microseconds, which shows on millisecond-sized forwards)
  static std::atomic<int> cached_clusters[64];   // zero-initialised; races only repeat the query
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
int dispatch_h3(int H, const TcParams& p, cudaStream_t st) {
  const bool bias = DOUT == 1 && p.bias_image != nullptr && bias_in_mma_enabled();
  switch (H) {
    case 768: return bias ? launch_tc3<768, 1, true>(p, st) : launch_tc3<768, DOUT, false>(p, st);
    case 1024: return bias ? launch_tc3<1024, 1, true>(p, st) : launch_tc3<1024, DOUT, false>(p, st);
  }
  set_error("bf16 wide-net kernel: unsupported hidden width %d", H);
  return UQ_ERR_UNSUPPORTED;
}

}  // namespace

bool tc3_supported(int hidden) { return hidden == 768 || hidden == 1024; }
int tc3_rows_per_tile() { return ROWS; }

int tc3_launch(const tc::TcParams& p, int hidden, int dout_pad, cudaStream_t st) {
  return dout_pad == 1 ? dispatch_h3<1>(hidden, p, st) : dispatch_h3<tc::MAX_DOUT>(hidden, p, st);
}

}  // namespace uq
