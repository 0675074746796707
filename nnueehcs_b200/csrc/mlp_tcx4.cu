// fp32-parity split mode (scaled fp16 x 2, see mlp_tcx.cu) for NARROW nets: hidden width 64 / 128,
// one output, CTA pairs with FOUR 64-row tile slots in flight per CTA.
//
// The reference's own architecture (examples/binomial_options/config.yaml:16-54) is 6 x 128, and
// UQ_PREC_FP32 is the wrappers' default precision -- this is the kernel a model built from the
// repo's YAML runs.  At H = 128 the three MMAs of a layer of one 64-row tile are 768 tensor-core
// cycles against ~3000 cycles of  commit -> epilogue wake-up -> drain -> barrier -> MMA resume,
// so one tile per CTA (mlp_tcx.cu) leaves the tensor core idle three quarters of the time.  As in
// mlp_tc4.cu, a narrow net leaves room for more tiles: four slots x (main + correction tile) fill
// the 512 TMEM columns, four slots x two fp16 pieces are 128 KB of shared memory, and the 2 KC
// weight stages of a layer stay in the ring while all four slots use them.  The MMA warp walks
// the slots in order; epilogue warp group g owns the slots t with t % 2 == g, so two slots are
// drained while the tensor core works on the others.
//
// Per slot: D_FULL[t] (layer accumulated, commit multicast), DRAINED[t] (the slot's 4 epilogue warps
// of each CTA are done with its accumulators and have rewritten its A pieces), X_READY[t].
// Inside a slot the layout is mlp_tcx.cu's: tcgen05.mma.cta_group::2 with M = 128 stores a CTA's
// [64 x H] slice of D as 128 lanes x H/2 columns (lanes 64.. = columns H/2..), so warp q of the
// group drains rows 32 (q & 1) + lane, column half q >> 1; the two threads of a row exchange their
// partial sums of squares (row scale of the next layer) and partial dot products (last Linear)
// through shared memory.
//
// Same weight image and layer statistics as mlp_tcx.cu (tcx_pack).
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
constexpr int ROWS = 64;                       // sample rows per CTA and slot
constexpr int CHUNKX_BYTES = ROWS * 128;       // one activation chunk [64 x 64] fp16
constexpr int TS = 4;                          // tile slots per CTA
constexpr int NG = 2;                          // epilogue warp groups
constexpr int OWN = TS / NG;                   // slots per group
constexpr int EPI_THREADS = NG * 128;
constexpr int NUM_THREADS = 64 + EPI_THREADS;

template <int H>
struct GeoX4 {
  static_assert(H == 64 || H == 128, "narrow split kernel: hidden width 64 or 128");
  static constexpr int KC = H / CHUNK_K;                 // 1 or 2
  static constexpr int NT = H;                           // MMA N of the pair
  static constexpr int TCOLS = H / 2;                    // TMEM columns of one accumulator tile
  static constexpr int BPT = TCOLS / 32;                 // 32-column drain blocks per thread and slot
  static constexpr int SLOT_COLS = H;                    // main tile, then correction tile
  static constexpr int TMEM_COLS = TS * H <= 256 ? 256 : 512;
  static constexpr int STAGE_BYTES = NT * 128;
  static constexpr int HALF_BYTES = STAGE_BYTES / 2;
  static constexpr int LAYER_STAGES = 2 * KC;            // g1, g2 of every K chunk
  static constexpr int A_SLOT_BYTES = 2 * KC * CHUNKX_BYTES;   // h1 chunks, then h2 chunks
  static constexpr int A_BYTES = TS * A_SLOT_BYTES;
  static constexpr int AUX_FLOATS = 2 * H + STAT_FLOATS;       // bias, w_last, layer statistics
  static constexpr int AUX_BYTES = 2 * AUX_FLOATS * 4;
  static constexpr int XS_SLOT_BYTES = ROWS * 64;              // x stash (K0 <= 32)
  static constexpr int SS_BYTES = 2 * TS * 2 * ROWS * 4;       // [2][slot][2 parts][64]
  static constexpr int XI_BYTES = 2 * TS * 2 * ROWS * 4;       // [2][slot][{1/s_x, |x|^2}][64]
  static constexpr int DOT_BYTES = 2 * TS * ROWS * 4;          // [2][slot][64] partial dot of ch = 1
  static constexpr int BAR_BYTES = 384;
  static constexpr int MISC_BYTES =
      1024 + BAR_BYTES + TS * XS_SLOT_BYTES + SS_BYTES + XI_BYTES + DOT_BYTES;
  static constexpr int BUDGET = SMEM_LIMIT - A_BYTES - AUX_BYTES - MISC_BYTES;
  static constexpr int NS_RAW = BUDGET / HALF_BYTES;
  static constexpr int NSTAGES = NS_RAW > 8 ? 8 : NS_RAW;
  static_assert(NSTAGES >= 2 * LAYER_STAGES, "the ring must hold two layers of weight stages");
  static constexpr int SMEM_BYTES = A_BYTES + NSTAGES * HALF_BYTES + AUX_BYTES + MISC_BYTES;
};

constexpr uint32_t B4_W_FULL = 0;       // 8 x 8 B
constexpr uint32_t B4_W_EMPTY = 64;     // 8 x 8 B
constexpr uint32_t B4_D_FULL = 128;     // TS x 8 B   commit multicast
constexpr uint32_t B4_DRAINED = 160;    // TS x 8 B   leader only: 4 warps of each CTA
constexpr uint32_t B4_X_READY = 192;    // TS x 8 B   leader only: 2 warps of each CTA
constexpr uint32_t B4_TMEM_PTR = 224;

// drain this thread's blocks of one slot (row 32 (q & 1) + lane, column half ch), as 16-column
// sub-blocks: main and correction tile loaded together and summed in fp32 (mlp_tcx.cu: drain_x)
template <int H, bool RELU, bool DROP, bool LAST>
__device__ __forceinline__ void drain_x4(uint32_t lane_addr, uint32_t a_row, int rx, int ch,
                                         const float* bias_s, const float* wl_s,
                                         const uint32_t (&keepw)[GeoX4<H>::BPT], float rs,
                                         float s_out, float (&dot)[1], float& ss) {
  using G = GeoX4<H>;
  uint32_t m[16], c[16];
  auto issue = [&](int u) {
    const uint32_t col = 32u * (uint32_t)(u >> 1) + 16u * (uint32_t)(u & 1);
    tmem_ld16(lane_addr + col, m);
    tmem_ld16(lane_addr + (uint32_t)G::TCOLS + col, c);
  };
  issue(0);
#pragma unroll
  for (int u = 0; u < 2 * G::BPT; ++u) {
    const int b = u >> 1, h = u & 1;
    const int f0 = ch * G::TCOLS + 32 * b;                  // first feature of the block
    const int cidx = f0 >> 6;                               // activation chunk
    const int piece0 = ((f0 & 63) >> 3) + 2 * h;
    const uint32_t a1_dst = a_row + (uint32_t)cidx * CHUNKX_BYTES;
    const uint32_t a2_dst = a1_dst + (uint32_t)G::KC * CHUNKX_BYTES;
    const uint32_t keep = DROP ? keepw[b] : 0xffffffffu;
    float a[16];
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 16; ++e) a[e] = __uint_as_float(m[e]) + __uint_as_float(c[e]);
    if (u + 1 < 2 * G::BPT) issue(u + 1);
    epi_block_x<H, 1, 16, RELU, DROP, LAST>(a, bias_s + f0 + 16 * h, keep >> (16 * h), rs, s_out,
                                            a1_dst, a2_dst, piece0, rx, wl_s + f0 + 16 * h,
                                            wl_s + f0 + 16 * h, dot, ss);
  }
}

// MC: the launch has live dropout; the dropout-free instantiation carries no mask code (registers)
template <int H, bool MC>
__global__ void __launch_bounds__(NUM_THREADS, 1)
uq_mlp_tcx4_kernel(const __grid_constant__ TcParams p) {
  using G = GeoX4<H>;
  constexpr int KC = G::KC, NT = G::NT, NS = G::NSTAGES;
  constexpr uint32_t STAGE_BYTES = G::STAGE_BYTES, HALF_BYTES = G::HALF_BYTES;
  constexpr int SBASE = G::AUX_FLOATS - STAT_FLOATS;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* a_smem = smem;                                  // TS slots x 2 pieces x KC chunks
  uint8_t* w_smem = smem + G::A_BYTES;                     // NS half-stages
  float* aux_smem = reinterpret_cast<float*>(w_smem + NS * HALF_BYTES);
  uint8_t* bar_smem = reinterpret_cast<uint8_t*>(aux_smem) + G::AUX_BYTES;
  uint8_t* misc = bar_smem + G::BAR_BYTES;
  const uint32_t xstash = smem_u32(misc);                              // [TS][K0/8][64] x 16 B
  float* ss_smem = reinterpret_cast<float*>(misc + TS * G::XS_SLOT_BYTES);
  float* xi_smem = ss_smem + 2 * TS * 2 * ROWS;
  float* dot_smem = xi_smem + 2 * TS * 2 * ROWS;
  const uint32_t a_base = smem_u32(a_smem);
  const uint32_t w_base = smem_u32(w_smem);
  const uint32_t bars = smem_u32(bar_smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int n_tiles = (int)((p.n + ROWS - 1) / ROWS);
  const int n_pairs = (n_tiles + 1) >> 1;
  const int n_units = ((n_pairs + TS - 1) / TS) * p.splits;   // (TS tile pairs, member split)

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(bars + B4_W_FULL + 8 * s, leader ? 2 : 1);
      mbar_init(bars + B4_W_EMPTY + 8 * s, 1);
    }
    for (int t = 0; t < TS; ++t) {
      mbar_init(bars + B4_D_FULL + 8 * t, 1);
      mbar_init(bars + B4_DRAINED + 8 * t, 8);
      mbar_init(bars + B4_X_READY + 8 * t, 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(bars + B4_TMEM_PTR, (uint32_t)G::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(bar_smem + B4_TMEM_PTR);

  if (warp == 0) {
    // ===================================== producer =============================================
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
          for (int s = 0; s < p.stages_per_member; ++s) {
            mbar_wait(bars + B4_W_EMPTY + 8 * slot, phase ^ 1, p.error_flag, 1);
            mbar_arrive_expect_tx(bars + B4_W_FULL + 8 * slot, HALF_BYTES);
            bulk_g2s(w_base + slot * HALF_BYTES, src, HALF_BYTES, bars + B4_W_FULL + 8 * slot);
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
      const uint32_t full0 = mapa_shared(bars + B4_W_FULL, 0);
      for (int unit = cluster_id; unit < n_units; unit += n_clusters) {
        const int split = unit % p.splits;
        const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
        const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
        const int n_stages = (me - mb) * p.stages_per_member;
        for (int s = 0; s < n_stages; ++s) {
          mbar_wait(bars + B4_W_FULL + 8 * slot, phase, p.error_flag, 6);
          mbar_arrive_cluster(full0 + 8 * slot);
          if (++slot == NS) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader) ===================================
    // slot-major: the 2 KC stages of a layer stay in the ring while all TS slots use them
    constexpr uint32_t idesc = make_idesc_f16(2 * ROWS, NT);
    const uint64_t a_desc0 = make_sw128_desc(a_base);
    const uint64_t b_desc0 = make_sw128_desc(w_base);
    // the last K = 16 step of the layer-0 chunk is all zeros: one MMA over it with accumulate = 0
    // clears the correction tile, which layer 0 does not use (mlp_tcx.cu)
    const int k0_steps = p.K0 / 16 - 1;
    uint32_t slot = 0, phase = 0;
    uint32_t g = 0, xm = 0;
    auto wait_stage = [&](int i) -> uint32_t {   // i-th stage from the ring head
      uint32_t r = slot + (uint32_t)i, ph = phase;
      if (r >= (uint32_t)NS) { r -= NS; ph ^= 1; }
      mbar_wait_cluster_inline(bars + B4_W_FULL + 8 * r, ph, p.error_flag, 4);
      return r;
    };
    auto release_stages = [&](int n) {
      for (int i = 0; i < n; ++i) {
        if (elect_one()) umma_commit_pair(bars + B4_W_EMPTY + 8 * slot, 3);
        if (++slot == (uint32_t)NS) { slot = 0; phase ^= 1; }
      }
    };
    for (int unit = cluster_id; unit < n_units; unit += n_clusters) {
      const int split = unit % p.splits;
      const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
      const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
      for (int k = mb; k < me; ++k, ++xm) {
        // ---- layer 0 ----------------------------------------------------------------------------
        {
          const uint32_t prev_par = (g - 1) & 1;
          const uint32_t r0 = wait_stage(0);
#pragma unroll 1
          for (int t = 0; t < TS; ++t) {
            mbar_wait_cluster_inline(bars + B4_X_READY + 8 * t, xm & 1, p.error_flag, 2);
            if (g != 0) mbar_wait_cluster_inline(bars + B4_DRAINED + 8 * t, prev_par, p.error_flag, 3);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t ad = a_desc0 + (uint64_t)((t * G::A_SLOT_BYTES) >> 4);
              const uint64_t bd = b_desc0 + (uint64_t)((r0 * HALF_BYTES) >> 4);
              const uint32_t d_main = tmem_base + (uint32_t)(t * G::SLOT_COLS);
              for (int ks = 0; ks < k0_steps; ++ks)
                umma_bf16_pair(d_main, ad + 2 * ks, bd + 2 * ks, idesc, ks > 0 ? 1u : 0u);
              umma_bf16_pair(d_main + (uint32_t)G::TCOLS, ad + 2 * k0_steps, bd + 2 * k0_steps,
                             idesc, 0u);
              umma_commit_pair(bars + B4_D_FULL + 8 * t, 3);
            }
            __syncwarp();
          }
          release_stages(1);
          ++g;
        }
        // ---- hidden layers: per K chunk  g1 x (h1 -> main, h2 -> corr), g2 x h1 -> corr ----------
        for (int l = 1; l < p.L_mma; ++l) {
          const uint32_t prev_par = (g - 1) & 1;
          uint32_t rr[G::LAYER_STAGES];
#pragma unroll
          for (int i = 0; i < G::LAYER_STAGES; ++i) rr[i] = wait_stage(i);
#pragma unroll 1
          for (int t = 0; t < TS; ++t) {
            mbar_wait_cluster_inline(bars + B4_DRAINED + 8 * t, prev_par, p.error_flag, 3);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t d_main = tmem_base + (uint32_t)(t * G::SLOT_COLS);
              const uint32_t d_corr = d_main + (uint32_t)G::TCOLS;
#pragma unroll
              for (int kc = 0; kc < KC; ++kc) {
                const uint64_t ad1 =
                    a_desc0 + (uint64_t)((t * G::A_SLOT_BYTES + kc * CHUNKX_BYTES) >> 4);
                const uint64_t ad2 =
                    a_desc0 + (uint64_t)((t * G::A_SLOT_BYTES + (KC + kc) * CHUNKX_BYTES) >> 4);
                const uint64_t bd1 = b_desc0 + (uint64_t)((rr[2 * kc] * HALF_BYTES) >> 4);
                const uint64_t bd2 = b_desc0 + (uint64_t)((rr[2 * kc + 1] * HALF_BYTES) >> 4);
#pragma unroll
                for (int ks = 0; ks < CHUNK_K / 16; ++ks)
                  umma_bf16_pair(d_main, ad1 + 2 * ks, bd1 + 2 * ks, idesc,
                                 (kc > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < CHUNK_K / 16; ++ks)
                  umma_bf16_pair(d_corr, ad2 + 2 * ks, bd1 + 2 * ks, idesc,
                                 (kc > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < CHUNK_K / 16; ++ks)
                  umma_bf16_pair(d_corr, ad1 + 2 * ks, bd2 + 2 * ks, idesc, 1u);
              }
              umma_commit_pair(bars + B4_D_FULL + 8 * t, 3);
            }
            __syncwarp();
          }
          release_stages(G::LAYER_STAGES);
          ++g;
        }
      }
    }
  } else {
    // ===================================== epilogue =============================================
    const int ew = warp - 2;
    const int et = threadIdx.x - 64;
    const int q = warp & 3;              // TMEM lane quarter
    const int grp = ew >> 2;             // owns the slots t with t % NG == grp
    const int rh = q & 1, ch = q >> 1;   // row half / column half
    const int row = rh * 32 + lane;      // row of the slot's 64-row tile
    const uint32_t a_row0 = a_base + (row >> 3) * 1024 + (row & 7) * 128;   // slot 0, chunk 0
    const int rx = row & 7;
    const bool x_owner = ch == 0;        // writes x, owns the row's Welford state
    const uint32_t lane_addr0 = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t drained0 = mapa_shared(bars + B4_DRAINED, 0);
    const uint32_t xready0 = mapa_shared(bars + B4_X_READY, 0);
    uint32_t g = 0, mcount = 0, ucount = 0;
    constexpr int AUX_PER_THREAD = (G::AUX_FLOATS + EPI_THREADS - 1) / EPI_THREADS;

    const bool use_stash = p.K0 <= 32;
    auto xi_at = [&](int par, int t, int which) {
      return xi_smem + ((par * TS + t) * 2 + which) * ROWS + row;
    };
    auto build_x = [&](int t, int tile, bool to_stash, int par) {
      if ((t % NG) == grp && x_owner)
        build_x_row_x(p, (int64_t)tile * ROWS + row, to_stash,
                      xstash + (uint32_t)(t * G::XS_SLOT_BYTES + (row << 4)), (uint32_t)(ROWS << 4),
                      a_row0 + (uint32_t)(t * G::A_SLOT_BYTES), rx, xi_at(par, t, 0),
                      xi_at(par, t, 1));
    };
    auto publish_x = [&](int t, int tile, int par) {
      if ((t % NG) == grp && x_owner) {
        if (use_stash) {
          for (int piece = 0; piece < p.K0 / 8; ++piece) {
            uint32_t a, b, c, d;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
                         : "r"(xstash + (uint32_t)(t * G::XS_SLOT_BYTES + ((piece * ROWS + row) << 4))));
            st_shared_v4(a_row0 + (uint32_t)(t * G::A_SLOT_BYTES) + (uint32_t)((piece ^ rx) << 4), a,
                         b, c, d);
          }
        } else {
          build_x(t, tile, false, par);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(xready0 + 8 * t);
      }
    };

    float aux_pf[AUX_PER_THREAD];
    auto aux_prefetch = [&](int member_global, int l) {
      const int wslot = p.shared_weights ? 0 : member_global;
      const bool last = (l == p.L_mma - 1);
      const float* bias =
          p.bias[l] + (size_t)((l == 0 && p.bias0_per_member) ? member_global : wslot) * H;
      const float* wl = p.w_last + (size_t)wslot * H;
      const float* st = p.lstats + ((size_t)wslot * p.L_mma + l) * STAT_FLOATS;
#pragma unroll
      for (int j = 0; j < AUX_PER_THREAD; ++j) {
        const int i = et + j * EPI_THREADS;
        float v = 0.f;
        if (i < H) v = __ldg(bias + i);
        else if (i < 2 * H) { if (last) v = __ldg(wl + (i - H)); }
        else if (i < SBASE + STAT_FLOATS) {
          v = __ldg(st + (i - SBASE));
          if (i == SBASE + 2 && l == 0 && p.bmax0) v = __ldg(p.bmax0 + member_global);
        }
        aux_pf[j] = v;
      }
    };
    // tile of slot t in a unit
    auto tile_of = [&](int unit, int t) { return 2 * ((unit / p.splits) * TS + t) + (int)rank; };

    bool first_step = true;
    for (int unit = cluster_id; unit < n_units; unit += n_clusters, ++ucount) {
      const int split = unit % p.splits;
      const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
      const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
      const int xpar = (int)(ucount & 1);

      float wf_n = 0.f, wf_mean[OWN], wf_m2[OWN];
#pragma unroll
      for (int j = 0; j < OWN; ++j) wf_mean[j] = 0.f, wf_m2[j] = 0.f;

      if (first_step) {
#pragma unroll 1
        for (int t = 0; t < TS; ++t) {
          if (use_stash) build_x(t, tile_of(unit, t), true, xpar);
          publish_x(t, tile_of(unit, t), xpar);
        }
        aux_prefetch(p.member_begin + mb, 0);
        first_step = false;
      }

      for (int k = mb; k < me; ++k, ++mcount) {
        const int kg = p.member_begin + k;
        const int wslot = p.shared_weights ? 0 : kg;
        float dot[OWN];      // last-Linear partial dot product of each owned slot's row
        float a_inv[OWN];    // 1 / (row scale of the A operand feeding the current layer)
#pragma unroll
        for (int j = 0; j < OWN; ++j) dot[j] = 0.f, a_inv[j] = 1.f;
        int drop_ord = 0;
        const uint8_t* mask_layer = p.masks;

        int nk = k + 1, nunit = unit;
        bool have_next = true;
        if (nk >= me) {
          nunit = unit + n_clusters;
          have_next = nunit < n_units;
          nk = (int)(((int64_t)p.member_count * (nunit % p.splits)) / p.splits);
        }
        const bool next_unit = (k + 1 >= me);

        for (int l = 0; l < p.L_mma; ++l, ++g) {
          const bool last = (l == p.L_mma - 1);
          const bool relu = (p.relu_mask >> l) & 1u;
          const bool has_drop = (p.dropout_mask >> l) & 1u;
          const int drop = (MC && has_drop) ? p.drop_mode : 0;
          const float in_scale =
              (l > 0 && ((p.dropout_mask >> (l - 1)) & 1u) && p.drop_mode) ? p.drop_scale : 1.f;

          float* aux = aux_smem + (g & 1) * G::AUX_FLOATS;
#pragma unroll
          for (int j = 0; j < AUX_PER_THREAD; ++j) {
            const int i = et + j * EPI_THREADS;
            if (i < G::AUX_FLOATS) aux[i] = aux_pf[j];
          }
          epi_bar_sync_n<EPI_THREADS>();
          if (!last) aux_prefetch(kg, l + 1);
          else if (have_next) aux_prefetch(p.member_begin + nk, 0);
          // x (and its row statistics) of the NEXT unit's tiles while the last layer runs
          if (last && have_next && next_unit && use_stash) {
#pragma unroll 1
            for (int t = 0; t < TS; ++t) build_x(t, tile_of(nunit, t), true, xpar ^ 1);
          }

#pragma unroll
          for (int j = 0; j < OWN; ++j) {
            const int t = grp + NG * j;
            const int64_t grow = (int64_t)tile_of(unit, t) * ROWS + row;
            // ---- row scales of this slot and step (mlp_tcx.cu) -------------------------------
            float n2;
            if (l == 0) {
              n2 = *xi_at(xpar, t, 1);
              a_inv[j] = *xi_at(xpar, t, 0);
            } else {
              const float* sp = ss_smem + ((((g - 1) & 1) * TS + t) * 2) * ROWS + row;
              n2 = sp[0] + sp[ROWS];
            }
            const float rs = a_inv[j] * aux[SBASE + 0] * in_scale;
            float s_out = 1.f, s_inv = 1.f;
            if (!last) {
              const float bound = fmaf(aux[SBASE + 1] * in_scale, sqrtf(n2), aux[SBASE + 2]);
              pow2_scale(bound, s_out, s_inv);
            }
            // keep masks do not depend on the activations: before the wait
            uint32_t keepw[G::BPT];
            if (drop) {
#pragma unroll
              for (int b = 0; b < G::BPT; ++b)
                keepw[b] = keep_bits32(p, drop, kg, drop_ord, grow, ch * G::TCOLS + 32 * b,
                                       mask_layer, H);
            }
            if (lane == 0) mbar_wait(bars + B4_D_FULL + 8 * t, g & 1, p.error_flag, 5);
            __syncwarp();
            tc_fence_after();
            if (last && have_next) publish_x(t, tile_of(nunit, t), next_unit ? (xpar ^ 1) : xpar);

            const uint32_t lane_addr = lane_addr0 + (uint32_t)(t * G::SLOT_COLS);
            const uint32_t a_row = a_row0 + (uint32_t)(t * G::A_SLOT_BYTES);
            float dslot[1] = {0.f};
            float ss = 0.f;
#define UQ_DRAINX4(R, D, L) \
  drain_x4<H, R, (D) && MC, L>(lane_addr, a_row, rx, ch, aux, aux + H, keepw, rs, s_out, dslot, ss)
            if (last) {
              if (relu) { if (drop) UQ_DRAINX4(true, true, true); else UQ_DRAINX4(true, false, true); }
              else { if (drop) UQ_DRAINX4(false, true, true); else UQ_DRAINX4(false, false, true); }
            } else {
              if (relu) { if (drop) UQ_DRAINX4(true, true, false); else UQ_DRAINX4(true, false, false); }
              else { if (drop) UQ_DRAINX4(false, true, false); else UQ_DRAINX4(false, false, false); }
            }
#undef UQ_DRAINX4
            if (last) {
              dot[j] = dslot[0];
            } else {
              ss_smem[(((g & 1) * TS + t) * 2 + ch) * ROWS + row] = ss;
              a_inv[j] = s_inv;
            }
            // this warp is done with slot t: accumulators drained, A pieces rewritten
            tc_fence_before();
            if (!last) fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(drained0 + 8 * t);
          }
          if (has_drop) {
            if (p.masks) mask_layer += (size_t)p.total_members * (size_t)p.n * (size_t)H;
            ++drop_ord;
          }
        }

        // ---- the row's two partial dot products, then Welford over members -------------------
        float* dx = dot_smem + (mcount & 1) * TS * ROWS;
        if (!x_owner) {
#pragma unroll
          for (int j = 0; j < OWN; ++j) dx[(grp + NG * j) * ROWS + row] = dot[j];
        }
        epi_bar_sync_n<EPI_THREADS>();
        if (x_owner) {
          wf_n += 1.f;
          const float inv_n = 1.f / wf_n;
          const float bl = __ldg(p.b_last + wslot);
          const float out_scale = final_dropout_scale(p);
#pragma unroll
          for (int j = 0; j < OWN; ++j) {
            float y = dot[j] + dx[(grp + NG * j) * ROWS + row];
            y = fmaf(y, out_scale, bl);
            if (p.last_relu) y = fmaxf(y, 0.f);
            member_fold(p, kg, 0, y, inv_n, wf_mean[j], wf_m2[j]);
          }
        }
      }

      if (x_owner) {
#pragma unroll
        for (int j = 0; j < OWN; ++j) {
          const int64_t grow = (int64_t)tile_of(unit, grp + NG * j) * ROWS + row;
          if (grow < p.n) {
            if (p.splits > 1) {
              p.part_mean[(size_t)split * (size_t)p.n + grow] = wf_mean[j];
              p.part_m2[(size_t)split * (size_t)p.n + grow] = wf_m2[j];
            } else {
              p.out0[grow] = wf_mean[j];
              p.out1[grow] = second_output(p, wf_m2[j], wf_n, grow);
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

template <int H, bool MC>
int launch_tcx4(const TcParams& p, cudaStream_t st) {
  using G = GeoX4<H>;
  auto kern = uq_mlp_tcx4_kernel<H, MC>;
  static std::atomic<int> cached_clusters[64];
  int dev = 0;
  cudaGetDevice(&dev);
  const int64_t n_tiles = (p.n + ROWS - 1) / ROWS;
  const int64_t n_pairs = (n_tiles + 1) / 2;
  const int64_t units = ((n_pairs + TS - 1) / TS) * p.splits;
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
    max_clusters = sms / 2;
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

bool tcx4_supported(int hidden, int dout_pad) {
  return (hidden == 64 || hidden == 128) && dout_pad == 1;
}
int tcx4_rows_per_unit() { return 2 * TS * ROWS; }

int tcx4_launch(const tc::TcParams& p, int hidden, cudaStream_t st) {
  const bool mc = p.drop_mode != 0 && p.dropout_mask != 0;
  if (hidden == 64) return mc ? launch_tcx4<64, true>(p, st) : launch_tcx4<64, false>(p, st);
  if (hidden == 128) return mc ? launch_tcx4<128, true>(p, st) : launch_tcx4<128, false>(p, st);
  set_error("narrow split kernel: unsupported hidden width %d", hidden);
  return UQ_ERR_UNSUPPORTED;
}

}  // namespace uq
