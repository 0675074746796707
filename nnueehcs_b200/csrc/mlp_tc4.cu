// bf16 tcgen05 path of the UQ forward for NARROW nets (hidden width 64 / 128, one output):
// CTA pairs with FOUR sample tiles in flight per CTA.
//
// At H = 128 a layer of one 128-row tile is 512 tensor-core cycles, but the chain
//   tcgen05.commit -> epilogue wake-up -> drain -> barrier -> MMA resume
// costs ~3000 cycles, so one tile per CTA (mlp_tc2.cu) leaves the tensor core idle 80 % of the
// time (18 % of peak on BASELINE configs[2]).  A narrow net leaves room for more: four [128 x H]
// fp32 accumulators fit the 512 TMEM columns and four activation tiles fit shared memory.  The
// four "tile slots" of a CTA run the same member and layer; the MMA warp walks them slot-major
// with the layer's weight stages resident in the ring (loaded once, used by all four slots:
// weight traffic per flop drops 4x as a side effect), so while the epilogue warps drain slot t
// the tensor core is already working on slots t+1.. -- the latency chain of one slot is hidden
// behind the MMAs of the other three.
//
// Per slot: D_FULL[t] (layer accumulated, commit multicast), DRAINED[t] (the slot's 8 epilogue warps
// of the pair are done with its accumulator and have rewritten its A chunks), X_READY[t].  Epilogue
// warp group g owns the slots t with t % 2 == g (whole rows: no cross-warp exchange for the last
// Linear's dot product), so two slots are drained concurrently.
// Everything else (split weight stages, peer relay, bias in the MMA or staged for the epilogue,
// in-place bf16 write-back, CUDA-core last Linear + per-row Welford) follows mlp_tc2_impl.cuh.
//
// Replaces: MCDropoutModel.forward (models.py:147-163), EnsembleModel.forward (:99-108) and the
// anchored forward behind DeltaUQMLP.forward (:313-341) for hidden widths 64 and 128, d_out = 1
// (the binomial-options surrogate of examples/binomial_options/config.yaml:16-54 is 6 x 128).
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
constexpr int TS = 4;                          // tile slots per CTA
constexpr int NG4 = 2;                         // epilogue warp groups
constexpr int EPI4_THREADS = NG4 * 128;
constexpr int NUM4_THREADS = 64 + EPI4_THREADS;

// BIAS: the folded bias is accumulated by the tensor core (per layer one bias stage in the ring,
// per slot one extra K = 16 MMA of an all-ones A tile against it), not added by the epilogue;
// with live dropout the activations are then stored with their 1/(1-p) (see epi_math).
template <int H, bool BIAS = false>
struct Geo4 {
  static_assert(H == 64 || H == 128, "narrow-net kernel: hidden width 64 or 128");
  static constexpr int KC = H / CHUNK_K;                 // 1 or 2
  static constexpr int NT = H;                           // one accumulator per slot, MMA N = H
  static constexpr int TMEM_COLS = TS * H <= 256 ? 256 : 512;
  static constexpr int STAGE_BYTES = NT * 128;
  static constexpr int HALF_BYTES = STAGE_BYTES / 2;
  static constexpr int A_SLOT_BYTES = KC * CHUNK_BYTES;
  static constexpr int A_BYTES = TS * A_SLOT_BYTES;
  static constexpr int WL_OFF = BIAS ? 0 : H;            // w_last inside a step's aux block
  static constexpr int AUX_FLOATS = WL_OFF + H;          // [bias +] w_last of one step
  static constexpr int AUX_BYTES = 2 * AUX_FLOATS * 4;
  static constexpr int ONES_BYTES = BIAS ? CONST_TILE_BYTES : 0;
  static constexpr int LSTAGES = KC + (BIAS ? 1 : 0);    // ring stages of one hidden layer
  static constexpr int XS_SLOT_BYTES = TILE_M * 64;      // x stash (K0 <= 32)
  static constexpr int BAR_BYTES = 384;
  static constexpr int MISC_BYTES = 1024 + BAR_BYTES + TS * XS_SLOT_BYTES;
  static constexpr int BUDGET = SMEM_LIMIT - A_BYTES - AUX_BYTES - MISC_BYTES - ONES_BYTES;
  static constexpr int NS_RAW = BUDGET / HALF_BYTES;
  static constexpr int NSTAGES = NS_RAW > 8 ? 8 : NS_RAW;
  static_assert(NSTAGES >= 2 * LSTAGES, "the ring must hold two layers of weight stages");
  static constexpr int SMEM_BYTES =
      A_BYTES + NSTAGES * HALF_BYTES + ONES_BYTES + AUX_BYTES + MISC_BYTES;
};

constexpr uint32_t B4_W_FULL = 0;       // 8 x 8 B
constexpr uint32_t B4_W_EMPTY = 64;     // 8 x 8 B
constexpr uint32_t B4_D_FULL = 128;     // TS x 8 B   commit multicast
constexpr uint32_t B4_DRAINED = 160;    // TS x 8 B   leader only: 4 warps of each CTA
constexpr uint32_t B4_X_READY = 192;    // TS x 8 B   leader only: 4 warps of each CTA
constexpr uint32_t B4_TMEM_PTR = 224;

// drain all chunks of one slot (this warp: 32 rows of it)
// keepw: the keep-mask words of the slot's 2 KC blocks, computed by the caller BEFORE it waits for
// the slot's MMAs (they do not depend on the activations; several independent Philox chains in
// flight hide each other's latency -- with two warps per scheduler a single chain does not)
template <int H, bool RELU, bool DROP, bool LAST, bool BIAS>
__device__ __forceinline__ void drain4(uint32_t lane_addr, uint32_t a_row, int rx,
                                       const float* bias_s, const float* wl_s,
                                       const uint32_t (&keepw)[2 * (H / CHUNK_K)], float in_scale,
                                       float (&dot)[1]) {
  constexpr int KC = H / CHUNK_K;
  uint32_t acc0[32], acc1[32];
  tmem_ld32(lane_addr, acc0);
#pragma unroll
  for (int c = 0; c < KC; ++c) {
    const int col0 = c * CHUNK_K;
    const uint32_t a_dst = a_row + (uint32_t)c * CHUNK_BYTES;
    float4 bv[8];
    if (!BIAS) {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) bv[j4] = reinterpret_cast<const float4*>(bias_s + col0)[j4];
    }
    tmem_ld_wait();
    tmem_ld32(lane_addr + (uint32_t)(col0 + 32), acc1);
    epi_block2<H, 1, 32, RELU, DROP, LAST, BIAS>(acc0, bv, DROP ? keepw[2 * c] : 0xffffffffu,
                                                 in_scale, a_dst, 0, rx, wl_s + col0, wl_s + col0,
                                                 dot);
    if (!BIAS) {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        bv[j4] = reinterpret_cast<const float4*>(bias_s + col0 + 32)[j4];
    }
    tmem_ld_wait();
    if (c + 1 < KC) tmem_ld32(lane_addr + (uint32_t)(col0 + CHUNK_K), acc0);
    epi_block2<H, 1, 32, RELU, DROP, LAST, BIAS>(acc1, bv, DROP ? keepw[2 * c + 1] : 0xffffffffu,
                                                 in_scale, a_dst, 4, rx, wl_s + col0 + 32,
                                                 wl_s + col0 + 32, dot);
  }
}

// MC: the launch has live dropout.  The dropout-free instantiation carries no mask code (with it the
// bias-in-the-MMA kernel needs 168 registers and spills; without, 154 and none: 19.9 -> 18.7 ms on
// deltauq32_binomial_4M).
template <int H, bool BIAS, bool MC>
__global__ void __launch_bounds__(NUM4_THREADS, 1)
uq_mlp_tc4_kernel(const __grid_constant__ TcParams p) {
  using G = Geo4<H, BIAS>;
  constexpr int KC = G::KC, NT = G::NT, NS = G::NSTAGES;
  constexpr uint32_t STAGE_BYTES = G::STAGE_BYTES, HALF_BYTES = G::HALF_BYTES;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* a_smem = smem;                                  // TS slots x KC chunks of 16 KB
  uint8_t* w_smem = smem + G::A_BYTES;                     // NS half-stages
  uint8_t* ones_smem = w_smem + NS * HALF_BYTES;           // BIAS: constant-1 A tile (bf16)
  float* aux_smem = reinterpret_cast<float*>(ones_smem + G::ONES_BYTES);
  uint8_t* bar_smem = reinterpret_cast<uint8_t*>(aux_smem) + G::AUX_BYTES;
  const uint32_t xstash = smem_u32(bar_smem + G::BAR_BYTES);         // [TS][K0/8][128] x 16 B
  const uint32_t a_base = smem_u32(a_smem);
  const uint32_t w_base = smem_u32(w_smem);
  const uint32_t bars = smem_u32(bar_smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int n_pairs = (p.n_tiles + 1) >> 1;
  const int n_units = ((n_pairs + TS - 1) / TS) * p.splits;   // (TS tile pairs, member split)

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(bars + B4_W_FULL + 8 * s, leader ? 2 : 1);
      mbar_init(bars + B4_W_EMPTY + 8 * s, 1);
    }
    for (int t = 0; t < TS; ++t) {
      mbar_init(bars + B4_D_FULL + 8 * t, 1);
      mbar_init(bars + B4_DRAINED + 8 * t, 8);
      mbar_init(bars + B4_X_READY + 8 * t, 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(bars + B4_TMEM_PTR, (uint32_t)G::TMEM_COLS);
  if (BIAS) {   // the all-ones A tile of the bias MMAs (read by the tensor core: async proxy)
    for (int i = threadIdx.x; i < G::ONES_BYTES / 4; i += NUM4_THREADS)
      reinterpret_cast<uint32_t*>(ones_smem)[i] = 0x3F803F80u;   // bf16 (1.0, 1.0)
    fence_proxy_async_smem();
  }
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
          auto load = [&](const uint8_t* from) {
            mbar_wait(bars + B4_W_EMPTY + 8 * slot, phase ^ 1, p.error_flag, 1);
            mbar_arrive_expect_tx(bars + B4_W_FULL + 8 * slot, HALF_BYTES);
            bulk_g2s(w_base + slot * HALF_BYTES, from, HALF_BYTES, bars + B4_W_FULL + 8 * slot);
            if (++slot == NS) { slot = 0; phase ^= 1; }
          };
          if (BIAS) {   // per layer: its weight stages, then its bias stage
            const uint8_t* bsrc =
                p.bias_image +
                (p.shared_weights ? 0 : (size_t)(p.member_begin + k) * p.L_mma * STAGE_BYTES) +
                rank * HALF_BYTES;
            for (int l = 0; l < p.L_mma; ++l, bsrc += STAGE_BYTES) {
              for (int s = 0; s < (l == 0 ? 1 : KC); ++s, src += STAGE_BYTES) load(src);
              if (l == 0 && p.bias0_image != nullptr)   // per-anchor layer-0 bias
                load(p.bias0_image + (size_t)(p.member_begin + k) * STAGE_BYTES + rank * HALF_BYTES);
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
      const uint32_t full0 = mapa_shared(bars + B4_W_FULL, 0);
      for (int unit = cluster_id; unit < n_units; unit += n_clusters) {
        const int split = unit % p.splits;
        const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
        const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
        const int n_stages = (me - mb) * (p.stages_per_member + (BIAS ? p.L_mma : 0));
        for (int s = 0; s < n_stages; ++s) {
          mbar_wait(bars + B4_W_FULL + 8 * slot, phase, p.error_flag, 6);
          mbar_arrive_cluster(full0 + 8 * slot);
          if (++slot == NS) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader) ===================================
    // slot-major: the KC stages of a layer stay in the ring while all TS slots use them
    constexpr uint32_t idesc = make_idesc_bf16(2 * TILE_M, NT);
    const uint64_t a_desc0 = make_sw128_desc(a_base);
    const uint64_t b_desc0 = make_sw128_desc(w_base);
    const uint64_t ones_desc = make_sw128_const_desc(smem_u32(ones_smem));
    const int k0_steps = p.K0 / 16;
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
          const uint32_t rb = BIAS ? wait_stage(1) : 0;
#pragma unroll 1
          for (int t = 0; t < TS; ++t) {
            mbar_wait_cluster_inline(bars + B4_X_READY + 8 * t, xm & 1, p.error_flag, 2);
            if (g != 0) mbar_wait_cluster_inline(bars + B4_DRAINED + 8 * t, prev_par, p.error_flag, 3);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t ad = a_desc0 + (uint64_t)((t * G::A_SLOT_BYTES) >> 4);
              const uint64_t bd = b_desc0 + (uint64_t)((r0 * HALF_BYTES) >> 4);
              for (int ks = 0; ks < k0_steps; ++ks)
                umma_bf16_pair(tmem_base + t * H, ad + 2 * ks, bd + 2 * ks, idesc, ks > 0 ? 1u : 0u);
              if (BIAS)   // accumulator += ones . bias^T
                umma_bf16_pair(tmem_base + t * H, ones_desc,
                               b_desc0 + (uint64_t)((rb * HALF_BYTES) >> 4), idesc, 1u);
              umma_commit_pair(bars + B4_D_FULL + 8 * t, 3);
            }
            __syncwarp();
          }
          release_stages(1 + (BIAS ? 1 : 0));
          ++g;
        }
        // ---- hidden layers ----------------------------------------------------------------------
        for (int l = 1; l < p.L_mma; ++l) {
          const uint32_t prev_par = (g - 1) & 1;
          uint32_t rr[KC];
#pragma unroll
          for (int kc = 0; kc < KC; ++kc) rr[kc] = wait_stage(kc);
          const uint32_t rb = BIAS ? wait_stage(KC) : 0;
#pragma unroll 1
          for (int t = 0; t < TS; ++t) {
            mbar_wait_cluster_inline(bars + B4_DRAINED + 8 * t, prev_par, p.error_flag, 3);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int kc = 0; kc < KC; ++kc) {
                const uint64_t ad =
                    a_desc0 + (uint64_t)((t * G::A_SLOT_BYTES + kc * CHUNK_BYTES) >> 4);
                const uint64_t bd = b_desc0 + (uint64_t)((rr[kc] * HALF_BYTES) >> 4);
#pragma unroll
                for (int ks = 0; ks < CHUNK_K / 16; ++ks)
                  umma_bf16_pair(tmem_base + t * H, ad + 2 * ks, bd + 2 * ks, idesc,
                                 (kc > 0 || ks > 0) ? 1u : 0u);
              }
              if (BIAS)
                umma_bf16_pair(tmem_base + t * H, ones_desc,
                               b_desc0 + (uint64_t)((rb * HALF_BYTES) >> 4), idesc, 1u);
              umma_commit_pair(bars + B4_D_FULL + 8 * t, 3);
            }
            __syncwarp();
          }
          release_stages(G::LSTAGES);
          ++g;
        }
      }
    }
  } else {
    // ===================================== epilogue =============================================
    const int ew = warp - 2;
    const int et = threadIdx.x - 64;
    const int q = warp & 3;
    const int grp = ew >> 2;
    const int row = q * 32 + lane;
    const uint32_t a_row0 = a_base + (row >> 3) * 1024 + (row & 7) * 128;   // slot 0, chunk 0
    const int rx = row & 7;
    const uint32_t lane_addr0 = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t drained0 = mapa_shared(bars + B4_DRAINED, 0);
    const uint32_t xready0 = mapa_shared(bars + B4_X_READY, 0);
    uint32_t g = 0, mcount = 0;
    constexpr int AUX_PER_THREAD = (G::AUX_FLOATS + EPI4_THREADS - 1) / EPI4_THREADS;

    const bool use_stash = p.K0 <= 32;
    auto build_x = [&](int t, int tile, int member_global, bool to_stash) {
      if ((t % NG4) == grp)
        build_x_row(p, (int64_t)tile * TILE_M + row, member_global, to_stash,
                    xstash + (uint32_t)(t * G::XS_SLOT_BYTES + (row << 4)), (uint32_t)(TILE_M << 4),
                    a_row0 + (uint32_t)(t * G::A_SLOT_BYTES), rx);
    };
    auto publish_x = [&](int t, int tile, int member_global) {
      if ((t % NG4) == grp) {
        if (use_stash) {
          for (int piece = 0; piece < p.K0 / 8; ++piece) {
            uint32_t a, b, c, d;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
                         : "r"(xstash + (uint32_t)(t * G::XS_SLOT_BYTES + ((piece * TILE_M + row) << 4))));
            st_shared_v4(a_row0 + (uint32_t)(t * G::A_SLOT_BYTES) + (uint32_t)((piece ^ rx) << 4), a,
                         b, c, d);
          }
        } else {
          build_x(t, tile, member_global, false);
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
#pragma unroll
      for (int j = 0; j < AUX_PER_THREAD; ++j) {
        const int i = et + j * EPI4_THREADS;
        float v = 0.f;
        if (!BIAS && i < H) v = __ldg(bias + i);
        else if (last && i >= G::WL_OFF && i < G::WL_OFF + H) v = __ldg(wl + (i - G::WL_OFF));
        aux_pf[j] = v;
      }
    };
    // tile of slot t in a unit
    auto tile_of = [&](int unit, int t) { return 2 * ((unit / p.splits) * TS + t) + (int)rank; };

    bool first_step = true;
    for (int unit = cluster_id; unit < n_units; unit += n_clusters) {
      const int split = unit % p.splits;
      const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
      const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);

      constexpr int OWN = TS / NG4;   // slots owned by this warp's group: t = grp + NG4 * j
      float wf_n = 0.f, wf_mean[OWN], wf_m2[OWN];
#pragma unroll
      for (int j = 0; j < OWN; ++j) wf_mean[j] = 0.f, wf_m2[j] = 0.f;

      if (first_step) {
#pragma unroll 1
        for (int t = 0; t < TS; ++t) {
          if (use_stash) build_x(t, tile_of(unit, t), p.member_begin + mb, true);
          publish_x(t, tile_of(unit, t), p.member_begin + mb);
        }
        aux_prefetch(p.member_begin + mb, 0);
        first_step = false;
      }

      for (int k = mb; k < me; ++k, ++mcount) {
        const int kg = p.member_begin + k;
        const int wslot = p.shared_weights ? 0 : kg;
        float dot[OWN];  // last-Linear dot product of each owned slot's row
#pragma unroll
        for (int j = 0; j < OWN; ++j) dot[j] = 0.f;
        int drop_ord = 0;
        const uint8_t* mask_layer = p.masks;

        int nk = k + 1, nunit = unit;
        bool have_next = true;
        if (nk >= me) {
          nunit = unit + n_clusters;
          have_next = nunit < n_units;
          nk = (int)(((int64_t)p.member_count * (nunit % p.splits)) / p.splits);
        }
        const bool new_tiles = (nunit / p.splits) != (unit / p.splits);

        for (int l = 0; l < p.L_mma; ++l, ++g) {
          const bool last = (l == p.L_mma - 1);
          const bool relu = (p.relu_mask >> l) & 1u;
          const bool has_drop = (p.dropout_mask >> l) & 1u;
          const int drop = (MC && has_drop) ? p.drop_mode : 0;
          // epilogue-bias: 1/(1-p) owed by the previous layer's dropout; bias in the MMA: this layer's
          const float in_scale =
              BIAS ? (drop ? p.drop_scale : 1.f)
                   : (l > 0 && ((p.dropout_mask >> (l - 1)) & 1u) && p.drop_mode) ? p.drop_scale : 1.f;

          float* aux = aux_smem + (g & 1) * G::AUX_FLOATS;
#pragma unroll
          for (int j = 0; j < AUX_PER_THREAD; ++j) {
            const int i = et + j * EPI4_THREADS;
            if (i < G::AUX_FLOATS) aux[i] = aux_pf[j];
          }
          epi_bar_sync_n<EPI4_THREADS>();
          if (!last) aux_prefetch(kg, l + 1);
          else if (have_next) aux_prefetch(p.member_begin + nk, 0);
          if (last && have_next && use_stash && new_tiles) {
#pragma unroll 1
            for (int t = 0; t < TS; ++t) build_x(t, tile_of(nunit, t), p.member_begin + nk, true);
          }

#pragma unroll 1   // one copy of the drain code for all slots (it is ~1.5 k instructions)
          for (int t = grp; t < TS; t += NG4) {
            float dslot[1] = {0.f};
            const int64_t grow = (int64_t)tile_of(unit, t) * TILE_M + row;
            uint32_t keepw[2 * KC];
            if (drop) {
#pragma unroll
              for (int b = 0; b < 2 * KC; ++b)
                keepw[b] = keep_bits32(p, drop, kg, drop_ord, grow, 32 * b, mask_layer, H);
            }
            if (lane == 0) mbar_wait(bars + B4_D_FULL + 8 * t, g & 1, p.error_flag, 5);
            __syncwarp();
            tc_fence_after();
            if (last && have_next) publish_x(t, tile_of(nunit, t), p.member_begin + nk);

            const uint32_t lane_addr = lane_addr0 + (uint32_t)(t * H);
            const uint32_t a_row = a_row0 + (uint32_t)(t * G::A_SLOT_BYTES);
#define UQ_DRAIN4(R, D, L) \
  drain4<H, R, (D) && MC, L, BIAS>(lane_addr, a_row, rx, aux, aux + G::WL_OFF, keepw, in_scale, dslot)
            if (last) {
              if (relu) { if (drop) UQ_DRAIN4(true, true, true); else UQ_DRAIN4(true, false, true); }
              else { if (drop) UQ_DRAIN4(false, true, true); else UQ_DRAIN4(false, false, true); }
            } else {
              if (relu) { if (drop) UQ_DRAIN4(true, true, false); else UQ_DRAIN4(true, false, false); }
              else { if (drop) UQ_DRAIN4(false, true, false); else UQ_DRAIN4(false, false, false); }
            }
#undef UQ_DRAIN4
            if (last) {
#pragma unroll
              for (int j = 0; j < OWN; ++j)
                if (grp + NG4 * j == t) dot[j] = dslot[0];
            }
            // this warp is done with slot t: accumulator drained, A chunks rewritten
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

        // ---- Welford over members, per owned slot ---------------------------------------------
        {
          wf_n += 1.f;
          const float inv_n = 1.f / wf_n;
          const float bl = __ldg(p.b_last + wslot);
#pragma unroll
          for (int j = 0; j < OWN; ++j) {
            float y = fmaf(dot[j], BIAS ? 1.f : final_dropout_scale(p), bl);
            if (p.last_relu) y = fmaxf(y, 0.f);
            member_fold(p, kg, 0, y, inv_n, wf_mean[j], wf_m2[j]);
          }
        }
      }

#pragma unroll
      for (int j = 0; j < OWN; ++j) {
        const int64_t grow = (int64_t)tile_of(unit, grp + NG4 * j) * TILE_M + row;
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

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, (uint32_t)G::TMEM_COLS);
  }
}

template <int H, bool BIAS, bool MC>
int launch_tc4(const TcParams& p, cudaStream_t st) {
  using G = Geo4<H, BIAS>;
  auto kern = uq_mlp_tc4_kernel<H, BIAS, MC>;
  // per-device launch geometry of this instantiation, queried once (the occupancy query and the
  // attribute call cost tens of microseconds, which shows on millisecond-sized forwards)
  static std::atomic<int> cached_clusters[64];   // zero-initialised; races only repeat the query
  int dev = 0;
  cudaGetDevice(&dev);
  const int64_t n_pairs = (p.n_tiles + 1) / 2;
  const int64_t units = ((n_pairs + TS - 1) / TS) * p.splits;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(NUM4_THREADS, 1, 1);
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

bool tc4_supported(int hidden, int dout_pad) { return (hidden == 64 || hidden == 128) && dout_pad == 1; }
int tc4_rows_per_unit() { return 2 * TS * tc::TILE_M; }

int tc4_launch(const tc::TcParams& p, int hidden, cudaStream_t st) {
  const bool bias = p.bias_image != nullptr && bias_in_mma_enabled();
  const bool mc = p.drop_mode != 0 && p.dropout_mask != 0;
#define UQ_TC4_DISPATCH(HH)                                                                  \
  if (hidden == HH)                                                                          \
    return bias ? (mc ? launch_tc4<HH, true, true>(p, st) : launch_tc4<HH, true, false>(p, st))  \
                : (mc ? launch_tc4<HH, false, true>(p, st) : launch_tc4<HH, false, false>(p, st));
  UQ_TC4_DISPATCH(64)
  UQ_TC4_DISPATCH(128)
#undef UQ_TC4_DISPATCH
  set_error("bf16 narrow-net kernel: unsupported hidden width %d", hidden);
  return UQ_ERR_UNSUPPORTED;
}

}  // namespace uq
