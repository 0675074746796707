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
//   warps 2..9  epilogue   tcgen05.ld the accumulator, add the (BN-folded) bias, ReLU, apply the
//                          Philox / injected dropout mask, round to bf16 and write the next
//                          layer's A operand back into the same smem chunks (in place), chunk by
//                          chunk so the next layer's MMAs start while the tail is still draining.
//                          The last Linear (H -> d_out, d_out <= 8) is a CUDA-core dot product
//                          in the same pass, feeding a per-row Welford (count, mean, M2) across
//                          members that lives in registers -- the [K, N, out] stack of
//                          nnueehcs/models.py:103,159 never exists.
//
// Layer 0 (d_in <= 21 inputs) is also an MMA: x is split into bf16 hi + lo parts and the folded
// first-layer weights likewise, laid out as [x_hi | x_lo | x_hi] . [w_hi | w_hi | w_lo] inside
// one K <= 64 chunk, which recovers ~16 mantissa bits for one extra K=16 step.
//
// Replaces: EnsembleModel.forward (models.py:99-108), MCDropoutModel.forward (:147-163) and the
// anchored forward behind DeltaUQMLP.forward (:313-341) for MLPs whose hidden widths are equal.
#include <stdlib.h>

#include "common.cuh"
#include "philox.cuh"

namespace uq {

namespace {

constexpr int TILE_M = 128;
constexpr int CHUNK_K = 64;                    // bf16 columns per SW128 row (128 bytes)
constexpr int CHUNK_BYTES = TILE_M * 128;      // one activation chunk [128 x 64] bf16
constexpr int MAX_MMA_LAYERS = 16;
constexpr int MAX_DOUT = 8;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;  // 320
constexpr int EPI_THREADS = NUM_EPI_WARPS * 32;
constexpr int MAX_STAGES = 8;
constexpr int MAX_CHUNKS = 8;
constexpr int TRACE_LEN = 2048;

struct TcParams {
  const float* x;        // [n][d_x]
  int64_t n;
  int d_x;               // features of x (d_in, or d_in/2 for Delta-UQ)
  int d_in;              // network input features
  int mode;
  int n_tiles;
  int splits;            // member-axis splits (partial moments when > 1)
  int member_begin, member_count, total_members;
  int H, n_tile, NH, KC, K0, split_s, L_mma, d_out;
  int n_stages;          // smem ring depth
  uint32_t stage_bytes;
  int stages_per_member;
  int shared_weights;
  int tmem_cols;
  const __nv_bfloat16* image;
  const float* bias[MAX_MMA_LAYERS];  // [K or 1][H] folded bias per MMA layer
  uint32_t relu_mask, dropout_mask;   // bit l: MMA layer l has ReLU / dropout on its output
  const float* w_last;                // [K or 1][d_out][H]
  const float* b_last;                // [K or 1][d_out]
  int last_relu;
  int drop_mode;                      // 0 none, 1 injected, 2 philox
  float drop_scale;
  uint32_t thr16;
  PhiloxKey key;
  const uint8_t* masks;               // injected base
  const float* anchors;               // [total_members][d_x]
  float* out0;
  float* out1;
  int output;                         // UQ_OUT_*
  float* part_mean;                   // [splits][n*d_out] when splits > 1
  float* part_m2;
  unsigned int* error_flag;
  unsigned long long* prof;          // optional [grid][16] cycle counters (UQ_TC_PROFILE=1)
  int debug_flags;                   // bring-up only (UQ_TC_DEBUG): 1 skip epilogue math, 2 skip MMA issue
  unsigned long long* trace;         // bring-up only (UQ_TC_TRACE): [3 roles][TRACE_LEN][2] (tag, clock) of CTA 0
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (clean CUDA error), never as a
// hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned int* err,
                                          int tag, long long* waited = nullptr) {
  if (waited) {  // profiling build of the wait: the first (potentially blocking) probe counts too
    const long long tp = clock64();
    const bool done = mbar_try_wait(bar, parity);
    long long te;
    asm volatile("{\n\t.reg .b32 t;\n\tmov.b32 t, %1;\n\tmov.u64 %0, %%clock64;\n\t}"
                 : "=l"(te) : "r"((uint32_t)done) : "memory");
    *waited += te - tp;
    if (done) return;
  } else if (mbar_try_wait(bar, parity)) {
    return;
  }
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      if (err) atomicExch(err, 0x80000000u | (unsigned)tag);
      __threadfence_system();
      __trap();
    }
  }
  if (waited) *waited += clock64() - t0;
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols)
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate, issued by one thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(addr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void epi_bar_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
}

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor):
// start address >> 4 in [0,14), LBO (unused for swizzled K-major) = 1 in [16,30),
// SBO = 1024 B (8 rows x 128 B) >> 4 in [32,46), version = 1 in [46,48), layout 2 in [61,64).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16: D=F32, A=B=BF16, K-major.
__host__ __device__ inline uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// byte offset of (row, 16-byte piece) inside a [rows x 64] bf16 SW128 K-major chunk
__host__ __device__ inline uint32_t sw128_offset(int row, int piece) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((piece ^ (row & 7)) << 4));
}

struct __align__(8) Barriers {
  uint64_t w_full[MAX_STAGES];
  uint64_t w_empty[MAX_STAGES];
  uint64_t chunk_done[MAX_CHUNKS];
  uint64_t d_full;
  uint64_t x_ready;
  uint32_t tmem_base;
  uint32_t pad;
};

// input feature i of the network for (sample row, member) -- x, or cat(x - a_k, a_k) for Delta-UQ
__device__ __forceinline__ float net_input(const TcParams& p, int64_t row, int member_global, int i) {
  if (row >= p.n) return 0.f;
  if (p.mode == UQ_MODE_DELTA_UQ) {
    const int d = p.d_x;
    const float a = __ldg(p.anchors + (int64_t)member_global * d + (i < d ? i : i - d));
    return i < d ? __ldg(p.x + row * d + i) - a : a;
  }
  return __ldg(p.x + row * p.d_x + i);
}

// ------------------------------------------------------------------------------------------------
// the fused kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NUM_THREADS, 1)
uq_mlp_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_smem = smem;                                    // KC chunks of 16 KB
  uint8_t* w_smem = smem + (size_t)p.KC * CHUNK_BYTES;       // n_stages x stage_bytes
  Barriers* bars = reinterpret_cast<Barriers*>(w_smem + (size_t)p.n_stages * p.stage_bytes);
  float* xchg = reinterpret_cast<float*>(bars + 1);          // [2][128][d_out] dot exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_units = p.n_tiles * p.splits;
  const bool prof = p.prof != nullptr;
  const long long t_start = prof ? clock64() : 0;
  long long wt[4] = {0, 0, 0, 0};  // waited cycles per barrier kind (profiling only)
  int tr_n = 0;
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;
  auto trace = [&](int role, unsigned tag) {
    if (tracing && tr_n < TRACE_LEN) {
      unsigned long long* t = p.trace + ((size_t)role * TRACE_LEN + tr_n) * 2;
      t[0] = tag;
      t[1] = (unsigned long long)clock64();
      ++tr_n;
    }
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(&bars->w_full[s], 1);
      mbar_init(&bars->w_empty[s], 1);
    }
    for (int c = 0; c < MAX_CHUNKS; ++c) mbar_init(&bars->chunk_done[c], 4);
    mbar_init(&bars->d_full, 1);
    mbar_init(&bars->x_ready, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================================== producer =============================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int split = unit % p.splits;
        const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
        const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
        for (int k = mb; k < me; ++k) {
          const int wslot = p.shared_weights ? 0 : (p.member_begin + k);
          const uint8_t* src = reinterpret_cast<const uint8_t*>(p.image) +
                               (size_t)wslot * p.stages_per_member * p.stage_bytes;
          for (int s = 0; s < p.stages_per_member; ++s, ++it) {
            const int slot = it % p.n_stages;
            const uint32_t par = (it / p.n_stages) & 1;
            mbar_wait(&bars->w_empty[slot], par ^ 1, p.error_flag, 1, prof ? &wt[0] : nullptr);
            trace(0, (1u << 24) | it);
            mbar_arrive_expect_tx(&bars->w_full[slot], p.stage_bytes);
            bulk_g2s(w_smem + (size_t)slot * p.stage_bytes, src + (size_t)s * p.stage_bytes,
                     p.stage_bytes, &bars->w_full[slot]);
            trace(0, (2u << 24) | it);
          }
        }
      }
      if (prof) p.prof[blockIdx.x * 16 + 1] = (unsigned long long)wt[0];  // producer: w_empty
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ===========================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(TILE_M, p.n_tile);
      const uint32_t a_base = smem_u32(a_smem);
      const uint32_t w_base = smem_u32(w_smem);
      uint32_t it = 0;     // weight stage counter
      uint32_t g = 0;      // layer-step counter (d_full / chunk_done phases)
      uint32_t xm = 0;     // member counter (x_ready phase)
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int split = unit % p.splits;
        const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
        const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
        for (int k = mb; k < me; ++k, ++xm) {
          for (int l = 0; l < p.L_mma; ++l, ++g) {
            // chunks of the previous layer-step's epilogue already waited for in this step
            uint32_t waited = (g == 0) ? 0xffffffffu : 0u;
            const uint32_t prev_par = (g - 1) & 1;
            if (l == 0) mbar_wait(&bars->x_ready, xm & 1, p.error_flag, 2, prof ? &wt[0] : nullptr);
            for (int nh = 0; nh < p.NH; ++nh) {
              const int kc_count = (l == 0) ? 1 : p.KC;
              for (int kc = 0; kc < kc_count; ++kc, ++it) {
                // the accumulator columns of this N-half must have been drained, and (for
                // l >= 1) the A chunk kc must have been written, by the previous epilogue
                const int need_hi = ((nh + 1) * p.n_tile + CHUNK_K - 1) / CHUNK_K - 1;
                const int upto = (l == 0) ? need_hi : (kc > need_hi ? kc : need_hi);
                for (int c = 0; c <= upto && c < p.KC; ++c) {
                  if (!(waited & (1u << c))) {
                    mbar_wait(&bars->chunk_done[c], prev_par, p.error_flag, 3,
                              prof ? &wt[1] : nullptr);
                    waited |= 1u << c;
                  }
                }
                const int slot = it % p.n_stages;
                const uint32_t par = (it / p.n_stages) & 1;
                trace(1, (1u << 24) | it);
                mbar_wait(&bars->w_full[slot], par, p.error_flag, 4, prof ? &wt[2] : nullptr);
                trace(1, (2u << 24) | it);
                tc_fence_after();
                const uint32_t a_addr = a_base + (uint32_t)kc * CHUNK_BYTES;
                const uint32_t b_addr = w_base + (uint32_t)slot * p.stage_bytes;
                const int ksteps = (l == 0) ? p.K0 / 16 : CHUNK_K / 16;
                const uint32_t d_addr = tmem_base + (uint32_t)(nh * p.n_tile);
                for (int ks = 0; ks < ksteps && !(p.debug_flags & 2); ++ks) {
                  umma_bf16(d_addr, make_sw128_desc(a_addr + ks * 32),
                            make_sw128_desc(b_addr + ks * 32), idesc,
                            (kc > 0 || ks > 0) ? 1u : 0u);
                }
                if ((p.debug_flags & 6) == 6) mbar_arrive(&bars->w_empty[slot]);  // bring-up probe
                else umma_commit(&bars->w_empty[slot]);  // frees the weight stage when MMAs retire
                trace(1, (3u << 24) | it);
              }
            }
            umma_commit(&bars->d_full);  // whole layer accumulated
          }
        }
      }
      if (prof) {
        p.prof[blockIdx.x * 16 + 2] = (unsigned long long)wt[0];  // MMA: x_ready
        p.prof[blockIdx.x * 16 + 3] = (unsigned long long)wt[1];  // MMA: chunk_done
        p.prof[blockIdx.x * 16 + 4] = (unsigned long long)wt[2];  // MMA: w_full
        p.prof[blockIdx.x * 16 + 0] = (unsigned long long)(clock64() - t_start);
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
    const float keep_scale = p.drop_scale;

    // writes the layer-0 A operand (split input row) into chunk 0, pieces [0, K0/8)
    auto write_x = [&](int tile, int member_global) {
      if (hf == 0) {
        const int64_t grow = (int64_t)tile * TILE_M + row;
        const int d = p.d_in;
        for (int piece = 0; piece < p.K0 / 8; ++piece) {
          uint32_t w4[4];
#pragma unroll
          for (int h2 = 0; h2 < 4; ++h2) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = piece * 8 + h2 * 2 + e;
              const int seg = col / d, i = col - seg * d;
              float out = 0.f;
              if (seg < p.split_s) {
                const float f = net_input(p, grow, member_global, i);
                const float hi = __bfloat162float(__float2bfloat16_rn(f));
                out = (seg == 1) ? (f - hi) : hi;   // [hi | lo | hi]
              }
              v[e] = out;
            }
            w4[h2] = pack_bf16x2(v[0], v[1]);
          }
          *reinterpret_cast<uint4*>(a_smem + sw128_offset(row, piece)) =
              make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->x_ready);
      }
    };

    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const int tile = unit / p.splits, split = unit % p.splits;
      const int mb = (int)(((int64_t)p.member_count * split) / p.splits);
      const int me = (int)(((int64_t)p.member_count * (split + 1)) / p.splits);
      const int64_t grow = (int64_t)tile * TILE_M + row;

      float wf_n = 0.f, wf_mean[MAX_DOUT], wf_m2[MAX_DOUT];
#pragma unroll
      for (int o = 0; o < MAX_DOUT; ++o) wf_mean[o] = 0.f, wf_m2[o] = 0.f;

      if (unit == (int)blockIdx.x) write_x(tile, p.member_begin + mb);  // very first member

      for (int k = mb; k < me; ++k, ++mcount) {
        const int kg = p.member_begin + k;                 // global member / pass id
        const int wslot = p.shared_weights ? 0 : kg;
        float dot[MAX_DOUT];
#pragma unroll
        for (int o = 0; o < MAX_DOUT; ++o) dot[o] = 0.f;
        int drop_ord = 0;
        const uint8_t* mask_layer = p.masks;

        for (int l = 0; l < p.L_mma; ++l, ++g) {
          const bool last = (l == p.L_mma - 1);
          const bool relu = (p.relu_mask >> l) & 1u;
          const bool drop = ((p.dropout_mask >> l) & 1u) && p.drop_mode != 0;
          const float* bias = p.bias[l] + (size_t)wslot * p.H;

          // one lane polls (32 lanes spinning on one mbarrier saturate the barrier unit), the
          // rest of the warp parks on the warp barrier
          if (lane == 0) mbar_wait(&bars->d_full, g & 1, p.error_flag, 5, prof ? &wt[0] : nullptr);
          __syncwarp();
          tc_fence_after();
          if (warp == 2 && lane == 0) trace(2, (1u << 24) | g);

          if (last) {
            // every MMA that reads the A chunks has retired: stage the next member's input row
            int nk = k + 1, ntile = tile, nsplit = split;
            bool have_next = true;
            if (nk >= me) {
              const int nunit = unit + gridDim.x;
              have_next = nunit < n_units;
              ntile = nunit / p.splits;
              nsplit = nunit % p.splits;
              nk = (int)(((int64_t)p.member_count * nsplit) / p.splits);
            }
            const long long tx0 = prof ? clock64() : 0;
            if (have_next) write_x(ntile, p.member_begin + nk);
            if (prof) wt[1] += clock64() - tx0;
          }

          for (int c = hf; c < p.KC; c += 2) {
#pragma unroll 1
            for (int b = 0; b < 2 && !(p.debug_flags & 1); ++b) {
              const int col0 = c * CHUNK_K + b * 32;
              uint32_t acc[32];
              tmem_ld32(lane_addr + (uint32_t)col0, acc);
              tmem_ld_wait();
              uint32_t keep = 0xffffffffu;
              if (drop) {
                if (p.drop_mode == 2) {
                  keep = 0;
#pragma unroll
                  for (int gq = 0; gq < 4; ++gq)
                    keep |= dropout_keep8(p.key, p.thr16, (uint32_t)kg, (uint32_t)drop_ord,
                                          (uint32_t)grow, (uint32_t)(col0 / 8 + gq))
                            << (8 * gq);
                } else {
                  keep = 0;
                  if (grow < p.n) {
                    const uint8_t* mrow =
                        mask_layer + ((size_t)kg * (size_t)p.n + (size_t)grow) * p.H + col0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) keep |= (mrow[j] ? 1u : 0u) << j;
                  }
                }
              }
              float v[32];
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + col0) + j4);
                v[j4 * 4 + 0] = __uint_as_float(acc[j4 * 4 + 0]) + bv.x;
                v[j4 * 4 + 1] = __uint_as_float(acc[j4 * 4 + 1]) + bv.y;
                v[j4 * 4 + 2] = __uint_as_float(acc[j4 * 4 + 2]) + bv.z;
                v[j4 * 4 + 3] = __uint_as_float(acc[j4 * 4 + 3]) + bv.w;
              }
              if (relu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
              }
              if (drop) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = ((keep >> j) & 1u) ? v[j] * keep_scale : 0.f;
              }
              if (!last) {
#pragma unroll
                for (int pc = 0; pc < 4; ++pc) {
                  const uint4 pk = make_uint4(pack_bf16x2(v[pc * 8 + 0], v[pc * 8 + 1]),
                                              pack_bf16x2(v[pc * 8 + 2], v[pc * 8 + 3]),
                                              pack_bf16x2(v[pc * 8 + 4], v[pc * 8 + 5]),
                                              pack_bf16x2(v[pc * 8 + 6], v[pc * 8 + 7]));
                  *reinterpret_cast<uint4*>(a_smem + (size_t)c * CHUNK_BYTES +
                                            sw128_offset(row, b * 4 + pc)) = pk;
                }
              } else {
                const float* wl = p.w_last + (size_t)wslot * p.d_out * p.H + col0;
                for (int o = 0; o < p.d_out; ++o) {
                  float s = dot[o];
#pragma unroll
                  for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 wv = __ldg(reinterpret_cast<const float4*>(wl + (size_t)o * p.H) + j4);
                    s = fmaf(v[j4 * 4 + 0], wv.x, s);
                    s = fmaf(v[j4 * 4 + 1], wv.y, s);
                    s = fmaf(v[j4 * 4 + 2], wv.z, s);
                    s = fmaf(v[j4 * 4 + 3], wv.w, s);
                  }
                  dot[o] = s;
                }
              }
            }
            // chunk c: accumulator columns drained (+ A chunk rewritten) -> release to the MMA warp
            tc_fence_before();
            if (!last) fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->chunk_done[c]);
            if (warp == 2 && lane == 0) trace(2, (2u << 24) | (g << 4) | c);
          }
          // a warp whose parity has no chunk (KC == 1) still owes nothing: barrier counts 4
          if (((p.dropout_mask >> l) & 1u)) {
            if (p.masks) mask_layer += (size_t)p.total_members * (size_t)p.n * (size_t)p.H;
            ++drop_ord;
          }
        }

        // ---- combine the two column-parity halves of the dot products, then Welford ----------
        float* xb = xchg + (size_t)(mcount & 1) * TILE_M * p.d_out;
        if (hf == 1) {
          for (int o = 0; o < p.d_out; ++o) xb[row * p.d_out + o] = dot[o];
        }
        const long long tb0 = prof ? clock64() : 0;
        epi_bar_sync();
        if (prof) wt[2] += clock64() - tb0;
        if (hf == 0) {
          wf_n += 1.f;
          const float inv_n = 1.f / wf_n;
          const float* bl = p.b_last + (size_t)wslot * p.d_out;
          for (int o = 0; o < p.d_out; ++o) {
            float y = dot[o] + (p.KC > 1 ? xb[row * p.d_out + o] : 0.f) + __ldg(bl + o);
            if (p.last_relu) y = fmaxf(y, 0.f);
            const float dlt = y - wf_mean[o];
            wf_mean[o] += dlt * inv_n;
            wf_m2[o] = fmaf(dlt, y - wf_mean[o], wf_m2[o]);
          }
        }
      }

      // ---- tile done: publish (mean, std) / (mean, M2) ------------------------------------------
      if (hf == 0 && grow < p.n) {
        for (int o = 0; o < p.d_out; ++o) {
          const int64_t idx = grow * p.d_out + o;
          if (p.splits > 1) {
            p.part_mean[(size_t)split * (size_t)p.n * p.d_out + idx] = wf_mean[o];
            p.part_m2[(size_t)split * (size_t)p.n * p.d_out + idx] = wf_m2[o];
          } else {
            p.out0[idx] = wf_mean[o];
            p.out1[idx] = (p.output == UQ_OUT_MOMENTS) ? wf_m2[o] : sqrtf(wf_m2[o] / (wf_n - 1.f));
          }
        }
      }
    }
  }

  if (prof && warp >= 2 && lane == 0 && (warp == 2 || warp == 6)) {
    const int o = warp == 2 ? 5 : 9;
    p.prof[blockIdx.x * 16 + o + 0] = (unsigned long long)wt[0];  // epilogue: d_full wait
    p.prof[blockIdx.x * 16 + o + 1] = (unsigned long long)wt[1];  // epilogue: write_x
    p.prof[blockIdx.x * 16 + o + 2] = (unsigned long long)wt[2];  // epilogue: pair barrier
    p.prof[blockIdx.x * 16 + o + 3] = (unsigned long long)(clock64() - t_start);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
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

size_t tc_smem_bytes(const TcPlan& t, int n_stages) {
  const int KC = t.hidden / 64;
  return 1024 + (size_t)KC * CHUNK_BYTES + (size_t)n_stages * t.stage_bytes + sizeof(Barriers) +
         2 * TILE_M * (size_t)t.d_out * sizeof(float) + 64;
}

int pick_stages(const TcPlan& t) {
  int s = MAX_STAGES;
  while (s > 2 && tc_smem_bytes(t, s) > 232448) --s;
  return s;
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
  if (H % 64 != 0 || H < 64 || H > 512) {
    t.why_not = "hidden width must be a multiple of 64 in [64, 512] (got " + std::to_string(H) + ")";
    return;
  }
  const int NH = (H + 255) / 256;
  const int n_tile = H / NH;
  if (n_tile % 16 != 0) { t.why_not = "hidden width not tileable"; return; }
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
  const Layer& last = m->layers[m->n_layers - 1];
  t.w_last = last.w;
  t.b_last = last.bias_folded;
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
  size_t b = 256 + 148 * 4 * 16 * sizeof(unsigned long long);  // error flag + profile counters
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
  p.H = t.hidden;
  p.n_tile = t.n_tile;
  p.NH = t.hidden / t.n_tile;
  p.KC = t.hidden / 64;
  p.K0 = t.k0;
  p.split_s = split_factor(t);
  p.L_mma = t.n_mma_layers;
  p.d_out = t.d_out;
  p.n_stages = pick_stages(t);
  p.stage_bytes = (uint32_t)t.stage_bytes;
  p.stages_per_member = t.stages_per_member;
  p.shared_weights = (a->mode != UQ_MODE_ENSEMBLE) ? 1 : 0;
  int cols = 32;
  while (cols < t.hidden) cols *= 2;
  p.tmem_cols = cols;
  p.image = t.image;
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
  const size_t prof_bytes = 148 * 4 * 16 * sizeof(unsigned long long);
  const char* dbg_env = getenv("UQ_TC_DEBUG");
  p.debug_flags = dbg_env ? atoi(dbg_env) : 0;
  const char* prof_env = getenv("UQ_TC_PROFILE");
  const bool want_prof = prof_env && prof_env[0] == '1';
  p.prof = want_prof ? reinterpret_cast<unsigned long long*>(wsb + 256) : nullptr;
  if (p.splits > 1) {
    const size_t part = (((size_t)p.splits * (size_t)n * m->d_out * sizeof(float)) + 255) & ~(size_t)255;
    p.part_mean = reinterpret_cast<float*>(wsb + 256 + prof_bytes);
    p.part_m2 = reinterpret_cast<float*>(wsb + 256 + prof_bytes + part);
  }
  UQ_CUDA(cudaMemsetAsync(p.error_flag, 0, sizeof(unsigned int), st));
  const char* trace_env = getenv("UQ_TC_TRACE");
  unsigned long long* d_trace = nullptr;
  if (trace_env && trace_env[0]) {
    UQ_CUDA(cudaMalloc(&d_trace, 3 * TRACE_LEN * 2 * sizeof(unsigned long long)));
    UQ_CUDA(cudaMemsetAsync(d_trace, 0, 3 * TRACE_LEN * 2 * sizeof(unsigned long long), st));
  }
  p.trace = d_trace;

  const size_t smem = tc_smem_bytes(t, p.n_stages);
  UQ_REQUIRE(smem <= 232448, UQ_ERR_UNSUPPORTED, "bf16 kernel needs %zu bytes of shared memory",
             smem);
  UQ_CUDA(cudaFuncSetAttribute(uq_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, uq_mlp_tc_kernel, NUM_THREADS, smem);
  if (per_sm < 1) per_sm = 1;
  if (per_sm * p.tmem_cols > 512) per_sm = 512 / p.tmem_cols;  // TMEM columns are per SM
  if (per_sm > 2) per_sm = 2;
  const int64_t units = (int64_t)p.n_tiles * p.splits;
  int grid = (int)(units < (int64_t)sms * per_sm ? units : (int64_t)sms * per_sm);
  uq_mlp_tc_kernel<<<grid, NUM_THREADS, smem, st>>>(p);
  UQ_LAUNCH_CHECK();
  if (d_trace) {  // bring-up aid: dump CTA 0's event timeline as CSV (role, tag, clock)
    std::vector<unsigned long long> h(3 * TRACE_LEN * 2);
    UQ_CUDA(cudaMemcpyAsync(h.data(), d_trace, h.size() * sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost, st));
    UQ_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_trace);
    FILE* f = fopen(trace_env, "w");
    if (f) {
      for (int r = 0; r < 3; ++r)
        for (int i = 0; i < TRACE_LEN; ++i) {
          const unsigned long long tag = h[((size_t)r * TRACE_LEN + i) * 2];
          if (tag) fprintf(f, "%d,%llu,%llu,%llu\n", r, tag >> 24, tag & 0xFFFFFF,
                           h[((size_t)r * TRACE_LEN + i) * 2 + 1]);
        }
      fclose(f);
    }
  }
  if (want_prof) {  // debugging aid: synchronises and prints the per-role wait breakdown
    std::vector<unsigned long long> h((size_t)grid * 16);
    UQ_CUDA(cudaMemcpyAsync(h.data(), p.prof, h.size() * sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost, st));
    UQ_CUDA(cudaStreamSynchronize(st));
    double a[16] = {0};
    for (int b = 0; b < grid; ++b)
      for (int i = 0; i < 16; ++i) a[i] += (double)h[(size_t)b * 16 + i] / grid;
    fprintf(stderr,
            "[uq_tc_profile] grid %d cycles/CTA: total %.0f | producer wait w_empty %.0f | MMA wait "
            "x_ready %.0f chunk_done %.0f w_full %.0f | epi(w2) wait d_full %.0f write_x %.0f pairbar "
            "%.0f total %.0f | epi(w6) wait d_full %.0f pairbar %.0f\n",
            grid, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[11]);
  }
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
