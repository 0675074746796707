// Host side of the bf16 tcgen05 path of the UQ forward: eligibility (tc_plan), the packed weight
// image (tc_pack) and launch parameters / dispatch (tc_forward).  The kernels live in
//   mlp_tc2.cu  CTA pairs, 128 sample rows per CTA, hidden width 64 .. 512 (the throughput path)
//   mlp_tc3.cu  CTA pairs,  64 sample rows per CTA, hidden width 768 / 1024
// and share one weight image: per member, the stages [N x 64] bf16 of every MMA layer in the order
// the MMA warp consumes them, already in the UMMA K-major SWIZZLE_128B shared-memory layout (8-row
// x 128-byte atoms) with eval-BatchNorm folded in, so a producer warp moves a stage -- or the
// N/2-row half one CTA of a pair needs -- with a single linear cp.async.bulk.
//
// Layer 0 (d_in <= 21 inputs) is also an MMA: x is split into bf16 hi + lo parts and the folded
// first-layer weights likewise, laid out as [x_hi | x_lo | x_hi] . [w_hi | w_hi | w_lo] inside
// one K <= 64 chunk, which recovers ~16 mantissa bits for one extra K = 16 step.
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

// ------------------------------------------------------------------------------------------------
// weight image packing
// ------------------------------------------------------------------------------------------------
// Writes one member's stages.  Stage order == consumption order of the kernel:
//   layer 0: NH stages (K0 real columns, [w_hi | w_hi | w_lo] split), layer l>=1: NH x KC stages.
__global__ void pack_image_kernel(__nv_bfloat16* __restrict__ image, const float* __restrict__ w,
                                  const float* __restrict__ alpha, int layer, int in, int ld, int H,
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
        const float f = w[(int64_t)n * ld + ii] * sc;
        const float hi = __bfloat162float(__float2bfloat16_rn(f));
        out = (seg == 2) ? (f - hi) : hi;        // [hi | hi | lo]
      }
    } else {
      const int kk = kc * 64 + col;
      out = w[(int64_t)n * ld + kk] * sc;
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

// Bias stages of the bias-in-the-MMA variants (mlp_tc2.cu, mlp_tc4.cu): stage (k, l, nh) is a
// [n_tile x 64] SW128 stage, zero except that row r carries the folded bias of output feature
// nh * n_tile + r as three bf16 pieces (hi + mid + lo == the fp32 bias to 24 bits) in K columns
// 0..2; the kernel multiplies it by an all-ones A tile as one K = 16 step, which leaves acc + bias
// in the accumulator.  Rows k_begin .. k_begin + K - 1 of bias [..][H] -> stages of the same k; one
// thread per (k, feature, 16-byte piece) writes the whole stage (no memset needed).
__global__ void pack_bias_image_kernel(__nv_bfloat16* __restrict__ image,
                                       const float* __restrict__ bias, int k_begin, int K, int l,
                                       int L_mma, int H, int n_tile, size_t stage_elems) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)K * H * 8) return;
  const int piece = (int)(i & 7);
  const int n = (int)((i >> 3) % H);
  const int k = k_begin + (int)((i >> 3) / H);
  const int nh = n / n_tile, r = n % n_tile;
  const int NH = H / n_tile;
  __nv_bfloat16 vals[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) vals[e] = __float2bfloat16_rn(0.f);
  if (piece == 0) {
    const float b = bias[(size_t)k * H + n];
    vals[0] = __float2bfloat16_rn(b);
    const float r1 = b - __bfloat162float(vals[0]);
    vals[1] = __float2bfloat16_rn(r1);
    vals[2] = __float2bfloat16_rn(r1 - __bfloat162float(vals[1]));
  }
  uint8_t* dst =
      reinterpret_cast<uint8_t*>(image + (((size_t)k * L_mma + l) * NH + nh) * stage_elems) +
      sw128_offset(r, piece);
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(vals);
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
  t.k0_delta = 0;
  if (m->n_members == 1 && d_in % 2 == 0) {
    const int dx = d_in / 2;
    int sd = 3;
    while (sd > 1 && sd * dx > 64) --sd;
    t.k0_delta = ((sd * dx + 15) / 16) * 16;
  }
  t.ok = true;
}

static int split_factor_of(int d_in) {
  int s = 3;
  while (s > 1 && s * d_in > 64) --s;
  return s;
}
static int split_factor(const TcPlan& t) { return split_factor_of(t.d_in); }

// Delta-UQ: bias0[k][h] = b0'[h] + sum_i (W0'[h][dx + i] - W0'[h][i]) a_k[i]   (W0', b0' = layer 0
// with eval-BatchNorm folded), so that  W0' [x - a_k; a_k] + b0' = W0'[:, :dx] x + bias0[k]
// PAGER (pager = 1) swaps the roles:  W0' [a_k - x; a... x] = (W0'[:, dx:] - W0'[:, :dx]) x + bias0[k]
// with bias0[k][h] = b0'[h] + sum_i W0'[h][i] a_k[i].
// diff_at: first column of the half of W0 that multiplies the difference (0, or dx with
// UQ_MODEL_ANCHOR_FIRST); "W0a" above is that half, "W0b" the other one.
__global__ void delta_bias_kernel(const float* __restrict__ w0, const float* __restrict__ alpha,
                                  const float* __restrict__ bias_folded,
                                  const float* __restrict__ anchors, int H, int dx, int n_anchors,
                                  int anchor_begin, int pager, int diff_at,
                                  float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_anchors * H) return;
  const int h = i % H, k = anchor_begin + i / H;
  const float sc = alpha ? alpha[h] : 1.0f;
  float acc = bias_folded[h];
  for (int j = 0; j < dx; ++j) {
    const float wa = w0[(int64_t)h * 2 * dx + diff_at + j];
    const float wb = w0[(int64_t)h * 2 * dx + (dx - diff_at) + j];
    acc = fmaf((pager ? wa : wb - wa) * sc, anchors[(int64_t)k * dx + j], acc);
  }
  out[(int64_t)k * H + h] = acc;
}

// [H][dx] column differences W0[:, dx:] - W0[:, :dx] (the PAGER image's layer 0)
__global__ void column_diff_kernel(const float* __restrict__ w0, int H, int dx, int diff_at,
                                   float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * dx) return;
  const int h = i / dx, j = i % dx;
  out[i] = w0[(int64_t)h * 2 * dx + (dx - diff_at) + j] - w0[(int64_t)h * 2 * dx + diff_at + j];
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
          ly.has_bn ? ly.alpha + (size_t)k * ly.out : nullptr, l, ly.in, ly.in, H, t.n_tile, NH,
          KC, t.k0, s, stage_elems, stage_base);
      UQ_LAUNCH_CHECK();
      stage_base += stages;
    }
  }
  if (t.k0_delta > 0) {
    // same stages, except that layer 0 keeps only the d_in/2 columns that multiply x
    void* pd = nullptr;
    UQ_CUDA(cudaMalloc(&pd, member_elems * sizeof(__nv_bfloat16)));
    m->allocations.push_back(pd);
    t.image_delta = static_cast<__nv_bfloat16*>(pd);
    UQ_CUDA(cudaMemcpyAsync(pd, t.image, member_elems * sizeof(__nv_bfloat16),
                            cudaMemcpyDeviceToDevice, st));
    UQ_CUDA(cudaMemsetAsync(pd, 0, (size_t)NH * t.stage_bytes, st));
    const Layer& l0 = m->layers[0];
    const int dx = l0.in / 2;
    const int64_t total = (int64_t)NH * t.n_tile * 8;
    const int diff_at = m->anchor_first ? dx : 0;
    pack_image_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        t.image_delta, l0.w + diff_at, l0.has_bn ? l0.alpha : nullptr, 0, dx, l0.in, H, t.n_tile,
        NH, KC, t.k0_delta, split_factor_of(dx), stage_elems, 0);
    UQ_LAUNCH_CHECK();
    // PAGER image: layer 0 = column differences, everything else shared with the Delta-UQ image
    void *pp = nullptr, *diff = nullptr;
    UQ_CUDA(cudaMalloc(&pp, member_elems * sizeof(__nv_bfloat16)));
    m->allocations.push_back(pp);
    UQ_CUDA(cudaMalloc(&diff, sizeof(float) * (size_t)H * dx));
    m->allocations.push_back(diff);
    t.image_pager = static_cast<__nv_bfloat16*>(pp);
    UQ_CUDA(cudaMemcpyAsync(pp, pd, member_elems * sizeof(__nv_bfloat16), cudaMemcpyDeviceToDevice,
                            st));
    UQ_CUDA(cudaMemsetAsync(pp, 0, (size_t)NH * t.stage_bytes, st));
    column_diff_kernel<<<(H * dx + 255) / 256, 256, 0, st>>>(l0.w, H, dx, diff_at,
                                                              static_cast<float*>(diff));
    UQ_LAUNCH_CHECK();
    pack_image_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        t.image_pager, static_cast<const float*>(diff), l0.has_bn ? l0.alpha : nullptr, 0, dx, dx, H,
        t.n_tile, NH, KC, t.k0_delta, split_factor_of(dx), stage_elems, 0);
    UQ_LAUNCH_CHECK();
  }
  for (int l = 0; l < m->n_layers; ++l) {
    const Layer& ly = m->layers[l];
    const int n = ly.out * K;
    fold_bias_kernel<<<(n + 255) / 256, 256, 0, st>>>(ly.bias, ly.has_bn ? ly.alpha : nullptr,
                                                     ly.has_bn ? ly.beta : nullptr,
                                                     ly.bias_folded, n);
    UQ_LAUNCH_CHECK();
  }
  {   // bias stages of the bias-in-the-MMA variants
    void* pb = nullptr;
    const size_t bias_elems = (size_t)K * t.n_mma_layers * NH * stage_elems;
    UQ_CUDA(cudaMalloc(&pb, bias_elems * sizeof(__nv_bfloat16)));
    m->allocations.push_back(pb);
    t.bias_image = static_cast<__nv_bfloat16*>(pb);
    for (int l = 0; l < t.n_mma_layers; ++l) {
      pack_bias_image_kernel<<<(unsigned)(((int64_t)K * H * 8 + 255) / 256), 256, 0, st>>>(
          t.bias_image, m->layers[l].bias_folded, 0, K, l, t.n_mma_layers, H, t.n_tile, stage_elems);
      UQ_LAUNCH_CHECK();
    }
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

static int choose_splits(const uq_model* m, int64_t n, const uq_forward_args* a, bool split) {
  // sample rows one cluster (CTA pair) works on at a time
  const bool narrow = tc4_supported(m->tc.hidden, dout_pad(m->tc.d_out)) && !getenv("UQ_TC_NO_SLOTS");
  const bool narrow_x = tcx4_supported(m->tc.hidden, dout_pad(m->tc.d_out)) && !getenv("UQ_TC_NO_SLOTS");
  const int64_t rows = split ? (narrow_x ? tcx4_rows_per_unit() : tcx_rows_per_unit())
                       : m->tc.hidden > 512 ? TILE_M : narrow ? tc4_rows_per_unit() : 2 * TILE_M;
  const int64_t units = (n + rows - 1) / rows;
  int splits = 1;
  if (a->output == UQ_OUT_MOMENTS) return 1;  // a K-shard hands raw moments to the caller
  if (a->mode == UQ_MODE_PAGER) return 1;      // the max over anchors is not a moment merge
  // (1) Ensembles whose weights exceed the L2 (each die caches its own copy: ~63 MB): cut the
  // members into ~32 MB groups, which mlp_tc3.cu / mlp_tcx.cu walk split-major -- every cluster on
  // the same group at the same time (ensemble8x1024_256k: 2.86 GB of DRAM reads per launch and
  // 26.7 ms -> 22.7 ms).
  const bool group_kernel = split ? !narrow_x : m->tc.hidden > 512;   // mlp_tcx.cu / mlp_tc3.cu
  if (group_kernel && a->mode == UQ_MODE_ENSEMBLE) {
    const double member_mb = split
        ? (double)m->tc.x_stages_per_member * (double)m->tc.x_stage_bytes / 1048576.0
        : (double)m->tc.stages_per_member * (double)m->tc.stage_bytes / 1048576.0;
    if (member_mb * a->member_count > 48.0) {
      int s = (int)((member_mb * a->member_count + 31.9) / 32.0);
      if (s > a->member_count) s = a->member_count;
      if (s > 64) s = 64;
      if (s > 1 && units >= 74) return s;
    }
  }
  // (2) Split the member axis over clusters when that shortens the schedule: one launch is
  // ceil(units * s / 74) rounds of the 74 SM pairs, a round costs members / s member-forwards plus
  // a fixed per-unit part (input staging, Welford write-back; ~half a member-forward).  Few sample
  // tiles with many passes spread out (10 units x 100 passes -> 7 splits = 70 of the 74 pairs, one
  // round), and a ragged last round is evened out (256 units x 1000 passes: 4 rounds -> 7
  // half-rounds, -12 %).
  double best = 0.0;
  for (int s = 1; s <= 64; ++s) {
    if (s > 1 && a->member_count / s < 4) break;
    const double rounds = (double)((units * s + 73) / 74);
    const double cost = rounds * ((double)((a->member_count + s - 1) / s) + 0.5);
    if (s == 1 || cost < best * 0.97) best = cost, splits = s;   // a split must buy >= 3 %
  }
  return splits;
}

size_t tc_workspace_bytes(const uq_model* m, int64_t n, const uq_forward_args* a, bool split) {
  const int splits = choose_splits(m, n, a, split);
  size_t b = 256;  // error flag
  if (splits > 1) b += 2 * (size_t)splits * (size_t)n * m->d_out * sizeof(float) + 512;
  if (a->mode == UQ_MODE_DELTA_UQ || a->mode == UQ_MODE_PAGER) {  // per-anchor layer-0 bias [K][H]
    b += (((size_t)a->total_members * m->tc.hidden * sizeof(float)) + 255) & ~(size_t)255;
    if (split) b += (((size_t)a->total_members * sizeof(float)) + 255) & ~(size_t)255;  // max |bias0[k]|
    // per-anchor layer-0 bias stages of the bias-in-the-MMA variants
    else if (m->tc.bias_image) b += (size_t)a->total_members * (m->tc.hidden / m->tc.n_tile) * m->tc.stage_bytes;
  }
  return b;
}

// max |v[k][:]| of every row k in [begin, begin + count): the per-anchor term of the split mode's
// output bound (mlp_tcx.cu)
__global__ void row_absmax_kernel(const float* __restrict__ v, int H, int begin,
                                  float* __restrict__ out) {
  const int k = begin + blockIdx.x;
  float m = 0.f;
  for (int h = threadIdx.x; h < H; h += blockDim.x) m = fmaxf(m, fabsf(v[(int64_t)k * H + h]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
    out[k] = m;
  }
}

int tc_forward(const uq_model* m, const float* x, int64_t n, const uq_forward_args* a,
               float* out0, float* out1, void* ws, size_t ws_bytes, cudaStream_t st, bool split) {
  const TcPlan& t = m->tc;
  const size_t need = tc_workspace_bytes(m, n, a, split);
  UQ_REQUIRE(ws != nullptr && ws_bytes >= need, UQ_ERR_WORKSPACE,
             "tensor-core forward needs %zu workspace bytes, got %zu", need, ws_bytes);
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.x = x;
  p.n = n;
  p.row_base = a->row_base;
  p.d_in = m->d_in;
  const bool anchored = a->mode == UQ_MODE_DELTA_UQ || a->mode == UQ_MODE_PAGER;
  const bool pager = a->mode == UQ_MODE_PAGER;
  p.d_x = anchored ? m->d_in / 2 : m->d_in;
  p.mode = a->mode;
  p.n_tiles = (int)((n + TILE_M - 1) / TILE_M);
  p.splits = choose_splits(m, n, a, split);
  {
    const bool nx = tcx4_supported(t.hidden, dout_pad(t.d_out)) && !getenv("UQ_TC_NO_SLOTS");
    const bool group_kernel = split ? !nx : t.hidden > 512;
    p.split_major = (group_kernel && a->mode == UQ_MODE_ENSEMBLE && p.splits > 1) ? 1 : 0;
  }
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
  // bias stages of the bias-in-the-MMA variants (mlp_tc2.cu, mlp_tc4.cu; d_out 1)
  p.bias_image = !split ? reinterpret_cast<const uint8_t*>(t.bias_image) : nullptr;
  if (split) {  // fp32-parity split mode: its own image, three input segments always
    p.K0 = ((3 * t.d_in + 15) / 16) * 16 + 16;   // + the all-zero step (see mlp_tcx.cu)
    p.split_s = 3;
    p.stages_per_member = t.x_stages_per_member;
    p.image = reinterpret_cast<const uint8_t*>(t.x_image);
    p.lstats = t.x_stats;
  }
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
  if (pager) p.targets = a->anchor_targets, p.score_floor = a->score_floor;
  if (anchored) {
    // anchors -> per-anchor layer-0 bias; the kernels then see plain x (d_in / 2 inputs), shared
    // weights and no per-member input rebuild
    UQ_REQUIRE(t.image_delta != nullptr, UQ_ERR_UNSUPPORTED,
               "Delta-UQ bf16 path needs a single packed network with an even input width");
    size_t off = 256;
    if (p.splits > 1)
      off += 2 * ((((size_t)p.splits * (size_t)n * m->d_out * sizeof(float)) + 255) & ~(size_t)255);
    float* bias0 = reinterpret_cast<float*>(wsb + off);
    const Layer& l0 = m->layers[0];
    const int dx = m->d_in / 2;
    const int total = a->member_count * t.hidden;
    delta_bias_kernel<<<(total + 255) / 256, 256, 0, st>>>(
        l0.w, l0.has_bn ? l0.alpha : nullptr, l0.bias_folded, a->anchors, t.hidden, dx,
        a->member_count, a->member_begin, pager ? 1 : 0, m->anchor_first ? dx : 0, bias0);
    UQ_LAUNCH_CHECK();
    p.bias[0] = bias0;
    p.bias0_per_member = 1;
    p.image = reinterpret_cast<const uint8_t*>(pager ? t.image_pager : t.image_delta);
    p.d_in = dx;
    p.K0 = t.k0_delta;
    p.split_s = split_factor_of(dx);
    if (split) {
      float* bmax0 = reinterpret_cast<float*>(
          wsb + off + ((((size_t)a->total_members * t.hidden * sizeof(float)) + 255) & ~(size_t)255));
      row_absmax_kernel<<<a->member_count, 128, 0, st>>>(bias0, t.hidden, a->member_begin, bmax0);
      UQ_LAUNCH_CHECK();
      p.bmax0 = bmax0;
      p.image = reinterpret_cast<const uint8_t*>(pager ? t.x_image_pager : t.x_image_delta);
      p.lstats = pager ? t.x_stats_pager : t.x_stats_delta;
      p.K0 = ((3 * dx + 15) / 16) * 16 + 16;
      p.split_s = 3;
    } else if (p.bias_image != nullptr) {
      // the per-anchor biases as layer-0 bias stages (indexed by the global member id)
      __nv_bfloat16* b0img = reinterpret_cast<__nv_bfloat16*>(
          wsb + off + ((((size_t)a->total_members * t.hidden * sizeof(float)) + 255) & ~(size_t)255));
      const int64_t threads = (int64_t)a->member_count * t.hidden * 8;
      pack_bias_image_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
          b0img, bias0, a->member_begin, a->member_count, 0, 1, t.hidden, t.n_tile,
          t.stage_bytes / sizeof(__nv_bfloat16));
      UQ_LAUNCH_CHECK();
      p.bias0_image = reinterpret_cast<const uint8_t*>(b0img);
    }
    p.mode = UQ_MODE_MC_DROPOUT;   // shared weights, members differ only in bias0
    p.anchors = nullptr;
  }
#ifdef UQ_TC_TRACE
  const char* trace_env = getenv("UQ_TC_TRACE_FILE");
  unsigned long long* d_trace = nullptr;
  if (trace_env && trace_env[0]) {
    UQ_CUDA(cudaMalloc(&d_trace, TRACE_ROLES * TRACE_LEN * 2 * sizeof(unsigned long long)));
    UQ_CUDA(cudaMemsetAsync(d_trace, 0, TRACE_ROLES * TRACE_LEN * 2 * sizeof(unsigned long long), st));
  }
  p.trace = d_trace;
#endif

  const bool narrow = tc4_supported(t.hidden, dout_pad(t.d_out)) && !getenv("UQ_TC_NO_SLOTS");
  const bool narrow_x = tcx4_supported(t.hidden, dout_pad(t.d_out)) && !getenv("UQ_TC_NO_SLOTS");
  const int rc = split && narrow_x ? tcx4_launch(p, t.hidden, st)
                 : split          ? tcx_launch(p, t.hidden, dout_pad(t.d_out), st)
                 : t.hidden > 512 ? tc3_launch(p, t.hidden, dout_pad(t.d_out), st)
                 : narrow         ? tc4_launch(p, t.hidden, st)
                                  : tc2_launch(p, t.hidden, dout_pad(t.d_out), st);
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
