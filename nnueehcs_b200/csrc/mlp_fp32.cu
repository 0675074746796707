// fp32 parity path of the UQ forward: CUDA-core FFMA, fp32 accumulate.
//
// This is the arithmetic the reference performs (eager fp32 Linear / BatchNorm1d(eval) / ReLU /
// Dropout per member or pass, nnueehcs/models.py:103,156-158, then stack/mean/std :105-107,
// :159-162) restructured for the GPU:
//   * all K members (or passes / anchors) of a sample chunk go through one grouped GEMM launch
//     per layer (blockIdx.z = member); Linear bias, eval-BN (x*alpha+beta, the same op order as
//     ATen's batch_norm_cpu_transform_input), ReLU and the dropout mask are the GEMM epilogue;
//   * sample chunks are sized so the [K][chunk][width] activations of two consecutive layers
//     stay L2-resident (126 MB) instead of round-tripping HBM;
//   * the [K, N, out] stack never exists in full: the last Linear writes a [K][chunk][out] slab
//     and a finalise kernel reduces it to (mean, unbiased std) or (mean, M2) in float64.
// The bf16 tcgen05 kernel in mlp_tc.cu is the throughput path; this file is the 1e-5 mode.
#include "common.cuh"
#include "philox.cuh"

namespace uq {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, PADM = BM + 4;
constexpr int GEMM_THREADS = 256;

struct Epilogue {
  const float* bias;   // [K][N] (member-major) or nullptr
  const float* alpha;  // eval-BN scale or nullptr
  const float* beta;
  int relu;
  int drop_mode;       // 0 none, 1 injected byte masks, 2 native philox
  const uint8_t* mask; // injected: base of this dropout layer's [total_members][n_total][N] block
  int64_t n_total;     // samples of the whole call (mask row stride)
  int64_t sample0;     // index (inside this call) of this chunk's first sample
  int64_t philox_row0; // global index of the call's first row (Philox sample counter)
  int pass0;           // global id of this launch's first member (blockIdx.z = 0)
  int drop_layer;      // dropout layer ordinal (philox counter)
  float drop_scale;    // fl32(1 / fl32(1-p))
  uint32_t thr16;
  PhiloxKey key;
};

// out[z][m][n] = epi( sum_k A[z][m][k] * W[z][n][k] ),  A and W both K-contiguous.
template <bool VEC>
__global__ void __launch_bounds__(GEMM_THREADS)
sgemm_tn_kernel(const float* __restrict__ A, int64_t a_member_stride, int lda,
                const float* __restrict__ W, int64_t w_member_stride,
                float* __restrict__ C, int64_t c_member_stride, int ldc,
                int M, int N, int Kd, int member0, int member_step, Epilogue ep) {
  __shared__ __align__(16) float As[2][BK][PADM];
  __shared__ __align__(16) float Bs[2][BK][PADM];

  const int z = blockIdx.z;
  const int member = member0 + z * member_step;  // weight/bias slot (0 when weights are shared)
  A += (int64_t)z * a_member_stride;
  W += (int64_t)member * w_member_stride;
  C += (int64_t)z * c_member_stride;

  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int lrow = t >> 2;         // 0..63
  const int lk = (t & 3) * 4;      // 0,4,8,12

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int gm = m0 + lrow + h * 64, gn = n0 + lrow + h * 64, gk = k0 + lk;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (gm < M) {
        const float* p = A + (int64_t)gm * lda + gk;
        if (VEC) {
          if (gk < Kd) va = *reinterpret_cast<const float4*>(p);
        } else {
          if (gk + 0 < Kd) va.x = p[0];
          if (gk + 1 < Kd) va.y = p[1];
          if (gk + 2 < Kd) va.z = p[2];
          if (gk + 3 < Kd) va.w = p[3];
        }
      }
      if (gn < N) {
        const float* p = W + (int64_t)gn * Kd + gk;
        if (VEC) {
          if (gk < Kd) vb = *reinterpret_cast<const float4*>(p);
        } else {
          if (gk + 0 < Kd) vb.x = p[0];
          if (gk + 1 < Kd) vb.y = p[1];
          if (gk + 2 < Kd) vb.z = p[2];
          if (gk + 3 < Kd) vb.w = p[3];
        }
      }
      ra[h] = va;
      rb[h] = vb;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lrow + h * 64;
      As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y;
      As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
      Bs[buf][lk + 0][r] = rb[h].x; Bs[buf][lk + 1][r] = rb[h].y;
      Bs[buf][lk + 2][r] = rb[h].z; Bs[buf][lk + 3][r] = rb[h].w;
    }
  };

  const int nk = (Kd + BK - 1) / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue: bias -> eval-BN -> ReLU -> dropout ------------------------------------------
  const float* bias = ep.bias ? ep.bias + (int64_t)member * N : nullptr;
  const float* alpha = ep.alpha ? ep.alpha + (int64_t)member * N : nullptr;
  const float* beta = ep.beta ? ep.beta + (int64_t)member * N : nullptr;
  const uint32_t pass = (uint32_t)(ep.pass0 + z);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
    const int64_t sample = ep.sample0 + gm;
    // native masks: the thread's two runs of four columns sit in one 32-feature group each
    uint32_t keep_lo = 0, keep_hi = 0;
    if (ep.drop_mode == 2) {
      const uint32_t prow = (uint32_t)(sample + ep.philox_row0);
      keep_lo = dropout_keep32(ep.key, ep.thr16, pass, (uint32_t)ep.drop_layer, prow,
                               (uint32_t)((n0 + tx * 4) >> 5));
      keep_hi = dropout_keep32(ep.key, ep.thr16, pass, (uint32_t)ep.drop_layer, prow,
                               (uint32_t)((n0 + 64 + tx * 4) >> 5));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[gn];
      if (alpha) v = v * alpha[gn] + beta[gn];
      if (ep.relu) v = fmaxf(v, 0.f);
      if (ep.drop_mode == 1) {
        const uint8_t keep = ep.mask[((int64_t)pass * ep.n_total + sample) * N + gn];
        v = v * (keep ? ep.drop_scale : 0.f);
      } else if (ep.drop_mode == 2) {
        const bool keep = ((j < 4 ? keep_lo : keep_hi) >> (gn & 31)) & 1u;
        v = v * (keep ? ep.drop_scale : 0.f);
      }
      C[(int64_t)gm * ldc + gn] = v;
    }
  }
}

// Last Linear with a handful of outputs: one warp per (member, sample) row, shuffle reduction.
__global__ void __launch_bounds__(256)
linear_small_out_kernel(const float* __restrict__ A, int64_t a_member_stride, int lda,
                        const float* __restrict__ W, int64_t w_member_stride,
                        const float* __restrict__ bias, float* __restrict__ C,
                        int64_t c_member_stride, int M, int N, int Kd, int member0,
                        int member_step, int relu) {
  const int z = blockIdx.z, member = member0 + z * member_step;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row >= M) return;
  const float* a = A + (int64_t)z * a_member_stride + row * lda;
  const float* w = W + (int64_t)member * w_member_stride;
  for (int n = 0; n < N; ++n) {
    float s = 0.f;
    for (int k = lane; k < Kd; k += 32) s = fmaf(a[k], w[(int64_t)n * Kd + k], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      if (bias) s += bias[(int64_t)member * N + n];
      if (relu) s = fmaxf(s, 0.f);
      C[(int64_t)z * c_member_stride + row * N + n] = s;
    }
  }
}

// Delta-UQ input assembly: in[z][m] = cat(x[m] - a_k, a_k)   (oracle/uq_oracle.py restatement of
// deltaUQ_MLP.create_anchored_batch; parity unpinned).
// PAGER (swap = 1): the anchor is the input and the sample the anchor, in = cat(a_k - x[m], x[m])
// (PAGERMLP._anchored_predictions, models.py:396-424).
// anchor_first: the halves in the other order, cat(a_k, x[m] - a_k) / cat(x[m], a_k - x[m])
// (UQ_MODEL_ANCHOR_FIRST: the public deltauq package's channel order).
__global__ void delta_input_kernel(const float* __restrict__ x, const float* __restrict__ anchors,
                                   float* __restrict__ out, int M, int d, int member0, int swap,
                                   int anchor_first) {
  const int z = blockIdx.z;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * d) return;
  const int64_t m = i / d;
  const int c = (int)(i - m * d);
  const float a = anchors[(int64_t)(member0 + z) * d + c];
  float* o = out + ((int64_t)z * M + m) * (2 * d);
  const int diff_at = anchor_first ? d : 0, base_at = anchor_first ? 0 : d;
  o[diff_at + c] = swap ? a - x[i] : x[i] - a;
  o[base_at + c] = swap ? x[i] : a;
}

// PAGER: out0 = mean over anchors of the predictions, out1 = max_k |P[k] - Y_k| (float32 like the
// reference's torch.abs / torch.max), raised to floor[i] when given (torch.maximum, :389-390).
__global__ void finalize_pager_kernel(const float* __restrict__ slab, int64_t member_stride, int K,
                                      int64_t len, int d_out, const float* __restrict__ targets,
                                      const float* __restrict__ floor_, float* __restrict__ out0,
                                      float* __restrict__ out1) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  const int o = (int)(i % d_out);
  double s = 0.0;
  float mx = 0.f;
  for (int k = 0; k < K; ++k) {
    const float y = slab[(int64_t)k * member_stride + i];
    s += (double)y;
    mx = fmaxf(mx, fabsf(y - targets[(int64_t)k * d_out + o]));
  }
  out0[i] = (float)(s / (double)K);
  out1[i] = floor_ ? fmaxf(mx, floor_[i]) : mx;
}

// (mean, std) or (mean, M2) over the member axis of slab[K][M*out]; float64 accumulation.
__global__ void finalize_kernel(const float* __restrict__ slab, int64_t member_stride, int K,
                                int64_t len, float* __restrict__ out0, float* __restrict__ out1,
                                int moments) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  double s = 0.0;
  for (int k = 0; k < K; ++k) s += (double)slab[(int64_t)k * member_stride + i];
  const double mean = s / (double)K;
  double m2 = 0.0;
  for (int k = 0; k < K; ++k) {
    const double d = (double)slab[(int64_t)k * member_stride + i] - mean;
    m2 += d * d;
  }
  out0[i] = (float)mean;
  // K == 1: unbiased std is 0/0 = NaN, exactly what torch.std returns (models.py:106).
  out1[i] = moments ? (float)m2 : (float)sqrt(m2 / (double)(K - 1));
}

struct Plan {
  int64_t chunk;     // samples per chunk
  int group;         // members per grouped launch
  size_t act_bytes;  // one activation buffer [group][chunk][max_width]
  size_t slab_bytes; // [member_count][chunk][d_out]
  size_t total;
};

Plan make_plan(const uq_model* m, int64_t n, const uq_forward_args* a) {
  Plan p;
  const int K = a->member_count;
  int width = m->max_width;
  if (m->d_in > width) width = m->d_in;  // Delta-UQ stages cat(x - a, a) in an activation buffer
  p.group = K < 32 ? K : 32;
  const size_t budget = (size_t)96 << 20;  // two activation buffers ~ L2-resident
  int64_t c = (int64_t)(budget / ((size_t)p.group * width * sizeof(float)));
  int64_t chunk = 128;
  while (chunk * 2 <= c && chunk < 65536) chunk *= 2;
  if (chunk > n) chunk = ((n + 127) / 128) * 128;
  if (chunk < 128) chunk = 128;
  p.chunk = chunk;
  p.act_bytes = (size_t)p.group * chunk * width * sizeof(float);
  p.act_bytes = (p.act_bytes + 255) & ~(size_t)255;
  p.slab_bytes = ((size_t)K * chunk * m->d_out * sizeof(float) + 255) & ~(size_t)255;
  p.total = 2 * p.act_bytes + p.slab_bytes;
  return p;
}

}  // namespace

size_t fp32_workspace_bytes(const uq_model* m, int64_t n, const uq_forward_args* a) {
  return make_plan(m, n, a).total;
}

int fp32_forward(const uq_model* m, const float* x, int64_t n, const uq_forward_args* a,
                 float* out0, float* out1, void* ws, size_t ws_bytes, cudaStream_t st) {
  const Plan p = make_plan(m, n, a);
  UQ_REQUIRE(ws != nullptr && ws_bytes >= p.total, UQ_ERR_WORKSPACE,
             "fp32 forward needs %zu workspace bytes, got %zu", p.total, ws_bytes);
  char* base = static_cast<char*>(ws);
  float* act[2] = {reinterpret_cast<float*>(base), reinterpret_cast<float*>(base + p.act_bytes)};
  float* slab = reinterpret_cast<float*>(base + 2 * p.act_bytes);

  const int K = a->member_count;
  const bool shared_weights = (a->mode != UQ_MODE_ENSEMBLE);
  const bool mc = (a->mode == UQ_MODE_MC_DROPOUT) && a->dropout_active;
  const float keep_f = (float)(1.0 - a->dropout_p);
  const float drop_scale = 1.0f / keep_f;
  PhiloxKey key{(uint32_t)(a->philox_seed & 0xffffffffu), (uint32_t)(a->philox_seed >> 32),
                (uint32_t)(a->philox_offset & 0xffffffffu)};
  const uint32_t thr16 = dropout_thr16((float)a->dropout_p);
  const int L = m->n_layers;

  for (int64_t s0 = 0; s0 < n; s0 += p.chunk) {
    const int M = (int)((n - s0) < p.chunk ? (n - s0) : p.chunk);
    const int64_t slab_stride = (int64_t)M * m->d_out;
    for (int g0 = 0; g0 < K; g0 += p.group) {
      const int G = (K - g0) < p.group ? (K - g0) : p.group;
      const int gm0 = a->member_begin + g0;  // global id of the first member of this launch
      const float* in = x + s0 * m->d_in;
      int64_t in_stride = 0;
      int ld_in = m->d_in;
      int cur = 0;
      if (a->mode == UQ_MODE_DELTA_UQ || a->mode == UQ_MODE_PAGER) {
        const int d = m->d_in / 2;
        in = x + s0 * d;
        const int64_t tot = (int64_t)M * d;
        dim3 grid((unsigned)((tot + 255) / 256), 1, G);
        delta_input_kernel<<<grid, 256, 0, st>>>(in, a->anchors, act[0], M, d, gm0,
                                                 a->mode == UQ_MODE_PAGER ? 1 : 0,
                                                 m->anchor_first ? 1 : 0);
        UQ_LAUNCH_CHECK();
        in = act[0];
        in_stride = (int64_t)M * m->d_in;
        cur = 1;
      }
      int drop_ord = 0;
      const uint8_t* mask_base = a->masks;
      for (int l = 0; l < L; ++l) {
        const Layer& ly = m->layers[l];
        const bool last = (l == L - 1);
        const int64_t w_stride = shared_weights ? 0 : (int64_t)ly.out * ly.in;
        const int wm0 = shared_weights ? 0 : gm0;
        const int wstep = shared_weights ? 0 : 1;
        float* outp = last ? slab + (int64_t)g0 * slab_stride : act[cur];
        const int64_t out_stride = last ? slab_stride : (int64_t)M * ly.out;
        const bool simple = last && !ly.has_bn && !(ly.dropout && mc) && ly.out <= 8;
        if (simple) {
          dim3 grid((unsigned)((M + 7) / 8), 1, G);
          linear_small_out_kernel<<<grid, 256, 0, st>>>(
              in, in_stride, ld_in, ly.w, w_stride, ly.bias, outp, out_stride, M, ly.out, ly.in,
              wm0, wstep, ly.relu ? 1 : 0);
          UQ_LAUNCH_CHECK();
        } else {
          Epilogue ep;
          ep.bias = ly.bias;
          ep.alpha = ly.has_bn ? ly.alpha : nullptr;
          ep.beta = ly.has_bn ? ly.beta : nullptr;
          ep.relu = ly.relu ? 1 : 0;
          ep.drop_mode = 0;
          ep.mask = nullptr;
          ep.n_total = n;
          ep.sample0 = s0;
          ep.philox_row0 = a->row_base;
          ep.pass0 = gm0;
          ep.drop_layer = drop_ord;
          ep.drop_scale = drop_scale;
          ep.thr16 = thr16;
          ep.key = key;
          if (ly.dropout && mc) {
            if (a->masks) {
              ep.drop_mode = 1;
              ep.mask = mask_base;
            } else {
              ep.drop_mode = 2;
            }
          }
          dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((ly.out + BN - 1) / BN), G);
          const bool vec = (ly.in % 4 == 0) && (ld_in % 4 == 0) &&
                           ((reinterpret_cast<uintptr_t>(in) & 15) == 0) &&
                           ((reinterpret_cast<uintptr_t>(ly.w) & 15) == 0);
          // bias/alpha/beta are indexed by the *weight* member (slot 0 when weights are shared);
          // masks / philox use the global pass id ep.pass0 + blockIdx.z.
          if (vec)
            sgemm_tn_kernel<true><<<grid, GEMM_THREADS, 0, st>>>(
                in, in_stride, ld_in, ly.w, w_stride, outp, out_stride, ly.out, M, ly.out,
                ly.in, wm0, wstep, ep);
          else
            sgemm_tn_kernel<false><<<grid, GEMM_THREADS, 0, st>>>(
                in, in_stride, ld_in, ly.w, w_stride, outp, out_stride, ly.out, M, ly.out,
                ly.in, wm0, wstep, ep);
          UQ_LAUNCH_CHECK();
        }
        if (ly.dropout) {
          if (a->masks) mask_base += (size_t)a->total_members * (size_t)n * (size_t)ly.out;
          ++drop_ord;
        }
        in = outp;
        in_stride = out_stride;
        ld_in = ly.out;
        cur ^= 1;
      }
    }
    {
      const int64_t len = (int64_t)M * m->d_out;
      if (a->mode == UQ_MODE_PAGER) {
        finalize_pager_kernel<<<(unsigned)((len + 255) / 256), 256, 0, st>>>(
            slab, len, K, len, m->d_out, a->anchor_targets + (size_t)a->member_begin * m->d_out,
            a->score_floor ? a->score_floor + s0 * m->d_out : nullptr, out0 + s0 * m->d_out,
            out1 + s0 * m->d_out);
        UQ_LAUNCH_CHECK();
      } else
      finalize_kernel<<<(unsigned)((len + 255) / 256), 256, 0, st>>>(
          slab, len, K, len, out0 + s0 * m->d_out, out1 + s0 * m->d_out,
          a->output == UQ_OUT_MOMENTS ? 1 : 0);
      UQ_LAUNCH_CHECK();
    }
  }
  return UQ_OK;
}

}  // namespace uq
