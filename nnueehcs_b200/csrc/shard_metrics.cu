// Building blocks of the multi-GPU (sample-sorted) distribution metrics, SURVEY.md section 8e:
//   * uq_sample_stats      -- (min, max, mean, M2) of one rank's shard, merged on the host with
//                             Chan's formula to get the global Scott bandwidth and grid range;
//   * uq_key_histogram     -- counts per coarse order-preserving key bin (top 14 bits of the
//                             radix key: sign, exponent, 5 mantissa bits), all-reduced to pick
//                             balanced value-range splitters;
//   * uq_partition_by_bin  -- scatters a shard into per-destination-rank segments (the send
//                             buffer of the one all-to-all); order inside a segment is arbitrary
//                             because the receiver sorts.
// The reference has no counterpart: it is single-process scipy (nnueehcs/evaluation.py:182,268).
#include "common.cuh"

namespace uq {
namespace {

constexpr int KEY_BINS = 16384;
constexpr int STAT_BLOCKS = 592;

__device__ __forceinline__ uint32_t key_bin(float x) {
  const uint32_t b = __float_as_uint(x);
  const uint32_t k = b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);  // same map as the radix sort
  return k >> 18;
}

// per-block (min, max, sum(x - x0), sum((x - x0)^2)) in float64, x0 = x[0].  float4 loads, four
// in flight per thread (the scalar-load version ran at 2.1 TB/s: 4 KB in flight per SM); min / max
// stay in float32 (exact), only the two sums are float64.
__global__ void __launch_bounds__(256)
shard_stats_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ partials) {
  __shared__ double sh[4][8];
  const float x0f = x[0];
  const double x0 = (double)x0f;
  double s1 = 0.0, s2 = 0.0;
  float mnf = x0f, mxf = x0f;
  auto add = [&](float v) {
    const double d = (double)v - x0;
    s1 += d;
    s2 = fma(d, d, s2);
    mnf = fminf(mnf, v);
    mxf = fmaxf(mxf, v);
  };
  const int64_t head = min(n, (int64_t)((16 - ((uintptr_t)x & 15)) & 15) / 4);
  const int64_t n4 = (n - head) / 4;
  const float4* x4 = reinterpret_cast<const float4*>(x + head);
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  if (gtid < head) add(__ldg(x + gtid));
  int64_t i = gtid;
  for (; i + 3 * gstride < n4; i += 4 * gstride) {
    const float4 a0 = __ldg(x4 + i), a1 = __ldg(x4 + i + gstride);
    const float4 a2 = __ldg(x4 + i + 2 * gstride), a3 = __ldg(x4 + i + 3 * gstride);
    add(a0.x); add(a0.y); add(a0.z); add(a0.w);
    add(a1.x); add(a1.y); add(a1.z); add(a1.w);
    add(a2.x); add(a2.y); add(a2.z); add(a2.w);
    add(a3.x); add(a3.y); add(a3.z); add(a3.w);
  }
  for (; i < n4; i += gstride) {
    const float4 a0 = __ldg(x4 + i);
    add(a0.x); add(a0.y); add(a0.z); add(a0.w);
  }
  const int64_t tail0 = head + 4 * n4;
  if (tail0 + gtid < n) add(__ldg(x + tail0 + gtid));
  double mn = (double)mnf, mx = (double)mxf;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_down_sync(0xffffffffu, s1, o);
    s2 += __shfl_down_sync(0xffffffffu, s2, o);
    mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) sh[0][w] = s1, sh[1][w] = s2, sh[2][w] = mn, sh[3][w] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) {
      sh[0][0] += sh[0][k];
      sh[1][0] += sh[1][k];
      sh[2][0] = fmin(sh[2][0], sh[2][k]);
      sh[3][0] = fmax(sh[3][0], sh[3][k]);
    }
    for (int k = 0; k < 4; ++k) partials[4 * blockIdx.x + k] = sh[k][0];
  }
}

// block-private histogram in shared memory (64 KB), flushed with one atomic per non-empty bin
__global__ void __launch_bounds__(512)
key_histogram_kernel(const float* __restrict__ x, int64_t n, uint32_t* __restrict__ hist) {
  extern __shared__ uint32_t sh_hist[];
  for (int i = threadIdx.x; i < KEY_BINS; i += blockDim.x) sh_hist[i] = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    atomicAdd(&sh_hist[key_bin(__ldg(x + i))], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < KEY_BINS; i += blockDim.x) {
    const uint32_t c = sh_hist[i];
    if (c) atomicAdd(&hist[i], c);
  }
}

constexpr int PART_THREADS = 256;
constexpr int PART_ITEMS = 16;
constexpr int PART_MAX = 64;

// Each block counts its tile per destination, reserves one contiguous run per destination with
// a single atomic on the destination cursor, then writes its values into the runs.
__global__ void __launch_bounds__(PART_THREADS)
partition_kernel(const float* __restrict__ x, int64_t n, const uint8_t* __restrict__ bin_to_part,
                 int n_parts, float* __restrict__ out, unsigned long long* __restrict__ cursors) {
  __shared__ uint32_t cnt[PART_MAX];
  __shared__ unsigned long long base[PART_MAX];
  const int64_t tile0 = (int64_t)blockIdx.x * PART_THREADS * PART_ITEMS;
  if (threadIdx.x < PART_MAX) cnt[threadIdx.x] = 0;
  __syncthreads();
  float v[PART_ITEMS];
  uint32_t dst[PART_ITEMS], slot[PART_ITEMS];
#pragma unroll
  for (int r = 0; r < PART_ITEMS; ++r) {
    const int64_t i = tile0 + (int64_t)r * PART_THREADS + threadIdx.x;
    dst[r] = 0xFFFFFFFFu;
    if (i < n) {
      v[r] = __ldg(x + i);
      dst[r] = __ldg(bin_to_part + key_bin(v[r]));
      slot[r] = atomicAdd(&cnt[dst[r]], 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x < n_parts && cnt[threadIdx.x])
    base[threadIdx.x] = atomicAdd(&cursors[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < PART_ITEMS; ++r)
    if (dst[r] != 0xFFFFFFFFu) out[base[dst[r]] + slot[r]] = v[r];
}

}  // namespace

// (min, max, mean, M2) of one or two device samples with ONE stream synchronisation.
// out_host: 4 doubles per sample; workspace: uq_sample_stats_workspace_bytes() per sample.
int sample_stats_multi(const float* const* xs, const int64_t* ns, int count, double* out_host,
                       void* workspace, cudaStream_t st) {
  static thread_local double h[2][4 * STAT_BLOCKS];
  float x0f[2] = {0.f, 0.f};
  double* partials = static_cast<double*>(workspace);
  for (int s = 0; s < count; ++s) {
    shard_stats_kernel<<<STAT_BLOCKS, 256, 0, st>>>(xs[s], ns[s], partials + (size_t)s * 4 * STAT_BLOCKS);
    UQ_LAUNCH_CHECK();
  }
  for (int s = 0; s < count; ++s) {
    UQ_CUDA(cudaMemcpyAsync(h[s], partials + (size_t)s * 4 * STAT_BLOCKS, sizeof(h[s]),
                            cudaMemcpyDeviceToHost, st));
    UQ_CUDA(cudaMemcpyAsync(&x0f[s], xs[s], sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  UQ_CUDA(cudaStreamSynchronize(st));
  for (int s = 0; s < count; ++s) {
    double s1 = 0.0, s2 = 0.0, mn = h[s][2], mx = h[s][3];
    for (int b = 0; b < STAT_BLOCKS; ++b) {
      s1 += h[s][4 * b], s2 += h[s][4 * b + 1];
      if (h[s][4 * b + 2] < mn) mn = h[s][4 * b + 2];
      if (h[s][4 * b + 3] > mx) mx = h[s][4 * b + 3];
    }
    const double dn = (double)ns[s];
    out_host[4 * s + 0] = mn;
    out_host[4 * s + 1] = mx;
    out_host[4 * s + 2] = (double)x0f[s] + s1 / dn;   // mean
    out_host[4 * s + 3] = s2 - s1 * s1 / dn;          // M2 = sum (x - mean)^2
  }
  return UQ_OK;
}

}  // namespace uq

using namespace uq;

extern "C" {

int32_t uq_key_bins(void) { return KEY_BINS; }

size_t uq_sample_stats_workspace_bytes(void) { return sizeof(double) * 4 * STAT_BLOCKS; }

int uq_sample_stats(const float* x, int64_t n, double* out_host, void* workspace,
                    size_t workspace_bytes, void* stream) {
  UQ_REQUIRE(x && out_host && n >= 1, UQ_ERR_INVALID, "uq_sample_stats: NULL argument or n < 1");
  UQ_REQUIRE(workspace && workspace_bytes >= uq_sample_stats_workspace_bytes(), UQ_ERR_WORKSPACE,
             "uq_sample_stats: workspace too small");
  return sample_stats_multi(&x, &n, 1, out_host, workspace, static_cast<cudaStream_t>(stream));
}

int uq_key_histogram(const float* x, int64_t n, uint32_t* hist, void* stream) {
  UQ_REQUIRE(x && hist && n >= 1, UQ_ERR_INVALID, "uq_key_histogram: NULL argument or n < 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static PerDeviceOnce opted;
  if (int rc = smem_opt_in(key_histogram_kernel, KEY_BINS * (int)sizeof(uint32_t), opted)) return rc;
  int64_t blocks = (n + 512 * 32 - 1) / (512 * 32);
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (blocks < 1) blocks = 1;
  key_histogram_kernel<<<(unsigned)blocks, 512, KEY_BINS * sizeof(uint32_t), st>>>(x, n, hist);
  UQ_LAUNCH_CHECK();
  return UQ_OK;
}

int uq_partition_by_bin(const float* x, int64_t n, const uint8_t* bin_to_part, int32_t n_parts,
                        float* out, unsigned long long* cursors, void* stream) {
  UQ_REQUIRE(x && bin_to_part && out && cursors && n >= 1, UQ_ERR_INVALID,
             "uq_partition_by_bin: NULL argument or n < 1");
  UQ_REQUIRE(n_parts >= 1 && n_parts <= PART_MAX, UQ_ERR_INVALID,
             "uq_partition_by_bin: n_parts %d outside [1, %d]", n_parts, PART_MAX);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t tile = (int64_t)PART_THREADS * PART_ITEMS;
  partition_kernel<<<(unsigned)((n + tile - 1) / tile), PART_THREADS, 0, st>>>(
      x, n, bin_to_part, n_parts, out, cursors);
  UQ_LAUNCH_CHECK();
  return UQ_OK;
}

}  // extern "C"
