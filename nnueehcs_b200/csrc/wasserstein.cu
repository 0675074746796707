// 1-D Wasserstein distance between two unweighted samples:  W1 = integral |F_u - F_v| dx.
//
// Replaces scipy.stats.wasserstein_distance as called by
// WassersteinEvaluation._evaluate_uncertainties (nnueehcs/evaluation.py:175-188); the arithmetic
// follows scipy's _cdf_distance(p=1): the float32 scores are widened to float64 first
// (_validate_distribution), then successive differences of the merged sorted values times
// |cdf_u - cdf_v| with cdf = (#values <= x)/size, all in float64.
//
// GPU shape: two radix sorts (sort.cu), then one merge-path kernel -- each block binary-searches
// its diagonal of the (u, v) merge grid, stages its two input runs in shared memory with
// coalesced loads, each thread merges 8 consecutive positions sequentially, and the partial sums
// go through warp shuffles to one float64 per block; a last single-block kernel adds the block
// partials in a fixed order (deterministic result).
#include "common.cuh"
#include "sort.cuh"

namespace uq {
namespace {

constexpr int MRG_THREADS = 256;
constexpr int MRG_ITEMS = 8;
constexpr int MRG_TILE = MRG_THREADS * MRG_ITEMS;  // merged positions per block

// number of u's among the first k merged elements (u before v on ties)
__device__ __forceinline__ int64_t merge_path(const float* __restrict__ U, int64_t nu,
                                              const float* __restrict__ V, int64_t nv, int64_t k) {
  int64_t lo = k > nv ? k - nv : 0;
  int64_t hi = k < nu ? k : nu;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (U[mid] <= V[k - mid - 1]) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ int merge_path_smem(const float* U, int nu, const float* V, int nv, int k) {
  int lo = k > nv ? k - nv : 0;
  int hi = k < nu ? k : nu;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (U[mid] <= V[k - mid - 1]) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// merge-path split of every block boundary k = b * MRG_TILE, one thread each: done up front so the
// ~25 dependent global loads of a binary search are paid once in parallel instead of serially by
// thread 0 of each of the ~50 k integral blocks (that cost 0.6 of the kernel's 0.97 ms)
__global__ void __launch_bounds__(256)
merge_partition_kernel(const float* __restrict__ U, int64_t nu, const float* __restrict__ V,
                       int64_t nv, int64_t blocks, int64_t* __restrict__ splits) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= blocks) return;
  splits[b] = merge_path(U, nu, V, nv, b * MRG_TILE);
}

__global__ void __launch_bounds__(MRG_THREADS)
cdf_integral_kernel(const float* __restrict__ U, int64_t nu, const float* __restrict__ V,
                    int64_t nv, int64_t u_below, int64_t v_below, int64_t nu_total,
                    int64_t nv_total, const int64_t* __restrict__ splits,
                    double* __restrict__ block_partials) {
  __shared__ float su[MRG_TILE + 1];
  __shared__ float sv[MRG_TILE + 1];
  __shared__ double warp_part[MRG_THREADS / 32];
  const int64_t total = nu + nv;
  const int64_t k0 = (int64_t)blockIdx.x * MRG_TILE;           // first merged index of the block
  int64_t k1 = k0 + MRG_TILE;                                   // contributions k in [k0, k1)
  if (k1 > total - 1) k1 = total - 1;
  const int t = threadIdx.x;
  const int64_t i0 = splits[blockIdx.x], j0 = k0 - i0;
  const int len = (int)(k1 - k0);
  const int lu = (int)((nu - i0) < (int64_t)(len + 1) ? (nu - i0) : (int64_t)(len + 1));
  const int lv = (int)((nv - j0) < (int64_t)(len + 1) ? (nv - j0) : (int64_t)(len + 1));
  for (int i = t; i < lu; i += MRG_THREADS) su[i] = U[i0 + i];
  for (int i = t; i < lv; i += MRG_THREADS) sv[i] = V[j0 + i];
  __syncthreads();

  double acc = 0.0;
  const int ka = t * MRG_ITEMS;
  if (ka < len) {
    const int kb = (ka + MRG_ITEMS) < len ? (ka + MRG_ITEMS) : len;
    int i = merge_path_smem(su, lu, sv, lv, ka);
    int j = ka - i;
    // consume merged element ka
    float cur;
    if (j >= lv || (i < lu && su[i] <= sv[j])) cur = su[i++]; else cur = sv[j++];
    for (int k = ka; k < kb; ++k) {
      // next merged element (exists: k <= total - 2)
      float nxt;
      const bool take_u = (j >= lv) || (i < lu && su[i] <= sv[j]);
      nxt = take_u ? su[i] : sv[j];
      const double delta = (double)nxt - (double)cur;  // exact: scipy widens to float64 first
      // numpy: idx / size in float64
      const double cu = (double)(u_below + i0 + i) / (double)nu_total;
      const double cv = (double)(v_below + j0 + j) / (double)nv_total;
      acc += fabs(cu - cv) * delta;
      if (take_u) ++i; else ++j;
      cur = nxt;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((t & 31) == 0) warp_part[t >> 5] = acc;
  __syncthreads();
  if (t == 0) {
    double s = 0.0;
    for (int w = 0; w < MRG_THREADS / 32; ++w) s += warp_part[w];
    block_partials[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(1024)
sum_partials_kernel(const double* __restrict__ parts, int64_t n, double* __restrict__ out) {
  __shared__ double sh[1024];
  const int t = threadIdx.x;
  double s = 0.0;
  for (int64_t i = t; i < n; i += 1024) s += parts[i];
  sh[t] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (t < o) sh[t] += sh[t + o];
    __syncthreads();
  }
  if (t == 0) *out = sh[0];
}

struct WsLayout {
  size_t u, ut, v, vt, scratch, parts, splits, result, total;
  int64_t blocks;
};

WsLayout layout(int64_t nu, int64_t nv) {
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  WsLayout L;
  size_t o = 0;
  L.u = o; o += al(sizeof(float) * (size_t)nu);
  L.ut = o; o += al(sizeof(float) * (size_t)nu);
  L.v = o; o += al(sizeof(float) * (size_t)nv);
  L.vt = o; o += al(sizeof(float) * (size_t)nv);
  const size_t su = radix_sort_scratch_bytes(nu), sv = radix_sort_scratch_bytes(nv);
  L.scratch = o; o += al(su > sv ? su : sv);
  L.blocks = (nu + nv - 1 + MRG_TILE - 1) / MRG_TILE;
  if (L.blocks < 1) L.blocks = 1;
  L.parts = o; o += al(sizeof(double) * (size_t)L.blocks);
  L.splits = o; o += al(sizeof(int64_t) * (size_t)L.blocks);
  L.result = o; o += 256;
  L.total = o;
  return L;
}

}  // namespace

size_t wasserstein_workspace_bytes(int64_t nu, int64_t nv) {
  if (nu < 1 || nv < 1) return 0;
  return layout(nu, nv).total;
}

int wasserstein_1d(const float* u, int64_t nu, const float* v, int64_t nv, double* out_host,
                   void* ws, size_t ws_bytes, cudaStream_t st) {
  const WsLayout L = layout(nu, nv);
  UQ_REQUIRE(ws != nullptr && ws_bytes >= L.total, UQ_ERR_WORKSPACE,
             "wasserstein needs %zu workspace bytes, got %zu", L.total, ws_bytes);
  UQ_REQUIRE(nu + nv < ((int64_t)1 << 31), UQ_ERR_INVALID, "wasserstein: too many values");
  char* b = static_cast<char*>(ws);
  float* du = reinterpret_cast<float*>(b + L.u);
  float* dut = reinterpret_cast<float*>(b + L.ut);
  float* dv = reinterpret_cast<float*>(b + L.v);
  float* dvt = reinterpret_cast<float*>(b + L.vt);
  double* parts = reinterpret_cast<double*>(b + L.parts);
  double* result = reinterpret_cast<double*>(b + L.result);
  UQ_CUDA(cudaMemcpyAsync(du, u, sizeof(float) * (size_t)nu, cudaMemcpyDeviceToDevice, st));
  UQ_CUDA(cudaMemcpyAsync(dv, v, sizeof(float) * (size_t)nv, cudaMemcpyDeviceToDevice, st));
  float *su = nullptr, *sv = nullptr;
  int rc = radix_sort_f32(du, dut, nu, b + L.scratch, radix_sort_scratch_bytes(nu), &su, st);
  if (rc != UQ_OK) return rc;
  rc = radix_sort_f32(dv, dvt, nv, b + L.scratch, radix_sort_scratch_bytes(nv), &sv, st);
  if (rc != UQ_OK) return rc;
  if (nu + nv - 1 > 0) {
    int64_t* splits = reinterpret_cast<int64_t*>(b + L.splits);
    merge_partition_kernel<<<(unsigned)((L.blocks + 255) / 256), 256, 0, st>>>(su, nu, sv, nv,
                                                                               L.blocks, splits);
    UQ_LAUNCH_CHECK();
    cdf_integral_kernel<<<(unsigned)L.blocks, MRG_THREADS, 0, st>>>(su, nu, sv, nv, 0, 0, nu, nv,
                                                                    splits, parts);
    UQ_LAUNCH_CHECK();
    sum_partials_kernel<<<1, 1024, 0, st>>>(parts, L.blocks, result);
    UQ_LAUNCH_CHECK();
  } else {
    UQ_CUDA(cudaMemsetAsync(result, 0, sizeof(double), st));
  }
  UQ_CUDA(cudaMemcpyAsync(out_host, result, sizeof(double), cudaMemcpyDeviceToHost, st));
  UQ_CUDA(cudaStreamSynchronize(st));
  return UQ_OK;
}

// One value range of a sample-sorted (multi-GPU) Wasserstein: this rank holds every u and v
// value of its range; `u_below` / `v_below` values of each sample lie in lower ranges.  Writes
// {partial integral over the local merged sequence, first merged value, last merged value}.
int wasserstein_1d_range(const float* u, int64_t nu, const float* v, int64_t nv, int64_t u_below,
                         int64_t v_below, int64_t nu_total, int64_t nv_total, double* out_host,
                         void* ws, size_t ws_bytes, cudaStream_t st) {
  const WsLayout L = layout(nu > 0 ? nu : 1, nv > 0 ? nv : 1);
  UQ_REQUIRE(ws != nullptr && ws_bytes >= L.total, UQ_ERR_WORKSPACE,
             "wasserstein range needs %zu workspace bytes, got %zu", L.total, ws_bytes);
  UQ_REQUIRE(nu + nv >= 1 && nu + nv < ((int64_t)1 << 31), UQ_ERR_INVALID,
             "wasserstein range: %lld values", (long long)(nu + nv));
  char* b = static_cast<char*>(ws);
  float* du = reinterpret_cast<float*>(b + L.u);
  float* dut = reinterpret_cast<float*>(b + L.ut);
  float* dv = reinterpret_cast<float*>(b + L.v);
  float* dvt = reinterpret_cast<float*>(b + L.vt);
  double* parts = reinterpret_cast<double*>(b + L.parts);
  double* result = reinterpret_cast<double*>(b + L.result);
  float *su = du, *sv = dv;
  int rc;
  if (nu > 0) {
    UQ_CUDA(cudaMemcpyAsync(du, u, sizeof(float) * (size_t)nu, cudaMemcpyDeviceToDevice, st));
    rc = radix_sort_f32(du, dut, nu, b + L.scratch, radix_sort_scratch_bytes(nu), &su, st);
    if (rc != UQ_OK) return rc;
  }
  if (nv > 0) {
    UQ_CUDA(cudaMemcpyAsync(dv, v, sizeof(float) * (size_t)nv, cudaMemcpyDeviceToDevice, st));
    rc = radix_sort_f32(dv, dvt, nv, b + L.scratch, radix_sort_scratch_bytes(nv), &sv, st);
    if (rc != UQ_OK) return rc;
  }
  if (nu + nv - 1 > 0) {
    const int64_t blocks = (nu + nv - 1 + MRG_TILE - 1) / MRG_TILE;
    int64_t* splits = reinterpret_cast<int64_t*>(b + L.splits);
    merge_partition_kernel<<<(unsigned)((blocks + 255) / 256), 256, 0, st>>>(su, nu, sv, nv, blocks,
                                                                             splits);
    UQ_LAUNCH_CHECK();
    cdf_integral_kernel<<<(unsigned)blocks, MRG_THREADS, 0, st>>>(su, nu, sv, nv, u_below, v_below,
                                                                  nu_total, nv_total, splits, parts);
    UQ_LAUNCH_CHECK();
    sum_partials_kernel<<<1, 1024, 0, st>>>(parts, blocks, result);
    UQ_LAUNCH_CHECK();
  } else {
    UQ_CUDA(cudaMemsetAsync(result, 0, sizeof(double), st));
  }
  float ends[4] = {0.f, 0.f, 0.f, 0.f};  // u first, u last, v first, v last
  UQ_CUDA(cudaMemcpyAsync(out_host, result, sizeof(double), cudaMemcpyDeviceToHost, st));
  if (nu > 0) {
    UQ_CUDA(cudaMemcpyAsync(&ends[0], su, sizeof(float), cudaMemcpyDeviceToHost, st));
    UQ_CUDA(cudaMemcpyAsync(&ends[1], su + nu - 1, sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  if (nv > 0) {
    UQ_CUDA(cudaMemcpyAsync(&ends[2], sv, sizeof(float), cudaMemcpyDeviceToHost, st));
    UQ_CUDA(cudaMemcpyAsync(&ends[3], sv + nv - 1, sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  UQ_CUDA(cudaStreamSynchronize(st));
  double first, last;
  if (nu > 0 && nv > 0) {
    first = ends[0] < ends[2] ? ends[0] : ends[2];
    last = ends[1] > ends[3] ? ends[1] : ends[3];
  } else if (nu > 0) {
    first = ends[0], last = ends[1];
  } else {
    first = ends[2], last = ends[3];
  }
  out_host[1] = first;
  out_host[2] = last;
  return UQ_OK;
}

}  // namespace uq
