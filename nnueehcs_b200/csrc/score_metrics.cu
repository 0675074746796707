// Score consumers on the device: everything the reference computes
// from the two uncertainty-score vectors besides the distribution distances --
//   MeanScoreEvaluation / MaxScoreEvaluation / PercentileScoreEvaluation   evaluation.py:292-381
//   TNRatTPX (a Python loop over every unique score, two .item() syncs each) evaluation.py:538-580
//   AUROC (sklearn roc_auc_score on the host)                               evaluation.py:614-624
//   PercentileBasedClassifier (torch.quantile + four counts)  evaluation.py:637-662,
//                                                             classification.py:103-143
// -- from ONE pair of radix sorts: closed forms on the sorted arrays (oracle/metrics_oracle.py
// states and pins them).  The scores never leave the GPU; one small struct comes back.
#include <math.h>

#include "common.cuh"
#include "sort.cuh"

namespace uq {
namespace {

__device__ __forceinline__ int64_t lower_bound_f(const float* __restrict__ a, int64_t n, float x) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int64_t upper_bound_f(const float* __restrict__ a, int64_t n, float x) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] <= x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// sum of the ID scores in float64 (deterministic: fixed grid, block partials added in order later)
__global__ void __launch_bounds__(256)
sum_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ partials) {
  __shared__ double sh[8];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    s += (double)__ldg(x + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    partials[blockIdx.x] = t;
  }
}

// Mann-Whitney: sum over OOD scores v of (#ID < v) + (#ID <= v)  (= 2U, exact in uint64).
// Both arrays are sorted, so the counts are read off a MERGE instead of two 26-step binary
// searches per OOD score (2.1 ms at 50 M + 50 M): with ties broken "ID first" the number of ID
// scores merged before an OOD score v is #ID <= v, with ties broken "OOD first" it is #ID < v --
// one launch per rule.  Persistent blocks own contiguous ranges of 2048-position tiles of the
// merged sequence; the first split of a range comes from one warp-cooperative 33-ary search, every
// later one from where the previous tile ended; a thread finds its 8-position diagonal in shared
// memory and merges sequentially (the structure of wasserstein.cu's cdf_integral_kernel).
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;

template <bool ID_FIRST>
__device__ __forceinline__ bool takes_id(float a, float b) { return ID_FIRST ? a <= b : a < b; }

// number of ID scores among the first k merged elements
template <bool ID_FIRST>
__device__ __forceinline__ int64_t split_warp(const float* __restrict__ A, int64_t na,
                                              const float* __restrict__ B, int64_t nb, int64_t k) {
  const int lane = threadIdx.x & 31;
  int64_t lo = k > nb ? k - nb : 0;
  int64_t hi = k < na ? k : na;
  while (lo < hi) {   // takes_id(A[mid], B[k - mid - 1]) holds exactly for mid < answer
    const int64_t span = hi - lo;
    const int64_t mid = lo + (span * (lane + 1)) / 33;
    const bool pred = takes_id<ID_FIRST>(A[mid], B[k - mid - 1]);
    const int c = __popc(__ballot_sync(0xffffffffu, pred));
    const int64_t last_true = __shfl_sync(0xffffffffu, mid, c > 0 ? c - 1 : 0);
    const int64_t first_false = __shfl_sync(0xffffffffu, mid, c < 32 ? c : 31);
    if (c > 0) lo = last_true + 1;
    if (c < 32) hi = first_false;
  }
  return lo < hi ? lo : hi;
}
template <bool ID_FIRST>
__device__ __forceinline__ int split_smem(const float* A, int na, const float* B, int nb, int k) {
  int lo = k > nb ? k - nb : 0;
  int hi = k < na ? k : na;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (takes_id<ID_FIRST>(A[mid], B[k - mid - 1])) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <bool ID_FIRST>
__global__ void __launch_bounds__(RS_THREADS, 4)
rank_sum_kernel(const float* __restrict__ A /*sorted ID*/, int64_t na,
                const float* __restrict__ B /*sorted OOD*/, int64_t nb, int64_t tiles,
                unsigned long long* __restrict__ acc) {
  __shared__ float sa[RS_TILE + 1];
  __shared__ float sb[RS_TILE + 1];
  __shared__ unsigned long long warp_part[RS_THREADS / 32];
  __shared__ long long split_s;
  const int64_t total = na + nb;
  const int t = threadIdx.x;
  const int64_t tile_begin = (tiles * blockIdx.x) / gridDim.x;
  const int64_t tile_end = (tiles * (blockIdx.x + 1)) / gridDim.x;
  if (t < 32) {
    const int64_t s0 = split_warp<ID_FIRST>(A, na, B, nb, tile_begin * RS_TILE);
    if (t == 0) split_s = s0;
  }
  __syncthreads();
  unsigned long long sum = 0;
  for (int64_t tile = tile_begin; tile < tile_end; ++tile) {
    const int64_t i0 = split_s;
    const int64_t k0 = tile * RS_TILE, j0 = k0 - i0;
    const int len = (int)((total - k0) < (int64_t)RS_TILE ? (total - k0) : (int64_t)RS_TILE);
    const int la = (int)((na - i0) < (int64_t)len ? (na - i0) : (int64_t)len);
    const int lb = (int)((nb - j0) < (int64_t)len ? (nb - j0) : (int64_t)len);
    {
      float ra[RS_ITEMS], rb[RS_ITEMS];   // every global load in flight before the first store
#pragma unroll
      for (int r = 0; r < RS_ITEMS; ++r) {
        const int i = t + r * RS_THREADS;
        ra[r] = i < la ? __ldg(A + i0 + i) : 0.f;
        rb[r] = i < lb ? __ldg(B + j0 + i) : 0.f;
      }
#pragma unroll
      for (int r = 0; r < RS_ITEMS; ++r) {
        const int i = t + r * RS_THREADS;
        sa[i] = ra[r];
        sb[i] = rb[r];
      }
    }
    __syncthreads();
    const int ka = t * RS_ITEMS;
    if (ka < len) {
      const int kb = (ka + RS_ITEMS) < len ? (ka + RS_ITEMS) : len;
      int i = split_smem<ID_FIRST>(sa, la, sb, lb, ka);
      int j = ka - i;
      for (int k = ka; k < kb; ++k) {
        const bool id = (j >= lb) || (i < la && takes_id<ID_FIRST>(sa[i], sb[j]));
        if (id) ++i;
        else { sum += (unsigned long long)(i0 + i); ++j; }   // ID scores merged before this OOD score
      }
      if (kb == len) split_s = i0 + i;            // exactly one thread: the next tile's split
    }
    __syncthreads();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
  if ((t & 31) == 0) warp_part[t >> 5] = sum;
  __syncthreads();
  if (t == 0) {
    unsigned long long s = 0;
    for (int w = 0; w < RS_THREADS / 32; ++w) s += warp_part[w];
    if (s) atomicAdd(acc, s);
  }
}

constexpr int SUM_BLOCKS = 592;

__device__ double ratio(int64_t x, int64_t y) { return (x + y) ? (double)x / (double)(x + y) : 0.0; }

// everything scalar, one thread: a few dozen binary searches on the sorted arrays
__global__ void score_scalar_kernel(const float* __restrict__ a /*sorted ID*/, int64_t n_id,
                                    const float* __restrict__ b /*sorted OOD*/, int64_t n_ood,
                                    uq_score_request req, const double* __restrict__ sum_partials,
                                    const unsigned long long* __restrict__ auroc_acc,
                                    uq_score_result* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uq_score_result r;
  // ---- mean / max / np.percentile(method='linear') of the ID scores ------------------------------
  double s = 0.0;
  for (int i = 0; i < SUM_BLOCKS; ++i) s += sum_partials[i];
  r.mean_score = s / (double)n_id;
  r.max_score = (double)a[n_id - 1];
  {
    // np.percentile on a float32 array works in float32 throughout (NEP 50: the Python-float q
    // is "weak"): q / float32(100), virtual index (n - 1) * q, weight = index - floor(index)
    const float qf = __fdiv_rn((float)req.percentile_q, 100.0f);
    const float vi = __fmul_rn((float)(n_id - 1), qf);
    float fl = floorf(vi);
    if (fl > (float)(n_id - 1)) fl = (float)(n_id - 1);
    const int64_t i0 = (int64_t)fl, i1 = i0 + 1 < n_id ? i0 + 1 : n_id - 1;
    // a + (b - a) t, or b - (b - a)(1 - t) when t >= 0.5 (numpy/lib/_function_base_impl.py:_lerp)
    const float t = __fsub_rn(vi, fl);
    const float lo = a[i0], hi = a[i1];
    const float diff = __fsub_rn(hi, lo);
    const float v = (t >= 0.5f) ? __fsub_rn(hi, __fmul_rn(diff, __fsub_rn(1.0f, t)))
                                : __fadd_rn(lo, __fmul_rn(diff, t));
    r.percentile_score = (double)v;
  }
  // ---- AUROC --------------------------------------------------------------------------------------
  r.auroc = (double)(*auroc_acc) / (2.0 * (double)n_id * (double)n_ood);
  // ---- TNR at TPR (reference quirks kept: see oracle/metrics_oracle.py:tnr_at_tpr) -----------------
  {
    const bool rev = req.tnr_reversed != 0;
    const float* pos = rev ? a : b;
    const float* neg = rev ? b : a;
    const int64_t n_pos = rev ? n_id : n_ood, n_neg = rev ? n_ood : n_id;
    double tnr;
    if (rev ? (a[0] > b[n_ood - 1]) : (a[n_id - 1] < b[0])) {
      tnr = 1.0;
    } else {
      // smallest m with m / n_ood >= target in float64
      const double dn = (double)n_ood;
      int64_t m = (int64_t)ceil(req.target_tpr * dn);
      while (m > 0 && (double)(m - 1) / dn >= req.target_tpr) --m;
      while ((double)m / dn < req.target_tpr) ++m;
      if (m > n_pos) {
        tnr = 0.0;
      } else if (m == 0) {
        tnr = (double)n_neg / (double)n_id;
      } else {
        const float c = pos[n_pos - m];
        const float mn = a[0] < b[0] ? a[0] : b[0];
        tnr = (mn < c) ? (double)lower_bound_f(neg, n_neg, c) / (double)n_id : 0.0;
      }
    }
    r.tnr_at_tpr = tnr;
  }
  // ---- percentile classifier: threshold = torch.quantile(id, p) in float32 ------------------------
  {
    const bool rev = req.classifier_reversed != 0;
    // sorted(-id)[k] = -a[n_id - 1 - k]
    auto sid = [&](int64_t k) -> float { return rev ? -a[n_id - 1 - k] : a[k]; };
    const float rank = __fmul_rn((float)req.classifier_percentile, (float)(n_id - 1));
    const float below = floorf(rank), above = ceilf(rank);
    const float w = __fsub_rn(rank, below);
    const float lo = sid((int64_t)below), hi = sid((int64_t)above);
    const float diff = __fsub_rn(hi, lo);
    const float thr = (w < 0.5f) ? __fadd_rn(lo, __fmul_rn(w, diff))
                                 : __fsub_rn(hi, __fmul_rn(diff, __fsub_rn(1.0f, w)));
    int64_t id_above, ood_above;
    if (!rev) {
      id_above = n_id - upper_bound_f(a, n_id, thr);
      ood_above = n_ood - upper_bound_f(b, n_ood, thr);
    } else {  // #(-x > thr) = #(x < -thr)
      id_above = lower_bound_f(a, n_id, -thr);
      ood_above = lower_bound_f(b, n_ood, -thr);
    }
    const int64_t id_below = n_id - id_above, ood_below = n_ood - ood_above;
    r.sensitivity = ratio(ood_above, ood_below);
    r.specificity = ratio(id_below, id_above);
    r.fpr = ratio(id_above, id_below);
    r.fnr = ratio(ood_below, ood_above);
  }
  *out = r;
}

struct WsLayout {
  size_t a, at, b, bt, scratch, partials, acc, result, total;
};

WsLayout layout(int64_t n_id, int64_t n_ood) {
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  WsLayout L;
  size_t o = 0;
  L.a = o; o += al(sizeof(float) * (size_t)n_id);
  L.at = o; o += al(sizeof(float) * (size_t)n_id);
  L.b = o; o += al(sizeof(float) * (size_t)n_ood);
  L.bt = o; o += al(sizeof(float) * (size_t)n_ood);
  const size_t sa = radix_sort_scratch_bytes(n_id), sb = radix_sort_scratch_bytes(n_ood);
  L.scratch = o; o += al(sa > sb ? sa : sb);
  L.partials = o; o += al(sizeof(double) * SUM_BLOCKS);
  L.acc = o; o += 256;
  L.result = o; o += al(sizeof(uq_score_result));
  L.total = o;
  return L;
}

}  // namespace
}  // namespace uq

using namespace uq;

extern "C" {

size_t uq_score_metrics_workspace_bytes(int64_t n_id, int64_t n_ood) {
  if (n_id < 1 || n_ood < 1) return 0;
  return layout(n_id, n_ood).total;
}

int uq_score_metrics(const float* id_scores, int64_t n_id, const float* ood_scores, int64_t n_ood,
                     const uq_score_request* req, uq_score_result* out_host, void* workspace,
                     size_t workspace_bytes, void* stream) {
  UQ_REQUIRE(id_scores && ood_scores && req && out_host && n_id >= 1 && n_ood >= 1, UQ_ERR_INVALID,
             "uq_score_metrics: NULL argument or empty score vector");
  UQ_REQUIRE(req->percentile_q >= 0.0 && req->percentile_q <= 100.0, UQ_ERR_INVALID,
             "percentile must be between 0 and 100, got %g", req->percentile_q);
  UQ_REQUIRE(req->target_tpr >= 0.0 && req->target_tpr <= 1.0, UQ_ERR_INVALID,
             "target_tpr must be between 0 and 1, got %g", req->target_tpr);
  UQ_REQUIRE(req->classifier_percentile >= 0.0 && req->classifier_percentile <= 1.0,
             UQ_ERR_INVALID, "Percentile must be between 0 and 1, got %g",
             req->classifier_percentile);
  const WsLayout L = layout(n_id, n_ood);
  UQ_REQUIRE(workspace != nullptr && workspace_bytes >= L.total, UQ_ERR_WORKSPACE,
             "uq_score_metrics needs %zu workspace bytes, got %zu", L.total, workspace_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* w = static_cast<char*>(workspace);
  float* da = reinterpret_cast<float*>(w + L.a);
  float* dat = reinterpret_cast<float*>(w + L.at);
  float* db = reinterpret_cast<float*>(w + L.b);
  float* dbt = reinterpret_cast<float*>(w + L.bt);
  double* partials = reinterpret_cast<double*>(w + L.partials);
  unsigned long long* acc = reinterpret_cast<unsigned long long*>(w + L.acc);
  uq_score_result* result = reinterpret_cast<uq_score_result*>(w + L.result);
  float *sa = nullptr, *sb = nullptr;
  int rc = radix_sort_f32_copy(id_scores, da, dat, n_id, w + L.scratch,
                               radix_sort_scratch_bytes(n_id), &sa, st);
  if (rc != UQ_OK) return rc;
  rc = radix_sort_f32_copy(ood_scores, db, dbt, n_ood, w + L.scratch,
                           radix_sort_scratch_bytes(n_ood), &sb, st);
  if (rc != UQ_OK) return rc;
  sum_kernel<<<SUM_BLOCKS, 256, 0, st>>>(sa, n_id, partials);
  UQ_LAUNCH_CHECK();
  UQ_CUDA(cudaMemsetAsync(acc, 0, sizeof(unsigned long long), st));
  const int64_t tiles = (n_id + n_ood + RS_TILE - 1) / RS_TILE;
  const unsigned grid = (unsigned)(tiles < 148 * 4 ? tiles : 148 * 4);
  rank_sum_kernel<true><<<grid, RS_THREADS, 0, st>>>(sa, n_id, sb, n_ood, tiles, acc);   // #ID <= v
  UQ_LAUNCH_CHECK();
  rank_sum_kernel<false><<<grid, RS_THREADS, 0, st>>>(sa, n_id, sb, n_ood, tiles, acc);  // #ID < v
  UQ_LAUNCH_CHECK();
  score_scalar_kernel<<<1, 32, 0, st>>>(sa, n_id, sb, n_ood, *req, partials, acc, result);
  UQ_LAUNCH_CHECK();
  UQ_CUDA(cudaMemcpyAsync(out_host, result, sizeof(uq_score_result), cudaMemcpyDeviceToHost, st));
  UQ_CUDA(cudaStreamSynchronize(st));
  return UQ_OK;
}

}  // extern "C"
