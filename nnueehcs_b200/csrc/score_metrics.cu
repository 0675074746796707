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

// Mann-Whitney: sum over OOD scores of (#ID < v) + (#ID <= v)  (= 2U, exact in uint64)
__global__ void __launch_bounds__(256)
auroc_kernel(const float* __restrict__ id_sorted, int64_t n_id, const float* __restrict__ ood,
             int64_t n_ood, unsigned long long* __restrict__ acc) {
  __shared__ unsigned long long sh[8];
  unsigned long long s = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_ood;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float v = __ldg(ood + i);
    s += (unsigned long long)(lower_bound_f(id_sorted, n_id, v) + upper_bound_f(id_sorted, n_id, v));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    atomicAdd(acc, t);
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
  int64_t blocks = (n_ood + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  auroc_kernel<<<(unsigned)blocks, 256, 0, st>>>(sa, n_id, sb, n_ood, acc);
  UQ_LAUNCH_CHECK();
  score_scalar_kernel<<<1, 32, 0, st>>>(sa, n_id, sb, n_ood, *req, partials, acc, result);
  UQ_LAUNCH_CHECK();
  UQ_CUDA(cudaMemcpyAsync(out_host, result, sizeof(uq_score_result), cudaMemcpyDeviceToHost, st));
  UQ_CUDA(cudaStreamSynchronize(st));
  return UQ_OK;
}

}  // extern "C"
