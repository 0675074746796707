// Input-density uncertainty score of KDEMLPModel (nnueehcs/models.py:191-222):
//   kde = sklearn.neighbors.KernelDensity(bandwidth='scott', rtol=...).fit(train)   (fit_kde)
//   dens = -exp(kde.score_samples(x))                                                (forward)
// i.e. for every query row x_n the negated d-dimensional Gaussian KDE of the M fitted rows,
//   dens[n] = -(1 / (M h^d (2 pi)^(d/2))) * sum_j exp(-|x_n - y_j|^2 / (2 h^2)),
// with sklearn's 'scott' bandwidth h = M^(-1/(d+4)) (a plain scalar, not scaled by the data's
// spread).  sklearn walks a KD-tree in float64 and stops refining at its rtol; this kernel adds
// every term (CUDA-core FMAs + MUFU.EX2), so it sits inside the reference's own tolerance.
//
// Shape: N x M pair evaluations, d small (binomial options: 5) -> compute-bound on the FP32 and
// MUFU pipes, not on HBM (inputs are read once per tile and re-used from registers / shared
// memory).  Each thread keeps Q = 4 query rows in registers, a block of 256 threads therefore a
// 1024-row query tile; the fitted rows stream through shared memory in tiles of 512 (every lane
// reads the same row: a broadcast, no bank conflicts).  A tile's terms are summed in float32
// and folded into a float64 accumulator per query; with few query tiles the fitted rows are
// additionally split over blockIdx.y and the per-split partials added in a fixed order
// (deterministic result).
#include "common.cuh"

namespace uq {
namespace {

constexpr int KD_THREADS = 256;
constexpr int KD_Q = 4;                       // query rows per thread
constexpr int KD_TILE = 512;                  // fitted rows per shared-memory tile
constexpr int KD_ROWS = KD_THREADS * KD_Q;    // query rows per block
constexpr int KD_MAX_D = 32;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// D > 0: compile-time width; D == 0: runtime width d <= KD_MAX_D (query rows re-read from shared)
template <int D>
__global__ void __launch_bounds__(KD_THREADS)
kde_density_kernel(const float* __restrict__ fit, int64_t m, const float* __restrict__ x, int64_t n,
                   int d_rt, float neg_half_inv_h2_log2e, int64_t rows_per_split,
                   double* __restrict__ partial) {
  const int d = D > 0 ? D : d_rt;
  extern __shared__ float kd_sh[];            // [KD_TILE][d] fitted rows (+ [KD_ROWS][d] queries if D == 0)
  float* sy = kd_sh;
  float* sx = kd_sh + KD_TILE * d;
  const int t = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * KD_ROWS;
  const int64_t j_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t j_end = (j_begin + rows_per_split) < m ? (j_begin + rows_per_split) : m;

  // query rows: thread t owns rows row0 + t + q * KD_THREADS (coalesced across the block)
  float xq[KD_Q][D > 0 ? D : 1];
  if (D > 0) {
#pragma unroll
    for (int q = 0; q < KD_Q; ++q) {
      const int64_t r = row0 + t + (int64_t)q * KD_THREADS;
#pragma unroll
      for (int i = 0; i < D; ++i) xq[q][i] = r < n ? __ldg(x + r * D + i) : 0.f;
    }
  } else {
    for (int i = t; i < KD_ROWS * d; i += KD_THREADS) {
      const int64_t g = row0 * d + i;
      sx[i] = g < n * d ? __ldg(x + g) : 0.f;
    }
  }
  double acc[KD_Q];
#pragma unroll
  for (int q = 0; q < KD_Q; ++q) acc[q] = 0.0;

  for (int64_t j0 = j_begin; j0 < j_end; j0 += KD_TILE) {
    const int cnt = (int)((j_end - j0) < KD_TILE ? (j_end - j0) : KD_TILE);
    __syncthreads();
    for (int i = t; i < cnt * d; i += KD_THREADS) sy[i] = __ldg(fit + j0 * d + i);
    __syncthreads();
    float part[KD_Q];
#pragma unroll
    for (int q = 0; q < KD_Q; ++q) part[q] = 0.f;
    if (D > 0) {
#pragma unroll 4
      for (int j = 0; j < cnt; ++j) {
        float y[D > 0 ? D : 1];
#pragma unroll
        for (int i = 0; i < D; ++i) y[i] = sy[j * D + i];
#pragma unroll
        for (int q = 0; q < KD_Q; ++q) {
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < D; ++i) {
            const float df = xq[q][i] - y[i];
            s = fmaf(df, df, s);
          }
          part[q] += ex2_approx(s * neg_half_inv_h2_log2e);
        }
      }
    } else {
      for (int j = 0; j < cnt; ++j) {
#pragma unroll
        for (int q = 0; q < KD_Q; ++q) {
          const float* xr = sx + (t + q * KD_THREADS) * d;
          float s = 0.f;
          for (int i = 0; i < d; ++i) {
            const float df = xr[i] - sy[j * d + i];
            s = fmaf(df, df, s);
          }
          part[q] += ex2_approx(s * neg_half_inv_h2_log2e);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < KD_Q; ++q) acc[q] += (double)part[q];
  }
#pragma unroll
  for (int q = 0; q < KD_Q; ++q) {
    const int64_t r = row0 + t + (int64_t)q * KD_THREADS;
    if (r < n) partial[(int64_t)blockIdx.y * n + r] = acc[q];
  }
}

// Queries whose kernel sum is so small that float32 terms flushed to zero (below 2^-126) could
// matter are listed for the float64 pass below: with s >= m 2^-100 the at most m lost terms are
// below 2^-26 of the sum; anything smaller is re-done.  (The reference keeps such far-OOD scores
// apart down to 1e-308 -- score_samples works in log space -- and every score consumer ranks them.)
__global__ void kde_density_finish_kernel(const double* __restrict__ partial, int splits, int64_t n,
                                          double scale, double rescue_below,
                                          double* __restrict__ out,
                                          unsigned long long* __restrict__ n_far,
                                          int64_t* __restrict__ far_rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int k = 0; k < splits; ++k) s += partial[(int64_t)k * n + i];
  out[i] = -s * scale;
  if (s < rescue_below) far_rows[atomicAdd(n_far, 1ull)] = i;
}

// Far queries, one block each, in float64 and in log space: every thread folds its share of the
// fitted rows into a running (max exponent, sum of exp(e - max)) pair, the pairs are merged in a
// fixed order, and dens = -exp(log_norm + max) * sum -- which underflows to -0.0 only where the
// reference's exp(score_samples) does.  Terms more than 45 below the running maximum (4e-20 of
// it) are skipped without evaluating exp.
constexpr int KF_THREADS_FAR = 256;
__global__ void __launch_bounds__(KF_THREADS_FAR)
kde_density_far_kernel(const float* __restrict__ fit, int64_t m, const float* __restrict__ x, int d,
                       double neg_half_inv_h2, double log_norm,
                       const unsigned long long* __restrict__ n_far,
                       const int64_t* __restrict__ far_rows, double* __restrict__ out) {
  __shared__ double sh_m[KF_THREADS_FAR], sh_s[KF_THREADS_FAR];
  __shared__ float xs[KD_MAX_D];
  const unsigned long long count = *n_far;
  for (unsigned long long q = blockIdx.x; q < count; q += gridDim.x) {
    const int64_t row = far_rows[q];
    __syncthreads();
    if (threadIdx.x < d) xs[threadIdx.x] = x[row * d + threadIdx.x];
    __syncthreads();
    double mx = -1.0e300, sum = 0.0;
    for (int64_t j = threadIdx.x; j < m; j += KF_THREADS_FAR) {
      double s2 = 0.0;
      for (int i = 0; i < d; ++i) {
        const double df = (double)xs[i] - (double)__ldg(fit + j * d + i);
        s2 = fma(df, df, s2);
      }
      const double e = s2 * neg_half_inv_h2;
      if (e > mx) {
        sum = (mx - e > -45.0 ? sum * exp(mx - e) : 0.0) + 1.0;
        mx = e;
      } else if (e - mx > -45.0) {
        sum += exp(e - mx);
      }
    }
    sh_m[threadIdx.x] = mx, sh_s[threadIdx.x] = sum;
    __syncthreads();
    for (int o = KF_THREADS_FAR / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) {
        const double ma = sh_m[threadIdx.x], mb = sh_m[threadIdx.x + o];
        const double sa = sh_s[threadIdx.x], sb = sh_s[threadIdx.x + o];
        const double mm = ma > mb ? ma : mb;
        sh_s[threadIdx.x] = sa * exp(ma - mm) + sb * exp(mb - mm);   // empty shares: s = 0
        sh_m[threadIdx.x] = mm;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) out[row] = -exp(log_norm + sh_m[0]) * sh_s[0];
  }
}

int choose_splits(int64_t n, int64_t m) {
  const int64_t tiles = (n + KD_ROWS - 1) / KD_ROWS;
  int64_t splits = 1;
  while (tiles * splits < 2 * 148 && (m / (splits * 2)) >= 4 * KD_TILE) splits *= 2;
  return (int)splits;
}

template <int D>
int launch(const float* fit, int64_t m, const float* x, int64_t n, int d, float c, int splits,
           int64_t rows_per_split, double* partial, cudaStream_t st) {
  const size_t smem = sizeof(float) * ((size_t)KD_TILE * d + (D > 0 ? 0 : (size_t)KD_ROWS * d));
  if (smem > 48 * 1024)
    UQ_CUDA(cudaFuncSetAttribute(kde_density_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  dim3 grid((unsigned)((n + KD_ROWS - 1) / KD_ROWS), (unsigned)splits);
  kde_density_kernel<D><<<grid, KD_THREADS, smem, st>>>(fit, m, x, n, d, c, rows_per_split, partial);
  UQ_LAUNCH_CHECK();
  return UQ_OK;
}

}  // namespace

namespace {
size_t partial_bytes(int64_t n, int64_t m) {
  return (sizeof(double) * (size_t)choose_splits(n, m) * (size_t)n + 255) & ~(size_t)255;
}
}  // namespace

// [partial sums | far-query counter | far-query rows]
size_t kde_density_workspace_bytes(int64_t n, int64_t m) {
  if (n < 1 || m < 1) return 0;
  return partial_bytes(n, m) + 256 + sizeof(int64_t) * (size_t)n;
}

int kde_density(const float* fit, int64_t m, const float* x, int64_t n, int d, double bandwidth,
                double* out, void* ws, size_t ws_bytes, cudaStream_t st) {
  UQ_REQUIRE(d >= 1 && d <= KD_MAX_D, UQ_ERR_UNSUPPORTED,
             "kde_density: %d features (supported: 1..%d)", d, KD_MAX_D);
  UQ_REQUIRE(bandwidth > 0.0, UQ_ERR_INVALID, "kde_density: bandwidth must be positive");
  const size_t need = kde_density_workspace_bytes(n, m);
  UQ_REQUIRE(ws != nullptr && ws_bytes >= need, UQ_ERR_WORKSPACE,
             "kde_density needs %zu workspace bytes, got %zu", need, ws_bytes);
  const int splits = choose_splits(n, m);
  const int64_t rows_per_split = ((m + splits - 1) / splits + KD_TILE - 1) / KD_TILE * KD_TILE;
  double* partial = static_cast<double*>(ws);
  const float c = (float)(-0.5 / (bandwidth * bandwidth) * 1.4426950408889634);
  int rc;
  switch (d) {
    case 1: rc = launch<1>(fit, m, x, n, d, c, splits, rows_per_split, partial, st); break;
    case 2: rc = launch<2>(fit, m, x, n, d, c, splits, rows_per_split, partial, st); break;
    case 3: rc = launch<3>(fit, m, x, n, d, c, splits, rows_per_split, partial, st); break;
    case 4: rc = launch<4>(fit, m, x, n, d, c, splits, rows_per_split, partial, st); break;
    case 5: rc = launch<5>(fit, m, x, n, d, c, splits, rows_per_split, partial, st); break;
    case 6: rc = launch<6>(fit, m, x, n, d, c, splits, rows_per_split, partial, st); break;
    case 7: rc = launch<7>(fit, m, x, n, d, c, splits, rows_per_split, partial, st); break;
    case 8: rc = launch<8>(fit, m, x, n, d, c, splits, rows_per_split, partial, st); break;
    default: rc = launch<0>(fit, m, x, n, d, c, splits, rows_per_split, partial, st); break;
  }
  if (rc != UQ_OK) return rc;
  // normalisation of sklearn's Gaussian kernel and the 1/M of score_samples, float64
  const double log_norm = -(double)d * log(bandwidth) - 0.5 * (double)d * log(2.0 * M_PI) -
                          log((double)m);
  char* wsb = static_cast<char*>(ws);
  unsigned long long* n_far = reinterpret_cast<unsigned long long*>(wsb + partial_bytes(n, m));
  int64_t* far_rows = reinterpret_cast<int64_t*>(wsb + partial_bytes(n, m) + 256);
  UQ_CUDA(cudaMemsetAsync(n_far, 0, sizeof(unsigned long long), st));
  kde_density_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
      partial, splits, n, exp(log_norm), (double)m * 7.888609052210118e-31 /* m 2^-100 */, out,
      n_far, far_rows);
  UQ_LAUNCH_CHECK();
  // far queries (none for in-distribution inputs: the blocks read the counter and leave)
  int64_t far_blocks = n < 148 * 8 ? n : 148 * 8;
  kde_density_far_kernel<<<(unsigned)far_blocks, KF_THREADS_FAR, 0, st>>>(
      fit, m, x, d, -0.5 / (bandwidth * bandwidth), log_norm, n_far, far_rows, out);
  UQ_LAUNCH_CHECK();
  return UQ_OK;
}

}  // namespace uq
