// Jensen-Shannon distance between two Scott-bandwidth Gaussian KDEs on a shared linspace grid.
//
// Replaces JensenShannonEvaluation.pdf_jsd (nnueehcs/evaluation.py:268-276):
//     kde1 = gaussian_kde(dist1); kde2 = gaussian_kde(dist2)
//     x    = linspace(min(both), max(both), 20000)
//     jensenshannon(kde1(x), kde2(x))
// scipy's KDE is pdf(g) = 1/(n h sqrt(2 pi)) sum_i exp(-((x_i-g)/h)^2/2) with
// h = sqrt(unbiased var) * n^(-1/5); jensenshannon renormalises each pdf vector to sum 1, so the
// constant in front cancels and only the raw kernel sums are needed.
//
// Work is N x G Gaussian evaluations if done naively (2e12 at 100 M values).  Here the samples
// are radix-sorted first, so a chunk of 2048 consecutive values only touches the grid points
// within 9 h of its value range (a Gaussian term beyond 9 sigma is < 3e-18 of the peak) -- about
// 500 of the 20 000 points at 50 M values per sample -- and the inner loop runs in float32 on
// coordinates taken relative to the chunk's window origin (computed in float64, so no
// cancellation), one MUFU.EX2 per pair, flushed to float64 every 256 terms.  Each block adds its
// window's partial sums to the float64 grid with atomics; a last block turns the two grids into
// the JS distance.
//
// Moment method (default when it applies; uq_kde_jsd_ex's `method`): no sort and no N x G work.
// The value range is cut into bins of width h/4; for a value at offset eps*h from its bin centre c
//   exp(-((g - c)/h - eps)^2 / 2) = K(z) * exp(z eps - eps^2/2) = K(z) * sum_m He_m(z) eps^m / m!,
// z = (g - c)/h, so the kernel sum at every grid point follows from six numbers per bin -- the
// count and sum eps^m, m = 1..5 -- and a [grid point x 77 bins] evaluation in float64 that no
// longer depends on N.  uq_kde_jsd runs this as ONE cooperative launch (csrc/kde_fused.cu: three
// shared atomics per value on finer bins, folded into the six moments).  The six-atomic kernel
// below (kde_moments_kernel) serves the sharded path (kde_grid_accumulate), where bandwidth and
// range come from the all-reduced statistics.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "sort.cuh"

namespace uq {
namespace {

constexpr int KDE_THREADS = 256;
constexpr int KDE_CHUNK = 2048;      // sorted samples per block
constexpr int KDE_SUB = 256;         // float32 accumulation run before flushing to float64
constexpr double KDE_Z = 9.0;        // truncation in kernel standard deviations
constexpr int STAT_BLOCKS = 296;

struct KdeParams {       // written by setup_kernel, read by eval_kernel
  double lo, step;       // grid: g_j = lo + j * step
  double h[2];           // bandwidth (kernel std) of each sample
  double scale[2];       // sqrt(log2(e)/2) / h : exp(-z^2/2) == exp2(-((x-g)*scale)^2)
};

// per-block float64 partial sums of (x - x0) and (x - x0)^2 on the sorted array (x0 = minimum)
__global__ void __launch_bounds__(256)
stats_kernel(const float* __restrict__ xs, int64_t n, double* __restrict__ partials) {
  __shared__ double sh[2][8];
  const double x0 = (double)xs[0];
  double s1 = 0.0, s2 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double d = (double)xs[i] - x0;
    s1 += d;
    s2 += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_down_sync(0xffffffffu, s1, o);
    s2 += __shfl_down_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = s1;
    sh[1][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; ++w) a += sh[0][w], b += sh[1][w];
    partials[2 * blockIdx.x] = a;
    partials[2 * blockIdx.x + 1] = b;
  }
}

__global__ void setup_kernel(const float* __restrict__ us, int64_t nu, const float* __restrict__ vs,
                             int64_t nv, const double* __restrict__ pu,
                             const double* __restrict__ pv, int grid_pts, KdeParams* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s1 = 0, s2 = 0;
  for (int b = 0; b < STAT_BLOCKS; ++b) s1 += pu[2 * b], s2 += pu[2 * b + 1];
  const double var_u = (s2 - s1 * s1 / (double)nu) / (double)(nu - 1);
  s1 = 0, s2 = 0;
  for (int b = 0; b < STAT_BLOCKS; ++b) s1 += pv[2 * b], s2 += pv[2 * b + 1];
  const double var_v = (s2 - s1 * s1 / (double)nv) / (double)(nv - 1);
  KdeParams p;
  p.h[0] = sqrt(var_u) * pow((double)nu, -0.2);
  p.h[1] = sqrt(var_v) * pow((double)nv, -0.2);
  const double lo = fmin((double)us[0], (double)vs[0]);
  const double hi = fmax((double)us[nu - 1], (double)vs[nv - 1]);
  p.lo = lo;
  p.step = (hi - lo) / (double)(grid_pts - 1);
  const double c = sqrt(0.5 * 1.4426950408889634);
  p.scale[0] = c / p.h[0];
  p.scale[1] = c / p.h[1];
  *out = p;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One pass over the staged chunk for GP grid points per thread (columns jb + g * 256 + tid).
template <int GP>
__device__ __forceinline__ void kde_pass(const float* sx, int jb, int jlo, int jhi, double step_s,
                                         double* __restrict__ out) {
  float b[GP];
  double acc64[GP];
#pragma unroll
  for (int g = 0; g < GP; ++g) {
    const int j = jb + g * KDE_THREADS + threadIdx.x;
    b[g] = (float)((double)(j - jlo) * step_s);
    acc64[g] = 0.0;
  }
  for (int s0 = 0; s0 < KDE_CHUNK; s0 += KDE_SUB) {
    float acc[GP];
#pragma unroll
    for (int g = 0; g < GP; ++g) acc[g] = 0.f;
#pragma unroll 4
    for (int i = 0; i < KDE_SUB; i += 4) {
      const float4 a = *reinterpret_cast<const float4*>(&sx[s0 + i]);  // warp-broadcast
#pragma unroll
      for (int g = 0; g < GP; ++g) {
        const float d0 = a.x - b[g], d1 = a.y - b[g], d2 = a.z - b[g], d3 = a.w - b[g];
        acc[g] += ex2_approx(-d0 * d0) + ex2_approx(-d1 * d1) +
                  (ex2_approx(-d2 * d2) + ex2_approx(-d3 * d3));
      }
    }
#pragma unroll
    for (int g = 0; g < GP; ++g) acc64[g] += (double)acc[g];
  }
#pragma unroll
  for (int g = 0; g < GP; ++g) {
    const int j = jb + g * KDE_THREADS + threadIdx.x;
    if (j <= jhi && acc64[g] != 0.0) atomicAdd(&out[j], acc64[g]);
  }
}

// blockIdx.x < chunks_u : chunk of the first sample, else of the second
__global__ void __launch_bounds__(KDE_THREADS)
kde_eval_kernel(const float* __restrict__ us, int64_t nu, const float* __restrict__ vs, int64_t nv,
                int64_t chunks_u, int grid_pts, const KdeParams* __restrict__ params,
                double* __restrict__ pdf /* [2][grid_pts] */) {
  __shared__ __align__(16) float sx[KDE_CHUNK];
  const int which = (int64_t)blockIdx.x < chunks_u ? 0 : 1;
  const float* xs = which ? vs : us;
  const int64_t n = which ? nv : nu;
  const int64_t c0 = ((int64_t)blockIdx.x - (which ? chunks_u : 0)) * KDE_CHUNK;
  const int64_t c1 = (c0 + KDE_CHUNK) < n ? (c0 + KDE_CHUNK) : n;
  const KdeParams p = *params;
  const double h = p.h[which], scale = p.scale[which];
  double* out = pdf + (size_t)which * grid_pts;

  // window of grid indices within KDE_Z bandwidths of this chunk's value range
  const double x_first = (double)xs[c0], x_last = (double)xs[c1 - 1];
  int jlo = 0, jhi = grid_pts - 1;
  if (p.step > 0.0) {
    const double a = ceil((x_first - KDE_Z * h - p.lo) / p.step);
    const double b = floor((x_last + KDE_Z * h - p.lo) / p.step);
    if (a > 0.0) jlo = a > (double)(grid_pts - 1) ? grid_pts - 1 : (int)a;
    if (b < (double)(grid_pts - 1)) jhi = b < 0.0 ? 0 : (int)b;
  }
  const double g0 = p.lo + (double)jlo * p.step;  // window origin, float64

  // stage the chunk as scaled coordinates relative to the window origin
  for (int i = threadIdx.x; i < KDE_CHUNK; i += KDE_THREADS) {
    const int64_t gi = c0 + i;
    sx[i] = gi < c1 ? (float)(((double)xs[gi] - g0) * scale) : 3.0e18f;  // pad: exp2(-inf) = 0
  }
  __syncthreads();

  // passes of GP grid points per thread; GP follows what is left of the window, so a ~480-point
  // window costs 512 point-columns of MUFU work, not 1024
  const double step_s = p.step * scale;
  int jb = jlo;
  while (jb <= jhi) {
    const int rem = jhi - jb + 1;
    if (rem > 3 * KDE_THREADS) kde_pass<4>(sx, jb, jlo, jhi, step_s, out), jb += 4 * KDE_THREADS;
    else if (rem > 2 * KDE_THREADS) kde_pass<3>(sx, jb, jlo, jhi, step_s, out), jb += 3 * KDE_THREADS;
    else if (rem > KDE_THREADS) kde_pass<2>(sx, jb, jlo, jhi, step_s, out), jb += 2 * KDE_THREADS;
    else kde_pass<1>(sx, jb, jlo, jhi, step_s, out), jb += KDE_THREADS;
  }
}

// ---- moment method ---------------------------------------------------------------------------------

constexpr int KM_PER_H = 4;          // bins per bandwidth
constexpr int KM_ORDER = 5;          // highest eps power kept
constexpr int KM_WORDS = KM_ORDER + 1;
constexpr int KM_MAX_BINS = 8192;    // 8192 x 6 x 4 B = 192 KB of shared memory
constexpr int KM_THREADS = 1024;

// One pass over a sample: per bin the count and sum eps^m (m = 1..5), block-private in shared
// memory, flushed to float64 global tables [bin][6].  A float atomicAdd on shared memory is a CAS
// loop in SASS (ATOMS.CAST.SPIN; first version: 440 us per 50 M values), so the moments are
// accumulated in fixed point with native 32-bit ATOMS.ADD: term_m = (eps^m + o_m) * S_m with
// o_m = 2 * 8^-m (keeps it positive; |eps| <= 1/8) and S_m = 2^19 * 8^m, i.e. 2^19 .. 3 * 2^19 per
// term, converted with the 2^23 magic-number add (FADD + LOP instead of a quarter-rate F2I).  A
// word wraps at most once per 2730 adds; the add that sees the wrap (old + term < old) credits
// 2^32 / S_m to the global table directly.  The offset c * o_m leaves at the flush (exact: the
// count is an integer).
__device__ __forceinline__ uint32_t km_fixed(float term) {   // 0 <= term < 2^23, round to nearest
  return __float_as_uint(term + 8388608.0f) & 0x7FFFFFu;
}

__global__ void __launch_bounds__(KM_THREADS, 1)
kde_moments_kernel(const float* __restrict__ x, int64_t n, double lo, double inv_w, int nb,
                   double* __restrict__ tables) {
  extern __shared__ uint32_t km_sh[];
  uint32_t* cnt = km_sh;                       // [nb]
  uint32_t* mom = km_sh + nb;                  // [KM_ORDER][nb]
  for (int i = threadIdx.x; i < KM_WORDS * nb; i += KM_THREADS) km_sh[i] = 0;
  __syncthreads();
  constexpr float S1 = 4194304.f, S2 = S1 * 8.f, S3 = S2 * 8.f, S4 = S3 * 8.f, S5 = S4 * 8.f;
  auto add = [&](float v) {
    const double t = ((double)v - lo) * inv_w;                      // float64: bin offsets stay exact
    int b = (int)t;
    b = b < 0 ? 0 : (b >= nb ? nb - 1 : b);
    float e = (float)((t - (double)b - 0.5) * (1.0 / KM_PER_H));
    e = fminf(fmaxf(e, -0.126f), 0.126f);      // |eps| <= 1/8 by construction; keeps every term in (0, 2^21)
    const float e2 = e * e;
    // o_m * S_m = 2 * 8^-m * 2^19 * 8^m = 2^20 for every m
    const uint32_t t1 = km_fixed(fmaf(e, S1, 1048576.f));
    const uint32_t t2 = km_fixed(fmaf(e2, S2, 1048576.f));
    const uint32_t t3 = km_fixed(fmaf(e2 * e, S3, 1048576.f));
    const uint32_t t4 = km_fixed(fmaf(e2 * e2, S4, 1048576.f));
    const uint32_t t5 = km_fixed(fmaf(e2 * e2 * e, S5, 1048576.f));
    uint32_t* cell = mom + b;
    atomicAdd(&cnt[b], 1u);
    const uint32_t o1 = atomicAdd(cell, t1);
    const uint32_t o2 = atomicAdd(cell + nb, t2);
    const uint32_t o3 = atomicAdd(cell + 2 * nb, t3);
    const uint32_t o4 = atomicAdd(cell + 3 * nb, t4);
    const uint32_t o5 = atomicAdd(cell + 4 * nb, t5);
    // a word wraps at most once per 2730 adds: one test for all five, the credits in the rare branch
    if ((o1 + t1 < o1) | (o2 + t2 < o2) | (o3 + t3 < o3) | (o4 + t4 < o4) | (o5 + t5 < o5)) {
      double* T = tables + (size_t)b * KM_WORDS + 1;
      if (o1 + t1 < o1) atomicAdd(T + 0, 4294967296.0 / S1);
      if (o2 + t2 < o2) atomicAdd(T + 1, 4294967296.0 / S2);
      if (o3 + t3 < o3) atomicAdd(T + 2, 4294967296.0 / S3);
      if (o4 + t4 < o4) atomicAdd(T + 3, 4294967296.0 / S4);
      if (o5 + t5 < o5) atomicAdd(T + 4, 4294967296.0 / S5);
    }
  };
  const int64_t head = min(n, (int64_t)((16 - ((uintptr_t)x & 15)) & 15) / 4);
  const int64_t n4 = (n - head) / 4;
  const float4* x4 = reinterpret_cast<const float4*>(x + head);
  const int64_t gtid = (int64_t)blockIdx.x * KM_THREADS + threadIdx.x;
  const int64_t gstride = (int64_t)gridDim.x * KM_THREADS;
  if (gtid < head) add(__ldg(x + gtid));
  int64_t i = gtid;
  for (; i + gstride < n4; i += 2 * gstride) {
    const float4 a0 = __ldg(x4 + i), a1 = __ldg(x4 + i + gstride);
    add(a0.x); add(a0.y); add(a0.z); add(a0.w);
    add(a1.x); add(a1.y); add(a1.z); add(a1.w);
  }
  for (; i < n4; i += gstride) {
    const float4 a0 = __ldg(x4 + i);
    add(a0.x); add(a0.y); add(a0.z); add(a0.w);
  }
  const int64_t tail0 = head + 4 * n4;
  if (tail0 + gtid < n) add(__ldg(x + tail0 + gtid));
  __syncthreads();
  const double inv_s[KM_ORDER] = {1.0 / S1, 1.0 / S2, 1.0 / S3, 1.0 / S4, 1.0 / S5};
  for (int b = threadIdx.x; b < nb; b += KM_THREADS) {
    const uint32_t c = cnt[b];
    if (c) {
      atomicAdd(&tables[(size_t)b * KM_WORDS], (double)c);
#pragma unroll
      for (int m = 0; m < KM_ORDER; ++m)  // (low word - c * 2^20) / S_m; the wraps were credited
        atomicAdd(&tables[(size_t)b * KM_WORDS + 1 + m],
                  ((double)mom[m * nb + b] - (double)c * 1048576.0) * inv_s[m]);
    }
  }
}

// grid[j] += sum over the bins within KDE_Z bandwidths of K(z) * sum_m He_m(z) M_m / m!
// Eight lanes share a grid point (each takes every eighth bin of the ~77 in reach) and combine
// with shuffles in a fixed order: 625 blocks instead of 79, a 10-step dependent chain instead of 77.
constexpr int KE_LANES = 8;
__global__ void __launch_bounds__(256)
kde_eval_bins_kernel(const double* __restrict__ tables, int nb, double lo, double w, double h,
                     double grid_lo, double grid_step, int grid_pts, double* __restrict__ grid) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = tid / KE_LANES, part = tid % KE_LANES;
  double acc = 0.0;
  if (j < grid_pts) {
    const double g = grid_lo + (double)j * grid_step;
    const double reach = KDE_Z * h + 0.5 * w;
    int b0 = (int)floor((g - reach - lo) / w), b1 = (int)floor((g + reach - lo) / w);
    b0 = b0 < 0 ? 0 : b0;
    b1 = b1 >= nb ? nb - 1 : b1;
    for (int b = b0 + part; b <= b1; b += KE_LANES) {
      const double* T = tables + (size_t)b * KM_WORDS;
      const double c = T[0];
      if (c == 0.0) continue;
      const double z = (g - (lo + ((double)b + 0.5) * w)) / h;
      const double z2 = z * z;
      const double he2 = z2 - 1.0, he3 = z * (z2 - 3.0), he4 = z2 * (z2 - 6.0) + 3.0,
                   he5 = z * (z2 * (z2 - 10.0) + 15.0);
      const double series = c + z * T[1] + he2 * T[2] * (1.0 / 2.0) + he3 * T[3] * (1.0 / 6.0) +
                            he4 * T[4] * (1.0 / 24.0) + he5 * T[5] * (1.0 / 120.0);
      acc += exp(-0.5 * z2) * series;
    }
  }
#pragma unroll
  for (int o = KE_LANES / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, KE_LANES);
  if (j < grid_pts && part == 0) grid[j] += acc;
}

// number of bins the moment method needs for one sample on [lo, hi], or 0 if it does not apply
int km_bins(double lo, double hi, double h) {
  if (!(h > 0.0) || !(hi >= lo)) return 0;
  const double nb = floor((hi - lo) / (h / KM_PER_H)) + 1.0;
  return nb <= (double)KM_MAX_BINS ? (int)nb : 0;
}

// adds the kernel sums of sample x (bandwidth h) to grid[grid_pts] = linspace(lo, hi); tables:
// KM_MAX_BINS * KM_WORDS doubles of scratch
int km_accumulate(const float* x, int64_t n, double lo, double hi, double h, int nb, int grid_pts,
                  double* grid, double* tables, cudaStream_t st) {
  static PerDeviceOnce opted;
  if (int rc = smem_opt_in(kde_moments_kernel, KM_MAX_BINS * KM_WORDS * (int)sizeof(uint32_t), opted))
    return rc;
  const double w = h / KM_PER_H;
  UQ_CUDA(cudaMemsetAsync(tables, 0, sizeof(double) * (size_t)nb * KM_WORDS, st));
  int64_t blocks = (n + (int64_t)KM_THREADS * 8 - 1) / ((int64_t)KM_THREADS * 8);
  if (blocks > 148) blocks = 148;
  if (blocks < 1) blocks = 1;
  kde_moments_kernel<<<(unsigned)blocks, KM_THREADS, (size_t)nb * KM_WORDS * sizeof(uint32_t), st>>>(
      x, n, lo, 1.0 / w, nb, tables);
  UQ_LAUNCH_CHECK();
  kde_eval_bins_kernel<<<(grid_pts * KE_LANES + 255) / 256, 256, 0, st>>>(
      tables, nb, lo, w, h, lo, (hi - lo) / (double)(grid_pts - 1), grid_pts, grid);
  UQ_LAUNCH_CHECK();
  return UQ_OK;
}

// scipy.spatial.distance.jensenshannon on the two raw kernel-sum vectors, two launches of
// JSD_BLOCKS blocks (one block doing all 20 000 float64 logs took 34 us): per-block sums of each
// vector, then per-block sums of p log(p/m) and q log(q/m) with the totals rebuilt from the
// block sums in a fixed order; the last block to finish (ticket counter) adds the per-block terms
// in a fixed order and writes the distance.  Deterministic.
constexpr int JSD_BLOCKS = 80;
constexpr int JSD_THREADS = 256;
constexpr size_t JSD_SCRATCH_BYTES = sizeof(double) * 4 * JSD_BLOCKS + 256;  // sums | terms | ticket

__device__ __forceinline__ void jsd_block_reduce2(double& a, double& b, double (*sh)[JSD_THREADS]) {
  const int t = threadIdx.x;
  sh[0][t] = a, sh[1][t] = b;
  __syncthreads();
  for (int o = JSD_THREADS / 2; o > 0; o >>= 1) {
    if (t < o) sh[0][t] += sh[0][t + o], sh[1][t] += sh[1][t + o];
    __syncthreads();
  }
  a = sh[0][0], b = sh[1][0];
  __syncthreads();
}

__global__ void __launch_bounds__(JSD_THREADS)
jsd_sums_kernel(const double* __restrict__ pdf, int grid_pts, double* __restrict__ scratch) {
  __shared__ double sh[2][JSD_THREADS];
  const int per = (grid_pts + JSD_BLOCKS - 1) / JSD_BLOCKS;
  const int j0 = blockIdx.x * per, j1 = min(grid_pts, j0 + per);
  double a = 0.0, b = 0.0;
  for (int j = j0 + threadIdx.x; j < j1; j += JSD_THREADS) a += pdf[j], b += pdf[grid_pts + j];
  jsd_block_reduce2(a, b, sh);
  if (threadIdx.x == 0) scratch[2 * blockIdx.x] = a, scratch[2 * blockIdx.x + 1] = b;
}

__global__ void __launch_bounds__(JSD_THREADS)
jsd_terms_kernel(const double* __restrict__ pdf, int grid_pts, double* __restrict__ scratch,
                 double* __restrict__ result) {
  __shared__ double sh[2][JSD_THREADS];
  __shared__ bool last;
  double* terms = scratch + 2 * JSD_BLOCKS;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + 4 * JSD_BLOCKS);
  double su = 0.0, sv = 0.0;
  for (int k = 0; k < JSD_BLOCKS; ++k) su += scratch[2 * k], sv += scratch[2 * k + 1];
  const int per = (grid_pts + JSD_BLOCKS - 1) / JSD_BLOCKS;
  const int j0 = blockIdx.x * per, j1 = min(grid_pts, j0 + per);
  double left = 0.0, right = 0.0;
  for (int j = j0 + threadIdx.x; j < j1; j += JSD_THREADS) {
    const double p = pdf[j] / su, q = pdf[grid_pts + j] / sv;
    const double m = (p + q) / 2.0;
    if (p > 0.0 && m > 0.0) left += p * log(p / m);
    if (q > 0.0 && m > 0.0) right += q * log(q / m);
  }
  jsd_block_reduce2(left, right, sh);
  if (threadIdx.x == 0) {
    terms[2 * blockIdx.x] = left, terms[2 * blockIdx.x + 1] = right;
    __threadfence();
    last = atomicAdd(ticket, 1u) == (unsigned int)(JSD_BLOCKS - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double l = 0.0, r = 0.0;
    for (int k = 0; k < JSD_BLOCKS; ++k) {
      l += reinterpret_cast<volatile double*>(terms)[2 * k];
      r += reinterpret_cast<volatile double*>(terms)[2 * k + 1];
    }
    *result = sqrt((l + r) / 2.0);
    *ticket = 0;  // ready for the next call on this scratch
  }
}

// scratch: JSD_SCRATCH_BYTES, its ticket word zero on first use (zeroed here)
int jsd_launch(const double* pdf, int grid_pts, double* scratch, double* result, cudaStream_t st) {
  UQ_CUDA(cudaMemsetAsync(scratch + 4 * JSD_BLOCKS, 0, 8, st));
  jsd_sums_kernel<<<JSD_BLOCKS, JSD_THREADS, 0, st>>>(pdf, grid_pts, scratch);
  UQ_LAUNCH_CHECK();
  jsd_terms_kernel<<<JSD_BLOCKS, JSD_THREADS, 0, st>>>(pdf, grid_pts, scratch, result);
  UQ_LAUNCH_CHECK();
  return UQ_OK;
}

struct WsLayout {
  size_t u, ut, v, vt, scratch, partials, params, pdf, result, jsd, fused, fused_bytes, total;
};

WsLayout layout(int64_t nu, int64_t nv, int grid_pts) {
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  WsLayout L;
  size_t o = 0;
  L.u = o; o += al(sizeof(float) * (size_t)nu);
  L.ut = o; o += al(sizeof(float) * (size_t)nu);
  L.v = o; o += al(sizeof(float) * (size_t)nv);
  L.vt = o; o += al(sizeof(float) * (size_t)nv);
  const size_t su = radix_sort_scratch_bytes(nu), sv = radix_sort_scratch_bytes(nv);
  L.scratch = o; o += al(su > sv ? su : sv);
  L.partials = o; o += al(sizeof(double) * 4 * STAT_BLOCKS);
  L.params = o; o += al(sizeof(KdeParams));
  L.pdf = o; o += al(sizeof(double) * 2 * (size_t)grid_pts);
  L.result = o; o += 256;
  L.jsd = o; o += al(JSD_SCRATCH_BYTES);
  L.fused_bytes = kde_jsd_fused_workspace_bytes(nu, nv, grid_pts);   // 0 if the device cannot run it
  L.fused = o; o += al(L.fused_bytes);
  L.total = o;
  return L;
}

}  // namespace

size_t kde_jsd_workspace_bytes(int64_t nu, int64_t nv, int grid_pts) {
  if (nu < 1 || nv < 1 || grid_pts < 2) return 0;
  return layout(nu, nv, grid_pts).total;
}

// enqueue / finish (see wasserstein_1d_enqueue): the single-launch moment method without a
// synchronisation inside the call; finish falls back to the synchronous call when the kernel
// declined (range / bandwidth / non-finite data) or the device cannot run it.
int kde_jsd(const float* u, int64_t nu, const float* v, int64_t nv, int grid_pts, int method,
            double* out_host, int* method_used_host, void* ws, size_t ws_bytes, cudaStream_t st);

int kde_jsd_enqueue(const float* u, int64_t nu, const float* v, int64_t nv, int grid_pts,
                    void* record, void* ws, size_t ws_bytes, cudaStream_t st) {
  const WsLayout L = layout(nu, nv, grid_pts);
  UQ_REQUIRE(ws != nullptr && ws_bytes >= L.total, UQ_ERR_WORKSPACE,
             "kde_jsd needs %zu workspace bytes, got %zu", L.total, ws_bytes);
  UQ_REQUIRE(record != nullptr, UQ_ERR_INVALID, "kde_jsd enqueue: record is NULL");
  if (L.fused_bytes == 0) {          // finish will run the synchronous call
    memset(record, 0, UQ_METRIC_RECORD_BYTES);
    return UQ_OK;
  }
  return kde_jsd_fused_enqueue(u, nu, v, nv, grid_pts, record, static_cast<char*>(ws) + L.fused,
                               L.fused_bytes, st);
}

int kde_jsd_finish(const float* u, int64_t nu, const float* v, int64_t nv, int grid_pts,
                   const void* record, double* out_host, int* method_used_host, void* ws,
                   size_t ws_bytes, cudaStream_t st) {
  if (kde_jsd_fused_read(record, out_host) == 1) {
    if (method_used_host) *method_used_host = UQ_KDE_MOMENTS;
    return UQ_OK;
  }
  return kde_jsd(u, nu, v, nv, grid_pts, UQ_KDE_AUTO, out_host, method_used_host, ws, ws_bytes, st);
}

// method: UQ_KDE_AUTO / UQ_KDE_WINDOW (sorted samples, 9-sigma windows) / UQ_KDE_MOMENTS.
// method_used_host (may be NULL) receives the method that ran.
int kde_jsd(const float* u, int64_t nu, const float* v, int64_t nv, int grid_pts, int method,
            double* out_host, int* method_used_host, void* ws, size_t ws_bytes, cudaStream_t st) {
  const WsLayout L = layout(nu, nv, grid_pts);
  UQ_REQUIRE(ws != nullptr && ws_bytes >= L.total, UQ_ERR_WORKSPACE,
             "kde_jsd needs %zu workspace bytes, got %zu", L.total, ws_bytes);
  UQ_REQUIRE(method >= UQ_KDE_AUTO && method <= UQ_KDE_MOMENTS, UQ_ERR_INVALID,
             "kde_jsd: unknown method %d", method);
  char* b = static_cast<char*>(ws);
  if (method_used_host) *method_used_host = UQ_KDE_WINDOW;
  if (method != UQ_KDE_WINDOW && nu >= 2 && nv >= 2) {
    // the single-launch moment method (kde_fused.cu); it declines (status 2) when the range
    // exceeds ~4000 bandwidths, a bandwidth is zero or a value is not finite
    UQ_REQUIRE(L.fused_bytes > 0 || method != UQ_KDE_MOMENTS, UQ_ERR_UNSUPPORTED,
               "kde_jsd: this device cannot run the moment method's cooperative launch");
    if (L.fused_bytes > 0) {
      int status = 0;
      double info[6];
      const int rc = kde_jsd_fused(u, nu, v, nv, grid_pts, out_host, &status, info, b + L.fused,
                                   L.fused_bytes, st);
      if (rc != UQ_OK) return rc;
      if (status == 1) {
        if (method_used_host) *method_used_host = UQ_KDE_MOMENTS;
        return UQ_OK;
      }
      UQ_REQUIRE(method != UQ_KDE_MOMENTS, UQ_ERR_UNSUPPORTED,
                 "kde_jsd: the moment method needs (max - min) <= 4000 bandwidths, h > 0 and "
                 "finite values (range %g, bandwidths %g / %g)", info[3] - info[2], info[0],
                 info[1]);
    }
  }
  float* du = reinterpret_cast<float*>(b + L.u);
  float* dut = reinterpret_cast<float*>(b + L.ut);
  float* dv = reinterpret_cast<float*>(b + L.v);
  float* dvt = reinterpret_cast<float*>(b + L.vt);
  double* partials = reinterpret_cast<double*>(b + L.partials);
  KdeParams* params = reinterpret_cast<KdeParams*>(b + L.params);
  double* pdf = reinterpret_cast<double*>(b + L.pdf);
  double* result = reinterpret_cast<double*>(b + L.result);
  float *su = nullptr, *sv = nullptr;
  int rc = radix_sort_f32_copy(u, du, dut, nu, b + L.scratch, radix_sort_scratch_bytes(nu), &su, st);
  if (rc != UQ_OK) return rc;
  rc = radix_sort_f32_copy(v, dv, dvt, nv, b + L.scratch, radix_sort_scratch_bytes(nv), &sv, st);
  if (rc != UQ_OK) return rc;
  stats_kernel<<<STAT_BLOCKS, 256, 0, st>>>(su, nu, partials);
  UQ_LAUNCH_CHECK();
  stats_kernel<<<STAT_BLOCKS, 256, 0, st>>>(sv, nv, partials + 2 * STAT_BLOCKS);
  UQ_LAUNCH_CHECK();
  setup_kernel<<<1, 32, 0, st>>>(su, nu, sv, nv, partials, partials + 2 * STAT_BLOCKS, grid_pts,
                                 params);
  UQ_LAUNCH_CHECK();
  UQ_CUDA(cudaMemsetAsync(pdf, 0, sizeof(double) * 2 * (size_t)grid_pts, st));
  const int64_t chunks_u = (nu + KDE_CHUNK - 1) / KDE_CHUNK;
  const int64_t chunks_v = (nv + KDE_CHUNK - 1) / KDE_CHUNK;
  kde_eval_kernel<<<(unsigned)(chunks_u + chunks_v), KDE_THREADS, 0, st>>>(
      su, nu, sv, nv, chunks_u, grid_pts, params, pdf);
  UQ_LAUNCH_CHECK();
  rc = jsd_launch(pdf, grid_pts, reinterpret_cast<double*>(b + L.jsd), result, st);
  if (rc != UQ_OK) return rc;
  UQ_CUDA(cudaMemcpyAsync(out_host, result, sizeof(double), cudaMemcpyDeviceToHost, st));
  UQ_CUDA(cudaStreamSynchronize(st));
  return UQ_OK;
}

// ---- sharded KDE-JS (one GPU's part of each sample) ---------------------------------------------
size_t kde_grid_workspace_bytes(int64_t n) {
  if (n < 1) return 0;
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  return 2 * al(sizeof(float) * (size_t)n) + al(radix_sort_scratch_bytes(n)) + al(sizeof(KdeParams)) +
         al(sizeof(double) * KM_MAX_BINS * KM_WORDS);
}

// Adds this shard's Gaussian kernel sums to `grid` (float64 [grid_pts], device).  The grid
// (lo, hi) and the bandwidth come from the caller: they are properties of the WHOLE sample
// (global min / max / Scott factor), reduced across ranks before this call.
int kde_grid_accumulate(const float* x, int64_t n, double lo, double hi, double bandwidth,
                        int grid_pts, double* grid, void* ws, size_t ws_bytes, cudaStream_t st) {
  const size_t need = kde_grid_workspace_bytes(n);
  UQ_REQUIRE(ws != nullptr && ws_bytes >= need, UQ_ERR_WORKSPACE,
             "kde grid accumulate needs %zu workspace bytes, got %zu", need, ws_bytes);
  UQ_REQUIRE(bandwidth > 0.0 && hi >= lo, UQ_ERR_INVALID,
             "kde grid accumulate: bandwidth %g, range [%g, %g]", bandwidth, lo, hi);
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  char* b = static_cast<char*>(ws);
  float* dx = reinterpret_cast<float*>(b);
  float* dt = reinterpret_cast<float*>(b + al(sizeof(float) * (size_t)n));
  char* scratch = b + 2 * al(sizeof(float) * (size_t)n);
  KdeParams* params = reinterpret_cast<KdeParams*>(scratch + al(radix_sort_scratch_bytes(n)));
  if (const int nb = getenv("UQ_KDE_WINDOW_ONLY") ? 0 : km_bins(lo, hi, bandwidth)) {
    // moment method: one pass over the shard, then a [grid x 77 bins] evaluation (no sort)
    double* tables = reinterpret_cast<double*>(reinterpret_cast<char*>(params) + al(sizeof(KdeParams)));
    return km_accumulate(x, n, lo, hi, bandwidth, nb, grid_pts, grid, tables, st);
  }
  float* sx = nullptr;
  int rc = radix_sort_f32_copy(x, dx, dt, n, scratch, radix_sort_scratch_bytes(n), &sx, st);
  if (rc != UQ_OK) return rc;
  KdeParams hp;
  hp.lo = lo;
  hp.step = (hi - lo) / (double)(grid_pts - 1);
  hp.h[0] = hp.h[1] = bandwidth;
  hp.scale[0] = hp.scale[1] = sqrt(0.5 * 1.4426950408889634) / bandwidth;
  // pageable source: the copy is staged by the driver before the call returns
  UQ_CUDA(cudaMemcpyAsync(params, &hp, sizeof(hp), cudaMemcpyHostToDevice, st));
  const int64_t chunks = (n + KDE_CHUNK - 1) / KDE_CHUNK;
  kde_eval_kernel<<<(unsigned)chunks, KDE_THREADS, 0, st>>>(sx, n, sx, 0, chunks, grid_pts, params,
                                                           grid);
  UQ_LAUNCH_CHECK();
  return UQ_OK;
}

// Jensen-Shannon distance of two raw kernel-sum vectors ([2][grid_pts] float64, device)
int jsd_from_grids(const double* grids, int grid_pts, double* out_host, cudaStream_t st) {
  char* buf = nullptr;
  UQ_CUDA(cudaMallocAsync((void**)&buf, JSD_SCRATCH_BYTES + 256, st));
  double* result = reinterpret_cast<double*>(buf);
  double* scratch = reinterpret_cast<double*>(buf + 256);
  int rc = jsd_launch(grids, grid_pts, scratch, result, st);
  cudaError_t e = cudaSuccess;
  if (rc == UQ_OK) e = cudaMemcpyAsync(out_host, result, sizeof(double), cudaMemcpyDeviceToHost, st);
  cudaFreeAsync(buf, st);
  if (rc != UQ_OK) return rc;
  if (e != cudaSuccess) return cuda_fail(e, "jsd_from_grids", __FILE__, __LINE__);
  UQ_CUDA(cudaStreamSynchronize(st));
  return UQ_OK;
}

}  // namespace uq
