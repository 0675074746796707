// KDE Jensen-Shannon distance (JensenShannonEvaluation.pdf_jsd, nnueehcs/evaluation.py:268-276)
// as ONE cooperative launch: the moment method of kde_jsd.cu with every host round trip and
// launch boundary replaced by a grid barrier, and with half the shared atomics per value.
//
//   phase 1  (min, max, shifted float64 sums) of both samples      -> Scott bandwidths, grid, bins
//   phase 2  one pass over each sample: per FINE bin (width h/4 .. h/256) the count, sum eps and
//            sum eps^2 in block-private shared-memory tables (fixed point, native 32-bit ATOMS:
//            count and sum eps share one word, 2.3 atomics per value against the six of
//            kde_moments_kernel); every block stores its tables as a plain
//            coalesced slab -- no global atomics, so the result does not depend on block order
//   phase 3  fold: fine bins -> the coarse bins (width h/4) of kde_eval_bins: the slabs are added
//            up as integers, and the six coarse moments sum (delta + eps)^m follow from the
//            binomial shift of the three fine ones; the fine bins' third to fifth moments are
//            closed with the linear-density values E eps^3 = 3/5 a^2 E eps, E eps^4 = a^4 / 5,
//            E eps^5 = 3/7 a^4 E eps (a = the fine half width in bandwidths).  Against the exact
//            six-moment tables this moves the JS distance by <= 1e-9 relative at 2 M values per
//            sample even with fine = coarse bins (<= 1e-11 at h/16), and by <= 8e-7 at 2000 values
//            with fine = coarse (<= 4e-9 at h/16; small samples get h/256) -- numpy study in
//            DESIGN.md; the parity tests state 2e-5
//   phase 4  the [grid point x ~77 coarse bins] Hermite-series evaluation in float64
//   phase 5  scipy's jensenshannon on the two kernel-sum vectors; the last block writes the
//            distance into mapped host memory.
// All reductions are fixed-order trees, so the value is bit-reproducible for a given grid size.
// The kernel declines (status 2, nothing computed) when (max - min) exceeds ~4000 bandwidths of a
// sample, a bandwidth is not positive or a value is not finite; the caller then takes the older
// routes.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace uq {
namespace {

constexpr int KF_THREADS = 1024;
constexpr int KF_WARPS = KF_THREADS / 32;
constexpr int KF_COARSE_PER_H = 4;      // coarse bins are h/4 wide, as in kde_jsd.cu
// a coarse bin is cut into 2^k fine bins, k = 0..6 (h/4 .. h/256): the largest k whose bins fit
// shared memory, but k <= 2 for samples of 2^20 values and more (the closure error of the fold
// is statistical: large samples do not need finer bins and keep their slabs short)
constexpr int KF_MAX_REFINE = 6;
constexpr int KF_BIG_REFINE = 2;
constexpr int64_t KF_BIG_SAMPLE = (int64_t)1 << 20;
constexpr int KF_MAX_FINE = 16384;      // 16384 x 3 words x 4 B = 192 KB of shared memory
constexpr int KF_WORDS = 3;
constexpr int KF_TABLE_BYTES = KF_WORDS * KF_MAX_FINE * (int)sizeof(uint32_t);
constexpr int64_t KF_MAX_PER_BLOCK = (int64_t)1 << 19;  // values per block and slab (W2 <= 2^31)
constexpr int KF_CW = 6;                // coarse moments 0..5
constexpr int KF_MAX_COARSE = KF_MAX_FINE;   // k = 0: range up to 4096 bandwidths
constexpr int KF_MIN_GRID = 16;         // the fold stages <= 16384/16 + 64 fine bins per block
constexpr double KF_Z = 9.0;            // truncation in kernel standard deviations
constexpr int KF_LANES = 8;             // lanes per grid point in phase 4
// fixed point of the offset e (in FINE-BIN units, |e| <= 1/2) from the bin centre:
constexpr float KF_S1 = 262144.f;       // t1 = e 2^18 + 2^17 in [0, 2^18]
constexpr float KF_S2 = 16384.f;        // t2 = e^2 2^14 in [0, 2^12]

struct KfCtl {
  unsigned int barrier, ticket;
};

struct KfRecord {   // mapped host memory
  double jsd;
  double h[2], lo, hi;
  int status;       // 1 = computed, 2 = declined
  int nbf[2];
  int pad;
  unsigned long long t_ns[6];   // block 0's globaltimer at the start and after phases 1, 2, 3, 4, 5
};

struct KfParams {
  double lo, hi, h[2];
  float lo_f, inv_wf[2];
  double origin[2], wf[2];   // fine bin f of sample s = origin + [f, f + 1) wf  (origin = lo - wf)
  int nbf[2], nbc[2];
  int refine[2];    // k: fine bins are h / (4 * 2^k) wide
  int ok;
};

struct KfWs {       // device pointers into the workspace
  double* partials;          // [grid][2][4]
  uint32_t* slabs;           // [slabs of u | slabs of v][KF_WORDS * KF_MAX_FINE]
  double* coarse;            // [2][KF_MAX_COARSE][KF_CW]
  double* pdf;               // [2][grid_pts]
  double* blocksums;         // [grid][2]
  double* terms;             // [grid][2]
  KfCtl* ctl;                // (zeroed)
  KfRecord* record;
};

__device__ __forceinline__ unsigned long long kf_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ uint32_t kf_fixed(float term) {   // 0 <= term < 2^23, round to nearest
  return __float_as_uint(term + 8388608.0f) & 0x7FFFFFu;
}

// block-wide sums of two doubles in a fixed order; every thread gets the totals
__device__ __forceinline__ void kf_block_sum2(double& a, double& b, double (*sh)[KF_WARPS]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  __syncthreads();
  if (lane == 0) sh[0][w] = a, sh[1][w] = b;
  __syncthreads();
  a = sh[0][lane], b = sh[1][lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
}

// phase 1 for one sample: this block's (sum d, sum d^2, min, max), d = x - x[0]
__device__ __forceinline__ void kf_stats(const float* __restrict__ x, int64_t n,
                                         double* __restrict__ out4, double (*sh)[KF_WARPS]) {
  const float x0f = __ldg(x);
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
  for_each_value<KF_THREADS>(x, n, add);
  // NaN must survive the min / max (fminf drops it): fold it into the sums, which the
  // applicability test looks at
  kf_block_sum2(s1, s2, sh);
  double mn = (double)mnf, mx = (double)mxf;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __syncthreads();
  if (lane == 0) sh[0][w] = mn, sh[1][w] = mx;
  __syncthreads();
  mn = sh[0][lane], mx = sh[1][lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (threadIdx.x == 0) out4[0] = s1, out4[1] = s2, out4[2] = mn, out4[3] = mx;
  __syncthreads();
}

__device__ __forceinline__ uint32_t kf_atoms_add(uint32_t addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}
template <int OFF>
__device__ __forceinline__ void kf_reds_add(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0+%2], %1;" ::"r"(addr), "r"(v), "n"(OFF) : "memory");
}

// rare path of kf_add: did this add carry out of W2's low field / wrap W2?
__device__ __forceinline__ void kf_event2(uint32_t addr, uint32_t old, uint32_t inc) {
  const uint32_t nw = old + inc;
  const bool carry = (nw & 0x3FFFFFu) < (inc & 0x3FFFFFu), wrap = nw < old;
  if (carry | wrap)
    kf_reds_add<2 * KF_MAX_FINE * 4>(addr, (carry ? 1u << 8 : 0u) + (wrap ? 1u << 20 : 0u));
}

// phase 2, one value: two native shared atomics on the value's fine bin
//   W1 += t1,          t1 = e 2^18 + 2^17 in [0, 2^18]: wraps at most once per 2^14 adds;
//   W2 += 2^22 + t2,   t2 = e^2 2^14 in [0, 2^12]: low 22 bits = sum t2 mod 2^22, high 10 bits =
//                      (count + low-field carries) mod 2^10 -- count and second moment in ONE
//                      atomic; packing the count with the SMALL increment keeps the carry / wrap
//                      events rare (three adds in 1000: packed with t1 they fired in every second
//                      warp and cost more instructions than they saved);
//   SD += wrap of W1 | carry of W2 << 8 | wrap of W2 << 20, only from the add whose returned
//                      old word shows the event.
// The first version of this pass spent 45 instructions per value and was issue-bound (ncu: INT32
// pipe), so everything here is shaped to stay off the integer pipe:
//   * bin index = round(t + 1/2) by the 2^23 magic-number add -- no quarter-rate F2I / I2F; the
//     bin grid starts one fine bin below the minimum, so the index is never negative and the
//     rounding remainder e is in [-1/2, 1/2] by construction -- no clamps;
//   * the word address is one IMAD on the raw float bits (4 * (bits - 0x4B000000) mod 2^32);
//   * the increments are produced directly as bit patterns by FMAs with denormal results
//     (gradual underflow rounds to nearest at 2^-149): m 2^-149 has the bits m;
//   * a returned word is only tested for "a field is within one increment of full" (LOP3)
//     before the exact test.
__device__ __forceinline__ void kf_add(float v, float lo_f, float inv_wf, uint32_t w_rel) {
  const float a = fmaf(v - lo_f, inv_wf, 0.5f);           // t + 1/2, t = position in fine bins >= 0
  const float tf = a + 8388608.0f;                        // bits 0x4B000000 + round(t + 1/2)
  const float e = a - (tf - 8388608.0f);                  // offset from that bin's centre
  const uint32_t addr = __float_as_uint(tf) * 4u + w_rel; // w_rel = &W1[0] - 4 * 0x4B000000
  const uint32_t inc1 = __float_as_uint(fmaf(e, 0x1p-131f, 0x1p-132f));      // e 2^18 + 2^17
  const uint32_t inc2 = __float_as_uint(fmaf(e * e, 0x1p-135f, 0x1p-127f));  // 2^22 + e^2 2^14
  const uint32_t o1 = kf_atoms_add(addr, inc1);
  const uint32_t o2 = kf_atoms_add(addr + KF_MAX_FINE * 4, inc2);
  if ((~o1 & 0xFFFC0000u) == 0u)
    if (o1 + inc1 < o1) kf_reds_add<2 * KF_MAX_FINE * 4>(addr, 1u);
  if (((~o2 & 0x003FF000u) == 0u) | ((~o2 & 0xFF800000u) == 0u))
    kf_event2(addr, o2, inc2);
}

// this block's grid-stride share of x[0 .. n)
__device__ __forceinline__ void kf_accumulate(const float* __restrict__ x, int64_t n, float lo_f,
                                              float inv_wf, uint32_t w_rel) {
  for_each_value<KF_THREADS>(x, n, [&](float v) { kf_add(v, lo_f, inv_wf, w_rel); });
}

__global__ void __launch_bounds__(KF_THREADS, 1)
kde_jsd_fused_kernel(const float* __restrict__ u, int64_t nu, const float* __restrict__ v,
                     int64_t nv, int grid_pts, KfWs ws) {
  extern __shared__ __align__(16) uint32_t kf_sh[];
  __shared__ double red[2][KF_WARPS];
  __shared__ KfParams P;
  __shared__ double tot[2];
  __shared__ bool last;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int G = gridDim.x;
  const float* xs[2] = {u, v};
  const int64_t ns[2] = {nu, nv};

  const bool stamp = blockIdx.x == 0 && t == 0;
  if (stamp) ws.record->t_ns[0] = kf_now();
  const uint32_t w_rel = (uint32_t)__cvta_generic_to_shared(kf_sh) - 4u * 0x4B000000u;

  // ---- phase 1: statistics
  for (int s = 0; s < 2; ++s)
    kf_stats(xs[s], ns[s], ws.partials + ((size_t)blockIdx.x * 2 + s) * 4, red);
  grid_barrier(&ws.ctl->barrier, 1);
  if (warp == 0) {   // every block derives the same parameters from the same partials
    double st[2][4];
    for (int s = 0; s < 2; ++s) {
      double s1 = 0.0, s2 = 0.0, mn = INFINITY, mx = -INFINITY;
      for (int b = lane; b < G; b += 32) {
        const double* p = ws.partials + ((size_t)b * 2 + s) * 4;
        s1 += __ldcg(p), s2 += __ldcg(p + 1);
        mn = fmin(mn, __ldcg(p + 2)), mx = fmax(mx, __ldcg(p + 3));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      }
      st[s][0] = s1, st[s][1] = s2, st[s][2] = mn, st[s][3] = mx;
    }
    if (lane == 0) {
      KfParams p;
      p.ok = G >= KF_MIN_GRID;
      for (int s = 0; s < 2; ++s) {
        const double dn = (double)ns[s];
        const double m2 = st[s][1] - st[s][0] * st[s][0] / dn;   // sum (x - mean)^2
        // scipy.stats.gaussian_kde: h = sqrt(unbiased variance) * n^(-1/5)
        p.h[s] = sqrt(m2 / (dn - 1.0)) * pow(dn, -0.2);
      }
      p.lo = fmin(st[0][2], st[1][2]);
      p.hi = fmax(st[0][3], st[1][3]);
      // a NaN in the data poisons the sums (fmin / fmax drop it), an inf the range
      if (!(isfinite(p.lo) && isfinite(p.hi) && p.hi >= p.lo)) p.ok = 0;
      p.lo_f = (float)p.lo;   // exact: the minimum of float32 values
      for (int s = 0; s < 2; ++s) {
        if (!(p.h[s] > 0.0) || !isfinite(p.h[s])) { p.ok = 0; p.nbc[s] = p.nbf[s] = 0; continue; }
        const double w = p.h[s] / KF_COARSE_PER_H;
        const double nbc = floor((p.hi - p.lo) / w) + 3.0;   // + the fine bin below the minimum, + slack
        if (!(nbc <= (double)KF_MAX_COARSE)) { p.ok = 0; p.nbc[s] = p.nbf[s] = 0; continue; }
        p.nbc[s] = (int)nbc;
        const int kmax = ns[s] >= KF_BIG_SAMPLE ? KF_BIG_REFINE : KF_MAX_REFINE;
        int k = 0;
        while (k < kmax && (p.nbc[s] << (k + 1)) <= KF_MAX_FINE) ++k;
        p.refine[s] = k;
        p.nbf[s] = p.nbc[s] << k;
        p.inv_wf[s] = (float)((double)(KF_COARSE_PER_H << k) / p.h[s]);
        // the bins are those of the float32 factor the pass multiplies by
        p.wf[s] = 1.0 / (double)p.inv_wf[s];
        p.origin[s] = p.lo - p.wf[s];
      }
      P = p;
    }
  }
  __syncthreads();
  if (stamp) ws.record->t_ns[1] = kf_now();
  if (!P.ok) {
    if (blockIdx.x == 0 && t == 0) {
      ws.record->status = 2;
      ws.record->jsd = 0.0;
      ws.record->h[0] = P.h[0], ws.record->h[1] = P.h[1];
      ws.record->lo = P.lo, ws.record->hi = P.hi;
      ws.record->nbf[0] = ws.record->nbf[1] = 0;
      __threadfence_system();
    }
    return;   // the same decision in every block: nobody waits at a later barrier
  }

  // ---- phase 2: fine-bin tables; one plain slab per block, sample and round of G * 2^19 values
  {
    int slab_no = blockIdx.x;
    const int64_t round = (int64_t)G * KF_MAX_PER_BLOCK;   // a multiple of 4: alignment is kept
    for (int s = 0; s < 2; ++s) {
      const int nbf = P.nbf[s];
      for (int64_t r0 = 0; r0 < ns[s]; r0 += round, slab_no += G) {
        for (int w = 0; w < KF_WORDS; ++w)
          for (int i = t; i < nbf; i += KF_THREADS) kf_sh[w * KF_MAX_FINE + i] = 0;
        __syncthreads();
        kf_accumulate(xs[s] + r0, min(round, ns[s] - r0), P.lo_f, P.inv_wf[s], w_rel);
        __syncthreads();
        uint32_t* slab = ws.slabs + (size_t)slab_no * (KF_WORDS * KF_MAX_FINE);
        for (int w = 0; w < KF_WORDS; ++w)
          for (int i = t; i < nbf; i += KF_THREADS) slab[w * nbf + i] = kf_sh[w * KF_MAX_FINE + i];
        __syncthreads();
      }
    }
  }
  grid_barrier(&ws.ctl->barrier, 2);
  if (stamp) ws.record->t_ns[2] = kf_now();

  // ---- phase 3: fold the slabs into the coarse six-moment tables
  {
    unsigned long long* sums = reinterpret_cast<unsigned long long*>(kf_sh);  // [3][nfine]
    int slab_first = 0;
    for (int s = 0; s < 2; ++s) {
      const int nbf = P.nbf[s], nbc = P.nbc[s];
      const int fpc = 1 << P.refine[s];                    // fine bins per coarse bin
      const int fph = fpc * KF_COARSE_PER_H;               // fine bins per bandwidth
      const int chunk = (nbc + G - 1) / G;
      const int c0 = min(nbc, (int)blockIdx.x * chunk), c1 = min(nbc, c0 + chunk);
      const int nfine = (c1 - c0) * fpc;
      const int64_t round = (int64_t)G * KF_MAX_PER_BLOCK;
      const int nslabs = (int)((ns[s] + round - 1) / round) * G;
      const uint32_t* slab0 = ws.slabs + (size_t)slab_first * (KF_WORDS * KF_MAX_FINE);
      slab_first += nslabs;
      // thread = (fine bin, part): the parts of a bin take every PARTS-th slab and meet in shared
      // memory (one thread per bin walking all the slabs was a 148-deep chain of L2 round trips)
      for (int i = t; i < 5 * nfine; i += KF_THREADS) sums[i] = 0;
      __syncthreads();
      if (nfine > 0) {
        const int parts = max(1, min(nslabs, KF_THREADS / nfine));
        for (int it = t; it < nfine * parts; it += KF_THREADS) {
          const int part = it / nfine, fi = it - part * nfine;
          const uint32_t* p = slab0 + c0 * fpc + fi;
          unsigned long long t1 = 0, top = 0, low = 0, wraps1 = 0, carries2 = 0, wraps2 = 0;
#pragma unroll 4
          for (int b = part; b < nslabs; b += parts) {
            const uint32_t* q = p + (size_t)b * (KF_WORDS * KF_MAX_FINE);
            const uint32_t w1 = __ldcg(q), w2 = __ldcg(q + nbf), sd = __ldcg(q + 2 * nbf);
            t1 += w1, top += w2 >> 22, low += w2 & 0x3FFFFFu;
            wraps1 += sd & 0xFFu, carries2 += (sd >> 8) & 0xFFFu, wraps2 += sd >> 20;
          }
          // integer sums: the order of the shared atomics does not matter
          atomicAdd(&sums[fi], (wraps2 << 10) + top);              // count + carries2
          atomicAdd(&sums[nfine + fi], carries2);
          atomicAdd(&sums[2 * nfine + fi], (wraps1 << 32) + t1);   // sum t1
          atomicAdd(&sums[3 * nfine + fi], low);                   // sum t2 - carries2 2^22
        }
      }
      __syncthreads();
      // one warp per coarse bin, lanes over its fine bins, fixed-order butterfly
      for (int ci = warp; ci < c1 - c0; ci += KF_WARPS) {
        const double A = 0.5 / (double)fph;   // fine half width in bandwidths
        const double inv_s1 = 1.0 / ((double)KF_S1 * (double)fph);          // fine units -> h
        const double inv_s2 = 1.0 / ((double)KF_S2 * (double)fph * (double)fph);
        double M[KF_CW] = {0, 0, 0, 0, 0, 0};
        for (int i = lane; i < fpc; i += 32) {
          const int fi = ci * fpc + i;
          const unsigned long long carries = sums[nfine + fi];
          const double c = (double)(sums[fi] - carries);
          if (c == 0.0) continue;
          const double s1 = ((double)sums[2 * nfine + fi] - c * 131072.0) * inv_s1;
          const double s2 = (double)((carries << 22) + sums[3 * nfine + fi]) * inv_s2;
          const double S[KF_CW] = {c, s1, s2, 0.6 * A * A * s1, c * (A * A * A * A / 5.0),
                                   (3.0 / 7.0) * A * A * A * A * s1};
          const double d = ((double)i - 0.5 * (double)(fpc - 1)) / (double)fph;
          const double d2 = d * d, d3 = d2 * d, d4 = d2 * d2, d5 = d4 * d;
          M[0] += S[0];
          M[1] += d * S[0] + S[1];
          M[2] += d2 * S[0] + 2.0 * d * S[1] + S[2];
          M[3] += d3 * S[0] + 3.0 * d2 * S[1] + 3.0 * d * S[2] + S[3];
          M[4] += d4 * S[0] + 4.0 * d3 * S[1] + 6.0 * d2 * S[2] + 4.0 * d * S[3] + S[4];
          M[5] += d5 * S[0] + 5.0 * d4 * S[1] + 10.0 * d3 * S[2] + 10.0 * d2 * S[3] +
                  5.0 * d * S[4] + S[5];
        }
#pragma unroll
        for (int m = 0; m < KF_CW; ++m)
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) M[m] += __shfl_xor_sync(0xffffffffu, M[m], o);
        if (lane == 0) {
          double* T = ws.coarse + ((size_t)s * KF_MAX_COARSE + c0 + ci) * KF_CW;
#pragma unroll
          for (int m = 0; m < KF_CW; ++m) T[m] = M[m];
        }
      }
      __syncthreads();
    }
  }
  grid_barrier(&ws.ctl->barrier, 3);
  if (stamp) ws.record->t_ns[3] = kf_now();

  // ---- phase 4: kernel sums on the grid (the arithmetic of kde_eval_bins_kernel)
  {
    const double step = (P.hi - P.lo) / (double)(grid_pts - 1);
    const long long per_sample = (long long)grid_pts * KF_LANES;
    const long long total = 2 * per_sample;
    const long long stride = (long long)G * KF_THREADS;
    double psum[2] = {0.0, 0.0};
    for (long long base = 0; base < total; base += stride) {   // warp-uniform trip count
      const long long it = base + (long long)blockIdx.x * KF_THREADS + t;
      const bool on = it < total;
      const int s = on && it >= per_sample ? 1 : 0;
      const long long rem = it - (long long)s * per_sample;
      const int j = (int)(rem / KF_LANES), part = (int)(rem % KF_LANES);
      double acc = 0.0;
      if (on) {
        // the bins are those of the float32 scale factor the accumulate pass multiplied by
        const double h = P.h[s], w = (double)(1 << P.refine[s]) * P.wf[s], org = P.origin[s];
        const int nb = P.nbc[s];
        const double g = P.lo + (double)j * step;
        const double reach = KF_Z * h + 0.5 * w;   // (w = h / 4 up to float32 rounding)
        int b0 = (int)floor((g - reach - org) / w), b1 = (int)floor((g + reach - org) / w);
        b0 = b0 < 0 ? 0 : b0;
        b1 = b1 >= nb ? nb - 1 : b1;
        const double* tab = ws.coarse + (size_t)s * KF_MAX_COARSE * KF_CW;
        for (int b = b0 + part; b <= b1; b += KF_LANES) {
          const double* T = tab + (size_t)b * KF_CW;
          const double c = __ldcg(T);
          if (c == 0.0) continue;
          const double z = (g - (org + ((double)b + 0.5) * w)) / h;
          const double z2 = z * z;
          const double he2 = z2 - 1.0, he3 = z * (z2 - 3.0), he4 = z2 * (z2 - 6.0) + 3.0,
                       he5 = z * (z2 * (z2 - 10.0) + 15.0);
          const double series = c + z * __ldcg(T + 1) + he2 * __ldcg(T + 2) * (1.0 / 2.0) +
                                he3 * __ldcg(T + 3) * (1.0 / 6.0) +
                                he4 * __ldcg(T + 4) * (1.0 / 24.0) +
                                he5 * __ldcg(T + 5) * (1.0 / 120.0);
          acc += exp(-0.5 * z2) * series;
        }
      }
#pragma unroll
      for (int o = KF_LANES / 2; o > 0; o >>= 1)
        acc += __shfl_down_sync(0xffffffffu, acc, o, KF_LANES);
      if (on && part == 0) {
        ws.pdf[(size_t)s * grid_pts + j] = acc;
        psum[s] += acc;
      }
    }
    kf_block_sum2(psum[0], psum[1], red);
    if (t == 0) ws.blocksums[2 * blockIdx.x] = psum[0], ws.blocksums[2 * blockIdx.x + 1] = psum[1];
  }
  grid_barrier(&ws.ctl->barrier, 4);
  if (stamp) ws.record->t_ns[4] = kf_now();

  // ---- phase 5: scipy.spatial.distance.jensenshannon on the two raw vectors
  if (warp == 0) {
    double a = 0.0, b = 0.0;
    for (int k = lane; k < G; k += 32)
      a += __ldcg(ws.blocksums + 2 * k), b += __ldcg(ws.blocksums + 2 * k + 1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) tot[0] = a, tot[1] = b;
  }
  __syncthreads();
  {
    const double su = tot[0], sv = tot[1];
    double left = 0.0, right = 0.0;
    for (int j = blockIdx.x * KF_THREADS + t; j < grid_pts; j += G * KF_THREADS) {
      const double p = __ldcg(ws.pdf + j) / su, q = __ldcg(ws.pdf + grid_pts + j) / sv;
      const double m = (p + q) / 2.0;
      if (p > 0.0 && m > 0.0) left += p * log(p / m);
      if (q > 0.0 && m > 0.0) right += q * log(q / m);
    }
    kf_block_sum2(left, right, red);
    if (t == 0) {
      ws.terms[2 * blockIdx.x] = left, ws.terms[2 * blockIdx.x + 1] = right;
      __threadfence();
      last = atomicAdd(&ws.ctl->ticket, 1u) == (unsigned int)(G - 1);
    }
  }
  __syncthreads();
  if (last && warp == 0) {
    __threadfence();
    double l = 0.0, r = 0.0;
    for (int k = lane; k < G; k += 32)
      l += __ldcg(ws.terms + 2 * k), r += __ldcg(ws.terms + 2 * k + 1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l += __shfl_xor_sync(0xffffffffu, l, o);
      r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    if (lane == 0) {
      KfRecord* rec = ws.record;
      rec->jsd = sqrt((l + r) / 2.0);
      rec->h[0] = P.h[0], rec->h[1] = P.h[1], rec->lo = P.lo, rec->hi = P.hi;
      rec->nbf[0] = P.nbf[0], rec->nbf[1] = P.nbf[1];
      rec->t_ns[5] = kf_now();
      rec->status = 1;
      __threadfence_system();
    }
  }
}

struct KfLayout {
  size_t zeroed, ctl, zeroed_end, partials, slabs, coarse, pdf, blocksums, terms, total;
};

constexpr int KF_GRID_CAP = 1024;   // workspace is sized for at most this many co-resident blocks

KfLayout kf_layout(int64_t nu, int64_t nv, int grid_pts, int grid) {
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  KfLayout L;
  size_t o = 0;
  L.zeroed = o;
  L.ctl = o; o += 256;
  L.zeroed_end = o;
  L.partials = o; o += al(sizeof(double) * 8 * (size_t)grid);
  const int64_t round = (int64_t)grid * KF_MAX_PER_BLOCK;
  const size_t nslabs = (size_t)(((nu + round - 1) / round + (nv + round - 1) / round) * grid);
  L.slabs = o; o += al(sizeof(uint32_t) * nslabs * KF_WORDS * KF_MAX_FINE);
  L.coarse = o; o += al(sizeof(double) * 2 * KF_MAX_COARSE * KF_CW);
  L.pdf = o; o += al(sizeof(double) * 2 * (size_t)grid_pts);
  L.blocksums = o; o += al(sizeof(double) * 2 * (size_t)grid);
  L.terms = o; o += al(sizeof(double) * 2 * (size_t)grid);
  L.total = o;
  return L;
}

PerDeviceOnce kf_opted;
PerDeviceInt kf_grid_cache;
constexpr int KF_SMEM = KF_TABLE_BYTES;

int kf_grid(int* grid) {
  if (int rc = smem_opt_in(kde_jsd_fused_kernel, KF_SMEM, kf_opted)) return rc;
  if (int rc = coop_grid_limit(kde_jsd_fused_kernel, KF_THREADS, KF_SMEM, kf_grid_cache, grid))
    return rc;
  if (*grid > KF_GRID_CAP) *grid = KF_GRID_CAP;
  return UQ_OK;
}

}  // namespace

// Workspace for kde_jsd_fused on the current device (0 if the device cannot run it).
size_t kde_jsd_fused_workspace_bytes(int64_t nu, int64_t nv, int grid_pts) {
  int grid = 0;
  if (kf_grid(&grid) != UQ_OK) return 0;
  return kf_layout(nu, nv, grid_pts, grid).total;
}

// memset + cooperative launch; the last block writes the KfRecord to `record_dev`.  No
// synchronisation.
namespace {
int kde_jsd_fused_launch(const float* u, int64_t nu, const float* v, int64_t nv, int grid_pts,
                         KfRecord* record_dev, void* ws, size_t ws_bytes, cudaStream_t st) {
  int grid = 0;
  if (int rc = kf_grid(&grid)) return rc;
  const KfLayout L = kf_layout(nu, nv, grid_pts, grid);
  UQ_REQUIRE(ws != nullptr && ws_bytes >= L.total, UQ_ERR_WORKSPACE,
             "kde_jsd_fused needs %zu workspace bytes, got %zu", L.total, ws_bytes);
  UQ_REQUIRE(nu >= 2 && nv >= 2 && grid_pts >= 2, UQ_ERR_INVALID,
             "kde_jsd_fused: each sample needs at least 2 values and the grid 2 points");
  char* b = static_cast<char*>(ws);
  KfWs w;
  w.partials = reinterpret_cast<double*>(b + L.partials);
  w.slabs = reinterpret_cast<uint32_t*>(b + L.slabs);
  w.coarse = reinterpret_cast<double*>(b + L.coarse);
  w.pdf = reinterpret_cast<double*>(b + L.pdf);
  w.blocksums = reinterpret_cast<double*>(b + L.blocksums);
  w.terms = reinterpret_cast<double*>(b + L.terms);
  w.ctl = reinterpret_cast<KfCtl*>(b + L.ctl);
  w.record = record_dev;
  UQ_CUDA(cudaMemsetAsync(b + L.zeroed, 0, L.zeroed_end - L.zeroed, st));
  void* args[] = {(void*)&u, (void*)&nu, (void*)&v, (void*)&nv, (void*)&grid_pts, (void*)&w};
  UQ_CUDA(cudaLaunchCooperativeKernel((const void*)kde_jsd_fused_kernel, dim3(grid),
                                      dim3(KF_THREADS), args, KF_SMEM, st));
  UQ_LAUNCH_CHECK();
  return UQ_OK;
}
}  // namespace

// One memset + one cooperative launch + one stream synchronisation.  *status_host = 1: *out_host
// holds the distance; 2: the kernel declined (range / bandwidth / non-finite data) and nothing
// was computed.  info_host (may be NULL): {h_u, h_v, lo, hi, fine bins u, fine bins v}.
int kde_jsd_fused(const float* u, int64_t nu, const float* v, int64_t nv, int grid_pts,
                  double* out_host, int* status_host, double* info_host, void* ws, size_t ws_bytes,
                  cudaStream_t st) {
  void *slot_h = nullptr, *slot_d = nullptr;
  if (int rc = result_slot(&slot_h, &slot_d)) return rc;
  static_cast<KfRecord*>(slot_h)->status = 0;
  if (int rc = kde_jsd_fused_launch(u, nu, v, nv, grid_pts, static_cast<KfRecord*>(slot_d), ws,
                                    ws_bytes, st))
    return rc;
  UQ_CUDA(cudaStreamSynchronize(st));
  KfRecord rec;
  memcpy(&rec, slot_h, sizeof(rec));
  UQ_REQUIRE(rec.status == 1 || rec.status == 2, UQ_ERR_CUDA,
             "kde_jsd_fused: the kernel left no result record (status %d)", rec.status);
  *status_host = rec.status;
  *out_host = rec.jsd;
  if (info_host) {
    info_host[0] = rec.h[0], info_host[1] = rec.h[1], info_host[2] = rec.lo, info_host[3] = rec.hi;
    info_host[4] = (double)rec.nbf[0], info_host[5] = (double)rec.nbf[1];
  }
  return UQ_OK;
}

// enqueue-only form (see wasserstein_1d_enqueue): `record` = caller-owned mapped pinned host
// memory; kde_jsd_fused_read returns 1 (distance in *out_host), 2 (declined) or 0 (no record yet)
int kde_jsd_fused_enqueue(const float* u, int64_t nu, const float* v, int64_t nv, int grid_pts,
                          void* record, void* ws, size_t ws_bytes, cudaStream_t st) {
  void* record_dev = nullptr;
  if (!record || cudaHostGetDevicePointer(&record_dev, record, 0) != cudaSuccess) {
    (void)cudaGetLastError();   // the failed query must not show up in the next launch check
    set_error("kde_jsd enqueue: the record must be mapped pinned host memory");
    return UQ_ERR_INVALID;
  }
  static_assert(sizeof(KfRecord) <= UQ_METRIC_RECORD_BYTES, "record size");
  static_cast<KfRecord*>(record)->status = 0;
  return kde_jsd_fused_launch(u, nu, v, nv, grid_pts, static_cast<KfRecord*>(record_dev), ws,
                              ws_bytes, st);
}

int kde_jsd_fused_read(const void* record, double* out_host) {
  KfRecord rec;
  memcpy(&rec, record, sizeof(rec));
  *out_host = rec.jsd;
  return rec.status;
}

// diagnostics: microseconds block 0 spent in phases 1..5 of this thread's last kde_jsd_fused call
// on the current device
int kde_jsd_fused_phase_us(double* out5) {
  void *slot_h = nullptr, *slot_d = nullptr;
  if (int rc = result_slot(&slot_h, &slot_d)) return rc;
  KfRecord rec;
  memcpy(&rec, slot_h, sizeof(rec));
  for (int i = 0; i < 5; ++i)
    out5[i] = rec.t_ns[i + 1] >= rec.t_ns[i] ? (double)(rec.t_ns[i + 1] - rec.t_ns[i]) * 1e-3 : 0.0;
  return UQ_OK;
}

}  // namespace uq
