// 1-D Wasserstein distance between two unweighted samples:  W1 = integral |F_u - F_v| dx.
//
// Replaces scipy.stats.wasserstein_distance as called by
// WassersteinEvaluation._evaluate_uncertainties (nnueehcs/evaluation.py:175-188); the arithmetic
// follows scipy's _cdf_distance(p=1): the float32 scores are widened to float64 first
// (_validate_distribution), then successive differences of the merged sorted values times
// |cdf_u - cdf_v| with cdf = (#values <= x)/size, all in float64.
//
// Two methods (uq_wasserstein_1d_ex's `method`; uq_wasserstein_1d = automatic choice):
//
//  * binned (default): ONE pass over each sample at HBM speed accumulates, per order-preserving
//    key bin (top 14 key bits: sign, exponent, 5 mantissa bits -> 16384 bins), the count c_b and
//    the integer sum K_b of the low 18 key bits.  Inside a bin every float32 is t_b + k ulp_b, so
//    integral_bin F_u = ((C_b + c_b) w_b - K_b ulp_b) / n exactly (C_b = count below the bin,
//    w_b = bin width).  Where D = F_u - F_v provably keeps one sign across the bin
//    (C_u n_v >= (C_v + c_v) n_u  or  (C_u + c_u) n_v <= C_v n_u, in integers), the bin contributes
//    |integral D| and its values are never looked at again.  The values of the remaining
//    ("ambiguous") bins are compacted, radix-sorted and integrated exactly by a merge-path kernel
//    that clips every step at the bin edges.  ID-vs-OOD score distributions resolve almost every
//    bin in the first pass, so the whole metric costs one read of the inputs; if more than half of
//    the values are ambiguous (e.g. the two samples come from the same distribution) the sort
//    method below is used instead.
//
//  * sort: two radix sorts (sort.cu), then one merge-path kernel -- each block binary-searches
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


// ---- binned method -----------------------------------------------------------------------------

constexpr int WB_BINS = 16384;      // == KEY_BINS of shard_metrics.cu (top 14 key bits)
constexpr int WB_LOW_BITS = 18;
constexpr uint32_t WB_LOW_MASK = (1u << WB_LOW_BITS) - 1u;
constexpr int BM_THREADS = 1024;
constexpr int64_t BM_MAX_PER_BLOCK = (int64_t)1 << 22;  // 9-bit partial sums stay below 2^32

__device__ __forceinline__ uint32_t wb_key(float x) {
  const uint32_t b = __float_as_uint(x + 0.0f);  // -0.0 -> +0.0: one bin for the value zero
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

// lower edge of key bin b (b in [0, WB_BINS]) as float64; the inf/NaN bins decode as if the
// exponent range went on, so every bin has a finite width
__device__ __forceinline__ double wb_edge(int b, double* ulp) {
  const uint32_t key_hi = (uint32_t)b;                       // key >> 18, 15 bits for b = WB_BINS
  const bool pos = key_hi >= (1u << 13);
  const uint32_t bits = pos ? ((key_hi - (1u << 13)) << WB_LOW_BITS)
                            : (~(key_hi << WB_LOW_BITS)) & 0x7FFFFFFFu;
  const int e = (int)(bits >> 23);
  const double m = (double)(bits & 0x7FFFFFu);
  const double mag = e == 0 ? scalbn(m, -149) : scalbn(m + 8388608.0, e - 150);
  if (ulp) *ulp = scalbn(1.0, (e > 1 ? e : 1) - 150);
  return pos ? mag : -mag;
}

// One pass over a sample: block-private shared-memory tables (count, sum of the low 9 and of the
// high 9 of the 18 low key bits -- 32-bit shared atomics are native, 64-bit ones are a CAS loop),
// flushed with one 64-bit reduction per non-empty bin.
__global__ void __launch_bounds__(BM_THREADS, 1)
bin_moments_kernel(const float* __restrict__ x, int64_t n, unsigned long long* __restrict__ cnt,
                   unsigned long long* __restrict__ ksum) {
  extern __shared__ uint32_t bm_sh[];
  uint32_t* c = bm_sh;
  uint32_t* lo = bm_sh + WB_BINS;
  uint32_t* hi = bm_sh + 2 * WB_BINS;
  for (int i = threadIdx.x; i < 3 * WB_BINS; i += BM_THREADS) bm_sh[i] = 0;
  __syncthreads();
  auto add = [&](float v) {
    const uint32_t k = wb_key(v);
    const uint32_t b = k >> WB_LOW_BITS, low = k & WB_LOW_MASK;
    atomicAdd(&c[b], 1u);
    atomicAdd(&lo[b], low & 511u);
    atomicAdd(&hi[b], low >> 9);
  };
  // scalar head up to 16-byte alignment, float4 body (4 loads in flight per thread), scalar tail
  const int64_t head = min(n, (int64_t)((16 - ((uintptr_t)x & 15)) & 15) / 4);
  const int64_t n4 = (n - head) / 4;
  const float4* x4 = reinterpret_cast<const float4*>(x + head);
  const int64_t gtid = (int64_t)blockIdx.x * BM_THREADS + threadIdx.x;
  const int64_t gstride = (int64_t)gridDim.x * BM_THREADS;
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
  __syncthreads();
  for (int b = threadIdx.x; b < WB_BINS; b += BM_THREADS) {
    const uint32_t cb = c[b];
    if (cb) {
      atomicAdd(&cnt[b], (unsigned long long)cb);
      atomicAdd(&ksum[b], (unsigned long long)lo[b] + ((unsigned long long)hi[b] << 9));
    }
  }
}

struct BinnedResult {
  double resolved;     // sum of the sign-definite bins' contributions
  long long amb_u, amb_v;  // values of each sample in ambiguous bins
  long long nonfinite;     // values in the inf / NaN bins (the caller falls back to the sort method)
};

// Single block: prefix counts, per-bin contribution or ambiguity flag, rank offsets of the exact
// pass (skip_x[b] = values of x below bin b that sit in resolved bins), bin edges.
__global__ void __launch_bounds__(1024, 1)
bin_resolve_kernel(const unsigned long long* __restrict__ cnt_u,
                   const unsigned long long* __restrict__ ks_u,
                   const unsigned long long* __restrict__ cnt_v,
                   const unsigned long long* __restrict__ ks_v, long long nu, long long nv,
                   uint8_t* __restrict__ flags, long long* __restrict__ skip_u,
                   long long* __restrict__ skip_v, double* __restrict__ edges,
                   BinnedResult* __restrict__ res) {
  constexpr int PER = WB_BINS / 1024;  // consecutive bins per thread
  __shared__ long long sc[4][1024];
  __shared__ double sd[1024];
  const int t = threadIdx.x;
  const int b0 = t * PER;
  long long tot_u = 0, tot_v = 0;
  for (int q = 0; q < PER; ++q) tot_u += (long long)cnt_u[b0 + q], tot_v += (long long)cnt_v[b0 + q];
  sc[0][t] = tot_u, sc[1][t] = tot_v;
  __syncthreads();
  // exclusive scan over the 1024 thread totals (Hillis-Steele, two arrays at once)
  auto scan2 = [&](int a, int b) {
    for (int o = 1; o < 1024; o <<= 1) {
      long long xa = 0, xb = 0;
      if (t >= o) xa = sc[a][t - o], xb = sc[b][t - o];
      __syncthreads();
      sc[a][t] += xa, sc[b][t] += xb;
      __syncthreads();
    }
  };
  scan2(0, 1);
  long long Cu = sc[0][t] - tot_u, Cv = sc[1][t] - tot_v;
  const double dnu = (double)nu, dnv = (double)nv;
  double acc = 0.0;
  long long au = 0, av = 0, nonfin = 0;
  uint8_t fl[PER];
  for (int q = 0; q < PER; ++q) {
    const int b = b0 + q;
    const long long cu = (long long)cnt_u[b], cv = (long long)cnt_v[b];
    double ulp;
    const double tb = wb_edge(b, &ulp);
    const double w = wb_edge(b + 1, nullptr) - tb;
    edges[b] = tb;
    if (b == WB_BINS - 1) edges[WB_BINS] = tb + w;
    if ((b < 32 || b >= WB_BINS - 32) && (cu | cv)) nonfin += cu + cv;
    const bool d_ge0 = Cu * nv - (Cv + cv) * nu >= 0;  // D >= 0 on the whole bin (|.| < 2^62)
    const bool d_le0 = (Cu + cu) * nv - Cv * nu <= 0;  // D <= 0 on the whole bin
    fl[q] = 0;
    if (d_ge0 || d_le0) {
      if ((cu | cv) || Cu * nv != Cv * nu) {
        const double a = ((double)(Cu + cu) * w - (double)ks_u[b] * ulp) / dnu;
        const double c = ((double)(Cv + cv) * w - (double)ks_v[b] * ulp) / dnv;
        acc += fabs(a - c);
      }
    } else {
      fl[q] = 1;
      au += cu, av += cv;
    }
    flags[b] = fl[q];
    Cu += cu, Cv += cv;
  }
  __syncthreads();
  sc[0][t] = au, sc[1][t] = av, sc[2][t] = nonfin, sd[t] = acc;
  __syncthreads();
  scan2(0, 1);
  // skip tables: values below bin b in resolved bins = C_b - (ambiguous values below b)
  long long ambu = sc[0][t] - au, ambv = sc[1][t] - av;
  Cu -= tot_u, Cv -= tot_v;  // back to the count below bin b0
  for (int q = 0; q < PER; ++q) {
    const int b = b0 + q;
    skip_u[b] = Cu - ambu, skip_v[b] = Cv - ambv;
    const long long cu = (long long)cnt_u[b], cv = (long long)cnt_v[b];
    if (fl[q]) ambu += cu, ambv += cv;
    Cu += cu, Cv += cv;
  }
  if (t == 1023) skip_u[WB_BINS] = Cu - ambu, skip_v[WB_BINS] = Cv - ambv;
  // fixed-order reductions
  for (int o = 512; o > 0; o >>= 1) {
    __syncthreads();
    if (t < o) sd[t] += sd[t + o], sc[2][t] += sc[2][t + o];
  }
  if (t == 0) {
    res->resolved = sd[0];
    res->amb_u = sc[0][1023];
    res->amb_v = sc[1][1023];
    res->nonfinite = sc[2][0];
  }
}

// values whose bin is flagged -> out (order arbitrary: they are sorted next)
__global__ void __launch_bounds__(256)
compact_flagged_kernel(const float* __restrict__ x, int64_t n, const uint8_t* __restrict__ flags,
                       float* __restrict__ out, unsigned long long* __restrict__ cursor) {
  constexpr int ITEMS = 16;
  __shared__ uint32_t cnt;
  __shared__ unsigned long long base;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const int64_t tile0 = (int64_t)blockIdx.x * 256 * ITEMS;
  float v[ITEMS];
  uint32_t slot[ITEMS];
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const int64_t i = tile0 + (int64_t)r * 256 + threadIdx.x;
    slot[r] = 0xFFFFFFFFu;
    if (i < n) {
      v[r] = __ldg(x + i);
      if (__ldg(flags + (wb_key(v[r]) >> WB_LOW_BITS))) slot[r] = atomicAdd(&cnt, 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && cnt) base = atomicAdd(cursor, (unsigned long long)cnt);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < ITEMS; ++r)
    if (slot[r] != 0xFFFFFFFFu) out[base + slot[r]] = v[r];
}

// Exact integral of |F_u - F_v| over the ambiguous bins.  U / V hold only the ambiguous values
// (sorted); the rank of a value in its whole sample is its index here plus skip_x[its bin].
// Every merged element contributes [cur, min(next, upper edge of cur's bin)) and, when it is the
// first of its bin, [lower edge, cur) as well, so nothing outside ambiguous bins is counted.
__global__ void __launch_bounds__(MRG_THREADS)
cdf_integral_binned_kernel(const float* __restrict__ U, int64_t nu, const float* __restrict__ V,
                           int64_t nv, const long long* __restrict__ skip_u,
                           const long long* __restrict__ skip_v, const double* __restrict__ edges,
                           int64_t nu_total, int64_t nv_total, const int64_t* __restrict__ splits,
                           double* __restrict__ block_partials) {
  __shared__ float su[MRG_TILE + 1];
  __shared__ float sv[MRG_TILE + 1];
  __shared__ double warp_part[MRG_THREADS / 32];
  const int64_t total = nu + nv;
  const int64_t k0 = (int64_t)blockIdx.x * MRG_TILE;  // elements [k0, k1) belong to this block
  const int64_t k1 = (k0 + MRG_TILE) < total ? (k0 + MRG_TILE) : total;
  const int t = threadIdx.x;
  const int64_t i0 = splits[blockIdx.x], j0 = k0 - i0;
  const int len = (int)(k1 - k0);
  const int lu = (int)((nu - i0) < (int64_t)(len + 1) ? (nu - i0) : (int64_t)(len + 1));
  const int lv = (int)((nv - j0) < (int64_t)(len + 1) ? (nv - j0) : (int64_t)(len + 1));
  for (int i = t; i < lu; i += MRG_THREADS) su[i] = U[i0 + i];
  for (int i = t; i < lv; i += MRG_THREADS) sv[i] = V[j0 + i];
  __syncthreads();

  const double dnu = (double)nu_total, dnv = (double)nv_total;
  double acc = 0.0;
  const int ka = t * MRG_ITEMS;
  if (ka < len) {
    const int kb = (ka + MRG_ITEMS) < len ? (ka + MRG_ITEMS) : len;
    int i = merge_path_smem(su, lu, sv, lv, ka);
    int j = ka - i;
    int pb = -1;  // bin of merged element k0 + ka - 1
    {
      const bool hu = i0 + i > 0, hv = j0 + j > 0;
      float pu = 0.f, pv = 0.f;
      if (hu) pu = i > 0 ? su[i - 1] : U[i0 - 1];
      if (hv) pv = j > 0 ? sv[j - 1] : V[j0 - 1];
      if (hu || hv) pb = (int)(wb_key(hu && hv ? fmaxf(pu, pv) : (hu ? pu : pv)) >> WB_LOW_BITS);
    }
    for (int k = ka; k < kb; ++k) {
      const bool take_u = (j >= lv) || (i < lu && su[i] <= sv[j]);
      const float cur = take_u ? su[i] : sv[j];
      const int b = (int)(wb_key(cur) >> WB_LOW_BITS);
      const long long ou = __ldg(skip_u + b) + i0, ov = __ldg(skip_v + b) + j0;
      if (b != pb) {
        const double d = (double)(ou + i) / dnu - (double)(ov + j) / dnv;
        acc += fabs(d) * ((double)cur - __ldg(edges + b));
      }
      if (take_u) ++i; else ++j;
      double seg;
      bool same = false;
      float nxt = 0.f;
      if (k0 + k + 1 < total) {
        nxt = ((j >= lv) || (i < lu && su[i] <= sv[j])) ? su[i] : sv[j];
        same = (int)(wb_key(nxt) >> WB_LOW_BITS) == b;
      }
      seg = same ? (double)nxt - (double)cur : __ldg(edges + b + 1) - (double)cur;
      const double d = (double)(ou + i) / dnu - (double)(ov + j) / dnv;
      acc += fabs(d) * seg;
      pb = b;
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

struct WsLayout {
  size_t u, ut, v, vt, scratch, parts, splits, result, tables, flags, skips, edges, bres, cursors, total;
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
  L.tables = o; o += al(sizeof(unsigned long long) * 4 * WB_BINS);  // cnt_u, ks_u, cnt_v, ks_v
  L.flags = o; o += al(WB_BINS);
  L.skips = o; o += al(sizeof(long long) * 2 * (WB_BINS + 1));
  L.edges = o; o += al(sizeof(double) * (WB_BINS + 1));
  L.bres = o; o += 256;
  L.cursors = o; o += 256;
  L.total = o;
  return L;
}

}  // namespace

size_t wasserstein_workspace_bytes(int64_t nu, int64_t nv) {
  if (nu < 1 || nv < 1) return 0;
  return layout(nu, nv).total;
}

namespace {

int bin_moments_launch(const float* x, int64_t n, unsigned long long* cnt, unsigned long long* ks,
                       cudaStream_t st) {
  static bool attr_set = false;
  constexpr int SMEM = 3 * WB_BINS * (int)sizeof(uint32_t);
  if (!attr_set) {
    UQ_CUDA(cudaFuncSetAttribute(bin_moments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 SMEM));
    attr_set = true;
  }
  // one block per SM; more only to keep a block's 9-bit partial sums below 2^32
  int64_t blocks = (n + (int64_t)BM_THREADS * 16 - 1) / ((int64_t)BM_THREADS * 16);
  if (blocks > 148) blocks = 148;
  const int64_t need = (n + BM_MAX_PER_BLOCK - 1) / BM_MAX_PER_BLOCK;
  if (blocks < need) blocks = (need + 147) / 148 * 148;
  if (blocks < 1) blocks = 1;
  bin_moments_kernel<<<(unsigned)blocks, BM_THREADS, SMEM, st>>>(x, n, cnt, ks);
  UQ_LAUNCH_CHECK();
  return UQ_OK;
}

// The sort method on device buffers that already hold copies of the samples.
int sort_and_integrate(float* du, float* dut, int64_t nu, float* dv, float* dvt, int64_t nv,
                       char* b, const WsLayout& L, double* result, cudaStream_t st) {
  double* parts = reinterpret_cast<double*>(b + L.parts);
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
  return UQ_OK;
}

}  // namespace

// method: UQ_WASSERSTEIN_AUTO / _SORT / _BINNED (binned even when most values are ambiguous).
// info_host (may be NULL): {method used, ambiguous u values, ambiguous v values}.
int wasserstein_1d(const float* u, int64_t nu, const float* v, int64_t nv, int method,
                   double* out_host, int64_t* info_host, void* ws, size_t ws_bytes,
                   cudaStream_t st) {
  const WsLayout L = layout(nu, nv);
  UQ_REQUIRE(ws != nullptr && ws_bytes >= L.total, UQ_ERR_WORKSPACE,
             "wasserstein needs %zu workspace bytes, got %zu", L.total, ws_bytes);
  UQ_REQUIRE(nu + nv < ((int64_t)1 << 31), UQ_ERR_INVALID, "wasserstein: too many values");
  UQ_REQUIRE(method >= UQ_WASSERSTEIN_AUTO && method <= UQ_WASSERSTEIN_BINNED, UQ_ERR_INVALID,
             "wasserstein: unknown method %d", method);
  char* b = static_cast<char*>(ws);
  float* du = reinterpret_cast<float*>(b + L.u);
  float* dut = reinterpret_cast<float*>(b + L.ut);
  float* dv = reinterpret_cast<float*>(b + L.v);
  float* dvt = reinterpret_cast<float*>(b + L.vt);
  double* result = reinterpret_cast<double*>(b + L.result);
  if (info_host) info_host[0] = UQ_WASSERSTEIN_SORT, info_host[1] = nu, info_host[2] = nv;

  if (method != UQ_WASSERSTEIN_SORT) {
    unsigned long long* tables = reinterpret_cast<unsigned long long*>(b + L.tables);
    unsigned long long *cnt_u = tables, *ks_u = tables + WB_BINS, *cnt_v = tables + 2 * WB_BINS,
                       *ks_v = tables + 3 * WB_BINS;
    uint8_t* flags = reinterpret_cast<uint8_t*>(b + L.flags);
    long long* skip_u = reinterpret_cast<long long*>(b + L.skips);
    long long* skip_v = skip_u + WB_BINS + 1;
    double* edges = reinterpret_cast<double*>(b + L.edges);
    BinnedResult* bres = reinterpret_cast<BinnedResult*>(b + L.bres);
    UQ_CUDA(cudaMemsetAsync(tables, 0, sizeof(unsigned long long) * 4 * WB_BINS, st));
    int rc = bin_moments_launch(u, nu, cnt_u, ks_u, st);
    if (rc != UQ_OK) return rc;
    rc = bin_moments_launch(v, nv, cnt_v, ks_v, st);
    if (rc != UQ_OK) return rc;
    bin_resolve_kernel<<<1, 1024, 0, st>>>(cnt_u, ks_u, cnt_v, ks_v, (long long)nu, (long long)nv,
                                           flags, skip_u, skip_v, edges, bres);
    UQ_LAUNCH_CHECK();
    BinnedResult h;
    UQ_CUDA(cudaMemcpyAsync(&h, bres, sizeof(h), cudaMemcpyDeviceToHost, st));
    UQ_CUDA(cudaStreamSynchronize(st));
    const int64_t amb = h.amb_u + h.amb_v;
    const bool use = h.nonfinite == 0 &&
                     (method == UQ_WASSERSTEIN_BINNED || amb <= (nu + nv) / 2);
    if (use) {
      if (info_host) info_host[0] = UQ_WASSERSTEIN_BINNED, info_host[1] = h.amb_u, info_host[2] = h.amb_v;
      if (amb == 0) {
        *out_host = h.resolved;
        return UQ_OK;
      }
      // exact pass over the ambiguous bins: compact -> sort -> clipped merge integral
      unsigned long long* cursors = reinterpret_cast<unsigned long long*>(b + L.cursors);
      UQ_CUDA(cudaMemsetAsync(cursors, 0, 2 * sizeof(unsigned long long), st));
      const int64_t tile = 256 * 16;
      compact_flagged_kernel<<<(unsigned)((nu + tile - 1) / tile), 256, 0, st>>>(u, nu, flags, du,
                                                                                 cursors);
      UQ_LAUNCH_CHECK();
      compact_flagged_kernel<<<(unsigned)((nv + tile - 1) / tile), 256, 0, st>>>(v, nv, flags, dv,
                                                                                 cursors + 1);
      UQ_LAUNCH_CHECK();
      float *su = du, *sv = dv;
      if (h.amb_u > 0) {
        rc = radix_sort_f32(du, dut, h.amb_u, b + L.scratch, radix_sort_scratch_bytes(nu), &su, st);
        if (rc != UQ_OK) return rc;
      }
      if (h.amb_v > 0) {
        rc = radix_sort_f32(dv, dvt, h.amb_v, b + L.scratch, radix_sort_scratch_bytes(nv), &sv, st);
        if (rc != UQ_OK) return rc;
      }
      const int64_t blocks = (amb + MRG_TILE - 1) / MRG_TILE;
      int64_t* splits = reinterpret_cast<int64_t*>(b + L.splits);
      double* parts = reinterpret_cast<double*>(b + L.parts);
      merge_partition_kernel<<<(unsigned)((blocks + 255) / 256), 256, 0, st>>>(
          su, h.amb_u, sv, h.amb_v, blocks, splits);
      UQ_LAUNCH_CHECK();
      cdf_integral_binned_kernel<<<(unsigned)blocks, MRG_THREADS, 0, st>>>(
          su, h.amb_u, sv, h.amb_v, skip_u, skip_v, edges, nu, nv, splits, parts);
      UQ_LAUNCH_CHECK();
      sum_partials_kernel<<<1, 1024, 0, st>>>(parts, blocks, result);
      UQ_LAUNCH_CHECK();
      double exact = 0.0;
      UQ_CUDA(cudaMemcpyAsync(&exact, result, sizeof(double), cudaMemcpyDeviceToHost, st));
      UQ_CUDA(cudaStreamSynchronize(st));
      *out_host = h.resolved + exact;
      return UQ_OK;
    }
  }

  UQ_CUDA(cudaMemcpyAsync(du, u, sizeof(float) * (size_t)nu, cudaMemcpyDeviceToDevice, st));
  UQ_CUDA(cudaMemcpyAsync(dv, v, sizeof(float) * (size_t)nv, cudaMemcpyDeviceToDevice, st));
  const int rc = sort_and_integrate(du, dut, nu, dv, dvt, nv, b, L, result, st);
  if (rc != UQ_OK) return rc;
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
