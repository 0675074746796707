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
#include <string.h>

#include "common.cuh"
#include "sort.cuh"

namespace uq {
namespace {

constexpr int MRG_THREADS = 256;
constexpr int MRG_ITEMS = 8;
constexpr int MRG_TILE = MRG_THREADS * MRG_ITEMS;  // merged positions per block

// a <= b in the order the radix sort leaves the arrays in: NaNs (canonicalised by the sort) after
// everything else, all NaNs equal.  A plain "a <= b" is false for any NaN, which would merge a
// sample's NaNs AFTER the other sample's sentinel and turn scipy's nan (nan in, nan out) into inf.
__device__ __forceinline__ bool le_total(float a, float b) { return (a <= b) || (b != b); }

__device__ __forceinline__ int merge_path_smem(const float* U, int nu, const float* V, int nv, int k) {
  int lo = k > nv ? k - nv : 0;
  int hi = k < nu ? k : nu;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (le_total(U[mid], V[k - mid - 1])) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// |a / na - b / nb| for ranks a <= na, b <= nb.  scipy divides each rank by its sample size in
// float64 and subtracts (two roundings of 1.1e-16 each on numbers near 1, so its difference
// carries an ABSOLUTE error of ~1e-16); here the difference is the exact integer a nb - b na
// times one rounded reciprocal: closer to the true value, and no float64 division per merged
// element (two of them were 60 % of the integral kernel's instructions).  Sample sizes of 2^31
// and more (sharded totals) keep scipy's form.
struct CdfScale {
  long long na, nb;
  double dna, dnb, inv;
  bool exact;
};
__device__ __forceinline__ CdfScale cdf_scale(long long na, long long nb) {
  CdfScale s;
  s.na = na, s.nb = nb, s.dna = (double)na, s.dnb = (double)nb;
  s.exact = na < (1ll << 31) && nb < (1ll << 31);
  s.inv = 1.0 / (s.dna * s.dnb);
  return s;
}
__device__ __forceinline__ double cdf_gap(const CdfScale& s, long long a, long long b) {
  if (s.exact) return fabs((double)(a * s.nb - b * s.na)) * s.inv;
  return fabs((double)a / s.dna - (double)b / s.dnb);
}

// number of u's among the first k merged elements (u before v on ties)
__device__ __forceinline__ int64_t merge_path(const float* __restrict__ U, int64_t nu,
                                              const float* __restrict__ V, int64_t nv, int64_t k) {
  int64_t lo = k > nv ? k - nv : 0;
  int64_t hi = k < nu ? k : nu;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (le_total(U[mid], V[k - mid - 1])) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// The same split found by a whole warp, 33-ary: every lane probes one point of the range and a
// ballot finds the first failing probe, so a 50 M-value range closes in 6 rounds of one load
// instead of 26 dependent ones.  (Used once per persistent block.  As a replacement of the
// one-thread search for EVERY tile boundary it was slower -- 83 us against 34 for 48 829 splits:
// the probes of a round touch 64 distinct sectors.)
__device__ __forceinline__ int64_t merge_path_warp(const float* __restrict__ U, int64_t nu,
                                                   const float* __restrict__ V, int64_t nv,
                                                   int64_t k) {
  const int lane = threadIdx.x & 31;
  int64_t lo = k > nv ? k - nv : 0;
  int64_t hi = k < nu ? k : nu;
  // the answer is in [lo, hi]; pred(mid) = U[mid] <= V[k - mid - 1] holds exactly for mid < answer
  while (lo < hi) {
    const int64_t span = hi - lo;
    const int64_t mid = lo + (span * (lane + 1)) / 33;   // lo <= mid < hi, non-decreasing in lane
    const bool pred = le_total(U[mid], V[k - mid - 1]);
    const int c = __popc(__ballot_sync(0xffffffffu, pred));   // the true probes are a prefix
    const int64_t mid_last_true = __shfl_sync(0xffffffffu, mid, c > 0 ? c - 1 : 0);
    const int64_t mid_first_false = __shfl_sync(0xffffffffu, mid, c < 32 ? c : 31);
    if (c > 0) lo = mid_last_true + 1;
    if (c < 32) hi = mid_first_false;
  }
  return lo < hi ? lo : hi;   // lo > hi only if NaNs break the monotone order
}

// merge-path split of every block boundary k = b * MRG_TILE of the binned integral, one thread
// each, done up front so the ~25 dependent global loads of a search are paid once in parallel
__global__ void __launch_bounds__(256)
merge_partition_kernel(const float* __restrict__ U, int64_t nu, const float* __restrict__ V,
                       int64_t nv, int64_t blocks, int64_t* __restrict__ splits) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= blocks) return;
  splits[b] = merge_path(U, nu, V, nv, b * MRG_TILE);
}

// A block's two input runs (at most MRG_TILE + 1 values each) into shared memory with every
// global load in flight before the first store: as a plain strided loop the 2 x 9 rounds each
// waited a full L2 round trip, which was most of a block's life.
__device__ __forceinline__ void stage_runs(const float* __restrict__ U, int lu,
                                           const float* __restrict__ V, int lv, float* su,
                                           float* sv) {
  constexpr int R = MRG_ITEMS + 1;   // MRG_TILE + 1 values = MRG_ITEMS rounds of the block + 1
  const int t = threadIdx.x;
  float ru[R], rv[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = t + r * MRG_THREADS;
    ru[r] = i < lu ? __ldg(U + i) : 0.f;
    rv[r] = i < lv ? __ldg(V + i) : 0.f;
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = t + r * MRG_THREADS;
    if (i < lu) su[i] = ru[r];
    if (i < lv) sv[i] = rv[r];
  }
}

// The integral of |F_u - F_v| over the merged sorted samples.  Persistent blocks: block b owns a
// contiguous range of MRG_TILE-position tiles, finds the merge-path split of its first tile with
// one warp-cooperative search and gets every later split for free (a tile ends where the next one
// starts), so there is no partition pass; a thread keeps its partial sum in a register across all
// its tiles.  Inside a tile every thread merges MRG_ITEMS consecutive positions from shared
// memory; the runs are closed by NaN sentinels (see merge_tile: no index tests for exhausted runs
// in the common case, and real NaNs merge last as they do in scipy), the value is widened to float64 once per
// element, and the rank difference a nb - b na is carried as an exactly-updated float64 (DBL:
// na nb < 2^53) instead of being rebuilt from two 64-bit products per element.
constexpr int MRG_PAD = 16;

template <bool DBL>
__device__ __forceinline__ double merge_tile(const float* su, int lu, const float* sv, int lv,
                                             int ka, int kb, long long rank_u0, long long rank_v0,
                                             const CdfScale& cs, int* i_at_kb) {
  int i = merge_path_smem(su, lu, sv, lv, ka);
  int j = ka - i;
  float a = su[i], b = sv[j];
  // both runs end in a NaN sentinel: "a <= b, or b is a NaN and u is not exhausted" is the whole
  // take-u test -- real NaNs merge last (u's before v's), an exhausted run never wins
  bool tu = (a <= b) || ((b != b) && i < lu);   // merged element ka
  double cur = (double)(tu ? a : b);
  i += tu ? 1 : 0, j += tu ? 0 : 1;
  a = su[i], b = sv[j];
  long long ru = rank_u0 + i, rv = rank_v0 + j;
  double dd = DBL ? (double)(ru * cs.nb - rv * cs.na) : 0.0;
  const double up = cs.dnb, down = -cs.dna;
  double acc = 0.0;
#pragma unroll
  for (int q = 0; q < MRG_ITEMS; ++q) {
    if (ka + q < kb) {
      tu = (a <= b) || ((b != b) && i < lu);   // merged element ka + q + 1 (exists: kb <= total - 1)
      const double nxt = (double)(tu ? a : b);
      const double gap = DBL ? fabs(dd) : cdf_gap(cs, ru, rv);
      acc = fma(gap, nxt - cur, acc);
      cur = nxt;
      if (tu) { ++i; a = su[i]; } else { ++j; b = sv[j]; }
      if (DBL) dd += tu ? up : down; else { ru += tu ? 1 : 0; rv += tu ? 0 : 1; }
    }
  }
  *i_at_kb = i - (tu ? 1 : 0);           // u's among the first kb merged elements of the tile
  return DBL ? acc * cs.inv : acc;
}

__global__ void __launch_bounds__(MRG_THREADS, 4)
cdf_integral_kernel(const float* __restrict__ U, int64_t nu, const float* __restrict__ V,
                    int64_t nv, int64_t u_below, int64_t v_below, int64_t nu_total,
                    int64_t nv_total, int64_t tiles, double* __restrict__ block_partials) {
  __shared__ float su[MRG_TILE + MRG_PAD];
  __shared__ float sv[MRG_TILE + MRG_PAD];
  __shared__ double warp_part[MRG_THREADS / 32];
  __shared__ long long split_s;
  const int64_t total = nu + nv;
  const int t = threadIdx.x;
  const int64_t tile_begin = (tiles * blockIdx.x) / gridDim.x;
  const int64_t tile_end = (tiles * (blockIdx.x + 1)) / gridDim.x;
  if (t < 32) {
    const int64_t s0 = merge_path_warp(U, nu, V, nv, tile_begin * MRG_TILE);
    if (t == 0) split_s = s0;
  }
  __syncthreads();
  const CdfScale cs = cdf_scale(nu_total, nv_total);
  const bool dbl = cs.dna * cs.dnb < 9007199254740992.0;   // 2^53: the rank difference is exact
  double acc = 0.0;
  for (int64_t tile = tile_begin; tile < tile_end; ++tile) {
    const int64_t i0 = split_s;
    const int64_t k0 = tile * MRG_TILE, j0 = k0 - i0;
    int64_t k1 = k0 + MRG_TILE;                     // contributions k in [k0, k1)
    if (k1 > total - 1) k1 = total - 1;
    const int len = (int)(k1 - k0);
    int lu = (int)((nu - i0) < (int64_t)(len + 1) ? (nu - i0) : (int64_t)(len + 1));
    int lv = (int)((nv - j0) < (int64_t)(len + 1) ? (nv - j0) : (int64_t)(len + 1));
    // NaN scores compare false either way, so a merge over them can run past the end of v (the
    // result is NaN, as scipy's is); the run lengths must stay valid array extents regardless
    if (lu < 0) lu = 0;
    if (lv < 0) lv = 0;
    stage_runs(U + i0, lu, V + j0, lv, su, sv);
    if (t < MRG_PAD) {
      if (lu + t < MRG_TILE + MRG_PAD) su[lu + t] = __int_as_float(0x7fc00000);   // NaN
      if (lv + t < MRG_TILE + MRG_PAD) sv[lv + t] = __int_as_float(0x7fc00000);   // NaN
    }
    __syncthreads();
    const int ka = t * MRG_ITEMS;
    if (ka < len) {
      const int kb = (ka + MRG_ITEMS) < len ? (ka + MRG_ITEMS) : len;
      int i_at_kb;
      acc += dbl ? merge_tile<true>(su, lu, sv, lv, ka, kb, u_below + i0, v_below + j0, cs, &i_at_kb)
                 : merge_tile<false>(su, lu, sv, lv, ka, kb, u_below + i0, v_below + j0, cs, &i_at_kb);
      if (kb == len) split_s = i0 + i_at_kb;        // exactly one thread: the next tile's split
    }
    __syncthreads();
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

// persistent grid of cdf_integral_kernel for `tiles` tiles (four resident blocks per SM)
inline unsigned integral_grid(int64_t tiles) {
  const int64_t cap = 148 * 4;
  return (unsigned)(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
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
constexpr int64_t BM_MAX_PER_BLOCK = (int64_t)1 << 19;  // values per block between flushes (16-bit event counts)

__device__ __forceinline__ uint32_t wb_key(float x) {
  const uint32_t b = __float_as_uint(x + 0.0f);  // -0.0 -> +0.0: one bin for the value zero
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

// lower edge of key bin b (b in [0, WB_BINS]) as float64; the inf/NaN bins decode as if the
// exponent range went on, so every bin has a finite width.  *ulp = float32 spacing inside bin b.
__device__ __forceinline__ double wb_edge(int b, double* ulp) {
  const uint32_t key_hi = (uint32_t)b;  // key >> 18 (15 bits for b = WB_BINS)
  const bool pos = key_hi >= (1u << 13);
  const uint32_t bits = pos ? ((key_hi - (1u << 13)) << WB_LOW_BITS)
                            : (~(key_hi << WB_LOW_BITS)) & 0x7FFFFFFFu;
  const int e = (int)(bits >> 23);      // 0 .. 256
  const int es = e > 1 ? e : 1;
  const double u = __longlong_as_double((long long)(es - 150 + 1023) << 52);  // 2^(es - 150)
  const double mag = (double)((bits & 0x7FFFFFu) + (e ? 0x800000u : 0u)) * u;   // exact
  if (ulp) *ulp = u;
  return pos ? mag : -mag;
}

// One pass over a sample: block-private shared-memory tables, two native 32-bit shared-memory
// operations per value and as few instructions around them as possible -- ncu showed the pass
// bound by instruction issue / the INT32 pipe (64 lanes per clock and SM), not by HBM (40 %) or
// by the ATOMS wavefronts (36 %):
//   L[b] += low 18 key bits   (atom with return: the adding thread sees whether ITS add wrapped
//                              the word -- possible only when the old word's top 14 bits were all
//                              ones, one add in 16384 -- and then credits 2^20 to H[b]);
//   H[b] += 1                 (red, fire and forget; immediate address offset).
// So count = H & 0xFFFFF and the offset sum K = (H >> 20) 2^32 + L, exactly; a block sees at most
// 2^19 values between flushes (count < 2^20, at most 32 wraps).  History: two dependent atomics
// with the carry computed for every value (55.6 us per 50 M values), one packed atomic with the
// count in the top byte (same time: its carry / wrap test fired in every second warp and cost
// more instructions than the second atomic), this (see DESIGN.md for the numbers).  Flushed with
// one 64-bit reduction per non-empty bin and table.
constexpr int BM_SMEM = 2 * WB_BINS * (int)sizeof(uint32_t);

__device__ __forceinline__ uint32_t atoms_add(uint32_t addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}
template <int OFF>
__device__ __forceinline__ void reds_add(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0+%2], %1;" ::"r"(addr), "r"(v), "n"(OFF) : "memory");
}

// key = bits ^ (sign ? 0xFFFFFFFF : 0x80000000); bin = key >> 18, low = key & 0x3FFFF.  The sign
// mask and the word address are IMAD.HI (FMA pipe) instead of shifts (INT32 pipe).
__device__ __forceinline__ void bm_add(float v, uint32_t l_base) {
  const uint32_t b = __float_as_uint(v + 0.0f);   // -0.0 -> +0.0: one bin for the value zero
  uint32_t sign, addr;
  asm("mul.hi.s32 %0, %1, 1;" : "=r"(sign) : "r"(b));                     // b < 0 ? ~0 : 0
  const uint32_t k = b ^ (sign | 0x80000000u);
  asm("mad.hi.u32 %0, %1, 65536, %2;" : "=r"(addr) : "r"(k & 0xFFFC0000u), "r"(l_base));
  const uint32_t low = k & WB_LOW_MASK;
  const uint32_t old = atoms_add(addr, low);
  reds_add<WB_BINS * 4>(addr, 1u);
  if ((~old & 0xFFFC0000u) == 0u)                  // necessary for a wrap; exact test inside
    if (old + low < old) reds_add<WB_BINS * 4>(addr, 1u << 20);
}

// this block's grid-stride share of x[0 .. n); l_base = shared-memory address of L (H follows)
__device__ __forceinline__ void bm_accumulate(const float* __restrict__ x, int64_t n,
                                              uint32_t w_base) {
  for_each_value<BM_THREADS>(x, n, [&](float v) { bm_add(v, w_base); });
}

// adds the block-private tables to the global ones and leaves them zeroed
__device__ __forceinline__ void bm_flush(uint32_t* L, uint32_t* H, unsigned long long* cnt,
                                         unsigned long long* ksum) {
  __syncthreads();
  for (int b = threadIdx.x; b < WB_BINS; b += BM_THREADS) {
    const uint32_t h = H[b];
    if (h) {
      atomicAdd(&cnt[b], (unsigned long long)(h & 0xFFFFFu));
      atomicAdd(&ksum[b], ((unsigned long long)(h >> 20) << 32) + (unsigned long long)L[b]);
      H[b] = 0;
      L[b] = 0;
    }
  }
  __syncthreads();
}

// the pass as a kernel of its own: the sharded method's per-rank step (uq_bin_moments)
__global__ void __launch_bounds__(BM_THREADS, 1)
bin_moments_kernel(const float* __restrict__ x, int64_t n, unsigned long long* __restrict__ cnt,
                   unsigned long long* __restrict__ ksum) {
  extern __shared__ __align__(16) uint32_t bm_sh[];
  for (int i = threadIdx.x; i < 2 * WB_BINS; i += BM_THREADS) bm_sh[i] = 0;
  __syncthreads();
  bm_accumulate(x, n, (uint32_t)__cvta_generic_to_shared(bm_sh));
  bm_flush(bm_sh, bm_sh + WB_BINS, cnt, ksum);
}

constexpr int RS_BLOCKS = 64;  // bin_contrib_kernel: 64 blocks x 256 threads, one bin per thread
struct BinnedResult {          // zeroed together with the tables
  long long amb_u, amb_v;      // values of each sample in ambiguous bins
  long long nonfinite;         // values in the inf / NaN bins (the caller falls back to sorting)
  double parts[RS_BLOCKS];     // per-block sums of the sign-definite bins' contributions
};

// exclusive block-wide scan of a pair of values (1024 threads); ta / tb = block totals
__device__ __forceinline__ void block_scan_pair(long long a, long long b, long long& ea,
                                                long long& eb, long long& ta, long long& tb,
                                                long long* sm /* [64] */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  long long ia = a, ib = b;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long xa = __shfl_up_sync(0xffffffffu, ia, o), xb = __shfl_up_sync(0xffffffffu, ib, o);
    if (lane >= o) ia += xa, ib += xb;
  }
  if (lane == 31) sm[w] = ia, sm[32 + w] = ib;
  __syncthreads();
  const long long wa = sm[lane], wb = sm[32 + lane];  // every warp scans the 32 warp totals
  long long sa = wa, sb = wb;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long xa = __shfl_up_sync(0xffffffffu, sa, o), xb = __shfl_up_sync(0xffffffffu, sb, o);
    if (lane >= o) sa += xa, sb += xb;
  }
  const long long base_a = __shfl_sync(0xffffffffu, sa - wa, w);
  const long long base_b = __shfl_sync(0xffffffffu, sb - wb, w);
  ta = __shfl_sync(0xffffffffu, sa, 31);
  tb = __shfl_sync(0xffffffffu, sb, 31);
  ea = base_a + ia - a;
  eb = base_b + ib - b;
  __syncthreads();
}

// Exclusive prefix sums of a pair of per-bin tables, one bin per thread, 16 blocks of 1024.
// Block j first adds up the bins of the blocks below it (thread t takes bin i * 1024 + t of every
// lower block i: coalesced, independent loads), then scans its own 1024 bins.  A single block
// walking all 16384 bins is limited by one SM's sector rate (measured 27-43 us; this: a few us).
// With `flags` only flagged bins count (prefix of the ambiguous values) and the output is
// `minus - prefix` (the rank offsets skip_x of the exact pass).
constexpr int SCAN_BLOCKS = WB_BINS / 1024;
__global__ void __launch_bounds__(1024, 1)
bin_scan_kernel(const unsigned long long* __restrict__ cnt_a,
                const unsigned long long* __restrict__ cnt_b, const uint8_t* __restrict__ flags,
                const long long* __restrict__ minus_a, const long long* __restrict__ minus_b,
                long long* __restrict__ out_a, long long* __restrict__ out_b) {
  __shared__ long long sm[64];
  const int t = threadIdx.x;
  const int j = blockIdx.x;
  const int b = j * 1024 + t;
  long long sa = 0, sb = 0;  // counts stay below 2^31: only the low words are loaded
#pragma unroll 4
  for (int i = 0; i < j; ++i) {
    const int bb = i * 1024 + t;
    if (flags == nullptr || flags[bb]) sa += (uint32_t)cnt_a[bb], sb += (uint32_t)cnt_b[bb];
  }
  const bool on = flags == nullptr || flags[b];
  const long long ca = on ? (long long)(uint32_t)cnt_a[b] : 0;
  const long long cb = on ? (long long)(uint32_t)cnt_b[b] : 0;
  const long long ma = minus_a ? minus_a[b] : 0, mb = minus_b ? minus_b[b] : 0;
  long long ea, eb, base_a, base_b, ta, tb;
  block_scan_pair(sa, sb, ea, eb, base_a, base_b, sm);
  block_scan_pair(ca, cb, ea, eb, ta, tb, sm);
  ea += base_a, eb += base_b;
  out_a[b] = minus_a ? ma - ea : ea;
  out_b[b] = minus_b ? mb - eb : eb;
}

// One bin per thread: contribution of the bin if D = F_u - F_v provably keeps one sign on it,
// else the ambiguity flag; bin edges for the exact pass.
__global__ void __launch_bounds__(WB_BINS / RS_BLOCKS)
bin_contrib_kernel(const unsigned long long* __restrict__ cnt_u,
                   const unsigned long long* __restrict__ ks_u,
                   const unsigned long long* __restrict__ cnt_v,
                   const unsigned long long* __restrict__ ks_v,
                   const long long* __restrict__ pre_u, const long long* __restrict__ pre_v,
                   long long nu, long long nv, uint8_t* __restrict__ flags,
                   double* __restrict__ edges, BinnedResult* __restrict__ res) {
  constexpr int T = WB_BINS / RS_BLOCKS;
  __shared__ double sd[T];
  const int t = threadIdx.x;
  const int b = blockIdx.x * T + t;
  const long long cu = (long long)cnt_u[b], cv = (long long)cnt_v[b];
  const long long Cu = pre_u[b], Cv = pre_v[b];
  const unsigned long long ku = ks_u[b], kv = ks_v[b];
  double ulp;
  const double tb = wb_edge(b, &ulp);
  const double w = wb_edge(b + 1, nullptr) - tb;
  edges[b] = tb;
  if (b == WB_BINS - 1) edges[WB_BINS] = tb + w;
  const bool d_ge0 = Cu * nv - (Cv + cv) * nu >= 0;  // D >= 0 on the whole bin (|.| < 2^62)
  const bool d_le0 = (Cu + cu) * nv - Cv * nu <= 0;  // D <= 0 on the whole bin
  double acc = 0.0;
  const bool amb = !(d_ge0 || d_le0);
  if (!amb && ((cu | cv) || Cu * nv != Cv * nu)) {
    const double a = ((double)(Cu + cu) * w - (double)ku * ulp) / (double)nu;
    const double c = ((double)(Cv + cv) * w - (double)kv * ulp) / (double)nv;
    acc = fabs(a - c);
  }
  flags[b] = amb ? 1 : 0;
  if (amb) {  // integer atomics: order-independent
    atomicAdd(reinterpret_cast<unsigned long long*>(&res->amb_u), (unsigned long long)cu);
    atomicAdd(reinterpret_cast<unsigned long long*>(&res->amb_v), (unsigned long long)cv);
  }
  if ((b < 32 || b >= WB_BINS - 32) && (cu | cv))
    atomicAdd(reinterpret_cast<unsigned long long*>(&res->nonfinite), (unsigned long long)(cu + cv));
  sd[t] = acc;
  for (int o = T / 2; o > 0; o >>= 1) {  // fixed-order tree: deterministic
    __syncthreads();
    if (t < o) sd[t] += sd[t + o];
  }
  if (t == 0) res->parts[blockIdx.x] = sd[0];
}

// ---- the binned method as ONE cooperative launch ----------------------------------------------
// Phase 1: every block accumulates its grid-stride share of u, flushes, then of v, flushes (the
// block-private tables of bin_moments_kernel; a block never holds more than BM_MAX_PER_BLOCK
// values between flushes).  Grid barrier.  Phase 2: the 64 units of bin_contrib_kernel (256 bins
// each) are dealt to the blocks; a unit first adds up the counts of the bins below it (coalesced,
// <= 16 loads per thread), scans its own 256 bins and evaluates the sign tests and contributions
// with the arithmetic of bin_contrib_kernel (same fixed-order tree, so the result is bit-identical
// to the three-kernel path).  The last block to finish adds the 64 unit sums in a fixed order and
// writes the result record into mapped host memory.
struct FusedCtl {
  unsigned int barrier, ticket;
};
struct FusedRecord {     // what the host reads after the stream synchronisation
  double resolved;
  long long amb_u, amb_v, nonfinite;
};

__global__ void __launch_bounds__(BM_THREADS, 1)
wasserstein_binned_fused_kernel(const float* __restrict__ u, int64_t nu,
                                const float* __restrict__ v, int64_t nv,
                                unsigned long long* __restrict__ tables,  // cnt_u|ks_u|cnt_v|ks_v
                                long long* __restrict__ pre_u, long long* __restrict__ pre_v,
                                uint8_t* __restrict__ flags, double* __restrict__ edges,
                                BinnedResult* __restrict__ res, FusedCtl* __restrict__ ctl,
                                FusedRecord* __restrict__ record) {
  extern __shared__ __align__(16) uint32_t bm_sh[];
  __shared__ long long scan_sm[64];
  __shared__ double sd[WB_BINS / RS_BLOCKS];
  __shared__ bool last;
  uint32_t* L = bm_sh;              // offset sums
  uint32_t* H = bm_sh + WB_BINS;    // counts (+ 2^20 per wrap of L)
  unsigned long long* cnt_u = tables;
  unsigned long long* ks_u = tables + WB_BINS;
  unsigned long long* cnt_v = tables + 2 * WB_BINS;
  unsigned long long* ks_v = tables + 3 * WB_BINS;
  const int t = threadIdx.x;

  // ---- phase 1
  const uint32_t w_base = (uint32_t)__cvta_generic_to_shared(bm_sh);
  for (int i = t; i < 2 * WB_BINS; i += BM_THREADS) bm_sh[i] = 0;
  __syncthreads();
  const int64_t round = (int64_t)gridDim.x * BM_MAX_PER_BLOCK;  // values per flush round (x 4 | round)
  for (int64_t r0 = 0; r0 < nu; r0 += round) {
    bm_accumulate(u + r0, min(round, nu - r0), w_base);
    bm_flush(L, H, cnt_u, ks_u);
  }
  for (int64_t r0 = 0; r0 < nv; r0 += round) {
    bm_accumulate(v + r0, min(round, nv - r0), w_base);
    bm_flush(L, H, cnt_v, ks_v);
  }
  grid_barrier(&ctl->barrier, 1);

  // ---- phase 2
  constexpr int T = WB_BINS / RS_BLOCKS;  // 256 bins per unit
  for (int j = blockIdx.x; j < RS_BLOCKS; j += gridDim.x) {
    long long sa = 0, sb = 0;
    for (int i = t; i < j * T; i += BM_THREADS)
      sa += (long long)__ldcg(cnt_u + i), sb += (long long)__ldcg(cnt_v + i);
    const int b = j * T + t;
    long long cu = 0, cv = 0;
    if (t < T) cu = (long long)__ldcg(cnt_u + b), cv = (long long)__ldcg(cnt_v + b);
    long long ea, eb, base_a, base_b, ta, tb;
    block_scan_pair(sa, sb, ea, eb, base_a, base_b, scan_sm);
    block_scan_pair(cu, cv, ea, eb, ta, tb, scan_sm);
    double acc = 0.0;
    if (t < T) {
      const long long Cu = base_a + ea, Cv = base_b + eb;
      pre_u[b] = Cu;
      pre_v[b] = Cv;
      const unsigned long long ku = __ldcg(ks_u + b), kv = __ldcg(ks_v + b);
      double ulp;
      const double tb0 = wb_edge(b, &ulp);
      const double w = wb_edge(b + 1, nullptr) - tb0;
      edges[b] = tb0;
      if (b == WB_BINS - 1) edges[WB_BINS] = tb0 + w;
      const bool d_ge0 = Cu * nv - (Cv + cv) * nu >= 0;
      const bool d_le0 = (Cu + cu) * nv - Cv * nu <= 0;
      const bool amb = !(d_ge0 || d_le0);
      if (!amb && ((cu | cv) || Cu * nv != Cv * nu)) {
        const double a = ((double)(Cu + cu) * w - (double)ku * ulp) / (double)nu;
        const double c = ((double)(Cv + cv) * w - (double)kv * ulp) / (double)nv;
        acc = fabs(a - c);
      }
      flags[b] = amb ? 1 : 0;
      if (amb) {
        atomicAdd(reinterpret_cast<unsigned long long*>(&res->amb_u), (unsigned long long)cu);
        atomicAdd(reinterpret_cast<unsigned long long*>(&res->amb_v), (unsigned long long)cv);
      }
      if ((b < 32 || b >= WB_BINS - 32) && (cu | cv))
        atomicAdd(reinterpret_cast<unsigned long long*>(&res->nonfinite),
                  (unsigned long long)(cu + cv));
      sd[t] = acc;
    }
    for (int o = T / 2; o > 0; o >>= 1) {  // the fixed-order tree of bin_contrib_kernel
      __syncthreads();
      if (t < o) sd[t] += sd[t + o];
    }
    if (t == 0) res->parts[j] = sd[0];
    __syncthreads();
  }

  // ---- the last block to finish phase 2 writes the record
  if (t == 0) {
    __threadfence();
    last = atomicAdd(&ctl->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && t == 0) {
    __threadfence();
    double s = 0.0;
    for (int q = 0; q < RS_BLOCKS; ++q) s += __ldcg(&res->parts[q]);  // fixed order
    record->resolved = s;
    record->amb_u = __ldcg(&res->amb_u);
    record->amb_v = __ldcg(&res->amb_v);
    record->nonfinite = __ldcg(&res->nonfinite);
    __threadfence_system();
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
  stage_runs(U + i0, lu, V + j0, lv, su, sv);
  __syncthreads();

  const CdfScale cs = cdf_scale(nu_total, nv_total);
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
        acc += cdf_gap(cs, ou + i, ov + j) * ((double)cur - __ldg(edges + b));
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
      acc += cdf_gap(cs, ou + i, ov + j) * seg;
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
  size_t u, ut, v, vt, scratch, parts, splits, result, tables, bres, ctl, pre, flags, skips, edges, cursors, total;
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
  L.bres = o; o += al(sizeof(BinnedResult));                        // zeroed with the tables
  L.ctl = o; o += 256;                                              // barrier + ticket, zeroed too
  L.pre = o; o += al(sizeof(long long) * 2 * WB_BINS);
  L.flags = o; o += al(WB_BINS);
  L.skips = o; o += al(sizeof(long long) * 2 * (WB_BINS + 1));
  L.edges = o; o += al(sizeof(double) * (WB_BINS + 1));
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
  static PerDeviceOnce opted;
  constexpr int SMEM = BM_SMEM;
  if (int rc = smem_opt_in(bin_moments_kernel, SMEM, opted)) return rc;
  // one block per SM; more only to keep a block below BM_MAX_PER_BLOCK values
  int64_t blocks = (n + (int64_t)BM_THREADS * 16 - 1) / ((int64_t)BM_THREADS * 16);
  if (blocks > 148) blocks = 148;
  const int64_t need = (n + BM_MAX_PER_BLOCK - 1) / BM_MAX_PER_BLOCK;
  if (blocks < need) blocks = (need + 147) / 148 * 148;
  if (blocks < 1) blocks = 1;
  bin_moments_kernel<<<(unsigned)blocks, BM_THREADS, SMEM, st>>>(x, n, cnt, ks);
  UQ_LAUNCH_CHECK();
  return UQ_OK;
}

// The sort method: sorted copies of u and v in the workspace buffers (the inputs are only read).
int sort_and_integrate(const float* u, float* du, float* dut, int64_t nu, const float* v, float* dv,
                       float* dvt, int64_t nv, char* b, const WsLayout& L, double* result,
                       cudaStream_t st) {
  double* parts = reinterpret_cast<double*>(b + L.parts);
  float *su = nullptr, *sv = nullptr;
  int rc = radix_sort_f32_copy(u, du, dut, nu, b + L.scratch, radix_sort_scratch_bytes(nu), &su, st);
  if (rc != UQ_OK) return rc;
  rc = radix_sort_f32_copy(v, dv, dvt, nv, b + L.scratch, radix_sort_scratch_bytes(nv), &sv, st);
  if (rc != UQ_OK) return rc;
  if (nu + nv - 1 > 0) {
    const unsigned grid = integral_grid(L.blocks);
    cdf_integral_kernel<<<grid, MRG_THREADS, 0, st>>>(su, nu, sv, nv, 0, 0, nu, nv, L.blocks,
                                                      parts);
    UQ_LAUNCH_CHECK();
    sum_partials_kernel<<<1, 1024, 0, st>>>(parts, grid, result);
    UQ_LAUNCH_CHECK();
  } else {
    UQ_CUDA(cudaMemsetAsync(result, 0, sizeof(double), st));
  }
  return UQ_OK;
}

}  // namespace

namespace {

struct BinTables {
  unsigned long long *cnt_u, *ks_u, *cnt_v, *ks_v;
};
BinTables tables_at(unsigned long long* t) { return {t, t + WB_BINS, t + 2 * WB_BINS, t + 3 * WB_BINS}; }

// tables (already complete, e.g. all-reduced over ranks) -> flags, edges, prefix counts and the
// host copy of BinnedResult; synchronises the stream.  `bres` must be zeroed.
int resolve_bins(const BinTables& T, int64_t nu, int64_t nv, char* b, const WsLayout& L,
                 uint8_t* flags, BinnedResult* h, double* resolved, cudaStream_t st) {
  long long* pre_u = reinterpret_cast<long long*>(b + L.pre);
  long long* pre_v = pre_u + WB_BINS;
  double* edges = reinterpret_cast<double*>(b + L.edges);
  BinnedResult* bres = reinterpret_cast<BinnedResult*>(b + L.bres);
  bin_scan_kernel<<<SCAN_BLOCKS, 1024, 0, st>>>(T.cnt_u, T.cnt_v, nullptr, nullptr, nullptr, pre_u, pre_v);
  UQ_LAUNCH_CHECK();
  bin_contrib_kernel<<<RS_BLOCKS, WB_BINS / RS_BLOCKS, 0, st>>>(
      T.cnt_u, T.ks_u, T.cnt_v, T.ks_v, pre_u, pre_v, (long long)nu, (long long)nv, flags, edges,
      bres);
  UQ_LAUNCH_CHECK();
  UQ_CUDA(cudaMemcpyAsync(h, bres, sizeof(*h), cudaMemcpyDeviceToHost, st));
  UQ_CUDA(cudaStreamSynchronize(st));
  *resolved = 0.0;
  for (int q = 0; q < RS_BLOCKS; ++q) *resolved += h->parts[q];  // fixed order
  return UQ_OK;
}

// Exact integral over the ambiguous bins.  du / dv hold the (unsorted) ambiguous values and are
// clobbered; pre / edges / flags come from resolve_bins on the same workspace.  Synchronises.
int ambiguous_exact(const BinTables& T, const uint8_t* flags, float* du, float* dut, int64_t amb_u,
                    float* dv, float* dvt, int64_t amb_v, int64_t nu, int64_t nv, char* b,
                    const WsLayout& L, double* exact_host, cudaStream_t st) {
  long long* pre_u = reinterpret_cast<long long*>(b + L.pre);
  long long* pre_v = pre_u + WB_BINS;
  long long* skip_u = reinterpret_cast<long long*>(b + L.skips);
  long long* skip_v = skip_u + WB_BINS + 1;
  double* edges = reinterpret_cast<double*>(b + L.edges);
  double* result = reinterpret_cast<double*>(b + L.result);
  bin_scan_kernel<<<SCAN_BLOCKS, 1024, 0, st>>>(T.cnt_u, T.cnt_v, flags, pre_u, pre_v, skip_u, skip_v);
  UQ_LAUNCH_CHECK();
  float *su = du, *sv = dv;
  int rc;
  if (amb_u > 0) {
    rc = radix_sort_f32(du, dut, amb_u, b + L.scratch, radix_sort_scratch_bytes(amb_u), &su, st);
    if (rc != UQ_OK) return rc;
  }
  if (amb_v > 0) {
    rc = radix_sort_f32(dv, dvt, amb_v, b + L.scratch, radix_sort_scratch_bytes(amb_v), &sv, st);
    if (rc != UQ_OK) return rc;
  }
  const int64_t amb = amb_u + amb_v;
  const int64_t blocks = (amb + MRG_TILE - 1) / MRG_TILE;
  int64_t* splits = reinterpret_cast<int64_t*>(b + L.splits);
  double* parts = reinterpret_cast<double*>(b + L.parts);
  merge_partition_kernel<<<(unsigned)((blocks + 255) / 256), 256, 0, st>>>(su, amb_u, sv, amb_v,
                                                                           blocks, splits);
  UQ_LAUNCH_CHECK();
  cdf_integral_binned_kernel<<<(unsigned)blocks, MRG_THREADS, 0, st>>>(
      su, amb_u, sv, amb_v, skip_u, skip_v, edges, nu, nv, splits, parts);
  UQ_LAUNCH_CHECK();
  sum_partials_kernel<<<1, 1024, 0, st>>>(parts, blocks, result);
  UQ_LAUNCH_CHECK();
  UQ_CUDA(cudaMemcpyAsync(exact_host, result, sizeof(double), cudaMemcpyDeviceToHost, st));
  UQ_CUDA(cudaStreamSynchronize(st));
  return UQ_OK;
}

int compact_launch(const float* x, int64_t n, const uint8_t* flags, float* out,
                   unsigned long long* cursor, cudaStream_t st) {
  const int64_t tile = 256 * 16;
  compact_flagged_kernel<<<(unsigned)((n + tile - 1) / tile), 256, 0, st>>>(x, n, flags, out,
                                                                            cursor);
  UQ_LAUNCH_CHECK();
  return UQ_OK;
}

}  // namespace

namespace {

// tables, BinnedResult and the control words are zeroed and the fused kernel is launched; its last
// block writes the FusedRecord to `record` (device-visible: mapped pinned host memory).  No
// synchronisation.
int binned_fused_launch(const float* u, int64_t nu, const float* v, int64_t nv, char* b,
                        const WsLayout& L, FusedRecord* record, cudaStream_t st) {
  static PerDeviceOnce opted;
  static PerDeviceInt grid_cache;
  constexpr int SMEM = BM_SMEM;
  if (int rc = smem_opt_in(wasserstein_binned_fused_kernel, SMEM, opted)) return rc;
  int grid = 0;
  if (int rc = coop_grid_limit(wasserstein_binned_fused_kernel, BM_THREADS, SMEM, grid_cache, &grid))
    return rc;
  unsigned long long* tables = reinterpret_cast<unsigned long long*>(b + L.tables);
  long long* pre_u = reinterpret_cast<long long*>(b + L.pre);
  long long* pre_v = pre_u + WB_BINS;
  uint8_t* flags = reinterpret_cast<uint8_t*>(b + L.flags);
  double* edges = reinterpret_cast<double*>(b + L.edges);
  BinnedResult* bres = reinterpret_cast<BinnedResult*>(b + L.bres);
  FusedCtl* ctl = reinterpret_cast<FusedCtl*>(b + L.ctl);
  UQ_CUDA(cudaMemsetAsync(b + L.tables, 0, L.pre - L.tables, st));
  void* args[] = {(void*)&u, (void*)&nu, (void*)&v, (void*)&nv, (void*)&tables, (void*)&pre_u,
                  (void*)&pre_v, (void*)&flags, (void*)&edges, (void*)&bres, (void*)&ctl,
                  (void*)&record};
  UQ_CUDA(cudaLaunchCooperativeKernel((const void*)wasserstein_binned_fused_kernel, dim3(grid),
                                      dim3(BM_THREADS), args, SMEM, st));
  UQ_LAUNCH_CHECK();
  return UQ_OK;
}

// the synchronous form: this thread's mapped result slot, one stream synchronisation
int binned_fused(const float* u, int64_t nu, const float* v, int64_t nv, char* b, const WsLayout& L,
                 FusedRecord* h, cudaStream_t st) {
  void *slot_h = nullptr, *slot_d = nullptr;
  if (int rc = result_slot(&slot_h, &slot_d)) return rc;
  static_cast<FusedRecord*>(slot_h)->nonfinite = -1;   // the kernel overwrites it with a count
  if (int rc = binned_fused_launch(u, nu, v, nv, b, L, static_cast<FusedRecord*>(slot_d), st))
    return rc;
  UQ_CUDA(cudaStreamSynchronize(st));
  memcpy(h, slot_h, sizeof(*h));
  UQ_REQUIRE(h->nonfinite >= 0, UQ_ERR_CUDA, "wasserstein: the kernel left no result record");
  return UQ_OK;
}

}  // namespace

// ---- enqueue / finish: the same metric without a synchronisation inside the call ----------------
// enqueue = the one-pass binned method's memset + launch on `st`; `record` is caller-owned mapped
// pinned host memory (>= UQ_METRIC_RECORD_BYTES) that the kernel's last block fills in.  After the
// caller has synchronised the stream, finish reads the record: if every bin was resolved from
// the tables (the ID-vs-OOD case) the distance is there; otherwise (ambiguous bins, inf / NaN)
// the synchronous call runs.  Several metrics can so be in flight behind ONE synchronisation.
int wasserstein_1d_enqueue(const float* u, int64_t nu, const float* v, int64_t nv, void* record,
                           void* ws, size_t ws_bytes, cudaStream_t st) {
  const WsLayout L = layout(nu, nv);
  UQ_REQUIRE(ws != nullptr && ws_bytes >= L.total, UQ_ERR_WORKSPACE,
             "wasserstein needs %zu workspace bytes, got %zu", L.total, ws_bytes);
  UQ_REQUIRE(nu + nv < ((int64_t)1 << 31), UQ_ERR_INVALID, "wasserstein: too many values");
  void* record_dev = nullptr;
  if (!record || cudaHostGetDevicePointer(&record_dev, record, 0) != cudaSuccess) {
    (void)cudaGetLastError();   // the failed query must not show up in the next launch check
    set_error("wasserstein enqueue: the record must be mapped pinned host memory");
    return UQ_ERR_INVALID;
  }
  static_assert(sizeof(FusedRecord) <= UQ_METRIC_RECORD_BYTES, "record size");
  static_cast<FusedRecord*>(record)->nonfinite = -1;
  return binned_fused_launch(u, nu, v, nv, static_cast<char*>(ws), L,
                             static_cast<FusedRecord*>(record_dev), st);
}

int wasserstein_1d_finish(const float* u, int64_t nu, const float* v, int64_t nv,
                          const void* record, double* out_host, int64_t* info_host, void* ws,
                          size_t ws_bytes, cudaStream_t st) {
  FusedRecord h;
  memcpy(&h, record, sizeof(h));
  UQ_REQUIRE(h.nonfinite >= 0, UQ_ERR_INVALID,
             "wasserstein finish: no result record (synchronise the stream of the enqueue first)");
  if (h.nonfinite == 0 && h.amb_u + h.amb_v == 0) {
    *out_host = h.resolved;
    if (info_host) info_host[0] = UQ_WASSERSTEIN_BINNED, info_host[1] = 0, info_host[2] = 0;
    return UQ_OK;
  }
  return wasserstein_1d(u, nu, v, nv, UQ_WASSERSTEIN_AUTO, out_host, info_host, ws, ws_bytes, st);
}

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
    const BinTables T = tables_at(reinterpret_cast<unsigned long long*>(b + L.tables));
    uint8_t* flags = reinterpret_cast<uint8_t*>(b + L.flags);
    FusedRecord h;
    int rc = binned_fused(u, nu, v, nv, b, L, &h, st);  // one memset + one launch + one sync
    if (rc != UQ_OK) return rc;
    const double resolved = h.resolved;
    const int64_t amb = h.amb_u + h.amb_v;
    const bool use = h.nonfinite == 0 &&
                     (method == UQ_WASSERSTEIN_BINNED || amb <= (nu + nv) / 2);
    if (use) {
      if (info_host) info_host[0] = UQ_WASSERSTEIN_BINNED, info_host[1] = h.amb_u, info_host[2] = h.amb_v;
      double exact = 0.0;
      if (amb > 0) {
        unsigned long long* cursors = reinterpret_cast<unsigned long long*>(b + L.cursors);
        UQ_CUDA(cudaMemsetAsync(cursors, 0, 2 * sizeof(unsigned long long), st));
        rc = compact_launch(u, nu, flags, du, cursors, st);
        if (rc != UQ_OK) return rc;
        rc = compact_launch(v, nv, flags, dv, cursors + 1, st);
        if (rc != UQ_OK) return rc;
        rc = ambiguous_exact(T, flags, du, dut, h.amb_u, dv, dvt, h.amb_v, nu, nv, b, L, &exact, st);
        if (rc != UQ_OK) return rc;
      }
      *out_host = resolved + exact;
      return UQ_OK;
    }
  }

  const int rc = sort_and_integrate(u, du, dut, nu, v, dv, dvt, nv, b, L, result, st);
  if (rc != UQ_OK) return rc;
  UQ_CUDA(cudaMemcpyAsync(out_host, result, sizeof(double), cudaMemcpyDeviceToHost, st));
  UQ_CUDA(cudaStreamSynchronize(st));
  return UQ_OK;
}

// ---- per-rank steps of the sharded binned method (nnueehcs_b200.distributed) --------------------

int bin_moments_accumulate(const float* x, int64_t n, unsigned long long* cnt,
                           unsigned long long* ksum, cudaStream_t st) {
  return bin_moments_launch(x, n, cnt, ksum, st);
}

// tables = [cnt_u | ks_u | cnt_v | ks_v] of the WHOLE samples.  out_host = {sum of the resolved
// bins, ambiguous u values, ambiguous v values, values in inf/NaN bins}; flags_out[bin] = ambiguous.
int wasserstein_from_bins(const unsigned long long* tables, int64_t nu_total, int64_t nv_total,
                          uint8_t* flags_out, double* out_host, void* ws, size_t ws_bytes,
                          cudaStream_t st) {
  const WsLayout L = layout(1, 1);
  UQ_REQUIRE(ws != nullptr && ws_bytes >= L.total, UQ_ERR_WORKSPACE,
             "wasserstein_from_bins needs %zu workspace bytes, got %zu", L.total, ws_bytes);
  char* b = static_cast<char*>(ws);
  const BinTables T = tables_at(const_cast<unsigned long long*>(tables));
  UQ_CUDA(cudaMemsetAsync(b + L.bres, 0, sizeof(BinnedResult), st));
  BinnedResult h;
  double resolved;
  const int rc = resolve_bins(T, nu_total, nv_total, b, L, flags_out, &h, &resolved, st);
  if (rc != UQ_OK) return rc;
  out_host[0] = resolved;
  out_host[1] = (double)h.amb_u;
  out_host[2] = (double)h.amb_v;
  out_host[3] = (double)h.nonfinite;
  return UQ_OK;
}

int compact_flagged(const float* x, int64_t n, const uint8_t* flags, float* out,
                    int64_t* count_host, void* ws, size_t ws_bytes, cudaStream_t st) {
  UQ_REQUIRE(ws != nullptr && ws_bytes >= 8, UQ_ERR_WORKSPACE, "compact_flagged: workspace too small");
  unsigned long long* cursor = static_cast<unsigned long long*>(ws);
  UQ_CUDA(cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), st));
  const int rc = compact_launch(x, n, flags, out, cursor, st);
  if (rc != UQ_OK) return rc;
  unsigned long long c = 0;
  UQ_CUDA(cudaMemcpyAsync(&c, cursor, sizeof(c), cudaMemcpyDeviceToHost, st));
  UQ_CUDA(cudaStreamSynchronize(st));
  *count_host = (int64_t)c;
  return UQ_OK;
}

// Exact part: u_amb / v_amb = every ambiguous value of the whole samples (any order, left
// untouched), tables as above.  Writes the integral over the ambiguous bins.
int wasserstein_ambiguous(const float* u_amb, int64_t nu_amb, const float* v_amb, int64_t nv_amb,
                          const unsigned long long* tables, int64_t nu_total, int64_t nv_total,
                          double* out_host, void* ws, size_t ws_bytes, cudaStream_t st) {
  const WsLayout L = layout(nu_amb > 0 ? nu_amb : 1, nv_amb > 0 ? nv_amb : 1);
  UQ_REQUIRE(ws != nullptr && ws_bytes >= L.total, UQ_ERR_WORKSPACE,
             "wasserstein_ambiguous needs %zu workspace bytes, got %zu", L.total, ws_bytes);
  UQ_REQUIRE(nu_amb + nv_amb >= 1 && nu_amb + nv_amb < ((int64_t)1 << 31), UQ_ERR_INVALID,
             "wasserstein_ambiguous: %lld values", (long long)(nu_amb + nv_amb));
  char* b = static_cast<char*>(ws);
  float* du = reinterpret_cast<float*>(b + L.u);
  float* dut = reinterpret_cast<float*>(b + L.ut);
  float* dv = reinterpret_cast<float*>(b + L.v);
  float* dvt = reinterpret_cast<float*>(b + L.vt);
  uint8_t* flags = reinterpret_cast<uint8_t*>(b + L.flags);
  const BinTables T = tables_at(const_cast<unsigned long long*>(tables));
  UQ_CUDA(cudaMemsetAsync(b + L.bres, 0, sizeof(BinnedResult), st));
  BinnedResult h;
  double resolved;
  int rc = resolve_bins(T, nu_total, nv_total, b, L, flags, &h, &resolved, st);
  if (rc != UQ_OK) return rc;
  UQ_REQUIRE(h.amb_u == nu_amb && h.amb_v == nv_amb, UQ_ERR_INVALID,
             "wasserstein_ambiguous: got %lld + %lld values, the tables say %lld + %lld",
             (long long)nu_amb, (long long)nv_amb, h.amb_u, h.amb_v);
  if (nu_amb > 0)
    UQ_CUDA(cudaMemcpyAsync(du, u_amb, sizeof(float) * (size_t)nu_amb, cudaMemcpyDeviceToDevice, st));
  if (nv_amb > 0)
    UQ_CUDA(cudaMemcpyAsync(dv, v_amb, sizeof(float) * (size_t)nv_amb, cudaMemcpyDeviceToDevice, st));
  return ambiguous_exact(T, flags, du, dut, nu_amb, dv, dvt, nv_amb, nu_total, nv_total, b, L,
                         out_host, st);
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
    rc = radix_sort_f32_copy(u, du, dut, nu, b + L.scratch, radix_sort_scratch_bytes(nu), &su, st);
    if (rc != UQ_OK) return rc;
  }
  if (nv > 0) {
    rc = radix_sort_f32_copy(v, dv, dvt, nv, b + L.scratch, radix_sort_scratch_bytes(nv), &sv, st);
    if (rc != UQ_OK) return rc;
  }
  if (nu + nv - 1 > 0) {
    const int64_t blocks = (nu + nv - 1 + MRG_TILE - 1) / MRG_TILE;
    const unsigned grid = integral_grid(blocks);
    cdf_integral_kernel<<<grid, MRG_THREADS, 0, st>>>(su, nu, sv, nv, u_below, v_below, nu_total,
                                                      nv_total, blocks, parts);
    UQ_LAUNCH_CHECK();
    sum_partials_kernel<<<1, 1024, 0, st>>>(parts, grid, result);
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
