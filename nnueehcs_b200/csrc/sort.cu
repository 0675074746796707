// Stable LSD radix sort (8-bit digits, 4 passes) of float32 keys: upsweep histogram ->
// single-block spine scan -> downsweep scatter.  Each block owns a contiguous key range, so the
// spine is only 256 x (#blocks) counters; inside a tile, ranks come from lane-private byte
// counters in shared memory (stable: items, lanes, warps, tiles and blocks are all visited in key
// order), the tile is staged in sorted order in shared memory, and the 4-byte stores of one
// digit land in a run that the 126 MB L2 merges before write-back.
#include <stdlib.h>

#include "common.cuh"
#include "sort.cuh"

namespace uq {
namespace {

constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int RADIX = 256;
constexpr int MAX_SORT_BLOCKS = 148 * 2;   // <= SPINE_THREADS: one spine row is one block scan

// order-preserving key of a float32 bit pattern; every NaN (either sign, any payload) becomes the
// largest key, so NaNs sort last as np.sort puts them (and come back as 0x7FFFFFFF)
__device__ __forceinline__ uint32_t f2key(uint32_t b) {
  if ((b & 0x7FFFFFFFu) > 0x7F800000u) return 0xFFFFFFFFu;
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ uint32_t key2f(uint32_t k) {
  return k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu);
}

template <bool FIRST>
__global__ void __launch_bounds__(SORT_THREADS)
upsweep_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift, int64_t keys_per_block,
               uint32_t* __restrict__ spine, int num_blocks, uint32_t* __restrict__ digit_totals) {
  __shared__ uint32_t wh[SORT_WARPS][RADIX];
  const int t = threadIdx.x, w = t >> 5;
  for (int i = t; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&wh[0][0])[i] = 0;
  __syncthreads();
  const int64_t begin = (int64_t)blockIdx.x * keys_per_block;
  int64_t end = begin + keys_per_block;
  if (end > n) end = n;
  // 128-bit loads over the aligned bulk of the range
  const int64_t vbeg = begin / 4, vend = end / 4;   // begin is a multiple of the tile size
  const uint4* k4 = reinterpret_cast<const uint4*>(keys);
  for (int64_t i = vbeg + t; i < vend; i += SORT_THREADS) {
    const uint4 v = k4[i];
    uint32_t a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint32_t k = FIRST ? f2key(a[e]) : a[e];
      atomicAdd(&wh[w][(k >> shift) & 0xFF], 1u);
    }
  }
  for (int64_t i = vend * 4 + t; i < end; i += SORT_THREADS) {
    const uint32_t k = FIRST ? f2key(keys[i]) : keys[i];
    atomicAdd(&wh[w][(k >> shift) & 0xFF], 1u);
  }
  __syncthreads();
  for (int d = t; d < RADIX; d += SORT_THREADS) {
    uint32_t s = 0;
#pragma unroll
    for (int ww = 0; ww < SORT_WARPS; ++ww) s += wh[ww][d];
    spine[(int64_t)d * num_blocks + blockIdx.x] = s;
    if (s) atomicAdd(digit_totals + d, s);   // whole-array count of digit d (zeroed per sort)
  }
}

// exclusive scan of the digit-major spine: block d scans digit d's row (one counter per upsweep
// block, at most SPINE_THREADS of them) and starts it at the number of keys with a smaller digit,
// which it gets from the 256 whole-array digit totals the upsweep accumulated.  (One block walking
// all 256 x 291 counters took 24 us per pass, a tenth of the downsweep.)
constexpr int SPINE_THREADS = 320;
static_assert(MAX_SORT_BLOCKS <= SPINE_THREADS && RADIX <= SPINE_THREADS, "one row, one block");
__global__ void __launch_bounds__(SPINE_THREADS)
spine_scan_kernel(uint32_t* __restrict__ spine, int num_blocks,
                  const uint32_t* __restrict__ digit_totals) {
  __shared__ uint32_t warp_sums[SPINE_THREADS / 32];
  __shared__ uint32_t below_s[SPINE_THREADS / 32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5, d = blockIdx.x;
  uint32_t below = (t < d) ? digit_totals[t] : 0u;   // d < RADIX <= SPINE_THREADS
  const uint32_t v = t < num_blocks ? spine[(int64_t)d * num_blocks + t] : 0u;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) below += __shfl_down_sync(0xffffffffu, below, o);
  if (lane == 31) warp_sums[w] = incl;
  if (lane == 0) below_s[w] = below;
  __syncthreads();
  uint32_t base = 0;
#pragma unroll
  for (int ww = 0; ww < SPINE_THREADS / 32; ++ww) {
    base += below_s[ww];
    if (ww < w) base += warp_sums[ww];
  }
  if (t < num_blocks) spine[(int64_t)d * num_blocks + t] = base + incl - v;
}

// ---- downsweep: lane-private byte counters --------------------------------------------------------
// A warp owns ITEMS * 32 consecutive keys of the tile and each LANE owns ITEMS consecutive keys of
// those, so "key order" is (warp, lane, item).  Ranking needs no warp collectives at all:
//   count   every lane bumps its own byte counter of the key's digit (the old value is the key's
//           rank among the lane's earlier keys with that digit; <= ITEMS <= 32).  The counters of
//           lane l live in shared-memory bank l -- word (digit >> 2) * 32 + l, byte digit & 3 -- so
//           the read-modify-write of 32 random digits is ONE wavefront each way (in a [digit][lane]
//           byte matrix it was ~2.8: the kernel is bound by the LSU data pipe, 82 % busy, and half
//           of its shared-memory wavefronts were bank conflicts);
//   scan 1  one lane per (digit row j = digit >> 2, group g of four lanes): the four words of the
//           group (one 128-bit load) are four lanes x four digits of byte counts; their running
//           byte-wise sum is the exclusive prefix inside the group (<= 96: no byte overflows),
//           stored back in place; the group's total goes to the row's group-total slot;
//   scan 1b one lane per row: the eight group totals of a row added up -> the warp's count of the
//           row's four digits, for the block-wide scan (thread t = digit t) that follows;
//   scan 2  one lane per row again: the eight group totals turned into 16-bit exclusive bases that
//           start at the block offset of (digit, warp), two digits per word: base[j][g] = (digits
//           4j and 4j+2 | digits 4j+1 and 4j+3);
//   rank    position in the tile = group base + byte prefix + rank in the lane: two shared loads
//           per key, one of them conflict-free.
// The ballot version of this kernel (nine VOTEs and ~60 instructions per 32 keys and round, every
// round waiting for the previous one's counter update) issued 106 warp instructions per 32 keys
// and ran at 0.39 keys per clock and SM.
// Base rows are 64 bytes of data in 80-byte slots: the stride of 20 words makes the 128-bit row
// accesses of eight consecutive lanes hit eight different bank quads.  A row's group totals sit in
// the last 32 bytes of its own slot (16 of them overlap the data: scan 2 has the totals in
// registers before it writes the bases).
template <int ITEMS>
struct Down {
  static constexpr int TILE = SORT_THREADS * ITEMS;
  static constexpr int CNT_BYTES = SORT_WARPS * 8192;            // aliased by the staged tile
  static constexpr int BROW = 80;                                // bytes per base row slot
  static constexpr int BASE_BYTES = 64 * BROW;                   // per warp
  static constexpr int BASE_OFF = CNT_BYTES;
  static constexpr int WOFF_OFF = BASE_OFF + SORT_WARPS * BASE_BYTES;  // u16 [warp][256]
  static constexpr int DBASE_OFF = WOFF_OFF + SORT_WARPS * 512;  // u32 [256]
  static constexpr int GDELTA_OFF = DBASE_OFF + 1024;            // u32 [256]
  static constexpr int WTOT_OFF = GDELTA_OFF + 1024;             // u32 [warps]
  static constexpr int SMEM = WTOT_OFF + SORT_WARPS * 4;
  static_assert(TILE * 4 <= CNT_BYTES, "staged tile must fit the counter area");
  static_assert(ITEMS % 4 == 0 && ITEMS <= 32, "byte counters: at most 32 keys per lane");
  static_assert(2 * (SMEM + 1024) <= 233472, "two blocks per SM");
};

// One tile.  FULL = every key of the tile exists (all tiles but the last one of the array): no
// per-key validity tests anywhere.
template <int ITEMS, bool FIRST, bool LAST, bool FULL>
__device__ __forceinline__ void downsweep_tile(const uint32_t* __restrict__ in,
                                               uint32_t* __restrict__ out, int64_t tile0,
                                               int64_t end, int shift, unsigned char* smem) {
  using D = Down<ITEMS>;
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  unsigned char* cnt_w = smem + w * 8192;
  uint32_t* staged = reinterpret_cast<uint32_t*>(smem);
  unsigned char* base_w = smem + D::BASE_OFF + w * D::BASE_BYTES;
  unsigned short* woff = reinterpret_cast<unsigned short*>(smem + D::WOFF_OFF);
  unsigned short* woff_w = woff + w * RADIX;
  uint32_t* digit_base = reinterpret_cast<uint32_t*>(smem + D::DBASE_OFF);
  uint32_t* gdelta = reinterpret_cast<uint32_t*>(smem + D::GDELTA_OFF);
  uint32_t* warp_tot = reinterpret_cast<uint32_t*>(smem + D::WTOT_OFF);
  // counter byte of (digit d, this lane): cnt_w + (d >> 2) * 128 + (d & 3) + lane * 4
  const uint32_t lane_off = (uint32_t)lane * 4u;
  // 16-bit group base of (digit d, this lane's group): base_w + (d >> 2) * BROW + (d & 1) * 4 +
  // ((d >> 1) & 1) * 2 + (lane >> 2) * 8
  const uint32_t group_off = (uint32_t)(lane >> 2) * 8u;

  // ---- keys: lane l owns keys [l * ITEMS, (l + 1) * ITEMS) of the warp's run.  Loading them
  // straight from global memory (128-bit loads ITEMS * 4 bytes apart) costs one L1 wavefront per
  // lane and load -- a quarter of all the LSU wavefronts of this LSU-bound kernel -- so the warp
  // loads its run coalesced (4 wavefronts per load) and transposes it through its own counter
  // area: rows of ITEMS keys padded by 16 bytes, so both the coalesced stores and the per-lane
  // 128-bit loads are conflict-free.
  uint32_t key[ITEMS];
  const int64_t wbase = tile0 + (int64_t)w * (32 * ITEMS);
  const int64_t lbase = wbase + (int64_t)lane * ITEMS;
  int nvalid = ITEMS;
  if (FULL) {
    constexpr int ROW = ITEMS * 4 + 16;
    const uint4* src = reinterpret_cast<const uint4*>(in + wbase);
    uint4 v[ITEMS / 4];
#pragma unroll
    for (int r = 0; r < ITEMS / 4; ++r) v[r] = __ldg(src + r * 32 + lane);
#pragma unroll
    for (int r = 0; r < ITEMS / 4; ++r) {
      const int e = 128 * r + 4 * lane;            // first of the four keys, within the warp's run
      const int o = e / ITEMS;                     // owner lane
      *reinterpret_cast<uint4*>(cnt_w + o * ROW + (e - o * ITEMS) * 4) = v[r];
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < ITEMS / 4; ++q) {
      const uint4 k4 = *reinterpret_cast<const uint4*>(cnt_w + lane * ROW + q * 16);
      key[4 * q + 0] = k4.x, key[4 * q + 1] = k4.y, key[4 * q + 2] = k4.z, key[4 * q + 3] = k4.w;
    }
    __syncwarp();
  } else {
    const int64_t left = end - lbase;
    nvalid = left >= ITEMS ? ITEMS : (left > 0 ? (int)left : 0);
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) key[r] = r < nvalid ? __ldg(in + lbase + r) : 0u;
  }
  if (FIRST) {
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) key[r] = f2key(key[r]);
  }
  // ---- zero this warp's counters (conflict-free 128-bit stores)
  {
    uint4* z = reinterpret_cast<uint4*>(cnt_w);
#pragma unroll
    for (int i = 0; i < 16; ++i) z[lane + 32 * i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncwarp();
  // ---- count: rk = rank among the lane's earlier keys with the same digit
  uint32_t rk[ITEMS];
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const uint32_t d = (key[r] >> shift) & 0xFFu;
    rk[r] = 0;
    if (FULL || r < nvalid) {
      unsigned char* c = cnt_w + ((d >> 2) * 128u + (d & 3u) + lane_off);
      const uint32_t old = *c;
      *c = (unsigned char)(old + 1u);
      rk[r] = old;
    }
  }
  __syncwarp();
  // ---- scan 1: exclusive prefix inside every group of four lanes, group totals
#pragma unroll
  for (int it = 0; it < 16; ++it) {
    const int item = lane + 32 * it;               // (row j, group g) = (item >> 3, item & 7)
    uint4* grp = reinterpret_cast<uint4*>(cnt_w) + item;   // words 4g .. 4g+3 of row j
    const uint4 c = *grp;
    const uint32_t e2 = c.x + c.y, e3 = e2 + c.z;
    *grp = make_uint4(0u, c.x, e2, e3);
    *reinterpret_cast<uint32_t*>(base_w + (item >> 3) * D::BROW + 48 + (item & 7) * 4) = e3 + c.w;
  }
  __syncwarp();
  // ---- scan 1b: this warp's count of every digit
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int j = lane + 32 * i;
    const uint4 ga = *reinterpret_cast<const uint4*>(base_w + j * D::BROW + 48);
    const uint4 gb = *reinterpret_cast<const uint4*>(base_w + j * D::BROW + 64);
    const uint32_t g8[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
    uint32_t lo = 0, hi = 0;                       // digits (4j, 4j+2) and (4j+1, 4j+3), 16 bits each
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      lo += g8[g] & 0x00FF00FFu;
      hi += (g8[g] >> 8) & 0x00FF00FFu;
    }
    // woff_w[4j .. 4j+3] = counts of digits 4j, 4j+1, 4j+2, 4j+3
    *reinterpret_cast<uint2*>(woff_w + 4 * j) =
        make_uint2((lo & 0xFFFFu) | (hi << 16), (lo >> 16) | (hi & 0xFFFF0000u));
  }
  __syncthreads();
  // ---- thread t == digit t: where each (digit, warp) run starts in the tile and in `out`
  {
    uint32_t c[SORT_WARPS];
    uint32_t total = 0;
#pragma unroll
    for (int ww = 0; ww < SORT_WARPS; ++ww) {
      c[ww] = woff[ww * RADIX + t];
      total += c[ww];
    }
    uint32_t incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (int ww = 0; ww < SORT_WARPS; ++ww)
      if (ww < w) before += warp_tot[ww];
    const uint32_t off = before + incl - total;   // first tile-local position of digit t
    uint32_t run = off;
#pragma unroll
    for (int ww = 0; ww < SORT_WARPS; ++ww) {
      woff[ww * RADIX + t] = (unsigned short)run;
      run += c[ww];
    }
    gdelta[t] = digit_base[t] - off;              // global position = gdelta[d] + local position
    digit_base[t] += total;
  }
  __syncthreads();
  // ---- scan 2: 16-bit group bases that start at the (digit, warp) run's offset in the tile
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int j = lane + 32 * i;
    unsigned char* row = base_w + j * D::BROW;
    const uint4 ga = *reinterpret_cast<const uint4*>(row + 48);
    const uint4 gb = *reinterpret_cast<const uint4*>(row + 64);
    const uint2 st = *reinterpret_cast<const uint2*>(woff_w + 4 * j);   // starts of digits 4j..4j+3
    uint32_t lo = (st.x & 0xFFFFu) | (st.y << 16);          // digits 4j, 4j+2
    uint32_t hi = (st.x >> 16) | (st.y & 0xFFFF0000u);      // digits 4j+1, 4j+3
    const uint32_t g8[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
    uint32_t b[16];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      b[2 * g] = lo, b[2 * g + 1] = hi;
      lo += g8[g] & 0x00FF00FFu;
      hi += (g8[g] >> 8) & 0x00FF00FFu;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<uint4*>(row + 16 * q) = make_uint4(b[4 * q], b[4 * q + 1], b[4 * q + 2], b[4 * q + 3]);
  }
  __syncwarp();
  // ---- rank: tile-local destination of every key
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const uint32_t d = (key[r] >> shift) & 0xFFu;
    const uint32_t j = d >> 2;
    rk[r] += (uint32_t)cnt_w[j * 128u + (d & 3u) + lane_off] +
             (uint32_t)*reinterpret_cast<const unsigned short*>(
                 base_w + j * D::BROW + (d & 1u) * 4u + (d & 2u) + group_off);
  }
  __syncthreads();   // every warp is done with the counters: the staged tile takes their place
#pragma unroll
  for (int r = 0; r < ITEMS; ++r)
    if (FULL || r < nvalid) staged[rk[r]] = key[r];
  __syncthreads();
  // ---- consecutive threads write consecutive addresses of a digit's run
  if (FULL) {
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
      const int i = t + r * SORT_THREADS;
      const uint32_t k = staged[i];
      const uint32_t d = (k >> shift) & 0xFFu;
      out[gdelta[d] + (uint32_t)i] = LAST ? key2f(k) : k;
    }
  } else {
    const int tile_n = (int)(end - tile0);
    for (int i = t; i < tile_n; i += SORT_THREADS) {
      const uint32_t k = staged[i];
      const uint32_t d = (k >> shift) & 0xFFu;
      out[gdelta[d] + (uint32_t)i] = LAST ? key2f(k) : k;
    }
  }
  __syncthreads();
}

template <int ITEMS, bool FIRST, bool LAST>
__global__ void __launch_bounds__(SORT_THREADS, 2)
downsweep_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n, int shift,
                 int64_t tiles_per_block, const uint32_t* __restrict__ spine, int num_blocks) {
  using D = Down<ITEMS>;
  extern __shared__ __align__(16) unsigned char smem[];
  uint32_t* digit_base = reinterpret_cast<uint32_t*>(smem + D::DBASE_OFF);
  // SORT_THREADS == RADIX: thread t owns digit t's running output position
  digit_base[threadIdx.x] = spine[(int64_t)threadIdx.x * num_blocks + blockIdx.x];
  const int64_t begin = (int64_t)blockIdx.x * tiles_per_block * D::TILE;
  int64_t end = begin + tiles_per_block * D::TILE;
  if (end > n) end = n;
  for (int64_t tile0 = begin; tile0 < end; tile0 += D::TILE) {
    if (tile0 + D::TILE <= end)
      downsweep_tile<ITEMS, FIRST, LAST, true>(in, out, tile0, end, shift, smem);
    else
      downsweep_tile<ITEMS, FIRST, LAST, false>(in, out, tile0, end, shift, smem);
  }
}

struct SortPlan {
  int items;                 // keys per lane and tile: 16 or 32
  int num_blocks;
  int64_t tiles_per_block;
  int64_t tile() const { return (int64_t)SORT_THREADS * items; }
};

// 32 keys per lane (8192-key tiles: the row scan is paid once per 1024 keys of a warp and a
// digit's run in the output is twice as long) unless the array is too short to give every SM two
// blocks of them; UQ_SORT_ITEMS=16|32 overrides (A/B runs).
SortPlan sort_plan(int64_t n) {
  static const int forced = [] {
    const char* e = getenv("UQ_SORT_ITEMS");
    const int v = e ? atoi(e) : 0;
    return (v == 16 || v == 32) ? v : 0;
  }();
  SortPlan p;
  p.items = forced ? forced : (n >= (int64_t)148 * 2 * SORT_THREADS * 32 ? 32 : 16);
  const int64_t tiles = (n + p.tile() - 1) / p.tile();
  int64_t nb = tiles < MAX_SORT_BLOCKS ? tiles : MAX_SORT_BLOCKS;   // two resident blocks per SM, one wave
  if (nb < 1) nb = 1;
  p.tiles_per_block = (tiles + nb - 1) / nb;
  p.num_blocks = (int)((tiles + p.tiles_per_block - 1) / p.tiles_per_block);
  if (p.num_blocks < 1) p.num_blocks = 1;
  return p;
}

// pass 0: src -> b, then b -> a, a -> b, b -> a: the result is in `a`; src may be `a` itself
template <int ITEMS>
int sort_passes(const uint32_t* src, uint32_t* a, uint32_t* b, int64_t n, const SortPlan& p,
                uint32_t* spine, uint32_t* digit_totals, cudaStream_t st) {
  static PerDeviceOnce opted[3];
  constexpr int SMEM = Down<ITEMS>::SMEM;
  if (int rc = smem_opt_in(downsweep_kernel<ITEMS, true, false>, SMEM, opted[0])) return rc;
  if (int rc = smem_opt_in(downsweep_kernel<ITEMS, false, false>, SMEM, opted[1])) return rc;
  if (int rc = smem_opt_in(downsweep_kernel<ITEMS, false, true>, SMEM, opted[2])) return rc;
  const int64_t per_block = p.tiles_per_block * p.tile();
  UQ_CUDA(cudaMemsetAsync(digit_totals, 0, 4 * RADIX * sizeof(uint32_t), st));
  const uint32_t* in = src;
  uint32_t* out = b;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = pass * 8;
    uint32_t* totals = digit_totals + pass * RADIX;
    if (pass == 0)
      upsweep_kernel<true><<<p.num_blocks, SORT_THREADS, 0, st>>>(in, n, shift, per_block, spine,
                                                                p.num_blocks, totals);
    else
      upsweep_kernel<false><<<p.num_blocks, SORT_THREADS, 0, st>>>(in, n, shift, per_block, spine,
                                                                 p.num_blocks, totals);
    UQ_LAUNCH_CHECK();
    spine_scan_kernel<<<RADIX, SPINE_THREADS, 0, st>>>(spine, p.num_blocks, totals);
    UQ_LAUNCH_CHECK();
    if (pass == 0)
      downsweep_kernel<ITEMS, true, false><<<p.num_blocks, SORT_THREADS, SMEM, st>>>(
          in, out, n, shift, p.tiles_per_block, spine, p.num_blocks);
    else if (pass == 3)
      downsweep_kernel<ITEMS, false, true><<<p.num_blocks, SORT_THREADS, SMEM, st>>>(
          in, out, n, shift, p.tiles_per_block, spine, p.num_blocks);
    else
      downsweep_kernel<ITEMS, false, false><<<p.num_blocks, SORT_THREADS, SMEM, st>>>(
          in, out, n, shift, p.tiles_per_block, spine, p.num_blocks);
    UQ_LAUNCH_CHECK();
    in = out;
    out = (out == b) ? a : b;
  }
  return UQ_OK;
}

}  // namespace

namespace {
// spine for either tile size (the plan can be overridden at run time)
size_t spine_bytes(int64_t n) {
  const int64_t tiles = (n + SORT_THREADS * 16 - 1) / (SORT_THREADS * 16);
  const int64_t nb = tiles < MAX_SORT_BLOCKS ? (tiles < 1 ? 1 : tiles) : MAX_SORT_BLOCKS;
  return (((size_t)RADIX * (size_t)nb * sizeof(uint32_t)) + 255) & ~(size_t)255;
}
}  // namespace

size_t radix_sort_scratch_bytes(int64_t n) {
  return spine_bytes(n) + 4 * RADIX * sizeof(uint32_t);   // + the digit totals of the four passes
}

int radix_sort_f32_copy(const float* src, float* keys, float* tmp, int64_t n, void* scratch,
                        size_t scratch_bytes, float** sorted, cudaStream_t st) {
  UQ_REQUIRE(n >= 1 && n < ((int64_t)1 << 31), UQ_ERR_INVALID,
             "radix sort: n = %lld outside [1, 2^31)", (long long)n);
  UQ_REQUIRE(scratch && scratch_bytes >= radix_sort_scratch_bytes(n), UQ_ERR_WORKSPACE,
             "radix sort: scratch too small");
  UQ_REQUIRE(((uintptr_t)keys & 15) == 0 && ((uintptr_t)tmp & 15) == 0, UQ_ERR_INVALID,
             "radix sort: key buffers must be 16-byte aligned");
  if (src != keys && ((uintptr_t)src & 15) != 0) {   // e.g. a tensor slice: copy, then in place
    UQ_CUDA(cudaMemcpyAsync(keys, src, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    src = keys;
  }
  const SortPlan p = sort_plan(n);
  uint32_t* spine = static_cast<uint32_t*>(scratch);
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
  uint32_t* a = reinterpret_cast<uint32_t*>(keys);
  uint32_t* b = reinterpret_cast<uint32_t*>(tmp);
  uint32_t* totals = reinterpret_cast<uint32_t*>(static_cast<char*>(scratch) + spine_bytes(n));
  const int rc = p.items == 32 ? sort_passes<32>(s32, a, b, n, p, spine, totals, st)
                               : sort_passes<16>(s32, a, b, n, p, spine, totals, st);
  if (rc != UQ_OK) return rc;
  *sorted = keys;  // 4 passes: the result is in `keys`
  return UQ_OK;
}

int radix_sort_f32(float* keys, float* tmp, int64_t n, void* scratch, size_t scratch_bytes,
                   float** sorted, cudaStream_t st) {
  return radix_sort_f32_copy(keys, keys, tmp, n, scratch, scratch_bytes, sorted, st);
}

}  // namespace uq

// ---- C ABI: the sort on its own (tests, tools/bench_metrics.py) -------------------------------------
namespace {
size_t sort_al(size_t b) { return (b + 255) & ~(size_t)255; }
}  // namespace

size_t uq_sort_workspace_bytes(int64_t n) {
  if (n < 1) return 0;
  return sort_al(sizeof(float) * (size_t)n) + sort_al(uq::radix_sort_scratch_bytes(n));
}

int uq_sort_f32(const float* x, int64_t n, float* out, void* workspace, size_t workspace_bytes,
                void* stream) {
  UQ_REQUIRE(x && out && n >= 1, UQ_ERR_INVALID, "uq_sort_f32: NULL argument or n < 1");
  UQ_REQUIRE(workspace && workspace_bytes >= uq_sort_workspace_bytes(n), UQ_ERR_WORKSPACE,
             "uq_sort_f32: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* w = static_cast<char*>(workspace);
  float* tmp = reinterpret_cast<float*>(w);
  float* sorted = nullptr;
  const int rc = uq::radix_sort_f32_copy(x, out, tmp, n, w + sort_al(sizeof(float) * (size_t)n),
                                         uq::radix_sort_scratch_bytes(n), &sorted, st);
  if (rc != UQ_OK) return rc;
  if (sorted != out)
    UQ_CUDA(cudaMemcpyAsync(out, sorted, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, st));
  return UQ_OK;
}
