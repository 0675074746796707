// Stable LSD radix sort (8-bit digits, 4 passes) of float32 keys: upsweep histogram ->
// single-block spine scan -> downsweep scatter.  Each block owns a contiguous key range, so the
// spine is only 256 x (#blocks) counters; inside a tile, ranks come from warp-wide
// __match_any_sync multisplit (stable: lanes, rounds, warps, tiles and blocks are all visited in
// key order), loads are fully coalesced and the scattered 4-byte stores of one digit land in a
// run that the 126 MB L2 merges before write-back.
#include "common.cuh"
#include "sort.cuh"

namespace uq {
namespace {

constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 4096 keys
constexpr int RADIX = 256;

__device__ __forceinline__ uint32_t f2key(uint32_t b) {
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ uint32_t key2f(uint32_t k) {
  return k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu);
}

template <bool FIRST>
__global__ void __launch_bounds__(SORT_THREADS)
upsweep_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift, int64_t tiles_per_block,
               uint32_t* __restrict__ spine, int num_blocks) {
  __shared__ uint32_t wh[SORT_WARPS][RADIX];
  const int t = threadIdx.x, w = t >> 5;
  for (int i = t; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&wh[0][0])[i] = 0;
  __syncthreads();
  const int64_t begin = (int64_t)blockIdx.x * tiles_per_block * SORT_TILE;
  int64_t end = begin + tiles_per_block * SORT_TILE;
  if (end > n) end = n;
  // 128-bit loads over the aligned bulk of the range
  const int64_t vbeg = begin / 4, vend = end / 4;   // begin is a multiple of 4096
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
  }
}

// exclusive scan of the digit-major spine, one block of 1024 threads walking it in coalesced
// 4096-element slabs (uint4 per thread) with a running carry
__global__ void __launch_bounds__(1024) spine_scan_kernel(uint32_t* __restrict__ spine, int total) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < total; base += 4096) {
    const uint32_t carry = carry_s;   // written by warp 0 between the two barriers below
    const int i0 = base + t * 4;
    uint32_t v[4] = {0, 0, 0, 0};
    if (i0 + 3 < total) {
      const uint4 q = *reinterpret_cast<const uint4*>(spine + i0);
      v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (i0 + e < total) v[e] = spine[i0 + e];
    }
    const uint32_t s = v[0] + v[1] + v[2] + v[3];
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
      const uint32_t ws = warp_sums[lane];
      uint32_t wi = ws;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += y;
      }
      warp_sums[lane] = wi - ws;
      if (lane == 31) carry_s = carry + wi;   // read by everyone only after the next barrier
    }
    __syncthreads();
    uint32_t run = carry + warp_sums[w] + incl - s;
    uint32_t o4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      o4[e] = run;
      run += v[e];
    }
    if (i0 + 3 < total) {
      *reinterpret_cast<uint4*>(spine + i0) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (i0 + e < total) spine[i0 + e] = o4[e];
    }
    __syncthreads();   // warp_sums / carry_s are rewritten by the next slab
  }
}

// 4 blocks per SM (<= 64 registers): the ranking rounds are a chain of dependent warp-collectives
// and shared-memory updates, so the kernel is latency-bound and wants warps, not registers
template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(SORT_THREADS, 4)
downsweep_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n, int shift,
                 int64_t tiles_per_block, const uint32_t* __restrict__ spine, int num_blocks) {
  __shared__ uint32_t digit_base[RADIX];   // running global position of each digit's next key
  __shared__ uint32_t tile_off[RADIX];     // first tile-local position of each digit
  __shared__ uint32_t gdelta[RADIX];
  __shared__ uint32_t warp_tot[SORT_WARPS];
  __shared__ uint32_t wh[SORT_WARPS][RADIX];
  __shared__ uint32_t staged[SORT_TILE];   // the tile in sorted order
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int d = t; d < RADIX; d += SORT_THREADS)
    digit_base[d] = spine[(int64_t)d * num_blocks + blockIdx.x];
  const int64_t begin = (int64_t)blockIdx.x * tiles_per_block * SORT_TILE;
  int64_t end = begin + tiles_per_block * SORT_TILE;
  if (end > n) end = n;

  for (int64_t tile0 = begin; tile0 < end; tile0 += SORT_TILE) {
    for (int i = t; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&wh[0][0])[i] = 0;
    __syncthreads();
    uint32_t key[SORT_ITEMS];
    uint32_t rank[SORT_ITEMS];
    const int64_t wbase = tile0 + (int64_t)w * 32 * SORT_ITEMS;
    // all 16 loads of the tile are issued before any of them is consumed: inside the ranking loop
    // the warp-synchronous steps would otherwise serialise one HBM round trip per round
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
      const int64_t idx = wbase + r * 32 + lane;
      uint32_t k = 0xFFFFFFFFu;
      if (idx < end) k = FIRST ? f2key(__ldg(in + idx)) : __ldg(in + idx);
      key[r] = k;
    }
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
      const int64_t idx = wbase + r * 32 + lane;
      const bool valid = idx < end;
      const uint32_t k = key[r];
      // lanes holding the same digit, from nine ballots (one per digit bit + validity).
      // MATCH.ANY does the same in one instruction but runs on the ADU pipe at ~64 cycles per
      // warp on sm_100 (ncu: ADU 57 % busy, kernel 10x off the HBM roofline); ballots and LOP3s
      // issue at full rate.
      const uint32_t d = (k >> shift) & 0xFFu;
      uint32_t peers = __ballot_sync(0xffffffffu, valid);
      if (!valid) peers = ~peers;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
      }
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (valid && lane == leader) {
        old = wh[w][d];
        wh[w][d] = old + __popc(peers);
      }
      old = __shfl_sync(0xffffffffu, old, leader);
      rank[r] = old + __popc(peers & lt_mask);
      __syncwarp();
    }
    __syncthreads();
    // thread t == digit t: offsets of the warps inside the digit's run, the digit's run inside the
    // tile (block-wide exclusive scan), and where that run goes in global memory
    {
      uint32_t run = 0;
#pragma unroll
      for (int ww = 0; ww < SORT_WARPS; ++ww) {
        const uint32_t c = wh[ww][t];
        wh[ww][t] = run;
        run += c;
      }
      uint32_t incl = run;   // digits-in-tile inclusive scan
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
      const uint32_t off = before + incl - run;   // first tile-local position of digit t
      tile_off[t] = off;
      gdelta[t] = digit_base[t] - off;            // global position = gdelta[d] + local position
      digit_base[t] += run;
    }
    __syncthreads();
    // keys go to their tile-local sorted position in shared memory first ...
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
      const int64_t idx = wbase + r * 32 + lane;
      if (idx < end) {
        const uint32_t d = (key[r] >> shift) & 0xFFu;
        staged[tile_off[d] + wh[w][d] + rank[r]] = key[r];
      }
    }
    __syncthreads();
    // ... so that consecutive threads write consecutive addresses of a digit's run: a warp's store
    // touches a few sectors instead of 32 (the direct scatter was LSU-sector bound)
    const int tile_n = (int)((end - tile0) < (int64_t)SORT_TILE ? (end - tile0) : (int64_t)SORT_TILE);
#pragma unroll 4
    for (int i = t; i < tile_n; i += SORT_THREADS) {
      const uint32_t k = staged[i];
      const uint32_t d = (k >> shift) & 0xFFu;
      out[gdelta[d] + (uint32_t)i] = LAST ? key2f(k) : k;
    }
    __syncthreads();
  }
}

struct SortPlan {
  int num_blocks;
  int64_t tiles_per_block;
};

SortPlan sort_plan(int64_t n) {
  const int64_t tiles = (n + SORT_TILE - 1) / SORT_TILE;
  int64_t nb = tiles < 148 * 4 ? tiles : 148 * 4;
  if (nb < 1) nb = 1;
  SortPlan p;
  p.tiles_per_block = (tiles + nb - 1) / nb;
  p.num_blocks = (int)((tiles + p.tiles_per_block - 1) / p.tiles_per_block);
  if (p.num_blocks < 1) p.num_blocks = 1;
  return p;
}

}  // namespace

size_t radix_sort_scratch_bytes(int64_t n) {
  const SortPlan p = sort_plan(n);
  return (((size_t)RADIX * p.num_blocks * sizeof(uint32_t)) + 255) & ~(size_t)255;
}

int radix_sort_f32(float* keys, float* tmp, int64_t n, void* scratch, size_t scratch_bytes,
                   float** sorted, cudaStream_t st) {
  UQ_REQUIRE(n >= 1 && n < ((int64_t)1 << 31), UQ_ERR_INVALID,
             "radix sort: n = %lld outside [1, 2^31)", (long long)n);
  UQ_REQUIRE(scratch && scratch_bytes >= radix_sort_scratch_bytes(n), UQ_ERR_WORKSPACE,
             "radix sort: scratch too small");
  const SortPlan p = sort_plan(n);
  uint32_t* spine = static_cast<uint32_t*>(scratch);
  uint32_t* a = reinterpret_cast<uint32_t*>(keys);
  uint32_t* b = reinterpret_cast<uint32_t*>(tmp);
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = pass * 8;
    if (pass == 0)
      upsweep_kernel<true><<<p.num_blocks, SORT_THREADS, 0, st>>>(a, n, shift, p.tiles_per_block,
                                                                spine, p.num_blocks);
    else
      upsweep_kernel<false><<<p.num_blocks, SORT_THREADS, 0, st>>>(a, n, shift, p.tiles_per_block,
                                                                 spine, p.num_blocks);
    UQ_LAUNCH_CHECK();
    spine_scan_kernel<<<1, 1024, 0, st>>>(spine, RADIX * p.num_blocks);
    UQ_LAUNCH_CHECK();
    if (pass == 0)
      downsweep_kernel<true, false><<<p.num_blocks, SORT_THREADS, 0, st>>>(
          a, b, n, shift, p.tiles_per_block, spine, p.num_blocks);
    else if (pass == 3)
      downsweep_kernel<false, true><<<p.num_blocks, SORT_THREADS, 0, st>>>(
          a, b, n, shift, p.tiles_per_block, spine, p.num_blocks);
    else
      downsweep_kernel<false, false><<<p.num_blocks, SORT_THREADS, 0, st>>>(
          a, b, n, shift, p.tiles_per_block, spine, p.num_blocks);
    UQ_LAUNCH_CHECK();
    uint32_t* s = a;
    a = b;
    b = s;
  }
  *sorted = reinterpret_cast<float*>(a);  // 4 passes: result is back in `keys`
  return UQ_OK;
}

}  // namespace uq
