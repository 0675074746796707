// Probe: where does tcgen05.mma.cta_group::2 with M = 128 (64 rows per CTA) put D in tensor memory?
// A[r][0] = r_global + 1, B[n][0] = n + 1 (other K columns zero)  =>  D[r][n] = (r + 1)(n + 1),
// which identifies (row, column) of every TMEM cell read back with tcgen05.ld.32x32b.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I nnueehcs_b200/csrc -o tools/_bin/tmem_layout tools/microbench/tmem_layout.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "tc_ptx.cuh"

using namespace uq::tc;

constexpr int N = 256;

__global__ void __launch_bounds__(128, 1) probe(float* out /* [2][128 lanes][N cols] */, int m_pair) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* a_smem = smem;                // [rows x 64] bf16 SW128 (16 KB reserved)
  uint8_t* b_smem = smem + 16384;        // [128 x 64] bf16 SW128
  uint8_t* bar = smem + 16384 + 16384;
  const uint32_t rank = cluster_ctarank();
  const int rows = m_pair / 2;           // A rows per CTA
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < 32768 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  // A: element (r, k=0) = global row + 1
  if (t < rows) {
    __nv_bfloat16 v = __float2bfloat16_rn((float)(rank * rows + t + 1));
    *reinterpret_cast<__nv_bfloat16*>(a_smem + sw128_offset(t, 0)) = v;
  }
  // B: this CTA holds N rows [rank * N/2, +N/2); element (n, k=0) = n + 1
  {
    const int n = rank * (N / 2) + t;
    __nv_bfloat16 v = __float2bfloat16_rn((float)(n + 1));
    *reinterpret_cast<__nv_bfloat16*>(b_smem + sw128_offset(t, 0)) = v;
  }
  fence_proxy_async_smem();
  if (t == 0) {
    mbar_init(smem_u32(bar), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(smem_u32(bar + 16), 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(bar + 16);
  if (rank == 0 && warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(m_pair, N);
      umma_bf16_pair(tmem_base, make_sw128_desc(smem_u32(a_smem)), make_sw128_desc(smem_u32(b_smem)),
                     idesc, 0u);
      umma_commit_pair(smem_u32(bar), 3);
    }
    __syncwarp();
  }
  if (t == 0) {
    uint32_t spins = 0;
    while (!mbar_try_wait(smem_u32(bar), 0)) {
      if (++spins > (1u << 22)) __trap();
    }
  }
  __syncthreads();
  tc_fence_after();
  // every warp dumps its lane quarter, all N columns (only N/2 may be meaningful)
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j)
      out[((size_t)rank * 128 + t) * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

int main() {
  float* d;
  cudaMalloc(&d, sizeof(float) * 2 * 128 * N);
  for (int m_pair : {256, 128}) {
    cudaMemset(d, 0xFF, sizeof(float) * 2 * 128 * N);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 40960;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe, d, m_pair);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M=%d: error %s\n", m_pair, cudaGetErrorString(e)); return 1; }
    std::vector<float> h(2 * 128 * N);
    cudaMemcpy(h.data(), d, sizeof(float) * h.size(), cudaMemcpyDeviceToHost);
    printf("== pair M = %d, N = %d: TMEM cell (cta, lane, col) -> decoded (row, n) of D ==\n", m_pair, N);
    for (int cta = 0; cta < 2; ++cta)
      for (int lane : {0, 1, 15, 16, 31, 32, 33, 63, 64, 65, 95, 96, 127})
        for (int col : {0, 1, 63, 64, 127, 128, 129, 255}) {
          const float v = h[((size_t)cta * 128 + lane) * N + col];
          // decode v = (r+1)(n+1): try all r
          int fr = -1, fn = -1, cnt = 0;
          for (int r = 0; r < m_pair; ++r) {
            const float q = v / (float)(r + 1);
            const int n = (int)q - 1;
            if (q == (float)(int)q && n >= 0 && n < N) {
              // ambiguous factorizations exist; prefer consistency check by neighbours later
              if (cnt == 0) { fr = r; fn = n; }
              ++cnt;
            }
          }
          printf("cta %d lane %3d col %3d : v = %10.1f  first (r=%d, n=%d) of %d factorizations\n", cta, lane, col, v, fr, fn, cnt);
        }
    // unambiguous decode using two cells: D[r][n+1] - D[r][n] = r + 1 along a row of columns
    printf("-- per lane: row index from column differences, and n of col 0 / col N/2 --\n");
    for (int cta = 0; cta < 2; ++cta)
      for (int lane = 0; lane < 128; lane += (lane % 32 == 0 ? 1 : 15)) {
        const float* row = &h[((size_t)cta * 128 + lane) * N];
        const float d01 = row[1] - row[0];
        const float dh = row[N / 2 + 1] - row[N / 2];
        printf("cta %d lane %3d: cols 0..: r+1 = %6.1f, n(col0)+1 = %6.1f | cols N/2..: r+1 = %6.1f, n+1 = %6.1f\n", cta,
               lane, d01, d01 != 0 ? row[0] / d01 : -1.f, dh, dh != 0 ? row[N / 2] / dh : -1.f);
      }
  }
  return 0;
}
