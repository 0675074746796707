// Micro-benchmark: sustained cp.async.bulk (UBLKCP) global->shared ingest rate per SM on B200,
// as a function of stage size, ring depth and whether all CTAs stream the same addresses.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/bulk_bw tools/microbench/bulk_bw.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(128, 1)
bulk_bw(const uint8_t* src, size_t region, int stage_bytes, int n_stages, int iters, int distinct,
        int n_issuers, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint8_t* buf = smem + 1024;
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // n_issuers warps each own a disjoint subset of the stages (slot % n_issuers == warp)
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && warp < n_issuers) {
    const uint8_t* base = src + (distinct ? (size_t)blockIdx.x * region : 0);
    const long long t0 = clock64();
    for (int i = warp; i < iters + n_stages; i += n_issuers) {
      const int slot = i % n_stages;
      if (i >= n_stages) {
        const uint32_t par = ((i / n_stages) - 1) & 1;
        while (!try_wait(&bars[slot], par)) {}
      }
      if (i < iters) {
        const size_t off = ((size_t)i * stage_bytes) % (region - stage_bytes + 1);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[slot])), "r"(stage_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(buf + (size_t)slot * stage_bytes)), "l"(base + (off & ~(size_t)127)), "r"(stage_bytes), "r"(smem_u32(&bars[slot])) : "memory");
      }
    }
    if (warp == 0) out[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
}

int main() {
  const size_t region = 1 << 20;  // 1 MiB per CTA (or shared)
  uint8_t* src;
  cudaMalloc(&src, region * 160);
  cudaMemset(src, 1, region * 160);
  unsigned long long* out;
  cudaMalloc(&out, 160 * sizeof(unsigned long long));
  cudaFuncSetAttribute(bulk_bw, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const int iters = 4096;
  printf("stage_KB stages issuers distinct  B/clk/SM  GB/s(148 SMs, from events)\n");
  for (int distinct = 0; distinct <= 1; ++distinct)
    for (int issuers = 1; issuers <= 2; ++issuers)
      for (int kb : {4, 8, 16, 32})
        for (int stages : {2, 3, 4, 6, 8, 12}) {
          if ((size_t)kb * 1024 * stages + 1024 > 227 * 1024) continue;
          if (stages < issuers) continue;
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0); cudaEventCreate(&e1);
          const size_t sm = (size_t)kb * 1024 * stages + 1024;
          bulk_bw<<<148, 128, sm>>>(src, region, kb * 1024, stages, 64, distinct, issuers, out);  // warm L2
          cudaEventRecord(e0);
          bulk_bw<<<148, 128, sm>>>(src, region, kb * 1024, stages, iters, distinct, issuers, out);
          cudaEventRecord(e1);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          unsigned long long h[148];
          cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
          double cyc = 0; for (int b = 0; b < 148; ++b) cyc += (double)h[b] / 148;
          const double bytes = (double)iters * kb * 1024;
          printf("%7d %6d %7d %8d  %8.1f  %8.0f\n", kb, stages, issuers, distinct, bytes / cyc,
                 bytes * 148 / (ms * 1e-3) / 1e9);
        }
  return 0;
}
