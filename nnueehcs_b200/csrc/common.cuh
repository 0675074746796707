// Shared declarations of the nnueehcs_b200 native library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/nnueehcs_b200.h"

namespace uq {

// ---- error plumbing: never abort, hand a status + message back through the C ABI ------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void count_launch(int n = 1);

#define UQ_CUDA(expr)                                                          \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) return uq::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define UQ_LAUNCH_CHECK()                                                       \
  do {                                                                          \
    uq::count_launch();                                                         \
    cudaError_t _e = cudaGetLastError();                                        \
    if (_e != cudaSuccess) return uq::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

#define UQ_REQUIRE(cond, code, ...)   \
  do {                                \
    if (!(cond)) {                    \
      uq::set_error(__VA_ARGS__);     \
      return (code);                  \
    }                                 \
  } while (0)

// One-time, PER-DEVICE opt-in of a kernel to more than 48 KB of dynamic shared memory
// (cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device function attribute: a process
// that uses cuda:0 and then cuda:1 must set it on both).  Thread-safe; devices >= 64 set it on
// every call.
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0};
};
template <class Kernel>
inline int smem_opt_in(Kernel kernel, int bytes, PerDeviceOnce& once) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
  if (bit && (once.done.load(std::memory_order_acquire) & bit)) return UQ_OK;
  UQ_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (bit) once.done.fetch_or(bit, std::memory_order_release);
  return UQ_OK;
}

// ---- single-launch metric kernels: grid barrier, co-resident grid size, mapped result slot ------
// A metric that used to be 4-8 dependent launches runs as ONE cooperative launch whose phases are
// separated by this barrier (cooperative launch guarantees co-residency, so spinning is safe).
// `counter` is zeroed before the launch; barrier number k (1, 2, ...) waits for k * gridDim.x
// arrivals.  Release on arrive / acquire on the poll make every block's earlier global writes and
// atomics visible to every block's later reads.
#ifdef __CUDACC__
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int k) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int target = k * gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
  }
  __syncthreads();
}
#endif

#ifdef __CUDACC__
// A block's grid-stride share of x[0 .. n), one call of add(float) per value: scalar head up to
// 16-byte alignment, float4 body with four loads in flight per thread, scalar tail.  (A software-
// pipelined variant -- the next batch's loads issued before the current batch is processed --
// and a TMA-fed shared-memory ring were both measured and were not faster: the metric passes
// are bound by shared-memory atomic wavefronts, see DESIGN.md.)
template <int THREADS, class F>
__device__ __forceinline__ void for_each_value(const float* __restrict__ x, int64_t n, F&& add) {
  const int64_t head = min(n, (int64_t)((16 - ((uintptr_t)x & 15)) & 15) / 4);
  const int64_t n4 = (n - head) / 4;
  const float4* x4 = reinterpret_cast<const float4*>(x + head);
  const int64_t gtid = (int64_t)blockIdx.x * THREADS + threadIdx.x;
  const int64_t gstride = (int64_t)gridDim.x * THREADS;
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
}
#endif

// Largest grid of `threads`-thread blocks with `smem` dynamic bytes that is co-resident on the
// current device (cached per device like the shared-memory opt-in).
struct PerDeviceInt {
  std::atomic<int> v[64];
  PerDeviceInt() { for (auto& a : v) a.store(0); }
};
template <class Kernel>
inline int coop_grid_limit(Kernel kernel, int threads, size_t smem, PerDeviceInt& cache, int* out) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64) {
    const int c = cache.v[dev].load(std::memory_order_acquire);
    if (c > 0) { *out = c; return UQ_OK; }
  }
  int per_sm = 0, sms = 0, coop = 0;
  UQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
  UQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  UQ_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  UQ_REQUIRE(coop && per_sm * sms >= 1, UQ_ERR_UNSUPPORTED,
             "device %d cannot run the single-launch metric kernel (cooperative launch %d, "
             "%d blocks/SM)", dev, coop, per_sm);
  *out = per_sm * sms;
  if (dev >= 0 && dev < 64) cache.v[dev].store(*out, std::memory_order_release);
  return UQ_OK;
}

// 256 bytes of mapped pinned host memory per (thread, device): a metric kernel's last block writes
// the scalar result record straight into it, so the call ends with a stream synchronisation
// instead of a D2H copy + synchronisation.  Lives until process exit.
int result_slot(void** host_ptr, void** dev_ptr);

// ---- packed model ------------------------------------------------------------------------------
struct Layer {
  int in = 0, out = 0;
  bool has_bn = false, relu = false, dropout = false;
  // fp32 parity path, all members stacked on the leading axis
  float* w = nullptr;      // [K][out][in]
  float* bias = nullptr;   // [K][out]   (zeros when the Linear has no bias)
  float* alpha = nullptr;  // [K][out]   eval-BN scale  gamma / sqrt(var + eps)   (has_bn)
  float* beta = nullptr;   // [K][out]   eval-BN shift  beta_bn - mean * alpha     (has_bn)
  // bf16 tcgen05 path (only when the model is tensor-core eligible)
  float* bias_folded = nullptr;  // [K][out]  (bias * alpha + beta), fp32
};

// Geometry of the bf16 fused kernel's packed weight image (see mlp_tc.cu).
struct TcPlan {
  bool ok = false;
  std::string why_not;
  int d_in = 0;        // network input features
  int k0 = 0;          // padded K of the first MMA layer (multiple of 16, <= 64)
  int hidden = 0;      // H: common width of all hidden layers (multiple of 64, 64..512)
  int n_mma_layers = 0;  // layers executed as MMAs: Linear 0 .. L-2
  int d_out = 0;       // final Linear out features (1..8): CUDA-core epilogue dot
  size_t stage_bytes = 0;        // bytes of one weight stage [n_tile x 64] bf16
  int n_tile = 0;                // MMA N per stage (H if H <= 256 else 256)
  int stages_per_member = 0;
  __nv_bfloat16* image = nullptr;  // [K][stages_per_member][stage_bytes]
  // bias stages of mlp_tc2.cu's bias-in-the-MMA variant: [K][n_mma_layers][H / n_tile] stages whose
  // K columns 0..2 hold the folded bias of the stage's N rows as bf16 hi + mid + lo
  __nv_bfloat16* bias_image = nullptr;
  float* w_last = nullptr;         // [K][d_out][H] fp32 (dropout scale NOT folded; see kernel)
  float* b_last = nullptr;         // [K][d_out]
  // Delta-UQ variant of the image (single network with an even input width): layer 0 holds only
  // the columns that multiply x; the anchor enters as a per-anchor bias (see tc_forward)
  __nv_bfloat16* image_delta = nullptr;
  int k0_delta = 0;
  // PAGER variant: W0 [a - x; x] + b0 = (W0b - W0a) x + (b0 + W0a a), so layer 0 holds the column
  // differences and the anchor again enters as a per-anchor bias
  __nv_bfloat16* image_pager = nullptr;
  // fp32-parity split mode (mlp_tcx.cu): fp16 hi/lo images of the power-of-two scaled weights and
  // the per-(member, layer) statistics the kernel's row scaling needs
  bool x_ok = false;
  std::string x_why_not;
  int x_n_tile = 0;                // MMA N per stage
  size_t x_stage_bytes = 0;        // [x_n_tile x 64] fp16
  int x_stages_per_member = 0;
  uint16_t* x_image = nullptr;        // [K][x_stages_per_member][x_stage_bytes]
  uint16_t* x_image_delta = nullptr;  // layer 0 = the columns that multiply x (Delta-UQ)
  uint16_t* x_image_pager = nullptr;  // layer 0 = column differences (PAGER)
  float* x_stats = nullptr;           // [K][n_mma_layers][4]
  float* x_stats_delta = nullptr;     // [1][n_mma_layers][4]
  float* x_stats_pager = nullptr;
};

}  // namespace uq

struct uq_model {
  int device = 0;
  int n_members = 0;
  int n_layers = 0;
  int d_in = 0, d_out = 0, max_width = 0;
  int n_dropout = 0;
  bool anchor_first = false;   // UQ_MODEL_ANCHOR_FIRST: anchored input = cat([a, x - a])
  std::vector<uq::Layer> layers;
  std::vector<void*> allocations;
  uq::TcPlan tc;
};

namespace uq {

// fp32 parity path (mlp_fp32.cu)
size_t fp32_workspace_bytes(const uq_model* m, int64_t n, const uq_forward_args* a);
int fp32_forward(const uq_model* m, const float* x, int64_t n, const uq_forward_args* a,
                 float* out0, float* out1, void* ws, size_t ws_bytes, cudaStream_t st);

// bf16 tcgen05 path (mlp_tc.cu)
void tc_plan(uq_model* m);  // fills m->tc.ok / geometry (no allocation)
int tc_pack(uq_model* m, cudaStream_t st);
// split = true: the fp32-parity split mode (mlp_tcx.cu) instead of the bf16 kernels
size_t tc_workspace_bytes(const uq_model* m, int64_t n, const uq_forward_args* a,
                          bool split = false);
int tc_forward(const uq_model* m, const float* x, int64_t n, const uq_forward_args* a,
               float* out0, float* out1, void* ws, size_t ws_bytes, cudaStream_t st,
               bool split = false);
// fp32-parity split mode, host side (mlp_tcx.cu)
void tcx_plan(uq_model* m);  // fills m->tc.x_ok (needs tc_plan first)
int tcx_pack(uq_model* m, cudaStream_t st);

// moments (moments.cu)
int moments_merge(const float* means, const float* m2s, const double* counts, int n_shards,
                  int64_t len, float* out_mean, float* out_std, cudaStream_t st);
int moments_merge_strided(const float* means, const float* m2s, int64_t shard_stride,
                          const double* counts, int n_shards, int64_t len, float* out_mean,
                          float* out_second, int moments, cudaStream_t st);

// metrics (wasserstein.cu / kde_jsd.cu)
size_t wasserstein_workspace_bytes(int64_t nu, int64_t nv);
int wasserstein_1d(const float* u, int64_t nu, const float* v, int64_t nv, int method,
                   double* out_host, int64_t* info_host, void* ws, size_t ws_bytes,
                   cudaStream_t st);
int bin_moments_accumulate(const float* x, int64_t n, unsigned long long* cnt,
                           unsigned long long* ksum, cudaStream_t st);
int wasserstein_from_bins(const unsigned long long* tables, int64_t nu_total, int64_t nv_total,
                          uint8_t* flags_out, double* out_host, void* ws, size_t ws_bytes,
                          cudaStream_t st);
int compact_flagged(const float* x, int64_t n, const uint8_t* flags, float* out,
                    int64_t* count_host, void* ws, size_t ws_bytes, cudaStream_t st);
int wasserstein_ambiguous(const float* u_amb, int64_t nu_amb, const float* v_amb, int64_t nv_amb,
                          const unsigned long long* tables, int64_t nu_total, int64_t nv_total,
                          double* out_host, void* ws, size_t ws_bytes, cudaStream_t st);
int wasserstein_1d_range(const float* u, int64_t nu, const float* v, int64_t nv, int64_t u_below,
                         int64_t v_below, int64_t nu_total, int64_t nv_total, double* out_host,
                         void* ws, size_t ws_bytes, cudaStream_t st);
int sample_stats_multi(const float* const* xs, const int64_t* ns, int count, double* out_host,
                       void* workspace, cudaStream_t st);
size_t kde_grid_workspace_bytes(int64_t n);
int kde_grid_accumulate(const float* x, int64_t n, double lo, double hi, double bandwidth,
                        int grid_pts, double* grid, void* ws, size_t ws_bytes, cudaStream_t st);
int jsd_from_grids(const double* grids, int grid_pts, double* out_host, cudaStream_t st);
size_t kde_density_workspace_bytes(int64_t n, int64_t m);
int kde_density(const float* fit, int64_t m, const float* x, int64_t n, int d, double bandwidth,
                double* out, void* ws, size_t ws_bytes, cudaStream_t st);
size_t kde_jsd_fused_workspace_bytes(int64_t nu, int64_t nv, int grid_pts);
int kde_jsd_fused(const float* u, int64_t nu, const float* v, int64_t nv, int grid_pts,
                  double* out_host, int* status_host, double* info_host, void* ws, size_t ws_bytes,
                  cudaStream_t st);
int kde_jsd_fused_phase_us(double* out5);
int kde_jsd_fused_enqueue(const float* u, int64_t nu, const float* v, int64_t nv, int grid_pts,
                          void* record, void* ws, size_t ws_bytes, cudaStream_t st);
int kde_jsd_fused_read(const void* record, double* out_host);
int kde_jsd_enqueue(const float* u, int64_t nu, const float* v, int64_t nv, int grid_pts,
                    void* record, void* ws, size_t ws_bytes, cudaStream_t st);
int kde_jsd_finish(const float* u, int64_t nu, const float* v, int64_t nv, int grid_pts,
                   const void* record, double* out_host, int* method_used_host, void* ws,
                   size_t ws_bytes, cudaStream_t st);
int wasserstein_1d_enqueue(const float* u, int64_t nu, const float* v, int64_t nv, void* record,
                           void* ws, size_t ws_bytes, cudaStream_t st);
int wasserstein_1d_finish(const float* u, int64_t nu, const float* v, int64_t nv,
                          const void* record, double* out_host, int64_t* info_host, void* ws,
                          size_t ws_bytes, cudaStream_t st);
size_t kde_jsd_workspace_bytes(int64_t nu, int64_t nv, int grid_pts);
int kde_jsd(const float* u, int64_t nu, const float* v, int64_t nv, int grid_pts, int method,
            double* out_host, int* method_used_host, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace uq
