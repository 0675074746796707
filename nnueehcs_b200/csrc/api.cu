// C ABI of nnueehcs_b200 (include/nnueehcs_b200.h): model packing, forward dispatch, host-buffer
// end-to-end call.  No torch types; the Python side binds this with ctypes.
#include <stdarg.h>
#include <string.h>

#include <math.h>

#include "common.cuh"
#include "philox.cuh"

namespace uq {

static thread_local char g_err[1024] = "";
static thread_local uint64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return UQ_ERR_CUDA;
}

void count_launch(int n) { g_launches += (uint64_t)n; }

int result_slot(void** host_ptr, void** dev_ptr) {
  static thread_local void* slots[64] = {nullptr};
  static thread_local void* dslots[64] = {nullptr};
  int dev = 0;
  cudaGetDevice(&dev);
  UQ_REQUIRE(dev >= 0 && dev < 64, UQ_ERR_UNSUPPORTED, "result_slot: device %d >= 64", dev);
  if (!slots[dev]) {
    void* p = nullptr;
    UQ_CUDA(cudaHostAlloc(&p, 256, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(p, 0, 256);
    void* d = nullptr;
    UQ_CUDA(cudaHostGetDevicePointer(&d, p, 0));
    slots[dev] = p;
    dslots[dev] = d;
  }
  *host_ptr = slots[dev];
  *dev_ptr = dslots[dev];
  return UQ_OK;
}

namespace {

// eval-mode BatchNorm1d as ATen's CPU kernel factors it: alpha = weight / sqrt(var + eps),
// beta = bias - mean * alpha; out = in * alpha + beta.
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ bnb,
                               const float* __restrict__ mean, const float* __restrict__ var,
                               float eps, float* __restrict__ alpha, float* __restrict__ beta,
                               int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float invstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var[i], eps)));
  const float g = gamma ? gamma[i] : 1.0f;
  const float a = __fmul_rn(invstd, g);
  alpha[i] = a;
  beta[i] = __fsub_rn(bnb ? bnb[i] : 0.0f, __fmul_rn(mean[i], a));
}

int dev_alloc(uq_model* m, void** p, size_t bytes) {
  UQ_CUDA(cudaMalloc(p, bytes ? bytes : 16));
  m->allocations.push_back(*p);
  return UQ_OK;
}

}  // namespace
}  // namespace uq

using namespace uq;

extern "C" {

int uq_abi_version(void) { return UQ_ABI_VERSION; }
const char* uq_last_error(void) { return g_err; }
uint64_t uq_launch_count(void) { return g_launches; }
void uq_launch_count_reset(void) { g_launches = 0; }
int uq_kde_jsd_phase_us(double* out5) {
  UQ_REQUIRE(out5 != nullptr, UQ_ERR_INVALID, "uq_kde_jsd_phase_us: out is NULL");
  return kde_jsd_fused_phase_us(out5);
}

int uq_model_destroy(uq_model_t* model) {
  if (!model) return UQ_OK;
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(model->device);
  for (void* p : model->allocations) cudaFree(p);
  cudaSetDevice(cur);
  delete model;
  return UQ_OK;
}

int uq_model_create(uq_model_t** out, int32_t n_members, int32_t n_layers,
                    const uq_layer_desc* layers, void* stream) {
  return uq_model_create_ex(out, n_members, n_layers, layers, 0, stream);
}

int uq_model_create_ex(uq_model_t** out, int32_t n_members, int32_t n_layers,
                       const uq_layer_desc* layers, int32_t flags, void* stream) {
  UQ_REQUIRE(out != nullptr, UQ_ERR_INVALID, "uq_model_create: out is NULL");
  UQ_REQUIRE((flags & ~UQ_MODEL_ANCHOR_FIRST) == 0, UQ_ERR_INVALID,
             "uq_model_create_ex: unknown flags 0x%x", flags);
  *out = nullptr;
  UQ_REQUIRE(n_members >= 1 && n_layers >= 1 && layers != nullptr, UQ_ERR_INVALID,
             "uq_model_create: need >= 1 member and >= 1 Linear layer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int l = 0; l < n_layers; ++l) {
    const uq_layer_desc& d0 = layers[l];
    UQ_REQUIRE(d0.in_features >= 1 && d0.out_features >= 1 && d0.weight != nullptr,
               UQ_ERR_INVALID, "layer %d: bad shape or NULL weight", l);
    if (l > 0)
      UQ_REQUIRE(d0.in_features == layers[l - 1].out_features, UQ_ERR_INVALID,
                 "layer %d: in_features %d != previous out_features %d", l, d0.in_features,
                 layers[l - 1].out_features);
    for (int k = 1; k < n_members; ++k) {
      const uq_layer_desc& d = layers[(size_t)k * n_layers + l];
      UQ_REQUIRE(d.in_features == d0.in_features && d.out_features == d0.out_features &&
                     d.relu == d0.relu && d.dropout == d0.dropout &&
                     (d.bn_mean != nullptr) == (d0.bn_mean != nullptr) && d.weight != nullptr,
                 UQ_ERR_INVALID, "member %d layer %d differs structurally from member 0", k, l);
    }
    if (d0.bn_mean != nullptr || d0.bn_var != nullptr)
      UQ_REQUIRE(d0.bn_mean != nullptr && d0.bn_var != nullptr, UQ_ERR_INVALID,
                 "layer %d: BatchNorm needs running_mean and running_var (track_running_stats)",
                 l);
  }

  uq_model* m = new uq_model();
  cudaGetDevice(&m->device);
  m->n_members = n_members;
  m->n_layers = n_layers;
  m->anchor_first = (flags & UQ_MODEL_ANCHOR_FIRST) != 0;
  m->d_in = layers[0].in_features;
  m->d_out = layers[n_layers - 1].out_features;
  m->layers.resize(n_layers);
  int rc = UQ_OK;
  for (int l = 0; l < n_layers && rc == UQ_OK; ++l) {
    const uq_layer_desc& d0 = layers[l];
    Layer& ly = m->layers[l];
    ly.in = d0.in_features;
    ly.out = d0.out_features;
    ly.has_bn = d0.bn_mean != nullptr;
    ly.relu = d0.relu != 0;
    ly.dropout = d0.dropout != 0;
    if (ly.dropout) m->n_dropout++;
    if (ly.out > m->max_width) m->max_width = ly.out;
    const size_t wsz = (size_t)ly.out * ly.in;
    if ((rc = dev_alloc(m, (void**)&ly.w, sizeof(float) * wsz * n_members))) break;
    if ((rc = dev_alloc(m, (void**)&ly.bias, sizeof(float) * (size_t)ly.out * n_members))) break;
    if ((rc = dev_alloc(m, (void**)&ly.bias_folded, sizeof(float) * (size_t)ly.out * n_members)))
      break;
    if (ly.has_bn) {
      if ((rc = dev_alloc(m, (void**)&ly.alpha, sizeof(float) * (size_t)ly.out * n_members))) break;
      if ((rc = dev_alloc(m, (void**)&ly.beta, sizeof(float) * (size_t)ly.out * n_members))) break;
    }
    for (int k = 0; k < n_members && rc == UQ_OK; ++k) {
      const uq_layer_desc& d = layers[(size_t)k * n_layers + l];
      cudaError_t e = cudaMemcpyAsync(ly.w + wsz * k, d.weight, sizeof(float) * wsz,
                                      cudaMemcpyDeviceToDevice, st);
      if (e == cudaSuccess) {
        if (d.bias)
          e = cudaMemcpyAsync(ly.bias + (size_t)ly.out * k, d.bias, sizeof(float) * ly.out,
                              cudaMemcpyDeviceToDevice, st);
        else
          e = cudaMemsetAsync(ly.bias + (size_t)ly.out * k, 0, sizeof(float) * ly.out, st);
      }
      if (e != cudaSuccess) {
        rc = cuda_fail(e, "copy layer parameters", __FILE__, __LINE__);
        break;
      }
      if (ly.has_bn) {
        bn_fold_kernel<<<(ly.out + 127) / 128, 128, 0, st>>>(
            d.bn_weight, d.bn_bias, d.bn_mean, d.bn_var, d.bn_eps,
            ly.alpha + (size_t)ly.out * k, ly.beta + (size_t)ly.out * k, ly.out);
        count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) rc = cuda_fail(e, "bn_fold_kernel", __FILE__, __LINE__);
      }
    }
  }
  if (rc == UQ_OK) {
    tc_plan(m);
    if (m->tc.ok) rc = tc_pack(m, st);
    tcx_plan(m);
    if (rc == UQ_OK && m->tc.x_ok) rc = tcx_pack(m, st);
  }
  if (rc == UQ_OK) {
    // the caller may free its tensors once this returns
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
  }
  if (rc != UQ_OK) {
    uq_model_destroy(m);
    return rc;
  }
  *out = m;
  return UQ_OK;
}

int uq_model_supports_bf16(const uq_model_t* model) {
  if (!model) return 0;
  if (!model->tc.ok) set_error("bf16 path unavailable: %s", model->tc.why_not.c_str());
  return model->tc.ok ? 1 : 0;
}

int uq_model_supports_fp32_tc(const uq_model_t* model) {
  if (!model) return 0;
  if (!model->tc.x_ok)
    set_error("fp32 tensor-core split path unavailable: %s", model->tc.x_why_not.c_str());
  return model->tc.x_ok ? 1 : 0;
}

static int check_args(const uq_model_t* m, const float* x, int64_t n, const uq_forward_args* a,
                      const float* out0, const float* out1) {
  UQ_REQUIRE(m && a, UQ_ERR_INVALID, "uq_forward: NULL model or args");
  UQ_REQUIRE(n >= 1 && x && out0 && out1, UQ_ERR_INVALID,
             "uq_forward: need n >= 1 and non-NULL x/out pointers (n = %lld)", (long long)n);
  UQ_REQUIRE(n < ((int64_t)1 << 31), UQ_ERR_INVALID, "uq_forward: n = %lld exceeds 2^31-1",
             (long long)n);
  UQ_REQUIRE(a->mode >= UQ_MODE_ENSEMBLE && a->mode <= UQ_MODE_PAGER, UQ_ERR_INVALID,
             "uq_forward: unknown mode %d", a->mode);
  UQ_REQUIRE(a->precision == UQ_PREC_FP32 || a->precision == UQ_PREC_BF16 ||
                 a->precision == UQ_PREC_FP32_FFMA,
             UQ_ERR_INVALID,
             "uq_forward: unknown precision %d", a->precision);
  UQ_REQUIRE(a->output == UQ_OUT_MEAN_STD || a->output == UQ_OUT_MOMENTS, UQ_ERR_INVALID,
             "uq_forward: unknown output kind %d", a->output);
  UQ_REQUIRE(a->member_count >= 1 && a->member_begin >= 0 &&
                 a->member_begin + a->member_count <= a->total_members,
             UQ_ERR_INVALID, "uq_forward: member shard [%d, %d) outside [0, %d)", a->member_begin,
             a->member_begin + a->member_count, a->total_members);
  if (a->mode == UQ_MODE_ENSEMBLE) {
    UQ_REQUIRE(a->total_members == m->n_members, UQ_ERR_INVALID,
               "ensemble: total_members %d != packed members %d", a->total_members, m->n_members);
  } else {
    UQ_REQUIRE(m->n_members == 1, UQ_ERR_INVALID,
               "MC-dropout / Delta-UQ need a single packed network, got %d", m->n_members);
  }
  UQ_REQUIRE(a->row_base >= 0, UQ_ERR_INVALID, "uq_forward: row_base %d < 0", a->row_base);
  if (a->mode == UQ_MODE_MC_DROPOUT && a->dropout_active) {
    UQ_REQUIRE(a->dropout_p >= 0.0 && a->dropout_p < 1.0, UQ_ERR_INVALID,
               "dropout_p must be in [0, 1), got %g", a->dropout_p);
  }
  if (a->mode == UQ_MODE_DELTA_UQ || a->mode == UQ_MODE_PAGER) {
    UQ_REQUIRE(a->anchors != nullptr, UQ_ERR_INVALID, "Delta-UQ: anchors pointer is NULL");
    UQ_REQUIRE(m->d_in % 2 == 0, UQ_ERR_INVALID,
               "Delta-UQ: network input width %d is not 2 * d_in", m->d_in);
  }
  if (a->mode == UQ_MODE_PAGER) {
    UQ_REQUIRE(a->anchor_targets != nullptr, UQ_ERR_INVALID,
               "PAGER: anchor_targets (anchors_Y) pointer is NULL");
    UQ_REQUIRE(a->output == UQ_OUT_MEAN_STD, UQ_ERR_INVALID,
               "PAGER: the conformal score is a max over anchors, not a moment (output must be "
               "UQ_OUT_MEAN_STD; K-shards combine with an element-wise max)");
  }
  return UQ_OK;
}

size_t uq_forward_workspace_bytes(const uq_model_t* model, int64_t n,
                                  const uq_forward_args* args) {
  if (!model || !args || n < 1) return 0;
  if (args->precision == UQ_PREC_BF16) return tc_workspace_bytes(model, n, args, false);
  if (args->precision == UQ_PREC_FP32 && model->tc.x_ok)
    return tc_workspace_bytes(model, n, args, true);
  return fp32_workspace_bytes(model, n, args);
}

int uq_forward(const uq_model_t* model, const float* x, int64_t n, const uq_forward_args* args,
               float* out0, float* out1, void* workspace, size_t workspace_bytes,
               double* out_count, void* stream) {
  int rc = check_args(model, x, n, args, out0, out1);
  if (rc != UQ_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (args->precision == UQ_PREC_BF16) {
    UQ_REQUIRE(model->tc.ok, UQ_ERR_UNSUPPORTED, "bf16 path unavailable for this model: %s",
               model->tc.why_not.c_str());
    rc = tc_forward(model, x, n, args, out0, out1, workspace, workspace_bytes, st, false);
  } else if (args->precision == UQ_PREC_FP32 && model->tc.x_ok) {
    // the 1e-5 parity mode on the tensor cores (scaled fp16 x 2 split, mlp_tcx.cu)
    rc = tc_forward(model, x, n, args, out0, out1, workspace, workspace_bytes, st, true);
  } else {
    rc = fp32_forward(model, x, n, args, out0, out1, workspace, workspace_bytes, st);
  }
  if (rc == UQ_OK && out_count) *out_count = (double)args->member_count;
  return rc;
}

int uq_forward_host(const uq_model_t* model, const float* x_host, int64_t n,
                    const uq_forward_args* args, float* out0_host, float* out1_host,
                    void* stream) {
  UQ_REQUIRE(model && args && x_host && out0_host && out1_host && n >= 1, UQ_ERR_INVALID,
             "uq_forward_host: NULL argument or n < 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int d_x = (args->mode == UQ_MODE_DELTA_UQ || args->mode == UQ_MODE_PAGER)
                      ? model->d_in / 2 : model->d_in;
  const size_t xb = sizeof(float) * (size_t)n * d_x;
  const size_t ob = sizeof(float) * (size_t)n * model->d_out;
  const size_t wsb = uq_forward_workspace_bytes(model, n, args);
  const size_t xb_a = (xb + 255) & ~(size_t)255, ob_a = (ob + 255) & ~(size_t)255;
  char* buf = nullptr;
  UQ_CUDA(cudaMallocAsync((void**)&buf, xb_a + 2 * ob_a + wsb + 256, st));
  float* dx = reinterpret_cast<float*>(buf);
  float* d0 = reinterpret_cast<float*>(buf + xb_a);
  float* d1 = reinterpret_cast<float*>(buf + xb_a + ob_a);
  void* ws = buf + xb_a + 2 * ob_a;
  int rc = UQ_OK;
  cudaError_t e = cudaMemcpyAsync(dx, x_host, xb, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) rc = cuda_fail(e, "H2D x", __FILE__, __LINE__);
  if (rc == UQ_OK) rc = uq_forward(model, dx, n, args, d0, d1, ws, wsb, nullptr, st);
  if (rc == UQ_OK) {
    e = cudaMemcpyAsync(out0_host, d0, ob, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out1_host, d1, ob, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) rc = cuda_fail(e, "D2H outputs", __FILE__, __LINE__);
  }
  cudaFreeAsync(buf, st);
  e = cudaStreamSynchronize(st);
  if (rc == UQ_OK && e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
  return rc;
}

int uq_moments_merge(const float* means, const float* m2s, const double* counts,
                     int32_t n_shards, int64_t len, float* out_mean, float* out_std,
                     void* stream) {
  UQ_REQUIRE(means && m2s && counts && out_mean && out_std && n_shards >= 1 && len >= 1,
             UQ_ERR_INVALID, "uq_moments_merge: NULL argument or empty input");
  UQ_REQUIRE(n_shards <= 64, UQ_ERR_INVALID, "uq_moments_merge: at most 64 shards");
  return moments_merge(means, m2s, counts, n_shards, len, out_mean, out_std,
                       static_cast<cudaStream_t>(stream));
}

int uq_moments_merge_ex(const float* means, const float* m2s, int64_t shard_stride,
                        const double* counts, int32_t n_shards, int64_t len, float* out_mean,
                        float* out_second, int32_t output, void* stream) {
  UQ_REQUIRE(means && m2s && counts && out_mean && out_second && n_shards >= 1 && len >= 1,
             UQ_ERR_INVALID, "uq_moments_merge_ex: NULL argument or empty input");
  UQ_REQUIRE(n_shards <= 64, UQ_ERR_INVALID, "uq_moments_merge_ex: at most 64 shards");
  UQ_REQUIRE(shard_stride >= 1, UQ_ERR_INVALID, "uq_moments_merge_ex: shard_stride < 1");
  UQ_REQUIRE(output == UQ_OUT_MEAN_STD || output == UQ_OUT_MOMENTS, UQ_ERR_INVALID,
             "uq_moments_merge_ex: unknown output kind %d", output);
  return moments_merge_strided(means, m2s, shard_stride, counts, n_shards, len, out_mean,
                               out_second, output == UQ_OUT_MOMENTS ? 1 : 0,
                               static_cast<cudaStream_t>(stream));
}

// Export the native Philox keep-masks in the injected-mask layout (uq_forward_args.masks), so a
// native-RNG MC-dropout run can be replayed bit-for-bit through the oracle.
__global__ void philox_export_kernel(uint8_t* __restrict__ out, int64_t n, int width, int passes,
                                     int layer, uq::PhiloxKey key, uint32_t thr16) {
  const int64_t groups = (width + 31) / 32;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)passes * n * groups;
  if (i >= total) return;
  const int64_t g = i % groups;
  const int64_t s = (i / groups) % n;
  const int64_t p = i / (groups * n);
  const uint32_t m = uq::dropout_keep32(key, thr16, (uint32_t)p, (uint32_t)layer, (uint32_t)s,
                                        (uint32_t)g);
  uint8_t* o = out + (p * n + s) * width + g * 32;
  for (int b = 0; b < 32 && g * 32 + b < width; ++b) o[b] = (m >> b) & 1u;
}

int uq_philox_keep_masks(uint8_t* out, int64_t n, int32_t width, int32_t total_members,
                         int32_t dropout_layer, double dropout_p, uint64_t seed, uint64_t offset,
                         void* stream) {
  UQ_REQUIRE(out && n >= 1 && width >= 1 && total_members >= 1 && dropout_layer >= 0,
             UQ_ERR_INVALID, "uq_philox_keep_masks: bad argument");
  uq::PhiloxKey key{(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32),
                    (uint32_t)(offset & 0xffffffffu)};
  const int64_t total = (int64_t)total_members * n * ((width + 31) / 32);
  philox_export_kernel<<<(unsigned)((total + 255) / 256), 256, 0,
                         static_cast<cudaStream_t>(stream)>>>(
      out, n, width, total_members, dropout_layer, key, uq::dropout_thr16((float)dropout_p));
  UQ_LAUNCH_CHECK();
  return UQ_OK;
}

size_t uq_wasserstein_workspace_bytes(int64_t nu, int64_t nv) {
  return wasserstein_workspace_bytes(nu, nv);
}

int uq_wasserstein_1d(const float* u, int64_t nu, const float* v, int64_t nv, double* out_host,
                      void* workspace, size_t workspace_bytes, void* stream) {
  UQ_REQUIRE(out_host != nullptr, UQ_ERR_INVALID, "uq_wasserstein_1d: out is NULL");
  UQ_REQUIRE(u && v && nu >= 1 && nv >= 1, UQ_ERR_INVALID,
             "uq_wasserstein_1d: Distribution can't be empty.");
  return wasserstein_1d(u, nu, v, nv, UQ_WASSERSTEIN_AUTO, out_host, nullptr, workspace,
                        workspace_bytes, static_cast<cudaStream_t>(stream));
}

int uq_wasserstein_1d_ex(const float* u, int64_t nu, const float* v, int64_t nv, int32_t method,
                         double* out_host, int64_t* info_host, void* workspace,
                         size_t workspace_bytes, void* stream) {
  UQ_REQUIRE(out_host != nullptr, UQ_ERR_INVALID, "uq_wasserstein_1d_ex: out is NULL");
  UQ_REQUIRE(u && v && nu >= 1 && nv >= 1, UQ_ERR_INVALID,
             "uq_wasserstein_1d_ex: Distribution can't be empty.");
  return wasserstein_1d(u, nu, v, nv, method, out_host, info_host, workspace, workspace_bytes,
                        static_cast<cudaStream_t>(stream));
}

int uq_wasserstein_1d_enqueue(const float* u, int64_t nu, const float* v, int64_t nv, void* record,
                              void* workspace, size_t workspace_bytes, void* stream) {
  UQ_REQUIRE(u && v && nu >= 1 && nv >= 1, UQ_ERR_INVALID,
             "uq_wasserstein_1d_enqueue: Distribution can't be empty.");
  return wasserstein_1d_enqueue(u, nu, v, nv, record, workspace, workspace_bytes,
                                static_cast<cudaStream_t>(stream));
}

int uq_wasserstein_1d_finish(const float* u, int64_t nu, const float* v, int64_t nv,
                             const void* record, double* out_host, int64_t* info_host,
                             void* workspace, size_t workspace_bytes, void* stream) {
  UQ_REQUIRE(u && v && record && out_host && nu >= 1 && nv >= 1, UQ_ERR_INVALID,
             "uq_wasserstein_1d_finish: NULL argument or empty distribution");
  return wasserstein_1d_finish(u, nu, v, nv, record, out_host, info_host, workspace,
                               workspace_bytes, static_cast<cudaStream_t>(stream));
}

int uq_kde_jsd_enqueue(const float* u, int64_t nu, const float* v, int64_t nv, int32_t grid_pts,
                       void* record, void* workspace, size_t workspace_bytes, void* stream) {
  UQ_REQUIRE(u && v && nu >= 2 && nv >= 2, UQ_ERR_INVALID,
             "uq_kde_jsd_enqueue: each sample needs at least 2 values (got %lld, %lld)",
             (long long)nu, (long long)nv);
  UQ_REQUIRE(grid_pts >= 2, UQ_ERR_INVALID, "uq_kde_jsd_enqueue: grid_pts must be >= 2");
  return kde_jsd_enqueue(u, nu, v, nv, grid_pts, record, workspace, workspace_bytes,
                         static_cast<cudaStream_t>(stream));
}

int uq_kde_jsd_finish(const float* u, int64_t nu, const float* v, int64_t nv, int32_t grid_pts,
                      const void* record, double* out_host, int32_t* method_used, void* workspace,
                      size_t workspace_bytes, void* stream) {
  UQ_REQUIRE(u && v && record && out_host && nu >= 2 && nv >= 2 && grid_pts >= 2, UQ_ERR_INVALID,
             "uq_kde_jsd_finish: NULL argument, fewer than 2 values or fewer than 2 grid points");
  return kde_jsd_finish(u, nu, v, nv, grid_pts, record, out_host, method_used, workspace,
                        workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t uq_kde_jsd_workspace_bytes(int64_t nu, int64_t nv, int32_t grid_pts) {
  return kde_jsd_workspace_bytes(nu, nv, grid_pts);
}

int uq_kde_jsd(const float* u, int64_t nu, const float* v, int64_t nv, int32_t grid_pts,
               double* out_host, void* workspace, size_t workspace_bytes, void* stream) {
  UQ_REQUIRE(out_host != nullptr, UQ_ERR_INVALID, "uq_kde_jsd: out is NULL");
  UQ_REQUIRE(u && v && nu >= 2 && nv >= 2, UQ_ERR_INVALID,
             "uq_kde_jsd: each sample needs at least 2 values (got %lld, %lld)", (long long)nu,
             (long long)nv);
  UQ_REQUIRE(grid_pts >= 2 && grid_pts <= (1 << 20), UQ_ERR_INVALID,
             "uq_kde_jsd: grid_pts %d outside [2, 2^20]", grid_pts);
  return kde_jsd(u, nu, v, nv, grid_pts, UQ_KDE_AUTO, out_host, nullptr, workspace,
                 workspace_bytes, static_cast<cudaStream_t>(stream));
}

int uq_kde_jsd_ex(const float* u, int64_t nu, const float* v, int64_t nv, int32_t grid_pts,
                  int32_t method, double* out_host, int32_t* method_used, void* workspace,
                  size_t workspace_bytes, void* stream) {
  UQ_REQUIRE(out_host != nullptr, UQ_ERR_INVALID, "uq_kde_jsd_ex: out is NULL");
  UQ_REQUIRE(u && v && nu >= 2 && nv >= 2, UQ_ERR_INVALID,
             "uq_kde_jsd_ex: each sample needs at least 2 values (got %lld, %lld)", (long long)nu,
             (long long)nv);
  UQ_REQUIRE(grid_pts >= 2 && grid_pts <= (1 << 20), UQ_ERR_INVALID,
             "uq_kde_jsd_ex: grid_pts %d outside [2, 2^20]", grid_pts);
  int used = 0;
  const int rc = kde_jsd(u, nu, v, nv, grid_pts, method, out_host, &used, workspace,
                         workspace_bytes, static_cast<cudaStream_t>(stream));
  if (method_used) *method_used = used;
  return rc;
}

size_t uq_kde_grid_workspace_bytes(int64_t n) { return kde_grid_workspace_bytes(n); }

int uq_kde_grid_accumulate(const float* x, int64_t n, double lo, double hi, double bandwidth,
                           int32_t grid_pts, double* grid, void* workspace,
                           size_t workspace_bytes, void* stream) {
  UQ_REQUIRE(x && grid && n >= 1, UQ_ERR_INVALID, "uq_kde_grid_accumulate: NULL argument or n < 1");
  UQ_REQUIRE(grid_pts >= 2 && grid_pts <= (1 << 20), UQ_ERR_INVALID,
             "uq_kde_grid_accumulate: grid_pts %d outside [2, 2^20]", grid_pts);
  return kde_grid_accumulate(x, n, lo, hi, bandwidth, grid_pts, grid, workspace, workspace_bytes,
                             static_cast<cudaStream_t>(stream));
}

int uq_jsd_from_grids(const double* grids, int32_t grid_pts, double* out_host, void* stream) {
  UQ_REQUIRE(grids && out_host && grid_pts >= 2, UQ_ERR_INVALID,
             "uq_jsd_from_grids: NULL argument or grid_pts < 2");
  return jsd_from_grids(grids, grid_pts, out_host, static_cast<cudaStream_t>(stream));
}

int uq_wasserstein_1d_range(const float* u, int64_t nu, const float* v, int64_t nv,
                            int64_t u_below, int64_t v_below, int64_t nu_total, int64_t nv_total,
                            double* out_host, void* workspace, size_t workspace_bytes,
                            void* stream) {
  UQ_REQUIRE(out_host != nullptr, UQ_ERR_INVALID, "uq_wasserstein_1d_range: out is NULL");
  UQ_REQUIRE(nu >= 0 && nv >= 0 && nu + nv >= 1 && (nu == 0 || u) && (nv == 0 || v),
             UQ_ERR_INVALID, "uq_wasserstein_1d_range: empty range or NULL pointer");
  UQ_REQUIRE(nu_total >= 1 && nv_total >= 1 && u_below >= 0 && v_below >= 0 &&
                 u_below + nu <= nu_total && v_below + nv <= nv_total,
             UQ_ERR_INVALID, "uq_wasserstein_1d_range: inconsistent counts");
  return wasserstein_1d_range(u, nu, v, nv, u_below, v_below, nu_total, nv_total, out_host,
                              workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

double uq_kde_scott_bandwidth(int64_t m, int32_t d) {
  // sklearn.neighbors.KernelDensity.fit: bandwidth_ = n_samples ** (-1 / (n_features + 4))
  return (m < 1 || d < 1) ? 0.0 : pow((double)m, -1.0 / ((double)d + 4.0));
}

double uq_kde_silverman_bandwidth(int64_t m, int32_t d) {
  // sklearn.neighbors.KernelDensity.fit:
  // bandwidth_ = (n_samples * (n_features + 2) / 4) ** (-1 / (n_features + 4))
  return (m < 1 || d < 1) ? 0.0
                          : pow((double)m * ((double)d + 2.0) / 4.0, -1.0 / ((double)d + 4.0));
}

size_t uq_kde_density_workspace_bytes(int64_t n, int64_t m) {
  return kde_density_workspace_bytes(n, m);
}

int uq_kde_density(const float* fit, int64_t m, const float* x, int64_t n, int32_t d,
                   double bandwidth, double* out, void* workspace, size_t workspace_bytes,
                   void* stream) {
  UQ_REQUIRE(fit && x && out && m >= 1 && n >= 1, UQ_ERR_INVALID,
             "uq_kde_density: NULL argument, no fitted rows or no query rows");
  UQ_REQUIRE(n < ((int64_t)1 << 31) && m < ((int64_t)1 << 31), UQ_ERR_INVALID,
             "uq_kde_density: too many rows");
  return kde_density(fit, m, x, n, d, bandwidth, out, workspace, workspace_bytes,
                     static_cast<cudaStream_t>(stream));
}

int uq_bin_moments(const float* x, int64_t n, uint64_t* cnt, uint64_t* ksum, void* stream) {
  UQ_REQUIRE(x && cnt && ksum && n >= 1, UQ_ERR_INVALID, "uq_bin_moments: NULL argument or n < 1");
  return bin_moments_accumulate(x, n, reinterpret_cast<unsigned long long*>(cnt),
                                reinterpret_cast<unsigned long long*>(ksum),
                                static_cast<cudaStream_t>(stream));
}

int uq_wasserstein_from_bins(const uint64_t* tables, int64_t nu_total, int64_t nv_total,
                             uint8_t* flags_out, double* out_host, void* workspace,
                             size_t workspace_bytes, void* stream) {
  UQ_REQUIRE(tables && flags_out && out_host, UQ_ERR_INVALID,
             "uq_wasserstein_from_bins: NULL argument");
  UQ_REQUIRE(nu_total >= 1 && nv_total >= 1 && nu_total + nv_total < ((int64_t)1 << 31),
             UQ_ERR_INVALID, "uq_wasserstein_from_bins: Distribution can't be empty (or too large).");
  return wasserstein_from_bins(reinterpret_cast<const unsigned long long*>(tables), nu_total,
                               nv_total, flags_out, out_host, workspace, workspace_bytes,
                               static_cast<cudaStream_t>(stream));
}

int uq_compact_flagged(const float* x, int64_t n, const uint8_t* flags, float* out,
                       int64_t* count_host, void* workspace, size_t workspace_bytes,
                       void* stream) {
  UQ_REQUIRE(x && flags && out && count_host && n >= 1, UQ_ERR_INVALID,
             "uq_compact_flagged: NULL argument or n < 1");
  return compact_flagged(x, n, flags, out, count_host, workspace, workspace_bytes,
                         static_cast<cudaStream_t>(stream));
}

int uq_wasserstein_ambiguous(const float* u_amb, int64_t nu_amb, const float* v_amb,
                             int64_t nv_amb, const uint64_t* tables, int64_t nu_total,
                             int64_t nv_total, double* out_host, void* workspace,
                             size_t workspace_bytes, void* stream) {
  UQ_REQUIRE(tables && out_host && nu_amb >= 0 && nv_amb >= 0 && (nu_amb == 0 || u_amb) &&
                 (nv_amb == 0 || v_amb),
             UQ_ERR_INVALID, "uq_wasserstein_ambiguous: NULL argument");
  UQ_REQUIRE(nu_total >= 1 && nv_total >= 1, UQ_ERR_INVALID,
             "uq_wasserstein_ambiguous: Distribution can't be empty.");
  return wasserstein_ambiguous(u_amb, nu_amb, v_amb, nv_amb,
                               reinterpret_cast<const unsigned long long*>(tables), nu_total,
                               nv_total, out_host, workspace, workspace_bytes,
                               static_cast<cudaStream_t>(stream));
}

}  // extern "C"
