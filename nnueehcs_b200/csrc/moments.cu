// Cross-shard combine of per-sample (count, mean, M2) moments -- the one exchange step of the
// K-axis (members / passes / anchors) sharding.  Each GPU reduces its own members with the
// in-kernel Welford of the forward; NCCL all-gathers the [mean | M2] slabs; this kernel folds the
// G shards with Chan's pairwise formula in float64 and finalises the unbiased std that
// evaluation.py consumes (sqrt(M2 / (n - 1)), torch.std default, nnueehcs/models.py:106,162).
// A sum-allreduce of raw power sums would cancel catastrophically in fp32 when std << |mean|
// (the in-distribution case), which is why moments, not sums, cross the wire.
#include "common.cuh"

namespace uq {

namespace {

struct Counts {
  double c[64];
};

// shard s holds its means at means + s * shard_stride (likewise m2s); shards with count 0 are
// skipped (more ranks than members); moments != 0 writes the merged M2 instead of the std
__global__ void __launch_bounds__(256)
moments_merge_kernel(const float* __restrict__ means, const float* __restrict__ m2s,
                     int64_t shard_stride, Counts counts, int n_shards, int64_t len,
                     float* __restrict__ out_mean, float* __restrict__ out_std, int moments) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
    double n = 0.0, mean = 0.0, m2 = 0.0;
    for (int s = 0; s < n_shards; ++s) {
      const double nb = counts.c[s];
      if (nb <= 0.0) continue;
      const double mb = (double)means[(int64_t)s * shard_stride + i];
      const double sb = (double)m2s[(int64_t)s * shard_stride + i];
      const double nt = n + nb;
      const double d = mb - mean;
      mean += d * (nb / nt);
      m2 += sb + d * d * (n * nb / nt);
      n = nt;
    }
    out_mean[i] = (float)mean;
    out_std[i] = moments ? (float)m2 : (float)sqrt(m2 / (n - 1.0));
  }
}

}  // namespace

int moments_merge(const float* means, const float* m2s, const double* counts, int n_shards,
                  int64_t len, float* out_mean, float* out_std, cudaStream_t st) {
  return moments_merge_strided(means, m2s, len, counts, n_shards, len, out_mean, out_std, 0, st);
}

int moments_merge_strided(const float* means, const float* m2s, int64_t shard_stride,
                          const double* counts, int n_shards, int64_t len, float* out_mean,
                          float* out_second, int moments, cudaStream_t st) {
  Counts c;
  for (int s = 0; s < 64; ++s) c.c[s] = s < n_shards ? counts[s] : 0.0;
  int64_t blocks = (len + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  moments_merge_kernel<<<(unsigned)blocks, 256, 0, st>>>(means, m2s, shard_stride, c, n_shards, len,
                                                         out_mean, out_second, moments);
  UQ_LAUNCH_CHECK();
  return UQ_OK;
}

}  // namespace uq
