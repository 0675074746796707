// Stable LSD radix sort of float32 keys (ascending), hand-written for the metric kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace uq {

// scratch bytes radix_sort_f32 needs besides the two key buffers
size_t radix_sort_scratch_bytes(int64_t n);

// Sorts `n` floats ascending.  `keys` is the input and is clobbered; the sorted result ends up in
// `*sorted`, which is either `keys` or `tmp` (both must hold n floats).  NaNs sort last.
// Returns a UQ_* status.
int radix_sort_f32(float* keys, float* tmp, int64_t n, void* scratch, size_t scratch_bytes,
                   float** sorted, cudaStream_t st);

// Same sort of the `n` floats at `src`, which is only read (pass 0 reads it directly: no copy
// unless `src` is not 16-byte aligned).  The result ends up in `keys`; src == keys is allowed.
int radix_sort_f32_copy(const float* src, float* keys, float* tmp, int64_t n, void* scratch,
                        size_t scratch_bytes, float** sorted, cudaStream_t st);

}  // namespace uq
