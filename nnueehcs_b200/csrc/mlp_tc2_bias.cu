// bf16 tcgen05 path of the UQ forward, CTA-pair variant: the bias-in-the-MMA instantiations
// (uq_mlp_tc2_kernel<H, 1, 2, MC, true>, the default for d_out 1) of the kernel in mlp_tc2_impl.cuh.
#include "mlp_tc2_impl.cuh"

namespace uq {

namespace {

template <int H>
int launch_tc2_bias(const TcParams& p, cudaStream_t st) {
  const bool mc = p.drop_mode != 0 && p.dropout_mask != 0;
  return mc ? launch_tc2_mc<H, 1, 2, true, true>(p, st) : launch_tc2_mc<H, 1, 2, false, true>(p, st);
}

}  // namespace

int tc2_bias_launch(const tc::TcParams& p, int hidden, cudaStream_t st) {
  switch (hidden) {
    case 64: return launch_tc2_bias<64>(p, st);
    case 128: return launch_tc2_bias<128>(p, st);
    case 192: return launch_tc2_bias<192>(p, st);
    case 256: return launch_tc2_bias<256>(p, st);
    case 320: return launch_tc2_bias<320>(p, st);
    case 384: return launch_tc2_bias<384>(p, st);
    case 448: return launch_tc2_bias<448>(p, st);
    case 512: return launch_tc2_bias<512>(p, st);
  }
  set_error("bf16 pair kernel: unsupported hidden width %d", hidden);
  return UQ_ERR_UNSUPPORTED;
}

}  // namespace uq
