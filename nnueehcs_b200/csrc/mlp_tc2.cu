// bf16 tcgen05 path of the UQ forward, CTA-pair variant (cta_group::2): dispatch, and the
// epilogue-bias instantiations of the kernel in mlp_tc2_impl.cuh (d_out > 1, the 16-warp A/B
// variant, UQ_TC_BIAS_MMA=0).  The bias-in-the-MMA instantiations live in mlp_tc2_bias.cu.
#include "mlp_tc2_impl.cuh"

namespace uq {

int tc2_bias_launch(const tc::TcParams& p, int hidden, cudaStream_t st);   // mlp_tc2_bias.cu

namespace {

template <int H, int DOUT, int NG>
int launch_tc2(const TcParams& p, cudaStream_t st) {
  const bool mc = p.drop_mode != 0 && p.dropout_mask != 0;
  return mc ? launch_tc2_mc<H, DOUT, NG, true>(p, st) : launch_tc2_mc<H, DOUT, NG, false>(p, st);
}

template <int DOUT, int NG>
int dispatch_h2(int H, const TcParams& p, cudaStream_t st) {
  switch (H) {
    case 64: return launch_tc2<64, DOUT, NG>(p, st);
    case 128: return launch_tc2<128, DOUT, NG>(p, st);
    case 192: return launch_tc2<192, DOUT, NG>(p, st);
    case 256: return launch_tc2<256, DOUT, NG>(p, st);
    case 320: return launch_tc2<320, DOUT, NG>(p, st);
    case 384: return launch_tc2<384, DOUT, NG>(p, st);
    case 448: return launch_tc2<448, DOUT, NG>(p, st);
    case 512: return launch_tc2<512, DOUT, NG>(p, st);
  }
  set_error("bf16 pair kernel: unsupported hidden width %d", H);
  return UQ_ERR_UNSUPPORTED;
}

}  // namespace

bool tc2_supported(int hidden) { return hidden % 64 == 0 && hidden >= 64 && hidden <= 512; }

int tc2_launch(const tc::TcParams& p, int hidden, int dout_pad, cudaStream_t st) {
  if (dout_pad != 1) return dispatch_h2<tc::MAX_DOUT, 2>(hidden, p, st);
  // 8 epilogue warps by default (measured faster: 16 warps overlap more of the drain with the
  // MMAs and slow them down); UQ_TC_EPI_WARPS=16 selects the 16-warp variant for A/B runs.
  // (16 epilogue warps with the bias in the MMA were measured too: 15.0 ms against 13.7 on
  // ensemble16x512_1M, 240 against 203 ms on mcdropout_1000x512_64k, profiles/r02_j_bench_epi16_*;
  // the 16-warp A/B variant therefore stays epilogue-bias only)
  const char* e = getenv("UQ_TC_EPI_WARPS");
  if (e && e[0] == '1') return dispatch_h2<1, 4>(hidden, p, st);
  if (p.bias_image != nullptr && bias_in_mma_enabled()) return tc2_bias_launch(p, hidden, st);
  return dispatch_h2<1, 2>(hidden, p, st);
}

}  // namespace uq
