// Launch parameters shared by the fused tcgen05 kernels (mlp_tc.cu, mlp_tc2.cu).
#pragma once
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "philox.cuh"

namespace uq {
namespace tc {

constexpr int MAX_MMA_LAYERS = 16;
constexpr int MAX_DOUT = 8;
constexpr int TRACE_LEN = 4096;
constexpr int TRACE_ROLES = 6;

struct TcParams {
  const float* x;        // [n][d_x]
  int64_t n;
  int64_t row_base;      // global index of row 0 (Philox sample counter = row_base + row)
  int d_x;               // features of x (d_in, or d_in/2 for Delta-UQ)
  int d_in;              // network input features
  int mode;
  int n_tiles;
  int splits;            // member-axis splits (partial moments when > 1)
  int split_major;       // mlp_tc3.cu / mlp_tcx.cu: unit order (split, tile pair), not (tile pair, split)
  int member_begin, member_count, total_members;
  int K0, split_s, L_mma, d_out;
  int stages_per_member;
  int shared_weights;
  int bias0_per_member;  // layer-0 bias is indexed by the global member id (Delta-UQ anchors)
  const uint8_t* image;
  // mlp_tc2.cu, bias-in-the-MMA variant: per weight slot, one stage [N x 64] bf16 per (layer,
  // accumulator half) whose K columns 0..2 hold the folded bias as bf16 hi + mid + lo; null = the
  // epilogue adds the bias
  const uint8_t* bias_image;
  // anchored modes: the layer-0 bias stages, one set per global member id (built per call from the
  // per-anchor bias); null = layer 0 uses bias_image like every other layer
  const uint8_t* bias0_image;
  const float* bias[MAX_MMA_LAYERS];  // [K or 1][H] folded bias per MMA layer
  uint32_t relu_mask, dropout_mask;   // bit l: MMA layer l has ReLU / dropout on its output
  const float* w_last;                // [K or 1][DOUT][H]  (zero rows beyond d_out)
  const float* b_last;                // [K or 1][DOUT]
  int last_relu;
  int drop_mode;                      // 0 none, 1 injected, 2 philox
  float drop_scale;
  uint32_t thr16;
  PhiloxKey key;
  const uint8_t* masks;               // injected base
  const float* anchors;               // [total_members][d_x]
  const float* targets;               // PAGER: [total_members][d_out]; the second moment slot then
  const float* score_floor;           //   holds max_k |y_k - target_k| (>= score_floor[n] at the end)
  float* out0;
  float* out1;
  int output;                         // UQ_OUT_*
  float* part_mean;                   // [splits][n*d_out] when splits > 1
  float* part_m2;
  unsigned int* error_flag;
  // fp32-parity split mode (mlp_tcx.cu) only:
  const float* lstats;                // [K or 1][L_mma][4] = {1/C, max_n |w_n|_2, max |bias|, C}
  const float* bmax0;                 // [total_members] max |bias0[k]| (Delta-UQ / PAGER) or null
  unsigned long long* trace;          // -DUQ_TC_TRACE builds only: [3 roles][TRACE_LEN][2] of CTA 0
};

}  // namespace tc

// The bias-in-the-MMA variants (mlp_tc2.cu, mlp_tc4.cu) are the default where they apply (no live
// dropout, d_out 1); UQ_TC_BIAS_MMA=0 selects the epilogue-bias variants for A/B runs.
inline bool bias_in_mma_enabled() {
  const char* e = getenv("UQ_TC_BIAS_MMA");
  return e ? e[0] == '1' : true;
}

// mlp_tc2.cu: CTA-pair (cta_group::2) variant of the fused kernel; same weight image
int tc2_launch(const tc::TcParams& p, int hidden, int dout_pad, cudaStream_t st);
bool tc2_supported(int hidden);
// mlp_tc3.cu: CTA pairs with 64 rows per CTA for hidden widths 768 / 1024; same weight image
int tc3_launch(const tc::TcParams& p, int hidden, int dout_pad, cudaStream_t st);
bool tc3_supported(int hidden);
// mlp_tc4.cu: CTA pairs with four sample tiles in flight per CTA for hidden widths 64 / 128, d_out 1
int tc4_launch(const tc::TcParams& p, int hidden, cudaStream_t st);
bool tc4_supported(int hidden, int dout_pad);
int tc4_rows_per_unit();
// mlp_tcx.cu: fp32-parity mode on the tensor cores (scaled fp16 x 2 split, 64 rows per CTA),
// hidden widths 64 .. 512; its own weight image (tcx_pack)
int tcx_launch(const tc::TcParams& p, int hidden, int dout_pad, cudaStream_t st);
bool tcx_supported(int hidden);
int tcx_rows_per_unit();
// mlp_tcx4.cu: the same split mode with four 64-row tile slots per CTA for hidden widths 64 / 128,
// d_out 1 (the reference's own 6 x 128 architecture); same weight image
int tcx4_launch(const tc::TcParams& p, int hidden, cudaStream_t st);
bool tcx4_supported(int hidden, int dout_pad);
int tcx4_rows_per_unit();

}  // namespace uq
