// edgeconv.cuh — declarations shared by the FP32-FMA (edgeconv.cu) and tensor-core (edgeconv_tc.cu) EdgeConv kernels.
#pragma once
#include "common.cuh"

namespace dg {

constexpr int kLdH = kCout + 4;          // smem row stride of 64-wide tiles
constexpr int kMaxP = 16;                // points per tile cap (sizes the per-point smem arrays)

enum { MODE_EVAL = 0, MODE_TRAIN_A = 1, MODE_TRAIN_B = 2 };

struct EcK {
  const float* rows; int ld;
  const int32_t* idx;
  int N, k, M, C, C4, P, num_tiles;
  const float* w_a; const float* w_b; int two;
  float* out; int ld_out;
  const float* aff_a; const float* aff_b;     // float[192]: mean | scale | beta
  float* y_a_out; const float* y_a_in;
  float* v_center;                            // [M,64] scratch: hoisted centre term Wa2 x_i (tensor-core kernels)
  int tc;                                     // 0 = FFMA kernels, 1 = serial tensor-core kernels, 2 = 1 + pipelined conv_b (default)
  double* partials;
  float* ymax; float* ymin; uint8_t* amax; uint8_t* amin;
};

__device__ __forceinline__ float bn_apply(float y, float mean, float scale, float beta) {
  return __fmaf_rn(y - mean, scale, beta);
}

// tensor-core forward kernels (edgeconv_tc.cu); mode = MODE_EVAL / MODE_TRAIN_A (need C == 64 and v_center) or MODE_TRAIN_B
int launch_ec_tc_fwd(const EcK& p, int mode, int grid, cudaStream_t st);
int launch_ec_tc_bwd_b(const EcK& p, const float* g, const uint8_t* arg, const float* coef_b, float* dz_a,
                       float* dw_partials, float* dw_b, int grid, cudaStream_t st);
// T partials [nparts][2][64][64] -> dW_b (edgeconv_tc.cu)
int launch_ec_dw_finish(const float* dw_partials, int nparts, const float* w_b, const float* coef_b, float* dw_b, cudaStream_t st);
// pipelined conv_b backward with the fused scatter + the point-level finish of the BN_a backward (edgeconv_bwd_pipe.cu)
int launch_ec_pipe_bwd_b(const EcK& p, const float* g, const uint8_t* arg, const float* coef_b, float* r_raw, float* s_raw, int ld_rs,
                         float* yin, float* cnt, float* dw_partials, float* dw_b, int sms, cudaStream_t st);
int launch_ec_bwd_point_stats(const float* U, int ld_u, const float* V, int ld_v, const float* R, const float* S, int ld_rs, int M,
                              double* partials, cudaStream_t st);
int launch_ec_bwd_point_apply(float* R, float* S, int ld_rs, const float* yin, const float* ysum, const float* cnt, const float* coef_a,
                              int M, int k, cudaStream_t st);
// pipelined conv_b training pass (edgeconv_pipe.cu); one CTA per SM
int launch_ec_pipe_conv_b(const EcK& p, int sms, cudaStream_t st);

}  // namespace dg
