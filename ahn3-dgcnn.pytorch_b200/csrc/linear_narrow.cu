// linear_narrow.cu — y = x W^T for a NARROW output (the classifier conv9: 256 -> num_classes = 5, model.py:141,187-193) on
// the FP32 FMA pipe, forward and backward.
//
// On the tensor-core path this layer is all padding: N = 5 outputs occupy a 128-wide tile, the backward contracts over
// K = 5 padded to a 64-wide atom, and dy[M,5] / W^T are split into planes first — 0.2 ms per step for 0.4 GFLOP.  It is a
// streaming problem: x[M,K] is read once (forward), once more with dy (backward), dx[M,K] is written once; 57 MB each at
// the training shape.  One warp per row, lane = 16-byte chunks of K (coalesced), W in shared memory.
//   forward   y[r,o]  = sum_k x[r,k] W[o,k]                      (per-lane partial dots, butterfly reduce)
//   backward  dx[r,k] = sum_o dy[r,o] W[o,k];   dW[o,k] += sum_r dy[r,o] x[r,k]   (per-warp register accumulators over its
//             rows -> per-CTA shared-memory reduce -> float reductions into the zeroed dW)
#include "common.cuh"

namespace dg {

constexpr int kLnMaxNo = 8;
constexpr int kLnWarps = 8;

template <int J>      // J = ceil(K / 128) chunks of 4 floats per lane
__global__ void __launch_bounds__(32 * kLnWarps)
linear_narrow_fwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w, int ldw, int M, int K, int NO,
                         float* __restrict__ y, int ldy) {
  extern __shared__ __align__(16) float ws[];          // [NO][J * 128], zero padded beyond K
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int KP = J * 128;
  for (int i = threadIdx.x; i < NO * KP; i += blockDim.x) {
    const int o = i / KP, k = i - o * KP;
    ws[i] = k < K ? w[static_cast<size_t>(o) * ldw + k] : 0.f;
  }
  __syncthreads();
#pragma unroll 2
  for (int r = blockIdx.x * kLnWarps + warp; r < M; r += gridDim.x * kLnWarps) {
    float4 xv[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int k = 4 * (lane + 32 * j);
      xv[j] = k < K ? *reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * ldx + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float mine = 0.f;
    for (int o = 0; o < NO; ++o) {
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const float4 wv = *reinterpret_cast<const float4*>(ws + o * KP + 4 * (lane + 32 * j));
        a = __fmaf_rn(xv[j].x, wv.x, a); a = __fmaf_rn(xv[j].y, wv.y, a);
        a = __fmaf_rn(xv[j].z, wv.z, a); a = __fmaf_rn(xv[j].w, wv.w, a);
      }
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
      if (lane == o) mine = a;
    }
    if (lane < NO) y[static_cast<size_t>(r) * ldy + lane] = mine;
  }
}

template <int J, int NOT>      // NOT = compile-time bound of NO (register accumulators NOT x J float4)
__global__ void __launch_bounds__(32 * kLnWarps)
linear_narrow_bwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int lddy, const float* __restrict__ w,
                         int ldw, int M, int K, int NO, float* __restrict__ dx, int lddx, float* __restrict__ dw) {
  extern __shared__ __align__(16) float ws[];          // [NO][KP] weights, then reused for the dW reduce
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int KP = J * 128;
  for (int i = threadIdx.x; i < NO * KP; i += blockDim.x) {
    const int o = i / KP, k = i - o * KP;
    ws[i] = k < K ? w[static_cast<size_t>(o) * ldw + k] : 0.f;
  }
  __syncthreads();
  float4 acc[NOT][J];
#pragma unroll
  for (int o = 0; o < NOT; ++o)
#pragma unroll
    for (int j = 0; j < J; ++j) acc[o][j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
  for (int r = blockIdx.x * kLnWarps + warp; r < M; r += gridDim.x * kLnWarps) {
    float g[NOT];
#pragma unroll
    for (int o = 0; o < NOT; ++o) g[o] = o < NO ? dy[static_cast<size_t>(r) * lddy + o] : 0.f;     // warp-uniform addresses: broadcast
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int k = 4 * (lane + 32 * j);
      if (k < K) {
        const float4 xv = *reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * ldx + k);
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int o = 0; o < NOT; ++o) {
          if (o < NO) {
            const float4 wv = *reinterpret_cast<const float4*>(ws + o * KP + k);
            d.x = __fmaf_rn(g[o], wv.x, d.x); d.y = __fmaf_rn(g[o], wv.y, d.y);
            d.z = __fmaf_rn(g[o], wv.z, d.z); d.w = __fmaf_rn(g[o], wv.w, d.w);
            acc[o][j].x = __fmaf_rn(g[o], xv.x, acc[o][j].x); acc[o][j].y = __fmaf_rn(g[o], xv.y, acc[o][j].y);
            acc[o][j].z = __fmaf_rn(g[o], xv.z, acc[o][j].z); acc[o][j].w = __fmaf_rn(g[o], xv.w, acc[o][j].w);
          }
        }
        if (dx) *reinterpret_cast<float4*>(dx + static_cast<size_t>(r) * lddx + k) = d;
      }
    }
  }
  // dW: warp accumulators -> shared memory (one [NO][KP] slab, warps add in turn: fixed order inside the CTA) -> global
  __syncthreads();
  for (int i = threadIdx.x; i < NO * KP; i += blockDim.x) ws[i] = 0.f;
  __syncthreads();
  for (int wv = 0; wv < kLnWarps; ++wv) {
    if (warp == wv) {
#pragma unroll
      for (int o = 0; o < NOT; ++o)
#pragma unroll
        for (int j = 0; j < J; ++j)
          if (o < NO) {
            float4* p = reinterpret_cast<float4*>(ws + o * KP + 4 * (lane + 32 * j));
            float4 t = *p;
            t.x += acc[o][j].x; t.y += acc[o][j].y; t.z += acc[o][j].z; t.w += acc[o][j].w;
            *p = t;
          }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < NO * KP; i += blockDim.x) {
    const int o = i / KP, k = i - o * KP;
    if (k < K) atomicAdd(dw + static_cast<size_t>(o) * K + k, ws[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// THIN input (K <= 8: the point-level GEMM of the input block, X[M,4] [Wa1 ; Wa2 - Wa1]^T -> [U|V][M,128]).  On the
// tensor-core path K = 4 is padded to a 64-wide atom (operand planes of 56000 x 64 for 56000 x 4 numbers); here the row is
// one 16/32-byte load.
//   forward   y[r, o] = sum_c x[r,c] W[o,c]            thread = (row, 4 outputs), W in shared memory
//   backward  dW[o,c] = sum_r g[r,o] x[r,c]            lane = 4 outputs, 4 x KT register accumulators per lane over the warp's
//             rows -> shared-memory reduce over the CTA's warps -> float reductions into the zeroed dW;
//             dx[r,c] = sum_o g[r,o] W[o,c] (optional: the input cloud rarely needs a gradient) by a butterfly over the lanes
// ---------------------------------------------------------------------------------------------
template <int KT>      // KT = 4 or 8 input columns (zero padded)
__global__ void __launch_bounds__(256)
linear_thin_fwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w, int ldw, int M, int K, int NO,
                       float* __restrict__ y, int ldy) {
  extern __shared__ __align__(16) float ws[];          // [NO][KT]
  for (int i = threadIdx.x; i < NO * KT; i += blockDim.x) {
    const int o = i / KT, c = i - o * KT;
    ws[i] = c < K ? w[static_cast<size_t>(o) * ldw + c] : 0.f;
  }
  __syncthreads();
  const int o4n = NO >> 2;
  const long long total = static_cast<long long>(M) * o4n;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / o4n), o = 4 * static_cast<int>(i - static_cast<long long>(r) * o4n);
    float xv[KT];
#pragma unroll
    for (int c = 0; c < KT; c += 4) {
      const float4 v = *reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * ldx + c);
      xv[c] = v.x; xv[c + 1] = v.y; xv[c + 2] = v.z; xv[c + 3] = v.w;
    }
    float out[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float a = 0.f;
#pragma unroll
      for (int c = 0; c < KT; ++c) a = __fmaf_rn(xv[c], ws[(o + q) * KT + c], a);
      out[q] = a;
    }
    *reinterpret_cast<float4*>(y + static_cast<size_t>(r) * ldy + o) = make_float4(out[0], out[1], out[2], out[3]);
  }
}

template <int KT>      // NO == 128: lane = outputs 4 lane .. 4 lane + 3
__global__ void __launch_bounds__(256)
linear_thin_bwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ g, int ldg, const float* __restrict__ w, int ldw,
                       int M, int K, float* __restrict__ dx, int lddx, float* __restrict__ dw) {
  __shared__ __align__(16) float red[8][128 * KT];
  __shared__ __align__(16) float ws[128 * KT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 128 * KT; i += blockDim.x) {
    const int o = i / KT, c = i - o * KT;
    ws[i] = c < K ? w[static_cast<size_t>(o) * ldw + c] : 0.f;
  }
  __syncthreads();
  float acc[4][KT];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int c = 0; c < KT; ++c) acc[q][c] = 0.f;
#pragma unroll 2
  for (int r = blockIdx.x * 8 + warp; r < M; r += gridDim.x * 8) {
    const float4 gv = *reinterpret_cast<const float4*>(g + static_cast<size_t>(r) * ldg + 4 * lane);
    const float gq[4] = {gv.x, gv.y, gv.z, gv.w};
    float xv[KT];
#pragma unroll
    for (int c = 0; c < KT; c += 4) {
      const float4 v = *reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * ldx + c);
      xv[c] = v.x; xv[c + 1] = v.y; xv[c + 2] = v.z; xv[c + 3] = v.w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int c = 0; c < KT; ++c) acc[q][c] = __fmaf_rn(gq[q], xv[c], acc[q][c]);
    if (dx) {
      float d[KT];
#pragma unroll
      for (int c = 0; c < KT; ++c) {
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) a = __fmaf_rn(gq[q], ws[(4 * lane + q) * KT + c], a);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
        d[c] = a;
      }
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < KT; c += 4)
          if (c < lddx) *reinterpret_cast<float4*>(dx + static_cast<size_t>(r) * lddx + c) = make_float4(d[c], d[c + 1], d[c + 2], d[c + 3]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int c = 0; c < KT; ++c) red[warp][(4 * lane + q) * KT + c] = acc[q][c];
  __syncthreads();
  for (int i = threadIdx.x; i < 128 * KT; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) t += red[wv][i];
    const int o = i / KT, c = i - o * KT;
    if (c < K) atomicAdd(dw + static_cast<size_t>(o) * K + c, t);
  }
}

static int ln_check(const void* x, int ldx, const void* w, int ldw, int M, int K, int NO, const char* who) {
  DG_REQUIRE(x && w, DGCNN_E_INVALID, "%s: null pointer", who);
  DG_REQUIRE(M > 0 && K > 0 && NO > 0 && NO <= kLnMaxNo && K % 4 == 0 && K <= 512 && ldx >= K && ldx % 4 == 0 && ldw >= K,
             DGCNN_E_UNSUPPORTED, "%s: needs 1 <= outputs <= 8, K %% 4 == 0, K <= 512, ld_x %% 4 == 0 (M=%d K=%d outputs=%d)", who, M, K, NO);
  DG_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, DGCNN_E_INVALID, "%s: x must be 16-byte aligned", who);
  return DGCNN_OK;
}

}  // namespace dg

using namespace dg;

extern "C" int dgcnn_linear_narrow_supported(int K, int NO) { return (NO >= 1 && NO <= kLnMaxNo && K % 4 == 0 && K >= 4 && K <= 512) ? 1 : 0; }

extern "C" int dgcnn_linear_narrow_fwd_f32(const float* x, int ld_x, const float* w, int ld_w, int M, int K, int NO, float* y, int ld_y,
                                           void* stream) {
  int rc = ln_check(x, ld_x, w, ld_w, M, K, NO, "linear_narrow_fwd");
  if (rc) return rc;
  DG_REQUIRE(y && ld_y >= NO, DGCNN_E_INVALID, "linear_narrow_fwd: bad output");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int J = (K + 127) / 128;
  const int blocks = std::max(1, std::min((M + kLnWarps - 1) / kLnWarps, 4 * std::max(1, num_sms())));
  const size_t smem = static_cast<size_t>(NO) * (J <= 1 ? 1 : (J <= 2 ? 2 : 4)) * 128 * 4;
  if (J <= 1) linear_narrow_fwd_kernel<1><<<blocks, 32 * kLnWarps, smem, st>>>(x, ld_x, w, ld_w, M, K, NO, y, ld_y);
  else if (J <= 2) linear_narrow_fwd_kernel<2><<<blocks, 32 * kLnWarps, smem, st>>>(x, ld_x, w, ld_w, M, K, NO, y, ld_y);
  else linear_narrow_fwd_kernel<4><<<blocks, 32 * kLnWarps, smem, st>>>(x, ld_x, w, ld_w, M, K, NO, y, ld_y);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_linear_narrow_bwd_f32(const float* x, int ld_x, const float* dy, int ld_dy, const float* w, int ld_w, int M, int K,
                                           int NO, float* dx, int ld_dx, float* dw, void* stream) {
  int rc = ln_check(x, ld_x, w, ld_w, M, K, NO, "linear_narrow_bwd");
  if (rc) return rc;
  DG_REQUIRE(dy && dw && ld_dy >= NO && (!dx || (ld_dx >= K && ld_dx % 4 == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0)),
             DGCNN_E_INVALID, "linear_narrow_bwd: bad dy / dx / dw");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int J = (K + 127) / 128;
  DG_CUDA(cudaMemsetAsync(dw, 0, static_cast<size_t>(NO) * K * 4, st));
  const int blocks = std::max(1, std::min((M + kLnWarps - 1) / kLnWarps, 4 * std::max(1, num_sms())));
  const size_t smem = static_cast<size_t>(NO) * (J <= 1 ? 1 : (J <= 2 ? 2 : 4)) * 128 * 4;
  // register accumulators NOT x J float4: 8 x 2 (K <= 256) or 4 x 4 (K <= 512, <= 4 outputs)
  if (J <= 1) linear_narrow_bwd_kernel<1, 8><<<blocks, 32 * kLnWarps, smem, st>>>(x, ld_x, dy, ld_dy, w, ld_w, M, K, NO, dx, ld_dx, dw);
  else if (J <= 2) linear_narrow_bwd_kernel<2, 8><<<blocks, 32 * kLnWarps, smem, st>>>(x, ld_x, dy, ld_dy, w, ld_w, M, K, NO, dx, ld_dx, dw);
  else {
    DG_REQUIRE(NO <= 4, DGCNN_E_UNSUPPORTED, "linear_narrow_bwd: K > 256 needs <= 4 outputs");
    linear_narrow_bwd_kernel<4, 4><<<blocks, 32 * kLnWarps, smem, st>>>(x, ld_x, dy, ld_dy, w, ld_w, M, K, NO, dx, ld_dx, dw);
  }
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_linear_thin_supported(int K, int NO) { return (K >= 1 && K <= 8 && NO % 4 == 0 && NO >= 4 && NO <= 1024) ? 1 : 0; }

/* y[M,NO] = x[M,K] W[NO,K]^T for K <= 8 (x rows: ld_x >= round4(K) floats, 16-byte aligned, columns beyond K zero). */
extern "C" int dgcnn_linear_thin_fwd_f32(const float* x, int ld_x, const float* w, int ld_w, int M, int K, int NO, float* y, int ld_y,
                                         void* stream) {
  DG_REQUIRE(x && w && y, DGCNN_E_INVALID, "linear_thin_fwd: null pointer");
  DG_REQUIRE(M > 0 && dgcnn_linear_thin_supported(K, NO) && ld_x >= (K + 3) / 4 * 4 && ld_x % 4 == 0 && ld_w >= K && ld_y >= NO &&
                 ld_y % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0,
             DGCNN_E_UNSUPPORTED, "linear_thin_fwd: needs K <= 8, outputs %% 4 == 0, 16-byte aligned rows (M=%d K=%d outputs=%d)", M, K, NO);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long items = static_cast<long long>(M) * (NO / 4);
  const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>((items + 255) / 256, 16LL * std::max(1, num_sms()))));
  if (K <= 4) linear_thin_fwd_kernel<4><<<blocks, 256, static_cast<size_t>(NO) * 4 * 4, st>>>(x, ld_x, w, ld_w, M, K, NO, y, ld_y);
  else linear_thin_fwd_kernel<8><<<blocks, 256, static_cast<size_t>(NO) * 8 * 4, st>>>(x, ld_x, w, ld_w, M, K, NO, y, ld_y);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

/* dW[128,K] = g[M,128]^T x[M,K] (zeroed, then accumulated with float reductions); dx[M, round4(K)] = g W (optional). */
extern "C" int dgcnn_linear_thin_bwd_f32(const float* x, int ld_x, const float* g, int ld_g, const float* w, int ld_w, int M, int K,
                                         int NO, float* dx, int ld_dx, float* dw, void* stream) {
  DG_REQUIRE(x && g && w && dw, DGCNN_E_INVALID, "linear_thin_bwd: null pointer");
  DG_REQUIRE(M > 0 && K >= 1 && K <= 8 && NO == 128 && ld_x >= (K + 3) / 4 * 4 && ld_x % 4 == 0 && ld_g >= 128 && ld_g % 4 == 0 &&
                 ld_w >= K && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g)) & 15) == 0,
             DGCNN_E_UNSUPPORTED, "linear_thin_bwd: needs K <= 8, 128 outputs, 16-byte aligned rows (M=%d K=%d outputs=%d)", M, K, NO);
  DG_REQUIRE(!dx || (ld_dx >= (K + 3) / 4 * 4 && ld_dx % 4 == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0), DGCNN_E_INVALID,
             "linear_thin_bwd: dx needs ld_dx >= round4(K), ld_dx %% 4 == 0, 16-byte alignment");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DG_CUDA(cudaMemsetAsync(dw, 0, static_cast<size_t>(NO) * K * 4, st));
  const int blocks = std::max(1, std::min((M + 7) / 8, 4 * std::max(1, num_sms())));
  if (K <= 4) linear_thin_bwd_kernel<4><<<blocks, 256, 0, st>>>(x, ld_x, g, ld_g, w, ld_w, M, K, dx, ld_dx, dw);
  else linear_thin_bwd_kernel<8><<<blocks, 256, 0, st>>>(x, ld_x, g, ld_g, w, ld_w, M, K, dx, ld_dx, dw);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}
