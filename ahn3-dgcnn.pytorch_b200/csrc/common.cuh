// common.cuh — shared device helpers for libdgcnn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dgcnn_b200.h"

namespace dg {

constexpr int kCout = DGCNN_COUT;        // 64 output channels of every EdgeConv conv
constexpr int kTileRows = 128;           // edge rows per CTA tile
constexpr int kThreads = 256;
constexpr float kSlope = 0.2f;           // LeakyReLU negative slope (model.py:118)
constexpr int kMaxPartialCtas = 1024;    // upper bound on persistent grid size

// ---- host-side error plumbing ------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int num_sms();

#define DG_REQUIRE(cond, code, ...)   \
  do {                                \
    if (!(cond)) {                    \
      dg::set_error(__VA_ARGS__);     \
      return (code);                  \
    }                                 \
  } while (0)

#define DG_CUDA(call)                                         \
  do {                                                        \
    cudaError_t e__ = (call);                                 \
    if (e__ != cudaSuccess) return dg::cuda_fail(e__, #call); \
  } while (0)

// ---- device helpers -----------------------------------------------------------------------------
// slope < 1: max(z, slope z) is z for z > 0 and slope z otherwise, bit for bit (also for +-0 and NaN); one FMUL + one FMNMX
__device__ __forceinline__ float lrelu(float z) { return fmaxf(z, kSlope * z); }
__device__ __forceinline__ float lrelu_grad(float z) { return z > 0.f ? 1.f : kSlope; }

// Order-preserving map float -> uint32 (larger float => larger uint).
__device__ __forceinline__ uint32_t ordered_bits(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// Selection key: larger is better; ties on score prefer the LOWER index.
__device__ __forceinline__ unsigned long long make_key(float score, int idx) {
  return (static_cast<unsigned long long>(ordered_bits(score)) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(idx));
}
__device__ __forceinline__ int key_index(unsigned long long key) {
  return static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull));
}

// ---- FP32 register-tiled GEMM micro-kernels over shared-memory tiles ---------------------------
// 256 threads; tx = tid & 15, ty = tid >> 4.
//
// gemm_rr: C[128 x 64] += A[128 x K] * B[64 x K]^T, both operands row-major ("K-major") in smem with
// row strides LDA / LDB (floats, multiple of 4, LD % 32 == 4 keeps 16-byte accesses conflict-free).
// Thread owns rows ty + 16*i (i < 8) and columns tx + 16*q (q < 4).
// The contraction index runs in INCREASING order with one FMA chain per output: this is the summation
// order of the reference's CPU sgemm (kNN), so kNN scores are reproduced bit-for-bit.
template <int K, int LDA, int LDB>
__device__ __forceinline__ void gemm_rr(const float* __restrict__ As, const float* __restrict__ Bs,
                                        float (&acc)[8][4], int ty, int tx) {
#pragma unroll 2
  for (int kk = 0; kk < K; kk += 4) {
    float4 b[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) b[q] = *reinterpret_cast<const float4*>(Bs + (tx + 16 * q) * LDB + kk);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 a = *reinterpret_cast<const float4*>(As + (ty + 16 * i) * LDA + kk);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float c = acc[i][q];
        c = __fmaf_rn(a.x, b[q].x, c);
        c = __fmaf_rn(a.y, b[q].y, c);
        c = __fmaf_rn(a.z, b[q].z, c);
        c = __fmaf_rn(a.w, b[q].w, c);
        acc[i][q] = c;
      }
    }
  }
}

// gemm_kk: C[64 x 64] += sum_r A[r][m] * B[r][n] over r < R, both operands stored [r][64+] (contraction
// index outermost).  Thread owns m = 4*ty' .. +3 (ty' = tid >> 4), n = 4*tx' .. +3 (tx' = tid & 15).
template <int LDA, int LDB>
__device__ __forceinline__ void gemm_kk(const float* __restrict__ As, const float* __restrict__ Bs, int R,
                                        float (&acc)[4][4], int ty, int tx) {
#pragma unroll 4
  for (int r = 0; r < R; ++r) {
    const float4 a = *reinterpret_cast<const float4*>(As + r * LDA + 4 * ty);
    const float4 b = *reinterpret_cast<const float4*>(Bs + r * LDB + 4 * tx);
    const float av[4] = {a.x, a.y, a.z, a.w};
    const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int n = 0; n < 4; ++n) acc[m][n] = __fmaf_rn(av[m], bv[n], acc[m][n]);
  }
}

}  // namespace dg
