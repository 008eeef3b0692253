// head_bwd.cu — the sparse half of the backward of conv6 + bn6 + LeakyReLU + global max (model.py:172-175).
//
// The dense BN gradient is dy = a dz - b' - kappa y with y = feat W^T (see ops.GlobalFeature); dz is non-zero at the
// B * E arg-max entries only (one row per block and output channel).  The dense parts are two 192-wide GEMMs; the sparse
// parts, formerly a chain of library gathers (`w.repeat`, `index_add_`, `feat[idx]`, multiply, sum), are one pass here:
//     dfeat[n*(b,e), :] += a_e * gdz[b,e] * W[e, :]          (scatter; several channels may select the same row: atomics)
//     t1[e, :]           = sum_b gdz[b,e] * feat[n*(b,e), :]   (= dz^T feat, fixed order over b: deterministic)
// plus the vector-matrix product c[k] = sum_e v[e] W[e, k] (for the b' W row constant).
// One warp per output channel e; lane = slice of the K (<= 1024) feature columns.
#include "common.cuh"

namespace dg {

__global__ void __launch_bounds__(256)
global_feature_bwd_sparse_kernel(const float* __restrict__ feat, int ldf, const float* __restrict__ w, int ldw,
                                 const int* __restrict__ nstar, const float* __restrict__ gdz, const float* __restrict__ a,
                                 int B, int E, int K, float* __restrict__ dfeat, int lddf, float* __restrict__ t1) {
  const int lane = threadIdx.x & 31;
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (e >= E) return;
  const float ae = a[e];
  for (int k0 = 0; k0 < K; k0 += 128) {            // 4 columns per lane and sweep: 16-byte accesses
    const int k = k0 + 4 * lane;
    if (k >= K) continue;
    const float4 wv = *reinterpret_cast<const float4*>(w + static_cast<size_t>(e) * ldw + k);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) {
      const int row = nstar[b * E + e];
      const float g = gdz[b * E + e];
      const float4 f = *reinterpret_cast<const float4*>(feat + static_cast<size_t>(row) * ldf + k);
      acc.x = __fmaf_rn(g, f.x, acc.x); acc.y = __fmaf_rn(g, f.y, acc.y);
      acc.z = __fmaf_rn(g, f.z, acc.z); acc.w = __fmaf_rn(g, f.w, acc.w);
      const float s = ae * g;
      if (dfeat)      // (null: the reproducible form scatters separately, dgcnn_global_feature_scatter_det_f32)
        atomicAdd(reinterpret_cast<float4*>(dfeat + static_cast<size_t>(row) * lddf + k), make_float4(s * wv.x, s * wv.y, s * wv.z, s * wv.w));
    }
    *reinterpret_cast<float4*>(t1 + static_cast<size_t>(e) * K + k) = acc;
  }
}

// c[k] = scale * sum_e v[e] * W[e, k].  CTA = 32 columns x 32 row groups: group g adds rows g, g+32, ... (coalesced
// 128-byte reads), the 32 group sums are added in order (deterministic).  A thread per column walking all E rows was a
// 100 us latency chain.
__global__ void __launch_bounds__(1024) vecmat_kernel(const float* __restrict__ v, const float* __restrict__ w, int ldw, int E,
                                                      int K, float scale, float* __restrict__ c) {
  __shared__ float red[32][33];
  const int cl = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + cl;
  float a = 0.f;
  if (k < K)
    for (int e = g; e < E; e += 32) a = __fmaf_rn(v[e], w[static_cast<size_t>(e) * ldw + k], a);
  red[g][cl] = a;
  __syncthreads();
  if (g == 0 && k < K) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) t += red[j][cl];
    c[k] = scale * t;
  }
}

// Per-channel coefficients of that backward, formerly ~20 elementwise / reduction launches of the tensor library:
//   a = gamma rstd, yhat = (ysel - mean) rstd, gdz = dg * LReLU'(gamma yhat + beta),
//   s1 = sum_b gdz (= dbeta), s2 = sum_b gdz yhat (= dgamma), kappa = a (s2 / M) rstd, b' = a (s1 / M) - kappa mean,
//   wneg[e, :] = -kappa_e W[e, :]   (left operand of the K x K matrix -W^T diag(kappa) W)
// One warp per output channel e; the sums over the B blocks run in block order (deterministic).
__global__ void __launch_bounds__(256)
global_feature_bwd_coef_kernel(const float* __restrict__ dg, const float* __restrict__ ysel, const float* __restrict__ mean,
                               const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ w, int ldw, int B, int E, int K, float inv_m, float slope,
                               float* __restrict__ gdz, float* __restrict__ coef, float* __restrict__ wneg) {
  const int lane = threadIdx.x & 31;
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (e >= E) return;
  const float mu = mean[e], rs = rstd[e], ga = gamma[e], be = beta[e];
  float s1 = 0.f, s2 = 0.f;
  for (int b = 0; b < B; ++b) {                       // every lane walks the same B entries: no shuffle, fixed order
    const float yhat = (ysel[b * E + e] - mu) * rs;
    const float z = __fmaf_rn(yhat, ga, be);
    const float g = dg[b * E + e] * (z > 0.f ? 1.0f : slope);
    if (lane == 0) gdz[b * E + e] = g;
    s1 += g;
    s2 = __fmaf_rn(g, yhat, s2);
  }
  const float a = ga * rs;
  const float kappa = a * (s2 * inv_m) * rs;
  const float bprime = a * (s1 * inv_m) - kappa * mu;
  if (lane == 0) {
    coef[e] = a; coef[E + e] = kappa; coef[2 * E + e] = bprime; coef[3 * E + e] = s1; coef[4 * E + e] = s2;
  }
  for (int k = lane; k < K; k += 32) wneg[static_cast<size_t>(e) * K + k] = -kappa * w[static_cast<size_t>(e) * ldw + k];
}

// dW[e, k] = a_e t1[e, k] - b'_e colsum[k] - kappa_e wf2[e, k]   (the three terms of ops.GlobalFeature's weight gradient)
__global__ void __launch_bounds__(256)
global_feature_bwd_dw_kernel(const float* __restrict__ coef, const float* __restrict__ t1, const float* __restrict__ colsum,
                             const float* __restrict__ wf2, int E, int K, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E * K) return;
  const int e = i / K, k = i - e * K;
  dw[i] = coef[e] * t1[i] - coef[2 * E + e] * colsum[k] - coef[E + e] * wf2[i];
}

}  // namespace dg

using namespace dg;

extern "C" int dgcnn_global_feature_bwd_sparse_f32(const float* feat, int ld_feat, const float* w, int ld_w, const int32_t* nstar,
                                                   const float* gdz, const float* a, int B, int E, int K, float* dfeat,
                                                   int ld_dfeat, float* t1, void* stream) {
  DG_REQUIRE(feat && w && nstar && gdz && a && t1, DGCNN_E_INVALID, "global_feature_bwd_sparse: null pointer");
  DG_REQUIRE(B > 0 && E > 0 && K > 0 && K % 4 == 0 && ld_feat % 4 == 0 && ld_w % 4 == 0 && ld_dfeat % 4 == 0, DGCNN_E_INVALID,
             "global_feature_bwd_sparse: K and the row strides must be multiples of 4 (K=%d)", K);
  DG_REQUIRE(((reinterpret_cast<uintptr_t>(feat) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(dfeat) |
               reinterpret_cast<uintptr_t>(t1)) & 15) == 0, DGCNN_E_INVALID, "global_feature_bwd_sparse: 16-byte alignment required");
  const int warps_per_cta = 8;
  global_feature_bwd_sparse_kernel<<<(E + warps_per_cta - 1) / warps_per_cta, 32 * warps_per_cta, 0, static_cast<cudaStream_t>(stream)>>>(
      feat, ld_feat, w, ld_w, nstar, gdz, a, B, E, K, dfeat, ld_dfeat, t1);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_vecmat_f32(const float* v, const float* w, int ld_w, int E, int K, float scale, float* c, void* stream) {
  DG_REQUIRE(v && w && c && E > 0 && K > 0 && ld_w >= K, DGCNN_E_INVALID, "vecmat: bad arguments");
  vecmat_kernel<<<(K + 31) / 32, 1024, 0, static_cast<cudaStream_t>(stream)>>>(v, w, ld_w, E, K, scale, c);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_global_feature_bwd_coef_f32(const float* dg, const float* ysel, const float* mean, const float* rstd,
                                                 const float* gamma, const float* beta, const float* w, int ld_w, int B, int E,
                                                 int K, int M, float slope, float* gdz, float* coef, float* wneg, void* stream) {
  DG_REQUIRE(dg && ysel && mean && rstd && gamma && beta && w && gdz && coef && wneg, DGCNN_E_INVALID,
             "global_feature_bwd_coef: null pointer");
  DG_REQUIRE(B > 0 && E > 0 && K > 0 && M > 0 && ld_w >= K, DGCNN_E_INVALID, "global_feature_bwd_coef: bad sizes");
  const int warps_per_cta = 8;
  global_feature_bwd_coef_kernel<<<(E + warps_per_cta - 1) / warps_per_cta, 32 * warps_per_cta, 0, static_cast<cudaStream_t>(stream)>>>(
      dg, ysel, mean, rstd, gamma, beta, w, ld_w, B, E, K, 1.0f / static_cast<float>(M), slope, gdz, coef, wneg);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_global_feature_bwd_dw_f32(const float* coef, const float* t1, const float* colsum, const float* wf2, int E,
                                               int K, float* dw, void* stream) {
  DG_REQUIRE(coef && t1 && colsum && wf2 && dw && E > 0 && K > 0, DGCNN_E_INVALID, "global_feature_bwd_dw: bad arguments");
  global_feature_bwd_dw_kernel<<<(E * K + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(coef, t1, colsum, wf2, E, K, dw);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}
