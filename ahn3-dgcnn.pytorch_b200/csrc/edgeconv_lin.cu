// edgeconv_lin.cu — the single-convolution EdgeConv block (conv5 + bn5 + LeakyReLU + max over k, model.py:128-130,161-163)
// without any per-edge GEMM and without any per-edge tensor in HBM.
//
// A 1x1 convolution of the edge feature is linear in the two points of the edge:
//     y[(i,j)] = W [x_j - x_i ; x_i] = Wa1 x_j + (Wa2 - Wa1) x_i = U_j + V_i,     U = X Wa1^T,  V = X (Wa2 - Wa1)^T
// so the k-fold redundant per-edge GEMM (27.5 GFLOP at the training shape) collapses into one point-level tensor-core GEMM
// X [Wa1 ; Wa2 - Wa1]^T (0.9 GFLOP).  With batch-statistics BatchNorm and LeakyReLU both monotone per channel,
//     max_k LReLU(BN(y)) = LReLU(BN(V_i + max_j U_j))   (min_j for a negative scale),
// so the block needs, per point and channel, only the extrema of U over the neighbours (+ their slots for the backward)
// and the batch sums of y and y^2: one gather pass over the L2-resident U (14 MB at the training shape), 16-byte loads,
// a 256-byte row per neighbour.  The backward recomputes y = U_j + V_i the same way (dgcnn_edgeconv_bwd_a_lin).
// Only used for the 64-channel feature-space blocks: features are O(1) there, so U_j + V_i loses at most a digit against
// forming x_j - x_i first (the input block keeps the explicit difference: absolute map coordinates would lose three).
#include "edgeconv.cuh"

namespace dg {

// thread = (point, 4 channels): 16 points per CTA pass; persistent CTAs; per-CTA FP64 partials (sum, sum of squares).
// STORE_Y: the first convolution of a two-conv block — writes y_a = U_j + V_i [E,64] for the conv_b pass and the
// backward (BatchNorm with batch statistics sits between the two convolutions, so y_a is needed per edge) instead of the
// extrema; the per-edge tensor-core GEMM (gather, bf16x3 split, 24 MMAs per tile) it replaces took 4x as long.
template <bool STORE_Y>
__global__ void __launch_bounds__(kThreads)
ec_lin_extrema_kernel(const float* __restrict__ U, int ld_u, const float* __restrict__ V, int ld_v, const int32_t* __restrict__ idx,
                      int M, int N, int k, float* __restrict__ ymax, float* __restrict__ ymin, uint8_t* __restrict__ amax,
                      uint8_t* __restrict__ amin, float* __restrict__ y_out, double* __restrict__ partials,
                      float* __restrict__ ysum, float* __restrict__ yin, float* __restrict__ cnt) {
  __shared__ float4 red[2][kThreads];
  const int c4 = threadIdx.x & 15, pl = threadIdx.x >> 4;
  const int c = 4 * c4;
  float4 t1 = make_float4(0.f, 0.f, 0.f, 0.f), t2 = t1;
  for (int gp = blockIdx.x * 16 + pl; gp < M; gp += gridDim.x * 16) {
    const size_t base = static_cast<size_t>(gp / N) * N;
    const float4 v = *reinterpret_cast<const float4*>(V + static_cast<size_t>(gp) * ld_v + c);
    const int32_t* nb = idx + static_cast<size_t>(gp) * k;
    float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), mn = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
    uchar4 ax = make_uchar4(0, 0, 0, 0), an = ax;
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
#pragma unroll 4
    for (int slot = 0; slot < k; ++slot) {
      const float4 u = *reinterpret_cast<const float4*>(U + (base + nb[slot]) * ld_u + c);
      float4 y;
      y.x = __fadd_rn(u.x, v.x); y.y = __fadd_rn(u.y, v.y); y.z = __fadd_rn(u.z, v.z); y.w = __fadd_rn(u.w, v.w);
      if (STORE_Y) {
        __stcs(reinterpret_cast<float4*>(y_out + (static_cast<size_t>(gp) * k + slot) * kCout + c), y);   // streaming: keep U / yin in L2
        if (yin) {      // graph-structure sums of the backward (see dgcnn_edgeconv_bwd_a_point): sum of y over the in-edges of j
          atomicAdd(reinterpret_cast<float4*>(yin + (base + nb[slot]) * kCout + c), y);
          if (c4 == 0) atomicAdd(cnt + base + nb[slot], 1.f);
        }
      } else {
        // first occurrence wins ties, like the per-edge kernels
        if (y.x > mx.x) { mx.x = y.x; ax.x = slot; }
        if (y.y > mx.y) { mx.y = y.y; ax.y = slot; }
        if (y.z > mx.z) { mx.z = y.z; ax.z = slot; }
        if (y.w > mx.w) { mx.w = y.w; ax.w = slot; }
        if (y.x < mn.x) { mn.x = y.x; an.x = slot; }
        if (y.y < mn.y) { mn.y = y.y; an.y = slot; }
        if (y.z < mn.z) { mn.z = y.z; an.z = slot; }
        if (y.w < mn.w) { mn.w = y.w; an.w = slot; }
      }
      s1.x += y.x; s1.y += y.y; s1.z += y.z; s1.w += y.w;
      s2.x = __fmaf_rn(y.x, y.x, s2.x); s2.y = __fmaf_rn(y.y, y.y, s2.y);
      s2.z = __fmaf_rn(y.z, y.z, s2.z); s2.w = __fmaf_rn(y.w, y.w, s2.w);
    }
    if (STORE_Y && ysum) *reinterpret_cast<float4*>(ysum + static_cast<size_t>(gp) * kCout + c) = s1;   // sum over the out-edges of gp
    if (!STORE_Y) {
      const size_t o = static_cast<size_t>(gp) * kCout + c;
      *reinterpret_cast<float4*>(ymax + o) = mx;
      *reinterpret_cast<float4*>(ymin + o) = mn;
      *reinterpret_cast<uchar4*>(amax + o) = ax;
      *reinterpret_cast<uchar4*>(amin + o) = an;
    }
    t1.x += s1.x; t1.y += s1.y; t1.z += s1.z; t1.w += s1.w;
    t2.x += s2.x; t2.y += s2.y; t2.z += s2.z; t2.w += s2.w;
  }
  if (!partials) return;
  // per-CTA statistics in a fixed order: 16 point lanes per channel group -> FP64
  red[0][threadIdx.x] = t1;
  red[1][threadIdx.x] = t2;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int w = threadIdx.x >> 4, cg = threadIdx.x & 15;
    double a[4] = {0, 0, 0, 0};
    for (int l = 0; l < 16; ++l) {
      const float4 q = red[w][l * 16 + cg];
      a[0] += q.x; a[1] += q.y; a[2] += q.z; a[3] += q.w;
    }
    double* dst = partials + 128 + static_cast<size_t>(blockIdx.x) * 128 + w * 64 + 4 * cg;
    dst[0] = a[0]; dst[1] = a[1]; dst[2] = a[2]; dst[3] = a[3];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) partials[0] = static_cast<double>(gridDim.x);
}

// backward edge pass through BN_a with y recomputed as U_j + V_i (an L2 gather instead of a 430 MB HBM read of y_a):
// dy = cA dz - cB - cC (U_j + V_i).  DENSE_DZ: dz is the per-edge dz_a of a two-conv block; otherwise (single conv) dz is
// non-zero at the arg rows only and comes from g / arg.
template <bool DENSE_DZ>
__global__ void __launch_bounds__(kThreads)
ec_lin_bwd_a_kernel(const float* __restrict__ U, int ld_u, const float* __restrict__ V, int ld_v, const int32_t* __restrict__ idx,
                    int M, int N, int k, const float* __restrict__ dz_a, const float* __restrict__ g, const uint8_t* __restrict__ arg,
                    const float* __restrict__ coef_a, float* __restrict__ s_out, float* __restrict__ r_out, int ld_rs,
                    float* __restrict__ dy_edges) {
  const int c4 = threadIdx.x & 15, pl = threadIdx.x >> 4;
  const int c = 4 * c4;
  const float4 cA = *reinterpret_cast<const float4*>(coef_a + c);
  const float4 cB = *reinterpret_cast<const float4*>(coef_a + 64 + c);
  const float4 cC = *reinterpret_cast<const float4*>(coef_a + 128 + c);
  for (int gp = blockIdx.x * 16 + pl; gp < M; gp += gridDim.x * 16) {
    const size_t base = static_cast<size_t>(gp / N) * N;
    float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
    uchar4 av = make_uchar4(255, 255, 255, 255);
    if (!DENSE_DZ) {
      gv = *reinterpret_cast<const float4*>(g + static_cast<size_t>(gp) * kCout + c);
      av = *reinterpret_cast<const uchar4*>(arg + static_cast<size_t>(gp) * kCout + c);
    }
    const float4 v = *reinterpret_cast<const float4*>(V + static_cast<size_t>(gp) * ld_v + c);
    const int32_t* nb = idx + static_cast<size_t>(gp) * k;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int slot = 0; slot < k; ++slot) {
      const size_t j = base + nb[slot];
      const float4 u = *reinterpret_cast<const float4*>(U + j * ld_u + c);
      float4 dz;
      if (DENSE_DZ) {
        dz = *reinterpret_cast<const float4*>(dz_a + (static_cast<size_t>(gp) * k + slot) * kCout + c);
      } else {
        dz.x = av.x == slot ? gv.x : 0.f; dz.y = av.y == slot ? gv.y : 0.f;
        dz.z = av.z == slot ? gv.z : 0.f; dz.w = av.w == slot ? gv.w : 0.f;
      }
      float4 dy;
      dy.x = cA.x * dz.x - cB.x - cC.x * __fadd_rn(u.x, v.x);
      dy.y = cA.y * dz.y - cB.y - cC.y * __fadd_rn(u.y, v.y);
      dy.z = cA.z * dz.z - cB.z - cC.z * __fadd_rn(u.z, v.z);
      dy.w = cA.w * dz.w - cB.w - cC.w * __fadd_rn(u.w, v.w);
      s.x += dy.x; s.y += dy.y; s.z += dy.z; s.w += dy.w;
      if (dy_edges)      // reproducible form: the per-edge gradient goes to HBM and is gathered in a fixed order (graph_transpose.cu)
        *reinterpret_cast<float4*>(dy_edges + (static_cast<size_t>(gp) * k + slot) * kCout + c) = dy;
      else
        atomicAdd(reinterpret_cast<float4*>(r_out + j * ld_rs + c), dy);
    }
    *reinterpret_cast<float4*>(s_out + static_cast<size_t>(gp) * ld_rs + c) = s;
  }
}

}  // namespace dg

using namespace dg;

static int lin_check(const dgcnn_edgeconv_args* a, const float* U, int ld_u, const float* V, int ld_v, const char* who) {
  DG_REQUIRE(a && a->idx && U && V, DGCNN_E_INVALID, "%s: null pointer", who);
  DG_REQUIRE(a->B > 0 && a->N > 0 && a->k > 0 && a->k <= 255, DGCNN_E_INVALID, "%s: bad shape", who);
  DG_REQUIRE(ld_u >= kCout && ld_v >= kCout && ld_u % 4 == 0 && ld_v % 4 == 0 &&
                 ((reinterpret_cast<uintptr_t>(U) | reinterpret_cast<uintptr_t>(V)) & 15) == 0,
             DGCNN_E_INVALID, "%s: U and V need 16-byte aligned rows of at least 64 floats", who);
  return DGCNN_OK;
}

extern "C" int dgcnn_edgeconv_lin_extrema(const dgcnn_edgeconv_args* a, const float* U, int ld_u, const float* V, int ld_v,
                                          double* partials, float* ymax, float* ymin, uint8_t* amax, uint8_t* amin, void* stream) {
  int rc = lin_check(a, U, ld_u, V, ld_v, "edgeconv_lin_extrema");
  if (rc) return rc;
  DG_REQUIRE(ymax && ymin && amax && amin, DGCNN_E_INVALID, "edgeconv_lin_extrema: null output");
  const int M = a->B * a->N;
  const int blocks = std::max(1, std::min((M + 15) / 16, std::min(kMaxPartialCtas, 8 * std::max(1, num_sms()))));
  ec_lin_extrema_kernel<false><<<blocks, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(U, ld_u, V, ld_v, a->idx, M, a->N, a->k, ymax,
                                                                                         ymin, amax, amin, nullptr, partials, nullptr, nullptr, nullptr);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

static int lin_conv_a(const dgcnn_edgeconv_args* a, const float* U, int ld_u, const float* V, int ld_v, float* y_a, double* partials,
                      float* ysum, float* yin, float* cnt, void* stream, const char* who) {
  int rc = lin_check(a, U, ld_u, V, ld_v, who);
  if (rc) return rc;
  DG_REQUIRE(y_a && partials && (reinterpret_cast<uintptr_t>(y_a) & 15) == 0, DGCNN_E_INVALID, "%s: null / misaligned output", who);
  const int M = a->B * a->N;
  const int blocks = std::max(1, std::min((M + 15) / 16, std::min(kMaxPartialCtas, 8 * std::max(1, num_sms()))));
  ec_lin_extrema_kernel<true><<<blocks, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(U, ld_u, V, ld_v, a->idx, M, a->N, a->k, nullptr,
                                                                                        nullptr, nullptr, nullptr, y_a, partials, ysum, yin, cnt);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_edgeconv_lin_conv_a(const dgcnn_edgeconv_args* a, const float* U, int ld_u, const float* V, int ld_v,
                                         float* y_a, double* partials, void* stream) {
  return lin_conv_a(a, U, ld_u, V, ld_v, y_a, partials, nullptr, nullptr, nullptr, stream, "edgeconv_lin_conv_a");
}

extern "C" int dgcnn_edgeconv_lin_conv_a_train(const dgcnn_edgeconv_args* a, const float* U, int ld_u, const float* V, int ld_v,
                                               float* y_a, double* partials, float* ysum, void* stream) {
  DG_REQUIRE(ysum && (reinterpret_cast<uintptr_t>(ysum) & 15) == 0, DGCNN_E_INVALID,
             "edgeconv_lin_conv_a_train: ysum must be non-null and 16-byte aligned");
  return lin_conv_a(a, U, ld_u, V, ld_v, y_a, partials, ysum, nullptr, nullptr, stream, "edgeconv_lin_conv_a_train");
}

static int bwd_a_lin(const dgcnn_edgeconv_args* a, const float* U, int ld_u, const float* V, int ld_v, const float* dz_a, const float* g,
                     const uint8_t* arg, const float* coef_a, float* s, float* r, int ld_rs, float* dy_edges, void* stream, const char* who) {
  int rc = lin_check(a, U, ld_u, V, ld_v, who);
  if (rc) return rc;
  DG_REQUIRE(coef_a && s && (r || dy_edges) && (dz_a || (g && arg)), DGCNN_E_INVALID, "%s: null pointer", who);
  DG_REQUIRE(!dz_a || (reinterpret_cast<uintptr_t>(dz_a) & 15) == 0, DGCNN_E_INVALID, "%s: dz_a must be 16-byte aligned", who);
  DG_REQUIRE(ld_rs >= kCout && ld_rs % 4 == 0 && ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(r) |
                                                   reinterpret_cast<uintptr_t>(dy_edges)) & 15) == 0,
             DGCNN_E_INVALID, "%s: s / r / dy_edges need 16-byte alignment and ld_rs %% 4 == 0, ld_rs >= 64", who);
  const int M = a->B * a->N;
  const int blocks = std::max(1, std::min((M + 15) / 16, 8 * std::max(1, num_sms())));
  if (dz_a)
    ec_lin_bwd_a_kernel<true><<<blocks, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(U, ld_u, V, ld_v, a->idx, M, a->N, a->k, dz_a,
                                                                                        g, arg, coef_a, s, r, ld_rs, dy_edges);
  else
    ec_lin_bwd_a_kernel<false><<<blocks, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(U, ld_u, V, ld_v, a->idx, M, a->N, a->k, dz_a,
                                                                                         g, arg, coef_a, s, r, ld_rs, dy_edges);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_edgeconv_bwd_a_lin(const dgcnn_edgeconv_args* a, const float* U, int ld_u, const float* V, int ld_v,
                                        const float* dz_a, const float* g, const uint8_t* arg, const float* coef_a, float* s,
                                        float* r, int ld_rs, void* stream) {
  DG_REQUIRE(r, DGCNN_E_INVALID, "edgeconv_bwd_a_lin: null r");
  return bwd_a_lin(a, U, ld_u, V, ld_v, dz_a, g, arg, coef_a, s, r, ld_rs, nullptr, stream, "edgeconv_bwd_a_lin");
}

/* the same pass with the per-edge gradient dy_a[E,64] written out instead of scattered (s as before): the reproducible form
 * gathers r from it in a fixed order (dgcnn_graph_transpose + dgcnn_gather_in_edges_f32) */
extern "C" int dgcnn_edgeconv_bwd_a_lin_edges(const dgcnn_edgeconv_args* a, const float* U, int ld_u, const float* V, int ld_v,
                                              const float* dz_a, const float* g, const uint8_t* arg, const float* coef_a, float* s,
                                              int ld_rs, float* dy_edges, void* stream) {
  DG_REQUIRE(dy_edges, DGCNN_E_INVALID, "edgeconv_bwd_a_lin_edges: null dy_edges");
  return bwd_a_lin(a, U, ld_u, V, ld_v, dz_a, g, arg, coef_a, s, nullptr, ld_rs, dy_edges, stream, "edgeconv_bwd_a_lin_edges");
}

// ---- the weight of the linearised first convolution and its adjoint (one launch each instead of slice/subtract/concatenate) ----
namespace dg {
__global__ void lin_weight_kernel(const float* __restrict__ wa, int C, float* __restrict__ wlin) {
  // wa [64, 2C] = [Wa1 | Wa2]  ->  wlin [128, C] = [Wa1 ; Wa2 - Wa1]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kCout * C) return;
  const int o = i / C, c = i - o * C;
  const float w1 = wa[o * 2 * C + c], w2 = wa[o * 2 * C + C + c];
  wlin[i] = w1;
  wlin[kCout * C + i] = w2 - w1;
}
__global__ void lin_weight_bwd_kernel(const float* __restrict__ G, int C, float* __restrict__ dwa) {
  // G [128, C] = [r | s]^T X  ->  dwa [64, 2C] = [ (r - s)^T X | s^T X ]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kCout * C) return;
  const int o = i / C, c = i - o * C;
  const float gr = G[i], gs = G[kCout * C + i];
  dwa[o * 2 * C + c] = gr - gs;
  dwa[o * 2 * C + C + c] = gs;
}
}  // namespace dg

extern "C" int dgcnn_edgeconv_lin_weight_f32(const float* wa, int C, float* wlin, void* stream) {
  DG_REQUIRE(wa && wlin && C > 0, DGCNN_E_INVALID, "edgeconv_lin_weight: bad arguments");
  dg::lin_weight_kernel<<<(kCout * C + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(wa, C, wlin);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_edgeconv_lin_weight_bwd_f32(const float* G, int C, float* dwa, void* stream) {
  DG_REQUIRE(G && dwa && C > 0, DGCNN_E_INVALID, "edgeconv_lin_weight_bwd: bad arguments");
  dg::lin_weight_bwd_kernel<<<(kCout * C + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(G, C, dwa);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}
