// seg_loss.cu — the two per-point passes that sit between the model's logits and the drivers' bookkeeping:
//   * dgcnn_seg_loss_f32      train.py:180-197 + util.py:17-34 — softmax over the classes, the per-sample mask ("focus"
//                             class num_classes for unmasked points), label-smoothed cross entropy applied to the
//                             PROBABILITIES (the reference's softmax-then-log_softmax quirk), mean over all points;
//                             forward and backward fused: one pass emits the loss partials and d(loss)/d(logits);
//   * dgcnn_seg_predict_f32   test.py:123-127 / test.py:305-308 — softmax, max over the classes, points whose best
//                             probability is below min_class_confidence become UNCLASSIFIED (31).
// The reference does both with Python loops over every point of every sample on the host (O(B*N) interpreter steps per
// training step); here a point is one thread, the class channels are read with the tensor's own strides ([B,C,N] as the
// model returns it, coalesced over n), and the loss is reduced in a fixed order (FP64 partials), so it is deterministic.
#include "common.cuh"

namespace dg {

constexpr int kMaxClasses = 32;

// One point: q[0..K) = the criterion's input row, t = smoothed target, returns the loss term; writes r = softmax(q).
template <int KMAX>
__device__ __forceinline__ float smoothed_ce_row(const float (&q)[KMAX], int K, int label, float eps, float (&r)[KMAX]) {
  float m = q[0];
#pragma unroll
  for (int c = 1; c < KMAX; ++c)
    if (c < K) m = fmaxf(m, q[c]);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < KMAX; ++c)
    if (c < K) {
      r[c] = expf(q[c] - m);
      s += r[c];
    }
  const float lse = m + logf(s), inv = 1.0f / s;
  const float t_on = 1.0f - eps, t_off = eps / static_cast<float>(K - 1);
  float loss = 0.f;
#pragma unroll
  for (int c = 0; c < KMAX; ++c)
    if (c < K) {
      const float t = c == label ? t_on : t_off;
      loss -= t * (q[c] - lse);
      r[c] *= inv;
    }
  return loss;
}

template <int CMAX>
__global__ void __launch_bounds__(256)
seg_loss_kernel(const float* __restrict__ logits, long long sb, long long sc, long long sn, int B, int C, int N,
                const long long* __restrict__ labels, const unsigned char* __restrict__ mask, float eps, float scale,
                float* __restrict__ dlogits, long long dsb, long long dsc, long long dsn, int* __restrict__ pred,
                double* __restrict__ partials) {
  __shared__ double red[8];
  const long long M = static_cast<long long>(B) * N;
  double acc = 0.0;
  for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < M; p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(p / N), n = static_cast<int>(p - static_cast<long long>(b) * N);
    const float* z = logits + b * sb + n * sn;
    // softmax over the classes (train.py:180)
    float prob[CMAX + 1];
    float m = z[0];
#pragma unroll
    for (int c = 1; c < CMAX; ++c)
      if (c < C) m = fmaxf(m, z[c * sc]);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) {
        prob[c] = expf(z[c * sc] - m);
        s += prob[c];
      }
    const float inv = 1.0f / s;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) prob[c] = c < C ? prob[c] * inv : 0.f;
    prob[CMAX] = 0.f;
    const bool focus = mask != nullptr;                 // train form: one extra class for the points outside the mask
    const bool on = !focus || mask[p] != 0;
    const int K = focus ? C + 1 : C;
    int label = static_cast<int>(labels[p]);
    float q[CMAX + 1];
#pragma unroll
    for (int c = 0; c <= CMAX; ++c) q[c] = on ? prob[c] : 0.f;          // masked: [p, 0]; unmasked: [0, ..., 0, 1]
    if (focus && !on) {
      // (the dynamic index C is resolved by the unrolled compare below)
#pragma unroll
      for (int c = 0; c <= CMAX; ++c)
        if (c == C) q[c] = 1.0f;
      label = C;
    }
    float r[CMAX + 1];
    acc += static_cast<double>(smoothed_ce_row<CMAX + 1>(q, K, label, eps, r));
    if (pred) {                                           // labels_pred.max(dim=2)[1] over the K channels (first maximum)
      int best = 0;
      float bv = q[0];
#pragma unroll
      for (int c = 1; c <= CMAX; ++c)
        if (c < K && q[c] > bv) { bv = q[c]; best = c; }
      pred[p] = best;
    }
    if (dlogits) {
      float* dz = dlogits + b * dsb + n * dsn;
      if (on) {
        // dq = softmax(q) - t (the targets sum to one); dz = p * (dq - <p, dq>) through the first softmax
        const float t_on = 1.0f - eps, t_off = eps / static_cast<float>(K - 1);
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < C) {
            r[c] -= c == label ? t_on : t_off;
            dot = __fmaf_rn(prob[c], r[c], dot);
          }
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < C) dz[c * dsc] = scale * prob[c] * (r[c] - dot);
      } else {
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < C) dz[c * dsc] = 0.f;
      }
    }
  }
  // fixed-order block reduction, one FP64 partial per CTA
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partials[blockIdx.x] = t;
  }
}

__global__ void seg_loss_finish_kernel(const double* __restrict__ partials, int n, double inv_count, float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < n; ++i) t += partials[i];
    *loss = static_cast<float>(t * inv_count);
  }
}

template <int CMAX>
__global__ void __launch_bounds__(256)
seg_predict_kernel(const float* __restrict__ logits, long long sb, long long sc, long long sn, int B, int C, int N,
                   float min_conf, long long unclassified, long long* __restrict__ pred, float* __restrict__ conf) {
  const long long M = static_cast<long long>(B) * N;
  for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < M; p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(p / N), n = static_cast<int>(p - static_cast<long long>(b) * N);
    const float* z = logits + b * sb + n * sn;
    float e[CMAX];
    float m = z[0];
#pragma unroll
    for (int c = 1; c < CMAX; ++c)
      if (c < C) m = fmaxf(m, z[c * sc]);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) {
        e[c] = expf(z[c * sc] - m);
        s += e[c];
      }
    // softmax(...).max(dim): value and FIRST index of the maximum probability (test.py:123-124)
    const float inv = 1.0f / s;
    int best = 0;
    float bv = e[0] * inv;
#pragma unroll
    for (int c = 1; c < CMAX; ++c)
      if (c < C) {
        const float v = e[c] * inv;
        if (v > bv) { bv = v; best = c; }
      }
    pred[p] = bv < min_conf ? unclassified : static_cast<long long>(best);      // test.py:126
    if (conf) conf[p] = bv;
  }
}

}  // namespace dg

using namespace dg;

extern "C" size_t dgcnn_seg_loss_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return static_cast<size_t>(kMaxPartialCtas) * sizeof(double);
}

extern "C" int dgcnn_seg_loss_f32(const float* logits, int64_t stride_b, int64_t stride_c, int64_t stride_n, int B, int C, int N,
                                  const int64_t* labels, const uint8_t* mask, float smoothing, float grad_scale, float* loss,
                                  float* dlogits, int64_t dstride_b, int64_t dstride_c, int64_t dstride_n, int32_t* pred,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  DG_REQUIRE(logits && labels && loss && workspace, DGCNN_E_INVALID, "seg_loss: null pointer");
  DG_REQUIRE(B > 0 && N > 0 && C >= 2, DGCNN_E_INVALID, "seg_loss: bad shape B=%d C=%d N=%d", B, C, N);
  DG_REQUIRE(C < kMaxClasses, DGCNN_E_UNSUPPORTED, "seg_loss: at most %d classes (C=%d)", kMaxClasses - 1, C);
  DG_REQUIRE(workspace_bytes >= dgcnn_seg_loss_workspace_bytes(B, N), DGCNN_E_WORKSPACE, "seg_loss: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long M = static_cast<long long>(B) * N;
  const int grid = static_cast<int>(std::min<long long>((M + 255) / 256, std::min(kMaxPartialCtas, 8 * std::max(1, num_sms()))));
  double* partials = static_cast<double*>(workspace);
  const float scale = grad_scale / static_cast<float>(M);
  const long long* lab = reinterpret_cast<const long long*>(labels);
  if (C <= 7)
    seg_loss_kernel<7><<<grid, 256, 0, st>>>(logits, stride_b, stride_c, stride_n, B, C, N, lab, mask, smoothing, scale, dlogits,
                                             dstride_b, dstride_c, dstride_n, pred, partials);
  else if (C <= 15)
    seg_loss_kernel<15><<<grid, 256, 0, st>>>(logits, stride_b, stride_c, stride_n, B, C, N, lab, mask, smoothing, scale, dlogits,
                                              dstride_b, dstride_c, dstride_n, pred, partials);
  else
    seg_loss_kernel<31><<<grid, 256, 0, st>>>(logits, stride_b, stride_c, stride_n, B, C, N, lab, mask, smoothing, scale, dlogits,
                                              dstride_b, dstride_c, dstride_n, pred, partials);
  DG_CUDA(cudaGetLastError());
  seg_loss_finish_kernel<<<1, 32, 0, st>>>(partials, grid, 1.0 / static_cast<double>(M), loss);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_seg_predict_f32(const float* logits, int64_t stride_b, int64_t stride_c, int64_t stride_n, int B, int C, int N,
                                     float min_conf, int64_t unclassified, int64_t* pred, float* conf, void* stream) {
  DG_REQUIRE(logits && pred, DGCNN_E_INVALID, "seg_predict: null pointer");
  DG_REQUIRE(B > 0 && N > 0 && C >= 1, DGCNN_E_INVALID, "seg_predict: bad shape B=%d C=%d N=%d", B, C, N);
  DG_REQUIRE(C <= kMaxClasses, DGCNN_E_UNSUPPORTED, "seg_predict: at most %d classes (C=%d)", kMaxClasses, C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long M = static_cast<long long>(B) * N;
  const int grid = static_cast<int>(std::min<long long>((M + 255) / 256, 16 * std::max(1, num_sms())));
  long long* out = reinterpret_cast<long long*>(pred);
  if (C <= 8)
    seg_predict_kernel<8><<<grid, 256, 0, st>>>(logits, stride_b, stride_c, stride_n, B, C, N, min_conf, unclassified, out, conf);
  else
    seg_predict_kernel<32><<<grid, 256, 0, st>>>(logits, stride_b, stride_c, stride_n, B, C, N, min_conf, unclassified, out, conf);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}
