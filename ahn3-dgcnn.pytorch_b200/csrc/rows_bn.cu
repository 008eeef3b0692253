// rows_bn.cu — BatchNorm1d (+ optional per-block bias) + LeakyReLU over point-major rows [M, C], forward and backward.
//
// Used by the shared-MLP head (reference model.py:131-139,176-187: conv7/conv8 -> bn7/bn8 -> LeakyReLU).  The library
// composition (add, batch_norm stats, transform, leaky_relu; backward: leaky_relu_backward, bn reduce, bn elemt) makes
// 5 + 8 passes over the [M, C] activations; these kernels make 2 + 2 (+1 for the per-block bias gradient):
//   forward   pass 1: per-CTA column sums of v = y + bias[block] and v^2  -> FP64 partials -> mean / rstd / running stats
//             pass 2: h = LReLU((v - mean) * gamma * rstd + beta)
//   backward  pass 1: dz = dh * LReLU'(z(v)); per-CTA column sums of dz and dz * vhat -> dgamma, dbeta
//             pass 2: dy = gamma * rstd * (dz - mean(dz) - vhat * mean(dz * vhat))
// All reductions are per-CTA partials in a fixed order (deterministic), finished in FP64.
// HBM-bound streaming kernels: 16-byte accesses, a warp covers 512 contiguous bytes of a row.
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc.cuh"

namespace dg {

constexpr int kRbThreads = 256;
constexpr int kRbMaxCtas = 1184;          // 148 SMs x 8
constexpr int kColsumMaxParts = 1024;     // CTAs per segment: a single segment (colsum of the whole batch) must still fill the GPU

struct RowsBn {
  const float* y; int M, C, N;            // rows, channels, points per block (bias row = row / N)
  const float* bias;                      // [M / N, C] or null
  const float* a;                         // float[3C]: mean | scale | beta
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// thread -> float4 column group cg = tid % (C/4), row lane rl = tid / (C/4); CTA b owns rows [b*rpc, (b+1)*rpc)
template <bool BWD>
__global__ void __launch_bounds__(kRbThreads) rows_bn_reduce_kernel(RowsBn p, const float* __restrict__ dh, int rows_per_cta,
                                                                   double* __restrict__ partials) {
  __shared__ float4 red[2][kRbThreads];
  const int c4n = p.C >> 2, RL = kRbThreads / c4n;
  const int cg = threadIdx.x % c4n, rl = threadIdx.x / c4n;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(p.M, r0 + rows_per_cta);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  float4 mean = s, scale = s, beta = s;
  if (BWD) {
    mean = ld4(p.a + 4 * cg);
    scale = ld4(p.a + p.C + 4 * cg);
    beta = ld4(p.a + 2 * p.C + 4 * cg);
  }
  for (int r = r0 + rl; r < r1; r += RL) {
    float4 v = ld4(p.y + static_cast<size_t>(r) * p.C + 4 * cg);
    if (p.bias) {
      const float4 b = ld4(p.bias + static_cast<size_t>(r / p.N) * p.C + 4 * cg);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (!BWD) {
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      q.x = __fmaf_rn(v.x, v.x, q.x); q.y = __fmaf_rn(v.y, v.y, q.y); q.z = __fmaf_rn(v.z, v.z, q.z); q.w = __fmaf_rn(v.w, v.w, q.w);
    } else {
      // s = sum dz, q = sum dz * (v - mean)   (the factor rstd of vhat is applied once in the finalize kernel)
      const float4 g = ld4(dh + static_cast<size_t>(r) * p.C + 4 * cg);
      const float dx = v.x - mean.x, dy = v.y - mean.y, dzc = v.z - mean.z, dw = v.w - mean.w;
      const float gx = g.x * lrelu_grad(__fmaf_rn(dx, scale.x, beta.x)), gy = g.y * lrelu_grad(__fmaf_rn(dy, scale.y, beta.y));
      const float gz = g.z * lrelu_grad(__fmaf_rn(dzc, scale.z, beta.z)), gw = g.w * lrelu_grad(__fmaf_rn(dw, scale.w, beta.w));
      s.x += gx; s.y += gy; s.z += gz; s.w += gw;
      q.x = __fmaf_rn(gx, dx, q.x); q.y = __fmaf_rn(gy, dy, q.y); q.z = __fmaf_rn(gz, dzc, q.z); q.w = __fmaf_rn(gw, dw, q.w);
    }
  }
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.x < c4n) {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int l = 0; l < RL; ++l) {           // fixed order over the row lanes
      const float4 a = red[0][l * c4n + cg], b = red[1][l * c4n + cg];
      acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
      acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
    }
    double* dst = partials + static_cast<size_t>(blockIdx.x) * 2 * p.C;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      dst[4 * cg + e] = acc[e];
      dst[p.C + 4 * cg + e] = acc[4 + e];
    }
  }
}

// CTA = 8 channels x 128 part groups (1024 threads): group g adds the per-CTA partials g, g+128, ... (a part row
// contributes 64 contiguous bytes = two fully used sectors per statistic), then the group sums are added in a fixed order in
// two levels, all in FP64 (deterministic).  C/8 CTAs: up to 9.7 MB of partials (C = 512) are read by 64 SMs, not by 16.
constexpr int kFinCols = 8;
__device__ __forceinline__ bool rows_bn_sum_partials(const double* __restrict__ partials, int nparts, int C, int& c, double& s, double& q) {
  __shared__ double red[2][128][kFinCols + 1];
  __shared__ double red2[2][8][kFinCols + 1];
  const int cl = threadIdx.x & 7, g = threadIdx.x >> 3;
  c = blockIdx.x * kFinCols + cl;
  double a = 0.0, b = 0.0;
  if (c < C) {
#pragma unroll 4
    for (int i = g; i < nparts; i += 128) {
      a += partials[static_cast<size_t>(i) * 2 * C + c];
      b += partials[static_cast<size_t>(i) * 2 * C + C + c];
    }
  }
  red[0][g][cl] = a;
  red[1][g][cl] = b;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int sub = threadIdx.x >> 3;
    double a2 = 0.0, b2 = 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) { a2 += red[0][sub * 16 + j][cl]; b2 += red[1][sub * 16 + j][cl]; }
    red2[0][sub][cl] = a2;
    red2[1][sub][cl] = b2;
  }
  __syncthreads();
  if (threadIdx.x >= kFinCols || c >= C) return false;
  s = 0.0; q = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { s += red2[0][j][cl]; q += red2[1][j][cl]; }
  return true;
}

__global__ void __launch_bounds__(1024) rows_bn_fwd_finalize_kernel(const double* __restrict__ partials, int nparts, int C, double count,
                                            const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                            float momentum, float* running_mean, float* running_var, float* __restrict__ mean,
                                            float* __restrict__ rstd, float* __restrict__ aff) {
  int c;
  double s, q;
  if (!rows_bn_sum_partials(partials, nparts, C, c, s, q)) return;
  const double m = s / count;
  double var = q / count - m * m;
  if (var < 0.0) var = 0.0;
  const double rs = 1.0 / sqrt(var + static_cast<double>(eps));
  mean[c] = static_cast<float>(m);
  rstd[c] = static_cast<float>(rs);
  aff[c] = static_cast<float>(m);
  aff[C + c] = static_cast<float>(static_cast<double>(gamma[c]) * rs);
  aff[2 * C + c] = beta[c];
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = static_cast<float>((1.0 - momentum) * running_mean[c] + momentum * m);
    running_var[c] = static_cast<float>((1.0 - momentum) * running_var[c] + momentum * unbiased);
  }
}

// inference: affine from the running statistics
__global__ void rows_bn_eval_affine_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                           const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                           float* __restrict__ aff) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  aff[c] = running_mean[c];
  aff[C + c] = gamma[c] * rsqrtf(running_var[c] + eps);
  aff[2 * C + c] = beta[c];
}

// coef = [kA | kB | kC] with dy = kA * dz - kB - kC * (v - mean):   kA = gamma rstd, kB = kA * s1 / M, kC = kA * rstd^2 * s2 / M
__global__ void __launch_bounds__(1024) rows_bn_bwd_finalize_kernel(const double* __restrict__ partials, int nparts, int C, double count,
                                            const float* __restrict__ gamma, const float* __restrict__ rstd,
                                            float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef) {
  int c;
  double s1, s2;
  if (!rows_bn_sum_partials(partials, nparts, C, c, s1, s2)) return;
  const double rs = rstd[c], kA = static_cast<double>(gamma[c]) * rs;
  dbeta[c] = static_cast<float>(s1);
  dgamma[c] = static_cast<float>(s2 * rs);                  // sum dz * vhat
  coef[c] = static_cast<float>(kA);
  coef[C + c] = static_cast<float>(kA * s1 / count);
  coef[2 * C + c] = static_cast<float>(kA * rs * rs * s2 / count);
}

// BWD = false: out = LReLU((v - mean) scale + beta);   BWD = true: out = kA dz - kB - kC (v - mean)
// NSP = 0: FP32 rows.  NSP = 2 / 3: the result goes straight into the tensor-core operand format (NSP bf16 planes in 64 x 64
// SWIZZLE_128B units, see gemm_tc.cu) for the GEMM that consumes it — the FP32 tensor and the separate split pass (4 B
// written + 4 B read + 4-6 B written per element) never exist.  Rows M .. Mpad-1 of the planes are written as zeros.
template <bool BWD, int NSP>
__global__ void __launch_bounds__(kRbThreads) rows_bn_apply_kernel(RowsBn p, const float* __restrict__ dh,
                                                                  const float* __restrict__ coef, float* __restrict__ out,
                                                                  unsigned char* __restrict__ planes, int Mpad, int seg_len,
                                                                  int rows_per_cta, double* __restrict__ colsum_partials) {
  __shared__ float4 red[kRbThreads];
  const int c4n = p.C >> 2, RL = kRbThreads / c4n;
  const int cg = threadIdx.x % c4n, rl = threadIdx.x / c4n;
  const float4 mean = ld4(p.a + 4 * cg), scale = ld4(p.a + p.C + 4 * cg), beta = ld4(p.a + 2 * p.C + 4 * cg);
  float4 kA = mean, kB = mean, kC = mean;
  if (BWD) {
    kA = ld4(coef + 4 * cg);
    kB = ld4(coef + p.C + 4 * cg);
    kC = ld4(coef + 2 * p.C + 4 * cg);
  }
  // NSP == 0: rows interleaved over the grid.  Planes: grid (parts, segments), CTA (q, seg) owns a contiguous slice of the
  // rows of segment seg (so that its column sums are per-segment partials); the zero rows M .. Mpad-1 go to CTA (0, 0).
  int r_begin, r_end, r_step;
  if (NSP == 0) {
    r_begin = blockIdx.x * RL + rl; r_end = p.M; r_step = gridDim.x * RL;
  } else {
    const int s0 = blockIdx.y * seg_len;
    r_begin = s0 + blockIdx.x * rows_per_cta + rl;
    r_end = min(s0 + seg_len, s0 + (static_cast<int>(blockIdx.x) + 1) * rows_per_cta);
    r_step = RL;
    if (blockIdx.x == 0 && blockIdx.y == 0 && rl < RL) {
      for (int r = p.M + rl; r < Mpad; r += RL) {
        const int c = 4 * cg;
        unsigned char* unit = planes + (static_cast<size_t>(r >> 6) * (p.C >> 6) + (c >> 6)) * NSP * tc::kUnitBytes +
                              tc::unit_elem_offset(r & 63, c & 63);
#pragma unroll
        for (int q = 0; q < NSP; ++q) *reinterpret_cast<uint2*>(unit + static_cast<size_t>(q) * tc::kUnitBytes) = make_uint2(0u, 0u);
      }
    }
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rl < RL) {
    for (int r = r_begin; r < r_end; r += r_step) {
      const size_t off = static_cast<size_t>(r) * p.C + 4 * cg;
      float4 v = ld4(p.y + off);
      if (p.bias) {
        const float4 b = ld4(p.bias + static_cast<size_t>(r / p.N) * p.C + 4 * cg);
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      }
      const float dx = v.x - mean.x, dy = v.y - mean.y, dz = v.z - mean.z, dw = v.w - mean.w;
      const float zx = __fmaf_rn(dx, scale.x, beta.x), zy = __fmaf_rn(dy, scale.y, beta.y);
      const float zz = __fmaf_rn(dz, scale.z, beta.z), zw = __fmaf_rn(dw, scale.w, beta.w);
      float4 o;
      if (!BWD) {
        o = make_float4(lrelu(zx), lrelu(zy), lrelu(zz), lrelu(zw));
      } else {
        const float4 g = ld4(dh + off);
        o.x = kA.x * (g.x * lrelu_grad(zx)) - kB.x - kC.x * dx;
        o.y = kA.y * (g.y * lrelu_grad(zy)) - kB.y - kC.y * dy;
        o.z = kA.z * (g.z * lrelu_grad(zz)) - kB.z - kC.z * dz;
        o.w = kA.w * (g.w * lrelu_grad(zw)) - kB.w - kC.w * dw;
      }
      if (NSP == 0) {
        *reinterpret_cast<float4*>(out + off) = o;
      } else {
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        const int c = 4 * cg;
        unsigned char* unit = planes + (static_cast<size_t>(r >> 6) * (p.C >> 6) + (c >> 6)) * NSP * tc::kUnitBytes +
                              tc::unit_elem_offset(r & 63, c & 63);
        float v4[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int q = 0; q < NSP; ++q) {
          const __nv_bfloat162 h0 = __floats2bfloat162_rn(v4[0], v4[1]), h1 = __floats2bfloat162_rn(v4[2], v4[3]);
          if (q + 1 < NSP) {
            const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
            v4[0] -= f0.x; v4[1] -= f0.y; v4[2] -= f1.x; v4[3] -= f1.y;       // exact in FP32
          }
          uint2 w;
          w.x = *reinterpret_cast<const uint32_t*>(&h0);
          w.y = *reinterpret_cast<const uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(unit + static_cast<size_t>(q) * tc::kUnitBytes) = w;
        }
      }
    }
  }
  if (NSP != 0 && colsum_partials) {        // per-segment column sums of the result (the per-block bias gradient), fixed order
    red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < c4n) {
      double a4[4] = {0, 0, 0, 0};
      for (int l = 0; l < RL; ++l) {
        const float4 a = red[l * c4n + cg];
        a4[0] += a.x; a4[1] += a.y; a4[2] += a.z; a4[3] += a.w;
      }
      double* dst = colsum_partials + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * p.C + 4 * cg;
#pragma unroll
      for (int e = 0; e < 4; ++e) dst[e] = a4[e];
    }
  }
}

// out[seg][c] = sum over the N rows of segment seg of x[row][c]  (per-block bias gradient of conv7, colsum(feat)).
// grid (parts, segments): CTA (p, seg) sums a contiguous slice of the segment's rows -> FP64 partial; a second pass adds
// the parts in order.  Same thread mapping as the BN reductions.
__global__ void __launch_bounds__(kRbThreads) rows_colsum_partial_kernel(const float* __restrict__ x, int C, int N, int rows_per_cta,
                                                                        double* __restrict__ partials) {
  __shared__ float4 red[kRbThreads];
  const int c4n = C >> 2, RL = kRbThreads / c4n;            // any C % 4 == 0, C <= 1024: threads >= c4n * RL idle
  const int cg = threadIdx.x % c4n, rl = threadIdx.x / c4n;
  const int seg = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(N, r0 + rows_per_cta);
  const float* base = x + static_cast<size_t>(seg) * N * C;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rl < RL) {
    // eight independent rows in flight per thread (a single dependent chain of loads made this pass latency-bound)
    int r = r0 + rl;
    for (; r + 7 * RL < r1; r += 8 * RL) {
      float4 v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = ld4(base + static_cast<size_t>(r + e * RL) * C + 4 * cg);
      s.x += ((v[0].x + v[1].x) + (v[2].x + v[3].x)) + ((v[4].x + v[5].x) + (v[6].x + v[7].x));
      s.y += ((v[0].y + v[1].y) + (v[2].y + v[3].y)) + ((v[4].y + v[5].y) + (v[6].y + v[7].y));
      s.z += ((v[0].z + v[1].z) + (v[2].z + v[3].z)) + ((v[4].z + v[5].z) + (v[6].z + v[7].z));
      s.w += ((v[0].w + v[1].w) + (v[2].w + v[3].w)) + ((v[4].w + v[5].w) + (v[6].w + v[7].w));
    }
    for (; r < r1; r += RL) {
      const float4 v = ld4(base + static_cast<size_t>(r) * C + 4 * cg);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < c4n) {
    double acc[4] = {0, 0, 0, 0};
    for (int l = 0; l < RL; ++l) {
      const float4 a = red[l * c4n + cg];
      acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
    }
    double* dst = partials + (static_cast<size_t>(seg) * gridDim.x + blockIdx.x) * C + 4 * cg;
#pragma unroll
    for (int e = 0; e < 4; ++e) dst[e] = acc[e];
  }
}
// 32 consecutive outputs x 8 part groups per CTA: group g adds parts g, g + 8, ... (coalesced 256-byte rows, loads
// independent), then the 8 group sums are added in order.
__global__ void __launch_bounds__(256) rows_colsum_finish_kernel(const double* __restrict__ partials, int nparts, int C, int nseg,
                                                                 float* __restrict__ out) {
  __shared__ double red[8][32];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  double s = 0.0;
  if (i < nseg * C) {
    const int seg = i / C, c = i - seg * C;
    const double* src = partials + static_cast<size_t>(seg) * nparts * C + c;
#pragma unroll 4
    for (int q = g; q < nparts; q += 8) s += src[static_cast<size_t>(q) * C];
  }
  red[g][lane] = s;
  __syncthreads();
  if (g == 0 && i < nseg * C) {
    double a = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) a += red[q][lane];
    out[i] = static_cast<float>(a);
  }
}

static bool rows_bn_shape_ok(int C) { return C == 64 || C == 128 || C == 256 || C == 512 || C == 1024; }

static int rows_bn_grid(int M, int* rows_per_cta) {
  const int sms = std::max(1, num_sms());
  int ctas = std::max(1, std::min(std::min(M, sms * 8), kRbMaxCtas));
  *rows_per_cta = (M + ctas - 1) / ctas;
  return (M + *rows_per_cta - 1) / *rows_per_cta;
}

static int rows_bn_apply_grid(int M, int C) {
  const int RL = kRbThreads / (C / 4);
  return std::max(1, std::min((M + RL - 1) / RL, std::max(1, num_sms()) * 8));
}

}  // namespace dg

using namespace dg;

extern "C" int dgcnn_rows_bn_supported(int C) { return rows_bn_shape_ok(C) ? 1 : 0; }

extern "C" size_t dgcnn_rows_bn_workspace_bytes(int C) {
  return (static_cast<size_t>(kRbMaxCtas) * 2 * C) * sizeof(double) + 3 * static_cast<size_t>(C) * sizeof(float);
}

extern "C" int dgcnn_rows_bn_lrelu_fwd(const float* y, int M, int C, const float* bias, int N, const float* gamma,
                                       const float* beta, float eps, float momentum, int use_batch_stats,
                                       float* running_mean, float* running_var, float* mean, float* rstd, float* aff,
                                       float* h, void* workspace, size_t workspace_bytes, void* stream) {
  DG_REQUIRE(y && gamma && beta && aff && h && workspace, DGCNN_E_INVALID, "rows_bn_lrelu_fwd: null pointer");
  DG_REQUIRE(rows_bn_shape_ok(C), DGCNN_E_UNSUPPORTED, "rows_bn_lrelu_fwd: C must be 64, 128, 256, 512 or 1024");
  DG_REQUIRE(M > 0 && N > 0 && (!bias || M % N == 0), DGCNN_E_INVALID, "rows_bn_lrelu_fwd: bad shape");
  DG_REQUIRE(workspace_bytes >= dgcnn_rows_bn_workspace_bytes(C), DGCNN_E_INVALID, "rows_bn_lrelu_fwd: workspace too small");
  DG_REQUIRE(use_batch_stats ? (mean && rstd && (!running_mean == !running_var)) : (running_mean && running_var),
             DGCNN_E_INVALID, "rows_bn_lrelu_fwd: statistics buffers missing");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RowsBn p{y, M, C, N, bias, aff};
  if (use_batch_stats) {
    double* partials = static_cast<double*>(workspace);
    int rpc;
    const int grid = rows_bn_grid(M, &rpc);
    rows_bn_reduce_kernel<false><<<grid, kRbThreads, 0, st>>>(p, nullptr, rpc, partials);
    rows_bn_fwd_finalize_kernel<<<(C + kFinCols - 1) / kFinCols, 1024, 0, st>>>(partials, grid, C, static_cast<double>(M), gamma, beta, eps,
                                                                 momentum, running_mean, running_var, mean, rstd, aff);
  } else {
    rows_bn_eval_affine_kernel<<<(C + 127) / 128, 128, 0, st>>>(C, gamma, beta, eps, running_mean, running_var, aff);
  }
  const int blocks = rows_bn_apply_grid(M, C);
  rows_bn_apply_kernel<false, 0><<<blocks, kRbThreads, 0, st>>>(p, nullptr, nullptr, h, nullptr, 0, 0, 0, nullptr);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_rows_bn_lrelu_bwd(const float* dh, const float* y, int M, int C, const float* bias, int N,
                                       const float* gamma, const float* rstd, const float* aff, float* dy, float* dgamma,
                                       float* dbeta, void* workspace, size_t workspace_bytes, void* stream) {
  DG_REQUIRE(dh && y && gamma && rstd && aff && dy && dgamma && dbeta && workspace, DGCNN_E_INVALID, "rows_bn_lrelu_bwd: null pointer");
  DG_REQUIRE(rows_bn_shape_ok(C), DGCNN_E_UNSUPPORTED, "rows_bn_lrelu_bwd: C must be 64, 128, 256, 512 or 1024");
  DG_REQUIRE(M > 0 && N > 0 && (!bias || M % N == 0), DGCNN_E_INVALID, "rows_bn_lrelu_bwd: bad shape");
  DG_REQUIRE(workspace_bytes >= dgcnn_rows_bn_workspace_bytes(C), DGCNN_E_INVALID, "rows_bn_lrelu_bwd: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RowsBn p{y, M, C, N, bias, aff};
  double* partials = static_cast<double*>(workspace);
  float* coef = reinterpret_cast<float*>(partials + static_cast<size_t>(kRbMaxCtas) * 2 * C);
  int rpc;
  const int grid = rows_bn_grid(M, &rpc);
  rows_bn_reduce_kernel<true><<<grid, kRbThreads, 0, st>>>(p, dh, rpc, partials);
  rows_bn_bwd_finalize_kernel<<<(C + kFinCols - 1) / kFinCols, 1024, 0, st>>>(partials, grid, C, static_cast<double>(M), gamma, rstd, dgamma, dbeta, coef);
  const int blocks = rows_bn_apply_grid(M, C);
  rows_bn_apply_kernel<true, 0><<<blocks, kRbThreads, 0, st>>>(p, dh, coef, dy, nullptr, 0, 0, 0, nullptr);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" size_t dgcnn_rows_colsum_workspace_bytes(int M, int C, int N);

// ---- the same two passes with the result written as GEMM operand planes (nsplit = 2 or 3; rows padded to a multiple of 128) ----
static int planes_args_ok(int M, int C, int nsplit, const void* planes, size_t planes_bytes, const char* who) {
  DG_REQUIRE(planes && (nsplit == 2 || nsplit == 3) && C % 64 == 0 && (reinterpret_cast<uintptr_t>(planes) & 15) == 0, DGCNN_E_INVALID,
             "%s: planes must be non-null and 16-byte aligned, nsplit 2 or 3, C a multiple of 64", who);
  const size_t need = static_cast<size_t>((M + 127) / 128 * 2) * (C / 64) * nsplit * tc::kUnitBytes;
  DG_REQUIRE(planes_bytes >= need, DGCNN_E_WORKSPACE, "%s: planes buffer too small (%zu < %zu)", who, planes_bytes, need);
  return DGCNN_OK;
}

// grid (parts, segments) of the planes-writing apply pass: as many CTAs as the interleaved form uses
static dim3 planes_apply_grid(int M, int seg_len, int* rows_per_cta) {
  const int nseg = M / seg_len, sms = std::max(1, num_sms());
  const int parts = std::max(1, std::min(kColsumMaxParts, std::min(seg_len, (sms * 8 + nseg - 1) / nseg)));
  *rows_per_cta = (seg_len + parts - 1) / parts;
  return dim3((seg_len + *rows_per_cta - 1) / *rows_per_cta, nseg, 1);
}

extern "C" int dgcnn_rows_bn_lrelu_fwd_planes(const float* y, int M, int C, const float* bias, int N, const float* gamma,
                                              const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                                              float* mean, float* rstd, float* aff, void* planes, size_t planes_bytes, int nsplit,
                                              void* workspace, size_t workspace_bytes, void* stream) {
  DG_REQUIRE(y && gamma && beta && aff && mean && rstd && workspace, DGCNN_E_INVALID, "rows_bn_lrelu_fwd_planes: null pointer");
  DG_REQUIRE(rows_bn_shape_ok(C), DGCNN_E_UNSUPPORTED, "rows_bn_lrelu_fwd_planes: C must be 64, 128, 256, 512 or 1024");
  DG_REQUIRE(M > 0 && N > 0 && (!bias || M % N == 0) && (!running_mean == !running_var), DGCNN_E_INVALID, "rows_bn_lrelu_fwd_planes: bad shape");
  DG_REQUIRE(workspace_bytes >= dgcnn_rows_bn_workspace_bytes(C), DGCNN_E_INVALID, "rows_bn_lrelu_fwd_planes: workspace too small");
  int rc = planes_args_ok(M, C, nsplit, planes, planes_bytes, "rows_bn_lrelu_fwd_planes");
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RowsBn p{y, M, C, N, bias, aff};
  double* partials = static_cast<double*>(workspace);
  int rpc;
  const int grid = rows_bn_grid(M, &rpc);
  rows_bn_reduce_kernel<false><<<grid, kRbThreads, 0, st>>>(p, nullptr, rpc, partials);
  rows_bn_fwd_finalize_kernel<<<(C + kFinCols - 1) / kFinCols, 1024, 0, st>>>(partials, grid, C, static_cast<double>(M), gamma, beta, eps, momentum,
                                                               running_mean, running_var, mean, rstd, aff);
  const int Mpad = (M + 127) / 128 * 128;
  int rpc2;
  const dim3 blocks = planes_apply_grid(M, M, &rpc2);
  unsigned char* pl = static_cast<unsigned char*>(planes);
  if (nsplit == 3)
    rows_bn_apply_kernel<false, 3><<<blocks, kRbThreads, 0, st>>>(p, nullptr, nullptr, nullptr, pl, Mpad, M, rpc2, nullptr);
  else
    rows_bn_apply_kernel<false, 2><<<blocks, kRbThreads, 0, st>>>(p, nullptr, nullptr, nullptr, pl, Mpad, M, rpc2, nullptr);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_rows_bn_lrelu_bwd_planes(const float* dh, const float* y, int M, int C, const float* bias, int N,
                                              const float* gamma, const float* rstd, const float* aff, void* planes,
                                              size_t planes_bytes, int nsplit, float* dgamma, float* dbeta, float* dbias,
                                              void* colsum_workspace, size_t colsum_workspace_bytes, void* workspace,
                                              size_t workspace_bytes, void* stream) {
  DG_REQUIRE(dh && y && gamma && rstd && aff && dgamma && dbeta && workspace, DGCNN_E_INVALID, "rows_bn_lrelu_bwd_planes: null pointer");
  DG_REQUIRE(rows_bn_shape_ok(C), DGCNN_E_UNSUPPORTED, "rows_bn_lrelu_bwd_planes: C must be 64, 128, 256, 512 or 1024");
  DG_REQUIRE(M > 0 && N > 0 && (!bias || M % N == 0), DGCNN_E_INVALID, "rows_bn_lrelu_bwd_planes: bad shape");
  DG_REQUIRE(workspace_bytes >= dgcnn_rows_bn_workspace_bytes(C), DGCNN_E_INVALID, "rows_bn_lrelu_bwd_planes: workspace too small");
  int rc = planes_args_ok(M, C, nsplit, planes, planes_bytes, "rows_bn_lrelu_bwd_planes");
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RowsBn p{y, M, C, N, bias, aff};
  double* partials = static_cast<double*>(workspace);
  float* coef = reinterpret_cast<float*>(partials + static_cast<size_t>(kRbMaxCtas) * 2 * C);
  int rpc;
  const int grid = rows_bn_grid(M, &rpc);
  rows_bn_reduce_kernel<true><<<grid, kRbThreads, 0, st>>>(p, dh, rpc, partials);
  rows_bn_bwd_finalize_kernel<<<(C + kFinCols - 1) / kFinCols, 1024, 0, st>>>(partials, grid, C, static_cast<double>(M), gamma, rstd, dgamma, dbeta, coef);
  const int Mpad = (M + 127) / 128 * 128;
  // dbias (the gradient of the per-block bias = per-block column sums of dy) comes out of the same pass
  DG_REQUIRE(!dbias || (bias && colsum_workspace && colsum_workspace_bytes >= dgcnn_rows_colsum_workspace_bytes(M, C, N)), DGCNN_E_INVALID,
             "rows_bn_lrelu_bwd_planes: dbias needs a bias and dgcnn_rows_colsum_workspace_bytes(M, C, N) of workspace");
  const int seg = dbias ? N : M;
  int rpc2;
  const dim3 blocks = planes_apply_grid(M, seg, &rpc2);
  double* cs = dbias ? static_cast<double*>(colsum_workspace) : nullptr;
  unsigned char* pl = static_cast<unsigned char*>(planes);
  if (nsplit == 3)
    rows_bn_apply_kernel<true, 3><<<blocks, kRbThreads, 0, st>>>(p, dh, coef, nullptr, pl, Mpad, seg, rpc2, cs);
  else
    rows_bn_apply_kernel<true, 2><<<blocks, kRbThreads, 0, st>>>(p, dh, coef, nullptr, pl, Mpad, seg, rpc2, cs);
  if (dbias) rows_colsum_finish_kernel<<<(M / N * C + 31) / 32, 256, 0, st>>>(cs, static_cast<int>(blocks.x), C, M / N, dbias);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" size_t dgcnn_rows_colsum_workspace_bytes(int M, int C, int N) {
  if (M <= 0 || C <= 0 || N <= 0) return 0;
  return static_cast<size_t>(M / N) * kColsumMaxParts * C * sizeof(double);
}

extern "C" int dgcnn_rows_colsum_f32(const float* x, int M, int C, int N, float* out, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  DG_REQUIRE(x && out && workspace, DGCNN_E_INVALID, "rows_colsum: null pointer");
  DG_REQUIRE(C % 4 == 0 && C <= 1024 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, DGCNN_E_UNSUPPORTED,
             "rows_colsum: C must be a multiple of 4, at most 1024, x 16-byte aligned");
  DG_REQUIRE(M > 0 && N > 0 && M % N == 0, DGCNN_E_INVALID, "rows_colsum: M must be a multiple of N");
  DG_REQUIRE(workspace_bytes >= dgcnn_rows_colsum_workspace_bytes(M, C, N), DGCNN_E_INVALID, "rows_colsum: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nseg = M / N;
  const int sms = std::max(1, num_sms());
  const int parts = std::max(1, std::min(kColsumMaxParts, std::min(N, (sms * 8 + nseg - 1) / nseg)));
  const int rpc = (N + parts - 1) / parts;
  const int nparts = (N + rpc - 1) / rpc;
  rows_colsum_partial_kernel<<<dim3(nparts, nseg), kRbThreads, 0, st>>>(x, C, N, rpc, static_cast<double*>(workspace));
  rows_colsum_finish_kernel<<<(nseg * C + 31) / 32, 256, 0, st>>>(static_cast<const double*>(workspace), nparts, C, nseg, out);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}
