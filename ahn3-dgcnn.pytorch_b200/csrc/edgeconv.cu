// edgeconv.cu — fused EdgeConv blocks (replaces model.py:43-83 + 116-130 + 147-163).
//
// FP32-FMA implementation ("v1"): one persistent CTA processes tiles of P = min(128/k, 16) points,
// i.e. P*k <= 128 edge rows.  The gathered edge features [x_j - x_i] live only in shared memory; the
// x_i half of the 1x1 conv is hoisted to one mat-vec per point (W [x_j-x_i ; x_i] = Wa1 (x_j-x_i) + Wa2 x_i).
// Inference is a single kernel (gather -> conv_a -> BN -> LReLU -> conv_b -> BN -> LReLU -> max_k).
// Training needs batch statistics before normalising, so the block is split at the BN boundaries; the
// per-edge pre-BN conv_a output y_a is the only per-edge tensor that touches HBM (kept for backward).
#include "edgeconv.cuh"

namespace dg {

// Per-CTA column statistics -> partials (double).  t1/t2 are per-thread partial sums for columns tx+16q.
__device__ __forceinline__ void flush_stats(const float (&t1)[4], const float (&t2)[4], float* scratch,
                                            double* partials, int tid, int tx, int ty) {
  __syncthreads();
  float* r1 = scratch;               // [16][64]
  float* r2 = scratch + 16 * 64;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    r1[ty * 64 + tx + 16 * q] = t1[q];
    r2[ty * 64 + tx + 16 * q] = t2[q];
  }
  __syncthreads();
  if (tid < 128) {
    const float* r = tid < 64 ? r1 : r2;
    const int c = tid & 63;
    double s = 0.0;
#pragma unroll
    for (int y = 0; y < 16; ++y) s += static_cast<double>(r[y * 64 + c]);
    partials[128 + static_cast<size_t>(blockIdx.x) * 128 + tid] = s;
  }
  if (blockIdx.x == 0 && tid == 0) partials[0] = static_cast<double>(gridDim.x);
}

template <int CP, int MODE>
struct EcSmem {
  static constexpr int LDA = CP + 4;
  static constexpr size_t h = static_cast<size_t>(kTileRows) * kLdH;
  static constexpr size_t wa = (MODE == MODE_TRAIN_B) ? 0 : static_cast<size_t>(kCout) * LDA;
  static constexpr size_t wb = (MODE == MODE_TRAIN_A) ? 0 : static_cast<size_t>(kCout) * kLdH;
  static constexpr size_t v = kMaxP * kCout;
  static constexpr size_t xi = (MODE == MODE_TRAIN_B) ? 0 : static_cast<size_t>(kMaxP) * CP;
  static constexpr size_t aff = 2 * 192;
  static constexpr size_t floats = h + 2 * wa + wb + v + xi + aff;
  static constexpr size_t bytes = floats * 4 + kTileRows * 4;
};

template <int CP, int MODE>
__global__ void __launch_bounds__(kThreads, 2) ec_fwd_kernel(const EcK p) {
  using SM = EcSmem<CP, MODE>;
  constexpr int LDA = SM::LDA;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Hs = reinterpret_cast<float*>(smem_raw);     // A tile (stride LDA) / H tile / Y tile (stride kLdH)
  float* Wa1s = Hs + SM::h;
  float* Wa2s = Wa1s + SM::wa;
  float* Wbs = Wa2s + SM::wa;
  float* Vs = Wbs + SM::wb;
  float* xis = Vs + SM::v;
  float* affA = xis + SM::xi;
  float* affB = affA + 192;
  int* nbr = reinterpret_cast<int*>(affB + 192);

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int k = p.k, P = p.P, M = p.M;

  // ---- one-time per CTA: weights and BN affine into smem -----------------------------------------
  if (MODE != MODE_TRAIN_B) {
    for (int i = tid; i < kCout * CP; i += kThreads) {
      const int o = i / CP, c = i % CP;
      Wa1s[o * LDA + c] = c < p.C ? p.w_a[static_cast<size_t>(o) * 2 * p.C + c] : 0.f;
      Wa2s[o * LDA + c] = c < p.C ? p.w_a[static_cast<size_t>(o) * 2 * p.C + p.C + c] : 0.f;
    }
  }
  if (MODE == MODE_TRAIN_B || (MODE == MODE_EVAL && p.two)) {
    for (int i = tid; i < kCout * kCout; i += kThreads) Wbs[(i >> 6) * kLdH + (i & 63)] = p.w_b[i];
  }
  if (tid < 192) {
    affA[tid] = p.aff_a ? p.aff_a[tid] : 0.f;
    affB[tid] = p.aff_b ? p.aff_b[tid] : 0.f;
  }

  int rp[8];  // point slot of each of this thread's rows (row = ty + 16 i)
#pragma unroll
  for (int i = 0; i < 8; ++i) rp[i] = (ty + 16 * i) / k;

  float t1[4] = {0.f, 0.f, 0.f, 0.f}, t2[4] = {0.f, 0.f, 0.f, 0.f};

  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const int gp0 = tile * P;
    __syncthreads();  // S0: previous tile fully consumed; one-time loads visible
    if (tid < kTileRows) {
      const int pp = tid / k, slot = tid - pp * k, gp = gp0 + pp;
      int j = -1;
      if (pp < P && gp < M) j = (gp / p.N) * p.N + p.idx[static_cast<size_t>(gp) * k + slot];
      nbr[tid] = j;
    }
    if (MODE != MODE_TRAIN_B) {
      for (int i = tid; i < P * (CP / 4); i += kThreads) {
        const int pp = i / (CP / 4), c4 = i % (CP / 4), gp = gp0 + pp;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gp < M && 4 * c4 < p.C4) v = *reinterpret_cast<const float4*>(p.rows + static_cast<size_t>(gp) * p.ld + 4 * c4);
        *reinterpret_cast<float4*>(xis + pp * CP + 4 * c4) = v;
      }
    }
    __syncthreads();  // S1
    if (MODE != MODE_TRAIN_B) {
      // gather: A[row] = x_j - x_i   (model.py:74-81, first half of the concat)
      for (int i = tid; i < kTileRows * (CP / 4); i += kThreads) {
        const int row = i / (CP / 4), c4 = i % (CP / 4), j = nbr[row];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j >= 0 && 4 * c4 < p.C4) {
          const float4 xj = *reinterpret_cast<const float4*>(p.rows + static_cast<size_t>(j) * p.ld + 4 * c4);
          const float4 xi = *reinterpret_cast<const float4*>(xis + (row / k) * CP + 4 * c4);
          v = make_float4(xj.x - xi.x, xj.y - xi.y, xj.z - xi.z, xj.w - xi.w);
        }
        *reinterpret_cast<float4*>(Hs + row * LDA + 4 * c4) = v;
      }
      // hoisted centre term V[p] = Wa2 x_i (second half of the concat, identical for the k edges of p)
      for (int i = tid; i < P * kCout; i += kThreads) {
        const int pp = i >> 6, o = i & 63;
        float a = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < CP / 4; ++c4) {
          const float4 w = *reinterpret_cast<const float4*>(Wa2s + o * LDA + 4 * c4);
          const float4 x = *reinterpret_cast<const float4*>(xis + pp * CP + 4 * c4);
          a = __fmaf_rn(w.x, x.x, a);
          a = __fmaf_rn(w.y, x.y, a);
          a = __fmaf_rn(w.z, x.z, a);
          a = __fmaf_rn(w.w, x.w, a);
        }
        Vs[i] = a;
      }
    } else {
      // conv_b input: h = LReLU(BN_a(y_a)) from the stored pre-BN tile
      const size_t e0 = static_cast<size_t>(gp0) * k;
      for (int i = tid; i < kTileRows * 16; i += kThreads) {
        const int row = i >> 4, c4 = i & 15;
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nbr[row] >= 0) {
          const float4 y = *reinterpret_cast<const float4*>(p.y_a_in + (e0 + row) * kCout + 4 * c4);
          const int c = 4 * c4;
          h.x = lrelu(bn_apply(y.x, affA[c + 0], affA[64 + c + 0], affA[128 + c + 0]));
          h.y = lrelu(bn_apply(y.y, affA[c + 1], affA[64 + c + 1], affA[128 + c + 1]));
          h.z = lrelu(bn_apply(y.z, affA[c + 2], affA[64 + c + 2], affA[128 + c + 2]));
          h.w = lrelu(bn_apply(y.w, affA[c + 3], affA[64 + c + 3], affA[128 + c + 3]));
        }
        *reinterpret_cast<float4*>(Hs + row * kLdH + 4 * c4) = h;
      }
    }
    __syncthreads();  // S2

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;

    if (MODE != MODE_TRAIN_B) {
      gemm_rr<CP, LDA, LDA>(Hs, Wa1s, acc, ty, tx);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int pp = rp[i] < P ? rp[i] : 0;
          acc[i][q] += Vs[pp * kCout + tx + 16 * q];
        }
      if (MODE == MODE_EVAL) {
        if (p.two) {
          __syncthreads();  // S3: all reads of the A tile are done before H overwrites it
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int row = ty + 16 * i, col = tx + 16 * q;
              const float z = bn_apply(acc[i][q], affA[col], affA[64 + col], affA[128 + col]);
              Hs[row * kLdH + col] = nbr[row] >= 0 ? lrelu(z) : 0.f;
              acc[i][q] = 0.f;
            }
          __syncthreads();  // S4
          gemm_rr<kCout, kLdH, kLdH>(Hs, Wbs, acc, ty, tx);
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int col = tx + 16 * q;
              acc[i][q] = lrelu(bn_apply(acc[i][q], affB[col], affB[64 + col], affB[128 + col]));
            }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int col = tx + 16 * q;
              acc[i][q] = lrelu(bn_apply(acc[i][q], affA[col], affA[64 + col], affA[128 + col]));
            }
        }
      }
    } else {
      gemm_rr<kCout, kLdH, kLdH>(Hs, Wbs, acc, ty, tx);
    }

    if (MODE != MODE_EVAL) {
      // batch statistics of the raw conv output (valid edge rows only)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool valid = nbr[ty + 16 * i] >= 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float y = valid ? acc[i][q] : 0.f;
          t1[q] += y;
          t2[q] = __fmaf_rn(y, y, t2[q]);
        }
      }
    }
    __syncthreads();  // S5: GEMM reads of Hs done
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) Hs[(ty + 16 * i) * kLdH + tx + 16 * q] = acc[i][q];
    __syncthreads();  // S6: Y tile complete

    if (MODE == MODE_EVAL) {
      for (int i = tid; i < P * kCout; i += kThreads) {
        const int pp = i >> 6, c = i & 63, gp = gp0 + pp;
        if (gp < M) {
          float m = Hs[(pp * k) * kLdH + c];
          for (int s = 1; s < k; ++s) m = fmaxf(m, Hs[(pp * k + s) * kLdH + c]);
          p.out[static_cast<size_t>(gp) * p.ld_out + c] = m;
        }
      }
    } else {
      if (MODE == MODE_TRAIN_A) {
        const size_t e0 = static_cast<size_t>(gp0) * k;
        const int nvalid = min(P, M - gp0) * k;
        for (int i = tid; i < nvalid * 16; i += kThreads) {
          const int row = i >> 4, c4 = i & 15;
          *reinterpret_cast<float4*>(p.y_a_out + (e0 + row) * kCout + 4 * c4) =
              *reinterpret_cast<const float4*>(Hs + row * kLdH + 4 * c4);
        }
      }
      if (MODE == MODE_TRAIN_B || !p.two) {
        for (int i = tid; i < P * kCout; i += kThreads) {
          const int pp = i >> 6, c = i & 63, gp = gp0 + pp;
          if (gp < M) {
            float mx = Hs[(pp * k) * kLdH + c], mn = mx;
            int ax = 0, an = 0;
            for (int s = 1; s < k; ++s) {
              const float v = Hs[(pp * k + s) * kLdH + c];
              if (v > mx) { mx = v; ax = s; }
              if (v < mn) { mn = v; an = s; }
            }
            const size_t o = static_cast<size_t>(gp) * kCout + c;
            p.ymax[o] = mx; p.ymin[o] = mn;
            p.amax[o] = static_cast<uint8_t>(ax); p.amin[o] = static_cast<uint8_t>(an);
          }
        }
      }
    }
  }
  if (MODE != MODE_EVAL) flush_stats(t1, t2, Hs, p.partials, tid, tx, ty);
}

// ---------------------------------------------------------------------------------------------
// small per-channel kernels
// ---------------------------------------------------------------------------------------------
// Sum partials[part][2][C] over the parts (fixed order -> deterministic).  The kernel is pure latency (up to 1024 parts of
// 1 KB), so it is spread over kFinCtas CTAs of 4 channels: thread = (channel, statistic, part lane), 32 part lanes with
// all their loads in flight; a part contributes two fully used 32-byte sectors per CTA.  Result valid for threadIdx.x < 4
// (channel blockIdx.x * 4 + threadIdx.x).  blockDim.x = 256, C = 64.
constexpr int kFinCtas = 16, kFinThreads = 256;
__device__ __forceinline__ void reduce_partials(const double* __restrict__ partials, int C, double& s, double& q) {
  __shared__ double red[2][4][33];
  const int cl = threadIdx.x & 3, stat = (threadIdx.x >> 2) & 1, pl = threadIdx.x >> 3;
  const int nparts = static_cast<int>(partials[0]);
  const double* src = partials + 128 + stat * C + blockIdx.x * 4 + cl;
  double a = 0.0;
#pragma unroll 8
  for (int i = pl; i < nparts; i += 32) a += src[static_cast<size_t>(i) * 2 * C];
  red[stat][cl][pl] = a;
  __syncthreads();
  s = 0.0; q = 0.0;
  if (threadIdx.x < 4) {
#pragma unroll
    for (int g2 = 0; g2 < 32; ++g2) {
      s += red[0][threadIdx.x][g2];
      q += red[1][threadIdx.x][g2];
    }
  }
}

__global__ void __launch_bounds__(kFinThreads) bn_finalize_kernel(const double* __restrict__ partials, int C, double count,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float eps, float momentum, float* running_mean,
                                                          float* running_var, float* mean, float* rstd, float* aff) {
  double s, q;
  reduce_partials(partials, C, s, q);
  if (threadIdx.x >= 4) return;
  const int c = blockIdx.x * 4 + threadIdx.x;
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

__global__ void select_kernel(int M, const float* __restrict__ ymax, const float* __restrict__ ymin,
                              const uint8_t* __restrict__ amax, const uint8_t* __restrict__ amin,
                              const float* __restrict__ aff, float* __restrict__ out, int ld_out,
                              float* __restrict__ ysel, uint8_t* __restrict__ arg) {
  const size_t total = static_cast<size_t>(M) * kCout;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i & 63);
    const size_t pnt = i >> 6;
    const float sc = aff[64 + c];
    const bool up = sc >= 0.f;      // BN affine is increasing iff scale >= 0; LeakyReLU is increasing
    const float y = up ? ymax[i] : ymin[i];
    ysel[i] = y;
    arg[i] = up ? amax[i] : amin[i];
    out[pnt * ld_out + c] = lrelu(bn_apply(y, aff[c], sc, aff[128 + c]));
  }
}

// g = dout * LReLU'(z(ysel)); per-channel sums of g and g*ysel.
__global__ void __launch_bounds__(kThreads) bwd_prep_kernel(int M, const float* __restrict__ dout, int ld_dout,
                                                            const float* __restrict__ ysel,
                                                            const float* __restrict__ aff, float* __restrict__ g,
                                                            double* __restrict__ partials) {
  __shared__ double red[2][4][64];
  const int c = threadIdx.x & 63, r = threadIdx.x >> 6;
  const float mean = aff[c], sc = aff[64 + c], be = aff[128 + c];
  double s1 = 0.0, s2 = 0.0;
  for (int pnt = blockIdx.x * 4 + r; pnt < M; pnt += gridDim.x * 4) {
    const size_t o = static_cast<size_t>(pnt) * kCout + c;
    const float y = ysel[o];
    const float gv = dout[static_cast<size_t>(pnt) * ld_dout + c] * lrelu_grad(bn_apply(y, mean, sc, be));
    g[o] = gv;
    s1 += gv;
    s2 += static_cast<double>(gv) * y;
  }
  red[0][r][c] = s1;
  red[1][r][c] = s2;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int w = threadIdx.x >> 6;
    partials[128 + static_cast<size_t>(blockIdx.x) * 128 + threadIdx.x] =
        red[w][0][c] + red[w][1][c] + red[w][2][c] + red[w][3][c];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) partials[0] = static_cast<double>(gridDim.x);
}

// From S1 = sum dz, S2 = sum dz*y:  dgamma = rstd (S2 - mean S1), dbeta = S1 and the coefficients of
// dy = cA dz - cB - cC y  (BatchNorm backward with batch statistics).
__global__ void __launch_bounds__(kFinThreads) bn_bwd_reduce_kernel(const double* __restrict__ partials, int C, double count,
                                                            const float* __restrict__ gamma, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, float* dgamma, float* dbeta,
                                                            float* coef) {
  double s1, s2;
  reduce_partials(partials, C, s1, s2);
  if (threadIdx.x >= 4) return;
  const int c = blockIdx.x * 4 + threadIdx.x;
  const double m = mean[c], rs = rstd[c], ga = gamma[c];
  const double dg = rs * (s2 - m * s1);
  dgamma[c] = static_cast<float>(dg);
  dbeta[c] = static_cast<float>(s1);
  const double a = ga * rs;
  const double cC = a * (dg / count) * rs;
  coef[c] = static_cast<float>(a);
  coef[C + c] = static_cast<float>(a * (s1 / count) - cC * m);
  coef[2 * C + c] = static_cast<float>(cC);
}

// ---------------------------------------------------------------------------------------------
// backward edge pass through conv_b
// ---------------------------------------------------------------------------------------------
struct EcBwdSmem {
  static constexpr size_t tile = static_cast<size_t>(kTileRows) * kLdH;
  static constexpr size_t w = static_cast<size_t>(kCout) * kLdH;
  static constexpr size_t floats = 2 * tile + 2 * w + kMaxP * kCout + 192 + 192;
  static constexpr size_t bytes = floats * 4 + kMaxP * kCout;
};

__global__ void __launch_bounds__(kThreads, 2)
ec_bwd_b_kernel(const EcK p, const float* __restrict__ g, const uint8_t* __restrict__ arg,
                const float* __restrict__ coef_b, float* __restrict__ dz_a, float* __restrict__ dw_partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Hs = reinterpret_cast<float*>(smem_raw);
  float* DYs = Hs + EcBwdSmem::tile;
  float* Wbs = DYs + EcBwdSmem::tile;      // [c][c'']  (as stored)
  float* WbTs = Wbs + EcBwdSmem::w;        // [c''][c]
  float* gs = WbTs + EcBwdSmem::w;         // [P][64]
  float* affA = gs + kMaxP * kCout;        // mean | scale | beta of BN_a
  float* cb = affA + 192;                  // cA | cB | cC of BN_b
  uint8_t* args = reinterpret_cast<uint8_t*>(cb + 192);

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int k = p.k, P = p.P, M = p.M;

  for (int i = tid; i < kCout * kCout; i += kThreads) {
    const int o = i >> 6, c = i & 63;
    const float w = p.w_b[i];
    Wbs[o * kLdH + c] = w;
    WbTs[c * kLdH + o] = w;
  }
  if (tid < 192) {
    affA[tid] = p.aff_a[tid];
    cb[tid] = coef_b[tid];
  }
  int rp[8], rs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    rp[i] = (ty + 16 * i) / k;
    rs[i] = (ty + 16 * i) - rp[i] * k;
  }
  float adw[4][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n) adw[m][n] = 0.f;
  float t1[4] = {0.f, 0.f, 0.f, 0.f}, t2[4] = {0.f, 0.f, 0.f, 0.f};

  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const int gp0 = tile * P;
    const size_t e0 = static_cast<size_t>(gp0) * k;
    const int npts = min(P, M - gp0);
    const int nvalid = npts * k;
    __syncthreads();  // S0
    for (int i = tid; i < kTileRows * 16; i += kThreads) {
      const int row = i >> 4, c4 = i & 15;
      float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < nvalid) {
        const float4 y = *reinterpret_cast<const float4*>(p.y_a_in + (e0 + row) * kCout + 4 * c4);
        const int c = 4 * c4;
        h.x = lrelu(bn_apply(y.x, affA[c + 0], affA[64 + c + 0], affA[128 + c + 0]));
        h.y = lrelu(bn_apply(y.y, affA[c + 1], affA[64 + c + 1], affA[128 + c + 1]));
        h.z = lrelu(bn_apply(y.z, affA[c + 2], affA[64 + c + 2], affA[128 + c + 2]));
        h.w = lrelu(bn_apply(y.w, affA[c + 3], affA[64 + c + 3], affA[128 + c + 3]));
      }
      *reinterpret_cast<float4*>(Hs + row * kLdH + 4 * c4) = h;
    }
    for (int i = tid; i < P * kCout; i += kThreads) {
      const int pp = i >> 6, gp = gp0 + pp;
      const bool ok = gp < M;
      gs[i] = ok ? g[static_cast<size_t>(gp) * kCout + (i & 63)] : 0.f;
      args[i] = ok ? arg[static_cast<size_t>(gp) * kCout + (i & 63)] : 255;
    }
    __syncthreads();  // S1

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
    gemm_rr<kCout, kLdH, kLdH>(Hs, Wbs, acc, ty, tx);  // y_b (recomputed)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = ty + 16 * i;
      const bool valid = row < nvalid;
      const int pp = valid ? rp[i] : 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = tx + 16 * q;
        const float dz = (args[pp * kCout + col] == rs[i]) ? gs[pp * kCout + col] : 0.f;
        const float dy = cb[col] * dz - cb[64 + col] - cb[128 + col] * acc[i][q];
        DYs[row * kLdH + col] = valid ? dy : 0.f;
        acc[i][q] = 0.f;
      }
    }
    __syncthreads();  // S2
    gemm_kk<kLdH, kLdH>(DYs, Hs, nvalid, adw, ty, tx);        // dW_b[c][c''] += sum_rows dy[row][c] h[row][c'']
    gemm_rr<kCout, kLdH, kLdH>(DYs, WbTs, acc, ty, tx);       // dh[row][c''] = sum_c dy[row][c] W_b[c][c'']
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = ty + 16 * i;
      const bool valid = row < nvalid;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = tx + 16 * q;
        float dz = 0.f;
        if (valid) {
          dz = acc[i][q] * (Hs[row * kLdH + col] > 0.f ? 1.f : kSlope);
          const float ya = p.y_a_in[(e0 + row) * kCout + col];
          t1[q] += dz;
          t2[q] = __fmaf_rn(dz, ya, t2[q]);
        }
        acc[i][q] = dz;
      }
    }
    __syncthreads();  // S3: DYs reads done
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) DYs[(ty + 16 * i) * kLdH + tx + 16 * q] = acc[i][q];
    __syncthreads();  // S4
    for (int i = tid; i < nvalid * 16; i += kThreads) {
      const int row = i >> 4, c4 = i & 15;
      *reinterpret_cast<float4*>(dz_a + (e0 + row) * kCout + 4 * c4) =
          *reinterpret_cast<const float4*>(DYs + row * kLdH + 4 * c4);
    }
  }
  float* dst = dw_partials + static_cast<size_t>(blockIdx.x) * kCout * kCout;
#pragma unroll
  for (int m = 0; m < 4; ++m)
    *reinterpret_cast<float4*>(dst + (4 * ty + m) * kCout + 4 * tx) = make_float4(adw[m][0], adw[m][1], adw[m][2], adw[m][3]);
  flush_stats(t1, t2, Hs, p.partials, tid, tx, ty);
}

__global__ void dw_reduce_kernel(const float* __restrict__ dw_partials, const double* __restrict__ partials,
                                 float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kCout * kCout) return;
  const int nparts = static_cast<int>(partials[0]);
  double s = 0.0;
  for (int c = 0; c < nparts; ++c) s += dw_partials[static_cast<size_t>(c) * kCout * kCout + i];
  dw[i] = static_cast<float>(s);
}

// ---------------------------------------------------------------------------------------------
// backward edge pass through BN_a: dy_a, per-point sum s, scatter r
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
ec_bwd_a_kernel(const EcK p, const float* __restrict__ dz_a, const float* __restrict__ g,
                const uint8_t* __restrict__ arg, const float* __restrict__ coef_a, float* __restrict__ s_out,
                float* __restrict__ r_out, int ld_rs) {
  const int c4 = threadIdx.x & 15, pl = threadIdx.x >> 4;
  const int c = 4 * c4;
  const float4 cA = *reinterpret_cast<const float4*>(coef_a + c);
  const float4 cB = *reinterpret_cast<const float4*>(coef_a + 64 + c);
  const float4 cC = *reinterpret_cast<const float4*>(coef_a + 128 + c);
  const int k = p.k;
  for (int gp = blockIdx.x * 16 + pl; gp < p.M; gp += gridDim.x * 16) {
    const size_t base = static_cast<size_t>(gp / p.N) * p.N;
    float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
    uchar4 av = make_uchar4(255, 255, 255, 255);
    if (!dz_a) {
      gv = *reinterpret_cast<const float4*>(g + static_cast<size_t>(gp) * kCout + c);
      av = *reinterpret_cast<const uchar4*>(arg + static_cast<size_t>(gp) * kCout + c);
    }
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int slot = 0; slot < k; ++slot) {
      const size_t e = static_cast<size_t>(gp) * k + slot;
      const float4 y = *reinterpret_cast<const float4*>(p.y_a_in + e * kCout + c);
      float4 dz;
      if (dz_a) {
        dz = *reinterpret_cast<const float4*>(dz_a + e * kCout + c);
      } else {
        dz.x = av.x == slot ? gv.x : 0.f;
        dz.y = av.y == slot ? gv.y : 0.f;
        dz.z = av.z == slot ? gv.z : 0.f;
        dz.w = av.w == slot ? gv.w : 0.f;
      }
      float4 dy;
      dy.x = cA.x * dz.x - cB.x - cC.x * y.x;
      dy.y = cA.y * dz.y - cB.y - cC.y * y.y;
      dy.z = cA.z * dz.z - cB.z - cC.z * y.z;
      dy.w = cA.w * dz.w - cB.w - cC.w * y.w;
      s.x += dy.x; s.y += dy.y; s.z += dy.z; s.w += dy.w;
      const size_t j = base + p.idx[e];
      atomicAdd(reinterpret_cast<float4*>(r_out + j * ld_rs + c), dy);
    }
    *reinterpret_cast<float4*>(s_out + static_cast<size_t>(gp) * ld_rs + c) = s;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int fill_common(const dgcnn_edgeconv_args* a, EcK& p, const char* who) {
  DG_REQUIRE(a, DGCNN_E_INVALID, "%s: null args", who);
  DG_REQUIRE(a->B > 0 && a->N > 0 && a->C > 0 && a->k > 0, DGCNN_E_INVALID, "%s: bad shape", who);
  DG_REQUIRE(a->C <= 64 && a->k <= kTileRows, DGCNN_E_UNSUPPORTED, "%s: C=%d (max 64) / k=%d (max 128) not compiled",
             who, a->C, a->k);
  DG_REQUIRE(static_cast<long long>(a->B) * a->N * a->k < (1LL << 31), DGCNN_E_UNSUPPORTED, "%s: too many edges", who);
  p = EcK{};
  p.rows = a->rows; p.ld = a->ld; p.idx = a->idx;
  p.N = a->N; p.k = a->k; p.M = a->B * a->N; p.C = a->C; p.C4 = (a->C + 3) / 4 * 4;
  p.P = std::min(kTileRows / a->k, kMaxP);
  p.num_tiles = (p.M + p.P - 1) / p.P;
  p.w_a = a->w_a; p.w_b = a->w_b; p.two = a->w_b != nullptr;
  p.out = a->out; p.ld_out = a->ld_out;
  p.v_center = a->v_scratch;
  DG_REQUIRE(a->impl >= 0 && a->impl <= 2, DGCNN_E_INVALID, "%s: impl must be 0 (default), 1 (FP32 FMA) or 2 (serial tcgen05)", who);
  p.tc = a->impl == 0 ? 2 : (a->impl == 1 ? 0 : 1);
  return DGCNN_OK;
}

static int check_rows(const dgcnn_edgeconv_args* a, const char* who) {
  DG_REQUIRE(a->rows && a->idx && a->w_a, DGCNN_E_INVALID, "%s: null pointer", who);
  DG_REQUIRE(a->ld % 4 == 0 && a->ld >= (a->C + 3) / 4 * 4 && (reinterpret_cast<uintptr_t>(a->rows) & 15) == 0,
             DGCNN_E_INVALID, "%s: rows must be 16-byte aligned, ld %% 4 == 0, ld >= round4(C)", who);
  return DGCNN_OK;
}

static int persistent_grid(int num_tiles) {
  const int sms = num_sms();
  return std::max(1, std::min(std::min(num_tiles, 2 * (sms > 0 ? sms : 148)), kMaxPartialCtas));
}

template <int CP, int MODE>
static int launch_fwd(const EcK& p, cudaStream_t st) {
  using SM = EcSmem<CP, MODE>;
  auto kern = ec_fwd_kernel<CP, MODE>;
  DG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(SM::bytes)));
  kern<<<persistent_grid(p.num_tiles), kThreads, SM::bytes, st>>>(p);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

template <int MODE>
static int dispatch_cp(const EcK& p, cudaStream_t st) {
  if (p.C <= 4) return launch_fwd<4, MODE>(p, st);
  if (p.C <= 8) return launch_fwd<8, MODE>(p, st);
  if (p.C <= 16) return launch_fwd<16, MODE>(p, st);
  if (p.C <= 32) return launch_fwd<32, MODE>(p, st);
  return launch_fwd<64, MODE>(p, st);
}

}  // namespace dg

using namespace dg;

extern "C" size_t dgcnn_edgeconv_partials_bytes(void) { return (128 + static_cast<size_t>(kMaxPartialCtas) * 128) * sizeof(double); }
// per-CTA partials [128][64] (tensor-core path: D'^T H over H^T H; the FMA path uses the first [64][64]) + one summed copy
extern "C" size_t dgcnn_edgeconv_dw_partials_bytes(void) {
  return (static_cast<size_t>(kMaxPartialCtas) + 1) * 2 * kCout * kCout * sizeof(float);
}

extern "C" int dgcnn_edgeconv_fwd_eval(const dgcnn_edgeconv_args* a, const float* aff_a, const float* aff_b,
                                       void* stream) {
  EcK p;
  int rc = fill_common(a, p, "edgeconv_fwd_eval");
  if (rc) return rc;
  if ((rc = check_rows(a, "edgeconv_fwd_eval"))) return rc;
  DG_REQUIRE(a->out && aff_a && (!p.two || aff_b), DGCNN_E_INVALID, "edgeconv_fwd_eval: null pointer");
  p.aff_a = aff_a; p.aff_b = aff_b;
  if (p.C == 64 && p.v_center && p.tc)
    return launch_ec_tc_fwd(p, MODE_EVAL, persistent_grid(p.num_tiles), static_cast<cudaStream_t>(stream));
  return dispatch_cp<MODE_EVAL>(p, static_cast<cudaStream_t>(stream));
}

extern "C" int dgcnn_edgeconv_conv_a_train(const dgcnn_edgeconv_args* a, float* y_a, double* partials, float* ymax,
                                           float* ymin, uint8_t* amax, uint8_t* amin, void* stream) {
  EcK p;
  int rc = fill_common(a, p, "edgeconv_conv_a_train");
  if (rc) return rc;
  if ((rc = check_rows(a, "edgeconv_conv_a_train"))) return rc;
  DG_REQUIRE(y_a && partials, DGCNN_E_INVALID, "edgeconv_conv_a_train: null pointer");
  DG_REQUIRE(p.two || (ymax && ymin && amax && amin), DGCNN_E_INVALID,
             "edgeconv_conv_a_train: single-conv block needs the extrema outputs");
  p.y_a_out = y_a; p.partials = partials;
  p.ymax = ymax; p.ymin = ymin; p.amax = amax; p.amin = amin;
  if (p.C == 64 && p.v_center && p.tc)
    return launch_ec_tc_fwd(p, MODE_TRAIN_A, persistent_grid(p.num_tiles), static_cast<cudaStream_t>(stream));
  return dispatch_cp<MODE_TRAIN_A>(p, static_cast<cudaStream_t>(stream));
}

extern "C" int dgcnn_edgeconv_conv_b_train(const dgcnn_edgeconv_args* a, const float* y_a, const float* aff_a,
                                           double* partials, float* ymax, float* ymin, uint8_t* amax, uint8_t* amin,
                                           void* stream) {
  EcK p;
  int rc = fill_common(a, p, "edgeconv_conv_b_train");
  if (rc) return rc;
  DG_REQUIRE(a->idx && a->w_b && y_a && aff_a && partials && ymax && ymin && amax && amin, DGCNN_E_INVALID,
             "edgeconv_conv_b_train: null pointer");
  p.y_a_in = y_a; p.aff_a = aff_a; p.partials = partials;
  p.ymax = ymax; p.ymin = ymin; p.amax = amax; p.amin = amin;
  if (p.tc)
    return launch_ec_tc_fwd(p, MODE_TRAIN_B, persistent_grid(p.num_tiles), static_cast<cudaStream_t>(stream));
  return launch_fwd<64, MODE_TRAIN_B>(p, static_cast<cudaStream_t>(stream));
}

extern "C" int dgcnn_bn_finalize(const double* partials, int C, double count, const float* gamma, const float* beta,
                                 float eps, float momentum, float* running_mean, float* running_var, float* mean,
                                 float* rstd, float* aff, void* stream) {
  DG_REQUIRE(partials && gamma && beta && mean && rstd && aff, DGCNN_E_INVALID, "bn_finalize: null pointer");
  DG_REQUIRE(C == kCout, DGCNN_E_UNSUPPORTED, "bn_finalize: C must be 64");
  DG_REQUIRE(count >= 1.0 && (!running_mean == !running_var), DGCNN_E_INVALID, "bn_finalize: bad count / running stats");
  bn_finalize_kernel<<<kFinCtas, kFinThreads, 0, static_cast<cudaStream_t>(stream)>>>(partials, C, count, gamma, beta, eps, momentum,
                                                                    running_mean, running_var, mean, rstd, aff);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_edgeconv_select(int M, const float* ymax, const float* ymin, const uint8_t* amax,
                                     const uint8_t* amin, const float* aff, float* out, int ld_out, float* ysel,
                                     uint8_t* arg, void* stream) {
  DG_REQUIRE(ymax && ymin && amax && amin && aff && out && ysel && arg, DGCNN_E_INVALID, "edgeconv_select: null pointer");
  DG_REQUIRE(M > 0 && ld_out >= kCout, DGCNN_E_INVALID, "edgeconv_select: bad shape");
  const size_t total = static_cast<size_t>(M) * kCout;
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, 148 * 16));
  select_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(M, ymax, ymin, amax, amin, aff, out, ld_out, ysel, arg);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_edgeconv_bwd_prep(int M, const float* dout, int ld_dout, const float* ysel, const float* aff,
                                       float* g, double* partials, void* stream) {
  DG_REQUIRE(dout && ysel && aff && g && partials, DGCNN_E_INVALID, "edgeconv_bwd_prep: null pointer");
  DG_REQUIRE(M > 0 && ld_dout >= kCout, DGCNN_E_INVALID, "edgeconv_bwd_prep: bad shape");
  // a latency-bound streaming pass (each thread walks M / (4 blocks) points with dependent 4-byte loads): many CTAs
  const int blocks = std::max(1, std::min((M + 3) / 4, 6 * 148));
  bwd_prep_kernel<<<blocks, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(M, dout, ld_dout, ysel, aff, g, partials);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_bn_bwd_reduce(const double* partials, int C, double count, const float* gamma, const float* mean,
                                   const float* rstd, float* dgamma, float* dbeta, float* coef, void* stream) {
  DG_REQUIRE(partials && gamma && mean && rstd && dgamma && dbeta && coef, DGCNN_E_INVALID, "bn_bwd_reduce: null pointer");
  DG_REQUIRE(C == kCout && count >= 1.0, DGCNN_E_UNSUPPORTED, "bn_bwd_reduce: C must be 64");
  bn_bwd_reduce_kernel<<<kFinCtas, kFinThreads, 0, static_cast<cudaStream_t>(stream)>>>(partials, C, count, gamma, mean, rstd, dgamma, dbeta, coef);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_edgeconv_bwd_b(const dgcnn_edgeconv_args* a, const float* y_a, const float* aff_a, const float* g,
                                    const uint8_t* arg, const float* coef_b, float* dz_a, float* dw_b,
                                    double* partials, float* dw_partials, void* stream) {
  EcK p;
  int rc = fill_common(a, p, "edgeconv_bwd_b");
  if (rc) return rc;
  DG_REQUIRE(a->w_b && y_a && aff_a && g && arg && coef_b && dz_a && dw_b && partials && dw_partials, DGCNN_E_INVALID,
             "edgeconv_bwd_b: null pointer");
  p.y_a_in = y_a; p.aff_a = aff_a; p.partials = partials;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p.tc) {
    // two CTAs per SM; every CTA gets at least one tile so that its dW partial is defined
    const int grid = std::max(1, std::min(p.num_tiles, 2 * std::max(1, num_sms())));
    return launch_ec_tc_bwd_b(p, g, arg, coef_b, dz_a, dw_partials, dw_b, grid, st);
  } else {
    DG_CUDA(cudaFuncSetAttribute(ec_bwd_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(EcBwdSmem::bytes)));
    ec_bwd_b_kernel<<<persistent_grid(p.num_tiles), kThreads, EcBwdSmem::bytes, st>>>(p, g, arg, coef_b, dz_a, dw_partials);
    DG_CUDA(cudaGetLastError());
  }
  dw_reduce_kernel<<<(kCout * kCout + 255) / 256, 256, 0, st>>>(dw_partials, partials, dw_b);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_edgeconv_bwd_b_fused(const dgcnn_edgeconv_args* a, const float* y_a, const float* aff_a, const float* g,
                                          const uint8_t* arg, const float* coef_b, float* r_raw, float* s_raw, int ld_rs, float* yin,
                                          float* cnt, float* dw_b, float* dw_partials, void* stream) {
  EcK p;
  int rc = fill_common(a, p, "edgeconv_bwd_b_fused");
  if (rc) return rc;
  DG_REQUIRE(a->idx && a->w_b && y_a && aff_a && g && arg && coef_b && r_raw && s_raw && yin && cnt && dw_b && dw_partials,
             DGCNN_E_INVALID, "edgeconv_bwd_b_fused: null pointer");
  DG_REQUIRE(a->N >= kMaxP, DGCNN_E_UNSUPPORTED, "edgeconv_bwd_b_fused: N >= 16 (a tile spans at most two clouds)");
  DG_REQUIRE(((reinterpret_cast<uintptr_t>(y_a) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(arg) |
               reinterpret_cast<uintptr_t>(r_raw) | reinterpret_cast<uintptr_t>(yin)) & 15) == 0 && ld_rs >= kCout && ld_rs % 4 == 0,
             DGCNN_E_INVALID, "edgeconv_bwd_b_fused: y_a, g, arg, r_raw, yin need 16-byte alignment; ld_rs %% 4 == 0, ld_rs >= 64");
  DG_REQUIRE(a->k >= 8, DGCNN_E_UNSUPPORTED, "edgeconv_bwd_b_fused: k >= 8 (a thread's 8 consecutive edge rows span at most two points)");
  p.y_a_in = y_a; p.aff_a = aff_a;
  return launch_ec_pipe_bwd_b(p, g, arg, coef_b, r_raw, s_raw, ld_rs, yin, cnt, dw_partials, dw_b, num_sms(), static_cast<cudaStream_t>(stream));
}

extern "C" int dgcnn_edgeconv_bwd_a_point_stats(int M, const float* U, int ld_u, const float* V, int ld_v, const float* r_raw,
                                                const float* s_raw, int ld_rs, double* partials, void* stream) {
  DG_REQUIRE(M > 0 && U && V && r_raw && s_raw && partials && ld_u >= kCout && ld_v >= kCout && ld_rs >= kCout, DGCNN_E_INVALID,
             "edgeconv_bwd_a_point_stats: null pointer / bad leading dimension");
  return launch_ec_bwd_point_stats(U, ld_u, V, ld_v, r_raw, s_raw, ld_rs, M, partials, static_cast<cudaStream_t>(stream));
}

extern "C" int dgcnn_edgeconv_bwd_a_point_apply(int M, int k, float* r, float* s, int ld_rs, const float* yin, const float* ysum,
                                                const float* cnt, const float* coef_a, void* stream) {
  DG_REQUIRE(M > 0 && k > 0 && r && s && yin && ysum && cnt && coef_a, DGCNN_E_INVALID, "edgeconv_bwd_a_point_apply: null pointer");
  DG_REQUIRE(ld_rs >= kCout && ld_rs % 4 == 0 &&
                 ((reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(yin) |
                   reinterpret_cast<uintptr_t>(ysum)) & 15) == 0,
             DGCNN_E_INVALID, "edgeconv_bwd_a_point_apply: r, s, yin, ysum need 16-byte alignment; ld_rs %% 4 == 0, ld_rs >= 64");
  return launch_ec_bwd_point_apply(r, s, ld_rs, yin, ysum, cnt, coef_a, M, k, static_cast<cudaStream_t>(stream));
}

extern "C" int dgcnn_edgeconv_bwd_a(const dgcnn_edgeconv_args* a, const float* y_a, const float* dz_a, const float* g,
                                    const uint8_t* arg, const float* coef_a, float* s, float* r, int ld_rs, void* stream) {
  EcK p;
  int rc = fill_common(a, p, "edgeconv_bwd_a");
  if (rc) return rc;
  DG_REQUIRE(a->idx && y_a && coef_a && s && r && (dz_a || (g && arg)), DGCNN_E_INVALID, "edgeconv_bwd_a: null pointer");
  DG_REQUIRE(ld_rs >= kCout && ld_rs % 4 == 0 && ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(r)) & 15) == 0,
             DGCNN_E_INVALID, "edgeconv_bwd_a: s/r need 16-byte alignment and ld_rs %% 4 == 0, ld_rs >= 64");
  p.y_a_in = y_a;
  const int blocks = std::max(1, std::min((p.M + 15) / 16, 8 * 148));
  ec_bwd_a_kernel<<<blocks, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p, dz_a, g, arg, coef_a, s, r, ld_rs);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}
