// edgeconv_tc.cu — EdgeConv training-forward passes on the 5th-gen tensor cores (C_in = 64 convolutions).
//
// Same tiles and outputs as edgeconv.cu (P*k <= 128 edge rows per tile, persistent CTAs, per-CTA statistics), but the
// 128 x 64 x 64 tile GEMM is a tcgen05.mma with the accumulator in TMEM:
//   * the A tile is produced in shared memory by the CTA itself (gathered x_j - x_i, or LReLU(BN(y_a))), split on the
//     fly into three bf16 planes in the SWIZZLE_128B K-major layout (fence.proxy.async hands it to the tensor core);
//   * the 64 x 64 weight is split once per CTA;
//   * six kind::f16 MMAs per k16 step (bf16x3 split, FP32-accurate, see tc.cuh) — 24 MMAs per tile;
//   * the epilogue reads TMEM (tcgen05.ld), adds the hoisted centre term (computed once per point by
//     center_term_kernel), and parks the FP32 tile in shared memory (aliasing the A planes) for the column statistics,
//     the coalesced y_a store and the max/min over k;
//   * MODE_EVAL is the fully fused inference block: BN_a + LeakyReLU in the epilogue, H re-split into the A planes,
//     conv_b as a second tcgen05 GEMM, BN_b + LeakyReLU, max over k — nothing per-edge touches HBM.
// Two CTAs per SM overlap one CTA's gather/epilogue with the other's MMA.
#include "ec_tc_common.cuh"

namespace dg {

struct EcTcSmem {
  static constexpr uint32_t a_off = 0;                       // 3 A planes, later the FP32 Y tile
  static constexpr uint32_t w_off = 3 * kPlA;                // 3 weight planes (conv_a: Wa1, conv_b: W_b)
  static constexpr uint32_t w2_off = w_off + 3 * kPlW;       // 3 planes of W_b (fused inference kernel only)
  static constexpr uint32_t v_off = w2_off + 3 * kPlW;       // Vs [16][64]: hoisted centre term Wa2 x_i of the tile's points
  static constexpr uint32_t xi_off = v_off + kMaxP * 64 * 4; // xis [16][64]
  static constexpr uint32_t aff_off = xi_off + kMaxP * 64 * 4;   // affine a | affine b (2 x 192 floats)
  static constexpr uint32_t nbr_off = aff_off + 2 * 192 * 4;
  static constexpr uint32_t red_off = nbr_off + 128 * 4;        // [2][4][64] floats for the final statistics flush
  static constexpr uint32_t bar_off = red_off + 2 * 4 * 64 * 4;
  static constexpr uint32_t bytes = bar_off + 64;
};

// V[p][o] = sum_c Wa2[o][c] x[p][c]: the centre half of conv_a (w_a[:, 64:128]), once per point instead of once per edge.
// 64 points per CTA iteration, 4 points x 4 outputs per thread, sequential FMA chain over c (same order as the FMA path).
__global__ void __launch_bounds__(256) center_term_kernel(const float* __restrict__ rows, int ld, int M,
                                                          const float* __restrict__ w_a, float* __restrict__ V) {
  __shared__ __align__(16) float Wt[64][68];      // [c][o]
  __shared__ float xs[64][65];                    // [p][c]
  const int tid = threadIdx.x;
  for (int i = tid; i < 64 * 64; i += 256) Wt[i & 63][i >> 6] = w_a[static_cast<size_t>(i >> 6) * 128 + 64 + (i & 63)];
  const int og = tid & 15, pg = tid >> 4;
  for (int p0 = blockIdx.x * 64; p0 < M; p0 += gridDim.x * 64) {
    __syncthreads();
    for (int i = tid; i < 64 * 16; i += 256) {
      const int pp = i >> 4, c4 = i & 15;
      const float4 v = p0 + pp < M ? *reinterpret_cast<const float4*>(rows + static_cast<size_t>(p0 + pp) * ld + 4 * c4)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
      xs[pp][4 * c4] = v.x; xs[pp][4 * c4 + 1] = v.y; xs[pp][4 * c4 + 2] = v.z; xs[pp][4 * c4 + 3] = v.w;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
    for (int c = 0; c < 64; ++c) {
      const float4 w = *reinterpret_cast<const float4*>(&Wt[c][4 * og]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float x = xs[4 * pg + i][c];
        acc[i][0] = __fmaf_rn(w.x, x, acc[i][0]);
        acc[i][1] = __fmaf_rn(w.y, x, acc[i][1]);
        acc[i][2] = __fmaf_rn(w.z, x, acc[i][2]);
        acc[i][3] = __fmaf_rn(w.w, x, acc[i][3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (p0 + 4 * pg + i < M)
        *reinterpret_cast<float4*>(V + static_cast<size_t>(p0 + 4 * pg + i) * 64 + 4 * og) =
            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 2) ec_tc_fwd_kernel(const EcK p) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* Ap = sm + EcTcSmem::a_off;
  float* Ys = reinterpret_cast<float*>(sm + EcTcSmem::a_off);
  unsigned char* Wp = sm + EcTcSmem::w_off;
  unsigned char* Wp2 = sm + EcTcSmem::w2_off;
  float* Vs = reinterpret_cast<float*>(sm + EcTcSmem::v_off);
  float* xis = reinterpret_cast<float*>(sm + EcTcSmem::xi_off);
  float* aff = reinterpret_cast<float*>(sm + EcTcSmem::aff_off);
  float* affB = aff + 192;
  int* nbr = reinterpret_cast<int*>(sm + EcTcSmem::nbr_off);
  float* red = reinterpret_cast<float*>(sm + EcTcSmem::red_off);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + EcTcSmem::bar_off);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = p.k, P = p.P, M = p.M;

  if (tid == 0) {
    if (tc::smem_u32(sm) & 1023u) __trap();
    tc::mbar_init(bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 64);
  // weights: B operand rows = output channel, K = input channel.  conv_a uses Wa1 = w_a[:, 0:64]; conv_b uses w_b.
  {
    const float* W = MODE == MODE_TRAIN_B ? p.w_b : p.w_a;
    const int ldw = MODE == MODE_TRAIN_B ? 64 : 128;
    for (int i = tid; i < 64 * 8; i += kThreads) {
      const int o = i >> 3, ch = i & 7;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = W[static_cast<size_t>(o) * ldw + ch * 8 + e];
      store_split3(Wp, kPlW, o, ch, v);
    }
    if (MODE == MODE_EVAL && p.two) {
      for (int i = tid; i < 64 * 8; i += kThreads) {
        const int o = i >> 3, ch = i & 7;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = p.w_b[static_cast<size_t>(o) * 64 + ch * 8 + e];
        store_split3(Wp2, kPlW, o, ch, v);
      }
    }
    if (tid < 192) {
      aff[tid] = p.aff_a ? p.aff_a[tid] : 0.f;
      affB[tid] = p.aff_b ? p.aff_b[tid] : 0.f;
    }
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t taddr = *tmem_slot;
  constexpr uint32_t idesc = tc::idesc_bf16(128, 64);
  uint32_t phase = 0;

  // epilogue mapping: warp -> TMEM lane quarter (warp & 3) and column half (warp >> 2); thread = one tile row
  const int quarter = warp & 3, half = warp >> 2;
  const int erow = quarter * 32 + lane;
  const int erp = min(erow / k, kMaxP - 1);
  // statistics mapping: thread -> column (tid & 63), row group (tid >> 6) of 32 rows
  const int sc = tid & 63, sg = tid >> 6;
  float t1 = 0.f, t2 = 0.f;

  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const int gp0 = tile * P;
    const size_t e0 = static_cast<size_t>(gp0) * k;
    const int nvalid = min(P, M - gp0) * k;
    __syncthreads();  // S0: previous tile's Y tile / Vs / nbr fully consumed
    if (MODE != MODE_TRAIN_B) {
      if (tid < kTileRows) {
        const int pp = tid / k, slot = tid - pp * k, gp = gp0 + pp;
        nbr[tid] = (pp < P && gp < M) ? (gp / p.N) * p.N + p.idx[static_cast<size_t>(gp) * k + slot] : -1;
      }
      for (int i = tid; i < P * 16; i += kThreads) {
        const int pp = i >> 4, c4 = i & 15, gp = gp0 + pp;
        *reinterpret_cast<float4*>(xis + pp * 64 + 4 * c4) =
            gp < M ? *reinterpret_cast<const float4*>(p.rows + static_cast<size_t>(gp) * p.ld + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(Vs + pp * 64 + 4 * c4) =
            gp < M ? *reinterpret_cast<const float4*>(p.v_center + static_cast<size_t>(gp) * 64 + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();  // S1
    }
    // ---- produce the A planes: all global loads of the thread are issued before the first use ------------------------
    {
      float4 ra[4], rb[4];
      bool ok[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int i = tid + it * kThreads, row = i >> 3, ch = i & 7;
        const float* src;
        if (MODE != MODE_TRAIN_B) {
          const int j = nbr[row];
          ok[it] = j >= 0;
          src = p.rows + static_cast<size_t>(max(j, 0)) * p.ld + ch * 8;
        } else {
          ok[it] = row < nvalid;
          src = p.y_a_in + (e0 + min(row, max(nvalid - 1, 0))) * kCout + ch * 8;
        }
        ra[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        rb[it] = ra[it];
        if (ok[it]) {
          ra[it] = *reinterpret_cast<const float4*>(src);
          rb[it] = *reinterpret_cast<const float4*>(src + 4);
        }
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int i = tid + it * kThreads, row = i >> 3, ch = i & 7;
        float v[8] = {ra[it].x, ra[it].y, ra[it].z, ra[it].w, rb[it].x, rb[it].y, rb[it].z, rb[it].w};
        if (MODE != MODE_TRAIN_B) {
          if (ok[it]) {
            const float* xi = xis + (row / k) * 64 + ch * 8;
            const float4 c = *reinterpret_cast<const float4*>(xi), d = *reinterpret_cast<const float4*>(xi + 4);
            v[0] -= c.x; v[1] -= c.y; v[2] -= c.z; v[3] -= c.w;       // x_j - x_i (model.py:81)
            v[4] -= d.x; v[5] -= d.y; v[6] -= d.z; v[7] -= d.w;
          }
        } else if (ok[it]) {
          float pm[8], ps[8], pb[8];
          ld8(aff + ch * 8, pm); ld8(aff + 64 + ch * 8, ps); ld8(aff + 128 + ch * 8, pb);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = lrelu(bn_apply(v[e], pm[e], ps[e], pb[e]));      // h = LReLU(BN_a(y_a))
        }
        store_split3(Ap, kPlA, row, ch, v);
      }
    }
    tc::fence_proxy_async();           // generic-proxy smem writes -> visible to the tensor core
    __syncthreads();                   // S2
    if (tid == 0) {
      tc::fence_after_sync();
      tc::mma_split_katom<3>(taddr, tc::smem_u32(Ap), kPlA, tc::smem_u32(Wp), kPlW, idesc, true);
      tc::mma_commit(bar);
    }
    __syncwarp();
    tc::mbar_wait(bar, phase);
    phase ^= 1;
    tc::fence_after_sync();
    // ---- epilogue: TMEM -> FP32 tile in smem (over the A planes, which the MMA has finished reading) ------------
    if (MODE == MODE_EVAL) {
      // fused inference block: BN_a + LeakyReLU [-> conv_b on the tensor cores -> BN_b + LeakyReLU] -> max over k
      float v[32];
      tc::tmem_ld32(taddr + (static_cast<uint32_t>(quarter * 32) << 16) + half * 32, v);
      const float* vr = Vs + erp * kCout + half * 32;
      const bool valid = erow < nvalid;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float c4[4], pm[4], ps[4], pb[4];
        ld4v(vr + i, c4); ld4v(aff + half * 32 + i, pm); ld4v(aff + 64 + half * 32 + i, ps); ld4v(aff + 128 + half * 32 + i, pb);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[i + e] = valid ? lrelu(bn_apply(v[i + e] + c4[e], pm[e], ps[e], pb[e])) : 0.f;
      }
      if (p.two) {
        tc::fence_before_sync();
        __syncthreads();               // every warp has drained the accumulator; the A planes can be overwritten with H
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          float h[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) h[e] = v[c8 * 8 + e];
          store_split3(Ap, kPlA, erow, half * 4 + c8, h);
        }
        tc::fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
          tc::fence_after_sync();
          tc::mma_split_katom<3>(taddr, tc::smem_u32(Ap), kPlA, tc::smem_u32(Wp2), kPlW, idesc, true);
          tc::mma_commit(bar);
        }
        __syncwarp();
        tc::mbar_wait(bar, phase);
        phase ^= 1;
        tc::fence_after_sync();
        tc::tmem_ld32(taddr + (static_cast<uint32_t>(quarter * 32) << 16) + half * 32, v);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float pm[4], ps[4], pb[4];
          ld4v(affB + half * 32 + i, pm); ld4v(affB + 64 + half * 32 + i, ps); ld4v(affB + 128 + half * 32 + i, pb);
#pragma unroll
          for (int e = 0; e < 4; ++e) v[i + e] = lrelu(bn_apply(v[i + e], pm[e], ps[e], pb[e]));
        }
      }
      float4* dst = reinterpret_cast<float4*>(Ys + erow * kYs + half * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
      float v[32];
      tc::tmem_ld32(taddr + (static_cast<uint32_t>(quarter * 32) << 16) + half * 32, v);
      float4* dst = reinterpret_cast<float4*>(Ys + erow * kYs + half * 32);
      if (MODE == MODE_TRAIN_A) {
        const float4* vr = reinterpret_cast<const float4*>(Vs + erp * kCout + half * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 c = vr[i];
          dst[i] = make_float4(v[4 * i] + c.x, v[4 * i + 1] + c.y, v[4 * i + 2] + c.z, v[4 * i + 3] + c.w);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    }
    tc::fence_before_sync();
    __syncthreads();                   // S3: Y tile complete
    if (MODE == MODE_EVAL) {
      for (int i = tid; i < P * kCout; i += kThreads) {
        const int pp = i >> 6, c = i & 63, gp = gp0 + pp;
        if (gp < M) {
          float m = Ys[(pp * k) * kYs + c];
          for (int s2 = 1; s2 < k; ++s2) m = fmaxf(m, Ys[(pp * k + s2) * kYs + c]);
          p.out[static_cast<size_t>(gp) * p.ld_out + c] = m;
        }
      }
      continue;
    }
    // ---- batch statistics of the raw conv output (valid rows only) ----------------------------------------------
    {
      const int r_lo = sg * 32, r_hi = min(r_lo + 32, nvalid);
      float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};      // four chains: the loop is latency-bound otherwise
      int r = r_lo;
      for (; r + 3 < r_hi; r += 4) {
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float y = Ys[(r + c4) * kYs + sc];
          a1[c4] += y;
          a2[c4] = __fmaf_rn(y, y, a2[c4]);
        }
      }
      for (; r < r_hi; ++r) {
        const float y = Ys[r * kYs + sc];
        a1[0] += y;
        a2[0] = __fmaf_rn(y, y, a2[0]);
      }
      t1 += (a1[0] + a1[1]) + (a1[2] + a1[3]);
      t2 += (a2[0] + a2[1]) + (a2[2] + a2[3]);
    }
    if (MODE == MODE_TRAIN_A) {
      for (int i = tid; i < nvalid * 16; i += kThreads) {
        const int row = i >> 4, c4 = i & 15;
        *reinterpret_cast<float4*>(p.y_a_out + (e0 + row) * kCout + 4 * c4) = *reinterpret_cast<const float4*>(Ys + row * kYs + 4 * c4);
      }
    }
    if (MODE == MODE_TRAIN_B || !p.two) {
      for (int i = tid; i < P * kCout; i += kThreads) {
        const int pp = i >> 6, c = i & 63, gp = gp0 + pp;
        if (gp < M) {
          float mx = Ys[(pp * k) * kYs + c], mn = mx;
          int ax = 0, an = 0;
          for (int s = 1; s < k; ++s) {
            const float v = Ys[(pp * k + s) * kYs + c];
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
  // ---- per-CTA statistics -> partials ---------------------------------------------------------------------------------
  __syncthreads();
  if (MODE != MODE_EVAL) {
    red[(0 * 4 + sg) * 64 + sc] = t1;
    red[(1 * 4 + sg) * 64 + sc] = t2;
    __syncthreads();
    if (tid < 128) {
      const int w = tid >> 6, c = tid & 63;
      const double s = ((static_cast<double>(red[(w * 4 + 0) * 64 + c]) + red[(w * 4 + 1) * 64 + c]) + red[(w * 4 + 2) * 64 + c]) +
                       red[(w * 4 + 3) * 64 + c];
      p.partials[128 + static_cast<size_t>(blockIdx.x) * 128 + tid] = s;
    }
    if (blockIdx.x == 0 && tid == 0) p.partials[0] = static_cast<double>(gridDim.x);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(taddr, 64);
}

// ---------------------------------------------------------------------------------------------
// backward edge pass through conv_b on the tensor cores (same contract as ec_bwd_b_kernel in edgeconv.cu)
//
// The dense BN_b backward is dy_b = cA dz - cB - cC y_b with y_b = H W_b^T (dz: the max-over-k gradient, non-zero at the
// arg-max rows only).  Substituting y_b removes the recomputation of y_b and the epilogue that depended on it:
//   dh   = dy_b W_b           = D' W_b - H G,        D' = cA dz - cB,   G = W_b^T diag(cC) W_b   (64 x 64, once per CTA)
//   dW_b = dy_b^T H           = D'^T H - diag(cC) W_b (H^T H)
// so a tile is ONE batch of MMAs after ONE producer phase:
//   MMA a  dh    = D' W_b      A = D' planes (K-major), B = the W_b planes read MN-major
//   MMA b  dh   -= H G         A = H planes (K-major),  B = planes of -G (symmetric, K-major)
//   MMA c  T    += [D' ; H]^T H  MN-major views of the same tiles (contraction over the 128 tile rows); A stacked to
//                              M = 128 through the descriptor's leading-dimension offset, accumulated in TMEM over all
//                              tiles of the CTA: lanes 0-63 = D'^T H, lanes 64-127 = H^T H
// all with the 2-plane bf16 split: this is the gradient path, 2^-17 relative error is far below its 1e-3 noise floor
// (tests/test_gpu_parity.py), and two planes per operand let two CTAs share an SM.  ec_tc_dw_finish_kernel turns the
// per-CTA T partials into dW_b.
// ---------------------------------------------------------------------------------------------
struct EcTcBwdSmem {
  static constexpr uint32_t dy_off = 0;                         // DY planes (2)
  static constexpr uint32_t h_off = 2 * kPlA;                   // H planes (2)
  static constexpr uint32_t w_off = h_off + 2 * kPlA;           // W_b planes (2): rows = c (out), 128 B of c'' (in): K-major
                                                                // for y_b = H W_b^T, MN-major view (N = c'', K = c) for dh
  static constexpr uint32_t gm_off = w_off + 2 * kPlW;          // planes (2) of -G = -W_b^T diag(cC) W_b (K-major, symmetric)
  static constexpr uint32_t g_off = gm_off + 2 * kPlW;          // gs [16][64]
  static constexpr uint32_t aff_off = g_off + kMaxP * 64 * 4;   // affine of BN_a (192) + coefficients of BN_b (192)
  static constexpr uint32_t arg_off = aff_off + 2 * 192 * 4;    // args [16][64] u8
  static constexpr uint32_t red_off = arg_off + kMaxP * 64;     // [2][4][64] floats
  static constexpr uint32_t bar_off = red_off + 2 * 4 * 64 * 4;
  static constexpr uint32_t bytes = bar_off + 64;
};

__global__ void __launch_bounds__(kThreads, 2)
ec_tc_bwd_b_kernel(const EcK p, const float* __restrict__ g, const uint8_t* __restrict__ arg,
                   const float* __restrict__ coef_b, float* __restrict__ dz_a, float* __restrict__ dw_partials) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* DYp = sm + EcTcBwdSmem::dy_off;
  unsigned char* Hp = sm + EcTcBwdSmem::h_off;
  unsigned char* Wp = sm + EcTcBwdSmem::w_off;
  unsigned char* Gp = sm + EcTcBwdSmem::gm_off;
  float* gs = reinterpret_cast<float*>(sm + EcTcBwdSmem::g_off);
  float* aff = reinterpret_cast<float*>(sm + EcTcBwdSmem::aff_off);
  float* cb = aff + 192;
  uint8_t* args = sm + EcTcBwdSmem::arg_off;
  float* red = reinterpret_cast<float*>(sm + EcTcBwdSmem::red_off);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + EcTcBwdSmem::bar_off);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = p.k, P = p.P, M = p.M;

  if (tid == 0) {
    if (tc::smem_u32(sm) & 1023u) __trap();
    tc::mbar_init(bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 128);     // columns 0-63: y_b / dh, columns 64-127: dW_b accumulator
  for (int i = tid; i < 64 * 8; i += kThreads) {
    const int o = i >> 3, ch = i & 7;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = p.w_b[static_cast<size_t>(o) * 64 + ch * 8 + e];   // row o = out channel c, K = c''
    store_split<2>(Wp, kPlW, o, ch, v);
  }
  if (tid < 192) {
    aff[tid] = p.aff_a[tid];
    cb[tid] = coef_b[tid];
  }
  {
    // -G[i][j] = -sum_c cC_c W_b[c][i] W_b[c][j]; FP32 copy of W_b parked in the (still unused) DY plane area
    float* Wf = reinterpret_cast<float*>(DYp);      // [c][65]
    for (int i = tid; i < 64 * 64; i += kThreads) Wf[(i >> 6) * 65 + (i & 63)] = p.w_b[i];
    __syncthreads();
    const int gi = tid >> 2, gj0 = (tid & 3) * 16;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    for (int c = 0; c < 64; ++c) {
      const float a = -cb[128 + c] * Wf[c * 65 + gi];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = __fmaf_rn(a, Wf[c * 65 + gj0 + j], acc[j]);
    }
    float lo[8], hi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { lo[j] = acc[j]; hi[j] = acc[8 + j]; }
    store_split<2>(Gp, kPlW, gi, (tid & 3) * 2, lo);          // row n = i, K chunk of j: -G is symmetric
    store_split<2>(Gp, kPlW, gi, (tid & 3) * 2 + 1, hi);
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t taddr = *tmem_slot;
  constexpr uint32_t idesc_k = tc::idesc_bf16(128, 64);
  constexpr uint32_t idesc_mn = tc::idesc_bf16(128, 64) | (1u << 15) | (1u << 16);
  constexpr uint32_t idesc_kb = tc::idesc_bf16(128, 64) | (1u << 16);            // A K-major, B MN-major
  uint32_t phase = 0;
  bool first_tile = true;

  const int quarter = warp & 3, half = warp >> 2;
  const int erow = quarter * 32 + lane;
  const int erp = min(erow / k, kMaxP - 1), ers = erow - (erow / k) * k;
  float t1 = 0.f, t2 = 0.f;     // sums of dz_a and dz_a*y_a for column half*32+lane over this warp's rows

  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const int gp0 = tile * P;
    const size_t e0 = static_cast<size_t>(gp0) * k;
    const int nvalid = min(P, M - gp0) * k;
    // (no barrier here: gs / args of the previous tile were last read before its S1)
    if (tid == 0 && tile + static_cast<int>(gridDim.x) < p.num_tiles) {      // DRAM -> L2 prefetch of this CTA's next y_a tile
      const size_t e2 = static_cast<size_t>(tile + gridDim.x) * P * k;
      const size_t rows_left = static_cast<size_t>(M) * k - e2;
      tc::prefetch_l2(p.y_a_in + e2 * kCout, static_cast<uint32_t>(min(rows_left, static_cast<size_t>(P) * k) * kCout * 4));
    }
    for (int i = tid; i < P * kCout; i += kThreads) {
      const int gp = gp0 + (i >> 6);
      const bool ok = gp < M;
      gs[i] = ok ? g[static_cast<size_t>(gp) * kCout + (i & 63)] : 0.f;
      args[i] = ok ? arg[static_cast<size_t>(gp) * kCout + (i & 63)] : 255;
    }
    // producer with the EPILOGUE's mapping (thread = tile row erow, 32 columns of half `half`): the thread's y_a values
    // stay in registers until the epilogue, so y_a is read from memory once per tile
    float ya[32];
    const bool valid = erow < nvalid;
    {
      const float* src = p.y_a_in + (e0 + (valid ? erow : 0)) * kCout + half * 32;
#pragma unroll
      for (int i = 0; i < 32; ++i) ya[i] = 0.f;
      if (valid) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) ld_global_v8(src + i, ya + i);      // y_a rows are 256-byte aligned (cudaMalloc'd [E,64])
      }
    }
    __syncthreads();  // S0b: gs / args staged
    // per-column constants and the point's (g, arg) come in as 16-byte / 8-byte shared loads: scalar loads here made
    // the kernel MIO-bound (ncu: mio_throttle was the top stall)
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
      const int c0 = half * 32 + c8 * 8;
      float pm[8], ps[8], pb[8], kA[8], kB[8], gg[8];
      ld8(aff + c0, pm); ld8(aff + 64 + c0, ps); ld8(aff + 128 + c0, pb);
      ld8(cb + c0, kA); ld8(cb + 64 + c0, kB);
      ld8(gs + erp * kCout + c0, gg);
      const uint2 a8 = *reinterpret_cast<const uint2*>(args + erp * kCout + c0);
      float h[8], d[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        h[e] = valid ? lrelu(bn_apply(ya[c8 * 8 + e], pm[e], ps[e], pb[e])) : 0.f;
        const int ab = static_cast<int>(((e < 4 ? a8.x : a8.y) >> (8 * (e & 3))) & 0xffu);
        const float dz = (ab == ers) ? gg[e] : 0.f;
        d[e] = valid ? __fmaf_rn(kA[e], dz, -kB[e]) : 0.f;                          // D' = cA dz - cB
      }
      store_split<2>(Hp, kPlA, erow, half * 4 + c8, h);
      store_split<2>(DYp, kPlA, erow, half * 4 + c8, d);
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();  // S1: H and D' planes complete, accumulator of the previous tile drained
    if (tid == 0) {
      tc::fence_after_sync();
      // dh = D' W_b: A = D' planes (K-major), B = the W_b planes read MN-major (N = c'' contiguous, K = c = rows)
      {
        uint32_t acc3 = 0u;
        const int pa[3] = {0, 1, 0}, pb[3] = {1, 0, 0};
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) {
            tc::mma_bf16(taddr, tc::smem_desc_sw128(tc::smem_u32(DYp) + pa[q] * kPlA + k16 * 32),
                         smem_desc_mn(tc::smem_u32(Wp) + pb[q] * kPlW + k16 * 2048, 1024), idesc_kb, acc3);
            acc3 = 1u;
          }
      }
      tc::mma_split_katom<2>(taddr, tc::smem_u32(Hp), kPlA, tc::smem_u32(Gp), kPlW, idesc_k, false);     // dh -= H G
      // T += [D' ; H]^T H over the 128 tile rows: pairs (plane i, plane j), i + j <= 3, smallest first
      const uint32_t lbo = tc::smem_u32(Hp) - tc::smem_u32(DYp);
      uint32_t acc = first_tile ? 0u : 1u;
      const int pi[3] = {0, 1, 0}, pj[3] = {1, 0, 0};
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int k16 = 0; k16 < 8; ++k16) {
          tc::mma_bf16(taddr + 64, smem_desc_mn(tc::smem_u32(DYp) + pi[q] * kPlA + k16 * 2048, lbo),
                       smem_desc_mn(tc::smem_u32(Hp) + pj[q] * kPlA + k16 * 2048, 1024), idesc_mn, acc);
          acc = 1u;
        }
      tc::mma_commit(bar);
    }
    first_tile = false;
    __syncwarp();
    tc::mbar_wait(bar, phase);
    phase ^= 1;
    tc::fence_after_sync();
    // ---- epilogue 2: dz_a = dh * LReLU'(BN_a(y_a)); store; column sums of dz_a and dz_a*y_a ------------------------
    {
      float v[32];
      tc::tmem_ld32(taddr + (static_cast<uint32_t>(quarter * 32) << 16) + half * 32, v);
      float (&w)[32] = ya;                      // dz_a * y_a overwrites the cached y_a values
      if (valid) {
        float* dst = dz_a + (e0 + erow) * kCout + half * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float o[4];
          const float4 pm = *reinterpret_cast<const float4*>(aff + half * 32 + i);
          const float4 ps = *reinterpret_cast<const float4*>(aff + 64 + half * 32 + i);
          const float4 pb = *reinterpret_cast<const float4*>(aff + 128 + half * 32 + i);
          const float pmv[4] = {pm.x, pm.y, pm.z, pm.w}, psv[4] = {ps.x, ps.y, ps.z, ps.w}, pbv[4] = {pb.x, pb.y, pb.z, pb.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float z = bn_apply(ya[i + e], pmv[e], psv[e], pbv[e]);
            o[e] = v[i + e] * (z > 0.f ? 1.f : kSlope);
            v[i + e] = o[e];
            w[i + e] = o[e] * ya[i + e];
          }
          if (i & 4) st_global_v8(dst + i - 4, v + i - 4);      // whole 32-byte sectors of the thread's own row
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) { v[i] = 0.f; w[i] = 0.f; }
      }
      warp_transpose_reduce(v, lane);
      warp_transpose_reduce(w, lane);
      t1 += v[0];
      t2 += w[0];
    }
    tc::fence_before_sync();
  }
  // ---- dW_b partial of this CTA (TMEM lanes 0-63, columns 64-127) and the BN_a statistics -----------------------------
  __syncthreads();
  if (warp < 4 && !first_tile) {            // T partial of this CTA: rows 0-63 = D'^T H, rows 64-127 = H^T H
    float* dst = dw_partials + static_cast<size_t>(blockIdx.x) * 2 * kCout * kCout + static_cast<size_t>(warp * 32 + lane) * kCout;
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 32) {
      float v[32];
      tc::tmem_ld32(taddr + (static_cast<uint32_t>(warp * 32) << 16) + 64 + c0, v);
#pragma unroll
      for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
  }
  red[(0 * 4 + quarter) * 64 + half * 32 + lane] = t1;
  red[(1 * 4 + quarter) * 64 + half * 32 + lane] = t2;
  __syncthreads();
  if (tid < 128) {
    const int w = tid >> 6, c = tid & 63;
    const double s = ((static_cast<double>(red[(w * 4 + 0) * 64 + c]) + red[(w * 4 + 1) * 64 + c]) + red[(w * 4 + 2) * 64 + c]) +
                     red[(w * 4 + 3) * 64 + c];
    p.partials[128 + static_cast<size_t>(blockIdx.x) * 128 + tid] = s;
  }
  if (blockIdx.x == 0 && tid == 0) p.partials[0] = static_cast<double>(gridDim.x);
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(taddr, 128);
}

// T[128][64] = sum over the CTAs of their partials (FP64, fixed order) -> scratch behind the partials.
// 32 outputs x 8 part-groups per CTA: group g adds parts g, g+8, ... (coalesced 128-byte rows), then the 8 group sums
// are added in order.
__global__ void __launch_bounds__(256) ec_tc_dw_sum_kernel(const float* __restrict__ dw_partials, int nparts, float* __restrict__ T) {
  __shared__ double red[8][32];
  const int o = blockIdx.x * 32 + (threadIdx.x & 31), g = threadIdx.x >> 5;
  double s = 0.0;
#pragma unroll 4
  for (int c = g; c < nparts; c += 8) s += dw_partials[static_cast<size_t>(c) * 2 * kCout * kCout + o];
  red[g][threadIdx.x & 31] = s;
  __syncthreads();
  if (g == 0) {
    double a = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) a += red[q][threadIdx.x];
    T[o] = static_cast<float>(a);
  }
}
// dW_b[c][j] = T[c][j] - cC_c * sum_i W_b[c][i] T[64 + i][j]
__global__ void __launch_bounds__(256) ec_tc_dw_finish_kernel(const float* __restrict__ T, const float* __restrict__ w_b,
                                                              const float* __restrict__ coef_b, float* __restrict__ dw) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= kCout * kCout) return;
  const int c = o >> 6, j = o & 63;
  double a = 0.0;
  for (int i = 0; i < 64; ++i) a += static_cast<double>(w_b[c * 64 + i]) * T[(64 + i) * 64 + j];
  dw[o] = static_cast<float>(static_cast<double>(T[o]) - static_cast<double>(coef_b[128 + c]) * a);
}

int launch_ec_dw_finish(const float* dw_partials, int nparts, const float* w_b, const float* coef_b, float* dw_b, cudaStream_t st) {
  float* T = const_cast<float*>(dw_partials) + static_cast<size_t>(kMaxPartialCtas) * 2 * kCout * kCout;
  ec_tc_dw_sum_kernel<<<2 * kCout * kCout / 32, 256, 0, st>>>(dw_partials, nparts, T);
  ec_tc_dw_finish_kernel<<<(kCout * kCout + 255) / 256, 256, 0, st>>>(T, w_b, coef_b, dw_b);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

int launch_ec_tc_bwd_b(const EcK& p, const float* g, const uint8_t* arg, const float* coef_b, float* dz_a,
                       float* dw_partials, float* dw_b, int grid, cudaStream_t st) {
  DG_REQUIRE((reinterpret_cast<uintptr_t>(p.y_a_in) & 31) == 0 && (reinterpret_cast<uintptr_t>(dz_a) & 31) == 0, DGCNN_E_INVALID,
             "edgeconv_bwd_b (tensor-core path): y_a and dz_a must be 32-byte aligned");
  DG_CUDA(cudaFuncSetAttribute(ec_tc_bwd_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(EcTcBwdSmem::bytes)));
  ec_tc_bwd_b_kernel<<<grid, kThreads, EcTcBwdSmem::bytes, st>>>(p, g, arg, coef_b, dz_a, dw_partials);
  DG_CUDA(cudaGetLastError());
  return launch_ec_dw_finish(dw_partials, grid, p.w_b, coef_b, dw_b, st);
}

int launch_ec_tc_fwd(const EcK& p, int mode, int grid, cudaStream_t st) {
  if (mode != MODE_TRAIN_B) {
    center_term_kernel<<<std::max(1, std::min((p.M + 63) / 64, 148 * 4)), 256, 0, st>>>(p.rows, p.ld, p.M, p.w_a, p.v_center);
    DG_CUDA(cudaGetLastError());
  }
  // conv_b streams a contiguous tile: the warp-specialised pipeline wins there.  conv_a is bound by the L2 gather and
  // shared-memory bandwidth either way and runs faster as two serial CTAs per SM.
  if (mode == MODE_TRAIN_B && p.tc >= 2) return launch_ec_pipe_conv_b(p, (grid + 1) / 2, st);
  if (mode == MODE_TRAIN_A) {
    DG_CUDA(cudaFuncSetAttribute(ec_tc_fwd_kernel<MODE_TRAIN_A>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(EcTcSmem::bytes)));
    ec_tc_fwd_kernel<MODE_TRAIN_A><<<grid, kThreads, EcTcSmem::bytes, st>>>(p);
  } else if (mode == MODE_TRAIN_B) {
    DG_CUDA(cudaFuncSetAttribute(ec_tc_fwd_kernel<MODE_TRAIN_B>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(EcTcSmem::bytes)));
    ec_tc_fwd_kernel<MODE_TRAIN_B><<<grid, kThreads, EcTcSmem::bytes, st>>>(p);
  } else {
    DG_CUDA(cudaFuncSetAttribute(ec_tc_fwd_kernel<MODE_EVAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(EcTcSmem::bytes)));
    ec_tc_fwd_kernel<MODE_EVAL><<<grid, kThreads, EcTcSmem::bytes, st>>>(p);
  }
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

}  // namespace dg
