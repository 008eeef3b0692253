// edgeconv_pipe.cu — warp-specialised, software-pipelined EdgeConv training passes (C_in = 64) for sm_100a.
//
// Same tiles, outputs and statistics contract as edgeconv_tc.cu (P*k <= 128 edge rows per tile, persistent CTAs,
// per-CTA FP64 partials), but the three phases of a tile run in different warps and overlap across tiles:
//
//   converter warps (8)   global -> registers (prefetched one tile ahead; neighbour indices two tiles ahead)
//                         -> x_j - x_i  |  LReLU(BN_a(y_a))  -> bf16x3 split -> A planes[buf]   (SWIZZLE_128B, K-major)
//   MMA warp (1 thread)   tcgen05.mma  A planes[buf] x W planes -> TMEM accumulator[buf]         (24 MMAs per tile)
//   epilogue warps (4)    tcgen05.ld accumulator[buf] -> + centre term -> y_a store / column statistics / max-min over k
//
// Two A-plane buffers and two TMEM accumulators, handed over with mbarriers (a_full/a_free, acc_full/acc_free).
// One CTA (416 threads) per SM.  The serial version in edgeconv_tc.cu exposed one DRAM/L2 round trip, one MMA latency
// and three CTA-wide barriers per tile (5.8k cycles per tile and SM); here a tile costs max(converter, epilogue) time.
#include "ec_tc_common.cuh"

namespace dg {

constexpr int kPipeThreads = 416;      // warps 0-3 epilogue (warp == TMEM lane quarter), warp 4 MMA, warps 5-12 converters
constexpr int kPipeMmaWarp = 4;
constexpr int kPipeConv0 = 160;        // first converter thread
constexpr uint32_t kBufA = 3 * kPlA;   // bytes of one A buffer (3 planes)
constexpr int kYsP = 68;               // row stride of the epilogue tile: 16-byte row stores and column reads are both conflict-free

struct EcPipeSmem {
  static constexpr uint32_t a_off = 0;                          // 2 buffers x 3 planes
  static constexpr uint32_t w_off = 2 * kBufA;                  // 3 weight planes
  static constexpr uint32_t ys_off = w_off + 3 * kPlW;          // FP32 tile [128][kYsP] of the epilogue
  static constexpr uint32_t aff_off = ys_off + 128 * kYsP * 4;   // affine of BN_a (192 floats)
  static constexpr uint32_t red_off = aff_off + 192 * 4;        // [2][2][64] floats for the final statistics flush
  static constexpr uint32_t bar_off = red_off + 2 * 2 * 64 * 4; // 8 mbarriers + TMEM slot
  static constexpr uint32_t bytes = bar_off + 128;
};

template <int MODE>
__global__ void __launch_bounds__(kPipeThreads, 1) ec_pipe_fwd_kernel(const EcK p) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* Ap = sm + EcPipeSmem::a_off;
  unsigned char* Wp = sm + EcPipeSmem::w_off;
  float* Ys = reinterpret_cast<float*>(sm + EcPipeSmem::ys_off);
  float* aff = reinterpret_cast<float*>(sm + EcPipeSmem::aff_off);
  float* red = reinterpret_cast<float*>(sm + EcPipeSmem::red_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + EcPipeSmem::bar_off);
  uint64_t* a_full = bars;          // [2] converters -> MMA      (8 arrivals: one per converter warp)
  uint64_t* a_free = bars + 2;      // [2] MMA commit -> converters
  uint64_t* acc_full = bars + 4;    // [2] MMA commit -> epilogue
  uint64_t* acc_free = bars + 6;    // [2] epilogue -> MMA        (4 arrivals: one per epilogue warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = p.k, P = p.P, M = p.M;
  const int grid = gridDim.x;

  if (tid == 0) {
    if (tc::smem_u32(sm) & 1023u) __trap();
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&a_full[i], 8);
      tc::mbar_init(&a_free[i], 1);
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_free[i], 4);
    }
    tc::fence_barrier_init();
  }
  if (warp == kPipeMmaWarp) tc::tmem_alloc(tmem_slot, 128);
  {
    const float* W = MODE == MODE_TRAIN_B ? p.w_b : p.w_a;      // conv_a uses Wa1 = w_a[:, 0:64]
    const int ldw = MODE == MODE_TRAIN_B ? 64 : 128;
    for (int i = tid; i < 64 * 8; i += kPipeThreads) {
      const int o = i >> 3, ch = i & 7;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = W[static_cast<size_t>(o) * ldw + ch * 8 + e];
      store_split3(Wp, kPlW, o, ch, v);
    }
    if (tid < 192) aff[tid] = p.aff_a ? p.aff_a[tid] : 0.f;
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t taddr = *tmem_slot;

  if (warp > kPipeMmaWarp) {
    // =========================================== converters ===========================================================
    const int c = tid - kPipeConv0, ch = c & 7, rb = c >> 3;      // rows rb + 32 j, 8-channel chunk ch
    int pp[4], slot[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = rb + 32 * j;
      pp[j] = r / k;
      slot[j] = r - pp[j] * k;
    }
    float bm[8], bs[8], bb[8];
    if (MODE == MODE_TRAIN_B) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        bm[e] = aff[ch * 8 + e];
        bs[e] = aff[64 + ch * 8 + e];
        bb[e] = aff[128 + ch * 8 + e];
      }
    }
    int jn[4];                       // TRAIN_A: global row of the neighbour (or -1); TRAIN_B: 0 / -1 validity
    auto load_idx = [&](int tile) {
      const int gp0 = tile * P;
      const int nvalid = tile < p.num_tiles ? min(P, M - gp0) * k : 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = rb + 32 * j;
        jn[j] = -1;
        if (r < nvalid) {
          if (MODE == MODE_TRAIN_B) {
            jn[j] = 0;
          } else {
            const int gp = gp0 + pp[j];
            jn[j] = (gp / p.N) * p.N + __ldg(p.idx + static_cast<size_t>(gp) * k + slot[j]);
          }
        }
      }
    };
    auto load_rows = [&](int tile, float4 (&d)[8], uint32_t& mask) {
      const size_t e0 = static_cast<size_t>(tile) * P * k;
      mask = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        d[2 * j] = make_float4(0.f, 0.f, 0.f, 0.f);
        d[2 * j + 1] = d[2 * j];
        if (jn[j] >= 0) {
          const float* src = MODE == MODE_TRAIN_B ? p.y_a_in + (e0 + rb + 32 * j) * kCout + ch * 8
                                                  : p.rows + static_cast<size_t>(jn[j]) * p.ld + ch * 8;
          d[2 * j] = *reinterpret_cast<const float4*>(src);
          d[2 * j + 1] = *reinterpret_cast<const float4*>(src + 4);
          mask |= 1u << j;
        }
      }
    };
    // one pipeline step: request tile+1 (rows) and tile+2 (indices), then convert the tile whose rows are in `cur`.
    // The two register sets alternate by unrolling (a register copy would wait for the loads just issued).
    auto step = [&](int tile, int it, float4 (&cur)[8], uint32_t& curmask, float4 (&nxt)[8], uint32_t& nxtmask) {
      const int b = it & 1;
      const uint32_t u = static_cast<uint32_t>(it) >> 1;
      load_rows(tile + grid, nxt, nxtmask);            // tile+1 rows: their indices were requested one tile ago
      load_idx(tile + 2 * grid);
      mbar_wait_wd(&a_free[b], (u & 1u) ^ 1u);    // the MMA that read this buffer two tiles ago has completed
      tc::fence_after_sync();
      unsigned char* A = Ap + b * kBufA;
      const int gp0 = tile * P;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v[8] = {cur[2 * j].x, cur[2 * j].y, cur[2 * j].z, cur[2 * j].w,
                      cur[2 * j + 1].x, cur[2 * j + 1].y, cur[2 * j + 1].z, cur[2 * j + 1].w};
        if (curmask & (1u << j)) {
          if (MODE != MODE_TRAIN_B) {
            const float* xi = p.rows + static_cast<size_t>(gp0 + pp[j]) * p.ld + ch * 8;
            const float4 a = __ldg(reinterpret_cast<const float4*>(xi)), d = __ldg(reinterpret_cast<const float4*>(xi + 4));
            v[0] -= a.x; v[1] -= a.y; v[2] -= a.z; v[3] -= a.w;       // x_j - x_i (model.py:81)
            v[4] -= d.x; v[5] -= d.y; v[6] -= d.z; v[7] -= d.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = lrelu(bn_apply(v[e], bm[e], bs[e], bb[e]));   // h = LReLU(BN_a(y_a))
          }
        }
        store_split3(A, kPlA, rb + 32 * j, ch, v);
      }
      tc::fence_proxy_async();           // generic-proxy smem writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&a_full[b]);
    };
    float4 ra[8], rb2[8];
    uint32_t ma, mb;
    int tile = blockIdx.x;
    load_idx(tile);
    load_rows(tile, ra, ma);
    load_idx(tile + grid);
    for (int it = 0; tile < p.num_tiles; tile += 2 * grid, it += 2) {
      step(tile, it, ra, ma, rb2, mb);
      if (tile + grid < p.num_tiles) step(tile + grid, it + 1, rb2, mb, ra, ma);
    }
  } else if (warp == kPipeMmaWarp) {
    // =========================================== MMA issuer ===========================================================
    constexpr uint32_t idesc = tc::idesc_bf16(128, 64);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += grid, ++it) {
      const int b = it & 1;
      const uint32_t u = static_cast<uint32_t>(it) >> 1;
      if (lane == 0) {
        mbar_wait_wd(&a_full[b], u & 1u);
        mbar_wait_wd(&acc_free[b], (u & 1u) ^ 1u);
        tc::fence_after_sync();
        tc::mma_split_katom<3>(taddr + b * 64, tc::smem_u32(Ap + b * kBufA), kPlA, tc::smem_u32(Wp), kPlW, idesc, true);
        tc::mma_commit(&a_free[b]);
        tc::mma_commit(&acc_full[b]);
      }
      __syncwarp();
    }
  } else {
    // =========================================== epilogue =============================================================
    const int quarter = warp, et = tid;
    float* Yg = Ys;
    const int erow = et;                                   // TMEM lane = tile row
    const int erp = min(erow / k, kMaxP - 1);
    const int sc = et & 63, sg = et >> 6;
    float t1 = 0.f, t2 = 0.f;
    for (int it = 0, tile = blockIdx.x; tile < p.num_tiles; tile += grid, ++it) {
      const uint32_t b = it & 1, u = static_cast<uint32_t>(it) >> 1;
      const int gp0 = tile * P;
      const size_t e0 = static_cast<size_t>(gp0) * k;
      const int nvalid = min(P, M - gp0) * k;
      const bool valid = erow < nvalid;
      float4 vc[16];                                       // centre term of the row's point
      const float* vrow = p.v_center + static_cast<size_t>(min(gp0 + erp, M - 1)) * kCout;
      if (MODE == MODE_TRAIN_A) {
#pragma unroll
        for (int i = 0; i < 16; ++i) vc[i] = __ldg(reinterpret_cast<const float4*>(vrow) + i);
      }
      mbar_wait_wd(&acc_full[b], u & 1u);
      tc::fence_after_sync();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float v[32];
        tc::tmem_ld32(taddr + (static_cast<uint32_t>(quarter * 32) << 16) + b * 64 + half * 32, v);
        if (half == 1) {
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&acc_free[b]);    // accumulator drained: the MMA of tile+2 may overwrite it
        }
        if (MODE == MODE_TRAIN_A) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 q = vc[half * 8 + i];
            v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
          }
          if (valid) {        // 32-byte stores: every lane writes whole sectors of its own row
            float* dst = p.y_a_out + (e0 + erow) * kCout + half * 32;
#pragma unroll
            for (int i = 0; i < 4; ++i) st_global_v8(dst + 8 * i, v + 8 * i);
          }
        }
        float4* ys = reinterpret_cast<float4*>(Yg + erow * kYsP + half * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) ys[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
      named_bar_sync(1, 128);                          // Y tile complete
      if (MODE == MODE_TRAIN_B || !p.two) {
        // thread = (column sc, points sg, sg+2, ...): batch statistics and max/min over the k rows of a point in one pass;
        // two points x even/odd rows = four independent chains (a single chain is latency-bound)
        for (int q0 = sg; q0 < P; q0 += 4) {
          float mx[4], mn[4], s1[4], s2[4];
          int ax[4], an[4];
          const float* base[2];
          bool ok[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int q = q0 + 2 * h;
            ok[h] = q < P && gp0 + q < M;
            base[h] = Yg + (ok[h] ? q : q0) * k * kYsP + sc;
          }
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const float y = base[c4 >> 1][(c4 & 1) * kYsP];          // rows 0 and 1 (k >= 2)
            mx[c4] = y; mn[c4] = y; ax[c4] = c4 & 1; an[c4] = c4 & 1;
            s1[c4] = y; s2[c4] = y * y;
          }
          int s = 2;
          for (; s + 1 < k; s += 2) {
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const int r = s + (c4 & 1);
              const float y = base[c4 >> 1][r * kYsP];
              if (y > mx[c4]) { mx[c4] = y; ax[c4] = r; }
              if (y < mn[c4]) { mn[c4] = y; an[c4] = r; }
              s1[c4] += y;
              s2[c4] = __fmaf_rn(y, y, s2[c4]);
            }
          }
          if (s < k) {                                               // odd k: last row joins the even chains
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float y = base[h][s * kYsP];
              if (y > mx[2 * h]) { mx[2 * h] = y; ax[2 * h] = s; }
              if (y < mn[2 * h]) { mn[2 * h] = y; an[2 * h] = s; }
              s1[2 * h] += y;
              s2[2 * h] = __fmaf_rn(y, y, s2[2 * h]);
            }
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (!ok[h]) continue;
            const int e = 2 * h, o2 = 2 * h + 1;
            // first occurrence wins ties (model.py:149 max semantics as implemented by the reference kernels)
            const bool tmx = mx[o2] > mx[e] || (mx[o2] == mx[e] && ax[o2] < ax[e]);
            const bool tmn = mn[o2] < mn[e] || (mn[o2] == mn[e] && an[o2] < an[e]);
            const size_t o = static_cast<size_t>(gp0 + q0 + 2 * h) * kCout + sc;
            p.ymax[o] = tmx ? mx[o2] : mx[e];
            p.ymin[o] = tmn ? mn[o2] : mn[e];
            p.amax[o] = static_cast<uint8_t>(tmx ? ax[o2] : ax[e]);
            p.amin[o] = static_cast<uint8_t>(tmn ? an[o2] : an[e]);
            t1 += s1[e] + s1[o2];
            t2 += s2[e] + s2[o2];
          }
        }
      } else {
        const int r_lo = sg * 64, r_hi = min(r_lo + 64, nvalid);
        float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
        int r = r_lo;
        for (; r + 3 < r_hi; r += 4) {
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const float y = Yg[(r + c4) * kYsP + sc];
            a1[c4] += y;
            a2[c4] = __fmaf_rn(y, y, a2[c4]);
          }
        }
        for (; r < r_hi; ++r) {
          const float y = Yg[r * kYsP + sc];
          a1[0] += y;
          a2[0] = __fmaf_rn(y, y, a2[0]);
        }
        t1 += (a1[0] + a1[1]) + (a1[2] + a1[3]);
        t2 += (a2[0] + a2[1]) + (a2[2] + a2[3]);
      }
      named_bar_sync(1, 128);                          // Y tile consumed
    }
    // ---- per-CTA statistics -> partials (fixed order) ---------------------------------------------------------------
    red[(0 * 2 + sg) * 64 + sc] = t1;
    red[(1 * 2 + sg) * 64 + sc] = t2;
    named_bar_sync(1, 128);
    {
      const int w = tid >> 6, cc = tid & 63;
      const double sum = static_cast<double>(red[(w * 2 + 0) * 64 + cc]) + red[(w * 2 + 1) * 64 + cc];
      p.partials[128 + static_cast<size_t>(blockIdx.x) * 128 + tid] = sum;
    }
    if (blockIdx.x == 0 && tid == 0) p.partials[0] = static_cast<double>(gridDim.x);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == kPipeMmaWarp) tc::tmem_dealloc(taddr, 128);
}

int launch_ec_pipe_fwd(const EcK& p, int mode, int sms, cudaStream_t st) {
  const int grid = std::max(1, std::min(p.num_tiles, std::max(1, sms)));
  if (mode == MODE_TRAIN_A) {
    DG_CUDA(cudaFuncSetAttribute(ec_pipe_fwd_kernel<MODE_TRAIN_A>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(EcPipeSmem::bytes)));
    ec_pipe_fwd_kernel<MODE_TRAIN_A><<<grid, kPipeThreads, EcPipeSmem::bytes, st>>>(p);
  } else {
    DG_CUDA(cudaFuncSetAttribute(ec_pipe_fwd_kernel<MODE_TRAIN_B>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(EcPipeSmem::bytes)));
    ec_pipe_fwd_kernel<MODE_TRAIN_B><<<grid, kPipeThreads, EcPipeSmem::bytes, st>>>(p);
  }
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

}  // namespace dg
