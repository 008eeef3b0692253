// edgeconv_pipe.cu — warp-specialised, software-pipelined conv_b training pass (h = LReLU(BN_a(y_a)); y_b = W_b h) for sm_100a.
//
// Same tiles, outputs and statistics contract as MODE_TRAIN_B of edgeconv_tc.cu (P*k <= 128 edge rows per tile,
// persistent CTAs, per-CTA FP64 partials), but the phases of a tile run in different warps and overlap across tiles:
//
//   producer warp (1)     cp.async.bulk (TMA engine): the contiguous y_a tile (one <= 32 KB copy) -> raw FP32 stage[buf],
//                         mbarrier complete_tx
//   converter warps (8)   raw stage -> LReLU(BN_a(y_a)) -> bf16x3 split -> A planes[buf]   (SWIZZLE_128B, K-major)
//   MMA warp (1 thread)   tcgen05.mma  A planes[buf] x W planes -> TMEM accumulator[buf]   (24 MMAs per tile)
//   epilogue warps (8)    tcgen05.ld accumulator[buf] (lane quarter x column half) -> FP32 tile in smem -> column
//                         statistics and max/min over the k rows of every point in one pass (4 independent chains)
//
// Two raw stages, two A-plane buffers and two TMEM accumulators, handed over with mbarriers (raw_full/raw_free,
// a_full/a_free, acc_full/acc_free).  One CTA (576 threads) per SM.
//
// Measured (B=8, N=7000, k=30; 14000 tiles): 0.17 ms against 0.28 ms for the serial kernel.  What bounds it now is
// shared-memory bandwidth: per tile 32 KB raw in + 32 KB raw out + 48 KB plane stores + 144 KB of operand reads by the
// 24 N=64 MMAs + 64 KB through the epilogue tile = 320 KB = 2500 cycles at 128 B/clk.  conv_a (gather through L2) was
// tried in this form with register prefetch, 256-byte bulk copies and 16-byte cp.async gathers and stayed slower than
// the serial two-CTA kernel (0.39-0.53 ms against 0.33 ms), so it keeps that one.  Lessons kept in the code:
//   * fence.proxy.async compiles to MEMBAR and waits for the thread's outstanding global loads, so a converter cannot
//     prefetch the next tile into registers across it — the tile is staged by the async engine instead;
//   * 256-byte cp.async.bulk copies cost ~65 cycles each (a 128-row gather took 8.6k cycles).
#include "ec_tc_common.cuh"

namespace dg {

constexpr int kPipeThreads = 576;      // warps 0-7 epilogue (TMEM lane quarter x column half), 8 MMA, 9 producer, 10-17 converters
constexpr int kPipeMmaWarp = 8;
constexpr int kPipeProdWarp = 9;
constexpr int kPipeConvWarp0 = 10;
constexpr int kPipeConv0 = kPipeConvWarp0 * 32;   // first converter thread
constexpr uint32_t kBufA = 3 * kPlA;   // bytes of one A buffer (3 planes)
constexpr int kYsP = 68;               // row stride of the epilogue tile: 16-byte row stores and column reads are both conflict-free

struct EcPipeSmem {
  static constexpr uint32_t a_off = 0;                          // 2 buffers x 3 planes
  static constexpr uint32_t w_off = 2 * kBufA;                  // 3 weight planes
  static constexpr uint32_t raw_off = w_off + 3 * kPlW;         // 2 raw FP32 stages [128][64] filled by bulk copies
  static constexpr uint32_t ys_off = raw_off + 2 * 128 * 64 * 4;   // FP32 tile [128][kYsP] of the epilogue
  static constexpr uint32_t aff_off = ys_off + 128 * kYsP * 4;   // affine of BN_a (192 floats)
  static constexpr uint32_t red_off = aff_off + 192 * 4;        // [2][4][64] floats for the final statistics flush
  static constexpr uint32_t bar_off = red_off + 2 * 4 * 64 * 4; // 12 mbarriers + TMEM slot
  static constexpr uint32_t bytes = bar_off + 128;
};

__global__ void __launch_bounds__(kPipeThreads, 1) ec_pipe_conv_b_kernel(const EcK p) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* Ap = sm + EcPipeSmem::a_off;
  unsigned char* Wp = sm + EcPipeSmem::w_off;
  float* Raw = reinterpret_cast<float*>(sm + EcPipeSmem::raw_off);
  float* Ys = reinterpret_cast<float*>(sm + EcPipeSmem::ys_off);
  float* aff = reinterpret_cast<float*>(sm + EcPipeSmem::aff_off);
  float* red = reinterpret_cast<float*>(sm + EcPipeSmem::red_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + EcPipeSmem::bar_off);
  uint64_t* a_full = bars;          // [2] converters -> MMA      (8 arrivals: one per converter warp)
  uint64_t* a_free = bars + 2;      // [2] MMA commit -> converters
  uint64_t* acc_full = bars + 4;    // [2] MMA commit -> epilogue
  uint64_t* acc_free = bars + 6;    // [2] epilogue -> MMA        (8 arrivals: one per epilogue warp)
  uint64_t* raw_full = bars + 8;    // [2] bulk copy (expect_tx) -> converters
  uint64_t* raw_free = bars + 10;   // [2] converters -> producer (8 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = p.k, P = p.P, M = p.M;
  const int grid = gridDim.x;

  if (tid == 0) {
    if (tc::smem_u32(sm) & 1023u) __trap();
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&a_full[i], 8);
      tc::mbar_init(&a_free[i], 1);
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_free[i], 8);
      tc::mbar_init(&raw_full[i], 1);
      tc::mbar_init(&raw_free[i], 8);
    }
    tc::fence_barrier_init();
  }
  if (warp == kPipeMmaWarp) tc::tmem_alloc(tmem_slot, 128);
  for (int i = tid; i < 64 * 8; i += kPipeThreads) {      // W_b: B operand rows = output channel, K = input channel
    const int o = i >> 3, ch = i & 7;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = p.w_b[static_cast<size_t>(o) * 64 + ch * 8 + e];
    store_split3(Wp, kPlW, o, ch, v);
  }
  if (tid < 192) aff[tid] = p.aff_a[tid];
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t taddr = *tmem_slot;

  if (warp >= kPipeConvWarp0) {
    // =========================================== converters ===========================================================
    const int c = tid - kPipeConv0, ch = c & 7, rb = c >> 3;      // rows rb + 32 j, 8-channel chunk ch
    float bm[8], bs[8], bb[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      bm[e] = aff[ch * 8 + e];
      bs[e] = aff[64 + ch * 8 + e];
      bb[e] = aff[128 + ch * 8 + e];
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += grid, ++it) {
      const int b = it & 1;
      const uint32_t u = static_cast<uint32_t>(it) >> 1;
      const int nvalid = min(P, M - tile * P) * k;
      const float* raw = Raw + b * (128 * 64);
      unsigned char* A = Ap + b * kBufA;
      mbar_wait_wd(&raw_full[b], u & 1u);                // the bulk copy of this tile has landed
      mbar_wait_wd(&a_free[b], (u & 1u) ^ 1u);           // the MMA that read this A buffer two tiles ago has completed
      tc::fence_after_sync();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = rb + 32 * j;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (r < nvalid) {
          const float4 lo = *reinterpret_cast<const float4*>(raw + r * 64 + ch * 8);
          const float4 hi = *reinterpret_cast<const float4*>(raw + r * 64 + ch * 8 + 4);
          v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = lrelu(bn_apply(v[e], bm[e], bs[e], bb[e]));   // h = LReLU(BN_a(y_a))
        }
        store_split3(A, kPlA, r, ch, v);
      }
      tc::fence_proxy_async();           // generic-proxy smem writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(&a_full[b]);
        tc::mbar_arrive(&raw_free[b]);
      }
    }
  } else if (warp == kPipeProdWarp) {
    // =========================================== producer: bulk copy of the y_a tile =====================================
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += grid, ++it) {
      const int b = it & 1;
      const uint32_t u = static_cast<uint32_t>(it) >> 1;
      const int gp0 = tile * P;
      const uint32_t bytes = static_cast<uint32_t>(min(P, M - gp0) * k) * 256u;
      if (lane == 0) {
        mbar_wait_wd(&raw_free[b], (u & 1u) ^ 1u);
        tc::mbar_arrive_expect_tx(&raw_full[b], bytes);
        tc::bulk_g2s(Raw + b * (128 * 64), p.y_a_in + static_cast<size_t>(gp0) * k * kCout, bytes, &raw_full[b]);
      }
      __syncwarp();
    }
  } else if (warp == kPipeMmaWarp) {
    // =========================================== MMA issuer ===========================================================
    constexpr uint32_t idesc = tc::idesc_bf16(128, 64);
    const bool leader = tc::elect_one();     // uniform control flow + elect.sync: the issue code stays on the uniform datapath
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += grid, ++it) {
      const int b = it & 1;
      const uint32_t u = static_cast<uint32_t>(it) >> 1;
      mbar_wait_wd(&a_full[b], u & 1u);
      mbar_wait_wd(&acc_free[b], (u & 1u) ^ 1u);
      tc::fence_after_sync();
      if (leader) {
        tc::mma_split_katom<3>(taddr + b * 64, tc::smem_u32(Ap + b * kBufA), kPlA, tc::smem_u32(Wp), kPlW, idesc, true);
        tc::mma_commit(&a_free[b]);
        tc::mma_commit(&acc_full[b]);
      }
      __syncwarp();
    }
  } else {
    // =========================================== epilogue =============================================================
    // warp w: TMEM lane quarter w & 3 (hardware rule), column half w >> 2; thread = one tile row x 32 columns
    const int quarter = warp & 3, half = warp >> 2;
    const int erow = quarter * 32 + lane;
    const int sc = tid & 63, sg = tid >> 6;                // second phase: column sc, points sg, sg + 4, ...
    float t1 = 0.f, t2 = 0.f;
    for (int it = 0, tile = blockIdx.x; tile < p.num_tiles; tile += grid, ++it) {
      const uint32_t b = it & 1, u = static_cast<uint32_t>(it) >> 1;
      const int gp0 = tile * P;
      mbar_wait_wd(&acc_full[b], u & 1u);
      tc::fence_after_sync();
      {
        float v[32];
        tc::tmem_ld32(taddr + (static_cast<uint32_t>(quarter * 32) << 16) + b * 64 + half * 32, v);
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&acc_free[b]);      // accumulator drained: the MMA of tile+2 may overwrite it
        float4* ys = reinterpret_cast<float4*>(Ys + erow * kYsP + half * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) ys[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
      named_bar_sync(1, 256);                              // Y tile complete
      // thread = (column sc, points sg, sg+4, ...): batch statistics and max/min over the k rows of a point in one pass;
      // rows r = c (mod 4) form four independent chains (a single chain is latency-bound)
      for (int q = sg; q < P; q += 4) {
        if (gp0 + q >= M) break;
        const float* base = Ys + q * k * kYsP + sc;
        float mx[4], mn[4], s1[4], s2[4];
        int ax[4], an[4];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          mx[c4] = -INFINITY; mn[c4] = INFINITY; ax[c4] = 0; an[c4] = 0; s1[c4] = 0.f; s2[c4] = 0.f;
        }
        int s = 0;
        for (; s + 3 < k; s += 4) {
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const float y = base[(s + c4) * kYsP];
            if (y > mx[c4]) { mx[c4] = y; ax[c4] = s + c4; }
            if (y < mn[c4]) { mn[c4] = y; an[c4] = s + c4; }
            s1[c4] += y;
            s2[c4] = __fmaf_rn(y, y, s2[c4]);
          }
        }
#pragma unroll
        for (int c4 = 0; c4 < 3; ++c4) {
          if (s + c4 < k) {
            const float y = base[(s + c4) * kYsP];
            if (y > mx[c4]) { mx[c4] = y; ax[c4] = s + c4; }
            if (y < mn[c4]) { mn[c4] = y; an[c4] = s + c4; }
            s1[c4] += y;
            s2[c4] = __fmaf_rn(y, y, s2[c4]);
          }
        }
        // merge the chains; the first occurrence (lowest row) wins ties, as in the serial and FMA kernels
#pragma unroll
        for (int c4 = 1; c4 < 4; ++c4) {
          if (mx[c4] > mx[0] || (mx[c4] == mx[0] && ax[c4] < ax[0])) { mx[0] = mx[c4]; ax[0] = ax[c4]; }
          if (mn[c4] < mn[0] || (mn[c4] == mn[0] && an[c4] < an[0])) { mn[0] = mn[c4]; an[0] = an[c4]; }
        }
        const size_t o = static_cast<size_t>(gp0 + q) * kCout + sc;
        p.ymax[o] = mx[0];
        p.ymin[o] = mn[0];
        p.amax[o] = static_cast<uint8_t>(ax[0]);
        p.amin[o] = static_cast<uint8_t>(an[0]);
        t1 += (s1[0] + s1[1]) + (s1[2] + s1[3]);
        t2 += (s2[0] + s2[1]) + (s2[2] + s2[3]);
      }
      named_bar_sync(1, 256);                              // Y tile consumed
    }
    // ---- per-CTA statistics -> partials (fixed order) ---------------------------------------------------------------
    red[(0 * 4 + sg) * 64 + sc] = t1;
    red[(1 * 4 + sg) * 64 + sc] = t2;
    named_bar_sync(1, 256);
    if (tid < 128) {
      const int w = tid >> 6, cc = tid & 63;
      const double sum = ((static_cast<double>(red[(w * 4 + 0) * 64 + cc]) + red[(w * 4 + 1) * 64 + cc]) + red[(w * 4 + 2) * 64 + cc]) +
                         red[(w * 4 + 3) * 64 + cc];
      p.partials[128 + static_cast<size_t>(blockIdx.x) * 128 + tid] = sum;
    }
    if (blockIdx.x == 0 && tid == 0) p.partials[0] = static_cast<double>(gridDim.x);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == kPipeMmaWarp) tc::tmem_dealloc(taddr, 128);
}

int launch_ec_pipe_conv_b(const EcK& p, int sms, cudaStream_t st) {
  const int grid = std::max(1, std::min(p.num_tiles, std::max(1, sms)));
  DG_CUDA(cudaFuncSetAttribute(ec_pipe_conv_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(EcPipeSmem::bytes)));
  ec_pipe_conv_b_kernel<<<grid, kPipeThreads, EcPipeSmem::bytes, st>>>(p);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

}  // namespace dg
