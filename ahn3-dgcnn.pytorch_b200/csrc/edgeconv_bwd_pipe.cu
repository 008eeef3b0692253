// edgeconv_bwd_pipe.cu — warp-specialised backward of a two-convolution EdgeConv block through conv_b and the max over k,
// with the scatter of the conv_a gradient fused in: nothing per-edge is written to HBM.
//
// Same algebra as ec_tc_bwd_b_kernel (edgeconv_tc.cu): with D' = cA dz - cB (dz: the max-over-k gradient, non-zero at the
// arg rows only), H = LReLU(BN_a(y_a)) and G = W_b^T diag(cC) W_b,
//     dh   = D' W_b - H G                         (two batches of MMAs into one accumulator)
//     T   += [D' ; H]^T H                         (dW_b = T_top - diag(cC) W_b T_bottom, finished by ec_tc_dw_finish_kernel)
//     dz_a = dh * LReLU'(BN_a(y_a))               (the gradient at conv_a's output BEFORE the BN_a backward)
// What is new is everything after dz_a.  BN_a's backward is dy_a = cA' dz_a - cB' - cC' y_a with coefficients that need
// the batch sums of dz_a and dz_a*y_a, i.e. a second pass over dz_a — unless the pass is linear, and it is: conv_a is
// y_a(i,j) = U_j + V_i, its backward only needs r_j = sum over the in-edges of j and s_i = sum over the out-edges of i of
// dy_a, and sums commute with the affine map:
//     r_j = cA' R_j - cnt_j cB' - cC' Yin_j,     R_j  = sum_in  dz_a,   Yin_j  = sum_in  y_a,   cnt_j = in-degree
//     s_i = cA' S_i - k     cB' - cC' Ysum_i,    S_i  = sum_out dz_a,   Ysum_i = sum_out y_a
//     sum dz_a     = sum_i S_i,          sum dz_a*y_a = sum_j U_j.R_j + sum_i V_i.S_i        (per channel)
// Ysum comes out of the forward conv_a pass (dgcnn_edgeconv_lin_conv_a_train); Yin and cnt are accumulated here, by the
// converter warps, from the y_a values that pass through their registers anyway.
// So this kernel scatters the RAW dz_a rows (float reductions into the L2-resident [M,64] targets R and S) and three
// point-level kernels (14 MB each) finish the job; dz_a[E,64] (430 MB at the training shape, written here and read back
// by ec_lin_bwd_a_kernel<true> in the two-pass form) no longer exists.
//
// Pipeline (one 576-thread CTA per SM, tiles of P*k <= 128 edge rows):
//   producer warp      cp.async.bulk: y_a tile in two 64-row halves (ring of three 16 KB slots) + the tile's g rows and arg
//                      bytes -> raw[slot] / aux[buf]
//   converter warps(8) raw -> H = LReLU(BN_a(y_a)) -> 2 bf16 planes H[buf] (+ one sign byte per 8 channels); then patch the
//                      SINGLE D' buffer: it holds the constant background -cB in every row, the tile's P*64 non-zeros are
//                      written over it (and the previous tile's are restored) — no dense D' generation
//   MMA warp           12 MMAs D' W_b (B read MN-major), 24 MMAs T += [D';H]^T H, commit d_free (D' may be re-patched
//                      while the next 12 run), 12 MMAs -H G, commit acc_full[buf]
//   epilogue warps(8)  tcgen05.ld dh -> * LReLU' (sign bytes) -> FP32 dz tile [128][64] in its own 32 KB buffer (XOR-swizzled
//                      16-byte pieces: row-per-thread stores and row-contiguous loads are both conflict-free) -> R scatter
//                      (two 256-byte rows per warp instruction; a thread walks 8 consecutive rows and keeps their running
//                      sum, flushed into S when the point changes)
// H buffers are released by the MMA commit, the accumulators by the epilogue's tcgen05.ld, the raw slots by the converters:
// no stage waits for a later one.
#include "ec_tc_common.cuh"

namespace dg {

constexpr int kBpThreads = 576;
constexpr int kBpMmaWarp = 8;
constexpr int kBpProdWarp = 9;
constexpr int kBpConvWarp0 = 10;
constexpr int kBpConv0 = kBpConvWarp0 * 32;
constexpr uint32_t kBufH = 2 * kPlA;       // one H buffer: 2 planes = 32 KB = one FP32 [128][64] tile

struct EcBpSmem {
  static constexpr uint32_t d_off = 0;                              // D' planes (2), single buffer
  static constexpr uint32_t h_off = d_off + 2 * kPlA;               // 2 H buffers (behind D': the stacked [D';H] descriptor needs H - D' > 0)
  static constexpr uint32_t w_off = h_off + 2 * kBufH;              // W_b planes (2)
  static constexpr uint32_t gm_off = w_off + 2 * kPlW;              // planes (2) of -G
  static constexpr uint32_t raw_off = gm_off + 2 * kPlW;            // ring of 3 raw FP32 half tiles [64][64]
  static constexpr uint32_t dz_off = raw_off + 3 * 64 * 64 * 4;     // FP32 dz tile [128][64] of the epilogue
  static constexpr uint32_t g_off = dz_off + 128 * 64 * 4;          // 2 x g rows [kMaxP][64]
  static constexpr uint32_t arg_off = g_off + 2 * kMaxP * 64 * 4;   // 2 x arg bytes [kMaxP][64]
  static constexpr uint32_t sign_off = arg_off + 2 * kMaxP * 64;    // 4 x sign bytes [128][8] (ring over 4 tiles)
  static constexpr uint32_t aff_off = sign_off + 4 * 128 * 8;       // affine of BN_a (192) + coefficients of BN_b (192)
  static constexpr uint32_t bar_off = aff_off + 2 * 192 * 4;        // 15 mbarriers + TMEM slot
  static constexpr uint32_t bytes = bar_off + 128;
};
static_assert(EcBpSmem::bytes <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void red_add_v4(float* dst, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(kBpThreads, 1)
ec_pipe_bwd_b_kernel(const EcK p, const float* __restrict__ g, const uint8_t* __restrict__ arg, const float* __restrict__ coef_b,
                     float* __restrict__ r_raw, float* __restrict__ s_raw, int ld_rs, float* __restrict__ dw_partials,
                     float* __restrict__ yin, float* __restrict__ cnt) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* Dp = sm + EcBpSmem::d_off;
  unsigned char* Hp = sm + EcBpSmem::h_off;
  unsigned char* Wp = sm + EcBpSmem::w_off;
  unsigned char* Gp = sm + EcBpSmem::gm_off;
  float* Raw = reinterpret_cast<float*>(sm + EcBpSmem::raw_off);
  float* Gs = reinterpret_cast<float*>(sm + EcBpSmem::g_off);
  uint8_t* Args = sm + EcBpSmem::arg_off;
  uint8_t* Signs = sm + EcBpSmem::sign_off;
  float* aff = reinterpret_cast<float*>(sm + EcBpSmem::aff_off);
  float* cb = aff + 192;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + EcBpSmem::bar_off);
  uint64_t* in_full = bars;         // [2] converters -> MMA: H[buf] planes and the D' patch of this tile (8 arrivals)
  uint64_t* a_free = bars + 2;      // [2] MMA commit -> converters: every MMA that reads H[buf] has completed
  uint64_t* acc_full = bars + 4;    // [2] MMA commit -> epilogue
  uint64_t* acc_free = bars + 6;    // [2] epilogue -> MMA (8 arrivals)
  uint64_t* raw_full = bars + 8;    // [3] bulk copies (expect_tx) -> converters
  uint64_t* raw_free = bars + 11;   // [3] converters -> producer (8 arrivals)
  uint64_t* d_free = bars + 14;     // MMA commit -> converters: every MMA that reads D' of the previous tile has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = p.k, P = p.P, M = p.M;
  const int grid = gridDim.x;

  if (tid == 0) {
    if (tc::smem_u32(sm) & 1023u) __trap();
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&in_full[i], 8);
      tc::mbar_init(&a_free[i], 1);
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_free[i], 8);
    }
    for (int i = 0; i < 3; ++i) {
      tc::mbar_init(&raw_full[i], 1);
      tc::mbar_init(&raw_free[i], 8);
    }
    tc::mbar_init(d_free, 1);
    tc::fence_barrier_init();
  }
  if (warp == kBpMmaWarp) tc::tmem_alloc(tmem_slot, 256);   // columns 0-127: two dh accumulators, 128-191: T
  for (int i = tid; i < 64 * 8; i += kBpThreads) {
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
    // -G[i][j] = -sum_c cC_c W_b[c][i] W_b[c][j]; FP32 copy of W_b parked in the (still unused) raw ring
    float* Wf = Raw;      // [c][65]
    for (int i = tid; i < 64 * 64; i += kBpThreads) Wf[(i >> 6) * 65 + (i & 63)] = p.w_b[i];
    __syncthreads();
    if (tid < 256) {
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
    // the constant background of D': -cB_c in every row
    for (int i = tid; i < 128 * 8; i += kBpThreads) {
      const int row = i >> 3, ch = i & 7;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = -cb[64 + ch * 8 + e];
      store_split<2>(Dp, kPlA, row, ch, v);
    }
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t taddr = *tmem_slot;

  if (warp >= kBpConvWarp0) {
    // =========================================== converters ===========================================================
    const int c = tid - kBpConv0, ch = c & 7, rb = c >> 3;      // rows rb + 32 j, 8-channel chunk ch
    const int col = c & 63;                                     // the D' column this thread patches
    float bm[8], bs[8], bb[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      bm[e] = aff[ch * 8 + e];
      bs[e] = aff[64 + ch * 8 + e];
      bb[e] = aff[128 + ch * 8 + e];
    }
    const float kA = cb[col], kB = cb[64 + col];
    const __nv_bfloat16 bg_hi = __float2bfloat16_rn(-kB);
    const __nv_bfloat16 bg_lo = __float2bfloat16_rn(-kB - __bfloat162float(bg_hi));
    uint32_t prev[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};   // byte offsets of the previous tile's patches
    // 16-byte pieces of the chunk in an order that keeps the 8 lanes of a row on 8 different bank groups
    const int first = (ch >= 4) ? 4 : 0, second = 4 - first;
    // Yin_j += y_a rows, cnt_j += 1 (the in-edge sums of y_a and the in-degrees that BN_a's backward needs per point): done
    // here, from the registers the raw tile passes through.  The forward conv_a pass, where y_a is produced, is bound by L2
    // throughput; this kernel leaves it unused.  Neighbour indices are fetched one tile ahead.
    int ppj[4], jn[4];
    {
      const int nv0 = min(P, M - static_cast<int>(blockIdx.x) * P) * k;
      const size_t en = static_cast<size_t>(blockIdx.x) * P * k;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ppj[j] = (rb + 32 * j) / k;
        jn[j] = (rb + 32 * j < nv0) ? p.idx[en + rb + 32 * j] : 0;
      }
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += grid, ++it) {
      const int b = it & 1;
      const uint32_t u = static_cast<uint32_t>(it) >> 1;
      const int pv = min(P, M - tile * P);
      const int nvalid = pv * k;
      int jc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) jc[j] = jn[j];
      if (tile + grid < p.num_tiles) {
        const int nvn = min(P, M - (tile + grid) * P) * k;
        const size_t en = static_cast<size_t>(tile + grid) * P * k;
#pragma unroll
        for (int j = 0; j < 4; ++j) jn[j] = (rb + 32 * j < nvn) ? p.idx[en + rb + 32 * j] : 0;
      }
      const int bat0 = (tile * P) / p.N * p.N, bat1 = bat0 + p.N;        // a tile spans at most two clouds (P <= 16 <= N)
      unsigned char* H = Hp + b * kBufH;
      uint8_t* signs = Signs + (it & 3) * (128 * 8);     // read by the epilogue of this tile before it frees the accumulator
      mbar_wait_wd(&a_free[b], (u & 1u) ^ 1u);           // the MMAs that read this H buffer two tiles ago have completed
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const uint32_t step = 2u * static_cast<uint32_t>(it) + hh, slot = step % 3u, use = step / 3u;
        const float* raw = Raw + slot * (64 * 64);
        mbar_wait_wd(&raw_full[slot], use & 1u);         // this half of the y_a tile (and, with the first, g / arg) has landed
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int rl = rb + 32 * j, r = rl + 64 * hh;
          float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          uint32_t sb = 0;
          if (r < nvalid) {
            const float4 q0 = *reinterpret_cast<const float4*>(raw + rl * 64 + ch * 8 + first);
            const float4 q1 = *reinterpret_cast<const float4*>(raw + rl * 64 + ch * 8 + second);
            {
              const int gp = tile * P + ppj[2 * hh + j];
              const size_t jrow = static_cast<size_t>((gp >= bat1 ? bat1 : bat0) + jc[2 * hh + j]);
              red_add_v4(yin + jrow * kCout + ch * 8 + first, q0);
              red_add_v4(yin + jrow * kCout + ch * 8 + second, q1);
              if (ch == 0) atomicAdd(cnt + jrow, 1.f);
            }
            const float4 lo = first ? q1 : q0, hi = first ? q0 : q1;
            v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float z = bn_apply(v[e], bm[e], bs[e], bb[e]);
              sb |= (z > 0.f ? 1u : 0u) << e;
              v[e] = lrelu(z);
            }
          }
          signs[r * 8 + ch] = static_cast<uint8_t>(sb);
          store_split<2>(H, kPlA, r, ch, v);
        }
        if (hh == 0) {                                   // (the second half's slot is released after the patch: it guards g / arg)
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&raw_free[slot]);
        }
      }
      // D' patch: restore the background at the previous tile's positions, write this tile's P*64 non-zeros
      mbar_wait_wd(d_free, (static_cast<uint32_t>(it) & 1u) ^ 1u);
      const float* gs = Gs + b * (kMaxP * 64);
      const uint8_t* args = Args + b * (kMaxP * 64);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (prev[q] != 0xFFFFFFFFu) {
          *reinterpret_cast<__nv_bfloat16*>(Dp + prev[q]) = bg_hi;
          *reinterpret_cast<__nv_bfloat16*>(Dp + kPlA + prev[q]) = bg_lo;
          prev[q] = 0xFFFFFFFFu;
        }
        const int pp = (c >> 6) + 4 * q;
        if (pp < pv) {
          const int row = pp * k + args[pp * 64 + col];
          const float d = __fmaf_rn(kA, gs[pp * 64 + col], -kB);
          const __nv_bfloat16 dh = __float2bfloat16_rn(d);
          const __nv_bfloat16 dl = __float2bfloat16_rn(d - __bfloat162float(dh));
          const uint32_t off = tc::unit_elem_offset(row, col);
          *reinterpret_cast<__nv_bfloat16*>(Dp + off) = dh;
          *reinterpret_cast<__nv_bfloat16*>(Dp + kPlA + off) = dl;
          prev[q] = off;
        }
      }
      tc::fence_proxy_async();           // generic-proxy smem writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(&in_full[b]);
        tc::mbar_arrive(&raw_free[(2u * static_cast<uint32_t>(it) + 1u) % 3u]);
      }
    }
  } else if (warp == kBpProdWarp) {
    // =========================================== producer ==============================================================
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += grid, ++it) {
      const int b = it & 1;
      const int gp0 = tile * P;
      const uint32_t pv = static_cast<uint32_t>(min(P, M - gp0));
      const uint32_t rows = pv * static_cast<uint32_t>(k);
      if (lane == 0) {
        for (uint32_t hh = 0; hh < 2; ++hh) {
          const uint32_t step = 2u * static_cast<uint32_t>(it) + hh, slot = step % 3u, use = step / 3u;
          const uint32_t r0 = 64u * hh, nr = rows > r0 ? min(rows - r0, 64u) : 0u;
          mbar_wait_wd(&raw_free[slot], (use & 1u) ^ 1u);
          tc::mbar_arrive_expect_tx(&raw_full[slot], nr * 256u + (hh == 0 ? pv * 320u : 0u));
          if (nr) tc::bulk_g2s(Raw + slot * (64 * 64), p.y_a_in + (static_cast<size_t>(gp0) * k + r0) * kCout, nr * 256u, &raw_full[slot]);
          if (hh == 0) {
            tc::bulk_g2s(Gs + b * (kMaxP * 64), g + static_cast<size_t>(gp0) * kCout, pv * 256u, &raw_full[slot]);
            tc::bulk_g2s(Args + b * (kMaxP * 64), arg + static_cast<size_t>(gp0) * kCout, pv * 64u, &raw_full[slot]);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == kBpMmaWarp) {
    // =========================================== MMA issuer ===========================================================
    constexpr uint32_t idesc_k = tc::idesc_bf16(128, 64);
    constexpr uint32_t idesc_mn = tc::idesc_bf16(128, 64) | (1u << 15) | (1u << 16);
    constexpr uint32_t idesc_kb = tc::idesc_bf16(128, 64) | (1u << 16);            // A K-major, B MN-major
    // the whole warp walks the loop (uniform control flow); the lane picked by elect.sync issues.  Descriptors are built
    // once and stepped by adding to their 14-bit address field, so an MMA costs a few uniform-datapath instructions
    // instead of ~18 (this warp is a single in-order stream on the critical path: 48 MMAs per tile).
    const bool leader = tc::elect_one();
    const uint32_t d_s = tc::smem_u32(Dp);
    const uint64_t dk = tc::smem_desc_sw128(d_s);                      // D' planes, K-major
    const uint64_t gk = tc::smem_desc_sw128(tc::smem_u32(Gp));         // -G planes, K-major
    const uint64_t wm = smem_desc_mn(tc::smem_u32(Wp), 1024);          // W_b planes read MN-major
    const uint32_t dk_lo = static_cast<uint32_t>(dk), k_hi = static_cast<uint32_t>(dk >> 32);
    const uint32_t gk_lo = static_cast<uint32_t>(gk);
    const uint32_t wm_lo = static_cast<uint32_t>(wm), m_hi = static_cast<uint32_t>(wm >> 32);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += grid, ++it) {
      const int b = it & 1;
      const uint32_t u = static_cast<uint32_t>(it) >> 1;
      mbar_wait_wd(&in_full[b], u & 1u);
      mbar_wait_wd(&acc_free[b], (u & 1u) ^ 1u);
      tc::fence_after_sync();
      if (leader) {
        const uint32_t acc_addr = taddr + b * 64;
        const uint32_t h_s = tc::smem_u32(Hp + b * kBufH);
        const uint32_t hk_lo = static_cast<uint32_t>(tc::smem_desc_sw128(h_s));          // H planes, K-major
        const uint32_t hm_lo = static_cast<uint32_t>(smem_desc_mn(h_s, 1024));           // H planes, MN-major (B of T)
        const uint32_t dm_lo = static_cast<uint32_t>(smem_desc_mn(d_s, h_s - d_s));      // [D' ; H] stacked, MN-major (A of T)
        const int pa[3] = {0, 1, 0}, pb[3] = {1, 0, 0};      // plane pairs, smallest products first
        {
          // dh = D' W_b: A = D' planes (K-major), B = the W_b planes read MN-major (N = c'' contiguous, K = c = rows)
          uint32_t acc = 0u;
#pragma unroll
          for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16) {
              tc::mma_bf16(acc_addr, tc::desc_at(dk_lo, k_hi, pa[q] * kPlA + k16 * 32), tc::desc_at(wm_lo, m_hi, pb[q] * kPlW + k16 * 2048),
                           idesc_kb, acc);
              acc = 1u;
            }
        }
        {
          // T += [D' ; H]^T H over the 128 tile rows (MN-major views; the leading-dimension offset stacks H behind D')
          uint32_t acc = it == 0 ? 0u : 1u;
#pragma unroll
          for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int k16 = 0; k16 < 8; ++k16) {
              tc::mma_bf16(taddr + 128, tc::desc_at(dm_lo, m_hi, pa[q] * kPlA + k16 * 2048), tc::desc_at(hm_lo, m_hi, pb[q] * kPlA + k16 * 2048),
                           idesc_mn, acc);
              acc = 1u;
            }
        }
        tc::mma_commit(d_free);                                                    // D' may be patched for the next tile
        {
          // dh -= H G
#pragma unroll
          for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16)
              tc::mma_bf16(acc_addr, tc::desc_at(hk_lo, k_hi, pa[q] * kPlA + k16 * 32), tc::desc_at(gk_lo, k_hi, pb[q] * kPlW + k16 * 32), idesc_k, 1u);
        }
        tc::mma_commit(&a_free[b]);
        tc::mma_commit(&acc_full[b]);
      }
      __syncwarp();
    }
  } else {
    // =========================================== epilogue + scatter ===================================================
    // warp w: TMEM lane quarter w & 3 (hardware rule), column half w >> 2; thread = one tile row x 32 columns
    const int quarter = warp & 3, half = warp >> 2;
    const int erow = quarter * 32 + lane;
    const int c4 = tid & 15, hw = lane >> 4;               // scatter: 16-byte piece c4 of the 8 rows r0 .. r0 + 7
    const int r0 = 16 * warp + 8 * hw;                     // (tile-independent, and so are the points they belong to)
    const int pp0 = r0 / k;
    const int nfirst = min(8, (pp0 + 1) * k - r0);         // rows of point pp0 among the 8; the others are point pp0 + 1 (k >= 8)
    unsigned char* dz = sm + EcBpSmem::dz_off;             // FP32 [128][64], 16-byte pieces XOR-swizzled by row & 7
    // neighbour indices of the thread's rows, fetched one tile ahead (a load consumed in the same tile stalled the warp)
    int jn[8];
    {
      const int nv0 = min(P, M - static_cast<int>(blockIdx.x) * P) * k;
      const size_t en = static_cast<size_t>(blockIdx.x) * P * k;
#pragma unroll
      for (int q = 0; q < 8; ++q) jn[q] = (r0 + q < nv0) ? p.idx[en + r0 + q] : 0;
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += grid, ++it) {
      const uint32_t b = it & 1, u = static_cast<uint32_t>(it) >> 1;
      const int gp0 = tile * P;
      const int nvalid = min(P, M - gp0) * k;
      int jc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) jc[q] = jn[q];
      if (tile + grid < p.num_tiles) {
        const int nvn = min(P, M - (tile + grid) * P) * k;
        const size_t en = static_cast<size_t>(tile + grid) * P * k;
#pragma unroll
        for (int q = 0; q < 8; ++q) jn[q] = (r0 + q < nvn) ? p.idx[en + r0 + q] : 0;
      }
      mbar_wait_wd(&acc_full[b], u & 1u);
      tc::fence_after_sync();
      {
        float v[32];
        const uint32_t sw = *reinterpret_cast<const uint32_t*>(Signs + (it & 3) * (128 * 8) + erow * 8 + half * 4);
        tc::tmem_ld32(taddr + (static_cast<uint32_t>(quarter * 32) << 16) + b * 64 + half * 32, v);
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&acc_free[b]);      // accumulator drained: the MMAs of tile + 2 may overwrite it
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 o;
          o.x = v[4 * i + 0] * ((sw >> (4 * i + 0)) & 1u ? 1.f : kSlope);
          o.y = v[4 * i + 1] * ((sw >> (4 * i + 1)) & 1u ? 1.f : kSlope);
          o.z = v[4 * i + 2] * ((sw >> (4 * i + 2)) & 1u ? 1.f : kSlope);
          o.w = v[4 * i + 3] * ((sw >> (4 * i + 3)) & 1u ? 1.f : kSlope);
          *reinterpret_cast<float4*>(dz + erow * 256 + (((half * 8 + i) ^ (erow & 7)) << 4)) = o;
        }
      }
      named_bar_sync(1, 256);                              // dz tile complete
      // R_j += dz rows (in-edge sums), two 256-byte rows per warp instruction; the sums of the thread's rows of point pp0 and
      // of point pp0 + 1 go to S (out-edge sums).  Straight-line: 8 loads, 8 + 2 reductions.
      {
        const int gpa = gp0 + pp0;
        const int basea = (gpa / p.N) * p.N, baseb = ((gpa + 1) / p.N) * p.N;
        float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int r = r0 + q;
          if (r < nvalid) {
            const float4 d = *reinterpret_cast<const float4*>(dz + r * 256 + ((c4 ^ (r & 7)) << 4));
            const bool fa = q < nfirst;
            red_add_v4(r_raw + static_cast<size_t>((fa ? basea : baseb) + jc[q]) * ld_rs + 4 * c4, d);
            sa.x += fa ? d.x : 0.f; sa.y += fa ? d.y : 0.f; sa.z += fa ? d.z : 0.f; sa.w += fa ? d.w : 0.f;
            sb.x += fa ? 0.f : d.x; sb.y += fa ? 0.f : d.y; sb.z += fa ? 0.f : d.z; sb.w += fa ? 0.f : d.w;
          }
        }
        if (r0 < nvalid) red_add_v4(s_raw + static_cast<size_t>(gpa) * ld_rs + 4 * c4, sa);
        if (nfirst < 8 && r0 + nfirst < nvalid) red_add_v4(s_raw + static_cast<size_t>(gpa + 1) * ld_rs + 4 * c4, sb);
      }
      named_bar_sync(1, 256);                              // dz tile consumed
    }
    // ---- T partial of this CTA (TMEM lanes 0-127, columns 128-191): rows 0-63 = D'^T H, rows 64-127 = H^T H -------------
    // (the last acc_full wait covers every MMA issued before its commit, the T MMAs of the last tile included)
    if (warp < 4 && it > 0) {
      float* dst = dw_partials + static_cast<size_t>(blockIdx.x) * 2 * kCout * kCout + static_cast<size_t>(warp * 32 + lane) * kCout;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        float v[32];
        tc::tmem_ld32(taddr + (static_cast<uint32_t>(warp * 32) << 16) + 128 + c0, v);
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == kBpMmaWarp) tc::tmem_dealloc(taddr, 256);
}

// ---- point-level finish of the BN_a backward --------------------------------------------------------------------------
// per-CTA FP64 partials of  sum_i S_i  and  sum_j U_j.R_j + sum_i V_i.S_i  (= sum dz_a and sum dz_a*y_a over the edges)
__global__ void __launch_bounds__(256)
ec_bwd_point_stats_kernel(const float* __restrict__ U, int ld_u, const float* __restrict__ V, int ld_v, const float* __restrict__ R,
                          const float* __restrict__ S, int ld_rs, int M, double* __restrict__ partials) {
  __shared__ float red[2][4][64];
  const int c = threadIdx.x & 63, sub = threadIdx.x >> 6;
  float t1 = 0.f, t2 = 0.f;
  for (int i = blockIdx.x * 4 + sub; i < M; i += gridDim.x * 4) {
    const float s = S[static_cast<size_t>(i) * ld_rs + c], r = R[static_cast<size_t>(i) * ld_rs + c];
    t1 += s;
    t2 = __fmaf_rn(U[static_cast<size_t>(i) * ld_u + c], r, __fmaf_rn(V[static_cast<size_t>(i) * ld_v + c], s, t2));
  }
  red[0][sub][c] = t1;
  red[1][sub][c] = t2;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int w = threadIdx.x >> 6;
    const double s = ((static_cast<double>(red[w][0][c]) + red[w][1][c]) + red[w][2][c]) + red[w][3][c];
    partials[128 + static_cast<size_t>(blockIdx.x) * 128 + threadIdx.x] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) partials[0] = static_cast<double>(gridDim.x);
}

// r_j = cA R_j - cnt_j cB - cC Yin_j,  s_i = cA S_i - k cB - cC Ysum_i  (in place)
__global__ void __launch_bounds__(256)
ec_bwd_point_apply_kernel(float* __restrict__ R, float* __restrict__ S, int ld_rs, const float* __restrict__ yin,
                          const float* __restrict__ ysum, const float* __restrict__ cnt, const float* __restrict__ coef_a, int M, float kf) {
  const int c = 4 * (threadIdx.x & 15);
  const float4 cA = *reinterpret_cast<const float4*>(coef_a + c);
  const float4 cB = *reinterpret_cast<const float4*>(coef_a + 64 + c);
  const float4 cC = *reinterpret_cast<const float4*>(coef_a + 128 + c);
  for (int i = blockIdx.x * 16 + (threadIdx.x >> 4); i < M; i += gridDim.x * 16) {
    float4* rp = reinterpret_cast<float4*>(R + static_cast<size_t>(i) * ld_rs + c);
    float4* sp = reinterpret_cast<float4*>(S + static_cast<size_t>(i) * ld_rs + c);
    const float4 yi = *reinterpret_cast<const float4*>(yin + static_cast<size_t>(i) * kCout + c);
    const float4 ys = *reinterpret_cast<const float4*>(ysum + static_cast<size_t>(i) * kCout + c);
    const float n = cnt[i];
    float4 r = *rp, s = *sp;
    r.x = cA.x * r.x - n * cB.x - cC.x * yi.x; r.y = cA.y * r.y - n * cB.y - cC.y * yi.y;
    r.z = cA.z * r.z - n * cB.z - cC.z * yi.z; r.w = cA.w * r.w - n * cB.w - cC.w * yi.w;
    s.x = cA.x * s.x - kf * cB.x - cC.x * ys.x; s.y = cA.y * s.y - kf * cB.y - cC.y * ys.y;
    s.z = cA.z * s.z - kf * cB.z - cC.z * ys.z; s.w = cA.w * s.w - kf * cB.w - cC.w * ys.w;
    *rp = r;
    *sp = s;
  }
}

int launch_ec_pipe_bwd_b(const EcK& p, const float* g, const uint8_t* arg, const float* coef_b, float* r_raw, float* s_raw, int ld_rs,
                         float* yin, float* cnt, float* dw_partials, float* dw_b, int sms, cudaStream_t st) {
  const int grid = std::max(1, std::min(p.num_tiles, std::min(std::max(1, sms), kMaxPartialCtas)));
  DG_CUDA(cudaFuncSetAttribute(ec_pipe_bwd_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(EcBpSmem::bytes)));
  ec_pipe_bwd_b_kernel<<<grid, kBpThreads, EcBpSmem::bytes, st>>>(p, g, arg, coef_b, r_raw, s_raw, ld_rs, dw_partials, yin, cnt);
  DG_CUDA(cudaGetLastError());
  return launch_ec_dw_finish(dw_partials, grid, p.w_b, coef_b, dw_b, st);
}

int launch_ec_bwd_point_stats(const float* U, int ld_u, const float* V, int ld_v, const float* R, const float* S, int ld_rs, int M,
                              double* partials, cudaStream_t st) {
  const int blocks = std::max(1, std::min((M + 3) / 4, std::min(kMaxPartialCtas, 4 * std::max(1, num_sms()))));
  ec_bwd_point_stats_kernel<<<blocks, 256, 0, st>>>(U, ld_u, V, ld_v, R, S, ld_rs, M, partials);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

int launch_ec_bwd_point_apply(float* R, float* S, int ld_rs, const float* yin, const float* ysum, const float* cnt, const float* coef_a,
                              int M, int k, cudaStream_t st) {
  const int blocks = std::max(1, std::min((M + 15) / 16, 8 * std::max(1, num_sms())));
  ec_bwd_point_apply_kernel<<<blocks, 256, 0, st>>>(R, S, ld_rs, yin, ysum, cnt, coef_a, M, static_cast<float>(k));
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

}  // namespace dg
