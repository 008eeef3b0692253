// knn_tc.cu — feature-space kNN (16 < C <= 64) on the tensor cores, with EXACT results.
//
// Gram tiles are computed by tcgen05.mma on a 2-plane bf16 split of the features (3 MMAs per product,
// |error| <= ~2e-5 * |x_i||x_j|) into TMEM; the approximate scores only act as a FILTER ("threshold - collect -
// re-score"), there are no running top-k lists and no CTA-wide barriers inside the key loop:
//   pass 1  scan the first ~N/200 key tiles; each of the 4 threads that share a query row keeps the 8 best
//           approximate scores of its 16 columns in registers; tau_i = mean of the four 8th-best values, an estimate
//           of the row's ~100-th best score overall;
//   pass 2  scan all keys; every key with approx > tau_i is appended to the row's candidate buffer (global memory,
//           4 private sub-buffers of 64 per row, no atomics);
//   pass 3  (warp per row) bisect the candidates' approximate scores to a value v exceeded by >= k+4 of them, keep the
//           candidates above v - 2 eps_i (eps_i = rigorous bound on |approx - exact| for the row), re-score those with
//           the reference's exact FP32 arithmetic (FMA chain in channel order, two rounded subtractions,
//           model.py:35-37) and emit the exact top-k;
//   certificate: every key outside the candidate set has exact score <= tau_i - |x_i|^2 + eps_i; if that is below the
//           k-th exact score found, the result is provably the exact top-k.  Rows that fail it (tau too tight, buffer
//           overflow, more than 64 near-ties) are redone by an exact brute-force kernel.
// The neighbour table is therefore identical to dgcnn_knn_f32's, bit for bit, and the B x N x N matrix never leaves
// the SM: scores go TMEM -> registers.
#include "common.cuh"
#include "tc.cuh"
#include "select.cuh"

namespace dg {

constexpr int kTcKeyTile = 64;          // keys per MMA tile (UMMA N)
constexpr uint32_t kTileBytes = 2 * tc::kUnitBytes + kTcKeyTile * 4;   // plane0 | plane1 | 64 negated key norms
constexpr uint32_t kStageBytes = 17 * 1024;                             // smem slot (planes stay 1024-aligned)
constexpr int kTcStages = 6;
// error model of the filter (see DESIGN.md "kNN filter bound")
constexpr float kC1 = 2.0e-5f * 2.0f;   // 2 * (split + accumulation + chain error) on 2*dot, relative to |x_i||x_j|
constexpr float kC2 = 4.8e-7f;          // final roundings, relative to (|x_i|+|x_j|)^2

// ---------------------------------------------------------------------------------------------
// prep: rows -> per-block tiles of 64 points: two swizzled bf16 planes + the 64 squared norms; also the squared
// norms themselves (torch.sum order, see knn.cu) and the per-block maximum norm.
// ---------------------------------------------------------------------------------------------
// per-block channel means (the filter works on centred features: distances are translation invariant and the error
// bound scales with the norms, which centring shrinks by 10-100x on post-LeakyReLU features)
__global__ void __launch_bounds__(256) knn_tc_mean_kernel(const float* __restrict__ rows, int ld, int N, int C4,
                                                          float* __restrict__ musum) {
  // grid (slices, B): each CTA sums a slice of the block's rows; musum (zero-initialised) receives the channel SUMS
  __shared__ float red[4][64];
  const int b = blockIdx.y, c = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int per = (N + gridDim.x - 1) / gridDim.x, n0 = blockIdx.x * per, n1 = min(N, n0 + per);
  float s = 0.f;
  if (c < C4)
    for (int n = n0 + grp; n < n1; n += 4) s += rows[(static_cast<size_t>(b) * N + n) * ld + c];
  red[grp][c] = s;
  __syncthreads();
  if (threadIdx.x < 64) atomicAdd(musum + b * 64 + c, red[0][c] + red[1][c] + red[2][c] + red[3][c]);
}

__global__ void knn_tc_prep_kernel(const float* __restrict__ rows, int ld, int N, int C4, int T,
                                   const float* __restrict__ mu, unsigned char* __restrict__ tiles,
                                   float* __restrict__ normc, unsigned int* __restrict__ maxnormc_bits) {
  // one thread per (tile row, 16-byte chunk); grid.y = block.  total is a multiple of 8, so the 8 chunk-threads of a
  // row are always in the same warp and all active.
  const int b = blockIdx.y;
  const int total = T * 64 * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ch = i & 7, rl = (i >> 3) & 63, t = i >> 9;
    const int n = t * 64 + rl;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (n < N && ch * 8 < C4) {
      const float* src = rows + (static_cast<size_t>(b) * N + n) * ld + ch * 8;
      const float* m = mu + b * 64 + ch * 8;       // channel sums; any value near the mean is a valid centre
      const float inv_n = 1.0f / static_cast<float>(N);
      const float4 lo = *reinterpret_cast<const float4*>(src);
      v[0] = lo.x - m[0] * inv_n; v[1] = lo.y - m[1] * inv_n; v[2] = lo.z - m[2] * inv_n; v[3] = lo.w - m[3] * inv_n;
      if (ch * 8 + 4 < C4) {
        const float4 hi = *reinterpret_cast<const float4*>(src + 4);
        v[4] = hi.x - m[4] * inv_n; v[5] = hi.y - m[5] * inv_n; v[6] = hi.z - m[6] * inv_n; v[7] = hi.w - m[7] * inv_n;
      }
    }
    float sq = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) sq = __fmaf_rn(v[e], v[e], sq);
    sq += __shfl_xor_sync(0xffffffffu, sq, 1);
    sq += __shfl_xor_sync(0xffffffffu, sq, 2);
    sq += __shfl_xor_sync(0xffffffffu, sq, 4);
    unsigned char* tile = tiles + (static_cast<size_t>(b) * T + t) * kTileBytes;
    const uint32_t off = tc::unit_elem_offset(rl, ch * 8);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      __nv_bfloat16 h[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        h[e] = __float2bfloat16_rn(v[e]);
        v[e] -= __bfloat162float(h[e]);
      }
      *reinterpret_cast<uint4*>(tile + p * tc::kUnitBytes + off) = *reinterpret_cast<const uint4*>(h);
    }
    if (ch == 0) {
      // MINUS the centred squared norm; -inf for padding keys: their score 2*0 - inf never passes a threshold, so the
      // scans need no per-key bounds check
      reinterpret_cast<float*>(tile + 2 * tc::kUnitBytes)[rl] = n < N ? -sq : -INFINITY;
      if (n < N) {
        normc[static_cast<size_t>(b) * N + n] = sq;
        atomicMax(maxnormc_bits + b, __float_as_uint(sq));
      }
    }
  }
}

__global__ void knn_norms_max_kernel(const float* __restrict__ rows, int ld, int N, int C, float* __restrict__ norms,
                                     unsigned int* __restrict__ maxnorm_bits) {
  // same summation order as norms_kernel in knn.cu; per-block max via atomicMax on the (non-negative) float bits
  const int b = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  float val = 0.f;
  if (n < N) {
    const float* r = rows + (static_cast<size_t>(b) * N + n) * ld;
    float lvl1 = 0.f;
    int c = 0;
    for (; c + 16 <= C; c += 16) {
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) a = __fadd_rn(a, __fmul_rn(r[c + j], r[c + j]));
      lvl1 = __fadd_rn(lvl1, a);
    }
    float a = 0.f;
    for (; c < C; ++c) a = __fadd_rn(a, __fmul_rn(r[c], r[c]));
    val = __fadd_rn(a, lvl1);
    norms[static_cast<size_t>(b) * N + n] = val;
  }
  unsigned int m = __float_as_uint(val);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(maxnorm_bits + b, m);
}

// exact reference score of (query row xi, key row xj): FMA chain over channels, then the two rounded subtractions
// The key row is a gather from L2: all of its (<= 16) 16-byte loads are issued before the first use — with a rolled
// loop every 4 channels paid a full L2 round trip (ncu: the re-scoring pass was 36 % of the kernel at 0.09 IPC).
__device__ __forceinline__ float exact_score(const float* __restrict__ xi, const float* __restrict__ xj, int C4,
                                             float norm_i, float norm_j) {
  float4 b[16];
  if (C4 == 64 && (reinterpret_cast<uintptr_t>(xj) & 31) == 0) {
    // one lane reads one whole key row: 32-byte loads fetch every sector once (16-byte loads request each sector twice,
    // and with 16 warps x 32 rows in flight the second request usually misses L1 again)
#pragma unroll
    for (int q = 0; q < 16; q += 2)
      asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=f"(b[q].x), "=f"(b[q].y), "=f"(b[q].z), "=f"(b[q].w), "=f"(b[q + 1].x), "=f"(b[q + 1].y), "=f"(b[q + 1].z),
                     "=f"(b[q + 1].w)
                   : "l"(xj + 4 * q));
  } else {
#pragma unroll
    for (int q = 0; q < 16; ++q)
      b[q] = 4 * q < C4 ? *reinterpret_cast<const float4*>(xj + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float d = 0.f;
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    if (4 * q < C4) {                 // C4 <= 64 on this path; the chain order is the reference's (channel by channel)
      const float4 a = *reinterpret_cast<const float4*>(xi + 4 * q);
      d = __fmaf_rn(a.x, b[q].x, d);
      d = __fmaf_rn(a.y, b[q].y, d);
      d = __fmaf_rn(a.z, b[q].z, d);
      d = __fmaf_rn(a.w, b[q].w, d);
    }
  }
  return __fsub_rn(__fmaf_rn(d, 2.0f, -norm_j), norm_i);
}

constexpr int kAccs = 4;                 // TMEM accumulators (64 columns each)
constexpr int kScanWarps = 16;
constexpr int kTcThreads = (kScanWarps + 1) * 32;   // + 1 producer / MMA-issue warp
constexpr int kSubCands = 64;            // candidate slots per (row, column quarter)
constexpr int kSurvivors = 64;           // exact re-scores per row (2 per lane)

struct KnnTcSmem {
  static constexpr uint32_t a_bytes = 2 * 128 * 128;                        // two planes of 128 query rows
  static constexpr uint32_t b_off = a_bytes;
  static constexpr uint32_t misc_off = b_off + kTcStages * kStageBytes;     // per-row scalars
  static constexpr uint32_t stage_off = misc_off + 128 * (4 * 4 + 4 * 4 + 4 + 4);   // tau4 | cnt4 | tau | flag
  static constexpr uint32_t bar_off = stage_off + kScanWarps * kSurvivors * 8;      // per-warp survivor staging
  static constexpr uint32_t bytes = bar_off + 256;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// insert v into the descending top-8 t[0..7] (registers)
__device__ __forceinline__ void top8_insert(float (&t)[8], float v) {
#pragma unroll
  for (int i = 7; i >= 1; --i) t[i] = v > t[i - 1] ? t[i - 1] : fmaxf(v, t[i]);
  t[0] = fmaxf(v, t[0]);
}

__global__ void __launch_bounds__(kTcThreads, 1)
knn_tc_kernel(const unsigned char* __restrict__ tiles, const float* __restrict__ rows, int ld,
              const float* __restrict__ norms, const unsigned int* __restrict__ maxnorm_bits,
              const float* __restrict__ normc, const unsigned int* __restrict__ maxnormc_bits, int N, int C4, int T,
              int k, int T1, unsigned long long* __restrict__ cand, int32_t* __restrict__ idx,
              int* __restrict__ fb_count, int* __restrict__ fb_rows) {
  // no static shared memory in this kernel: the dynamic window starts 1024-byte aligned (checked below), and plain
  // pointer arithmetic on it keeps the shared address space visible to the compiler (LDS/STS, not generic LD/ST)
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* As = sm;
  unsigned char* Bs = sm + KnnTcSmem::b_off;
  float* tau4 = reinterpret_cast<float*>(sm + KnnTcSmem::misc_off);      // [128][4]
  int* cnt4 = reinterpret_cast<int*>(tau4 + 128 * 4);                    // [128][4]
  float* tauS = reinterpret_cast<float*>(cnt4 + 128 * 4);                // [128]
  int* flagS = reinterpret_cast<int*>(tauS + 128);                       // [128] overflow of a sub-buffer
  unsigned long long* stageS = reinterpret_cast<unsigned long long*>(sm + KnnTcSmem::stage_off);
  uint64_t* abar = reinterpret_cast<uint64_t*>(sm + KnnTcSmem::bar_off);
  uint64_t* full = abar + 1;                 // [kTcStages]  tile landed in smem
  uint64_t* sfree = full + kTcStages;        // [kTcStages]  all scan warps are done with the slot
  uint64_t* accfull = sfree + kTcStages;     // [kAccs]      MMA of the tile finished
  uint64_t* accfree = accfull + kAccs;       // [kAccs]      all scan warps have read the accumulator
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfree + kAccs);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y, qt = blockIdx.x, q0 = qt * 128;
  const int Tk = (N + kTcKeyTile - 1) / kTcKeyTile;        // key tiles of the block
  const int TT = T1 + Tk;                                  // tiles streamed: pass 1 (first T1 tiles) then pass 2 (all)
  const unsigned char* btiles = tiles + static_cast<size_t>(b) * T * kTileBytes;

  if (tid == 0) {
    if (tc::smem_u32(sm) & 1023u) __trap();
    tc::mbar_init(abar, 1);
    for (int s = 0; s < kTcStages; ++s) {
      tc::mbar_init(full + s, 1);
      tc::mbar_init(sfree + s, kScanWarps);
    }
    for (int a = 0; a < kAccs; ++a) {
      tc::mbar_init(accfull + a, 1);
      tc::mbar_init(accfree + a, kScanWarps);
    }
    tc::fence_barrier_init();
  }
  if (warp == kScanWarps) tc::tmem_alloc(tmem_slot, kAccs * kTcKeyTile);
  if (tid < 128) flagS[tid] = 0;
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t taddr = *tmem_slot;

  if (warp == kScanWarps) {
    // ===== producer + MMA issuer (one thread) =================================================================
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::idesc_bf16(128, kTcKeyTile);
      tc::mbar_arrive_expect_tx(abar, KnnTcSmem::a_bytes);
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int p = 0; p < 2; ++p)
          tc::bulk_g2s(As + p * 16384 + h * tc::kUnitBytes,
                       btiles + static_cast<size_t>(2 * qt + h) * kTileBytes + p * tc::kUnitBytes, tc::kUnitBytes, abar);
      int loaded = 0;
      auto issue_copy = [&](int l) {
        const int s = l % kTcStages;
        const int tile = l < T1 ? l : l - T1;
        tc::mbar_arrive_expect_tx(full + s, kTileBytes);
        tc::bulk_g2s(Bs + s * kStageBytes, btiles + static_cast<size_t>(tile) * kTileBytes, kTileBytes, full + s);
      };
      tc::mbar_wait(abar, 0);
      for (int i = 0; i < TT; ++i) {
        // tile i must be on its way; further tiles are prefetched only if their slot is already free (never block
        // the MMA issue on a slot that the scan warps still hold)
        while (loaded < TT && loaded < i + kTcStages) {
          uint64_t* bar = sfree + loaded % kTcStages;
          const uint32_t par = ((loaded / kTcStages) & 1) ^ 1;
          if (loaded <= i) tc::mbar_wait(bar, par);
          else if (!tc::mbar_try_wait(bar, par)) break;
          issue_copy(loaded++);
        }
        const int s = i % kTcStages, a = i % kAccs;
        tc::mbar_wait(full + s, (i / kTcStages) & 1);
        tc::mbar_wait(accfree + a, ((i / kAccs) & 1) ^ 1);
        tc::fence_after_sync();
        tc::mma_split_katom<2>(taddr + a * kTcKeyTile, tc::smem_u32(As), 16384, tc::smem_u32(Bs + s * kStageBytes),
                               tc::kUnitBytes, idesc, true);
        tc::mma_commit(accfull + a);
      }
    }
  } else {
    // ===== scan warps: TMEM lane quarter (warp & 3), column quarter (warp >> 2); thread = 1 query row x 16 keys ====
    const int quarter = warp & 3, cq = warp >> 2;
    const int erow = quarter * 32 + lane;
    const bool row_ok = q0 + erow < N;
    const uint32_t tlane = static_cast<uint32_t>(quarter * 32) << 16;
    float top[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) top[i] = -INFINITY;

    // ---- pass 1: threshold estimate from the first T1 tiles ------------------------------------------------------
    for (int i = 0; i < T1; ++i) {
      const int s = i % kTcStages, a = i % kAccs;
      tc::mbar_wait(accfull + a, (i / kAccs) & 1);
      tc::fence_after_sync();
      float v[16];
      tc::tmem_ld16(taddr + tlane + a * kTcKeyTile + cq * 16, v);
      const float* kn = reinterpret_cast<const float*>(Bs + s * kStageBytes + 2 * tc::kUnitBytes) + cq * 16;
#pragma unroll
      for (int c = 0; c < 16; c += 2) {
        // two keys per FFMA2 (same roundings as the scalar form); padding keys score -inf
        const float2 kk = *reinterpret_cast<const float2*>(kn + c);
        const float2 sc2 = __ffma2_rn(make_float2(v[c], v[c + 1]), make_float2(2.0f, 2.0f), kk);
        if (sc2.x > top[7]) top8_insert(top, sc2.x);
        if (sc2.y > top[7]) top8_insert(top, sc2.y);
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(accfree + a);
        tc::mbar_arrive(sfree + s);
      }
    }
    tau4[erow * 4 + cq] = top[7];
    named_bar_sync(1, kScanWarps * 32);
    const float4 t4 = *reinterpret_cast<const float4*>(tau4 + erow * 4);
    const float tau = 0.25f * (t4.x + t4.y + t4.z + t4.w);     // -inf if a thread saw fewer than 8 keys: keep all
    if (cq == 0) tauS[erow] = tau;

    // ---- pass 2: collect every key above tau ------------------------------------------------------------------------
    unsigned long long* mycand = cand + ((static_cast<size_t>(b) * N + q0 + erow) * 4 + cq) * kSubCands;
    int n = 0;
    const float tau_eff = row_ok ? tau : INFINITY;       // rows past the end of the block collect nothing
    for (int i = T1; i < TT; ++i) {
      const int s = i % kTcStages, a = i % kAccs;
      tc::mbar_wait(accfull + a, (i / kAccs) & 1);
      tc::fence_after_sync();
      float v[16];
      tc::tmem_ld16(taddr + tlane + a * kTcKeyTile + cq * 16, v);
      const float* kn = reinterpret_cast<const float*>(Bs + s * kStageBytes + 2 * tc::kUnitBytes) + cq * 16;
      const int j0 = (i - T1) * kTcKeyTile + cq * 16;
#pragma unroll
      for (int c = 0; c < 16; c += 2) {
        // approx of 2 x_i.x_j - |x_j|^2 (|x_i|^2 is constant per row), two keys per FFMA2.  The warp-uniform vote keeps
        // the common case (no lane has a candidate) free of divergence bookkeeping.
        const float2 kk = *reinterpret_cast<const float2*>(kn + c);
        const float2 sc2 = __ffma2_rn(make_float2(v[c], v[c + 1]), make_float2(2.0f, 2.0f), kk);
        const bool h0 = sc2.x > tau_eff, h1 = sc2.y > tau_eff;
        if (__any_sync(0xffffffffu, h0 || h1)) {
          if (h0) {
            if (n < kSubCands) mycand[n] = make_key(sc2.x, j0 + c);
            ++n;
          }
          if (h1) {
            if (n < kSubCands) mycand[n] = make_key(sc2.y, j0 + c + 1);
            ++n;
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(accfree + a);
        tc::mbar_arrive(sfree + s);
      }
    }
    cnt4[erow * 4 + cq] = n;
    __threadfence_block();
    named_bar_sync(1, kScanWarps * 32);

    // ---- pass 3: bisect, re-score exactly, certify (warp per row, 8 rows per warp) ----------------------------------
    const float rmax = sqrtf(__uint_as_float(maxnormc_bits[b]));      // largest CENTRED norm of the block
    const float rx = sqrtf(__uint_as_float(maxnorm_bits[b]));         // largest raw norm (rounding of the centring)
    unsigned long long* stg = stageS + warp * kSurvivors;
    for (int rr = 0; rr < 8; ++rr) {
      const int row = warp * 8 + rr, q = q0 + row;
      if (q >= N) break;
      const size_t g = static_cast<size_t>(b) * N + q;
      const int4 c4 = *reinterpret_cast<const int4*>(cnt4 + row * 4);
      bool bad = max(max(c4.x, c4.y), max(c4.z, c4.w)) > kSubCands;          // a sub-buffer overflowed
      const int cs[4] = {min(c4.x, kSubCands), min(c4.y, kSubCands), min(c4.z, kSubCands), min(c4.w, kSubCands)};
      const int ntot = cs[0] + cs[1] + cs[2] + cs[3];
      bad = bad || ntot < k;                                                  // tau was too tight
      const unsigned long long* rc = cand + g * 4 * kSubCands;
      unsigned long long key[8];
      float ap[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int sub = r >> 1, slot = (r & 1) * 32 + lane;
        key[r] = slot < cs[sub] ? rc[sub * kSubCands + slot] : 0ull;
        ap[r] = key[r] != 0ull ? sel::key_score(key[r]) : -INFINITY;
      }
      const float tau_r = tauS[row];
      // bisection: v with at least min(ntot, k) candidates strictly above it
      float lo = tau_r, hi = ap[0];
#pragma unroll
      for (int r = 1; r < 8; ++r) hi = fmaxf(hi, ap[r]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      if (!(lo > -INFINITY)) {                       // tau = -inf (tiny blocks): start from the smallest candidate
        float mn = INFINITY;
#pragma unroll
        for (int r = 0; r < 8; ++r) mn = fminf(mn, key[r] != 0ull ? ap[r] : INFINITY);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        lo = mn - fabsf(mn) * 1e-6f - 1e-30f;
      }
      // k candidates above the cut suffice: with |approx - exact - const| <= eps, a key below lo - 2 eps has k keys
      // strictly above it in exact score.  Every extra survivor costs a 256-byte gather.
      const int want = min(ntot, k);
      for (int it = 0; it < 10; ++it) {
        const float mid = 0.5f * (lo + hi);
        int c = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) c += ap[r] > mid;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (c >= want) lo = mid; else hi = mid;
        if (c >= want && c <= want + 3) break;               // close enough: every extra survivor is a 256-byte gather
      }
      const float norm_i = norms[g];                 // reference |x_i|^2 (exact scores)
      const float nci = normc[g];                    // centred |y_i|^2: approx score of the row = exact pd + nci
      const float ni = sqrtf(nci);
      // Only keys with |y_j| <= |y_i| + d_tau can have a score near tau (score <= |y_i|^2 - (|y_j| - |y_i|)^2), so the
      // error bound uses R_i = min(R, |y_i| + 1.1 d_tau) instead of the block maximum R: one far outlier does not
      // inflate every row's margin.  Third term: rounding of y = fl(x - mu) (2^-22 Rx (|y_i| + 2 R_i)).
      const float dtau = sqrtf(fmaxf(nci - tau_r, 0.f));
      const float ri = tau_r > -INFINITY ? fminf(rmax, ni + 1.1f * dtau + 1e-3f * ni) : rmax;
      const float eps = 1.01f * (kC1 * ni * ri + kC2 * (ni + ri) * (ni + ri) + 2.4e-7f * rx * (ni + 2.0f * ri));
      const float cut = lo - 2.0f * eps;
      // compact the survivors (approx > cut) into the warp's staging area
      int base = 0;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const bool keep = ap[r] > cut;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        if (keep && pos < kSurvivors) stg[pos] = key[r];
        base += __popc(m);
      }
      bad = bad || base > kSurvivors;                                         // too many near-ties to re-score here
      const int ns = min(base, kSurvivors);
      __syncwarp();
      const float* xi = rows + g * ld;
      unsigned long long ex[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        ex[r] = 0ull;
        const int e = lane + 32 * r;
        if (e < ns) {
          const int j = key_index(stg[e]);
          const size_t gj = static_cast<size_t>(b) * N + j;
          ex[r] = make_key(exact_score(xi, rows + gj * ld, C4, norm_i, norms[gj]), j);
        }
      }
      __syncwarp();
      // bitonic sort of the (<= 64, distinct) exact keys, best first: element lane in ex[0], lane + 32 in ex[1]
      sel::sort64_desc(ex[0], ex[1], lane);
      if (ex[0] != 0ull && lane < k) idx[g * k + lane] = key_index(ex[0]);
      if (ex[1] != 0ull && lane + 32 < k) idx[g * k + lane + 32] = key_index(ex[1]);
      // certificate: no key outside the candidate set can reach the k-th exact score found
      const unsigned long long kth = sel::list_at(ex[0], ex[1], k - 1);
      const float pdk = kth != 0ull ? sel::key_score(kth) : -INFINITY;
      const bool cert = (tau_r - nci) + eps < pdk;
      if (lane == 0) {     // diagnostics: fb_count[1..4] count the reasons (overflow, too few, too many near-ties, certificate)
        if (max(max(c4.x, c4.y), max(c4.z, c4.w)) > kSubCands) atomicAdd(fb_count + 1, 1);
        else if (ntot < k) atomicAdd(fb_count + 2, 1);
        else if (base > kSurvivors) atomicAdd(fb_count + 3, 1);
        else if (!cert) atomicAdd(fb_count + 4, 1);
      }
      bad = bad || !cert;
      if (bad && lane == 0) fb_rows[atomicAdd(fb_count, 1)] = static_cast<int>(g);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == kScanWarps) tc::tmem_dealloc(taddr, kAccs * kTcKeyTile);
}

// warp-cooperative sorted insertion used by the brute-force fallback (2 keys per lane, position = r*32 + lane)
__device__ __forceinline__ unsigned long long list2_at(const unsigned long long (&lst)[2], int pos) {
  return sel::list_at(lst[0], lst[1], pos);
}
__device__ __forceinline__ void list2_insert(unsigned long long (&lst)[2], unsigned long long c, int lane) {
  const unsigned long long up0 = __shfl_up_sync(0xffffffffu, lst[0], 1);
  const unsigned long long up1 = __shfl_up_sync(0xffffffffu, lst[1], 1);
  const unsigned long long wrap = __shfl_sync(0xffffffffu, lst[0], 31);
  const unsigned long long p0 = lane == 0 ? ~0ull : up0;
  const unsigned long long p1 = lane == 0 ? wrap : up1;
  lst[0] = lst[0] >= c ? lst[0] : (p0 >= c ? c : p0);
  lst[1] = lst[1] >= c ? lst[1] : (p1 >= c ? c : p1);
}

// ---------------------------------------------------------------------------------------------
// exact brute force for the (rare) rows whose certificate failed: one CTA per flagged row
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
knn_fallback_kernel(const float* __restrict__ rows, int ld, const float* __restrict__ norms, int N, int C4, int k,
                    int32_t* __restrict__ idx, const int* __restrict__ fb_count, const int* __restrict__ fb_rows) {
  __shared__ __align__(16) float xi[64];
  __shared__ unsigned long long wl[8][64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int count = *fb_count;
  for (int f = blockIdx.x; f < count; f += gridDim.x) {
    const size_t g = static_cast<size_t>(fb_rows[f]);
    const size_t base = (g / N) * N;
    __syncthreads();
    if (tid < 64) xi[tid] = tid < C4 ? rows[g * ld + tid] : 0.f;
    __syncthreads();
    const float norm_i = norms[g];
    unsigned long long lst[2] = {0ull, 0ull};
    unsigned long long thr = 0ull;
    for (int j0 = warp * 32; j0 < N; j0 += 256) {
      const int j = j0 + lane;
      unsigned long long c = 0ull;
      if (j < N) c = make_key(exact_score(xi, rows + (base + j) * ld, C4, norm_i, norms[base + j]), j);
      unsigned m = __ballot_sync(0xffffffffu, c > thr);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const unsigned long long cc = __shfl_sync(0xffffffffu, c, src);
        if (cc > thr) {
          list2_insert(lst, cc, lane);
          thr = list2_at(lst, k - 1);
        }
      }
    }
    wl[warp][lane] = lst[0];
    wl[warp][lane + 32] = lst[1];
    __syncthreads();
    if (warp == 0) {
      for (int w = 1; w < 8; ++w)
        for (int e = 0; e < k; ++e) {
          const unsigned long long cc = wl[w][e];
          if (cc > thr) {
            list2_insert(lst, cc, lane);
            thr = list2_at(lst, k - 1);
          }
        }
      if (lane < k) idx[g * k + lane] = key_index(lst[0]);
      if (lane + 32 < k) idx[g * k + lane + 32] = key_index(lst[1]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// low-dimensional graphs (C <= 16: the xyz/agl input graph): same threshold - collect - select scheme with EXACT
// scores from an FP32 FMA chain (thread per query, coordinates in registers, keys broadcast from shared memory).
// No certificate is needed: every key above tau is a candidate, so the exact top-k is among them iff >= k were found.
// ---------------------------------------------------------------------------------------------
constexpr int kSmallKeys = 512;        // keys per shared-memory tile
constexpr int kSmallRows = 64;         // query rows per CTA; 4 threads per row, each scanning a quarter of every key tile

template <int CP>
__global__ void __launch_bounds__(kThreads)
knn_small_kernel(const float* __restrict__ rows, int ld, const float* __restrict__ norms, int N, int C4, int k,
                 int n_sample, unsigned long long* __restrict__ cand, int32_t* __restrict__ idx,
                 int* __restrict__ fb_count, int* __restrict__ fb_rows) {
  // keys in pairs: Ks[(pair * CP + c) * 2 + slot] so that one FFMA2 (fma.rn.f32x2) scores two keys; Kn holds -|x_j|^2
  __shared__ __align__(16) float Ks[kSmallKeys * CP];
  __shared__ __align__(8) float Kn[kSmallKeys];
  __shared__ __align__(16) int cntS[kSmallRows * 4];
  __shared__ __align__(16) float tau4[kSmallRows * 4];
  __shared__ unsigned long long stageS[8][kSurvivors];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row = tid & (kSmallRows - 1), kq = tid >> 6;          // a warp = 32 rows of one key quarter
  const int b = blockIdx.y, q0 = blockIdx.x * kSmallRows, q = q0 + row;
  const size_t base = static_cast<size_t>(b) * N;
  const bool row_ok = q < N;
  float xq[CP];
#pragma unroll
  for (int c = 0; c < CP; c += 4) {
    const float4 v = (row_ok && c < C4) ? *reinterpret_cast<const float4*>(rows + (base + q) * ld + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    xq[c] = v.x; xq[c + 1] = v.y; xq[c + 2] = v.z; xq[c + 3] = v.w;
  }
  const float nq = row_ok ? norms[base + q] : 0.f;
  unsigned long long* mycand = cand + ((base + q) * 4 + kq) * kSubCands;

  float top[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) top[i] = -INFINITY;
  float tau = -INFINITY;
  int n = 0;
  // pass 0: the first n_sample keys give tau (each thread's 8th best of its quarter ~ global rank 32 N / n_sample, the
  // four estimates are averaged); pass 1: all keys, every key above tau becomes a candidate
  for (int pass = 0; pass < 2; ++pass) {
    const int nkeys = pass == 0 ? n_sample : N;
    for (int t0 = 0; t0 < nkeys; t0 += kSmallKeys) {
      __syncthreads();
      for (int i = tid; i < kSmallKeys * (CP / 4); i += kThreads) {
        const int r = i / (CP / 4), c4 = i % (CP / 4), j = t0 + r;
        const float4 v = (j < nkeys && 4 * c4 < C4) ? *reinterpret_cast<const float4*>(rows + (base + j) * ld + 4 * c4)
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
        float* dst = Ks + ((r >> 1) * CP + 4 * c4) * 2 + (r & 1);
        dst[0] = v.x; dst[2] = v.y; dst[4] = v.z; dst[6] = v.w;
      }
      // keys past the end score -inf: no per-key bounds check in the scan
      for (int r = tid; r < kSmallKeys; r += kThreads) Kn[r] = t0 + r < nkeys ? -norms[base + t0 + r] : -INFINITY;
      __syncthreads();
      const int r0 = kq * (kSmallKeys / 4);
      const float2 two = make_float2(2.0f, 2.0f), mnq = make_float2(-nq, -nq);
      const float thr = pass == 0 ? 0.f : (row_ok ? tau : INFINITY);       // rows past the end of the block collect nothing
#pragma unroll 2
      for (int i = 0; i < kSmallKeys / 4; i += 2) {
        const int r = r0 + i;
        float2 d = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < CP; c += 2) {
          const float4 kv = *reinterpret_cast<const float4*>(Ks + ((r >> 1) * CP + c) * 2);      // channels c, c+1 of both keys
          d = __ffma2_rn(make_float2(xq[c], xq[c]), make_float2(kv.x, kv.y), d);
          d = __ffma2_rn(make_float2(xq[c + 1], xq[c + 1]), make_float2(kv.z, kv.w), d);
        }
        // model.py:37, one rounding per step: fl(fl(2 d - |x_j|^2) - |x_i|^2); the FMA chain order is the scalar one
        const float2 pd2 = __fadd2_rn(__ffma2_rn(d, two, *reinterpret_cast<const float2*>(Kn + r)), mnq);
        if (pass == 0) {
          if (pd2.x > top[7]) top8_insert(top, pd2.x);
          if (pd2.y > top[7]) top8_insert(top, pd2.y);
        } else {
          // the warp-uniform vote keeps the common case (no lane has a candidate) free of divergence bookkeeping
          const bool h0 = pd2.x > thr, h1 = pd2.y > thr;
          if (__any_sync(0xffffffffu, h0 || h1)) {
            if (h0) {
              if (n < kSubCands) mycand[n] = make_key(pd2.x, t0 + r);
              ++n;
            }
            if (h1) {
              if (n < kSubCands) mycand[n] = make_key(pd2.y, t0 + r + 1);
              ++n;
            }
          }
        }
      }
    }
    if (pass == 0) {
      tau4[row * 4 + kq] = top[7];
      __syncthreads();
      const float4 t4 = *reinterpret_cast<const float4*>(tau4 + row * 4);
      tau = 0.25f * (t4.x + t4.y + t4.z + t4.w);        // -inf when a thread saw fewer than 8 keys: keep everything
    }
  }
  cntS[row * 4 + kq] = n;
  __threadfence_block();
  __syncthreads();

  // selection: warp per query row (8 rows per warp); scores are exact, so the top-k of the candidates is the answer
  // iff at least k candidates exist
  unsigned long long* stg = stageS[warp];
  for (int rr = 0; rr < kSmallRows / 8; ++rr) {
    const int srow = warp * (kSmallRows / 8) + rr, qq = q0 + srow;
    if (qq >= N) break;
    const size_t g = base + qq;
    const int4 c4 = *reinterpret_cast<const int4*>(cntS + srow * 4);
    bool bad = max(max(c4.x, c4.y), max(c4.z, c4.w)) > kSubCands;
    const int cs[4] = {min(c4.x, kSubCands), min(c4.y, kSubCands), min(c4.z, kSubCands), min(c4.w, kSubCands)};
    const int ntot = cs[0] + cs[1] + cs[2] + cs[3];
    bad = bad || ntot < k;
    const unsigned long long* rc = cand + g * 4 * kSubCands;
    unsigned long long key[8];
    float ap[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int sub = r >> 1, slot = (r & 1) * 32 + lane;
      key[r] = slot < cs[sub] ? rc[sub * kSubCands + slot] : 0ull;
      ap[r] = key[r] != 0ull ? sel::key_score(key[r]) : -INFINITY;
    }
    const float4 t4 = *reinterpret_cast<const float4*>(tau4 + srow * 4);
    float lo = 0.25f * (t4.x + t4.y + t4.z + t4.w), hi = ap[0];
#pragma unroll
    for (int r = 1; r < 8; ++r) hi = fmaxf(hi, ap[r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    if (!(lo > -INFINITY)) {
      float mn = INFINITY;
#pragma unroll
      for (int r = 0; r < 8; ++r) mn = fminf(mn, key[r] != 0ull ? ap[r] : INFINITY);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      lo = mn - fabsf(mn) * 1e-6f - 1e-30f;
    }
    const int want = min(ntot, k + 4);
    for (int it = 0; it < 12; ++it) {            // v = lo: at least `want` candidates strictly above it
      const float mid = 0.5f * (lo + hi);
      int c = 0;
#pragma unroll
      for (int r = 0; r < 8; ++r) c += ap[r] > mid;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      if (c >= want) lo = mid; else hi = mid;
      if (c >= want && c <= kSurvivors) break;             // any cut that keeps [k+4, 64] candidates will do
    }
    int nb = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const bool keep = ap[r] > lo;
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      const int pos = nb + __popc(m & ((1u << lane) - 1u));
      if (keep && pos < kSurvivors) stg[pos] = key[r];
      nb += __popc(m);
    }
    bad = bad || nb > kSurvivors;                 // > 64 exact ties around the k-th score
    const int ns = min(nb, kSurvivors);
    __syncwarp();
    unsigned long long e0 = lane < ns ? stg[lane] : 0ull;
    unsigned long long e1 = lane + 32 < ns ? stg[lane + 32] : 0ull;
    __syncwarp();
    sel::sort64_desc(e0, e1, lane);                         // distinct keys, best first
    if (e0 != 0ull && lane < k) idx[g * k + lane] = key_index(e0);
    if (e1 != 0ull && lane + 32 < k) idx[g * k + lane + 32] = key_index(e1);
    if (bad && lane == 0) fb_rows[atomicAdd(fb_count, 1)] = static_cast<int>(g);
  }
}

template <int CP>
static int launch_knn_small(const float* rows, int ld, const float* norms, int B, int N, int C4, int k, int n_sample,
                            unsigned long long* cand, int32_t* idx, int* fb_count, int* fb_rows, cudaStream_t st) {
  knn_small_kernel<CP><<<dim3((N + kSmallRows - 1) / kSmallRows, B), kThreads, 0, st>>>(rows, ld, norms, N, C4, k, n_sample, cand,
                                                                                       idx, fb_count, fb_rows);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

}  // namespace dg

using namespace dg;

static size_t knn_tc_layout(int B, int N, size_t* o_norms, size_t* o_max, size_t* o_cnt, size_t* o_rows, size_t* o_tiles,
                            size_t* o_cand, size_t* o_normc = nullptr, size_t* o_mu = nullptr, size_t* o_maxc = nullptr) {
  const int T = 2 * ((N + 127) / 128);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 1023) & ~size_t(1023); return o; };
  *o_norms = take(static_cast<size_t>(B) * N * 4);
  *o_max = take(static_cast<size_t>(B) * 4);
  *o_cnt = take(4);
  *o_rows = take(static_cast<size_t>(B) * N * 4);
  *o_tiles = take(static_cast<size_t>(B) * T * kTileBytes);
  // candidate buffers; query tiles are padded to 128 rows, so allow for the rows past the end of the last block
  *o_cand = take((static_cast<size_t>(B) * N + 128) * 4 * kSubCands * 8);
  const size_t a = take(static_cast<size_t>(B) * N * 4), m = take(static_cast<size_t>(B) * 64 * 4), x = take(static_cast<size_t>(B) * 4);
  if (o_normc) *o_normc = a;
  if (o_mu) *o_mu = m;
  if (o_maxc) *o_maxc = x;
  return off + 1024;
}

extern "C" size_t dgcnn_knn_tc_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  size_t a, b, c, d, e, f;
  return knn_tc_layout(B, N, &a, &b, &c, &d, &e, &f);
}

extern "C" int dgcnn_knn_tc_supported(int C, int N, int k) { return (C > 16 && C <= 64 && k <= 48 && N >= 128) ? 1 : 0; }

extern "C" int dgcnn_knn_tc_f32(const float* rows, int ld, int B, int C, int N, int k, int32_t* idx, void* workspace,
                                size_t workspace_bytes, void* stream) {
  DG_REQUIRE(rows && idx && workspace, DGCNN_E_INVALID, "knn_tc: null pointer");
  DG_REQUIRE(B > 0 && N > 0 && C > 0 && B <= 65535, DGCNN_E_INVALID, "knn_tc: bad shape B=%d C=%d N=%d", B, C, N);
  DG_REQUIRE(k >= 1 && k <= N, DGCNN_E_INVALID, "knn_tc: selected index k out of range (k=%d, N=%d)", k, N);
  DG_REQUIRE(dgcnn_knn_tc_supported(C, N, k), DGCNN_E_UNSUPPORTED,
             "knn_tc: needs 16 < C <= 64, k <= 48, N >= 128 (C=%d N=%d k=%d); use dgcnn_knn_f32", C, N, k);
  const int C4 = (C + 3) / 4 * 4;
  DG_REQUIRE(ld >= C4 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(rows) & 15) == 0, DGCNN_E_INVALID,
             "knn_tc: rows must be 16-byte aligned with ld %% 4 == 0 and ld >= round4(C)");
  DG_REQUIRE(workspace_bytes >= dgcnn_knn_tc_workspace_bytes(B, N), DGCNN_E_WORKSPACE, "knn_tc: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t o_norms, o_max, o_cnt, o_rows, o_tiles, o_cand, o_normc, o_mu, o_maxc;
  knn_tc_layout(B, N, &o_norms, &o_max, &o_cnt, &o_rows, &o_tiles, &o_cand, &o_normc, &o_mu, &o_maxc);
  unsigned char* ws = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  float* norms = reinterpret_cast<float*>(ws + o_norms);
  unsigned int* maxnorm = reinterpret_cast<unsigned int*>(ws + o_max);
  int* fb_count = reinterpret_cast<int*>(ws + o_cnt);
  int* fb_rows = reinterpret_cast<int*>(ws + o_rows);
  unsigned char* tiles = ws + o_tiles;
  const int T = 2 * ((N + 127) / 128);

  DG_CUDA(cudaMemsetAsync(ws + o_max, 0, o_rows - o_max, st));     // maxnorm + fallback counter
  knn_norms_max_kernel<<<dim3((N + 255) / 256, B), 256, 0, st>>>(rows, ld, N, C, norms, maxnorm);
  DG_CUDA(cudaGetLastError());
  float* normc = reinterpret_cast<float*>(ws + o_normc);
  float* mu = reinterpret_cast<float*>(ws + o_mu);
  unsigned int* maxnormc = reinterpret_cast<unsigned int*>(ws + o_maxc);
  DG_CUDA(cudaMemsetAsync(maxnormc, 0, static_cast<size_t>(B) * 4, st));
  DG_CUDA(cudaMemsetAsync(mu, 0, static_cast<size_t>(B) * 64 * 4, st));
  knn_tc_mean_kernel<<<dim3(std::max(1, std::min(64, N / 256)), B), 256, 0, st>>>(rows, ld, N, C4, mu);
  DG_CUDA(cudaGetLastError());
  knn_tc_prep_kernel<<<dim3(std::min((T * 512 + 255) / 256, 4096), B), 256, 0, st>>>(rows, ld, N, C4, T, mu, tiles, normc, maxnormc);
  DG_CUDA(cudaGetLastError());
  DG_CUDA(cudaFuncSetAttribute(knn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(KnnTcSmem::bytes)));
  // pass-1 sample: the 8th best of a thread's 16*T1 keys estimates global rank N/(2*T1); aim at ~max(100, 3k)
  const int Tk = (N + kTcKeyTile - 1) / kTcKeyTile;
  const int target = std::max(100, 3 * k);
  const int T1 = std::max(1, std::min(Tk, (N + 2 * target - 1) / (2 * target)));
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(ws + o_cand);
  knn_tc_kernel<<<dim3((N + 127) / 128, B), kTcThreads, KnnTcSmem::bytes, st>>>(tiles, rows, ld, norms, maxnorm, normc, maxnormc, N,
                                                                               C4, T, k, T1, cand, idx, fb_count, fb_rows);
  DG_CUDA(cudaGetLastError());
  knn_fallback_kernel<<<64, kThreads, 0, st>>>(rows, ld, norms, N, C4, k, idx, fb_count, fb_rows);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}


// ---- low-dimensional exact kNN (C <= 16), threshold-collect-select ------------------------------------------------
static size_t knn_small_layout(int B, int N, size_t* o_norms, size_t* o_cnt, size_t* o_rows, size_t* o_cand) {
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 1023) & ~size_t(1023); return o; };
  *o_norms = take(static_cast<size_t>(B) * N * 4);
  *o_cnt = take(4);
  *o_rows = take(static_cast<size_t>(B) * N * 4);
  *o_cand = take((static_cast<size_t>(B) * N + kSmallRows) * 4 * kSubCands * 8);
  return off + 1024;
}

extern "C" int dgcnn_knn_small_supported(int C, int N, int k) { return (C <= 16 && k <= 48 && N >= 512) ? 1 : 0; }

extern "C" size_t dgcnn_knn_small_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  size_t a, b, c, d;
  return knn_small_layout(B, N, &a, &b, &c, &d);
}

extern "C" int dgcnn_knn_small_f32(const float* rows, int ld, int B, int C, int N, int k, int32_t* idx, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  DG_REQUIRE(rows && idx && workspace, DGCNN_E_INVALID, "knn_small: null pointer");
  DG_REQUIRE(B > 0 && N > 0 && C > 0 && B <= 65535, DGCNN_E_INVALID, "knn_small: bad shape B=%d C=%d N=%d", B, C, N);
  DG_REQUIRE(k >= 1 && k <= N, DGCNN_E_INVALID, "knn_small: selected index k out of range (k=%d, N=%d)", k, N);
  DG_REQUIRE(dgcnn_knn_small_supported(C, N, k), DGCNN_E_UNSUPPORTED,
             "knn_small: needs C <= 16, k <= 48, N >= 512 (C=%d N=%d k=%d); use dgcnn_knn_f32", C, N, k);
  const int C4 = (C + 3) / 4 * 4;
  DG_REQUIRE(ld >= C4 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(rows) & 15) == 0, DGCNN_E_INVALID,
             "knn_small: rows must be 16-byte aligned with ld %% 4 == 0 and ld >= round4(C)");
  DG_REQUIRE(workspace_bytes >= dgcnn_knn_small_workspace_bytes(B, N), DGCNN_E_WORKSPACE, "knn_small: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t o_norms, o_cnt, o_rows, o_cand;
  knn_small_layout(B, N, &o_norms, &o_cnt, &o_rows, &o_cand);
  unsigned char* ws = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  float* norms = reinterpret_cast<float*>(ws + o_norms);
  int* fb_count = reinterpret_cast<int*>(ws + o_cnt);
  int* fb_rows = reinterpret_cast<int*>(ws + o_rows);
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(ws + o_cand);
  DG_CUDA(cudaMemsetAsync(fb_count, 0, 4, st));
  // norms in torch.sum's order (rows beyond C are zero padded, so the padded chain is exact as well)
  knn_norms_max_kernel<<<dim3((N + 255) / 256, B), 256, 0, st>>>(rows, ld, N, C, norms, reinterpret_cast<unsigned int*>(ws + o_rows));
  DG_CUDA(cudaGetLastError());
  // sample size: a thread's 8th best of its n_sample/4 keys has expected global rank 32 N / n_sample; aim at max(100, 3k)
  const int target = std::max(100, 3 * k);
  const int n_sample = std::max(256, std::min(N, static_cast<int>(32LL * N / target)));
  int rc;
  if (C4 <= 4) rc = launch_knn_small<4>(rows, ld, norms, B, N, C4, k, n_sample, cand, idx, fb_count, fb_rows, st);
  else if (C4 <= 8) rc = launch_knn_small<8>(rows, ld, norms, B, N, C4, k, n_sample, cand, idx, fb_count, fb_rows, st);
  else rc = launch_knn_small<16>(rows, ld, norms, B, N, C4, k, n_sample, cand, idx, fb_count, fb_rows, st);
  if (rc) return rc;
  knn_fallback_kernel<<<64, kThreads, 0, st>>>(rows, ld, norms, N, C4, k, idx, fb_count, fb_rows);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}
