// knn_tc.cu — kNN graphs (model.py:25-40) on the tensor cores for every channel count C <= 128, with EXACT results.
//
// Gram tiles (128 queries x 128 keys; 128 x 64 for C > 64) are computed by tcgen05.mma on a 2-plane bf16 split of the
// block-centred features (3 MMAs per product, |error| <= ~2e-5 * |x_i||x_j|) into TMEM.  The approximate scores only act
// as a FILTER ("threshold - collect - re-score"); there are no running top-k lists and no CTA-wide barrier in the key loop:
//   pass 1  (scan kernel) scan ~16 % of the key tiles (spread over the block); every thread keeps the running maximum of
//           each of its 16 columns, so a query row owns 64 group maxima; tau_i = the r-th largest of them estimates the
//           row's ~R-th best score overall (R ~ max(100, 3k + 10));
//   pass 2  (scan kernel) scan all keys; the sign bits of (tau_i - score) are funnel-shifted into a 16-bit hit mask
//           (1 ALU op per score); only a thread with a non-zero mask parks its 16 scores in shared memory and appends
//           the hits to the row's candidate buffer (global memory, 4 private sub-buffers of 64 per row, no atomics);
//   pass 3  (select kernel, warp per row) bisect the candidates' approximate scores to a value v exceeded by >= k of
//           them, keep the candidates above v - 2 eps_i (eps_i = rigorous bound on |approx - exact| for the row),
//           re-score those with the reference's exact FP32 arithmetic (FMA chain in channel order, two rounded
//           subtractions, model.py:35-37) and emit the exact top-k;
//   certificate: every key outside the candidate set has exact score <= tau_i - |y_i|^2 + eps_i; if that is below the
//           k-th exact score found, the result is provably the exact top-k.  Rows that fail it (tau too tight, buffer
//           overflow, more than 64 near-ties) are redone by an exact brute-force kernel.
// The neighbour table is therefore identical to dgcnn_knn_f32's, bit for bit, and the B x N x N matrix never leaves
// the SM: scores go TMEM -> registers.
//
// Why this shape (B200 measurements, profiles/): the previous kernel (128 x 64 tiles, per-score compare/vote/branch,
// top-8 insertion lists in pass 1) was bound by the ALU pipe (FSETP/VOTE/FMNMX issue at half rate), not by the tensor
// pipe (14 % busy): ~350 warp instructions per 64-key tile and warp.  Here a score costs one FFMA2 half (2 v + kn),
// one FFMA2 half (tau - score) and one SHF; N = 128 MMAs halve the shared-memory operand bytes per FLOP.
#include "common.cuh"
#include "tc.cuh"
#include "select.cuh"

#include <cfloat>

namespace dg {

// error model of the filter (DESIGN.md "kNN filter bound"), RIGOROUS worst case per element:
//   x = x1 + x2 + ex with |ex| <= 2^-16 |x| (bf16 has 8 significant bits: unit roundoff 2^-8, applied twice); the three
//   MMAs x1*y1 + x1*y2 + x2*y1 miss x2*y2 + ex*y + (x1+x2)*ey <= 3 * 2^-16 |x_c y_c| = 4.58e-5 |x_c y_c|;
//   FP32 accumulation in TMEM (3 * C/16 steps) adds <= 0.07e-5 (C = 64) / 0.14e-5 (C = 128), and the reference's own FMA
//   chain is within C * 2^-24 = 0.38e-5 / 0.76e-5 of the exact dot product.  With Cauchy-Schwarz
//   |dot_tc - dot_ref| <= 5.03e-5 (5.48e-5) |x_i||x_j|, i.e. 1.01e-4 (1.10e-4) on 2*dot.
// (Random data stays ~10x below this — measured 4.2e-6 of sum|x_c y_c| at C = 64 — but a 4-5 term dot product (the input
// graph) comes within 35 % of it, and the SET mode of the select pass relies on eps being a true bound.)
constexpr float kC1 = 1.02e-4f;           // C <= 64
constexpr float kC1Wide = 1.11e-4f;       // 64 < C <= 128
constexpr float kC2 = 6.0e-7f;            // final roundings (incl. fl(tau - score)), relative to (|x_i|+|x_j|)^2

constexpr int kScanWarps = 16;
constexpr int kTcThreads = (kScanWarps + 1) * 32;   // + 1 producer / MMA-issue warp
constexpr int kSubCands = 64;            // candidate slots per (row, column quarter)
constexpr int kSurvivors = 64;           // exact re-scores per row (2 per lane)
constexpr int kGroups = 64;              // pass-1 group maxima per query row (4 threads x 16 columns)

// Key tiles in global memory, per block: tile t holds KEYS consecutive points as KA atoms (64 channels each) x 2 bf16
// planes in the SWIZZLE_128B K-major layout (row r of a plane at r*128 B), followed by the KEYS negated squared norms of
// the centred points (-inf for padding keys).  One tile is one contiguous bulk copy.
template <int KA, int KEYS>
struct KnnTile {
  static constexpr uint32_t plane_bytes = KEYS * 128;
  static constexpr uint32_t norm_off = KA * 2 * plane_bytes;
  static constexpr uint32_t bytes = norm_off + KEYS * 4;
};

// ---------------------------------------------------------------------------------------------
// prep: per-block channel means (the filter works on centred features: distances are translation invariant and the
// error bound scales with the norms, which centring shrinks by 10-100x on post-LeakyReLU features and on map coordinates)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) knn_tc_mean_kernel(const float* __restrict__ rows, int ld, int N, int C4,
                                                          float* __restrict__ musum) {
  // grid (slices, B): each CTA sums a slice of the block's rows; musum (zero-initialised) receives the channel SUMS.
  // thread = (16-byte channel chunk, row lane): coalesced float4 loads, 4 independent rows in flight per thread
  __shared__ float4 red[8][32];
  const int b = blockIdx.y, c4 = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int per = (N + gridDim.x - 1) / gridDim.x, n0 = blockIdx.x * per, n1 = min(N, n0 + per);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (4 * c4 < C4) {
    const float* base = rows + static_cast<size_t>(b) * N * ld + 4 * c4;
    int n = n0 + grp;
    for (; n + 24 < n1; n += 32) {
      const float4 a0 = *reinterpret_cast<const float4*>(base + static_cast<size_t>(n) * ld);
      const float4 a1 = *reinterpret_cast<const float4*>(base + static_cast<size_t>(n + 8) * ld);
      const float4 a2 = *reinterpret_cast<const float4*>(base + static_cast<size_t>(n + 16) * ld);
      const float4 a3 = *reinterpret_cast<const float4*>(base + static_cast<size_t>(n + 24) * ld);
      s.x += (a0.x + a1.x) + (a2.x + a3.x); s.y += (a0.y + a1.y) + (a2.y + a3.y);
      s.z += (a0.z + a1.z) + (a2.z + a3.z); s.w += (a0.w + a1.w) + (a2.w + a3.w);
    }
    for (; n < n1; n += 8) {
      const float4 a0 = *reinterpret_cast<const float4*>(base + static_cast<size_t>(n) * ld);
      s.x += a0.x; s.y += a0.y; s.z += a0.z; s.w += a0.w;
    }
  }
  red[grp][c4] = s;
  __syncthreads();
  if (threadIdx.x < 32 && 4 * c4 < C4) {
    float4 t = red[0][c4];
#pragma unroll
    for (int q = 1; q < 8; ++q) { t.x += red[q][c4].x; t.y += red[q][c4].y; t.z += red[q][c4].z; t.w += red[q][c4].w; }
    float* dst = musum + b * 128 + 4 * c4;
    atomicAdd(dst, t.x); atomicAdd(dst + 1, t.y); atomicAdd(dst + 2, t.z); atomicAdd(dst + 3, t.w);
  }
}

template <int KA, int KEYS>
__global__ void knn_tc_prep_kernel(const float* __restrict__ rows, int ld, int N, int C4, int T,
                                   const float* __restrict__ mu, unsigned char* __restrict__ tiles,
                                   float* __restrict__ normc, unsigned int* __restrict__ maxnormc_bits) {
  // one thread per (tile row, 8-channel chunk); grid.y = block.  The CPR chunk-threads of a row are consecutive lanes.
  using Tile = KnnTile<KA, KEYS>;
  constexpr int CPR = KA * 8;
  const int b = blockIdx.y;
  const int total = T * KEYS * CPR;
  const float inv_n = 1.0f / static_cast<float>(N);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ch = i % CPR, rl = (i / CPR) % KEYS, t = i / (CPR * KEYS);
    const int n = t * KEYS + rl;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (n < N && ch * 8 < C4) {
      const float* src = rows + (static_cast<size_t>(b) * N + n) * ld + ch * 8;
      const float* m = mu + b * 128 + ch * 8;       // channel sums; any value near the mean is a valid centre
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
#pragma unroll
    for (int o = 1; o < CPR; o <<= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    unsigned char* tile = tiles + (static_cast<size_t>(b) * T + t) * Tile::bytes;
    const uint32_t off = static_cast<uint32_t>(ch >> 3) * 2 * Tile::plane_bytes + tc::unit_elem_offset(rl, (ch & 7) * 8);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      __nv_bfloat16 h[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        h[e] = __float2bfloat16_rn(v[e]);
        v[e] -= __bfloat162float(h[e]);
      }
      *reinterpret_cast<uint4*>(tile + p * Tile::plane_bytes + off) = *reinterpret_cast<const uint4*>(h);
    }
    if (ch == 0) {
      // MINUS the centred squared norm; -inf for padding keys: their score 2*0 - inf never passes a threshold, so the
      // scans need no per-key bounds check
      reinterpret_cast<float*>(tile + Tile::norm_off)[rl] = n < N ? -sq : -INFINITY;
      if (n < N) {
        normc[static_cast<size_t>(b) * N + n] = sq;
        atomicMax(maxnormc_bits + b, __float_as_uint(sq));
      }
    }
  }
}

__global__ void __launch_bounds__(128) knn_norms_max_kernel(const float* __restrict__ rows, int ld, int N, int C,
                                                            float* __restrict__ norms, unsigned int* __restrict__ maxnorm_bits) {
  // |x|^2 in torch.sum's order (sequential in blocks of 16 channels, block sums added sequentially: the same order as
  // norms_kernel in knn.cu); per-block max via atomicMax on the (non-negative) float bits.
  // 128 rows per CTA: 64-channel slabs are staged through shared memory with coalesced 16-byte loads (a thread reading
  // its own 256-byte row touched 32 different lines per warp load), then thread = row sums its slab row by row.
  __shared__ float tile[128 * 65];
  const int b = blockIdx.y, n0 = blockIdx.x * 128, tid = threadIdx.x;
  const int n = n0 + tid;
  const int C4 = (C + 3) & ~3;
  float lvl1 = 0.f, tail = 0.f;
  for (int h = 0; h < C; h += 64) {
    const int cw = min(64, C - h), cw4 = min(64, C4 - h);
    __syncthreads();
    for (int i = tid; i < 128 * 16; i += 128) {
      const int r = i >> 4, c4 = i & 15;
      if (n0 + r < N && 4 * c4 < cw4) {
        const float4 v = *reinterpret_cast<const float4*>(rows + (static_cast<size_t>(b) * N + n0 + r) * ld + h + 4 * c4);
        float* t = tile + r * 65 + 4 * c4;
        t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
      }
    }
    __syncthreads();
    if (n < N) {
      const float* r = tile + tid * 65;
      int c = 0;
      for (; c + 16 <= cw; c += 16) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) a = __fadd_rn(a, __fmul_rn(r[c + j], r[c + j]));
        lvl1 = __fadd_rn(lvl1, a);
      }
      for (; c < cw; ++c) tail = __fadd_rn(tail, __fmul_rn(r[c], r[c]));
    }
  }
  float val = 0.f;
  if (n < N) {
    val = __fadd_rn(tail, lvl1);
    norms[static_cast<size_t>(b) * N + n] = val;
  }
  unsigned int m = __float_as_uint(val);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(maxnorm_bits + b, m);
}

// exact reference score of (query row xi, key row xj): FMA chain over channels, then the two rounded subtractions.
// The key row is a gather from L2: all 16-byte loads of a 64-channel half are issued before the first use — with a rolled
// loop every 4 channels paid a full L2 round trip.
template <int NB = 16>
__device__ __forceinline__ float exact_score(const float* __restrict__ xi, const float* __restrict__ xj, int C4,
                                             float norm_i, float norm_j) {
  // NB float4 chunks (4 NB channels) of the key row are in flight at a time: 16 for the passes that re-score two keys per
  // lane anyway, 8 where register pressure decides the occupancy
  constexpr int W = 4 * NB;
  float d = 0.f;
  for (int h = 0; h < C4; h += W) {
    const int cw = min(W, C4 - h);
    const float* xjh = xj + h;
    float4 b[NB];
    if (cw == W && (reinterpret_cast<uintptr_t>(xjh) & 31) == 0) {
      // one lane reads one whole key row: 32-byte loads fetch every sector once (16-byte loads request each sector twice,
      // and with many rows in flight the second request usually misses L1 again)
#pragma unroll
      for (int q = 0; q < NB; q += 2)
        asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=f"(b[q].x), "=f"(b[q].y), "=f"(b[q].z), "=f"(b[q].w), "=f"(b[q + 1].x), "=f"(b[q + 1].y), "=f"(b[q + 1].z),
                       "=f"(b[q + 1].w)
                     : "l"(xjh + 4 * q));
    } else {
#pragma unroll
      for (int q = 0; q < NB; ++q)
        b[q] = 4 * q < cw ? *reinterpret_cast<const float4*>(xjh + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      if (4 * q < cw) {                 // the chain order is the reference's (channel by channel)
        const float4 a = *reinterpret_cast<const float4*>(xi + h + 4 * q);
        d = __fmaf_rn(a.x, b[q].x, d);
        d = __fmaf_rn(a.y, b[q].y, d);
        d = __fmaf_rn(a.z, b[q].z, d);
        d = __fmaf_rn(a.w, b[q].w, d);
      }
    }
  }
  return __fsub_rn(__fmaf_rn(d, 2.0f, -norm_j), norm_i);
}

// candidate entry of the scan pass: approximate score (raw float bits) in the high word, key index in the low word.
// (No order-preserving key transform here: the scan's hit path is issue-bound; the select pass compares the floats.)
__device__ __forceinline__ float cand_score(unsigned long long e) { return __uint_as_float(static_cast<uint32_t>(e >> 32)); }
__device__ __forceinline__ int cand_index(unsigned long long e) { return static_cast<int>(static_cast<uint32_t>(e)); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// mask <- (mask << 1) | sign(d): one funnel shift per score
__device__ __forceinline__ uint32_t shift_in_sign(uint32_t mask, float d) {
  uint32_t r;
  asm("shf.l.wrap.b32 %0, %1, %2, 1;" : "=r"(r) : "r"(__float_as_uint(d)), "r"(mask));
  return r;
}

// ---------------------------------------------------------------------------------------------
// scan kernel: pass 1 (threshold) + pass 2 (collect).  One CTA = 128 query rows of one block.
// ---------------------------------------------------------------------------------------------
template <int KA, int KEYS>
struct KnnScan {
  using Tile = KnnTile<KA, KEYS>;
  static constexpr int kStages = KA == 1 ? 4 : 3;
  // TMEM accumulators = shared-memory stages.  The producer tests `sfree` of a slot by phase PARITY, which is only
  // unambiguous while the barrier is at most one phase behind: with A accumulators the MMA issue (and so the prefetch
  // test) runs up to A tiles ahead of the slowest scan warp, and a slot that is two releases behind reads as "free"
  // (seen on B200 with 4 accumulators over 3 stages: a tile was overwritten under its MMA).  A == stages, together with
  // the scan warps arriving on `sfree` BEFORE `accfree`, bounds the lag at one phase.
  static constexpr int kAccs = kStages;
  static constexpr int CH = KEYS / 64;                                      // 16-column chunks per thread and tile
  static constexpr int QC = KEYS / 4;                                       // columns per column quarter
  static constexpr uint32_t stage_bytes = (Tile::bytes + 1023u) & ~1023u;
  static constexpr uint32_t a_plane = 128 * 128;                            // one plane of one atom of the 128 query rows
  static constexpr uint32_t a_bytes = KA * 2 * a_plane;
  static constexpr uint32_t b_off = a_bytes;
  static constexpr uint32_t scr_off = b_off + kStages * stage_bytes;        // per-thread score park (16 floats + pad) /
  static constexpr uint32_t scr_stride = 20;                                //   pass-1 group maxima [128][68]
  static constexpr uint32_t scr_bytes = kScanWarps * 32 * scr_stride * 4;
  static constexpr uint32_t tau_off = scr_off + scr_bytes;
  static constexpr uint32_t bar_off = tau_off + 128 * 4;
  static constexpr uint32_t bytes = bar_off + 256;
  static constexpr uint32_t tmem_cols = kAccs * KEYS <= 256 ? 256 : 512;        // power of two >= kAccs * KEYS
  static_assert(128 * 68 * 4 <= scr_bytes, "group maxima must fit in the score park");
};

template <int KA, int KEYS>
__global__ void __launch_bounds__(kTcThreads, 1)
knn_tc_scan_kernel(const unsigned char* __restrict__ tiles, int N, int T, int nk16_last, int T1, int rsel,
                   unsigned long long* __restrict__ cand, int* __restrict__ cnt4, float* __restrict__ tau_out) {
  using S = KnnScan<KA, KEYS>;
  using Tile = KnnTile<KA, KEYS>;
  constexpr int kStages = S::kStages, kAccs = S::kAccs, CH = S::CH, QC = S::QC;
  // no static shared memory in this kernel: the dynamic window starts 1024-byte aligned (checked below)
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* As = sm;
  unsigned char* Bs = sm + S::b_off;
  float* scr = reinterpret_cast<float*>(sm + S::scr_off);
  float* tauS = reinterpret_cast<float*>(sm + S::tau_off);
  uint64_t* abar = reinterpret_cast<uint64_t*>(sm + S::bar_off);
  uint64_t* full = abar + 1;                 // [kStages]  tile landed in smem
  uint64_t* sfree = full + kStages;          // [kStages]  all scan warps have read the slot's key norms; MMA done (implied)
  uint64_t* accfull = sfree + kStages;       // [kAccs]    MMAs of the tile finished
  uint64_t* accfree = accfull + kAccs;       // [kAccs]    all scan warps have read the accumulator
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfree + kAccs);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y, qt = blockIdx.x, q0 = qt * 128;
  const int Tk = (N + KEYS - 1) / KEYS;                    // key tiles of the block
  const int TT = T1 + Tk;                                  // tiles streamed: pass 1 (T1 sampled tiles) then pass 2 (all)
  const unsigned char* btiles = tiles + static_cast<size_t>(b) * T * Tile::bytes;

  if (tid == 0) {
    if (tc::smem_u32(sm) & 1023u) __trap();
    tc::mbar_init(abar, 1);
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(full + s, 1);
      tc::mbar_init(sfree + s, kScanWarps);
    }
    for (int a = 0; a < kAccs; ++a) {
      tc::mbar_init(accfull + a, 1);
      tc::mbar_init(accfree + a, kScanWarps);
    }
    tc::fence_barrier_init();
  }
  if (warp == kScanWarps) tc::tmem_alloc(tmem_slot, S::tmem_cols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t taddr = *tmem_slot;

  if (warp == kScanWarps) {
    // ===== producer + MMA issuer (one thread) =================================================================
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::idesc_bf16(128, KEYS);
      constexpr int NT = 128 / KEYS;                        // source tiles that make up the 128 query rows
      tc::mbar_arrive_expect_tx(abar, S::a_bytes);
#pragma unroll
      for (int h = 0; h < NT; ++h)
#pragma unroll
        for (int ap = 0; ap < KA * 2; ++ap)
          tc::bulk_g2s(As + ap * S::a_plane + h * Tile::plane_bytes,
                       btiles + static_cast<size_t>(NT * qt + h) * Tile::bytes + ap * Tile::plane_bytes, Tile::plane_bytes, abar);
      int loaded = 0;
      auto issue_copy = [&](int l) {
        const int s = l % kStages;
        // pass-1 sample tiles are spread over the block (robust against spatially sorted inputs)
        const int tile = l < T1 ? static_cast<int>((static_cast<long long>(l) * Tk) / T1) : l - T1;
        tc::mbar_arrive_expect_tx(full + s, Tile::bytes);
        tc::bulk_g2s(Bs + s * S::stage_bytes, btiles + static_cast<size_t>(tile) * Tile::bytes, Tile::bytes, full + s);
      };
      const uint64_t da = tc::smem_desc_sw128(tc::smem_u32(As));
      const uint32_t a_lo = static_cast<uint32_t>(da), hi = static_cast<uint32_t>(da >> 32);
      tc::mbar_wait(abar, 0);
      for (int i = 0; i < TT; ++i) {
        // tile i must be on its way; further tiles are prefetched only if their slot is already free (never block
        // the MMA issue on a slot that the scan warps still hold)
        while (loaded < TT && loaded < i + kStages) {
          uint64_t* bar = sfree + loaded % kStages;
          const uint32_t par = ((loaded / kStages) & 1) ^ 1;
          if (loaded <= i) tc::mbar_wait(bar, par);
          else if (!tc::mbar_try_wait(bar, par)) break;
          issue_copy(loaded++);
        }
        const int s = i % kStages, a = i % kAccs;
        tc::mbar_wait(full + s, (i / kStages) & 1);
        tc::mbar_wait(accfree + a, ((i / kAccs) & 1) ^ 1);
        tc::fence_after_sync();
        const uint32_t b_lo = static_cast<uint32_t>(tc::smem_desc_sw128(tc::smem_u32(Bs + s * S::stage_bytes)));
        const uint32_t tacc = taddr + a * KEYS;
        uint32_t acc = 0u;
        // 2-plane split: x1*y2, x2*y1, x1*y1 (smallest magnitude first); only the k16 steps that hold channels are issued
#pragma unroll
        for (int atom = 0; atom < KA; ++atom) {
          const int steps = atom == KA - 1 ? nk16_last : 4;
#pragma unroll
          for (int pr = 0; pr < 3; ++pr) {
            const uint32_t pa = pr == 1 ? 1u : 0u, pb = pr == 0 ? 1u : 0u;
            for (int k16 = 0; k16 < steps; ++k16) {
              tc::mma_bf16(tacc, tc::desc_at(a_lo, hi, (atom * 2 + pa) * S::a_plane + k16 * 32),
                           tc::desc_at(b_lo, hi, (atom * 2 + pb) * Tile::plane_bytes + k16 * 32), idesc, acc);
              acc = 1u;
            }
          }
        }
        tc::mma_commit(accfull + a);
      }
    }
  } else {
    // ===== scan warps: TMEM lane quarter (warp & 3), column quarter (warp >> 2); thread = 1 query row x QC keys ====
    const int quarter = warp & 3, cq = warp >> 2;
    const int erow = quarter * 32 + lane;
    const bool row_ok = q0 + erow < N;
    const size_t g = static_cast<size_t>(b) * N + q0 + erow;          // valid only if row_ok
    const uint32_t tbase = taddr + (static_cast<uint32_t>(quarter * 32) << 16) + cq * QC;
    const float2 two = make_float2(2.0f, 2.0f);

    // One tile's scores sc[ch][c] = 2 * gram + kn (the row constant |y_i|^2 is left out) in two halves, so that the TMEM
    // loads of tile i + 1 are in flight while tile i is being scanned (software pipelining: with 4 warps per scheduler the
    // barrier wake-up + tcgen05.ld latency was exposed once per tile):
    //   issue_loads(i, r)   wait for the MMAs of tile i, issue the tcgen05.ld's (no wait)
    //   finish_loads(i, r)  wait::ld, add the key norms, hand the accumulator and the key slot back
    auto issue_loads = [&](int i, uint32_t (&r)[CH][16]) {
      const int a = i % kAccs;
      tc::mbar_wait(accfull + a, (i / kAccs) & 1);
      tc::fence_after_sync();
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) tc::tmem_ld16_issue(tbase + a * KEYS + ch * 16, r[ch]);
    };
    auto finish_loads = [&](int i, uint32_t (&r)[CH][16], float (&sc)[CH][16]) {
      const int s = i % kStages, a = i % kAccs;
      const float4* kn4 = reinterpret_cast<const float4*>(Bs + s * S::stage_bytes + Tile::norm_off) + cq * (QC / 4);
      float4 kk[CH][4];
#pragma unroll
      for (int ch = 0; ch < CH; ++ch)
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) kk[ch][q4] = kn4[ch * 4 + q4];
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) tc::tmem_ld_wait16(r[ch]);
#pragma unroll
      for (int ch = 0; ch < CH; ++ch)
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          // two keys per FFMA2 (same roundings as the scalar form); padding keys score -inf
          const float2 s0 = __ffma2_rn(make_float2(__uint_as_float(r[ch][4 * q4]), __uint_as_float(r[ch][4 * q4 + 1])), two,
                                       make_float2(kk[ch][q4].x, kk[ch][q4].y));
          const float2 s1 = __ffma2_rn(make_float2(__uint_as_float(r[ch][4 * q4 + 2]), __uint_as_float(r[ch][4 * q4 + 3])), two,
                                       make_float2(kk[ch][q4].z, kk[ch][q4].w));
          sc[ch][4 * q4] = s0.x; sc[ch][4 * q4 + 1] = s0.y; sc[ch][4 * q4 + 2] = s1.x; sc[ch][4 * q4 + 3] = s1.y;
        }
      // every lane has consumed its TMEM registers and key norms (convergent code above): hand both back
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(sfree + s);            // slot first: an observed `accfree` phase implies the slot's release
        tc::mbar_arrive(accfree + a);
      }
    };
    auto load_scores = [&](int i, float (&sc)[CH][16]) {
      uint32_t r[CH][16];
      issue_loads(i, r);
      finish_loads(i, r, sc);
    };

    // ---- pass 1: threshold estimate from T1 sampled tiles ----------------------------------------------------------
    float tau = -FLT_MAX;                         // T1 == 0 (tiny blocks): every key is a candidate
    if (T1 > 0) {
      float gm[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) gm[c] = -INFINITY;
      for (int i = 0; i < T1; ++i) {
        float sc[CH][16];
        load_scores(i, sc);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          if (CH == 2) gm[c] = fmaxf(fmaxf(gm[c], sc[0][c]), sc[CH - 1][c]);      // FMNMX3
          else gm[c] = fmaxf(gm[c], sc[0][c]);
        }
      }
      // group maxima of the row: [128][68] floats (16-byte stores, conflict-free), then a warp sorts each row's 64 values
      float4* dst = reinterpret_cast<float4*>(scr + erow * 68 + cq * 16);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) dst[q4] = make_float4(gm[4 * q4], gm[4 * q4 + 1], gm[4 * q4 + 2], gm[4 * q4 + 3]);
      named_bar_sync(1, kScanWarps * 32);
      for (int rr = 0; rr < 8; ++rr) {
        const int row = warp * 8 + rr;
        unsigned long long e0 = make_key(scr[row * 68 + lane], lane), e1 = make_key(scr[row * 68 + 32 + lane], lane + 32);
        sel::sort64_desc(e0, e1, lane);
        const unsigned long long kth = sel::list_at(e0, e1, rsel - 1);
        if (lane == 0) tauS[row] = fmaxf(sel::key_score(kth), -FLT_MAX);
      }
      named_bar_sync(1, kScanWarps * 32);         // tauS complete; the score park may be reused
      tau = tauS[erow];
    }
    if (row_ok) tau_out[g] = tau;

    // ---- pass 2: collect every key above tau --------------------------------------------------------------------------
    uint2* wp = reinterpret_cast<uint2*>(cand + (g * 4 + cq) * kSubCands);       // next free slot of this thread's sub-buffer
    float* park = scr + (warp * 32 + lane) * S::scr_stride;
    const float tau_eff = row_ok ? tau : INFINITY;       // rows past the end of the block collect nothing
    const float2 tau2 = make_float2(tau_eff, tau_eff), neg1 = make_float2(-1.0f, -1.0f);
    int n = 0;
    float sc[CH][16];
    load_scores(T1, sc);                                   // (TT > T1: a block has at least one key tile)
    for (int i = T1; i < TT; ++i) {
      uint32_t rn[CH][16];
      if (i + 1 < TT) issue_loads(i + 1, rn);              // next tile's accumulator -> registers while this one is scanned
      const int j0 = (i - T1) * KEYS + cq * QC;
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) {
        // sign(tau - score) of column c ends up in bit 15 - c.  Four independent 4-bit chains: one 16-deep chain of
        // dependent funnel shifts left the warp waiting on the 4-cycle ALU latency with only 4 warps per scheduler.
        uint32_t mq[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const float2 d0 = __ffma2_rn(make_float2(sc[ch][4 * q4], sc[ch][4 * q4 + 1]), neg1, tau2);
          const float2 d1 = __ffma2_rn(make_float2(sc[ch][4 * q4 + 2], sc[ch][4 * q4 + 3]), neg1, tau2);
          uint32_t mm = shift_in_sign(0u, d0.x);
          mm = shift_in_sign(mm, d0.y);
          mm = shift_in_sign(mm, d1.x);
          mq[q4] = shift_in_sign(mm, d1.y);
        }
        const uint32_t mask = (((mq[0] << 4) | mq[1]) << 8) | ((mq[2] << 4) | mq[3]);
        if (mask != 0u) {
          // rare per thread (a few % of the chunks): park the 16 scores so that the hit columns can be indexed
          float4* p4 = reinterpret_cast<float4*>(park);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) p4[q4] = make_float4(sc[ch][4 * q4], sc[ch][4 * q4 + 1], sc[ch][4 * q4 + 2], sc[ch][4 * q4 + 3]);
          uint32_t m = mask;
          const int jc = j0 + ch * 16 + 15;
          do {
            const int pos = 31 - __clz(m);                    // highest set bit first = lowest column first (c = 15 - pos)
            m ^= 1u << pos;
            if (n < kSubCands) *wp = make_uint2(static_cast<uint32_t>(jc - pos), __float_as_uint(park[15 - pos]));
            ++wp;
            ++n;
          } while (m);
        }
      }
      if (i + 1 < TT) finish_loads(i + 1, rn, sc);
    }
    if (row_ok) cnt4[g * 4 + cq] = n;
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == kScanWarps) tc::tmem_dealloc(taddr, S::tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// select kernel (pass 3): warp per query row — bisect, re-score exactly, certify
//
// With |approx_j - (exact_j + |y_i|^2)| <= eps for every key of the row, and a cut `lo` exceeded by c >= k candidates:
//   * a key with approx <= lo - 2 eps has k keys strictly above it in exact score: out;
//   * a_k = the k-th largest approximate score (lo < a_k): a key with approx > a_k + 2 eps can only be beaten by keys
//     whose approx exceeds a_k, of which there are fewer than k: certainly IN, no exact score needed ("sure");
//   * the keys in between ("band", typically 2-6) are re-scored with the reference's exact FP32 arithmetic and fill the
//     remaining k - #sure places by exact score (ties: lower index), exactly like dgcnn_knn_f32.
// ORDERED = false emits the neighbour SET (sure keys first, then the selected band keys): only the band rows are
// gathered — 10x less L2 traffic than re-scoring all k survivors, which is what bounded this pass.
// ORDERED = true (the public knn() / get_graph_feature(), whose order is the reference's topk order) re-scores every
// survivor and sorts: the table equals dgcnn_knn_f32's element for element.
// Certificate: every key that was never a candidate has exact score <= tau_i - |y_i|^2 + eps; the sure keys beat that
// by construction (a_k > tau_i), the selected band keys are checked with their exact scores.
// ---------------------------------------------------------------------------------------------
template <bool ORDERED>
__global__ void __launch_bounds__(kThreads, ORDERED ? 2 : 4)
knn_tc_select_kernel(const float* __restrict__ rows, int ld, const float* __restrict__ norms,
                     const unsigned int* __restrict__ maxnorm_bits, const float* __restrict__ normc,
                     const unsigned int* __restrict__ maxnormc_bits, int N, int C4, int k,
                     const unsigned long long* __restrict__ cand, const int* __restrict__ cnt4,
                     const float* __restrict__ tau_in, int32_t* __restrict__ idx, int* __restrict__ fb_count,
                     int* __restrict__ fb_rows) {
  __shared__ unsigned long long stageS[kThreads / 32][kSurvivors];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y, q = blockIdx.x * (kThreads / 32) + warp;
  if (q >= N) return;
  const unsigned lt_mask = (1u << lane) - 1u;
  const float c1 = C4 > 64 ? kC1Wide : kC1;
  const float rmax = sqrtf(__uint_as_float(maxnormc_bits[b]));      // largest CENTRED norm of the block
  const float rx = sqrtf(__uint_as_float(maxnorm_bits[b]));         // largest raw norm (rounding of the centring)
  unsigned long long* stg = stageS[warp];
  const size_t g = static_cast<size_t>(b) * N + q;
  const int4 c4 = *reinterpret_cast<const int4*>(cnt4 + g * 4);
  const int cmax = max(max(c4.x, c4.y), max(c4.z, c4.w));
  const bool overflow = cmax > kSubCands;                                   // a sub-buffer overflowed
  const bool big = cmax > 32;                                               // warp-uniform: second half of the sub-buffers in use
  const int cs[4] = {min(c4.x, kSubCands), min(c4.y, kSubCands), min(c4.z, kSubCands), min(c4.w, kSubCands)};
  const int ntot = cs[0] + cs[1] + cs[2] + cs[3];
  bool bad = overflow || ntot < k;                                          // ... or tau was too tight
  const unsigned long long* rc = cand + g * 4 * kSubCands;
  // key[r], r < 4: slot `lane` of sub-buffer r; r >= 4: slot 32 + lane of sub-buffer r - 4 (only if `big`)
  unsigned long long key[8];
  float ap[8];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    key[r] = lane < cs[r] ? rc[r * kSubCands + lane] : 0ull;
    ap[r] = lane < cs[r] ? cand_score(key[r]) : -INFINITY;
    key[4 + r] = 0ull;
    ap[4 + r] = -INFINITY;
  }
  if (big) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      key[4 + r] = 32 + lane < cs[r] ? rc[r * kSubCands + 32 + lane] : 0ull;
      ap[4 + r] = 32 + lane < cs[r] ? cand_score(key[4 + r]) : -INFINITY;
    }
  }
  auto count_above = [&](float t) {
    int c = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) c += __popc(__ballot_sync(0xffffffffu, ap[r] > t));
    if (big) {
#pragma unroll
      for (int r = 4; r < 8; ++r) c += __popc(__ballot_sync(0xffffffffu, ap[r] > t));
    }
    return c;
  };
  const float tau_r = tau_in[g];
  const bool tau_open = !(tau_r > -FLT_MAX);       // no threshold was applied (tiny blocks)
  // bisection: lo with at least min(ntot, k) candidates strictly above it (and fewer than that above hi)
  float lo = tau_r, hi = fmaxf(fmaxf(ap[0], ap[1]), fmaxf(ap[2], ap[3]));
  if (big) hi = fmaxf(hi, fmaxf(fmaxf(ap[4], ap[5]), fmaxf(ap[6], ap[7])));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  if (tau_open) {                                  // start from the smallest candidate
    float mn = INFINITY;
#pragma unroll
    for (int r = 0; r < 8; ++r) mn = fminf(mn, ap[r] > -INFINITY ? ap[r] : INFINITY);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    lo = mn - fabsf(mn) * 1e-6f - 1e-30f;
  }
  const int want = min(ntot, k);
  int cl = ntot;                                   // candidates strictly above lo
  for (int it = 0; it < 12; ++it) {
    const float mid = 0.5f * (lo + hi);
    const int c = count_above(mid);
    if (c >= want) { lo = mid; cl = c; } else hi = mid;
    if (c >= want && c <= want + 3) break;               // close enough: every extra survivor is a row gather
  }
  const float norm_i = norms[g];                 // reference |x_i|^2 (exact scores)
  const float nci = normc[g];                    // centred |y_i|^2: approx score of the row = exact pd + nci
  const float ni = sqrtf(nci);
  // Only keys with |y_j| <= |y_i| + d_tau can have a score near tau (score <= |y_i|^2 - (|y_j| - |y_i|)^2), so the
  // error bound uses R_i = min(R, |y_i| + 1.1 d_tau) instead of the block maximum R: one far outlier does not
  // inflate every row's margin.  Third term: rounding of y = fl(x - mu) (2^-22 Rx (|y_i| + 2 R_i)).
  const float dtau = sqrtf(fmaxf(nci - tau_r, 0.f));
  const float ri = tau_open ? rmax : fminf(rmax, ni + 1.1f * dtau + 1e-3f * ni);
  const float eps = 1.01f * (c1 * ni * ri + kC2 * (ni + ri) * (ni + ri) + 2.4e-7f * rx * (ni + 2.0f * ri));
  const float cut = lo - 2.0f * eps;
  const float* xi = rows + g * ld;

  if (!ORDERED) {
    if (bad || cl > k + 3) {          // overflow / too few candidates / a run of > 4 equal approximate scores at the cut
      if (lane == 0) {
        atomicAdd(fb_count + (overflow ? 1 : (ntot < k ? 2 : 3)), 1);
        fb_rows[atomicAdd(fb_count, 1)] = static_cast<int>(g);
      }
      return;
    }
    // ---- neighbour SET: a_k by peeling the cl - k smallest values above lo (groups of equal values at once) --------
    float ak = lo;
    int above = cl;                               // candidates strictly above ak
    for (int t = 0; t < 4; ++t) {
      float mn = INFINITY;
#pragma unroll
      for (int r = 0; r < 8; ++r) mn = fminf(mn, ap[r] > ak ? ap[r] : INFINITY);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      const int c = count_above(mn);
      ak = mn;
      above = c;
      if (c < k) break;                           // count(> mn) < k <= count(>= mn): mn is the k-th largest
    }
    const float sure_cut = ak + 2.0f * eps;
    // sure keys -> idx[0 .. s), band keys -> staging
    int s = 0, nb = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r >= 4 && !big) break;
      const bool is_sure = ap[r] > sure_cut;
      const bool is_band = !is_sure && ap[r] > cut;
      const unsigned ms = __ballot_sync(0xffffffffu, is_sure), mb = __ballot_sync(0xffffffffu, is_band);
      if (is_sure) idx[g * k + s + __popc(ms & lt_mask)] = cand_index(key[r]);      // s + ... < k: fewer than k keys exceed a_k
      const int pb = nb + __popc(mb & lt_mask);
      if (is_band && pb < kSurvivors) stg[pb] = key[r];
      s += __popc(ms);
      nb += __popc(mb);
    }
    (void)above;
    const int need = k - s;                       // >= 1: a_k itself is a band key
    bad = nb > kSurvivors || nb < need;           // too many near-ties to re-score here (nb < need cannot happen)
    const int ns = min(nb, kSurvivors);
    __syncwarp();
    unsigned long long ex[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      ex[r] = 0ull;
      const int e = lane + 32 * r;
      if (e < ns) {
        const int j = cand_index(stg[e]);
        const size_t gj = static_cast<size_t>(b) * N + j;
        ex[r] = make_key(exact_score<8>(xi, rows + gj * ld, C4, norm_i, norms[gj]), j);
      }
    }
    __syncwarp();
    unsigned long long last = ~0ull;              // smallest selected exact key
    if (ns <= 32) {
      // rank of each band key among the band (distinct keys): selected iff rank < need
      int rank = 0;
      for (int t = 0; t < ns; ++t) rank += __shfl_sync(0xffffffffu, ex[0], t) > ex[0];
      const bool selk = lane < ns && rank < need;
      if (selk) idx[g * k + s + rank] = key_index(ex[0]);
      last = selk ? ex[0] : ~0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, last, o);
        last = other < last ? other : last;
      }
    } else {
      sel::sort64_desc(ex[0], ex[1], lane);
      if (lane < need) idx[g * k + s + lane] = key_index(ex[0]);
      if (lane + 32 < need) idx[g * k + s + lane + 32] = key_index(ex[1]);
      last = sel::list_at(ex[0], ex[1], need - 1);
    }
    const float pdl = sel::key_score(last);
    const bool cert = tau_open || (tau_r - nci) + eps < pdl;
    if (lane == 0) {
      if (nb > kSurvivors) atomicAdd(fb_count + 3, 1);
      else if (!cert) atomicAdd(fb_count + 4, 1);
    }
    bad = bad || !cert;
    if (bad && lane == 0) fb_rows[atomicAdd(fb_count, 1)] = static_cast<int>(g);
    return;
  }

  // ---- ordered table (and the rows the set path cannot take): re-score every survivor above the cut, sort ----------
  int base = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const bool keep = ap[r] > cut;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int pos = base + __popc(m & lt_mask);
    if (keep && pos < kSurvivors) stg[pos] = key[r];
    base += __popc(m);
  }
  bad = bad || base > kSurvivors;                                         // too many near-ties to re-score here
  const int ns = min(base, kSurvivors);
  __syncwarp();
  unsigned long long ex[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    ex[r] = 0ull;
    const int e = lane + 32 * r;
    if (e < ns) {
      const int j = cand_index(stg[e]);
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
  const bool cert = tau_open || (tau_r - nci) + eps < pdk;
  if (lane == 0) {     // diagnostics: fb_count[1..4] count the reasons (overflow, too few, too many near-ties, certificate)
    if (overflow) atomicAdd(fb_count + 1, 1);
    else if (ntot < k) atomicAdd(fb_count + 2, 1);
    else if (base > kSurvivors) atomicAdd(fb_count + 3, 1);
    else if (!cert) atomicAdd(fb_count + 4, 1);
  }
  bad = bad || !cert;
  if (bad && lane == 0) fb_rows[atomicAdd(fb_count, 1)] = static_cast<int>(g);
}

// insert v into the descending top-8 t[0..7] (registers)
__device__ __forceinline__ void top8_insert(float (&t)[8], float v) {
#pragma unroll
  for (int i = 7; i >= 1; --i) t[i] = v > t[i - 1] ? t[i - 1] : fmaxf(v, t[i]);
  t[0] = fmaxf(v, t[0]);
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
  __shared__ __align__(16) float xi[128];
  __shared__ unsigned long long wl[8][64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int count = *fb_count;
  for (int f = blockIdx.x; f < count; f += gridDim.x) {
    const size_t g = static_cast<size_t>(fb_rows[f]);
    const size_t base = (g / N) * N;
    __syncthreads();
    if (tid < 128) xi[tid] = tid < C4 ? rows[g * ld + tid] : 0.f;
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

// workspace layout shared by the host entry points (all offsets 1024-byte aligned)
struct KnnTcLayout {
  size_t norms, maxn, cnt, fbrows, tiles, cand, normc, mu, maxc, tau, cnt4, total;
  int T;
  uint32_t tile_bytes;
};
static KnnTcLayout knn_tc_layout(int B, int N, int C) {
  KnnTcLayout L;
  const bool wide = C > 64;
  L.tile_bytes = wide ? KnnTile<2, 64>::bytes : KnnTile<1, 128>::bytes;
  L.T = (wide ? 2 : 1) * ((N + 127) / 128);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 1023) & ~size_t(1023); return o; };
  const size_t rows_pad = static_cast<size_t>(B) * N + 128;        // query tiles are padded to 128 rows
  L.norms = take(static_cast<size_t>(B) * N * 4);
  L.maxn = take(static_cast<size_t>(B) * 4);
  L.cnt = take(32);
  L.fbrows = take(static_cast<size_t>(B) * N * 4);
  L.tiles = take(static_cast<size_t>(B) * L.T * L.tile_bytes);
  L.cand = take(rows_pad * 4 * kSubCands * 8);
  L.normc = take(static_cast<size_t>(B) * N * 4);
  L.mu = take(static_cast<size_t>(B) * 128 * 4);
  L.maxc = take(static_cast<size_t>(B) * 4);
  L.tau = take(rows_pad * 4);
  L.cnt4 = take(rows_pad * 16);
  L.total = off + 1024;
  return L;
}

extern "C" size_t dgcnn_knn_tc_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return std::max(knn_tc_layout(B, N, 64).total, knn_tc_layout(B, N, 128).total);
}

extern "C" int dgcnn_knn_tc_supported(int C, int N, int k) { return (C >= 1 && C <= 128 && k <= 48 && N >= 128) ? 1 : 0; }

template <int KA, int KEYS>
static int launch_knn_tc(const float* rows, int ld, int B, int C4, int N, int k, int ordered, int32_t* idx, unsigned char* ws,
                         const KnnTcLayout& L, cudaStream_t st) {
  using S = KnnScan<KA, KEYS>;
  float* norms = reinterpret_cast<float*>(ws + L.norms);
  unsigned int* maxnorm = reinterpret_cast<unsigned int*>(ws + L.maxn);
  int* fb_count = reinterpret_cast<int*>(ws + L.cnt);
  int* fb_rows = reinterpret_cast<int*>(ws + L.fbrows);
  unsigned char* tiles = ws + L.tiles;
  float* normc = reinterpret_cast<float*>(ws + L.normc);
  float* mu = reinterpret_cast<float*>(ws + L.mu);
  unsigned int* maxnormc = reinterpret_cast<unsigned int*>(ws + L.maxc);
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(ws + L.cand);
  float* tau = reinterpret_cast<float*>(ws + L.tau);
  int* cnt4 = reinterpret_cast<int*>(ws + L.cnt4);
  const int T = L.T;

  knn_tc_mean_kernel<<<dim3(std::max(1, std::min(64, N / 128)), B), 256, 0, st>>>(rows, ld, N, C4, mu);
  DG_CUDA(cudaGetLastError());
  knn_tc_prep_kernel<KA, KEYS><<<dim3(std::min((T * KEYS * KA * 8 + 255) / 256, 4096), B), 256, 0, st>>>(rows, ld, N, C4, T, mu, tiles,
                                                                                                     normc, maxnormc);
  DG_CUDA(cudaGetLastError());
  // Threshold sampling.  R = target number of candidates per row; the pass-1 sample of S = T1*KEYS keys is split into 64
  // groups per row and tau is the r-th largest group maximum.  The r-th largest of the sample sits at global rank
  // ~ r' N / S with relative spread 1/sqrt(r'): r' ~ 24 with R = 3k keeps P(fewer than k candidates) ~ P(Poisson(8) >= 24)
  // ~ 1e-6 per row and a sub-buffer (mean R/4) more than 6 sigma below its 64 slots, so fallback rows are rare events.
  const int Tk = (N + KEYS - 1) / KEYS;
  const int R = std::max(90, 3 * k);
  int T1 = 0, rsel = 1;
  if (N > 4 * kSubCands) {
    T1 = static_cast<int>((24.0 * N / R) / KEYS + 0.5);
    T1 = std::max(1, std::min(Tk, T1));
    const double rp = static_cast<double>(R) * std::min(T1 * KEYS, N) / N;       // expected sample rank of the target
    const double r = rp - rp * rp / (2.0 * kGroups);                              // top sample values sharing a group
    rsel = std::max(2, std::min(40, static_cast<int>(r + 0.5)));
  }
  const int nk16_last = (C4 - (KA - 1) * 64 + 15) / 16;
  DG_CUDA(cudaFuncSetAttribute(knn_tc_scan_kernel<KA, KEYS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(S::bytes)));
  knn_tc_scan_kernel<KA, KEYS><<<dim3((N + 127) / 128, B), kTcThreads, S::bytes, st>>>(tiles, N, T, nk16_last, T1, rsel, cand, cnt4, tau);
  DG_CUDA(cudaGetLastError());
  if (ordered)
    knn_tc_select_kernel<true><<<dim3((N + 7) / 8, B), kThreads, 0, st>>>(rows, ld, norms, maxnorm, normc, maxnormc, N, C4, k, cand, cnt4,
                                                                          tau, idx, fb_count, fb_rows);
  else
    knn_tc_select_kernel<false><<<dim3((N + 7) / 8, B), kThreads, 0, st>>>(rows, ld, norms, maxnorm, normc, maxnormc, N, C4, k, cand, cnt4,
                                                                           tau, idx, fb_count, fb_rows);
  DG_CUDA(cudaGetLastError());
  knn_fallback_kernel<<<128, kThreads, 0, st>>>(rows, ld, norms, N, C4, k, idx, fb_count, fb_rows);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_knn_tc_f32(const float* rows, int ld, int B, int C, int N, int k, int ordered, int32_t* idx, void* workspace,
                                size_t workspace_bytes, void* stream) {
  DG_REQUIRE(rows && idx && workspace, DGCNN_E_INVALID, "knn_tc: null pointer");
  DG_REQUIRE(B > 0 && N > 0 && C > 0 && B <= 65535, DGCNN_E_INVALID, "knn_tc: bad shape B=%d C=%d N=%d", B, C, N);
  DG_REQUIRE(k >= 1 && k <= N, DGCNN_E_INVALID, "knn_tc: selected index k out of range (k=%d, N=%d)", k, N);
  DG_REQUIRE(dgcnn_knn_tc_supported(C, N, k), DGCNN_E_UNSUPPORTED,
             "knn_tc: needs C <= 128, k <= 48, N >= 128 (C=%d N=%d k=%d); use dgcnn_knn_f32", C, N, k);
  const int C4 = (C + 3) / 4 * 4;
  DG_REQUIRE(ld >= C4 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(rows) & 15) == 0, DGCNN_E_INVALID,
             "knn_tc: rows must be 16-byte aligned with ld %% 4 == 0 and ld >= round4(C)");
  DG_REQUIRE(workspace_bytes >= dgcnn_knn_tc_workspace_bytes(B, N), DGCNN_E_WORKSPACE, "knn_tc: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const KnnTcLayout L = knn_tc_layout(B, N, C);
  unsigned char* ws = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  // maxnorm + fallback counters, then centred max norm + channel sums
  DG_CUDA(cudaMemsetAsync(ws + L.maxn, 0, L.fbrows - L.maxn, st));
  DG_CUDA(cudaMemsetAsync(ws + L.mu, 0, L.tau - L.mu, st));
  knn_norms_max_kernel<<<dim3((N + 127) / 128, B), 128, 0, st>>>(rows, ld, N, C, reinterpret_cast<float*>(ws + L.norms),
                                                                 reinterpret_cast<unsigned int*>(ws + L.maxn));
  DG_CUDA(cudaGetLastError());
  if (C > 64) return launch_knn_tc<2, 64>(rows, ld, B, C4, N, k, ordered, idx, ws, L, st);
  return launch_knn_tc<1, 128>(rows, ld, B, C4, N, k, ordered, idx, ws, L, st);
}

/* diagnostics: copies the 5 fallback counters of the last dgcnn_knn_tc_f32 call on this workspace to host memory
 * (rows redone by brute force | sub-buffer overflow | too few candidates | too many near-ties | certificate failed).
 * Synchronises the stream. */
extern "C" int dgcnn_knn_tc_fallback_stats(const void* workspace, int B, int N, int C, int* out5, void* stream) {
  DG_REQUIRE(workspace && out5, DGCNN_E_INVALID, "knn_tc_fallback_stats: null pointer");
  const KnnTcLayout L = knn_tc_layout(B, N, C);
  const unsigned char* ws = reinterpret_cast<const unsigned char*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  DG_CUDA(cudaMemcpyAsync(out5, ws + L.cnt, 5 * sizeof(int), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  DG_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  return DGCNN_OK;
}


// ---- low-dimensional exact kNN (C <= 16), threshold-collect-select ------------------------------------------------
static size_t knn_small_layout(int B, int N, size_t* o_norms, size_t* o_cnt, size_t* o_rows, size_t* o_cand) {
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 1023) & ~size_t(1023); return o; };
  *o_norms = take(static_cast<size_t>(B) * N * 4);
  *o_cnt = take(4 + static_cast<size_t>(B) * 4);       // fallback counter | per-block max norm (unused by this path)
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
  DG_CUDA(cudaMemsetAsync(fb_count, 0, 4 + static_cast<size_t>(B) * 4, st));
  // norms in torch.sum's order (rows beyond C are zero padded, so the padded chain is exact as well)
  knn_norms_max_kernel<<<dim3((N + 127) / 128, B), 128, 0, st>>>(rows, ld, N, C, norms, reinterpret_cast<unsigned int*>(fb_count + 1));
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
