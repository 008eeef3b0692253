// knn_tc.cu — feature-space kNN (16 < C <= 64) on the tensor cores, with EXACT results.
//
// Gram tiles are computed by tcgen05.mma on a 2-plane bf16 split of the features (3 MMAs per product,
// |error| <= ~2e-5 * |x_i||x_j|) into TMEM; the approximate scores only act as a FILTER:
//   * every query keeps its 64 best approximate candidates (streaming: per-row candidate buffers + warp-level
//     bitonic merges, select.cuh);
//   * with eps_i a rigorous bound on |approx - exact| for row i, every true top-k neighbour has
//     approx >= t_k - 2 eps_i (t_k = k-th best approximate score), so if the 64-th candidate is below that
//     cut the candidate list provably contains the exact top-k ("certificate");
//   * the survivors are re-scored with the reference's exact FP32 arithmetic (FMA chain in channel order,
//     two rounded subtractions, model.py:35-37) and the top-k of the exact scores is emitted;
//   * rows whose certificate fails (more than 64-k near-ties, e.g. many duplicated points) are redone by an
//     exact brute-force kernel.  The neighbour table is therefore identical to dgcnn_knn_f32's, bit for bit.
// The B x N x N matrix never leaves the SM: scores go TMEM -> registers -> shared memory -> selection.
#include "common.cuh"
#include "tc.cuh"
#include "select.cuh"

namespace dg {

constexpr int kTcKeyTile = 64;          // keys per MMA tile (UMMA N)
constexpr uint32_t kTileBytes = 2 * tc::kUnitBytes + kTcKeyTile * 4;   // plane0 | plane1 | 64 key norms
constexpr uint32_t kStageBytes = 17 * 1024;                             // smem slot (planes stay 1024-aligned)
constexpr int kTcStages = 3;
// error model of the filter (see DESIGN.md "kNN filter bound")
constexpr float kC1 = 2.0e-5f * 2.0f;   // 2 * (split + accumulation + chain error) on 2*dot, relative to |x_i||x_j|
constexpr float kC2 = 4.8e-7f;          // final roundings, relative to (|x_i|+|x_j|)^2

// ---------------------------------------------------------------------------------------------
// prep: rows -> per-block tiles of 64 points: two swizzled bf16 planes + the 64 squared norms; also the squared
// norms themselves (torch.sum order, see knn.cu) and the per-block maximum norm.
// ---------------------------------------------------------------------------------------------
__global__ void knn_tc_prep_kernel(const float* __restrict__ rows, int ld, int N, int C4, int T,
                                   const float* __restrict__ norms, unsigned char* __restrict__ tiles) {
  // one thread per (tile row, 16-byte chunk); grid.y = block
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
      const float4 lo = *reinterpret_cast<const float4*>(src);
      v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
      if (ch * 8 + 4 < C4) {
        const float4 hi = *reinterpret_cast<const float4*>(src + 4);
        v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
      }
    }
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
    if (ch == 0) reinterpret_cast<float*>(tile + 2 * tc::kUnitBytes)[rl] = n < N ? norms[static_cast<size_t>(b) * N + n] : 0.f;
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
__device__ __forceinline__ float exact_score(const float* __restrict__ xi, const float* __restrict__ xj, int C4,
                                             float norm_i, float norm_j) {
  float d = 0.f;
  for (int c = 0; c < C4; c += 4) {
    const float4 a = *reinterpret_cast<const float4*>(xi + c);
    const float4 b = *reinterpret_cast<const float4*>(xj + c);
    d = __fmaf_rn(a.x, b.x, d);
    d = __fmaf_rn(a.y, b.y, d);
    d = __fmaf_rn(a.z, b.z, d);
    d = __fmaf_rn(a.w, b.w, d);
  }
  return __fsub_rn(__fmaf_rn(d, 2.0f, -norm_j), norm_i);
}

struct KnnTcSmem {
  static constexpr uint32_t a_bytes = 2 * 128 * 128;                        // two planes of 128 query rows
  static constexpr uint32_t b_off = a_bytes;
  static constexpr uint32_t sel_off = b_off + kTcStages * kStageBytes;
  static constexpr uint32_t bar_off = sel_off + static_cast<uint32_t>(sel::kSmemBytes);
  static constexpr uint32_t bytes = bar_off + 128;
};

// 512 threads.  Scan: warp w reads TMEM lane quarter (w & 3) and column quarter (w >> 2); a thread = one query row x
// 16 keys per tile, scores stay in registers.  Compaction / finalisation: warp w owns rows 8w .. 8w+7.
constexpr int kTcThreads = 512;
__global__ void __launch_bounds__(kTcThreads, 1)
knn_tc_kernel(const unsigned char* __restrict__ tiles, const float* __restrict__ rows, int ld,
              const float* __restrict__ norms, const unsigned int* __restrict__ maxnorm_bits, int N, int C4, int T,
              int k, int kp, int kinit, int32_t* __restrict__ idx, int* __restrict__ fb_count, int* __restrict__ fb_rows) {
  // no static shared memory in this kernel: the dynamic window starts 1024-byte aligned (checked below), and plain
  // pointer arithmetic on it keeps the shared address space visible to the compiler (LDS/STS instead of generic LD/ST)
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* As = sm;
  unsigned char* Bs = sm + KnnTcSmem::b_off;
  const sel::Smem S = sel::carve(sm + KnnTcSmem::sel_off);
  uint64_t* abar = reinterpret_cast<uint64_t*>(sm + KnnTcSmem::bar_off);
  uint64_t* full = abar + 1;               // [kTcStages]
  uint64_t* accfull = full + kTcStages;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfull + 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y, qt = blockIdx.x, q0 = qt * 128;
  const int Tk = (N + kTcKeyTile - 1) / kTcKeyTile;        // key tiles to visit
  const unsigned char* btiles = tiles + static_cast<size_t>(b) * T * kTileBytes;

  if (tid == 0) {
    if (tc::smem_u32(sm) & 1023u) __trap();
    tc::mbar_init(abar, 1);
    for (int s = 0; s < kTcStages; ++s) tc::mbar_init(full + s, 1);
    tc::mbar_init(accfull, 1);
    tc::mbar_init(accfull + 1, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 128);            // two 64-column accumulators
  sel::init(S, tid, kTcThreads);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t taddr = *tmem_slot;
  constexpr uint32_t idesc = tc::idesc_bf16(128, kTcKeyTile);

  if (tid == 0) {
    // queries: the two 64-row tiles 2*qt, 2*qt+1 -> planes of 128 rows
    tc::mbar_arrive_expect_tx(abar, KnnTcSmem::a_bytes);
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int p = 0; p < 2; ++p)
        tc::bulk_g2s(As + p * 16384 + h * tc::kUnitBytes, btiles + static_cast<size_t>(2 * qt + h) * kTileBytes + p * tc::kUnitBytes,
                     tc::kUnitBytes, abar);
    for (int t = 0; t < kTcStages && t < Tk; ++t) {
      tc::mbar_arrive_expect_tx(full + t, kTileBytes);
      tc::bulk_g2s(Bs + t * kStageBytes, btiles + static_cast<size_t>(t) * kTileBytes, kTileBytes, full + t);
    }
    tc::mbar_wait(abar, 0);
    tc::mbar_wait(full, 0);
    tc::fence_after_sync();
    tc::mma_split_katom<2>(taddr, tc::smem_u32(As), 16384, tc::smem_u32(Bs), tc::kUnitBytes, idesc, true);
    tc::mma_commit(accfull);
  }
  __syncwarp();

  const int quarter = warp & 3, cq = warp >> 2;
  const int erow = quarter * 32 + lane;                     // scan: this thread's query row
  const bool row_ok = q0 + erow < N;

  for (int t = 0; t < Tk; ++t) {
    if (tid == 0 && t + 1 < Tk) {                           // keep the tensor pipe one tile ahead of the selection
      const int s1 = (t + 1) % kTcStages;
      tc::mbar_wait(full + s1, ((t + 1) / kTcStages) & 1);
      tc::fence_after_sync();
      tc::mma_split_katom<2>(taddr + ((t + 1) & 1) * kTcKeyTile, tc::smem_u32(As), 16384, tc::smem_u32(Bs + s1 * kStageBytes),
                             tc::kUnitBytes, idesc, true);
      tc::mma_commit(accfull + ((t + 1) & 1));
    }
    __syncwarp();
    tc::mbar_wait(accfull + (t & 1), (t >> 1) & 1);
    tc::fence_after_sync();
    {
      float v[16];
      tc::tmem_ld16(taddr + (static_cast<uint32_t>(quarter * 32) << 16) + (t & 1) * kTcKeyTile + cq * 16, v);
      const float* kn = reinterpret_cast<const float*>(Bs + (t % kTcStages) * kStageBytes + 2 * tc::kUnitBytes) + cq * 16;
      const float thr = S.thr[erow];
      int n = S.cnt[erow * sel::kSub + cq];
      const int j0 = t * kTcKeyTile + cq * 16;
      if (t < Tk - 1) {                                     // every key of the tile exists
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float sc = __fmaf_rn(v[i], 2.0f, -kn[i]); // approx of 2 x_i.x_j - |x_j|^2 (|x_i|^2 is constant per row)
            if (sc > thr) sel::push(S, erow, cq, n, sc, j0 + i);
          }
        }
      } else {
        const int jn = row_ok ? N - j0 : 0;                 // keys j0+i with i < jn exist
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float sc = __fmaf_rn(v[i], 2.0f, -kn[i]);
          if (sc > thr && i < jn) sel::push(S, erow, cq, n, sc, j0 + i);
        }
      }
      S.cnt[erow * sel::kSub + cq] = n;
      // A sub-buffer (16 slots) receives at most 16 keys per tile, each with probability ~kp/(keys seen): compact early
      // and often while that probability is high so that overflow (-> exact fallback for the row) stays a non-event.
      const int trigger = (t == Tk - 1 || t == 0) ? 1 : (t < 32 ? 4 : 8);
      sel::request_compaction(S, erow, n, trigger, t & 1);
    }
    tc::fence_before_sync();
    __syncthreads();                                        // pushes visible; tile t's smem slot and accumulator are free
    if (tid == 0 && t + kTcStages < Tk) {
      const int s = t % kTcStages;
      tc::mbar_arrive_expect_tx(full + s, kTileBytes);
      tc::bulk_g2s(Bs + s * kStageBytes, btiles + static_cast<size_t>(t + kTcStages) * kTileBytes, kTileBytes, full + s);
    }
    __syncwarp();
    // Tile 0 bootstraps the threshold: the kinit-th best of the first 64 keys has an expected global rank of
    // ~kinit*N/65 >= 4*kp, which removes the initial flood of candidates; if it turns out tighter than the kp-th best
    // overall the certificate below fails and the row goes to the exact fallback (probability < 1e-3).
    if (tid == 0) S.nwork[(t + 1) & 1] = 0;
    sel::compact_queued(S, warp, kTcThreads / 32, lane, t == 0 ? kinit : kp, t & 1);
    __syncthreads();                                        // thresholds / buffers consistent for the next scan
  }

  // ---- certificate + exact re-scoring + exact top-k (warp per query row) ---------------------------------
  const float rmax = sqrtf(__uint_as_float(maxnorm_bits[b]));
  for (int rr = 0; rr < 8; ++rr) {
    const int row = warp * 8 + rr, q = q0 + row;
    if (q >= N) break;
    const sel::u64 l0 = S.lists[(row * 2 + 0) * 32 + lane];
    const sel::u64 l1 = S.lists[(row * 2 + 1) * 32 + lane];
    const size_t g = static_cast<size_t>(b) * N + q;
    const float norm_i = norms[g];
    const float ni = sqrtf(norm_i);
    const float eps = 1.01f * (kC1 * ni * rmax + kC2 * (ni + rmax) * (ni + rmax));
    const float tk = sel::key_score(sel::list_at(l0, l1, k - 1));
    const float cut = tk - 2.0f * eps;
    // every key whose approximate score exceeds the final pass threshold is in the list, so the list is complete
    // above `cut` iff that threshold is strictly below the cut
    const bool uncertified = !(S.thr[row] < cut) || S.overflow[row] != 0;
    const float* xi = rows + g * ld;
    sel::u64 ex[2];
    const sel::u64 ls[2] = {l0, l1};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      ex[r] = 0ull;
      if (ls[r] != 0ull && sel::key_score(ls[r]) >= cut) {
        const int j = key_index(ls[r]);
        const size_t gj = static_cast<size_t>(b) * N + j;
        ex[r] = make_key(exact_score(xi, rows + gj * ld, C4, norm_i, norms[gj]), j);
      }
    }
    int rank0 = 0, rank1 = 0;
#pragma unroll 4
    for (int s = 0; s < 32; ++s) {
      const sel::u64 o0 = __shfl_sync(0xffffffffu, ex[0], s);
      const sel::u64 o1 = __shfl_sync(0xffffffffu, ex[1], s);
      rank0 += (o0 > ex[0]) + (o1 > ex[0]);
      rank1 += (o0 > ex[1]) + (o1 > ex[1]);
    }
    if (ex[0] != 0ull && rank0 < k) idx[g * k + rank0] = key_index(ex[0]);
    if (ex[1] != 0ull && rank1 < k) idx[g * k + rank1] = key_index(ex[1]);
    if (uncertified && lane == 0) fb_rows[atomicAdd(fb_count, 1)] = static_cast<int>(g);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(taddr, 128);
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
  __shared__ unsigned long long wl[8][sel::kList];
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

}  // namespace dg

using namespace dg;

static size_t knn_tc_layout(int B, int N, size_t* o_norms, size_t* o_max, size_t* o_cnt, size_t* o_rows, size_t* o_tiles) {
  const int T = 2 * ((N + 127) / 128);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 1023) & ~size_t(1023); return o; };
  *o_norms = take(static_cast<size_t>(B) * N * 4);
  *o_max = take(static_cast<size_t>(B) * 4);
  *o_cnt = take(4);
  *o_rows = take(static_cast<size_t>(B) * N * 4);
  *o_tiles = take(static_cast<size_t>(B) * T * kTileBytes);
  return off + 1024;
}

extern "C" size_t dgcnn_knn_tc_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  size_t a, b, c, d, e;
  return knn_tc_layout(B, N, &a, &b, &c, &d, &e);
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
  size_t o_norms, o_max, o_cnt, o_rows, o_tiles;
  knn_tc_layout(B, N, &o_norms, &o_max, &o_cnt, &o_rows, &o_tiles);
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
  knn_tc_prep_kernel<<<dim3(std::min((T * 512 + 255) / 256, 4096), B), 256, 0, st>>>(rows, ld, N, C4, T, norms, tiles);
  DG_CUDA(cudaGetLastError());
  DG_CUDA(cudaFuncSetAttribute(knn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(KnnTcSmem::bytes)));
  const int kp = std::min(sel::kList, k + 8);     // list position that defines the pass threshold / certificate
  const int kinit = std::max(4, std::min(kp, (4 * kp * 65 + N - 1) / N));
  knn_tc_kernel<<<dim3((N + 127) / 128, B), kTcThreads, KnnTcSmem::bytes, st>>>(tiles, rows, ld, norms, maxnorm, N, C4, T, k, kp,
                                                                               kinit, idx, fb_count, fb_rows);
  DG_CUDA(cudaGetLastError());
  knn_fallback_kernel<<<64, kThreads, 0, st>>>(rows, ld, norms, N, C4, k, idx, fb_count, fb_rows);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}
