// knn.cu — fused exact-FP32 kNN graph construction (replaces model.py:25-40) + layout helpers.
//
// The B x N x N score matrix is never written to HBM: each CTA owns TQ query rows of one block,
// streams the block's points through shared memory in 64-key tiles, forms the score tile with an
// FP32 FMA-chain GEMM (bit-identical to the reference's CPU sgemm) and folds it into a
// per-query sorted top-k list kept in shared memory (warp-cooperative insertion via shuffles).
#include "common.cuh"

namespace dg {

// ---------------------------------------------------------------------------------------------
// layout helpers
// ---------------------------------------------------------------------------------------------
__global__ void pack_rows_kernel(const float* __restrict__ x, long long sb, long long sc, long long sn,
                                 int B, int C, int C4, int N, float* __restrict__ rows, int ld) {
  const long long total = static_cast<long long>(B) * N * C4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C4);
    const long long p = i / C4;
    const int n = static_cast<int>(p % N);
    const int b = static_cast<int>(p / N);
    rows[p * ld + c] = c < C ? x[b * sb + c * sc + n * sn] : 0.f;
  }
}

__global__ void unpack_rows_kernel(const float* __restrict__ rows, int ld, int B, int C, int N,
                                   float* __restrict__ y) {
  // 32x32 smem transpose: coalesced reads along channels, coalesced writes along points.
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int n = n0 + r, c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (n < N && c < C) ? rows[(static_cast<size_t>(b) * N + n) * ld + c] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, n = n0 + threadIdx.x;
    if (c < C && n < N) y[(static_cast<size_t>(b) * C + c) * N + n] = tile[threadIdx.x][r];
  }
}

// |x|^2 per point in torch.sum's order for a non-innermost reduction (aten SumKernel multi_row_sum):
// squares rounded first, sequential accumulation inside blocks of 16 channels, block sums accumulated
// sequentially, remainder added first.  For C <= 16 this is the plain left-to-right sum.
__global__ void norms_kernel(const float* __restrict__ rows, int ld, int M, int C, float* __restrict__ norms) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= M) return;
  const float* r = rows + static_cast<size_t>(p) * ld;
  float lvl1 = 0.f;
  int c = 0;
  for (; c + 16 <= C; c += 16) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float v = r[c + j];
      a = __fadd_rn(a, __fmul_rn(v, v));
    }
    lvl1 = __fadd_rn(lvl1, a);
  }
  float a = 0.f;
  for (; c < C; ++c) {
    const float v = r[c];
    a = __fadd_rn(a, __fmul_rn(v, v));
  }
  norms[p] = __fadd_rn(a, lvl1);
}

__global__ void graph_feature_kernel(const float* __restrict__ rows, int ld, const int32_t* __restrict__ idx,
                                     int N, int C, int k, long long total, float* __restrict__ out) {
  // out[p][slot][0:C) = x_j - x_i ; [C:2C) = x_i        (model.py:74-81)
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c2 = static_cast<int>(i % (2 * C));
    const long long e = i / (2 * C);
    const long long p = e / k;
    const float xi = rows[p * ld + (c2 < C ? c2 : c2 - C)];
    if (c2 < C) {
      const long long base = (p / N) * N;
      const long long j = base + idx[e];
      out[i] = __fsub_rn(rows[j * ld + c2], xi);
    } else {
      out[i] = xi;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// fused exact kNN
// ---------------------------------------------------------------------------------------------
template <int LPL>
__device__ __forceinline__ unsigned long long list_at(const unsigned long long (&lst)[LPL], int pos) {
  // value at sorted position pos = r*32 + lane (warp-uniform pos)
  unsigned long long v = 0;
#pragma unroll
  for (int r = 0; r < LPL; ++r) {
    const unsigned long long w = __shfl_sync(0xffffffffu, lst[r], pos & 31);
    if (r == (pos >> 5)) v = w;
  }
  return v;
}

template <int LPL>
__device__ __forceinline__ void list_insert(unsigned long long (&lst)[LPL], unsigned long long c, int lane) {
  // The list is sorted descending over positions r*32+lane; keys are unique (index in low bits).
  unsigned long long prev[LPL];
#pragma unroll
  for (int r = 0; r < LPL; ++r) {
    const unsigned long long up = __shfl_up_sync(0xffffffffu, lst[r], 1);
    unsigned long long wrap = ~0ull;
    if (r > 0) wrap = __shfl_sync(0xffffffffu, lst[r > 0 ? r - 1 : 0], 31);
    prev[r] = lane == 0 ? wrap : up;
  }
#pragma unroll
  for (int r = 0; r < LPL; ++r) lst[r] = lst[r] >= c ? lst[r] : (prev[r] >= c ? c : prev[r]);
}

constexpr int kKeyTile = 64;
constexpr int kSld = kKeyTile + 1;

template <int CP, int LPL, int TQ>
struct KnnSmem {
  static constexpr int LDA = CP + 4;
  static constexpr size_t a_floats = static_cast<size_t>(TQ) * LDA;
  static constexpr size_t b_floats = static_cast<size_t>(kKeyTile) * LDA;
  static constexpr size_t s_floats = static_cast<size_t>(TQ) * kSld;
  static constexpr size_t n_floats = TQ + kKeyTile;
  static constexpr size_t list_bytes = static_cast<size_t>(TQ) * 32 * LPL * 8;
  static constexpr size_t float_bytes = (a_floats + b_floats + s_floats + n_floats) * 4;
  static constexpr size_t bytes = ((float_bytes + 15) / 16) * 16 + list_bytes;
};

template <int CP, int LPL, int TQ>
__global__ void __launch_bounds__(kThreads) knn_kernel(const float* __restrict__ rows, int ld,
                                                       const float* __restrict__ norms, int N, int C4, int k,
                                                       int32_t* __restrict__ idx) {
  using SM = KnnSmem<CP, LPL, TQ>;
  constexpr int LDA = SM::LDA;
  constexpr int RI = TQ / 16;  // rows per thread in the GEMM
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* As = reinterpret_cast<float*>(smem_raw);
  float* Bs = As + SM::a_floats;
  float* Ss = Bs + SM::b_floats;
  float* qn = Ss + SM::s_floats;
  float* kn = qn + TQ;
  unsigned long long* Ls = reinterpret_cast<unsigned long long*>(smem_raw + ((SM::float_bytes + 15) / 16) * 16);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.y, q0 = blockIdx.x * TQ;
  const float* X = rows + static_cast<size_t>(b) * N * ld;
  const float* nb = norms + static_cast<size_t>(b) * N;

  // query tile, query norms, empty lists
  for (int i = tid; i < TQ * (CP / 4); i += kThreads) {
    const int r = i / (CP / 4), c4 = i % (CP / 4), q = q0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < N && 4 * c4 < C4) v = *reinterpret_cast<const float4*>(X + static_cast<size_t>(q) * ld + 4 * c4);
    *reinterpret_cast<float4*>(As + r * LDA + 4 * c4) = v;
  }
  for (int r = tid; r < TQ; r += kThreads) qn[r] = (q0 + r < N) ? nb[q0 + r] : 0.f;
  for (int i = tid; i < TQ * 32 * LPL; i += kThreads) Ls[i] = 0ull;

  for (int t0 = 0; t0 < N; t0 += kKeyTile) {
    __syncthreads();  // previous tile's Bs / Ss consumers are done (also covers the prologue stores)
    for (int i = tid; i < kKeyTile * (CP / 4); i += kThreads) {
      const int r = i / (CP / 4), c4 = i % (CP / 4), j = t0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < N && 4 * c4 < C4) v = *reinterpret_cast<const float4*>(X + static_cast<size_t>(j) * ld + 4 * c4);
      *reinterpret_cast<float4*>(Bs + r * LDA + 4 * c4) = v;
    }
    if (tid < kKeyTile) kn[tid] = (t0 + tid < N) ? nb[t0 + tid] : 0.f;
    __syncthreads();

    // score tile: dot via FMA chain in channel order, then the reference's two rounded subtractions
    float acc[RI][4];
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
#pragma unroll 2
    for (int kk = 0; kk < CP; kk += 4) {
      float4 bv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) bv[q] = *reinterpret_cast<const float4*>(Bs + (tx + 16 * q) * LDA + kk);
#pragma unroll
      for (int i = 0; i < RI; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(As + (ty + 16 * i) * LDA + kk);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float c = acc[i][q];
          c = __fmaf_rn(a.x, bv[q].x, c);
          c = __fmaf_rn(a.y, bv[q].y, c);
          c = __fmaf_rn(a.z, bv[q].z, c);
          c = __fmaf_rn(a.w, bv[q].w, c);
          acc[i][q] = c;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int row = ty + 16 * i, col = tx + 16 * q;
        // pd = (-|x_j|^2 - (-2 dot)) - |x_i|^2   (model.py:37): one rounding per step
        Ss[row * kSld + col] = __fsub_rn(__fmaf_rn(acc[i][q], 2.0f, -kn[col]), qn[row]);
      }
    __syncthreads();

    // fold the tile into the per-query lists: warp `warp` owns TQ/8 query rows
    for (int rr = 0; rr < TQ / 8; ++rr) {
      const int row = warp * (TQ / 8) + rr;
      if (q0 + row >= N) break;
      unsigned long long lst[LPL];
#pragma unroll
      for (int r = 0; r < LPL; ++r) lst[r] = Ls[(row * LPL + r) * 32 + lane];
      unsigned long long thr = list_at<LPL>(lst, k - 1);
      bool changed = false;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int col = lane + 32 * half, j = t0 + col;
        const unsigned long long c = (j < N) ? make_key(Ss[row * kSld + col], j) : 0ull;
        unsigned m = __ballot_sync(0xffffffffu, c > thr);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const unsigned long long cc = __shfl_sync(0xffffffffu, c, src);
          if (cc > thr) {
            list_insert<LPL>(lst, cc, lane);
            thr = list_at<LPL>(lst, k - 1);
            changed = true;
          }
        }
      }
      if (changed) {
#pragma unroll
        for (int r = 0; r < LPL; ++r) Ls[(row * LPL + r) * 32 + lane] = lst[r];
      }
    }
  }
  __syncthreads();
  // emit indices (sorted best-first; callers treat them as a set)
  for (int i = tid; i < TQ * k; i += kThreads) {
    const int row = i / k, pos = i % k, q = q0 + row;
    if (q < N) {
      const unsigned long long key = Ls[(row * LPL + (pos >> 5)) * 32 + (pos & 31)];
      idx[(static_cast<size_t>(b) * N + q) * k + pos] = key_index(key);
    }
  }
}

template <int CP, int LPL, int TQ>
static int launch_knn(const float* rows, int ld, const float* norms, int B, int N, int C4, int k, int32_t* idx,
                      cudaStream_t st) {
  using SM = KnnSmem<CP, LPL, TQ>;
  auto kern = knn_kernel<CP, LPL, TQ>;
  DG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(SM::bytes)));
  dim3 grid((N + TQ - 1) / TQ, B);
  kern<<<grid, kThreads, SM::bytes, st>>>(rows, ld, norms, N, C4, k, idx);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

template <int CP, int LPL>
static int dispatch_tq(const float* rows, int ld, const float* norms, int B, int N, int C4, int k, int32_t* idx,
                       cudaStream_t st) {
  if (KnnSmem<CP, LPL, 128>::bytes <= 200 * 1024) return launch_knn<CP, LPL, 128>(rows, ld, norms, B, N, C4, k, idx, st);
  return launch_knn<CP, LPL, 64>(rows, ld, norms, B, N, C4, k, idx, st);
}

template <int CP>
static int dispatch_lpl(const float* rows, int ld, const float* norms, int B, int N, int C4, int k, int32_t* idx,
                        cudaStream_t st) {
  if (k <= 32) return dispatch_tq<CP, 1>(rows, ld, norms, B, N, C4, k, idx, st);
  if (k <= 64) return dispatch_tq<CP, 2>(rows, ld, norms, B, N, C4, k, idx, st);
  return dispatch_tq<CP, 4>(rows, ld, norms, B, N, C4, k, idx, st);
}

}  // namespace dg

using namespace dg;

extern "C" int dgcnn_pack_rows_f32(const float* x, int64_t sb, int64_t sc, int64_t sn, int B, int C, int N,
                                   float* rows, int ld, void* stream) {
  DG_REQUIRE(x && rows, DGCNN_E_INVALID, "pack_rows: null pointer");
  const int C4 = (C + 3) / 4 * 4;
  DG_REQUIRE(B > 0 && C > 0 && N > 0 && ld >= C4 && ld % 4 == 0, DGCNN_E_INVALID,
             "pack_rows: need B,C,N > 0 and ld >= round4(C), ld %% 4 == 0 (B=%d C=%d N=%d ld=%d)", B, C, N, ld);
  const long long total = static_cast<long long>(B) * N * C4;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 16));
  pack_rows_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, sb, sc, sn, B, C, C4, N, rows, ld);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_unpack_rows_f32(const float* rows, int ld, int B, int C, int N, float* y, void* stream) {
  DG_REQUIRE(rows && y, DGCNN_E_INVALID, "unpack_rows: null pointer");
  DG_REQUIRE(B > 0 && C > 0 && N > 0 && ld >= C && B <= 65535, DGCNN_E_INVALID, "unpack_rows: bad shape");
  dim3 grid((N + 31) / 32, (C + 31) / 32, B), block(32, 8);
  unpack_rows_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(rows, ld, B, C, N, y);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_knn_f32(const float* rows, int ld, int B, int C, int N, int k, int32_t* idx, float* norms,
                             void* stream) {
  DG_REQUIRE(rows && idx && norms, DGCNN_E_INVALID, "knn: null pointer");
  DG_REQUIRE(B > 0 && N > 0 && C > 0 && B <= 65535, DGCNN_E_INVALID, "knn: bad shape B=%d C=%d N=%d", B, C, N);
  // the reference raises "selected index k out of range" for k > N
  DG_REQUIRE(k >= 1 && k <= N, DGCNN_E_INVALID, "knn: selected index k out of range (k=%d, N=%d)", k, N);
  DG_REQUIRE(C <= 128 && k <= 128, DGCNN_E_UNSUPPORTED, "knn: C=%d (max 128) / k=%d (max 128) not compiled", C, k);
  const int C4 = (C + 3) / 4 * 4;
  DG_REQUIRE(ld >= C4 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(rows) & 15) == 0, DGCNN_E_INVALID,
             "knn: rows must be 16-byte aligned with ld %% 4 == 0 and ld >= round4(C)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int M = B * N;
  norms_kernel<<<(M + 255) / 256, 256, 0, st>>>(rows, ld, M, C, norms);
  DG_CUDA(cudaGetLastError());
  if (C <= 4) return dispatch_lpl<4>(rows, ld, norms, B, N, C4, k, idx, st);
  if (C <= 8) return dispatch_lpl<8>(rows, ld, norms, B, N, C4, k, idx, st);
  if (C <= 16) return dispatch_lpl<16>(rows, ld, norms, B, N, C4, k, idx, st);
  if (C <= 32) return dispatch_lpl<32>(rows, ld, norms, B, N, C4, k, idx, st);
  if (C <= 64) return dispatch_lpl<64>(rows, ld, norms, B, N, C4, k, idx, st);
  return dispatch_lpl<128>(rows, ld, norms, B, N, C4, k, idx, st);
}

extern "C" int dgcnn_graph_feature_f32(const float* rows, int ld, const int32_t* idx, int B, int C, int N, int k,
                                       float* out, void* stream) {
  DG_REQUIRE(rows && idx && out, DGCNN_E_INVALID, "graph_feature: null pointer");
  DG_REQUIRE(B > 0 && C > 0 && N > 0 && k > 0 && ld >= C, DGCNN_E_INVALID, "graph_feature: bad shape");
  const long long total = static_cast<long long>(B) * N * k * 2 * C;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 32));
  graph_feature_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(rows, ld, idx, N, C, k, total, out);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}
