// gemm_tc.cu — FP32-accurate GEMM on the 5th-gen tensor cores:  D[M,N] = A[M,K] * B[N,K]^T  (FP32 in/out).
//
// Operands are pre-split into bf16 planes in the swizzled unit format of tc.cuh (split_planes_kernel), staged
// with bulk async copies through an mbarrier ring, multiplied with tcgen05.mma (accumulator in TMEM) by one
// elected thread, and read back with tcgen05.ld by four epilogue warps.  Used for the dense GEMMs of the
// shared MLP / segmentation head (model.py:167-193) and as the validation vehicle for the tcgen05 primitives
// that the tensor-core kNN and EdgeConv kernels build on.
#include "common.cuh"
#include "tc.cuh"
#include "debug_lab.h"

namespace dg {

// ---------------------------------------------------------------------------------------------
// FP32 [R,K] (element strides sr, sc) -> NS bf16 planes in 64x64 swizzled units, zero padded to R_pad x K_pad.
// One thread produces one 16-byte chunk (8 consecutive k) of every plane.
// ---------------------------------------------------------------------------------------------
template <int NS>
__global__ void split_planes_kernel(const float* __restrict__ X, long long sr, long long sc, int R, int K, int R_pad,
                                    int KA, unsigned char* __restrict__ out) {
  const long long total = static_cast<long long>(R_pad) * KA * 8;   // chunks
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // consecutive threads walk the 8 chunks of a row, then the rows of a unit: 128-byte coalesced stores
    const int ch = static_cast<int>(i & 7);
    const long long t = i >> 3;
    const int rl = static_cast<int>(t & 63);
    const long long u = t >> 6;              // unit index = rt * KA + ka
    const int ka = static_cast<int>(u % KA);
    const int rt = static_cast<int>(u / KA);
    const int r = rt * 64 + rl, c0 = ka * 64 + ch * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = c0 + e;
      v[e] = (r < R && c < K) ? X[r * sr + c * sc] : 0.f;
    }
    unsigned char* unit = out + static_cast<size_t>(u) * NS * tc::kUnitBytes;
    const uint32_t off = tc::unit_elem_offset(rl, ch * 8);
#pragma unroll
    for (int p = 0; p < NS; ++p) {
      __nv_bfloat16 h[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        h[e] = __float2bfloat16_rn(v[e]);
        v[e] -= __bfloat162float(h[e]);       // exact in FP32
      }
      *reinterpret_cast<uint4*>(unit + static_cast<size_t>(p) * tc::kUnitBytes + off) = *reinterpret_cast<const uint4*>(h);
    }
  }
}

// Same output format for inputs whose contraction index is the SLOW memory dimension (sr == 1, e.g. A = dY^T for a
// weight gradient): one CTA per 64x64 unit, coalesced loads along the rows, transposed through shared memory.
template <int NS>
__global__ void __launch_bounds__(256) split_planes_t_kernel(const float* __restrict__ X, long long sr, long long sc,
                                                             int R, int K, int KA, unsigned char* __restrict__ out) {
  __shared__ float tile[64][65];                 // [c][r]
  const int u = blockIdx.x, ka = u % KA, rt = u / KA;
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int rl = tid & 63, cl = (tid >> 6) + 4 * i;
    const int r = rt * 64 + rl, c = ka * 64 + cl;
    tile[cl][rl] = (r < R && c < K) ? X[r * sr + c * sc] : 0.f;
  }
  __syncthreads();
  unsigned char* unit = out + static_cast<size_t>(u) * NS * tc::kUnitBytes;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int rl = (tid >> 3) + 32 * h, ch = tid & 7;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = tile[ch * 8 + e][rl];
    const uint32_t off = tc::unit_elem_offset(rl, ch * 8);
#pragma unroll
    for (int p = 0; p < NS; ++p) {
      __nv_bfloat16 hv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        hv[e] = __float2bfloat16_rn(v[e]);
        v[e] -= __bfloat162float(hv[e]);
      }
      *reinterpret_cast<uint4*>(unit + static_cast<size_t>(p) * tc::kUnitBytes + off) = *reinterpret_cast<const uint4*>(hv);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GEMM kernel: one 128 x BN output tile per CTA, K streamed in 64-wide atoms through a STAGES-deep ring.
// warp 0: bulk-copy producer | warp 1: TMEM owner + MMA issuer | warps 2-5: epilogue (one TMEM lane quarter each)
// ---------------------------------------------------------------------------------------------
// Epilogue variants.  EPI_STORE writes D.  EPI_COLSTATS never writes D: per row-tile it emits the column sums and
// sums of squares (for batch-statistics BN) and, per segment of `seg_len` consecutive rows (= one point-cloud block),
// the column extremum that the monotone BN+LeakyReLU maps to the maximum (max of y for gamma >= 0, min otherwise) with
// the 32-row group where it is attained (64-bit atomicMax); the exact row is resolved by the finalize kernel.
// With them  max_n LReLU(BN(y[n,c]))  is known without ever materialising y (conv6 + bn6 + global max, model.py:172-175).
enum { EPI_STORE = 0, EPI_COLSTATS = 1 };
struct GemmEpi {
  int seg_len;
  float* stat_partials;            // [row tiles][2][N]
  unsigned long long* gext;        // [segments][N]  (ordered(sign*y) << 32 | 32-row group), zero-initialised by the caller
  const float* gamma;              // [N]: sign(gamma) = sign of the BN scale decides whether the max or the min is needed
};

__device__ __forceinline__ void warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}
__device__ __forceinline__ void warp_transpose_max(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, s));
    }
  }
}
template <int NS, int BN, int STAGES>
struct GemmSmem {
  static constexpr uint32_t a_plane = 128 * 128;          // 128 rows x 128 B
  static constexpr uint32_t b_plane = BN * 128;
  static constexpr uint32_t stage = NS * (a_plane + b_plane);
  static constexpr uint32_t stat_bytes = 4 * 2 * BN * 4;  // EPI_COLSTATS: [4 quarters][2][BN] floats
  static constexpr uint32_t bytes = STAGES * stage + stat_bytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

// Persistent: CTA b works on the 128 x BN output tiles b, b + gridDim.x, ... (column tile fastest, so that concurrently
// running CTAs share their A rows in L2) with TWO accumulators in TMEM — the epilogue of a tile (tcgen05.ld, global stores or
// column statistics) overlaps the loads and MMAs of the next one.  The stage ring runs on across tiles.
template <int NS, int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(192, 1)
tc_gemm_kernel(const unsigned char* __restrict__ Ap, const unsigned char* __restrict__ Bp, float* __restrict__ D,
               int ldd, int M, int N, int KA, int ka_per_split, size_t split_stride, GemmEpi epi, int ns_a, int ns_b, int NT, int MT,
               int Z) {
  using SM = GemmSmem<NS, BN, STAGES>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* statS = reinterpret_cast<float*>(tiles + STAGES * SM::stage);
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + STAGES * SM::stage + SM::stat_bytes);
  uint64_t* empty = full + STAGES;
  uint64_t* accfull = empty + STAGES;     // [2] MMA commit -> epilogue
  uint64_t* accfree = accfull + 2;        // [2] epilogue (4 warps) -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfree + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = NT * MT * Z;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(full + s, 1);
      tc::mbar_init(empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(accfull + s, 1);
      tc::mbar_init(accfree + s, 4);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 2 * BN);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t taddr = *tmem_slot;

  if (warp == 0) {
    if (tc::elect_one()) {             // one lane, chosen with elect.sync: the issue code stays on the uniform datapath
      int it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int nt = t % NT, mt = (t / NT) % MT, z = t / (NT * MT);
        // split-K: this tile contracts k-atoms [ka_lo, ka_hi) and goes to D + z * split_stride
        const int ka_lo = z * ka_per_split, ka_hi = min(KA, ka_lo + ka_per_split);
        for (int ka = ka_lo; ka < ka_hi; ++ka, ++it) {
          const int s = it % STAGES;
          tc::mbar_wait(empty + s, ((it / STAGES) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(full + s, SM::stage);
          unsigned char* a_s = tiles + s * SM::stage;
          unsigned char* b_s = a_s + NS * SM::a_plane;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const unsigned char* src = Ap + (static_cast<size_t>(2 * mt + h) * KA + ka) * ns_a * tc::kUnitBytes;
#pragma unroll
            for (int p = 0; p < NS; ++p)
              tc::bulk_g2s(a_s + p * SM::a_plane + h * tc::kUnitBytes, src + p * tc::kUnitBytes, tc::kUnitBytes, full + s);
          }
#pragma unroll
          for (int h = 0; h < BN / 64; ++h) {
            const unsigned char* src = Bp + (static_cast<size_t>(nt * (BN / 64) + h) * KA + ka) * ns_b * tc::kUnitBytes;
#pragma unroll
            for (int p = 0; p < NS; ++p)
              tc::bulk_g2s(b_s + p * SM::b_plane + h * tc::kUnitBytes, src + p * tc::kUnitBytes, tc::kUnitBytes, full + s);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::idesc_bf16(128, BN);
      int it = 0, tl = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++tl) {
        const int z = t / (NT * MT);
        const int ka_lo = z * ka_per_split, ka_hi = min(KA, ka_lo + ka_per_split);
        const int ab = tl & 1;
        tc::mbar_wait(accfree + ab, ((tl >> 1) & 1) ^ 1);     // the epilogue has drained this accumulator (two tiles ago)
        tc::fence_after_sync();
        for (int ka = ka_lo; ka < ka_hi; ++ka, ++it) {
          const int s = it % STAGES;
          tc::mbar_wait(full + s, (it / STAGES) & 1);
          tc::fence_after_sync();
          const uint32_t a_s = tc::smem_u32(tiles + s * SM::stage);
          tc::mma_split_katom<NS>(taddr + ab * BN, a_s, SM::a_plane, a_s + NS * SM::a_plane, SM::b_plane, idesc, ka == ka_lo);
          tc::mma_commit(empty + s);       // frees the stage once these MMAs have read it
        }
        tc::mma_commit(accfull + ab);
      }
    }
  } else {
    const int quarter = warp & 3;        // TMEM lanes this warp may read: 32*quarter .. +31
    int tl = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++tl) {
    const int nt = t % NT, mt = (t / NT) % MT, z = t / (NT * MT);
    const int ab = tl & 1;
    const uint32_t tacc = taddr + ab * BN;
    float* Dz = D + static_cast<size_t>(z) * split_stride;
    tc::mbar_wait(accfull + ab, (tl >> 1) & 1);
    tc::fence_after_sync();
    const int row = mt * 128 + quarter * 32 + lane;
    if (EPI == EPI_STORE) {
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tc::tmem_ld32(tacc + (static_cast<uint32_t>(quarter * 32) << 16) + c0, v);
        const int n0 = nt * BN + c0;
        if (row < M) {
          float* dst = Dz + static_cast<size_t>(row) * ldd + n0;
          if (n0 + 32 <= N && (ldd & 7) == 0 && (reinterpret_cast<uintptr_t>(Dz) & 31) == 0) {
            // thread = output row: 32-byte stores write whole sectors (16-byte stores would touch every sector twice)
#pragma unroll
            for (int i = 0; i < 32; i += 8)
              asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + i), "f"(v[i]), "f"(v[i + 1]),
                           "f"(v[i + 2]), "f"(v[i + 3]), "f"(v[i + 4]), "f"(v[i + 5]), "f"(v[i + 6]), "f"(v[i + 7])
                           : "memory");
          } else if (n0 + 32 <= N && (ldd & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (n0 + i < N) dst[i] = v[i];
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(accfree + ab);         // accumulator drained (the stores are in flight on their own)
    } else {
      const bool rvalid = row < M;
      const int seg = rvalid ? row / epi.seg_len : -1;
      const unsigned vmask = __ballot_sync(0xffffffffu, rvalid);
      const int seg_lo = __shfl_sync(0xffffffffu, seg, 0);
      const int seg_hi = vmask ? __shfl_sync(0xffffffffu, seg, 31 - __clz(vmask)) : -1;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tc::tmem_ld32(tacc + (static_cast<uint32_t>(quarter * 32) << 16) + c0, v);
        const int col = nt * BN + c0 + lane;
        {
          float a[32], q[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            a[i] = rvalid ? v[i] : 0.f;
            q[i] = a[i] * a[i];
          }
          warp_transpose_sum(a, lane);
          warp_transpose_sum(q, lane);
          statS[(quarter * 2 + 0) * BN + c0 + lane] = a[0];
          statS[(quarter * 2 + 1) * BN + c0 + lane] = q[0];
        }
        if (vmask) {
          const float sgn = (col < N && epi.gamma[col] < 0.f) ? -1.f : 1.f;     // column of lane after the transpose
          for (int pass = 0; pass < 2; ++pass) {
            const int sg = pass == 0 ? seg_lo : seg_hi;
            if (pass == 1 && seg_hi == seg_lo) break;
            const bool in_seg = rvalid && seg == sg;
            float t[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float si = __shfl_sync(0xffffffffu, sgn, i);                  // sign of column c0 + i
              t[i] = in_seg ? si * v[i] : -INFINITY;
            }
            warp_transpose_max(t, lane);
            if (col < N && t[0] > -INFINITY)
              atomicMax(epi.gext + static_cast<size_t>(sg) * N + col,
                        (static_cast<unsigned long long>(ordered_bits(t[0])) << 32) | static_cast<uint32_t>(mt * 4 + quarter));
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(accfree + ab);
      // combine the four lane quarters in a fixed order: one partial per (row tile, column)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int te = threadIdx.x - 64;                     // 0..127 over the four epilogue warps
      if (te < BN && nt * BN + te < N) {
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const float sum = ((statS[(0 * 2 + w) * BN + te] + statS[(1 * 2 + w) * BN + te]) + statS[(2 * 2 + w) * BN + te]) +
                            statS[(3 * 2 + w) * BN + te];
          epi.stat_partials[(static_cast<size_t>(mt) * 2 + w) * N + nt * BN + te] = sum;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");       // statS is rewritten by the next tile
    }
    }  // tiles
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(taddr, 2 * BN);
}


// ---------------------------------------------------------------------------------------------
// Row-contraction GEMM ("TN", weight gradients):  D[N1, N2] = sum_r A[r, n1] * B[r, n2]  from the SAME split planes that
// the K-major GEMMs consume (units of 64 rows x 64 columns, 128 B per row).  The contraction index is the unit ROW, i.e.
// both operands are read MN-major: the 64 columns of a unit are the contiguous M (resp. N) extent, 16 rows are one
// k16 step (2048 B apart), wider M / N extents are further units `lbo` bytes apart (descriptor fields pinned on the GPU by
// tools/labs/mn_lab.py and used by ec_tc_bwd_b_kernel).  A weight gradient dW = dY^T X therefore needs NO transposing
// re-split of dY and X: it reuses the planes made for the data-gradient / forward GEMMs.
// One 128 x BN output tile per CTA and split (blockIdx.z = contiguous range of 64-row tiles); partial tiles are reduced in
// a fixed order by splitk_reduce_kernel.  Units beyond an operand's width are left zero in shared memory.
// ---------------------------------------------------------------------------------------------
template <int NS, int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
tc_gemm_tn_kernel(const unsigned char* __restrict__ Ap, int KA_a, int ns_a, const unsigned char* __restrict__ Bp, int KA_b, int ns_b,
                  float* __restrict__ D, int ldd, int N1, int N2, int RT, int rt_per_split, size_t split_stride, int accumulate) {
  using SM = GemmSmem<NS, BN, STAGES>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + STAGES * SM::stage);
  uint64_t* empty = full + STAGES;
  uint64_t* accfull = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = blockIdx.x, mt = blockIdx.y;
  const int rt_lo = blockIdx.z * rt_per_split, rt_hi = min(RT, rt_lo + rt_per_split);
  D += static_cast<size_t>(blockIdx.z) * split_stride;
  // operand units of this tile that exist (the rest of the stage stays zero)
  const int a_units = max(0, min(2, KA_a - 2 * mt)), b_units = max(0, min(BN / 64, KA_b - nt * (BN / 64)));

  // skipped (out-of-range) units must read as zeros: zero the stages once, only when this tile has such units
  if (a_units < 2 || b_units < BN / 64)
    for (uint32_t i = threadIdx.x; i < STAGES * SM::stage / 16; i += blockDim.x) reinterpret_cast<uint4*>(tiles)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(full + s, 1);
      tc::mbar_init(empty + s, 1);
    }
    tc::mbar_init(accfull, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, BN);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t taddr = *tmem_slot;

  if (warp == 0) {
    if (tc::elect_one()) {
      const uint32_t bytes = static_cast<uint32_t>(NS * (a_units + b_units)) * tc::kUnitBytes;
      for (int rt = rt_lo; rt < rt_hi; ++rt) {
        const int it = rt - rt_lo, s = it % STAGES;
        tc::mbar_wait(empty + s, ((it / STAGES) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(full + s, bytes);
        unsigned char* a_s = tiles + s * SM::stage;
        unsigned char* b_s = a_s + NS * SM::a_plane;
        for (int h = 0; h < a_units; ++h) {
          const unsigned char* src = Ap + (static_cast<size_t>(rt) * KA_a + 2 * mt + h) * ns_a * tc::kUnitBytes;
#pragma unroll
          for (int p = 0; p < NS; ++p)
            tc::bulk_g2s(a_s + p * SM::a_plane + h * tc::kUnitBytes, src + p * tc::kUnitBytes, tc::kUnitBytes, full + s);
        }
        for (int h = 0; h < b_units; ++h) {
          const unsigned char* src = Bp + (static_cast<size_t>(rt) * KA_b + nt * (BN / 64) + h) * ns_b * tc::kUnitBytes;
#pragma unroll
          for (int p = 0; p < NS; ++p)
            tc::bulk_g2s(b_s + p * SM::b_plane + h * tc::kUnitBytes, src + p * tc::kUnitBytes, tc::kUnitBytes, full + s);
        }
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::idesc_bf16(128, BN) | (1u << 15) | (1u << 16);       // A and B MN-major
      auto desc_mn = [](uint32_t saddr) {      // SWIZZLE_128B, 64-wide MN atoms 8192 B apart, 8-row K groups 1024 B apart
        const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((tc::kUnitBytes >> 4) & 0x3FFFu) << 16);
        const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        return (static_cast<uint64_t>(hi) << 32) | lo;
      };
      uint32_t acc = 0u;
      for (int rt = rt_lo; rt < rt_hi; ++rt) {
        const int it = rt - rt_lo, s = it % STAGES;
        tc::mbar_wait(full + s, (it / STAGES) & 1);
        tc::fence_after_sync();
        const uint32_t a_s = tc::smem_u32(tiles + s * SM::stage), b_s = a_s + NS * SM::a_plane;
        // plane pairs (i, j), i + j <= NS + 1, smallest magnitude first; 4 k16 steps of 16 rows each
#pragma unroll
        for (int sum = NS + 1; sum >= 2; --sum) {
#pragma unroll
          for (int i = 1; i <= NS; ++i) {
            const int j = sum - i;
            if (j < 1 || j > NS) continue;
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16) {
              tc::mma_bf16(taddr, desc_mn(a_s + (i - 1) * SM::a_plane + k16 * 2048), desc_mn(b_s + (j - 1) * SM::b_plane + k16 * 2048),
                           idesc, acc);
              acc = 1u;
            }
          }
        }
        tc::mma_commit(empty + s);
      }
      tc::mma_commit(accfull);
    }
  } else {
    const int quarter = warp & 3;
    tc::mbar_wait(accfull, 0);
    tc::fence_after_sync();
    const int row = mt * 128 + quarter * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tc::tmem_ld32(taddr + (static_cast<uint32_t>(quarter * 32) << 16) + c0, v);
      const int n0 = nt * BN + c0;
      if (row < N1 && n0 < N2) {
        float* dst = D + static_cast<size_t>(row) * ldd + n0;
        const bool vec = n0 + 32 <= N2 && (ldd & 3) == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0;
        if (accumulate) {
          // every split adds its partial tile into the (zeroed) output: no partial tiles in HBM, no reduce pass; the
          // summation order over the splits is not fixed (like the scatter reductions of the EdgeConv backward)
          if (vec) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(v[i]), "f"(v[i + 1]), "f"(v[i + 2]), "f"(v[i + 3])
                           : "memory");
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (n0 + i < N2) atomicAdd(dst + i, v[i]);
          }
        } else if (vec) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (n0 + i < N2) dst[i] = v[i];
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(taddr, BN);
}

// D[i] = sum_z P[z][i] in a fixed order (deterministic split-K reduction)
__global__ void splitk_reduce_kernel(const float* __restrict__ P, size_t split_stride, int splits, float* __restrict__ D,
                                     int ldd, int M, int N) {
  const size_t total = static_cast<size_t>(M) * N;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t r = i / N, c = i % N;
    // fixed summation order (deterministic); four independent chains hide the load latency
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int z = 0;
    for (; z + 4 <= splits; z += 4) {
      a0 += P[(z + 0) * split_stride + i];
      a1 += P[(z + 1) * split_stride + i];
      a2 += P[(z + 2) * split_stride + i];
      a3 += P[(z + 3) * split_stride + i];
    }
    for (; z < splits; ++z) a0 += P[z * split_stride + i];
    D[r * ldd + c] = (a0 + a1) + (a2 + a3);
  }
}

static size_t planes_bytes(int rows_pad, int KA, int NS) {
  return static_cast<size_t>(rows_pad / 64) * KA * NS * tc::kUnitBytes;
}

template <int NS>
static int launch_split(const float* X, long long sr, long long sc, int R, int K, int R_pad, int KA, unsigned char* out,
                        cudaStream_t st) {
  if (sr == 1 && sc != 1 && R_pad / 64 * KA < (1 << 30)) {       // contraction index is the slow dimension: transpose via smem
    split_planes_t_kernel<NS><<<R_pad / 64 * KA, 256, 0, st>>>(X, sr, sc, R, K, KA, out);
    DG_CUDA(cudaGetLastError());
    return DGCNN_OK;
  }
  const long long total = static_cast<long long>(R_pad) * KA * 8;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 32));
  split_planes_kernel<NS><<<blocks, 256, 0, st>>>(X, sr, sc, R, K, R_pad, KA, out);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

static int splitk_factor(int M, int N, int KA) {
  const int tiles = ((N + 127) / 128) * ((M + 127) / 128);
  if (tiles >= 148 || KA < 16) return 1;
  return std::max(1, std::min(KA / 8, (2 * 148 + tiles - 1) / tiles));
}

template <int NS, int BN, int STAGES>
static int launch_gemm(const unsigned char* Ap, const unsigned char* Bp, float* D, int ldd, int M, int N, int KA,
                       float* partials, cudaStream_t st, int ns_a = NS, int ns_b = NS) {
  using SM = GemmSmem<NS, BN, STAGES>;
  auto kern = tc_gemm_kernel<NS, BN, STAGES, EPI_STORE>;
  DG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(SM::bytes)));
  const int splits = splitk_factor(M, N, KA);
  const int NT = (N + BN - 1) / BN, MT = (M + 127) / 128, sms = std::max(1, num_sms());
  if (splits == 1) {
    kern<<<std::min(NT * MT, sms), 192, SM::bytes, st>>>(Ap, Bp, D, ldd, M, N, KA, KA, 0, GemmEpi{}, ns_a, ns_b, NT, MT, 1);
    DG_CUDA(cudaGetLastError());
    return DGCNN_OK;
  }
  const int per = (KA + splits - 1) / splits;
  const size_t stride = static_cast<size_t>(M) * N;
  const int Z = (KA + per - 1) / per;
  kern<<<std::min(NT * MT * Z, sms), 192, SM::bytes, st>>>(Ap, Bp, partials, N, M, N, KA, per, stride, GemmEpi{}, ns_a, ns_b, NT, MT, Z);
  DG_CUDA(cudaGetLastError());
  const size_t total = stride;
  splitk_reduce_kernel<<<static_cast<int>(std::min<size_t>((total + 255) / 256, 148 * 8)), 256, 0, st>>>(partials, stride, Z, D,
                                                                                                     ldd, M, N);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

}  // namespace dg

using namespace dg;

extern "C" size_t dgcnn_tc_gemm_workspace_bytes(int M, int N, int K, int nsplit) {
  if (M <= 0 || N <= 0 || K <= 0 || nsplit < 2 || nsplit > 3) return 0;
  const int KA = (K + 63) / 64;
  const size_t part = static_cast<size_t>(splitk_factor(M, N, KA)) * M * N * sizeof(float);
  return planes_bytes((M + 127) / 128 * 128, KA, nsplit) + planes_bytes((N + 127) / 128 * 128, KA, nsplit) +
         (splitk_factor(M, N, KA) > 1 ? part : 0) + 4096;
}

extern "C" int dgcnn_tc_gemm_f32(const float* A, int64_t a_sr, int64_t a_sc, const float* B, int64_t b_sr, int64_t b_sc,
                                 float* D, int ldd, int M, int N, int K, int nsplit, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  DG_REQUIRE(A && B && D && workspace, DGCNN_E_INVALID, "tc_gemm: null pointer");
  DG_REQUIRE(M > 0 && N > 0 && K > 0 && ldd >= N, DGCNN_E_INVALID, "tc_gemm: bad shape M=%d N=%d K=%d ldd=%d", M, N, K, ldd);
  DG_REQUIRE(nsplit == 2 || nsplit == 3, DGCNN_E_UNSUPPORTED, "tc_gemm: nsplit must be 2 or 3");
  DG_REQUIRE(workspace_bytes >= dgcnn_tc_gemm_workspace_bytes(M, N, K, nsplit), DGCNN_E_WORKSPACE, "tc_gemm: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int KA = (K + 63) / 64, Mp = (M + 127) / 128 * 128, Np = (N + 127) / 128 * 128;
  unsigned char* Ap = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  unsigned char* Bp = Ap + planes_bytes(Mp, KA, nsplit);
  float* partials = reinterpret_cast<float*>(Bp + planes_bytes(Np, KA, nsplit));
  int rc;
  if (nsplit == 3) {
    if ((rc = launch_split<3>(A, a_sr, a_sc, M, K, Mp, KA, Ap, st))) return rc;
    if ((rc = launch_split<3>(B, b_sr, b_sc, N, K, Np, KA, Bp, st))) return rc;
    return launch_gemm<3, 128, 2>(Ap, Bp, D, ldd, M, N, KA, partials, st);
  }
  if ((rc = launch_split<2>(A, a_sr, a_sc, M, K, Mp, KA, Ap, st))) return rc;
  if ((rc = launch_split<2>(B, b_sr, b_sc, N, K, Np, KA, Bp, st))) return rc;
  // wide tiles where the grid stays full: a 128x256 MMA reads 12 KB of operands for 128 cycles of math (96 B/clk),
  // a 128x128 one 8 KB for 64 (125 B/clk, at the shared-memory limit)
  if (Np % 256 == 0 && splitk_factor(M, N, KA) == 1 && static_cast<long long>((M + 127) / 128) * (Np / 256) >= 2 * 148)
    return launch_gemm<2, 256, 2>(Ap, Bp, D, ldd, M, N, KA, partials, st);
  return launch_gemm<2, 128, 3>(Ap, Bp, D, ldd, M, N, KA, partials, st);
}


namespace dg {
static int tn_splits(int N1, int N2, int BN, int RT) {
  // one wave of CTAs, at least 4 row tiles each: every split costs a partial tile in HBM and a pass of the reduce kernel
  const int tiles = ((N1 + 127) / 128) * ((N2 + BN - 1) / BN);
  return std::max(1, std::min((RT + 3) / 4, (148 + tiles - 1) / tiles));
}
template <int NS, int BN, int STAGES>
static int launch_gemm_tn(const unsigned char* Ap, int KA_a, int ns_a, const unsigned char* Bp, int KA_b, int ns_b, float* D, int ldd,
                          int N1, int N2, int RT, float* partials, bool atomic, cudaStream_t st) {
  using SM = GemmSmem<NS, BN, STAGES>;
  auto kern = tc_gemm_tn_kernel<NS, BN, STAGES>;
  DG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(SM::bytes)));
  const int splits = tn_splits(N1, N2, BN, RT);
  const int per = (RT + splits - 1) / splits;
  dim3 grid((N2 + BN - 1) / BN, (N1 + 127) / 128, (RT + per - 1) / per);
  if (grid.z == 1) {
    kern<<<grid, 192, SM::bytes, st>>>(Ap, KA_a, ns_a, Bp, KA_b, ns_b, D, ldd, N1, N2, RT, per, 0, 0);
    DG_CUDA(cudaGetLastError());
    return DGCNN_OK;
  }
  if (atomic) {
    DG_CUDA(cudaMemset2DAsync(D, static_cast<size_t>(ldd) * 4, 0, static_cast<size_t>(N2) * 4, N1, st));
    kern<<<grid, 192, SM::bytes, st>>>(Ap, KA_a, ns_a, Bp, KA_b, ns_b, D, ldd, N1, N2, RT, per, 0, 1);
    DG_CUDA(cudaGetLastError());
    return DGCNN_OK;
  }
  const size_t stride = static_cast<size_t>(N1) * N2;
  kern<<<grid, 192, SM::bytes, st>>>(Ap, KA_a, ns_a, Bp, KA_b, ns_b, partials, N2, N1, N2, RT, per, stride, 0);
  DG_CUDA(cudaGetLastError());
  splitk_reduce_kernel<<<static_cast<int>(std::min<size_t>((stride + 255) / 256, 148 * 8)), 256, 0, st>>>(partials, stride, grid.z, D, ldd,
                                                                                                      N1, N2);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}
}  // namespace dg

// ---- plane-level interface: split once, multiply many times ------------------------------------------------------
extern "C" size_t dgcnn_tc_planes_bytes(int R, int K, int nsplit) {
  if (R <= 0 || K <= 0 || nsplit < 2 || nsplit > 3) return 0;
  return planes_bytes((R + 127) / 128 * 128, (K + 63) / 64, nsplit) + 1024;
}

extern "C" int dgcnn_tc_split_planes_f32(const float* X, int64_t row_stride, int64_t col_stride, int R, int K, int nsplit,
                                         void* planes, size_t planes_bytes_avail, void* stream) {
  DG_REQUIRE(X && planes, DGCNN_E_INVALID, "tc_split_planes: null pointer");
  DG_REQUIRE(R > 0 && K > 0 && (nsplit == 2 || nsplit == 3), DGCNN_E_INVALID, "tc_split_planes: bad arguments");
  DG_REQUIRE((reinterpret_cast<uintptr_t>(planes) & 127) == 0, DGCNN_E_INVALID, "tc_split_planes: planes must be 128-byte aligned");
  DG_REQUIRE(planes_bytes_avail >= dgcnn_tc_planes_bytes(R, K, nsplit) - 1024, DGCNN_E_WORKSPACE, "tc_split_planes: planes buffer too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int KA = (K + 63) / 64, Rp = (R + 127) / 128 * 128;
  unsigned char* out = static_cast<unsigned char*>(planes);
  return nsplit == 3 ? launch_split<3>(X, row_stride, col_stride, R, K, Rp, KA, out, st)
                     : launch_split<2>(X, row_stride, col_stride, R, K, Rp, KA, out, st);
}

extern "C" size_t dgcnn_tc_gemm_planes_workspace_bytes(int mode, int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (mode == 0) return static_cast<size_t>(splitk_factor(M, N, (K + 63) / 64)) * M * N * sizeof(float) + 1024;
  if (mode == 2) return 1024;      // the splits are added into D
  const int RT = (K + 63) / 64;
  return static_cast<size_t>(tn_splits(M, N, 128, RT)) * M * N * sizeof(float) + 1024;
}

extern "C" int dgcnn_tc_gemm_planes_f32(int mode, const void* a_planes, int a_cols, int a_nsplit, const void* b_planes, int b_cols,
                                        int b_nsplit, int nsplit, float* D, int ldd, int M, int N, int K, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  DG_REQUIRE(a_planes && b_planes && D && workspace, DGCNN_E_INVALID, "tc_gemm_planes: null pointer");
  DG_REQUIRE(M > 0 && N > 0 && K > 0 && ldd >= N, DGCNN_E_INVALID, "tc_gemm_planes: bad shape M=%d N=%d K=%d ldd=%d", M, N, K, ldd);
  DG_REQUIRE((nsplit == 2 || nsplit == 3) && a_nsplit >= nsplit && b_nsplit >= nsplit && a_nsplit <= 3 && b_nsplit <= 3, DGCNN_E_INVALID,
             "tc_gemm_planes: nsplit=%d must not exceed the stored planes (%d, %d)", nsplit, a_nsplit, b_nsplit);
  DG_REQUIRE(((reinterpret_cast<uintptr_t>(a_planes) | reinterpret_cast<uintptr_t>(b_planes)) & 127) == 0, DGCNN_E_INVALID,
             "tc_gemm_planes: planes must be 128-byte aligned");
  DG_REQUIRE(workspace_bytes >= dgcnn_tc_gemm_planes_workspace_bytes(mode, M, N, K), DGCNN_E_WORKSPACE, "tc_gemm_planes: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned char* Ap = static_cast<const unsigned char*>(a_planes);
  const unsigned char* Bp = static_cast<const unsigned char*>(b_planes);
  float* partials = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  if (mode == 0) {
    // D[M,N] = A[M,K] B[N,K]^T: planes of A over [M, K] (a_cols == K), of B over [N, K] (b_cols == K)
    DG_REQUIRE(a_cols == K && b_cols == K, DGCNN_E_INVALID, "tc_gemm_planes: mode 0 needs a_cols == b_cols == K");
    const int KA = (K + 63) / 64, Np = (N + 127) / 128 * 128;
    if (nsplit == 3) return launch_gemm<3, 128, 2>(Ap, Bp, D, ldd, M, N, KA, partials, st, a_nsplit, b_nsplit);
    if (Np % 256 == 0 && splitk_factor(M, N, KA) == 1 && static_cast<long long>((M + 127) / 128) * (Np / 256) >= 2 * 148)
      return launch_gemm<2, 256, 2>(Ap, Bp, D, ldd, M, N, KA, partials, st, a_nsplit, b_nsplit);
    return launch_gemm<2, 128, 3>(Ap, Bp, D, ldd, M, N, KA, partials, st, a_nsplit, b_nsplit);
  }
  DG_REQUIRE(mode == 1 || mode == 2, DGCNN_E_INVALID, "tc_gemm_planes: mode must be 0 (A B^T), 1 (A^T B) or 2 (A^T B, splits added atomically)");
  // D[M,N] = sum_r A[r, m] B[r, n]: planes of A over [K rows, a_cols = M], of B over [K rows, b_cols = N]
  DG_REQUIRE(a_cols == M && b_cols == N, DGCNN_E_INVALID, "tc_gemm_planes: mode 1 needs a_cols == M and b_cols == N");
  const int RT = (K + 63) / 64, KA_a = (M + 63) / 64, KA_b = (N + 63) / 64;
  if (nsplit == 3) return launch_gemm_tn<3, 128, 2>(Ap, KA_a, a_nsplit, Bp, KA_b, b_nsplit, D, ldd, M, N, RT, partials, mode == 2, st);
  return launch_gemm_tn<2, 128, 3>(Ap, KA_a, a_nsplit, Bp, KA_b, b_nsplit, D, ldd, M, N, RT, partials, mode == 2, st);
}

// ---------------------------------------------------------------------------------------------
// GEMM + batch statistics + per-block column extrema (conv6 + bn6 + LeakyReLU + max over the points), no D store
// ---------------------------------------------------------------------------------------------
namespace dg {
// 32 columns x 8 row-tile groups per CTA: group g adds the row-tile partials g, g+8, ... in FP64 (coalesced 128-byte
// reads), the 8 group sums are added in order (deterministic); then the BN scalars.  (One thread per column looping over
// all 438 row tiles was a 70 us latency chain.)
__global__ void __launch_bounds__(256) colstats_bn_kernel(const float* __restrict__ stat_partials, int num_mt, int N, double count,
                                   const float* __restrict__ gamma, float eps, float momentum, int use_batch_stats,
                                   float* running_mean, float* running_var, float* mean_out, float* rstd_out) {
  __shared__ double red[2][8][32];
  const int cl = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double s = 0.0, q = 0.0;
  if (use_batch_stats && c < N) {
#pragma unroll 4
    for (int i = g; i < num_mt; i += 8) {
      s += stat_partials[(static_cast<size_t>(i) * 2 + 0) * N + c];
      q += stat_partials[(static_cast<size_t>(i) * 2 + 1) * N + c];
    }
  }
  red[0][g][cl] = s;
  red[1][g][cl] = q;
  __syncthreads();
  if (g != 0 || c >= N) return;
  double m, rs;
  if (use_batch_stats) {
    s = 0.0; q = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += red[0][j][cl]; q += red[1][j][cl]; }
    m = s / count;
    double var = q / count - m * m;
    if (var < 0.0) var = 0.0;
    rs = 1.0 / sqrt(var + static_cast<double>(eps));
    if (running_mean) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = static_cast<float>((1.0 - momentum) * running_mean[c] + momentum * m);
      running_var[c] = static_cast<float>((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
  } else {
    m = running_mean[c];
    rs = 1.0 / sqrt(static_cast<double>(running_var[c]) + static_cast<double>(eps));
  }
  mean_out[c] = static_cast<float>(m);
  rstd_out[c] = static_cast<float>(rs);
}

// one warp per (segment, column): decode the extremum, apply BN + LeakyReLU, and resolve the attaining row inside the
// recorded 32-row group by recomputing the 32 dot products in FP32 (lane = row)
__global__ void segmax_finalize_kernel(const float* __restrict__ A, long long a_sr, long long a_sc, const float* __restrict__ Bw,
                                       long long b_sr, long long b_sc, int M, int N, int K, int segments, int seg_len,
                                       const unsigned long long* __restrict__ gext, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, const float* __restrict__ mean,
                                       const float* __restrict__ rstd, float* __restrict__ gout, float* __restrict__ ysel,
                                       int* __restrict__ nstar) {
  const int lane = threadIdx.x & 31;
  const long long item = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (item >= static_cast<long long>(segments) * N) return;
  const int b = static_cast<int>(item / N), c = static_cast<int>(item % N);
  const unsigned long long key = gext[item];
  const uint32_t u = static_cast<uint32_t>(key >> 32);
  const float sgn = gamma[c] < 0.f ? -1.f : 1.f;
  const float y = sgn * __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
  const int row0 = static_cast<int>(key & 0xFFFFFFFFull) * 32;
  // dot products of the 32 candidate rows with weight row c: lane = row (its 4 K bytes are whole 128-byte lines, the weight
  // row is a broadcast), four independent chains.  (Lane = slice of k with a 32 x 32 register transpose needed 168
  // registers: 10 % occupancy on a latency-bound kernel, 82 us.)
  const int rr = min(row0 + lane, M - 1);
  const float* ar = A + static_cast<long long>(rr) * a_sr;
  const float* wr = Bw + static_cast<long long>(c) * b_sr;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (a_sc == 1 && b_sc == 1 && (K & 3) == 0 && ((reinterpret_cast<uintptr_t>(ar) | reinterpret_cast<uintptr_t>(wr)) & 15) == 0) {
#pragma unroll 4
    for (int kk = 0; kk < K; kk += 4) {
      const float4 a4 = *reinterpret_cast<const float4*>(ar + kk), w4 = *reinterpret_cast<const float4*>(wr + kk);
      acc[0] = __fmaf_rn(a4.x, w4.x, acc[0]);
      acc[1] = __fmaf_rn(a4.y, w4.y, acc[1]);
      acc[2] = __fmaf_rn(a4.z, w4.z, acc[2]);
      acc[3] = __fmaf_rn(a4.w, w4.w, acc[3]);
    }
  } else {
    for (int kk = 0; kk < K; ++kk) acc[kk & 3] = __fmaf_rn(ar[kk * a_sc], wr[kk * b_sc], acc[kk & 3]);
  }
  const float dot = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  const int row = row0 + lane;
  float d = -INFINITY;
  if (row < M && row / seg_len == b) d = -fabsf(dot - y);       // the row whose value reproduces the recorded extremum best
  float best = d;
  int bl = lane;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
    if (ob > best || (ob == best && ol < bl)) { best = ob; bl = ol; }
  }
  if (lane == 0) {
    const float z = __fmaf_rn(y - mean[c], gamma[c] * rstd[c], beta[c]);
    gout[item] = z > 0.f ? z : kSlope * z;
    ysel[item] = y;
    nstar[item] = static_cast<int>(key & 0xFFFFFFFFull) * 32 + bl;
  }
}
}  // namespace dg

extern "C" size_t dgcnn_tc_gemm_colstats_workspace_bytes(int M, int N, int K, int segments) {
  if (M <= 0 || N <= 0 || K <= 0 || segments <= 0) return 0;
  const int KA = (K + 63) / 64;
  const size_t planes = planes_bytes((M + 127) / 128 * 128, KA, 3) + planes_bytes((N + 127) / 128 * 128, KA, 3);
  const size_t stats = static_cast<size_t>((M + 127) / 128) * 2 * N * sizeof(float);
  const size_t ext = static_cast<size_t>(segments) * N * 8;
  return planes + stats + ext + 8192;
}

// g[segments,N] = max over the rows of each segment of LeakyReLU(BN(A B^T)); BN with batch statistics over all M rows
// (use_batch_stats = 1: training, running stats updated when given) or with the running statistics (0: inference).
// Also returns mean/rstd [N] and, per (segment, column), the selected pre-BN value `ysel` and its row `nstar`.
extern "C" int dgcnn_tc_gemm_bn_lrelu_segmax_f32(const float* A, int64_t a_sr, int64_t a_sc, const float* B, int64_t b_sr,
                                                 int64_t b_sc, int M, int N, int K, int seg_len, const float* gamma,
                                                 const float* beta, float eps, float momentum, int use_batch_stats,
                                                 float* running_mean, float* running_var, float* mean, float* rstd,
                                                 float* g, float* ysel, int32_t* nstar, void* workspace,
                                                 size_t workspace_bytes, void* stream) {
  DG_REQUIRE(A && B && gamma && beta && mean && rstd && g && ysel && nstar && workspace, DGCNN_E_INVALID, "gemm_segmax: null pointer");
  DG_REQUIRE(M > 0 && N > 0 && K > 0 && seg_len >= 32 && M % seg_len == 0, DGCNN_E_INVALID,
             "gemm_segmax: need seg_len >= 32 and M a multiple of seg_len (M=%d seg_len=%d)", M, seg_len);
  DG_REQUIRE(use_batch_stats || (running_mean && running_var), DGCNN_E_INVALID, "gemm_segmax: inference needs running statistics");
  DG_REQUIRE(!running_mean == !running_var, DGCNN_E_INVALID, "gemm_segmax: running_mean/var must come together");
  const int segments = M / seg_len;
  DG_REQUIRE(workspace_bytes >= dgcnn_tc_gemm_colstats_workspace_bytes(M, N, K, segments), DGCNN_E_WORKSPACE, "gemm_segmax: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int KA = (K + 63) / 64, Mp = (M + 127) / 128 * 128, Np = (N + 127) / 128 * 128, num_mt = Mp / 128;
  unsigned char* Ap = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  unsigned char* Bp = Ap + planes_bytes(Mp, KA, 3);
  float* stats = reinterpret_cast<float*>(Bp + planes_bytes(Np, KA, 3));
  unsigned long long* gext = reinterpret_cast<unsigned long long*>(stats + static_cast<size_t>(num_mt) * 2 * N);
  int rc;
  if ((rc = launch_split<3>(A, a_sr, a_sc, M, K, Mp, KA, Ap, st))) return rc;
  if ((rc = launch_split<3>(B, b_sr, b_sc, N, K, Np, KA, Bp, st))) return rc;
  DG_CUDA(cudaMemsetAsync(gext, 0, static_cast<size_t>(segments) * N * 8, st));
  using SM = GemmSmem<3, 128, 2>;
  auto kern = tc_gemm_kernel<3, 128, 2, EPI_COLSTATS>;
  DG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(SM::bytes)));
  GemmEpi epi{seg_len, stats, gext, gamma};
  kern<<<std::min((Np / 128) * num_mt, std::max(1, num_sms())), 192, SM::bytes, st>>>(Ap, Bp, nullptr, 0, M, N, KA, KA, 0, epi, 3, 3,
                                                                                   Np / 128, num_mt, 1);
  DG_CUDA(cudaGetLastError());
  colstats_bn_kernel<<<(N + 31) / 32, 256, 0, st>>>(stats, num_mt, N, static_cast<double>(M), gamma, eps, momentum, use_batch_stats,
                                                      running_mean, running_var, mean, rstd);
  DG_CUDA(cudaGetLastError());
  const long long items = static_cast<long long>(segments) * N;
  segmax_finalize_kernel<<<static_cast<int>((items * 32 + 255) / 256), 256, 0, st>>>(A, a_sr, a_sc, B, b_sr, b_sc, M, N, K, segments,
                                                                                  seg_len, gext, gamma, beta, mean, rstd, g, ysel,
                                                                                  nstar);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

// ---------------------------------------------------------------------------------------------
// Descriptor lab (debug): D[128 x 64] = [X | Y]^T-contraction over the 128 tile rows using MN-major operand
// descriptors on K-major-stored tiles: D[m][n] = sum_r S[r][m] * Y[r][n], S = X for m < 64, Y for m >= 64.
// The descriptor fields are runtime parameters so that one run can scan the candidate encodings.
// ---------------------------------------------------------------------------------------------
namespace dg {
__global__ void __launch_bounds__(128, 1)
tc_mn_lab_kernel(const float* __restrict__ X, const float* __restrict__ Y, float* __restrict__ D, uint32_t lbo_a,
                 uint32_t sbo_a, uint32_t lbo_b, uint32_t sbo_b, uint32_t kstep, int a_major, int b_major) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* Xs = sm;            // [128 rows][128 B] bf16, swizzled like every other tile
  unsigned char* Ys = sm + 16384;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 98304 - 64);   // operands may be addressed up to ~64 KB by bad encodings
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc(slot, 64);
  for (int i = tid; i < 128 * 8; i += 128) {
    const int r = i >> 3, ch = i & 7;
    __nv_bfloat16 hx[8], hy[8];
    for (int e = 0; e < 8; ++e) {
      hx[e] = __float2bfloat16_rn(X[r * 64 + ch * 8 + e]);
      hy[e] = __float2bfloat16_rn(Y[r * 64 + ch * 8 + e]);
    }
    const uint32_t off = r * 128 + (((ch ^ (r & 7)) & 7) << 4);
    *reinterpret_cast<uint4*>(Xs + off) = *reinterpret_cast<const uint4*>(hx);
    *reinterpret_cast<uint4*>(Ys + off) = *reinterpret_cast<const uint4*>(hy);
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t taddr = *slot;
  if (tid == 0) {
    const uint32_t idesc = tc::idesc_bf16(128, 64) | (static_cast<uint32_t>(a_major) << 15) | (static_cast<uint32_t>(b_major) << 16);
    auto desc = [](uint32_t saddr, uint32_t lbo, uint32_t sbo) {
      const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
      const uint32_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
      return (static_cast<uint64_t>(hi) << 32) | lo;
    };
    for (int k16 = 0; k16 < 8; ++k16)
      tc::mma_bf16(taddr, desc(tc::smem_u32(Xs) + k16 * kstep, lbo_a, sbo_a), desc(tc::smem_u32(Ys) + k16 * kstep, lbo_b, sbo_b),
                   idesc, k16 > 0);
    tc::mma_commit(bar);
  }
  __syncwarp();
  tc::mbar_wait(bar, 0);
  tc::fence_after_sync();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    float v[32];
    tc::tmem_ld32(taddr + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * 64 + c0 + i] = v[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(taddr, 64);
}
}  // namespace dg

// debug: A operand in tensor memory.  D[128x64] = X[128x64] Y[64x64]^T with bf16-rounded inputs; X goes registers ->
// TMEM (tcgen05.st, thread = row, 32-bit column j = K elements 2j (low half unless `swap`) and 2j+1), Y -> smem planes.
namespace dg {
__global__ void __launch_bounds__(128, 1)
tc_tmem_a_lab_kernel(const float* __restrict__ X, const float* __restrict__ Y, float* __restrict__ D, int swap) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* Ys = sm;            // [64 rows][128 B] bf16, SWIZZLE_128B K-major
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc(slot, 128);          // columns 0-63: D, 64-95: A (32 words = 64 bf16 per row)
  for (int i = tid; i < 64 * 8; i += 128) {
    const int r = i >> 3, ch = i & 7;
    __nv_bfloat16 hy[8];
    for (int e = 0; e < 8; ++e) hy[e] = __float2bfloat16_rn(Y[r * 64 + ch * 8 + e]);
    *reinterpret_cast<uint4*>(Ys + r * 128 + (((ch ^ (r & 7)) & 7) << 4)) = *reinterpret_cast<const uint4*>(hy);
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t taddr = *slot;
  {
    uint32_t w[32];
    const float* xr = X + (warp * 32 + lane) * 64;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const __nv_bfloat162 h = swap ? __floats2bfloat162_rn(xr[2 * j + 1], xr[2 * j]) : __floats2bfloat162_rn(xr[2 * j], xr[2 * j + 1]);
      w[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    tc::tmem_st32(taddr + (static_cast<uint32_t>(warp * 32) << 16) + 64, w);
    tc::tmem_st_wait();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    tc::fence_after_sync();
    constexpr uint32_t idesc = tc::idesc_bf16(128, 64);
    for (int k16 = 0; k16 < 4; ++k16)
      tc::mma_bf16_ts(taddr, taddr + 64 + k16 * 8, tc::smem_desc_sw128(tc::smem_u32(Ys) + k16 * 32), idesc, k16 > 0);
    tc::mma_commit(bar);
  }
  __syncwarp();
  tc::mbar_wait(bar, 0);
  tc::fence_after_sync();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    float v[32];
    tc::tmem_ld32(taddr + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * 64 + c0 + i] = v[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(taddr, 128);
}
}  // namespace dg

extern "C" int dgcnn_debug_tmem_a_lab(const float* X, const float* Y, float* D, int swap, void* stream) {
  DG_CUDA(cudaFuncSetAttribute(dg::tc_tmem_a_lab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  dg::tc_tmem_a_lab_kernel<<<1, 128, 16384, static_cast<cudaStream_t>(stream)>>>(X, Y, D, swap);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_debug_mn_lab(const float* X, const float* Y, float* D, int lbo_a, int sbo_a, int lbo_b, int sbo_b,
                                  int kstep, int a_major, int b_major, void* stream) {
  DG_CUDA(cudaFuncSetAttribute(dg::tc_mn_lab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
  dg::tc_mn_lab_kernel<<<1, 128, 98304, static_cast<cudaStream_t>(stream)>>>(X, Y, D, lbo_a, sbo_a, lbo_b, sbo_b, kstep, a_major,
                                                                               b_major);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}
