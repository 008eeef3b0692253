// graph_transpose.cu — the REPRODUCIBLE form of the two scatters of the backward pass (opt-in: ops.DETERMINISTIC).
//
// The default kernels add conv_a's per-edge gradient into its neighbour's row with float reductions (atomics): fast, but the
// summation order — and with it the last bits of every gradient — changes from run to run.  Here the kNN table is
// transposed once per graph (CSR over the in-edges of every point, each segment sorted by edge id) and r_j is GATHERED in
// that fixed order from a per-edge gradient tensor; the global-feature scatter is done by one owner warp per target row.
//   in_start[j] .. in_start[j+1]   the in-edges of point j (edge id e = i * k + slot, i the source point), ascending
// Costs a per-edge tensor in HBM again (430 MB per block at the training shape): the price of bit-reproducible gradients.
#include "common.cuh"

namespace dg {

__global__ void __launch_bounds__(256) gt_count_kernel(const int32_t* __restrict__ idx, long long E, int N, int k, int* __restrict__ cnt) {
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < E; e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = e / k;
    atomicAdd(cnt + (i / N) * N + idx[e], 1);              // integer counts: order-independent
  }
}

// exclusive scan of cnt[0..M) -> start[0..M]; cnt is reset to zero (fill cursors).  One CTA.
__global__ void __launch_bounds__(1024) gt_scan_kernel(int M, int* __restrict__ cnt, int* __restrict__ start) {
  __shared__ int part[1024];
  const int tid = threadIdx.x;
  const int per = (M + 1023) / 1024, lo = min(M, tid * per), hi = min(M, lo + per);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += cnt[i];
  part[tid] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = tid >= o ? part[tid - o] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  int run = part[tid] - s;
  for (int i = lo; i < hi; ++i) {
    const int c = cnt[i];
    start[i] = run;
    run += c;
    cnt[i] = 0;
  }
  if (tid == 1023) start[M] = part[1023];
}

__global__ void __launch_bounds__(256) gt_fill_kernel(const int32_t* __restrict__ idx, long long E, int N, int k, const int* __restrict__ start,
                                                      int* __restrict__ cursor, int* __restrict__ in_edge) {
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < E; e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = e / k;
    const long long j = (i / N) * N + idx[e];
    in_edge[start[j] + atomicAdd(cursor + j, 1)] = static_cast<int>(e);
  }
}

// the fill order inside a segment is arbitrary: sort every segment by edge id (thread per point, insertion sort: ~k entries)
__global__ void __launch_bounds__(256) gt_sort_kernel(int M, const int* __restrict__ start, int* __restrict__ in_edge) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int lo = start[j], hi = start[j + 1];
  for (int a = lo + 1; a < hi; ++a) {
    const int v = in_edge[a];
    int b = a - 1;
    while (b >= lo && in_edge[b] > v) {
      in_edge[b + 1] = in_edge[b];
      --b;
    }
    in_edge[b + 1] = v;
  }
}

// r[j, :] = sum over the in-edges of j, in ascending edge id, of dy[e, :]      (thread = (point, 4 channels))
__global__ void __launch_bounds__(256) gt_gather_kernel(const float* __restrict__ dy, const int* __restrict__ start,
                                                        const int* __restrict__ in_edge, int M, float* __restrict__ r, int ld_r) {
  const int c = 4 * (threadIdx.x & 15);
  for (int j = blockIdx.x * 16 + (threadIdx.x >> 4); j < M; j += gridDim.x * 16) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = start[j]; q < start[j + 1]; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(dy + static_cast<size_t>(in_edge[q]) * kCout + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(r + static_cast<size_t>(j) * ld_r + c) = acc;
  }
}

// global-feature scatter, one owner per target row: warp per (block b, channel e); the warp whose e is the LOWEST channel that
// selected row n*(b, e) adds the contributions of every channel that selected the same row, in ascending e.
__global__ void __launch_bounds__(256)
gf_scatter_det_kernel(const float* __restrict__ w, int ldw, const int* __restrict__ nstar, const float* __restrict__ gdz,
                      const float* __restrict__ a, int B, int E, int K, float* __restrict__ dfeat, int lddf) {
  const int lane = threadIdx.x & 31;
  const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (item >= B * E) return;
  const int b = item / E, e = item - b * E;
  const int row = nstar[b * E + e];
  // is there a lower channel with the same row?  (then that warp owns the row)
  bool owned = false;
  for (int e0 = lane; e0 < e && !owned; e0 += 32) owned = nstar[b * E + e0] == row;
  if (__any_sync(0xffffffffu, owned)) return;
  for (int k0 = 0; k0 < K; k0 += 128) {
    const int kk = k0 + 4 * lane;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e1 = e; e1 < E; ++e1) {                       // warp-uniform loop: every lane sees the same matches, ascending
      if (nstar[b * E + e1] != row) continue;
      if (kk < K) {
        const float s = a[e1] * gdz[b * E + e1];
        const float4 wv = *reinterpret_cast<const float4*>(w + static_cast<size_t>(e1) * ldw + kk);
        acc.x = __fmaf_rn(s, wv.x, acc.x); acc.y = __fmaf_rn(s, wv.y, acc.y);
        acc.z = __fmaf_rn(s, wv.z, acc.z); acc.w = __fmaf_rn(s, wv.w, acc.w);
      }
    }
    if (kk < K) {
      float4* p = reinterpret_cast<float4*>(dfeat + static_cast<size_t>(row) * lddf + kk);
      float4 t = *p;
      t.x += acc.x; t.y += acc.y; t.z += acc.z; t.w += acc.w;
      *p = t;                                              // the only writer of this row in this kernel
    }
  }
}

}  // namespace dg

using namespace dg;

extern "C" size_t dgcnn_graph_transpose_workspace_bytes(int B, int N) {
  return (B > 0 && N > 0) ? static_cast<size_t>(B) * N * 4 + 256 : 0;
}

extern "C" int dgcnn_graph_transpose(const int32_t* idx, int B, int N, int k, int32_t* in_start, int32_t* in_edge, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  DG_REQUIRE(idx && in_start && in_edge && workspace, DGCNN_E_INVALID, "graph_transpose: null pointer");
  DG_REQUIRE(B > 0 && N > 0 && k > 0 && static_cast<long long>(B) * N * k < (1LL << 31), DGCNN_E_INVALID, "graph_transpose: bad shape");
  DG_REQUIRE(workspace_bytes >= dgcnn_graph_transpose_workspace_bytes(B, N), DGCNN_E_WORKSPACE, "graph_transpose: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int M = B * N;
  const long long E = static_cast<long long>(M) * k;
  int* cnt = static_cast<int*>(workspace);
  DG_CUDA(cudaMemsetAsync(cnt, 0, static_cast<size_t>(M) * 4, st));
  const int blocks = static_cast<int>(std::min<long long>((E + 255) / 256, 148 * 16));
  gt_count_kernel<<<blocks, 256, 0, st>>>(idx, E, N, k, cnt);
  gt_scan_kernel<<<1, 1024, 0, st>>>(M, cnt, in_start);
  gt_fill_kernel<<<blocks, 256, 0, st>>>(idx, E, N, k, in_start, cnt, in_edge);
  gt_sort_kernel<<<(M + 255) / 256, 256, 0, st>>>(M, in_start, in_edge);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_gather_in_edges_f32(const float* dy_edges, const int32_t* in_start, const int32_t* in_edge, int M, float* r, int ld_r,
                                         void* stream) {
  DG_REQUIRE(dy_edges && in_start && in_edge && r && M > 0 && ld_r >= kCout && ld_r % 4 == 0 &&
                 ((reinterpret_cast<uintptr_t>(dy_edges) | reinterpret_cast<uintptr_t>(r)) & 15) == 0,
             DGCNN_E_INVALID, "gather_in_edges: null pointer / alignment (dy_edges [E,64], r rows of >= 64 floats, 16-byte aligned)");
  gt_gather_kernel<<<std::max(1, std::min((M + 15) / 16, 148 * 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy_edges, in_start, in_edge,
                                                                                                          M, r, ld_r);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_global_feature_scatter_det_f32(const float* w, int ld_w, const int32_t* nstar, const float* gdz, const float* a, int B,
                                                    int E, int K, float* dfeat, int ld_dfeat, void* stream) {
  DG_REQUIRE(w && nstar && gdz && a && dfeat, DGCNN_E_INVALID, "global_feature_scatter_det: null pointer");
  DG_REQUIRE(B > 0 && E > 0 && K > 0 && K % 4 == 0 && ld_w % 4 == 0 && ld_dfeat % 4 == 0 &&
                 ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(dfeat)) & 15) == 0,
             DGCNN_E_INVALID, "global_feature_scatter_det: K and the row strides must be multiples of 4, 16-byte alignment");
  const long long warps = static_cast<long long>(B) * E;
  gf_scatter_det_kernel<<<static_cast<int>((warps * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(w, ld_w, nstar, gdz, a, B, E, K,
                                                                                                            dfeat, ld_dfeat);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}
