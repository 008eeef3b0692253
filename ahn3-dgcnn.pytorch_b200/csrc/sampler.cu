// sampler.cu — the input pipeline's per-block work on device (data.py:107-143 `sample_points`, data.py:169-182 `collate_pcs`
// and the `permute(0, 2, 1).float()` of train.py:176), for tiles that are resident in HBM.
//
// The reference draws `num_point` points of a tile WITHOUT replacement with probabilities inversely related to the class
// frequencies (`np.random.choice(..., p=point_weights, replace=False)`) in Python loops over every point, then copies the
// selected rows into a float64 batch on the host.  Here:
//   1. class histogram of the tile's labels                                             (sample_hist_kernel)
//   2. one exponential race key per point, key_i = -log(u_i) / w_class(i)               (sample_keys_kernel)
//      — the n smallest keys are a weighted sample without replacement (Efraimidis & Spirakis 2006), i.e. successive
//      draws proportional to the remaining weights, which is what `choice(replace=False, p)` samples;
//   3. the n-th smallest key by a 4-pass radix select over the key bits (HBM-bound: the 4-byte keys are re-read per pass)
//      and a deterministic two-level compaction of the selected indices (index order)   (sample_digit_*, sample_compact_*)
//   4. gather of the selected rows straight into the model's [F, n] float32 slot and the int64 label row   (gather kernel)
// Counter-based randomness (splitmix64 of seed and point index): the result is a pure function of (labels, seed).
// Everything is enqueued on the stream; nothing is synchronised.
#include "common.cuh"

namespace dg {

constexpr int kSampMaxClasses = 32;
constexpr int kSampChunk = 2048;            // points per CTA of the compaction passes (256 threads x 8)

struct SampleState {                         // lives in the workspace
  unsigned int counts[kSampMaxClasses];
  unsigned int hist[256];
  unsigned int prefix;                       // key bits decided so far (high to low)
  unsigned int rank;                         // 0-based rank of the wanted key among the keys matching the prefix
  unsigned int n_less;                       // keys strictly below the final threshold
  unsigned int need_equal;                   // keys equal to the threshold that are selected (lowest indices first)
  unsigned int pad[4];
};

__global__ void sample_init_kernel(SampleState* st, unsigned int rank) {
  unsigned int* w = reinterpret_cast<unsigned int*>(st);
  for (int i = threadIdx.x; i < static_cast<int>(sizeof(SampleState) / 4); i += blockDim.x) w[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) st->rank = rank;       // 0-based rank of the largest selected key
}

__global__ void sample_hist_kernel(const int* __restrict__ labels, int Nt, int num_classes, SampleState* st) {
  __shared__ unsigned int h[kSampMaxClasses];
  if (threadIdx.x < kSampMaxClasses) h[threadIdx.x] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Nt; i += gridDim.x * blockDim.x) {
    const int c = labels[i];
    if (c >= 0 && c < num_classes) atomicAdd(&h[c], 1u);
  }
  __syncthreads();
  if (threadIdx.x < num_classes && h[threadIdx.x]) atomicAdd(&st->counts[threadIdx.x], h[threadIdx.x]);
}

__device__ __forceinline__ double uniform01(unsigned long long seed, unsigned long long i) {
  unsigned long long z = seed + (i + 1ull) * 0x9E3779B97F4A7C15ull;            // splitmix64
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (static_cast<double>(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);    // (0, 1), 53 bits
}

// data.py:124-136: weights_c = (tot - count_c) / sum_c (tot - count_c); every point of a class carries the class weight;
// NaN (a single-class tile) falls back to uniform.  Points with weight 0 or an out-of-range label get key = +inf.
__global__ void sample_keys_kernel(const int* __restrict__ labels, int Nt, int num_classes, unsigned long long seed,
                                   const SampleState* st, unsigned int* __restrict__ keys) {
  __shared__ float inv_w[kSampMaxClasses];
  if (threadIdx.x == 0) {
    double tot = 0.0, shuf = 0.0;
    for (int c = 0; c < num_classes; ++c) tot += st->counts[c];
    for (int c = 0; c < num_classes; ++c) shuf += tot - st->counts[c];
    for (int c = 0; c < num_classes; ++c) {
      const double w = shuf > 0.0 ? (tot - st->counts[c]) / shuf : 1.0;          // uniform fallback (data.py:137-138)
      inv_w[c] = w > 0.0 ? static_cast<float>(1.0 / w) : INFINITY;
    }
  }
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Nt; i += gridDim.x * blockDim.x) {
    const int c = labels[i];
    float key = INFINITY;
    if (c >= 0 && c < num_classes) key = static_cast<float>(-log(uniform01(seed, static_cast<unsigned long long>(i)))) * inv_w[c];
    keys[i] = __float_as_uint(key);            // keys are >= 0: the unsigned bit pattern orders like the float
  }
}

// histogram of the 8-bit digit at `shift` over the keys whose higher bits equal the prefix
__global__ void sample_digit_hist_kernel(const unsigned int* __restrict__ keys, int Nt, int shift, SampleState* st) {
  __shared__ unsigned int h[256];
  h[threadIdx.x] = 0;                          // blockDim.x == 256
  __syncthreads();
  const unsigned int prefix = st->prefix;
  const unsigned int himask = shift == 24 ? 0u : ~((1u << (shift + 8)) - 1u);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Nt; i += gridDim.x * blockDim.x) {
    const unsigned int k = keys[i];
    if ((k & himask) == prefix) atomicAdd(&h[(k >> shift) & 255u], 1u);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
}

// pick the digit bucket that holds the wanted rank, extend the prefix, clear the histogram (one thread: 256 bins)
__global__ void sample_digit_pick_kernel(int shift, SampleState* st) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned int rank = st->rank, below = 0;
  int d = 0;
  for (; d < 256; ++d) {
    const unsigned int c = st->hist[d];
    if (rank < below + c) break;
    below += c;
  }
  if (d == 256) d = 255;                        // (rank beyond the population: caller's n > usable points)
  st->n_less += below;
  st->rank = rank - below;
  st->prefix |= static_cast<unsigned int>(d) << shift;
  if (shift == 0) st->need_equal = st->rank + 1;               // among keys equal to the threshold take rank + 1 of them
  for (int i = 0; i < 256; ++i) st->hist[i] = 0;
}

// compaction pass 1: per chunk, how many keys are below / equal to the threshold
__global__ void __launch_bounds__(256) sample_compact_count_kernel(const unsigned int* __restrict__ keys, int Nt, const SampleState* st,
                                                                   unsigned int* __restrict__ chunk_counts) {
  __shared__ unsigned int red[2][8];
  const unsigned int thr = st->prefix;
  unsigned int less = 0, equal = 0;
  const int base = blockIdx.x * kSampChunk;
#pragma unroll
  for (int j = 0; j < kSampChunk / 256; ++j) {
    const int i = base + j * 256 + threadIdx.x;
    if (i < Nt) {
      const unsigned int k = keys[i];
      less += k < thr;
      equal += k == thr;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    less += __shfl_xor_sync(0xffffffffu, less, o);
    equal += __shfl_xor_sync(0xffffffffu, equal, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = less; red[1][threadIdx.x >> 5] = equal; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int a = 0, b = 0;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
    chunk_counts[2 * blockIdx.x] = a;
    chunk_counts[2 * blockIdx.x + 1] = b;
  }
}
// exclusive prefix over the chunks (one CTA; the chunk count is Nt / 2048)
__global__ void __launch_bounds__(1024) sample_compact_scan_kernel(unsigned int* __restrict__ chunk_counts, int num_chunks) {
  __shared__ unsigned int carry[2];
  __shared__ unsigned int wsum[2][32];
  if (threadIdx.x == 0) carry[0] = carry[1] = 0;
  __syncthreads();
  for (int c0 = 0; c0 < num_chunks; c0 += 1024) {
    const int c = c0 + threadIdx.x;
    unsigned int v[2] = {c < num_chunks ? chunk_counts[2 * c] : 0u, c < num_chunks ? chunk_counts[2 * c + 1] : 0u};
    unsigned int inc[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      unsigned int x = v[q];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o) x += y;
      }
      inc[q] = x;
      if ((threadIdx.x & 31) == 31) wsum[q][threadIdx.x >> 5] = x;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        unsigned int x = wsum[q][threadIdx.x];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
          if (threadIdx.x >= o) x += y;
        }
        wsum[q][threadIdx.x] = x;                 // inclusive over the warps
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const unsigned int wbase = (threadIdx.x >> 5) ? wsum[q][(threadIdx.x >> 5) - 1] : 0u;
      if (c < num_chunks) chunk_counts[2 * c + q] = carry[q] + wbase + inc[q] - v[q];      // exclusive
    }
    __syncthreads();
    if (threadIdx.x == 0) { carry[0] += wsum[0][31]; carry[1] += wsum[1][31]; }
    __syncthreads();
  }
}
// compaction pass 2: write the selected point indices — all keys below the threshold in index order, then the first
// `need_equal` keys equal to it
__global__ void __launch_bounds__(256) sample_compact_write_kernel(const unsigned int* __restrict__ keys, int Nt, const SampleState* st,
                                                                   const unsigned int* __restrict__ chunk_offsets, int n_samples,
                                                                   int* __restrict__ out_idx) {
  __shared__ unsigned int wl[8], we[8];
  const unsigned int thr = st->prefix, n_less = st->n_less, need_equal = st->need_equal;
  const int base = blockIdx.x * kSampChunk;
  unsigned int run_l = chunk_offsets[2 * blockIdx.x], run_e = chunk_offsets[2 * blockIdx.x + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = 0; j < kSampChunk / 256; ++j) {
    const int i = base + j * 256 + threadIdx.x;
    const unsigned int k = i < Nt ? keys[i] : 0xFFFFFFFFu;
    const bool l = i < Nt && k < thr, e = i < Nt && k == thr;
    const unsigned int ml = __ballot_sync(0xffffffffu, l), me = __ballot_sync(0xffffffffu, e);
    if (lane == 0) { wl[warp] = __popc(ml); we[warp] = __popc(me); }
    __syncthreads();
    unsigned int bl = 0, be = 0, tl = 0, te = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (w < warp) { bl += wl[w]; be += we[w]; }
      tl += wl[w]; te += we[w];
    }
    if (l) {
      const unsigned int pos = run_l + bl + __popc(ml & ((1u << lane) - 1u));
      if (pos < static_cast<unsigned int>(n_samples)) out_idx[pos] = i;
    } else if (e) {
      const unsigned int r = run_e + be + __popc(me & ((1u << lane) - 1u));
      if (r < need_equal && n_less + r < static_cast<unsigned int>(n_samples)) out_idx[n_less + r] = i;
    }
    run_l += tl;
    run_e += te;
    __syncthreads();
  }
}

// collate + permute(0,2,1).float(): x[f, j] = points[idx[j], f] (row stride ldp) for f < F; y[j] = labels[idx[j]]
template <typename T>
__global__ void sample_gather_kernel(const T* __restrict__ points, int ldp, const int* __restrict__ labels, const int* __restrict__ idx,
                                     int n, int F, float* __restrict__ x, long long x_sf, long long x_sn,
                                     long long* __restrict__ y) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int i = idx[j];
  for (int f = 0; f < F; ++f) x[f * x_sf + j * x_sn] = static_cast<float>(points[static_cast<size_t>(i) * ldp + f]);
  if (y) y[j] = labels[i];
}

}  // namespace dg

using namespace dg;

static size_t sampler_layout(int Nt, size_t* o_keys, size_t* o_chunks) {
  size_t off = (sizeof(SampleState) + 255) & ~size_t(255);
  *o_keys = off;
  off += (static_cast<size_t>(Nt) * 4 + 255) & ~size_t(255);
  *o_chunks = off;
  off += static_cast<size_t>((Nt + kSampChunk - 1) / kSampChunk) * 8 + 256;
  return off;
}

extern "C" size_t dgcnn_sample_points_workspace_bytes(int Nt) {
  if (Nt <= 0) return 0;
  size_t a, b;
  return sampler_layout(Nt, &a, &b) + 256;
}

extern "C" int dgcnn_sample_points(const int32_t* labels, int Nt, int num_classes, int n_samples, uint64_t seed, int32_t* out_idx,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  DG_REQUIRE(labels && out_idx && workspace, DGCNN_E_INVALID, "sample_points: null pointer");
  DG_REQUIRE(Nt > 0 && n_samples > 0 && n_samples <= Nt, DGCNN_E_INVALID,
             "sample_points: cannot take %d samples without replacement from %d points", n_samples, Nt);
  DG_REQUIRE(num_classes >= 1 && num_classes <= kSampMaxClasses, DGCNN_E_UNSUPPORTED, "sample_points: at most %d classes", kSampMaxClasses);
  DG_REQUIRE(workspace_bytes >= dgcnn_sample_points_workspace_bytes(Nt), DGCNN_E_WORKSPACE, "sample_points: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned char* ws = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  size_t o_keys, o_chunks;
  sampler_layout(Nt, &o_keys, &o_chunks);
  SampleState* state = reinterpret_cast<SampleState*>(ws);
  unsigned int* keys = reinterpret_cast<unsigned int*>(ws + o_keys);
  unsigned int* chunks = reinterpret_cast<unsigned int*>(ws + o_chunks);
  const int grid = std::max(1, std::min((Nt + 255) / 256, 8 * std::max(1, num_sms())));
  const int num_chunks = (Nt + kSampChunk - 1) / kSampChunk;
  sample_init_kernel<<<1, 128, 0, st>>>(state, static_cast<unsigned int>(n_samples - 1));
  sample_hist_kernel<<<grid, 256, 0, st>>>(labels, Nt, num_classes, state);
  sample_keys_kernel<<<grid, 256, 0, st>>>(labels, Nt, num_classes, seed, state, keys);
  for (int shift = 24; shift >= 0; shift -= 8) {
    sample_digit_hist_kernel<<<grid, 256, 0, st>>>(keys, Nt, shift, state);
    sample_digit_pick_kernel<<<1, 32, 0, st>>>(shift, state);
  }
  sample_compact_count_kernel<<<num_chunks, 256, 0, st>>>(keys, Nt, state, chunks);
  sample_compact_scan_kernel<<<1, 1024, 0, st>>>(chunks, num_chunks);
  sample_compact_write_kernel<<<num_chunks, 256, 0, st>>>(keys, Nt, state, chunks, n_samples, out_idx);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}

extern "C" int dgcnn_gather_points(const void* points, int points_are_f64, int ld_points, const int32_t* labels, const int32_t* idx,
                                   int n, int F, float* x, int64_t x_stride_f, int64_t x_stride_n, int64_t* y, void* stream) {
  DG_REQUIRE(points && idx && x && (labels || !y), DGCNN_E_INVALID, "gather_points: null pointer");
  DG_REQUIRE(n > 0 && F > 0 && ld_points >= F, DGCNN_E_INVALID, "gather_points: bad shape n=%d F=%d ld=%d", n, F, ld_points);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long* yy = reinterpret_cast<long long*>(y);
  if (points_are_f64)
    sample_gather_kernel<double><<<(n + 255) / 256, 256, 0, st>>>(static_cast<const double*>(points), ld_points, labels, idx, n, F, x,
                                                                  x_stride_f, x_stride_n, yy);
  else
    sample_gather_kernel<float><<<(n + 255) / 256, 256, 0, st>>>(static_cast<const float*>(points), ld_points, labels, idx, n, F, x,
                                                                 x_stride_f, x_stride_n, yy);
  DG_CUDA(cudaGetLastError());
  return DGCNN_OK;
}
