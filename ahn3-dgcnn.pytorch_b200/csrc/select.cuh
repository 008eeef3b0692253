// select.cuh — warp-level selection helpers of the kNN kernels.
//
// Candidates are 64-bit keys (ordered score bits << 32 | ~index, see make_key in common.cuh): ties on the score prefer
// the lower index and all keys of a row are distinct, so sorting keys gives the reference's top-k order directly.
// A row's (<= 64) re-scored survivors are sorted by a bitonic network over the warp, two keys per lane.
#pragma once
#include "common.cuh"

namespace dg {
namespace sel {

typedef unsigned long long u64;

__device__ __forceinline__ float key_score(u64 key) {
  const uint32_t u = static_cast<uint32_t>(key >> 32);
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// ---- warp-level bitonic networks on keys distributed as element i = r*32 + lane ---------------------------
__device__ __forceinline__ void cmpx(u64& v, u64 pv, bool keep_max) {
  const bool gt = v > pv;
  v = (gt == keep_max) ? v : pv;
}
__device__ __forceinline__ void sort32_desc(u64& a0, int lane) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int d = size >> 1; d >= 1; d >>= 1) {
      const u64 p0 = __shfl_xor_sync(0xffffffffu, a0, d);
      cmpx(a0, p0, ((lane & d) == 0) == ((lane & size) == 0));
    }
  }
}
__device__ __forceinline__ void merge64_desc(u64& a0, u64& a1, int lane) {   // input bitonic, output sorted descending
  const u64 mx = a0 > a1 ? a0 : a1, mn = a0 > a1 ? a1 : a0;
  a0 = mx;
  a1 = mn;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    const u64 p0 = __shfl_xor_sync(0xffffffffu, a0, d), p1 = __shfl_xor_sync(0xffffffffu, a1, d);
    const bool lower = (lane & d) == 0;
    cmpx(a0, p0, lower);
    cmpx(a1, p1, lower);
  }
}
__device__ __forceinline__ void sort64_desc(u64& a0, u64& a1, int lane) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int d = size >> 1; d >= 1; d >>= 1) {
      const u64 p0 = __shfl_xor_sync(0xffffffffu, a0, d), p1 = __shfl_xor_sync(0xffffffffu, a1, d);
      const bool lower = (lane & d) == 0;
      const bool desc0 = (lane & size) == 0;              // element lane
      const bool desc1 = ((lane + 32) & size) == 0;       // element lane + 32 (ascending block when size == 32)
      cmpx(a0, p0, lower == desc0);
      cmpx(a1, p1, lower == desc1);
    }
  }
  merge64_desc(a0, a1, lane);
}

__device__ __forceinline__ u64 list_at(u64 l0, u64 l1, int pos) {
  const u64 w0 = __shfl_sync(0xffffffffu, l0, pos & 31);
  const u64 w1 = __shfl_sync(0xffffffffu, l1, pos & 31);
  return (pos >> 5) ? w1 : w0;
}

}  // namespace sel
}  // namespace dg
