// select.cuh — streaming top-K selection for 128 query rows of a CTA.
//
// Producers (four threads per query row, 16 keys of each tile each, scores in registers) append every candidate that
// beats the row's current threshold to a private sub-buffer in shared memory (no atomics).  Between tiles, warps
// compact the rows in which a sub-buffer is filling up: bitonic sort of the buffered keys (shuffles), bitonic merge with the row's sorted list of 64
// (2 keys per lane), new threshold = key at position kp-1.  Candidates are 64-bit keys (ordered score << 32 | ~idx),
// so ties on the score prefer the lower index and all keys are distinct.
#pragma once
#include "common.cuh"

namespace dg {
namespace sel {

constexpr int kRows = 128;
constexpr int kList = 64;            // sorted list length per row (2 per lane)
constexpr int kSub = 4;              // producers per row (each owns a private sub-buffer: no atomics)
constexpr int kSubCap = 16;          // candidates per sub-buffer (= keys one producer sees per tile)
constexpr int kBufCap = kSub * kSubCap;
constexpr int kBufLd = kBufCap + 1;  // row stride of the buffer (u64): producer stores of a warp hit distinct banks

typedef unsigned long long u64;

struct Smem {
  u64* lists;        // [kRows][2][32]
  u64* buf;          // [kRows][kBufLd]
  int* cnt;          // [kRows][kSub]
  float* thr;        // [kRows]  current pass threshold (score of list position kp-1, or -inf)
  int* overflow;     // [kRows]  a candidate was dropped because the buffer was full -> row needs the exact fallback
  int* flag;         // [kRows]  row is queued for compaction
  int* work;         // [2][kRows] double-buffered worklists of rows to compact
  int* nwork;        // [2 (+2 pad)]
};
constexpr size_t kSmemBytes = static_cast<size_t>(kRows) * (kList * 8 + kBufLd * 8 + 4 * kSub + 4 + 4 + 4 + 8) + 16;

__device__ __forceinline__ Smem carve(unsigned char* base) {   // base 8-byte aligned
  Smem s;
  s.lists = reinterpret_cast<u64*>(base);
  s.buf = s.lists + kRows * kList;
  s.cnt = reinterpret_cast<int*>(s.buf + kRows * kBufLd);
  s.thr = reinterpret_cast<float*>(s.cnt + kRows * kSub);
  s.overflow = reinterpret_cast<int*>(s.thr + kRows);
  s.flag = s.overflow + kRows;
  s.work = s.flag + kRows;
  s.nwork = s.work + 2 * kRows;
  return s;
}

__device__ __forceinline__ void init(const Smem& s, int tid, int nthreads) {
  for (int i = tid; i < kRows * kList; i += nthreads) s.lists[i] = 0ull;
  for (int i = tid; i < kRows * kSub; i += nthreads) s.cnt[i] = 0;
  for (int i = tid; i < kRows; i += nthreads) {
    s.thr[i] = -INFINITY;
    s.overflow[i] = 0;
    s.flag[i] = 0;
  }
  if (tid < 4) s.nwork[tid] = 0;
}

__device__ __forceinline__ float key_score(u64 key) {
  const uint32_t u = static_cast<uint32_t>(key >> 32);
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// producer side: the thread owning (row, sub) appends to its private sub-buffer; `n` is its running count
// (loaded from cnt[row][sub] before the tile, stored back after it).
__device__ __forceinline__ void push(const Smem& s, int row, int sub, int& n, float score, int j) {
  if (n < kSubCap) {
    s.buf[row * kBufLd + sub * kSubCap + n] = make_key(score, j);
    ++n;
  } else {
    s.overflow[row] = 1;
  }
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

// consumer side: one warp merges the buffer of `row` into its list and refreshes the threshold.
__device__ __forceinline__ void compact_row(const Smem& s, int row, int lane, int kp) {
  // flat index e = r*32 + lane over the valid entries of the four sub-buffers
  const int4 c4 = *reinterpret_cast<const int4*>(s.cnt + row * kSub);
  const int c0 = min(c4.x, kSubCap), c1 = min(c4.y, kSubCap), c2 = min(c4.z, kSubCap), c3 = min(c4.w, kSubCap);
  const int n01 = c0 + c1, n012 = n01 + c2, nb = n012 + c3;
  u64 e0 = 0ull, e1 = 0ull;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int e = lane + 32 * r;
    if (e < nb) {
      const int sub = e < c0 ? 0 : (e < n01 ? 1 : (e < n012 ? 2 : 3));
      const int slot = e - (sub == 0 ? 0 : (sub == 1 ? c0 : (sub == 2 ? n01 : n012)));
      const u64 v = s.buf[row * kBufLd + sub * kSubCap + slot];
      if (r == 0) e0 = v; else e1 = v;
    }
  }
  if (nb > 32) {
    sort64_desc(e0, e1, lane);
  } else {
    sort32_desc(e0, lane);       // the common case: half the network
  }
  // top-64 of (list U buffer): m[i] = max(list[i], buffer_sorted[63 - i]) is bitonic and holds it
  const u64 r0 = __shfl_sync(0xffffffffu, e1, 31 - lane);
  const u64 r1 = __shfl_sync(0xffffffffu, e0, 31 - lane);
  u64 l0 = s.lists[(row * 2 + 0) * 32 + lane];
  u64 l1 = s.lists[(row * 2 + 1) * 32 + lane];
  l0 = l0 > r0 ? l0 : r0;
  l1 = l1 > r1 ? l1 : r1;
  merge64_desc(l0, l1, lane);
  s.lists[(row * 2 + 0) * 32 + lane] = l0;
  s.lists[(row * 2 + 1) * 32 + lane] = l1;
  // the pass threshold only ever tightens (kp may be a smaller "bootstrap" position on the first tile)
  const u64 kth = list_at(l0, l1, kp - 1);
  if (lane == 0 && kth != 0ull) s.thr[row] = fmaxf(s.thr[row], key_score(kth));
  if (lane < kSub) s.cnt[row * kSub + lane] = 0;
}

// producer side, after a tile: queue the row for compaction if this producer's sub-buffer reached `trigger`
__device__ __forceinline__ void request_compaction(const Smem& s, int row, int n, int trigger, int parity) {
  if (n >= trigger && atomicExch(&s.flag[row], 1) == 0) s.work[parity * kRows + atomicAdd(&s.nwork[parity], 1)] = row;
}
// consumer side: the queued rows are dealt round-robin to the CTA's warps (balanced, unlike static row ownership)
__device__ __forceinline__ void compact_queued(const Smem& s, int warp, int nwarps, int lane, int kp, int parity) {
  const int nw = s.nwork[parity];
  for (int i = warp; i < nw; i += nwarps) {
    const int row = s.work[parity * kRows + i];
    compact_row(s, row, lane, kp);
    if (lane == 0) s.flag[row] = 0;
  }
}

}  // namespace sel
}  // namespace dg
