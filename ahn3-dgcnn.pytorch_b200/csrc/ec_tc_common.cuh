// ec_tc_common.cuh — helpers shared by the tensor-core EdgeConv kernels (edgeconv_tc.cu, edgeconv_pipe.cu).
#pragma once
#include <cuda_bf16.h>
#include "edgeconv.cuh"
#include "tc.cuh"

namespace dg {

constexpr uint32_t kPlA = 128 * 128;   // bytes of one A plane (128 rows x 64 bf16)
constexpr uint32_t kPlW = 64 * 128;    // bytes of one weight plane
constexpr int kYs = 68;                // row stride of the FP32 tile parked over the A planes: 16-byte row accesses
                                       // (thread = row) and 4-byte column accesses (thread = column) are both conflict-free

// 8 consecutive channels (chunk ch) of tile row r -> NP bf16 planes (x = x1 + x2 (+ x3)), swizzled.
// Packed conversions: one cvt.rn.bf16x2 per pair and plane; the residual is exact in FP32.
template <int NP>
__device__ __forceinline__ void store_split(unsigned char* planes, uint32_t plane_bytes, int r, int ch, float (&v)[8]) {
  const uint32_t off = static_cast<uint32_t>(r) * 128u + (static_cast<uint32_t>((ch ^ (r & 7)) & 7) << 4);
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    __nv_bfloat162 h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
      if (p + 1 < NP) {
        const float2 f = __bfloat1622float2(h[e]);
        v[2 * e] -= f.x;
        v[2 * e + 1] -= f.y;
      }
    }
    *reinterpret_cast<uint4*>(planes + p * plane_bytes + off) = *reinterpret_cast<const uint4*>(h);
  }
}
__device__ __forceinline__ void store_split3(unsigned char* planes, uint32_t plane_bytes, int r, int ch, float (&v)[8]) {
  store_split<3>(planes, plane_bytes, r, ch, v);
}

// MN-major SWIZZLE_128B descriptor: lbo = byte distance between 64-wide MN atoms, 8-row K groups 1024 B apart
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t saddr, uint32_t lbo) {
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

// column sums over the 32 lanes of a warp for 32 per-lane values: afterwards v[0] of lane l is the sum of column l
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[32], int lane) {
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

__device__ __forceinline__ void ld4v(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
// 8 consecutive floats (32-byte aligned in shared memory) as two 16-byte loads
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// ---- small synchronisation helpers of the warp-specialised kernels -----------------------------------------
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// mbarrier wait with a watchdog: a protocol bug traps (reported as a launch failure) instead of hanging the GPU.
// (nanosleep back-off was measured slower: the hand-overs of the pipelined kernels are latency-critical.)
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(tc::smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {      // suspended in hardware until the phase flips (or 20 us pass)
    if (++spins > (1u << 20)) __trap();
  }
}
// 256-bit global load (sm_100): 8 consecutive floats, 32-byte aligned.  When one lane reads a whole row, 16-byte loads
// request every 32-byte sector twice; this fetches it once.
__device__ __forceinline__ void ld_global_v8(const float* src, float* v) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(src));
}
// 256-bit global store (sm_100): 8 consecutive floats, 32-byte aligned
__device__ __forceinline__ void st_global_v8(float* dst, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

}  // namespace dg
