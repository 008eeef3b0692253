/* debug_lab.h — descriptor / tensor-memory lab entry points.  NOT part of the product C-ABI (include/dgcnn_b200.h):
 * they are exported for tools/mn_lab.py and tools/tmem_a_lab.py only, which pinned the MN-major descriptor fields and the
 * TMEM A-operand layout on the GPU. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
/* exercises MN-major tcgen05 operand descriptors on one 128x64 tile pair */
int dgcnn_debug_mn_lab(const float* X, const float* Y, float* D, int lbo_a, int sbo_a, int lbo_b, int sbo_b,
                       int kstep, int a_major, int b_major, void* stream);
/* tcgen05.mma with the A operand in tensor memory on one 128x64x64 tile */
int dgcnn_debug_tmem_a_lab(const float* X, const float* Y, float* D, int swap, void* stream);
#ifdef __cplusplus
}
#endif
