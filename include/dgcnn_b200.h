/* dgcnn_b200.h — C-ABI of libdgcnn_b200.so: the B200 (sm_100a) EdgeConv hot path of DGCNN.
 *
 * The reference (BenCurran98/AHN3-dgcnn.pytorch) has no FFI: its boundary is the Python surface of
 * model.py.  These entry points are what a ctypes binding for that path binds; the in-tree Python
 * host (`ahn3-dgcnn.pytorch_b200/model.py`) is such a binding and INTEGRATION.md shows the stub a
 * maintainer of the reference would add.  Each function names the reference lines it replaces.
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless stated; nothing is allocated, freed or synchronised here.
 *  - Work is enqueued on `stream` (a cudaStream_t passed as void*); NULL = legacy default stream.
 *  - "rows" = point-major feature matrix: row p = b*N + n holds the C channels of point n of block b,
 *    consecutive floats, `ld` floats between rows (ld >= C, ld % 4 == 0, 16-byte aligned base).
 *  - idx = int32 [B*N, k], neighbour indices LOCAL to the block (0..N-1), unordered (a set).
 *  - Return value: 0 ok; <0 error (DGCNN_E_*).  dgcnn_last_error() gives a thread-local message.
 *  - Every *_partials / *_ws buffer is caller-provided scratch sized by the matching *_bytes() call.
 */
#ifndef DGCNN_B200_H
#define DGCNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGCNN_OK 0
#define DGCNN_E_INVALID (-1)     /* bad argument (null pointer, k > N, misaligned ld ...)            */
#define DGCNN_E_UNSUPPORTED (-2) /* shape outside the compiled range (C > 128, k > 128 ...)           */
#define DGCNN_E_CUDA (-3)        /* a CUDA runtime call failed; message holds cudaGetErrorString      */
#define DGCNN_E_WORKSPACE (-4)   /* workspace too small                                                */

#define DGCNN_COUT 64            /* every EdgeConv conv has 64 output channels (model.py:116-130)     */

int dgcnn_version(void);                 /* 100*major + minor */
const char* dgcnn_last_error(void);      /* host pointer, thread-local, never NULL */
int dgcnn_num_sms(void);                 /* SM count of the current device (cached), <0 on error */

/* ---- layout: x[B,C,N] with arbitrary element strides  ->  rows[B*N, ld] (cols C..Cpad-1 zeroed) ----
 * replaces x.transpose(2,1).contiguous() at model.py:73.  Cpad = C rounded up to a multiple of 4.    */
int dgcnn_pack_rows_f32(const float* x, int64_t stride_b, int64_t stride_c, int64_t stride_n,
                        int B, int C, int N, float* rows, int ld, void* stream);
/* rows[B*N, ld] (C channels) -> y[B,C,N] contiguous (for the depth=1..3 early returns, model.py:151-165) */
int dgcnn_unpack_rows_f32(const float* rows, int ld, int B, int C, int N, float* y, void* stream);

/* ---- kNN graph (model.py:25-40) -------------------------------------------------------------------
 * score[i,j] = fl(fl(2*dot(x_i,x_j) - |x_j|^2) - |x_i|^2) with dot an FP32 FMA chain over channels in
 * increasing order and |x|^2 torch.sum's order (sequential in blocks of 16) — the exact roundings of the
 * reference's CPU path; idx[p,:] = the k largest per row (ties: lower index wins), self included.
 * `norms` = scratch float[B*N].  Exact FP32 path for every C <= 128.                                  */
int dgcnn_knn_f32(const float* rows, int ld, int B, int C, int N, int k,
                  int32_t* idx, float* norms, void* stream);

/* ---- tensor-core kNN (model.py:25-40; every graph of the model: C = 3..4 input graph, C = 64 feature graphs, C <= 128) ---
 * Same result as dgcnn_knn_f32, bit for bit: tcgen05 Gram tiles on a 2-plane bf16 split of the block-centred features act
 * as a filter (threshold from a sample of the keys, ~100 candidates per query, a rigorous error bound), the survivors are
 * re-scored with the exact FP32 expression, rows that cannot be certified are redone by brute force.
 * ordered != 0: rows best-first like the reference's topk (knn(), get_graph_feature()); ordered == 0: the same k
 * neighbours in unspecified order (all the EdgeConv blocks need; only the few keys whose approximate score is within
 * the error bound of the k-th are re-scored, 10x less gather traffic).
 * Needs C <= 128, k <= 48, N >= 128.  workspace: dgcnn_knn_tc_workspace_bytes(B, N) bytes (about 2.1 KB per point).   */
int dgcnn_knn_tc_supported(int C, int N, int k);
size_t dgcnn_knn_tc_workspace_bytes(int B, int N);
int dgcnn_knn_tc_f32(const float* rows, int ld, int B, int C, int N, int k, int ordered, int32_t* idx,
                     void* workspace, size_t workspace_bytes, void* stream);
/* Diagnostics of the last dgcnn_knn_tc_f32 call that used `workspace` (same B, N, C): out5 (HOST int[5]) = rows redone by
 * brute force | sub-buffer overflows | too few candidates | too many near-ties | certificate failures.  Synchronises.  */
int dgcnn_knn_tc_fallback_stats(const void* workspace, int B, int N, int C, int* out5, void* stream);

/* ---- low-dimensional kNN (C <= 16: the xyz/agl input graph), thread per query, exact FP32 scores, threshold-collect-
 * select; same result as dgcnn_knn_f32 (rows that cannot be resolved in the fast path are redone by brute force).
 * Needs C <= 16, k <= 48, N >= 512.                                                                       */
int dgcnn_knn_small_supported(int C, int N, int k);
size_t dgcnn_knn_small_workspace_bytes(int B, int N);
int dgcnn_knn_small_f32(const float* rows, int ld, int B, int C, int N, int k, int32_t* idx,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- get_graph_feature (model.py:43-83), materialising form kept for API parity -------------------
 * out[B,N,k,2C] = cat(x_j - x_i, x_i); the reference returns permute(0,3,1,2) of this buffer.        */
int dgcnn_graph_feature_f32(const float* rows, int ld, const int32_t* idx, int B, int C, int N, int k,
                            float* out, void* stream);

/* ---- EdgeConv block ---------------------------------------------------------------------------------
 * One block = gather + [x_j-x_i, x_i] + conv_a(2C->64) + BN + LeakyReLU(0.2) [+ conv_b(64->64) + BN +
 * LeakyReLU] + max over k (model.py:147-150, 154-157, 161-163).  Weights are the Conv2d tensors as
 * stored: w_a[64, 2C] row-major, w_b[64, 64] (NULL for the single-conv block 3).
 * The [B,2C,N,k] edge tensor is never formed in HBM.                                                   */
typedef struct dgcnn_edgeconv_args {
  const float* rows;      /* input features, point-major                                              */
  int ld;                 /* row stride of `rows` (floats)                                            */
  const int32_t* idx;     /* [B*N, k] local neighbour indices                                         */
  int B, C, N, k;         /* C = input channels (<= 64)                                               */
  const float* w_a;       /* [64, 2C]                                                                 */
  const float* w_b;       /* [64, 64] or NULL                                                         */
  float* out;             /* [B*N, 64] block output, row stride ld_out                                */
  int ld_out;
  float* v_scratch;       /* optional float[B*N*64] scratch: enables the tensor-core kernels for C == 64
                             (holds the per-point centre term Wa2 x_i); NULL selects the FP32-FMA kernels   */
  int impl;               /* kernel family of the 64-channel passes: 0 = default (tcgen05, warp-specialised pipelined
                             conv_b), 1 = FP32-FMA kernels, 2 = serial tcgen05 kernels only.  Per call (no process-wide
                             state); 1 and 2 exist for A/B checks.  The tensor-core paths are FP32-accurate (bf16x3).   */
} dgcnn_edgeconv_args;

/* BN "affine" vectors are float[192] = mean[64] | scale[64] | beta[64]; the kernels apply
 * z = (y - mean) * scale + beta (one subtraction + one FMA, the accurate form when |mean| >> std).
 * eval: mean = running_mean, scale = gamma / sqrt(running_var + eps);  train: from dgcnn_bn_finalize.   */

/* Inference: single fused kernel (gather -> conv_a -> BN -> LReLU -> [conv_b -> BN -> LReLU] -> max_k). */
int dgcnn_edgeconv_fwd_eval(const dgcnn_edgeconv_args* a, const float* affine_a, const float* affine_b,
                            void* stream);

/* Training forward (batch-statistics BN; SURVEY.md §7 hard part 2).  Sequence per block:
 *   conv_a_train  -> bn_finalize(a) -> [conv_b_train -> bn_finalize(b)] -> select
 * y_a: float[B*N*k, 64] pre-BN conv_a output (kept for backward).  partials: scratch of
 * dgcnn_edgeconv_partials_bytes().  ymax/ymin float[B*N,64], amax/amin uint8[B*N,64] = per point/channel
 * extrema over k of the LAST conv's raw output and their slots (NULL allowed in conv_a_train when w_b != NULL). */
size_t dgcnn_edgeconv_partials_bytes(void);
int dgcnn_edgeconv_conv_a_train(const dgcnn_edgeconv_args* a, float* y_a, double* partials,
                                float* ymax, float* ymin, uint8_t* amax, uint8_t* amin, void* stream);
int dgcnn_edgeconv_conv_b_train(const dgcnn_edgeconv_args* a, const float* y_a, const float* affine_a,
                                double* partials, float* ymax, float* ymin, uint8_t* amax, uint8_t* amin,
                                void* stream);
/* Reduce the per-CTA partial sums (sum, sum of squares) in double, produce batch mean / rstd and the
 * affine triple, and update running_mean/var (momentum, unbiased var) like nn.BatchNorm2d (model.py:107-111;
 * running_* may both be NULL).  count = values per channel (B*N*k).  C must be 64.                      */
int dgcnn_bn_finalize(const double* partials, int C, double count, const float* gamma, const float* beta,
                      float eps, float momentum, float* running_mean, float* running_var,
                      float* mean, float* rstd, float* affine, void* stream);
/* out = LeakyReLU(BN(sel)), sel = scale>=0 ? ymax : ymin (BN-affine and LeakyReLU are monotone, so
 * max_k LReLU(BN(y)) = LReLU(BN(max_k y)) for scale >= 0 and LReLU(BN(min_k y)) otherwise);
 * ysel float[B*N,64] and arg uint8[B*N,64] (slot of sel) are saved for backward.                       */
int dgcnn_edgeconv_select(int M, const float* ymax, const float* ymin, const uint8_t* amax, const uint8_t* amin,
                          const float* affine, float* out, int ld_out, float* ysel, uint8_t* arg, void* stream);

/* Training backward (dense because of batch-stat BN; SURVEY.md §7 hard part 3).  Sequence:
 *   bwd_prep -> bn_bwd_reduce(last BN) -> [bwd_b -> bn_bwd_reduce(BN a)] -> bwd_a
 * then dX = r*Wa1 + s*(Wa2-Wa1), dWa = [(r-s)^T X, s^T X] are small point-level GEMMs done by the host.  */
/* g = dout * LeakyReLU'(BN(ysel)); partials <- (sum g, sum g*ysel) per channel.                        */
int dgcnn_edgeconv_bwd_prep(int M, const float* dout, int ld_dout, const float* ysel, const float* affine,
                            float* g, double* partials, void* stream);
/* From partial sums (S1 = sum dz, S2 = sum dz*y) of one BN: dgamma, dbeta and the coefficients
 * dy = cA*dz - cB - cC*y  of the BN input gradient.  coef = float[3*64] = cA | cB | cC.                 */
int dgcnn_bn_bwd_reduce(const double* partials, int C, double count, const float* gamma,
                        const float* mean, const float* rstd, float* dgamma, float* dbeta, float* coef,
                        void* stream);
/* Edge pass through conv_b: recompute h = LReLU(BN_a(y_a)), y_b = W_b h; dy_b from (g, arg, coef_b);
 * dW_b (float[64*64], overwritten); dz_a = (dy_b W_b) * LReLU'(.) -> float[B*N*k,64]; partials <- BN_a sums.
 * dw_partials: scratch of dgcnn_edgeconv_dw_partials_bytes().                                           */
size_t dgcnn_edgeconv_dw_partials_bytes(void);
int dgcnn_edgeconv_bwd_b(const dgcnn_edgeconv_args* a, const float* y_a, const float* affine_a,
                         const float* g, const uint8_t* arg, const float* coef_b,
                         float* dz_a, float* dw_b, double* partials, float* dw_partials, void* stream);
/* Edge pass through BN_a: dy_a = cA*dz_a - cB - cC*y_a; s[p,:] = sum_k dy_a; r[j,:] += dy_a (scatter).
 * dz_a == NULL selects the single-conv form (dz_a given implicitly by g/arg).  r must be zero-filled.
 * s and r are [B*N, 64] with row stride ld_rs (so they can be the two halves of one [B*N,128] GEMM operand). */
int dgcnn_edgeconv_bwd_a(const dgcnn_edgeconv_args* a, const float* y_a, const float* dz_a,
                         const float* g, const uint8_t* arg, const float* coef_a,
                         float* s, float* r, int ld_rs, void* stream);

/* ---- single-convolution block without a per-edge GEMM (conv5 + bn5 + LeakyReLU + max over k, model.py:128-130,161-163) ----
 * y[(i,j)] = W [x_j - x_i ; x_i] = U_j + V_i with U = X Wa1^T and V = X (Wa2 - Wa1)^T (ONE point-level GEMM by the caller,
 * X [Wa1 ; Wa2 - Wa1]^T), and max_k LReLU(BN(y)) = LReLU(BN(V_i + max_j U_j)) (min for a negative BN scale).
 * dgcnn_edgeconv_lin_extrema: one gather pass over U — per point and channel the extrema of fl(U_j + V_i) over the k
 * neighbours (ymax, ymin float[B*N,64], slots amax, amin uint8) and the batch sums of y, y^2 (partials, as
 * dgcnn_edgeconv_conv_a_train writes them; NULL in inference).  Follow with dgcnn_bn_finalize (training) and
 * dgcnn_edgeconv_select.  dgcnn_edgeconv_bwd_a_lin is dgcnn_edgeconv_bwd_a with y recomputed as U_j + V_i (dz_a ==
 * NULL: single-conv block, dz from g / arg; dz_a != NULL: two-conv block, the y_a read becomes an L2 gather).
 * For the single-conv block no per-edge tensor exists in HBM in either direction.  Only `idx`, B, N, k of the args are used.
 * Intended for the 64-channel feature-space block (O(1) features); rows of U, V: >= 64 floats, 16-byte aligned.       */
int dgcnn_edgeconv_lin_extrema(const dgcnn_edgeconv_args* a, const float* U, int ld_u, const float* V, int ld_v,
                               double* partials, float* ymax, float* ymin, uint8_t* amax, uint8_t* amin, void* stream);
/* first convolution of a TWO-conv block the same way: y_a[B*N*k,64] = U_j + V_i (kept per edge because batch-statistics BN
 * sits between the convolutions) and its batch sums; replaces dgcnn_edgeconv_conv_a_train for 64-channel inputs.    */
int dgcnn_edgeconv_lin_conv_a(const dgcnn_edgeconv_args* a, const float* U, int ld_u, const float* V, int ld_v,
                              float* y_a, double* partials, void* stream);
int dgcnn_edgeconv_bwd_a_lin(const dgcnn_edgeconv_args* a, const float* U, int ld_u, const float* V, int ld_v,
                             const float* dz_a, const float* g, const uint8_t* arg, const float* coef_a, float* s, float* r,
                             int ld_rs, void* stream);
/* the weight of that point-level GEMM and its adjoint: wa[64,2C] = [Wa1 | Wa2] (the reference's conv weight, model.py:116-130)
 * -> wlin[128,C] = [Wa1 ; Wa2 - Wa1];  G[128,C] = [r|s]^T X -> dwa[64,2C] = [G_r - G_s | G_s].  Dense, contiguous.      */
int dgcnn_edgeconv_lin_weight_f32(const float* wa, int C, float* wlin, void* stream);
int dgcnn_edgeconv_lin_weight_bwd_f32(const float* G, int C, float* dwa, void* stream);

/* ---- reproducible gradients (opt-in; csrc/graph_transpose.cu) ------------------------------------------------------------
 * The default backward adds per-edge gradients into their neighbour's row with float reductions: the summation order, and
 * with it the last bits of the gradients, varies from run to run.  The reproducible form writes the per-edge gradient
 * (dgcnn_edgeconv_bwd_a_lin_edges: dy_edges[B*N*k,64], s as usual), transposes the kNN table once per graph
 * (dgcnn_graph_transpose: in_start[B*N+1], in_edge[B*N*k] = the in-edges e = i*k+slot of every point, ascending) and
 * gathers r in that order (dgcnn_gather_in_edges_f32).  dgcnn_global_feature_scatter_det_f32 is the owner-per-row form of
 * the scatter inside dgcnn_global_feature_bwd_sparse_f32 (call that one with dfeat = NULL for t1 only).               */
int dgcnn_edgeconv_bwd_a_lin_edges(const dgcnn_edgeconv_args* a, const float* U, int ld_u, const float* V, int ld_v,
                                   const float* dz_a, const float* g, const uint8_t* arg, const float* coef_a, float* s,
                                   int ld_rs, float* dy_edges, void* stream);
size_t dgcnn_graph_transpose_workspace_bytes(int B, int N);
int dgcnn_graph_transpose(const int32_t* idx, int B, int N, int k, int32_t* in_start, int32_t* in_edge, void* workspace,
                          size_t workspace_bytes, void* stream);
int dgcnn_gather_in_edges_f32(const float* dy_edges, const int32_t* in_start, const int32_t* in_edge, int M, float* r, int ld_r,
                              void* stream);
int dgcnn_global_feature_scatter_det_f32(const float* w, int ld_w, const int32_t* nstar, const float* gdz, const float* a, int B,
                                         int E, int K, float* dfeat, int ld_dfeat, void* stream);

/* ---- backward of a two-conv block with NO per-edge gradient tensor in HBM (csrc/edgeconv_bwd_pipe.cu) -------------------
 * BN_a's backward is affine in (dz_a, y_a) and conv_a's backward only needs in-edge / out-edge sums, so the sums are taken
 * of the raw quantities and the affine map is applied per POINT afterwards:
 *   r_j = cA R_j - cnt_j cB - cC Yin_j,   s_i = cA S_i - k cB - cC Ysum_i
 * dgcnn_edgeconv_lin_conv_a_train: dgcnn_edgeconv_lin_conv_a + ysum[B*N,64], the sum of y_a over the out-edges of i.
 * dgcnn_edgeconv_bwd_b_fused: dgcnn_edgeconv_bwd_b (tcgen05, 2 bf16 planes) that, instead of writing dz_a[E,64], adds its
 *   rows into r_raw[j] and their per-point sums into s_raw[i] (row stride ld_rs), and the y_a rows it streams into
 *   yin[j] ([B*N,64], the sum of y_a over the in-edges of j) and cnt[j] ([B*N], the in-degree) — all float reductions:
 *   zero the four targets first.  k >= 8, N >= 16.
 * dgcnn_edgeconv_bwd_a_point_stats: the batch sums BN_a's backward needs, from point-level data:
 *   sum dz_a = sum_i S_i,  sum dz_a*y_a = sum_j U_j.R_j + sum_i V_i.S_i  -> partials (then dgcnn_bn_bwd_reduce -> coef_a).
 * dgcnn_edgeconv_bwd_a_point_apply: the affine map above, in place on r_raw / s_raw.                                  */
int dgcnn_edgeconv_lin_conv_a_train(const dgcnn_edgeconv_args* a, const float* U, int ld_u, const float* V, int ld_v,
                                    float* y_a, double* partials, float* ysum, void* stream);
int dgcnn_edgeconv_bwd_b_fused(const dgcnn_edgeconv_args* a, const float* y_a, const float* aff_a, const float* g,
                               const uint8_t* arg, const float* coef_b, float* r_raw, float* s_raw, int ld_rs, float* yin,
                               float* cnt, float* dw_b, float* dw_partials, void* stream);
int dgcnn_edgeconv_bwd_a_point_stats(int M, const float* U, int ld_u, const float* V, int ld_v, const float* r_raw,
                                     const float* s_raw, int ld_rs, double* partials, void* stream);
int dgcnn_edgeconv_bwd_a_point_apply(int M, int k, float* r, float* s, int ld_rs, const float* yin, const float* ysum,
                                     const float* cnt, const float* coef_a, void* stream);

/* ---- FP32-accurate tensor-core GEMM (tcgen05, bf16 split planes):  D[M,N] = A[M,K] * B[N,K]^T ---------------
 * A and B are FP32 with arbitrary element strides (row stride, column stride), so transposed operands need no
 * copy.  nsplit = 3: six bf16 MMAs per product, relative error ~2^-24 ("FP32-accurate", the dense GEMMs of the
 * shared MLP / head, model.py:167-193);  nsplit = 2: three MMAs, ~2^-17 (filter precision).
 * workspace: dgcnn_tc_gemm_workspace_bytes(M,N,K,nsplit) bytes of device scratch.                          */
size_t dgcnn_tc_gemm_workspace_bytes(int M, int N, int K, int nsplit);
int dgcnn_tc_gemm_f32(const float* A, int64_t a_row_stride, int64_t a_col_stride,
                      const float* B, int64_t b_row_stride, int64_t b_col_stride,
                      float* D, int ldd, int M, int N, int K, int nsplit,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- plane-level form: split an operand ONCE, multiply it many times --------------------------------------------------
 * The tensor-core operand format ("split planes", csrc/tc.cuh): an FP32 matrix X[R,K] as nsplit bf16 planes x = x1 + x2
 * (+ x3) in units of 64 rows x 64 columns (128 B per row, SWIZZLE_128B).  dgcnn_tc_split_planes_f32 writes them (buffer of
 * dgcnn_tc_planes_bytes(R,K,nsplit) bytes, 128-byte aligned).  dgcnn_tc_gemm_planes_f32 multiplies two such operands:
 *   mode 0   D[M,N] = A[M,K] B[N,K]^T      planes of A over [M,K] (a_cols = K), of B over [N,K] (b_cols = K)
 *   mode 1   D[M,N] = sum_r A[r,m] B[r,n]  planes of A over [K rows, a_cols = M], of B over [K rows, b_cols = N]: the
 *            contraction runs over the unit ROWS (MN-major operand descriptors), so a weight gradient dW = dY^T X reuses
 *            the planes of dY (made for the data gradient) and of X (made for the forward) — no transposing re-split.
 *            The rows are split over the CTAs; partial tiles go through the workspace and are summed in a fixed order.
 *   mode 2   mode 1 with the partial tiles ADDED into D by float reductions instead (D is zeroed first): no partial tiles in
 *            HBM and no reduce pass, summation order over the splits not fixed.
 * a_nsplit / b_nsplit = planes stored per operand (2 or 3), nsplit <= both = planes used.  The layer sequence of a Linear:
 * forward split(X,3) -> mode 0; backward split(dY,2) -> mode 0 with split(W^T,2) for dX, mode 1 with X's planes for dW.  */
size_t dgcnn_tc_planes_bytes(int R, int K, int nsplit);
int dgcnn_tc_split_planes_f32(const float* X, int64_t row_stride, int64_t col_stride, int R, int K, int nsplit,
                              void* planes, size_t planes_bytes, void* stream);
size_t dgcnn_tc_gemm_planes_workspace_bytes(int mode, int M, int N, int K);
int dgcnn_tc_gemm_planes_f32(int mode, const void* a_planes, int a_cols, int a_nsplit, const void* b_planes, int b_cols,
                             int b_nsplit, int nsplit, float* D, int ldd, int M, int N, int K,
                             void* workspace, size_t workspace_bytes, void* stream);

/* ---- conv6 + bn6 + LeakyReLU + global max over the points (model.py:172-175) without materialising [B,emb,N] -------
 * g[segments,N] = max over the `seg_len` rows of each segment of LeakyReLU(BN(A B^T)).  The GEMM epilogue only emits
 * per-row-tile column sums / sums of squares and per-segment column max / min (with the attaining row); BN-affine and
 * LeakyReLU are monotone, so the max of the activations is the activation of the max (or min, for negative scale).
 * use_batch_stats = 1: training (batch statistics over all M rows, running stats updated if given); 0: inference.
 * Outputs for backward: mean, rstd [N]; ysel (selected pre-BN value) and nstar (its row) [segments, N].        */
size_t dgcnn_tc_gemm_colstats_workspace_bytes(int M, int N, int K, int segments);
int dgcnn_tc_gemm_bn_lrelu_segmax_f32(const float* A, int64_t a_row_stride, int64_t a_col_stride,
                                      const float* B, int64_t b_row_stride, int64_t b_col_stride,
                                      int M, int N, int K, int seg_len, const float* gamma, const float* beta,
                                      float eps, float momentum, int use_batch_stats,
                                      float* running_mean, float* running_var, float* mean, float* rstd,
                                      float* g, float* ysel, int32_t* nstar,
                                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- backward of conv6 + bn6 + LeakyReLU + global max, sparse half (model.py:172-175; see dgcnn_tc_gemm_bn_lrelu_segmax_f32) ----
 * gdz[B,E] = gradient at the B*E arg-max entries (row nstar[b,e] of block b), a[E] = gamma*rstd.
 *   dfeat[nstar[b,e], :] += a[e] * gdz[b,e] * w[e, :]   (atomic scatter into the caller-initialised dfeat[M,K])
 *   t1[e, :] = sum_b gdz[b,e] * feat[nstar[b,e], :]      ([E,K], overwritten; fixed order)
 * K % 4 == 0, 16-byte aligned rows.  dgcnn_vecmat_f32: c[K] = scale * v[E] W[E,K] (the b' W row constant).                */
int dgcnn_global_feature_bwd_sparse_f32(const float* feat, int ld_feat, const float* w, int ld_w, const int32_t* nstar,
                                        const float* gdz, const float* a, int B, int E, int K, float* dfeat, int ld_dfeat,
                                        float* t1, void* stream);
int dgcnn_vecmat_f32(const float* v, const float* w, int ld_w, int E, int K, float scale, float* c, void* stream);
/* the per-channel coefficients of that backward in one launch (reference: autograd through model.py:133-135,172-175):
 *   gdz[b,e] = dg[b,e] * LReLU'(gamma_e yhat + beta_e), yhat = (ysel[b,e] - mean_e) rstd_e;
 *   coef[5,E] = a = gamma rstd | kappa = a (s2 / M) rstd | b' = a (s1 / M) - kappa mean | s1 = sum_b gdz (dbeta) |
 *               s2 = sum_b gdz yhat (dgamma);   wneg[E,K] = -kappa_e w[e,:]  (dense, row stride K).
 * dgcnn_global_feature_bwd_dw_f32: dw[e,k] = a_e t1[e,k] - b'_e colsum[k] - kappa_e wf2[e,k] (all dense [E,K] / [K]).   */
int dgcnn_global_feature_bwd_coef_f32(const float* dg, const float* ysel, const float* mean, const float* rstd,
                                      const float* gamma, const float* beta, const float* w, int ld_w, int B, int E, int K,
                                      int M, float slope, float* gdz, float* coef, float* wneg, void* stream);
int dgcnn_global_feature_bwd_dw_f32(const float* coef, const float* t1, const float* colsum, const float* wf2, int E, int K,
                                    float* dw, void* stream);

/* ---- narrow-output linear layer on the FP32 FMA pipe (the classifier conv9: 256 -> num_classes, model.py:141,187-193) ----
 * y[M,NO] = x[M,K] W[NO,K]^T with 1 <= NO <= 8, K %% 4 == 0, K <= 512 (backward: K <= 256, or <= 512 with NO <= 4).  A
 * streaming pass per direction (x read once, dx written once) instead of tensor-core tiles that are 96 % padding.
 * Backward: dx (optional) = dy W, dW[NO,K] = dy^T x (zeroed, then accumulated with float reductions).              */
int dgcnn_linear_narrow_supported(int K, int NO);
int dgcnn_linear_narrow_fwd_f32(const float* x, int ld_x, const float* w, int ld_w, int M, int K, int NO, float* y, int ld_y,
                                void* stream);
int dgcnn_linear_narrow_bwd_f32(const float* x, int ld_x, const float* dy, int ld_dy, const float* w, int ld_w, int M, int K,
                                int NO, float* dx, int ld_dx, float* dw, void* stream);
/* ---- thin-input linear layer (K <= 8: the point-level GEMM of the input block, X[M,4] [Wa1 ; Wa2 - Wa1]^T -> [U|V]) -------
 * y[M,NO] = x[M,K] W[NO,K]^T, NO %% 4 == 0; x rows: ld_x >= round4(K) floats, columns beyond K zero.  Backward (NO = 128):
 * dW[128,K] = g^T x (zeroed, then accumulated with float reductions) and, optionally, dx[M, round4(K)] = g W.          */
int dgcnn_linear_thin_supported(int K, int NO);
int dgcnn_linear_thin_fwd_f32(const float* x, int ld_x, const float* w, int ld_w, int M, int K, int NO, float* y, int ld_y,
                              void* stream);
int dgcnn_linear_thin_bwd_f32(const float* x, int ld_x, const float* g, int ld_g, const float* w, int ld_w, int M, int K,
                              int NO, float* dx, int ld_dx, float* dw, void* stream);

/* ---- shared-MLP head: BatchNorm1d (+ per-block bias) + LeakyReLU(0.2) over point-major rows ---------------------------
 * Replaces `bn7(conv7(..))`, `bn8(conv8(..))` + LeakyReLU of reference model.py:131-139,179-184 on y[M,C] (row-major,
 * contiguous).  v = y + bias[row / N] (bias may be NULL: it carries the once-per-block global-feature half of conv7),
 * h = LReLU(BN(v)).  use_batch_stats != 0: training statistics over the M rows (written to mean/rstd; running_* updated
 * with `momentum` when given, unbiased variance); == 0: running statistics.  aff (float[3C], out) = mean | gamma*rstd |
 * beta, consumed by the backward.  workspace: dgcnn_rows_bn_workspace_bytes(C) bytes.  C in {64,128,256,512,1024}.     */
int dgcnn_rows_bn_supported(int C);
size_t dgcnn_rows_bn_workspace_bytes(int C);
int dgcnn_rows_bn_lrelu_fwd(const float* y, int M, int C, const float* bias, int N, const float* gamma, const float* beta,
                            float eps, float momentum, int use_batch_stats, float* running_mean, float* running_var,
                            float* mean, float* rstd, float* aff, float* h, void* workspace, size_t workspace_bytes,
                            void* stream);
/* backward of the training form: dy[M,C] (gradient w.r.t. y, equally w.r.t. v), dgamma[C], dbeta[C] from dh[M,C].    */
int dgcnn_rows_bn_lrelu_bwd(const float* dh, const float* y, int M, int C, const float* bias, int N, const float* gamma,
                            const float* rstd, const float* aff, float* dy, float* dgamma, float* dbeta,
                            void* workspace, size_t workspace_bytes, void* stream);
/* The same two passes with the result written directly as GEMM operand planes (dgcnn_tc_split_planes_f32's format, nsplit =
 * 2 or 3, rows padded with zeros to a multiple of 128; `planes` 128-byte aligned, dgcnn_tc_planes_bytes(M, C, nsplit) bytes):
 * the FP32 activation / gradient between a BatchNorm pass and the GEMM that consumes it never exists.  Training form only
 * (batch statistics).  C must be a multiple of 64.  dbias (optional, [M/N, C]): the per-block bias gradient, from the same
 * pass (colsum_workspace: dgcnn_rows_colsum_workspace_bytes(M, C, N) bytes).                                                                  */
int dgcnn_rows_bn_lrelu_fwd_planes(const float* y, int M, int C, const float* bias, int N, const float* gamma,
                                   const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                                   float* mean, float* rstd, float* aff, void* planes, size_t planes_bytes, int nsplit,
                                   void* workspace, size_t workspace_bytes, void* stream);
int dgcnn_rows_bn_lrelu_bwd_planes(const float* dh, const float* y, int M, int C, const float* bias, int N,
                                   const float* gamma, const float* rstd, const float* aff, void* planes,
                                   size_t planes_bytes, int nsplit, float* dgamma, float* dbeta, float* dbias,
                                   void* colsum_workspace, size_t colsum_workspace_bytes, void* workspace,
                                   size_t workspace_bytes, void* stream);

/* out[s][c] = sum_{n < N} x[s*N + n][c] for the M/N segments of x[M,C] (the per-block bias gradient `dy.view(B,N,C).sum(1)`
 * of the conv7 split, column sums of the features); deterministic two-pass FP64 reduction.  C % 4 == 0, C <= 1024.    */
size_t dgcnn_rows_colsum_workspace_bytes(int M, int C, int N);
int dgcnn_rows_colsum_f32(const float* x, int M, int C, int N, float* out, void* workspace, size_t workspace_bytes,
                          void* stream);

/* ---- training criterion on device (train.py:180-197 + util.py:17-34; the validation form train.py:265-270) -----------
 * loss = mean over the B*N points of the label-smoothed (eps = `smoothing`, 0.2 upstream) cross entropy that the reference
 * applies to softmax(logits) — i.e. log_softmax of PROBABILITIES, quirk kept.  mask != NULL (uint8 [B*N], 1 = point takes
 * part): the training form with num_classes + 1 channels — a masked point contributes [p_0..p_{C-1}, 0] with its label,
 * an unmasked one [0,..,0,1] with label C (constant, no gradient), replacing the per-point Python loops of train.py:183-192.
 * mask == NULL: the validation form (C channels).  logits / dlogits are addressed with element strides, so both the
 * model's [B,C,N] output and point-major rows work.  Forward and backward are one pass: dlogits (may be NULL) receives
 * grad_scale * d(loss)/d(logits); pred (may be NULL, int32 [B*N]) the arg-max channel the drivers log as prediction.
 * labels: int64 [B*N].  workspace: dgcnn_seg_loss_workspace_bytes(B, N).  Deterministic (fixed-order FP64 reduction). */
size_t dgcnn_seg_loss_workspace_bytes(int B, int N);
int dgcnn_seg_loss_f32(const float* logits, int64_t stride_b, int64_t stride_c, int64_t stride_n, int B, int C, int N,
                       const int64_t* labels, const uint8_t* mask, float smoothing, float grad_scale, float* loss,
                       float* dlogits, int64_t dstride_b, int64_t dstride_c, int64_t dstride_n, int32_t* pred,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- tile-inference epilogue (test.py:123-127, 305-308): softmax over the classes, (max probability, arg-max) per point,
 * pred = arg-max, or `unclassified` (31 upstream) where the max probability is below min_conf.  pred: int64 [B*N] like
 * the reference's tensor; conf (may be NULL): float [B*N] max probability.                                          */
int dgcnn_seg_predict_f32(const float* logits, int64_t stride_b, int64_t stride_c, int64_t stride_n, int B, int C, int N,
                          float min_conf, int64_t unclassified, int64_t* pred, float* conf, void* stream);

/* ---- input pipeline on device (data.py:107-143 sample_points, data.py:169-182 collate_pcs, train.py:176) -------------
 * dgcnn_sample_points: indices of n_samples points of one tile drawn WITHOUT replacement with the reference's class
 * weights (w_c = (tot - count_c) / sum_c (tot - count_c); uniform when that is 0/0), as an exponential race
 * (key = -log(u)/w, the n smallest keys; counter-based u from `seed` and the point index, so the draw is a pure function
 * of labels and seed).  labels: int32 [Nt] class ids (0..num_classes-1; other values are never drawn).  out_idx: int32
 * [n_samples]: the points whose key is below the n-th smallest in index order, then the one(s) at it.  Needs n_samples <= number of points with non-zero weight.
 * dgcnn_gather_points: x[f*x_stride_f + j*x_stride_n] = (float) points[idx[j]*ld_points + f] for f < F — the collate and
 * the permute(0,2,1).float() in one pass, straight into a [F, n] slot of the model's input batch — and y[j] =
 * labels[idx[j]] (int64, may be NULL).  points: float64 (points_are_f64 != 0, what np.load delivers) or float32.      */
size_t dgcnn_sample_points_workspace_bytes(int Nt);
int dgcnn_sample_points(const int32_t* labels, int Nt, int num_classes, int n_samples, uint64_t seed, int32_t* out_idx,
                        void* workspace, size_t workspace_bytes, void* stream);
int dgcnn_gather_points(const void* points, int points_are_f64, int ld_points, const int32_t* labels, const int32_t* idx,
                        int n, int F, float* x, int64_t x_stride_f, int64_t x_stride_n, int64_t* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DGCNN_B200_H */
