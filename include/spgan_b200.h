/*
 * spgan_b200.h -- C ABI of libspgan_b200.so: the SP-GAN hot path (kNN graph + EdgeConv
 * generator, PointNet critic, WGAN-GP penalty, optimizer step) as hand-written sm_100a CUDA.
 *
 * The reference (liruihui/SP-GAN) has no FFI for this path: it is Python nn.Modules over
 * torch ops.  Each entry point below therefore cites the reference *operator sequence* it
 * replaces (file:line relative to the reference tree); the Python host layer in
 * sp-gan_b200/ mirrors the reference module API on top of these.  INTEGRATION.md shows the
 * ctypes binding a reference maintainer would add.
 *
 * Conventions ("C-ABI"):
 *   - plain C types only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - the caller allocates every output and workspace; the library never allocates, frees
 *     or retains a pointer, keeps no mutable global state, and is re-entrant;
 *   - work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*) of the
 *     caller's current device; no implicit synchronisation; CUDA-graph capturable;
 *   - return 0 on success, a negative SPGAN_E_* code for argument errors, or a positive
 *     cudaError_t if the launch failed.  Nothing ever calls exit() or prints.
 *   - "rows" matrices are row-major fp32 [R, C] with an explicit leading dimension where
 *     noted; the reference's channel-first tensors are [B, C, N].
 *   - tuning knobs: a few entry points read an environment variable ONCE per process (never afterwards, so
 *     the library stays free of mutable state): SPGAN_TC_PF = 2|3|4 (A-operand prefetch ring of the tcgen05
 *     GEMM), SPGAN_TC_FENCE = producer|consumer (where its proxy fence runs), SPGAN_REDUCE_WAVES = 1..32
 *     (CTAs per SM of the column reductions), SPGAN_KNN_WS = 1 (opt-in warp-specialised kNN kernel, same
 *     results).  Defaults are the measured best; none changes a result except through fp32 summation order.
 */
#ifndef SPGAN_B200_H_
#define SPGAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPGAN_ABI_VERSION 1

#define SPGAN_OK 0
#define SPGAN_E_BADARG (-1)      /* null pointer / non-positive size / inconsistent shape */
#define SPGAN_E_UNSUPPORTED (-2) /* valid request outside the implemented envelope (e.g. k+1 > 32) */
#define SPGAN_E_ALIGN (-3)       /* pointer or leading dimension violates a stated alignment */

typedef void *spgan_stream_t; /* cudaStream_t */

int spgan_abi_version(void);
/* Static string for a code returned by any entry point (host pointer, never freed). */
const char *spgan_error_string(int code);

/* ------------------------------------------------------------------ kNN graph + grouping
 * Replaces Generation/modules.py:695-720 (bmm + sum + add + full sort + slice +
 * index_select loop + repeat + cat).  Arithmetic of the distance is the reference's CPU
 * rounding order (see oracle/knn_recipe.c); ties broken by (dist, index). */

/* xs[b,n] = sum_c x[b,c,n]^2, products rounded first, cascade-16 order (modules.py:697).
 * main_cols < 0 selects the reference default (N/32)*32; columns >= main_cols use the
 * 4-way interleaved order ATen applies to its vector tail. */
int spgan_sqnorm(const float *x_bcn, int B, int C, int N, int main_cols, float *xs, spgan_stream_t stream);

/* The same squared norms (same reduction orders) for point-major rows [B*N, C]. */
int spgan_sqnorm_pm(const float *x_rows, int B, int C, int N, int main_cols, float *xs, spgan_stream_t stream);

/* Neighbour lists of point-major rows [B*N, C] by tensor-core filter + exact refine (csrc/knn_tc.cu): approximate
 * distances (tcgen05, fp16x3 split, A operand in tensor memory) select, for every query, every candidate within a
 * proven error margin of its (k+1)-th smallest approximate distance; the reference's exact fp32 recipe then ranks only
 * those (a handful per query) by (dist, index).  Bit-identical to spgan_knn_group on the same points; queries whose
 * candidate list overflows or comes up short (duplicate-heavy or non-finite clouds) are ranked by an exact scan.
 * xs from spgan_sqnorm_pm.  N % 128 == 0, N <= 4096, 4 <= C <= 256, C % 4 == 0, k <= 15 (a channel count that is not
 * a multiple of 4 may be zero-padded by the caller: zero channels change neither the FMA chain nor the norms).
 * spgan_knn_rows_workspace returns the workspace bytes (256-byte aligned buffer) or 0 when the shape is outside the
 * envelope (the caller then uses spgan_knn_group).  workspace[1] (int32) counts the queries ranked by the exact scan. */
size_t spgan_knn_rows_workspace(int B, int C, int N, int k);
int spgan_knn_rows(const float *x_rows, const float *xs, int B, int C, int N, int k, int32_t *idx, void *workspace,
                   size_t workspace_bytes, spgan_stream_t stream);

/* idx[b,n,r] = rank r+1 of row n of dist (rank 0 dropped), int32 [B,N,k]; 1 <= k <= 31, k < N.
 * If ee != NULL also writes the grouped edge features ee[B,2C,N,k] (first C channels the
 * centre point, last C neighbour - centre; modules.py:717-720) in the same kernel. */
int spgan_knn_group(const float *x_bcn, const float *xs, int B, int C, int N, int k, int32_t *idx,
                    float *ee, spgan_stream_t stream);

/* Grouping only, for a caller-supplied neighbour list (the `idx=` argument of
 * get_edge_features, modules.py:683,694).  Indices must lie in [0, N). */
int spgan_group(const float *x_bcn, const int32_t *idx, int B, int C, int N, int k, float *ee,
                spgan_stream_t stream);

/* int32 <-> int64 index views (the reference API exposes int64 [B, N*k]). */
int spgan_idx32_to_idx64(const int32_t *src, int64_t *dst, int64_t n, spgan_stream_t stream);
int spgan_idx64_to_idx32(const int64_t *src, int32_t *dst, int64_t n, spgan_stream_t stream);

/* ------------------------------------------------------------------ the other kNN / grouping entry points
 *  that share the hot path's distance arithmetic.  Query and candidate clouds are
 * channel-first [B,C,Nq] / [B,C,Nc] with caller-supplied squared norms (spgan_sqnorm for the
 * channel-first reduction order of modules.py:642/697, spgan_sqnorm_rows for point-major rows).
 * dist(i,j) = (-2*dot + a) + b with (a,b) = (|q_i|^2, |c_j|^2), or (|c_j|^2, |q_i|^2) when
 * cand_norm_first != 0 (the rounding order of `knn`, modules.py:641-643).  Ties: (dist, index). */

/* xs[r] = sum_c x[r,c]^2 of point-major rows [R,C], rounded squares added in channel order
 * (Common/pointnet_util.py:38-39; bit-identical to torch CPU for C <= 3, i.e. xyz rows). */
int spgan_sqnorm_rows(const float *x_rows, int64_t R, int C, float *xs, spgan_stream_t stream);

/* idx[b,i,r] = candidate of rank first_rank + r (ascending distance) for query i; int32 [B,Nq,k],
 * first_rank in {0,1}, k + first_rank <= min(Nc, 32).  Replaces the distance + topk/sort of
 * `knn` (modules.py:640-646, Common/ops.py:129-135; first_rank 0, cand_norm_first 1) and
 * `knn_point` (Common/pointnet_util.py, Common/pointconv_util.py:107-118; first_rank 0). */
int spgan_knn_query(const float *xq_bcn, const float *xsq, int Nq, const float *xc_bcn, const float *xsc, int Nc,
                    int B, int C, int k, int first_rank, int cand_norm_first, int32_t *idx, spgan_stream_t stream);

/* dist[B,Nq,Nc] materialised: `pairwise_dist` (modules.py:629-637), `square_distance`
 * (Common/pointnet_util.py:19-40). */
int spgan_pairwise_sqdist(const float *xq_bcn, const float *xsq, int Nq, const float *xc_bcn, const float *xsc,
                          int Nc, int B, int C, int cand_norm_first, float *dist, spgan_stream_t stream);

/* spgan_group with a selectable channel order: diff_first != 0 writes [neighbour - centre, centre]
 * (`get_graph_feature`, modules.py:678), 0 writes [centre, neighbour - centre] (modules.py:720). */
int spgan_group_ex(const float *x_bcn, const int32_t *idx, int B, int C, int N, int k, int diff_first, float *ee,
                   spgan_stream_t stream);

/* out[b,s,:] = points[b, idx[b,s], :] for points [B,N,C], idx [B,S] (int32, or int64 when
 * idx_is_int64 != 0): `index_points` (Common/pointnet_util.py:43-59).  If status != NULL it is set
 * to 1 when an index falls outside [-N,N) (that row of out is left untouched; negative indices
 * wrap like torch fancy indexing).
 * spgan_scatter_add_rows is its adjoint: dpoints[b, idx[b,s], :] += g[b,s,:] (atomic). */
int spgan_gather_rows(const float *points, const void *idx, int idx_is_int64, int B, int N, int64_t S, int C,
                      float *out, int *status, spgan_stream_t stream);
int spgan_scatter_add_rows(const float *g, const void *idx, int idx_is_int64, int B, int N, int64_t S, int C,
                           float *dpoints, spgan_stream_t stream);

/* ------------------------------------------------------------------ approximate EMD
 * The synchronous auction of metrics/emd/emd_cuda.cu:95-282 behind emdModule (metrics/emd/emd_module.py:33-71;
 * called with eps 0.005, 300 iterations at Common/GAN_metrics.py:375-379, 406-407): xyz1, xyz2 [B,n,3] ->
 * dist [B,n] (squared distance of every point of xyz1 to its matched point of xyz2) and assignment [B,n].
 * One CTA per cloud pair runs all iterations out of shared memory: no global scratch (the reference passes nine
 * work arrays), n <= ~5200, n need not be a multiple of 1024, iters >= 1.  Arithmetic and tie rules:
 * oracle/emd_recipe.c (bit-identical results).  spgan_emd_grad: gxyz1 = 2 graddist (xyz1 - xyz2[assignment])
 * (emd_cuda.cu:283-316; the reference computes no gradient for xyz2 either). */
int spgan_emd_auction(const float *xyz1, const float *xyz2, int B, int n, float eps, int iters, float *dist,
                      int32_t *assignment, spgan_stream_t stream);
int spgan_emd_grad(const float *xyz1, const float *xyz2, const float *graddist, const int32_t *assignment, int B,
                   int n, float *gxyz1, spgan_stream_t stream);

/* ------------------------------------------------------------------ layout
 * [B,C,N] (arbitrary element strides sb, sc, sn) <-> point-major rows [B*N, C].
 * Replaces the transpose/contiguous calls of Generator.py:167,170 and the strided read of
 * Discriminator.forward's input (model.py:249). */
int spgan_bcn_to_rows(const float *src, int64_t sb, int64_t sc, int64_t sn, int B, int C, int N,
                      float *rows, spgan_stream_t stream);
int spgan_rows_to_bcn(const float *rows, int B, int C, int N, float *dst, spgan_stream_t stream);
/* out[r, :] = [a[r, :Ca], b[r, :Cb]]  (torch.cat([x, z], -1), Generator.py:166). sa/sb: row strides
 * in elements (0 broadcasts one row per segment of seg_rows rows: the tiled latent of model.py:131). */
int spgan_concat_cols(const float *a, int64_t lda, int Ca, const float *b, int64_t ldb, int64_t b_seg_stride,
                      int seg_rows, int Cb, int64_t R, float *out, spgan_stream_t stream);
int spgan_split_cols_add(const float *g, int64_t R, int Ca, int Cb, float *ga, float *gb, spgan_stream_t stream);

/* ------------------------------------------------------------------ dense contraction
 * C[M,N] = op(A) * op(B) (+ bias[N]) (+ C if accumulate); row-major, fp32 accumulate.
 * op(A) is A[M,K] (transA=0, lda >= K) or A^T with A stored [K,M] (transA=1, lda >= M);
 * likewise B stored [K,N] (transB=0) or [N,K] (transB=1).  Replaces every Conv1d(k=1) /
 * Conv2d(1x1) / Conv2d([1,k]) / Linear of Generator.py:56-71,107-135 and
 * Discriminator.py:55-94 and their autograd (dgrad: NN, wgrad: TN with split-K).
 * engine: 0 = fp32 CUDA-core tiles (exact fp32 products, sequential-k accumulation per tile);
 *         1 = tcgen05 tensor cores, TF32x3: operands split x = hi + lo in tf32, three MMAs per product
 *             (hi*hi + hi*lo + lo*hi), fp32 accumulation in TMEM; ~2^-21 relative per product, i.e.
 *             fp32-faithful.  2 = the same pipeline with a bf16 split (~2^-16 per product, twice the
 *             MMA rate; opt-in).  Engines 1/2 cover transA == 0 with M >= 128, N >= 16, K >= 16 and the
 *             weight-gradient form transA == 1, transB == 0 with K >= 4096 and M, N >= 16 (accumulated
 *             in TMEM per 1024 rows of K and flushed into C with fp32 atomics);
 *             both need a workspace of spgan_gemm_workspace(engine, N, K) bytes (256-byte aligned).
 *             Anything else runs on engine 0, which uses the workspace (if given) for deterministic
 *             split-K partial tiles when M <= 128 and 256 <= K < 2048 (the small-batch MLPs).
 * Weight reuse: the tensor engines first split op(B) into the workspace.  A caller that multiplies by the same weight
 * again (same B, N, K, transB, same engine route) may pass the SAME workspace with bit 1 of transB set (transB | 2):
 * the split is then skipped (also honoured by spgan_gemm_fused).
 * The first int of the workspace is a status word: non-zero after completion means the kernel
 * aborted on an internal pipeline timeout (never expected; checked by the tests). */
size_t spgan_gemm_workspace(int engine, int N, int K);
/* 1 when spgan_gemm (engine 3, transA = 0) runs this product on the chunked-K kernel of csrc/gemm_ts.cu (K > 256,
 * K % 128 == 0, 16-byte aligned A): the layout of the split weight kept in the workspace differs between the kernels,
 * so a caller that re-uses a workspace (transB | 2) must key it on this. */
size_t spgan_gemm_bigk_route(int64_t M, int N, int K, const float *A, int64_t lda);
/* Workspace of the weight-gradient form (transA = 1, transB = 0: C[Mo,No] = A^T B, K = rows) on engine 3: a 256-byte
 * status block plus the split-K partial tiles [k_chunks, Mo, No] that csrc/gemm_wg.cu adds up in a fixed order
 * (deterministic; the engine-1 kernel, taken for shapes / alignments outside gemm_wg.cu's envelope or when the
 * workspace is smaller, flushes with atomics).  The status word is written only by a pipeline timeout, which traps. */
size_t spgan_gemm_wgrad_workspace(int64_t Mo, int No, int64_t K);
/* Weight gradient with the forward's operand prologue folded in:  C[Mo,No] (+)= dY^T * pro(X),
 * pro(x)[k,n] = LeakyReLU_slope(x[k,n] * x_scale[n] + x_shift[n]) (both NULL: identity), dY [K,Mo], X [K,No] row-major
 * fp32, K = points / edges.  The weight gradient of conv(LeakyReLU(BatchNorm(x))) (Generator.py:56-62,
 * Discriminator.py:55-81 under autograd) needs the ACTIVATED input, which spgan_gemm_fused never wrote: it is re-formed
 * inside the converter of the tcgen05 weight-gradient kernel instead of by a spgan_norm_apply pass.  Engine-3 arithmetic,
 * deterministic split-K partials; workspace = spgan_gemm_wgrad_workspace bytes, 256-byte aligned.
 * SPGAN_E_UNSUPPORTED outside the kernel's envelope (Mo, No >= 16, K >= 4096, ld % 4 == 0, 16-byte aligned). */
int spgan_gemm_wgrad_fused(int64_t Mo, int No, int64_t K, const float *dY, int64_t ldy, const float *X, int64_t ldx,
                           const float *x_scale, const float *x_shift, float x_slope, float *C, int64_t ldc,
                           int accumulate, void *workspace, size_t workspace_bytes, spgan_stream_t stream);
int spgan_gemm(int transA, int transB, int64_t M, int N, int K, const float *A, int64_t lda, const float *B,
               int64_t ldb, float *C, int64_t ldc, const float *bias, int accumulate, int engine, void *workspace,
               size_t workspace_bytes, spgan_stream_t stream);

/* Fused point-wise product: a conv -> BatchNorm(train) -> LeakyReLU -> conv chain without the normalised tensor or
 * the statistics pass ever touching HBM (Discriminator.py:55-81 `mlps`/`fc2`, Generator.py:56-62 `conv_w`):
 *     C[M,N] = pro(A)[M,K] * op(B) (+ bias[N]) (+ C if accumulate)
 *     pro(a)[r,k] = LeakyReLU_{a_slope}(a[r,k] * a_scale[k] + a_shift[k])      (identity when a_scale == NULL)
 * and, when col_sum / col_sqsum are given, row-block partial sums of C and C*C per column:
 *     col_sum[j, n] = sum of C[r, n] over the rows r of block j, j < spgan_gemm_fused_stats_rows(M)
 * (one block of rows per (CTA, TMEM lane quarter), accumulated in a fixed order: deterministic; N <= 256;
 * spgan_bn_finalize turns them into the batch mean / variance and the next layer's a_scale / a_shift).  A [M,K] row-major with lda % 4 == 0 and 16-byte aligned;
 * op(B) as in spgan_gemm.  16 <= K <= 256, M >= 128, N >= 16.  sm_100a kernel: TMA-staged fp32 A tiles, operand
 * split x = hi + 2^-11 lo in fp16 (22 significant bits, 3 tcgen05 MMAs per product), A converted once per
 * 128-row tile into tensor memory and reused for all N/64 column tiles (csrc/gemm_ts.cu).
 * spgan_gemm_fused_workspace returns the workspace size in bytes (256-byte aligned buffer), or 0 when the
 * shape / alignment is not supported (the caller then composes spgan_gemm + spgan_colstats + spgan_norm_apply). */
size_t spgan_gemm_fused_workspace(int64_t M, int N, int K, const float *A, int64_t lda);
size_t spgan_gemm_fused_stats_rows(int64_t M);
int spgan_gemm_fused(int transB, int64_t M, int N, int K, const float *A, int64_t lda, const float *B, int64_t ldb,
                     float *C, int64_t ldc, const float *bias, int accumulate, const float *a_scale,
                     const float *a_shift, float a_slope, float *col_sum, float *col_sqsum, void *workspace,
                     size_t workspace_bytes, spgan_stream_t stream);

/* ------------------------------------------------------------------ elementwise
 * n = element count of flat fp32 tensors unless rows/cols are given. */
int spgan_fill(float *x, int64_t n, float v, spgan_stream_t stream);
int spgan_copy(const float *x, float *y, int64_t n, spgan_stream_t stream);
int spgan_axpby(float a, const float *x, float b, const float *y, float *out, int64_t n, spgan_stream_t stream);
int spgan_mul(const float *x, const float *y, float *out, int64_t n, spgan_stream_t stream);
/* LeakyReLU (slope 0 = ReLU): nn.LeakyReLU of Generator.py:59,62,68,111,114,155-156. */
int spgan_lrelu(const float *x, float slope, float *y, int64_t n, spgan_stream_t stream);
/* dx = g * (x > 0 ? 1 : slope) */
int spgan_lrelu_bwd(const float *g, const float *x, float slope, float *dx, int64_t n, spgan_stream_t stream);
int spgan_tanh(const float *x, float *y, int64_t n, spgan_stream_t stream);             /* Generator.py:135 */
int spgan_tanh_bwd(const float *g, const float *y, float *dx, int64_t n, spgan_stream_t stream);
/* out[r,c] = x[r,c] (+|*) v[seg(r), c]; seg(r) = r / seg_rows (seg_rows == R: one row vector). */
int spgan_add_segvec(const float *x, const float *v, int64_t R, int C, int64_t seg_rows, float *out,
                     spgan_stream_t stream);
int spgan_mul_segvec(const float *x, const float *v, int64_t R, int C, int64_t seg_rows, float *out,
                     spgan_stream_t stream);
/* out[i] = 1/sqrt(v[i] + eps): eval-mode BatchNorm scale from running_var. */
int spgan_rsqrt_eps(const float *v, float eps, int64_t n, float *out, spgan_stream_t stream);
/* out[r,:] = x[r,:] / (||x[r,:]||_2 + eps): the --z_norm latent normalisation (Generator.py:163-164). */
int spgan_row_l2_normalize(const float *x, int64_t R, int C, float eps, float *out, spgan_stream_t stream);
/* out[r,c] = v[seg(r), c] */
int spgan_bcast_segvec(const float *v, int64_t R, int C, int64_t seg_rows, float *out, spgan_stream_t stream);

/* ------------------------------------------------------------------ column reductions
 * Per segment of seg_rows consecutive rows (nseg = R / seg_rows) and per column.
 * workspace: >= spgan_colreduce_workspace(R, C, seg_rows, nvals) bytes. */
size_t spgan_colreduce_workspace(int64_t R, int C, int64_t seg_rows, int nvals);
int spgan_colsum(const float *x, int64_t R, int C, int64_t seg_rows, float *out /*[nseg,C]*/, void *workspace,
                 spgan_stream_t stream);
int spgan_coldot(const float *x, const float *y, int64_t R, int C, int64_t seg_rows, float *out, void *workspace,
                 spgan_stream_t stream);
/* Batch / instance statistics: mean[nseg,C], rstd = 1/sqrt(biased var + eps), var (biased; may be
 * NULL).  BatchNorm{1,2}d (one segment) and InstanceNorm1d (segment = cloud) of
 * Generator.py:29,58,61,67,121,124 and Discriminator.py:57-79. */
int spgan_colstats(const float *x, int64_t R, int C, int64_t seg_rows, float eps, float *mean, float *rstd,
                   float *var, void *workspace, spgan_stream_t stream);
/* y = ((x - mean[s]) * rstd[s]) * gamma + beta, then LeakyReLU(slope) if slope != 1.
 * gamma/beta: [C] (may be NULL = 1/0). */
int spgan_norm_apply(const float *x, int64_t R, int C, int64_t seg_rows, const float *mean, const float *rstd,
                     const float *gamma, const float *beta, float slope, float *y, spgan_stream_t stream);
/* Running-stat update of nn.BatchNorm (momentum m, unbiased variance): rm = (1-m) rm + m mean,
 * rv = (1-m) rv + m var * R/(R-1); *count += 1 (int64). */
int spgan_bn_update_running(const float *mean, const float *var, int C, int64_t R, float momentum, float *rm,
                            float *rv, int64_t *count, spgan_stream_t stream);
/* spgan_colstats over ONE segment (train-mode BatchNorm) with the running-statistics update of
 * spgan_bn_update_running folded into the finalize pass (one launch fewer per BatchNorm forward). */
int spgan_colstats_bn(const float *x, int64_t R, int C, float eps, float *mean, float *rstd, float *var,
                      float momentum, float *rm, float *rv, int64_t *count, void *workspace, spgan_stream_t stream);
/* The statistics side of spgan_gemm_fused: col_sum / col_sqsum [rows, C] partial sums of a [R, C] GEMM output ->
 * batch mean, rstd = 1/sqrt(biased var + eps), var; optionally (scale, shift != NULL) the prologue tables of the
 * consuming spgan_gemm_fused, scale = rstd * gamma, shift = beta - mean * scale (gamma / beta NULL = 1 / 0), and
 * optionally (rm, rv != NULL) the running-statistics update of nn.BatchNorm.  Partials are combined in double, in
 * row order (deterministic).  Replaces the separate statistics pass of BatchNorm1d/2d after a 1x1 conv
 * (Discriminator.py:56-64, Generator.py:59-61). */
int spgan_bn_finalize(const float *col_sum, const float *col_sqsum, int rows, int C, int64_t R, float eps,
                      const float *gamma, const float *beta, float *mean, float *rstd, float *var, float *scale,
                      float *shift, float momentum, float *rm, float *rv, int64_t *count, spgan_stream_t stream);
/* scale = rstd * gamma, shift = beta - mean * scale from given statistics (eval-mode BatchNorm feeding a fused GEMM). */
int spgan_bn_tables(const float *mean, const float *rstd, const float *gamma, const float *beta, int C, float *scale,
                    float *shift, spgan_stream_t stream);
/* spgan_norm_bwd_reduce over ONE segment that also accumulates the parameter gradients in place:
 * acc_dbeta[c] += sg[c], acc_dgamma[c] += sgx[c] (either may be NULL). */
int spgan_norm_bwd_reduce_acc(const float *g, const float *x, float slope, int64_t R, int C, const float *mean,
                              const float *rstd, const float *gamma, const float *beta, float *sg, float *sgx,
                              float *acc_dbeta, float *acc_dgamma, void *workspace, spgan_stream_t stream);
/* First-order backward of y = act(norm(x)*gamma+beta): given g = dL/dy (post-activation) it computes
 * sums sg[s,c] = sum g', sgx[s,c] = sum g' * xhat, where g' = g masked by the LeakyReLU(slope) of the
 * forward (slope 1 = no activation; the mask is recomputed from x, mean, rstd, gamma, beta with the
 * forward's arithmetic, nothing else is re-read), then dx.  dgamma = sum_s sgx, dbeta = sum_s sg. */
int spgan_norm_bwd_reduce(const float *g, const float *x, float slope, int64_t R, int C, int64_t seg_rows,
                          const float *mean, const float *rstd, const float *gamma, const float *beta, float *sg,
                          float *sgx, void *workspace, spgan_stream_t stream);
int spgan_norm_bwd_apply(const float *g, const float *x, float slope, int64_t R, int C, int64_t seg_rows,
                         const float *mean, const float *rstd, const float *gamma, const float *beta,
                         const float *sg, const float *sgx, float *dx, spgan_stream_t stream);
/* spgan_norm_bwd_apply (one segment) with dx = (the BatchNorm backward) + addend[R, C]: under the gradient penalty's
 * final backward pass the input of a BatchNorm receives a second gradient term from the double-backward node
 * (gradient_penalty.py:31-37); adding it here replaces autograd's separate accumulation pass.  C % 4 == 0. */
int spgan_norm_bwd_apply_add(const float *g, const float *x, float slope, int64_t R, int C, const float *mean,
                             const float *rstd, const float *gamma, const float *beta, const float *sg,
                             const float *sgx, const float *addend, float *dx, spgan_stream_t stream);
/* Second-order (double) backward of train-mode batch norm, needed by the gradient penalty
 * (gradient_penalty.py:31-33 with create_graph=True).  Inputs: first-backward operands
 * (g = dL/dy, x, gamma, mean, rstd) and u = d(loss)/d(dx).  Outputs: gg = d/dg, gx = d/dx,
 * ggamma[C] = d/dgamma.  Five column sums go through `sums` [5, C] (caller workspace). */
int spgan_bn_dbl_bwd_reduce(const float *g, const float *u, const float *x, int64_t R, int C, const float *mean,
                            float *sums /*[5,C]*/, void *workspace, spgan_stream_t stream);
int spgan_bn_dbl_bwd_apply(const float *g, const float *u, const float *x, int64_t R, int C, const float *mean,
                           const float *rstd, const float *gamma, const float *sums, float *gg, float *gx,
                           float *ggamma, spgan_stream_t stream);
/* The same for BatchNorm FUSED with LeakyReLU(slope): g is the gradient w.r.t. the ACTIVATED output; the mask
 * (slope where fma(xhat, gamma, beta) <= 0) is recomputed from x, applied to g on load and to gg on store
 * (it is piecewise constant in x, so gx needs no extra term).  C % 4 == 0, 16-byte aligned pointers. */
int spgan_bn_act_dbl_bwd_reduce(const float *g, const float *u, const float *x, float slope, int64_t R, int C,
                                const float *mean, const float *rstd, const float *gamma, const float *beta,
                                float *sums /*[5,C]*/, void *workspace, spgan_stream_t stream);
int spgan_bn_act_dbl_bwd_apply(const float *g, const float *u, const float *x, float slope, int64_t R, int C,
                               const float *mean, const float *rstd, const float *gamma, const float *beta,
                               const float *sums, float *gg, float *gx, float *ggamma, spgan_stream_t stream);

/* ------------------------------------------------------------------ pooling over points / neighbours
 * Max over each segment of seg_rows rows (torch.max(x2, 2) Generator.py:183;
 * adaptive_max_pool1d Discriminator.py:104); arg = row offset inside the segment of the first
 * maximum (torch semantics). */
int spgan_segmax(const float *x, int64_t R, int C, int64_t seg_rows, float *out, int32_t *arg,
                 spgan_stream_t stream);
/* dx = 0 except dx[s*seg_rows + arg[s,c], c] = g[s,c] */
int spgan_segmax_scatter(const float *g, const int32_t *arg, int64_t R, int C, int64_t seg_rows, float *dx,
                         spgan_stream_t stream);
/* out[s,c] = x[s*seg_rows + arg[s,c], c] */
int spgan_segmax_gather(const float *x, const int32_t *arg, int64_t R, int C, int64_t seg_rows, float *out,
                        spgan_stream_t stream);
/* Fused train-mode BatchNorm1d + LeakyReLU(slope > 0) + max over the points of every cloud, the tail of the critic's
 * point-wise stack (Generation/Discriminator.py:77-81,104).  x [R, C] is read once: column moments over all R rows
 * (mean, rstd, biased var as spgan_colstats with seg_rows = R) and, per segment of seg_rows rows, the pooled value
 * pooled[s,c] = max_r lrelu(bn(x[r,c])) with arg[s,c] = row (inside the segment) of the first extreme of x[:,c]
 * (maximum for gamma >= 0, minimum for gamma < 0: every fp32 step of bn + lrelu is monotone).  C % 4 == 0.
 * workspace: >= spgan_bn_pool_workspace(R, C, seg_rows) bytes, 16-byte aligned. */
size_t spgan_bn_pool_workspace(int64_t R, int C, int64_t seg_rows);
int spgan_bn_pool_fwd(const float *x, int64_t R, int C, int64_t seg_rows, const float *gamma, const float *beta,
                      float eps, float slope, float *mean, float *rstd, float *var, float *pooled, int32_t *arg,
                      void *workspace, spgan_stream_t stream);
/* Backward of spgan_bn_pool_fwd for a gradient gpooled [nseg, C]: gprime [nseg, C] (scratch: gradient at the
 * pre-activation of the selected rows), sg[c] = sum gprime (= d beta), sgx[c] = sum gprime * xhat (= d gamma),
 * dx [R, C] = gamma rstd (gprime at the selected row - sg / R - xhat sgx / R); dx may be NULL. */
int spgan_bn_pool_bwd(const float *gpooled, const float *x, const int32_t *arg, int64_t R, int C, int64_t seg_rows,
                      const float *mean, const float *rstd, const float *gamma, const float *beta, float slope,
                      float *gprime, float *sg, float *sgx, float *dx, spgan_stream_t stream);
/* softmax over the k neighbours of each (point, channel): x, y are [P, k, C] (F.softmax(w, -1),
 * Generator.py:79) and its backward dx = y * (g - sum_k g*y). */
int spgan_softmax_k(const float *x, int64_t P, int k, int C, float *y, spgan_stream_t stream);
int spgan_softmax_k_bwd(const float *g, const float *y, int64_t P, int k, int C, float *dx, spgan_stream_t stream);
/* The same with both train-mode BatchNorm + LeakyReLU applications folded into the loads (Generator.py:78-82):
 * w = softmax_k(lrelu(bn_w(xw))), prod = lrelu(bn_y(xy)) * w from the PRE-normalisation tensors xw, xy [P, k, C]
 * and per-channel (mean, rstd, gamma, beta) of the two BatchNorm2d layers; bit-identical to
 * spgan_norm_apply x2 + spgan_softmax_mul_k without writing the normalised tensors.  k <= 16.
 * w may be NULL (no backward pass will follow: the softmax weights are not written).
 * Backward: dwa / dya = gradients w.r.t. the two ACTIVATED tensors (feed spgan_norm_bwd_*); either may be NULL. */
int spgan_bn_softmax_mul_k(const float *xw, const float *xy, int64_t P, int k, int C, const float *mean_w,
                           const float *rstd_w, const float *gamma_w, const float *beta_w, const float *mean_y,
                           const float *rstd_y, const float *gamma_y, const float *beta_y, float slope, float *w,
                           float *prod, spgan_stream_t stream);
int spgan_bn_softmax_mul_k_bwd(const float *g, const float *xy, const float *w, int64_t P, int k, int C,
                               const float *mean_y, const float *rstd_y, const float *gamma_y, const float *beta_y,
                               float slope, float *dwa, float *dya, spgan_stream_t stream);
/* Row softmax y[r,:] = softmax(x[r,:]) over the last axis of [R, N] and its backward dx = y * (g - sum_j g_j y_j):
 * the N x N attention map of the optional Attention block (--attn; Generation/modules.py:549-553). */
int spgan_row_softmax(const float *x, int64_t R, int N, float *y, spgan_stream_t stream);
int spgan_row_softmax_bwd(const float *g, const float *y, int64_t R, int N, float *dx, spgan_stream_t stream);
/* Fused attention modulation of EdgeBlock (Generator.py:79,82): w = softmax_k(x), prod = y * w, and its
 * backward dx = w (g y - sum_k g y w), dy = g w (dx / dy may be NULL). */
int spgan_softmax_mul_k(const float *x, const float *y, int64_t P, int k, int C, float *w, float *prod,
                        spgan_stream_t stream);
int spgan_softmax_mul_k_bwd(const float *g, const float *y, const float *w, int64_t P, int k, int C, float *dx,
                            float *dy, spgan_stream_t stream);

/* ------------------------------------------------------------------ edge aggregation
 * out[(p*k + r), :] = (pc ? pc[p,:] : 0) + pn[j,:] - pn[p,:] + bias, j = b(p)*N + idx[p, r]:
 * the per-edge value of a 1x1 conv applied to [centre, neighbour - centre] expressed through
 * per-point projections (Generator.py:78,81 and modules.py:793). */
int spgan_edge_combine(const float *pc, const float *pn, const int32_t *idx, const float *bias, int64_t P, int N,
                       int k, int C, float *out, spgan_stream_t stream);
/* Backward: dpc[p] = sum_r g[p,r] (if dpc), dpn[j] += g[p,r], dpn[p] -= sum_r g[p,r] (atomic; dpn is
 * zeroed by the call). */
int spgan_edge_combine_bwd(const float *g, const int32_t *idx, int64_t P, int N, int k, int C, float *dpc,
                           float *dpn, spgan_stream_t stream);
/* EdgeBlock passes with the train-mode BatchNorm reductions folded in (Generator.py:75-88; csrc/edge_fused.cu).
 * spgan_edge_combine_stats = spgan_edge_combine that ALSO leaves per-column sums / sums of squares of what it writes
 * as `spgan_edge_stats_rows(P, C)` deterministic partial rows [rows, C] (feed spgan_bn_finalize with R = P*k): the
 * statistics pass of the BatchNorm2d that follows (Generator.py:58,68) disappears.  out may be NULL (statistics only).
 * Needs C % 4 == 0, 256 % (C/4) == 0 and 16-byte aligned pointers (else SPGAN_E_UNSUPPORTED; rows() returns 0). */
size_t spgan_edge_stats_rows(int64_t P, int C);
int spgan_edge_combine_stats(const float *pc, const float *pn, const int32_t *idx, const float *bias, int64_t P, int N,
                             int k, int C, float *out, float *col_sum, float *col_sqsum, spgan_stream_t stream);
/* spgan_bn_softmax_mul_k_bwd that also emits, as `spgan_attn_bwd_rows(P, k, C)` partial rows part[rows, 4, C], the
 * column sums the two BatchNorm backwards need: (sum g'_w, sum g'_w xhat_w, sum g'_y, sum g'_y xhat_y), g' = the
 * gradient after the LeakyReLU mask recomputed from xw / xy.  Replaces two spgan_norm_bwd_reduce passes.
 * 256 % C == 0, k <= 16. */
size_t spgan_attn_bwd_rows(int64_t P, int k, int C);
int spgan_bn_softmax_mul_k_bwd_stats(const float *g, const float *xw, const float *xy, const float *w, int64_t P, int k,
                                     int C, const float *mean_w, const float *rstd_w, const float *gamma_w,
                                     const float *beta_w, const float *mean_y, const float *rstd_y,
                                     const float *gamma_y, const float *beta_y, float slope, float *dwa, float *dya,
                                     float *part, spgan_stream_t stream);
/* part[rows, nvals, C] -> out[nvals, C]: fixed-tree fp64 sum over the partial rows (deterministic); acc0..acc3
 * (nullable, [C]) receive += out[v] (parameter gradients accumulated in place).  nvals <= 4. */
int spgan_partials_finalize(const float *part, int rows, int nvals, int C, float *out, float *acc0, float *acc1,
                            float *acc2, float *acc3, spgan_stream_t stream);
/* spgan_edge_combine_bwd applied to dx = gamma rstd (g' - sg/n - xhat sgx/n), n = P*k, formed on the fly from the
 * gradient g w.r.t. the ACTIVATED tensor and the pre-normalisation tensor x (= spgan_norm_bwd_apply followed by
 * spgan_edge_combine_bwd without the [P*k, C] intermediate).  C % 4 == 0. */
int spgan_edge_combine_bwd_bn(const float *g, const float *x, const int32_t *idx, int64_t P, int N, int k, int C,
                              const float *mean, const float *rstd, const float *gamma, const float *beta,
                              const float *sg, const float *sgx, float slope, float *dpc, float *dpn,
                              spgan_stream_t stream);
/* out[p,c] = max_r x[p,r,c] with argmax (torch.max(x, 3), modules.py:794) and its scatter. */
int spgan_kmax(const float *x, int64_t P, int k, int C, float *out, int32_t *arg, spgan_stream_t stream);
int spgan_kmax_scatter(const float *g, const int32_t *arg, int64_t P, int k, int C, float *dx, spgan_stream_t stream);
/* w[o, r, c] <-> w4[o, c, 0, r]: Conv2d(F, F, [1,k]) weight (Generator.py:71) as a [Fout, k*F] matrix. */
int spgan_permute_ock_to_okc(const float *src, int O, int Cc, int k, float *dst, spgan_stream_t stream);
int spgan_permute_okc_to_ock(const float *src, int O, int Cc, int k, float *dst, spgan_stream_t stream);
/* AdaIN apply: out[r,c] = s[r,c] * xhat[r,c] + s[r,C+c], xhat = (x - mean[b,c]) * rstd[b,c]
 * (Generator.py:38-45) and its backward pieces ds = [g*xhat, g], gxh = g * s[:, :C]. */
int spgan_adain_apply(const float *x, const float *s, int64_t R, int C, int64_t seg_rows, const float *mean,
                      const float *rstd, float *out, spgan_stream_t stream);
int spgan_adain_bwd(const float *g, const float *x, const float *s, int64_t R, int C, int64_t seg_rows,
                    const float *mean, const float *rstd, float *ds, float *gxh, spgan_stream_t stream);

/* ------------------------------------------------------------------ gradient penalty
 * mix = real + alpha[b] * (fake - real): gradient_penalty.py:26.  real/fake are [B,3,N]-shaped
 * with element strides; mix is contiguous [B,C,N]. */
int spgan_gp_interp(const float *real, int64_t rsb, int64_t rsc, int64_t rsn, const float *fake, int64_t fsb,
                    int64_t fsc, int64_t fsn, const float *alpha, int B, int C, int N, float *mix,
                    spgan_stream_t stream);
/* norms[b] = ||g[b,:]||_2; *penalty = lambda * mean_b(((norm - gamma)/gamma)^2): gradient_penalty.py:35. */
int spgan_gp_penalty(const float *g, int B, int64_t D, float gamma, float lambda, float *norms, float *penalty,
                     spgan_stream_t stream);
/* dg[b,:] = gout * lambda * 2 (norm-gamma) / (gamma^2 B norm) * g[b,:] */
int spgan_gp_penalty_bwd(const float *g, const float *norms, const float *gout, int B, int64_t D, float gamma,
                         float lambda, float *dg, spgan_stream_t stream);
/* out[0] = scale * mean(x[0:n]) (+ out[0] if accumulate): the wgan loss means, loss_utils.py:728-730,859-863 */
int spgan_mean(const float *x, int64_t n, float scale, int accumulate, float *out, spgan_stream_t stream);

/* ------------------------------------------------------------------ evaluation: pairwise Chamfer distance
 * cd[i*R + j - pair0] = mean_n min_m |a_i[n] - b_j[m]|^2 + mean_m min_n |a_i[n] - b_j[m]|^2 for the cloud pairs
 * pair0 <= i*R + j < pair0 + npairs of a [S, N, 3] x b [R, M, 3] (contiguous xyz): the CD half of
 * metrics/evaluation_metrics.py:89-125 (_pairwise_EMD_CD_) / Common/GAN_metrics.py:658-684 (pairwise_CD), with the
 * direct-difference arithmetic of metrics/CD_EMD/cd/chamferdist/chamfer.cu:12-134.  dl / dr (the two directed
 * means, distChamfer evaluation_metrics.py:37-49) are optional; any of cd, dl, dr may be NULL but not all.
 * The pair range lets ranks shard the S x R matrix (BASELINE configs[4]).  N, M <= ~7000 (shared-memory staging). */
int spgan_pairwise_chamfer(const float *a, const float *b, int S, int R, int N, int M, int64_t pair0, int64_t npairs,
                           float *cd, float *dl, float *dr, spgan_stream_t stream);

/* ------------------------------------------------------------------ optimizer
 * torch.optim.Adam semantics (model.py:94-97) over one flat parameter buffer: in-place update of
 * p, m, v from grad_scale * g; step is the 1-based step count.  grad_scale = 1/world_size turns the
 * summed all-reduce result into the data-parallel mean without a separate pass. */
int spgan_adam_step(float *p, const float *g, float *m, float *v, int64_t n, float lr, float beta1, float beta2,
                    float eps, int step, float grad_scale, spgan_stream_t stream);

/* The same update with the step count kept on the device: state = {int32 step, float, float, pad} (16 bytes, zeroed
 * by the caller once); every call increments state[0] first and derives the bias corrections from it.  No argument
 * changes between steps, so a captured CUDA graph of the whole training step can be replayed. */
int spgan_adam_step_dev(float *p, const float *g, float *m, float *v, int64_t n, float lr, float beta1, float beta2,
                        float eps, int32_t *state, float grad_scale, spgan_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SPGAN_B200_H_ */
