/*
 * pcm_b200.h -- C ABI of libpcm_b200.so, the B200-native (sm_100a) replacement for the native
 * layer of HaoyiZhu/PointCloudMatters' point-cloud behaviour-cloning training step.
 *
 * Conventions (all entry points):
 *   - plain C, `extern "C"`, raw DEVICE pointers + sizes; no torch / ATen types;
 *   - every call is asynchronous on the `stream` argument (a cudaStream_t passed as void*;
 *     NULL = legacy default stream, which is what the reference launchers always use);
 *   - the caller owns and allocates every buffer; nothing is allocated inside the library
 *     (a few launchers keep a tiny per-process scratch, documented where they do);
 *   - return value: 0 on success, a cudaError_t value (> 0) if the launch failed, or a
 *     negative PCM_E* code for an argument the kernel cannot honour;
 *   - `offset` / `new_offset` are int32 CUMULATIVE END indices per cloud (reference convention,
 *     libs/pointops/functions/query.py:20-22), -1 marks padding in index outputs.
 *
 * Each prototype cites the reference interface it replaces (paths relative to the reference
 * repository root).  The reference's own launchers are `extern "C" void ...(raw pointers)`
 * declared in libs/pointops/src/<op>/<op>_cuda_kernel.h and exported to Python through the
 * pybind11 module `pointops._C` (libs/pointops/src/pointops_api.cpp:15-32).
 */
#ifndef PCM_B200_H_
#define PCM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCM_OK 0
#define PCM_EINVAL (-1)      /* bad argument (null pointer, non-positive size, ...)            */
#define PCM_EUNSUPPORTED (-2) /* size outside what the kernel supports (e.g. nsample > 128)     */

typedef void *pcm_stream_t; /* cudaStream_t */

/* Library / build identification. */
int pcm_abi_version(void);
const char *pcm_build_info(void);
/* Number of kernel launches issued by this library in the current process (bench.py gpu_launches). */
long long pcm_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * pointops family
 * ------------------------------------------------------------------------------------------ */

/* Replaces farthest_point_sampling_cuda_launcher(int b, int n, const float *xyz,
 * const int *offset, const int *new_offset, float *tmp, int *idx)
 * (libs/pointops/src/sampling/sampling_cuda_kernel.h:13; kernel sampling_cuda_kernel.cu:14-129).
 * `n` is the size of the largest cloud, exactly as the reference caller computes it
 * (libs/pointops/functions/sampling.py:14-16): it selects the reference's thread-block size and
 * therefore its tie-break rule, which this kernel reproduces bit-exactly.
 * `tmp` (n_total floats pre-filled with 1e10, sampling.py:18) is only touched when n > 8192;
 * it may be NULL otherwise (running minima live in registers). */
int pcm_farthest_point_sampling(int b, int n, const float *xyz, const int *offset,
                                const int *new_offset, float *tmp, int *idx, pcm_stream_t stream);

/* Tuning hook (no reference counterpart): force the CTA width (128/256/512/1024, 0 = automatic)
 * of the register-resident FPS kernel; used by bench / profiling sweeps. */
int pcm_tune_fps_threads(int threads);

/* Replaces knn_query_cuda_launcher(int m, int nsample, const float *xyz, const float *new_xyz,
 * const int *offset, const int *new_offset, int *idx, float *dist2)
 * (libs/pointops/src/knn_query/knn_query_cuda_kernel.h:13; kernel knn_query_cuda_kernel.cu:60-104).
 * Adds `b` (number of clouds; the reference scans new_offset linearly instead).  Outputs are
 * bit-identical to the reference including tie order (exact replay of its binary heap).
 * dist2 receives SQUARED distances (the Python wrapper applies sqrt, query.py:23) and may be
 * NULL.  nsample <= 128 (reference hard limit, knn_query_cuda_kernel.cu:82-83). */
int pcm_knn_query(int b, int m, int nsample, const float *xyz, const float *new_xyz,
                  const int *offset, const int *new_offset, int *idx, float *dist2,
                  pcm_stream_t stream);

/* Replaces ball_query_cuda_launcher(int m, int nsample, float min_radius, float max_radius,
 * const float *xyz, const float *new_xyz, const int *offset, const int *new_offset, int *idx,
 * float *dist2) (libs/pointops/src/ball_query/ball_query_cuda_kernel.h:17-21; kernel
 * ball_query_cuda_kernel.cu:58-123).  Reproduces the reference's quirks (double-precision 1e-5
 * test, heap_sort without heapify, index written into dist2 on the strided-subsample branch);
 * stops collecting at 2048 candidates where the reference overflows its stack arrays. */
int pcm_ball_query(int b, int m, int nsample, float min_radius, float max_radius,
                   const float *xyz, const float *new_xyz, const int *offset,
                   const int *new_offset, int *idx, float *dist2, pcm_stream_t stream);

/* Replaces random_ball_query_cuda_launcher(int m, int nsample, float min_radius,
 * float max_radius, const int *order, const float *xyz, const float *new_xyz, const int *offset,
 * const int *new_offset, int *idx, float *dist2)
 * (libs/pointops/src/random_ball_query/random_ball_query_cuda_kernel.h; kernel .cu:58-108). */
int pcm_random_ball_query(int b, int m, int nsample, float min_radius, float max_radius,
                          const int *order, const float *xyz, const float *new_xyz,
                          const int *offset, const int *new_offset, int *idx, float *dist2,
                          pcm_stream_t stream);

/* Replace grouping_{forward,backward}_cuda_launcher
 * (libs/pointops/src/grouping/grouping_cuda_kernel.cu:5-25, launchers :27-41).
 * backward ACCUMULATES into grad_input (caller zero-fills, functions/grouping.py:30). */
int pcm_grouping_forward(int m, int nsample, int c, const float *input, const int *idx,
                         float *output, pcm_stream_t stream);
int pcm_grouping_backward(int m, int nsample, int c, const float *grad_output, const int *idx,
                          float *grad_input, pcm_stream_t stream);

/* Replace interpolation_{forward,backward}_cuda_launcher
 * (libs/pointops/src/interpolation/interpolation_cuda_kernel.cu:5-33).  forward ACCUMULATES into
 * output (caller zero-fills, functions/interpolation.py:39). */
int pcm_interpolation_forward(int n, int c, int k, const float *input, const int *idx,
                              const float *weight, float *output, pcm_stream_t stream);
int pcm_interpolation_backward(int n, int c, int k, const float *grad_output, const int *idx,
                               const float *weight, float *grad_input, pcm_stream_t stream);

/* Replace aggregation_{forward,backward}_cuda_launcher
 * (libs/pointops/src/aggregation/aggregation_cuda_kernel.cu:5-39). */
int pcm_aggregation_forward(int n, int nsample, int c, int w_c, const float *input,
                            const float *position, const float *weight, const int *idx,
                            float *output, pcm_stream_t stream);
int pcm_aggregation_backward(int n, int nsample, int c, int w_c, const float *input,
                             const float *position, const float *weight, const int *idx,
                             const float *grad_output, float *grad_input, float *grad_position,
                             float *grad_weight, pcm_stream_t stream);

/* Replace subtraction_{forward,backward}_cuda_launcher
 * (libs/pointops/src/subtraction/subtraction_cuda_kernel.cu:5-30). */
int pcm_subtraction_forward(int n, int nsample, int c, const float *input1, const float *input2,
                            const int *idx, float *output, pcm_stream_t stream);
int pcm_subtraction_backward(int n, int nsample, int c, const int *idx, const float *grad_output,
                             float *grad_input1, float *grad_input2, pcm_stream_t stream);

/* Replace attention_{relation,fusion}_step_{forward,backward}_cuda_launcher
 * (libs/pointops/src/attention/attention_cuda_kernel.cu:9-86, launchers :93-147). */
int pcm_attention_relation_step_forward(int m, int g, int c, const float *query, const float *key,
                                        const float *weight, const int *index_target,
                                        const int *index_refer, float *output,
                                        pcm_stream_t stream);
int pcm_attention_relation_step_backward(int m, int g, int c, const float *query,
                                         float *grad_query, const float *key, float *grad_key,
                                         const float *weight, float *grad_weight,
                                         const int *index_target, const int *index_refer,
                                         const float *grad_output, pcm_stream_t stream);
int pcm_attention_fusion_step_forward(int m, int g, int c, const float *weight, const float *value,
                                      const int *index_target, const int *index_refer,
                                      float *output, pcm_stream_t stream);
int pcm_attention_fusion_step_backward(int m, int g, int c, const float *weight,
                                       float *grad_weight, const float *value, float *grad_value,
                                       const int *index_target, const int *index_refer,
                                       const float *grad_output, pcm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Dense blocks of the policy.  The reference runs these through
 * torch.nn (cuBLAS SGEMM, ATen kernels): nn.Linear / MultiheadAttention in-proj / out-proj
 * (src/models/components/act/transformer.py:220-233,297-314), PointNet's k=1 SubMConv3d layers
 * (src/models/components/pcd_encoder/pointnet.py:31-55) and the set-abstraction Linear
 * (src/models/components/act/act.py:368-370,457-459).
 * ------------------------------------------------------------------------------------------ */

/* C[m,n] (+)= sum_k A(m,k) * B(n,k) (+ bias[n]) (ReLU) on tcgen05 tensor cores: bf16 operands,
 * fp32 accumulation in tensor memory, TMA-fed 128B-swizzled shared-memory ring.
 *   a_mn / b_mn = 0: operand stored row-major [rows, K], pitch ld elements (K contiguous);
 *               = 1: stored row-major [K, rows], pitch ld (rows contiguous) -- lets dX = dY * W and
 *                    dW = dY^T * X read tensors in place, without transposes.
 *   c_bf16: C is bf16 (else fp32), pitch ldc.  accumulate: atomically ADD into fp32 C
 *   (gradient accumulation; implied when split_k > 1).  bias / relu only with split_k == 1.
 *   split_k == 0 (accumulate only): the launcher picks tile width and K split together.
 * Base pointers must be 16-byte aligned and lda / ldb multiples of 8 (TMA). */
int pcm_gemm_bf16(int M, int N, int K, const void *A, int lda, int a_mn, const void *B, int ldb,
                  int b_mn, void *C, int ldc, int c_bf16, const float *bias, int relu,
                  int accumulate, int split_k, pcm_stream_t stream);

/* Batched / scaled form of pcm_gemm_bf16 used by the attention blocks (the reference goes
 * through nn.MultiheadAttention's math path, transformer.py:246-248,329-340): `batch` stacked
 * problems share one 2-D tensor per operand (a_rows_total x ., b_rows_total x .); batch z starts
 * a_batch_rows / b_batch_rows rows further down (for an MN-major operand those rows are the K
 * dimension, so its K tail must meet zeros in the other operand).  C = alpha * acc.
 * c_mode 0: plain rows (z * c_batch_rows + m); 1: head-split -- token-major rows r = l*hs_B + b,
 * columns h*64 + d are written to a (hs_B, hs_nh, hs_L, 64) tensor; 2: head-merge -- batch
 * z = b*hs_nh + h, row l, column d is written to token-major (l*hs_B + b, h*64 + d), pitch ldc. */
int pcm_gemm_bf16_ex(int M, int N, int K, int batch, const void *A, int lda, int a_mn,
                     long long a_rows_total, long long a_batch_rows, const void *B, int ldb, int b_mn,
                     long long b_rows_total, long long b_batch_rows, void *C, int ldc, int c_bf16,
                     int c_mode, long long c_batch_rows, int hs_B, int hs_nh, int hs_L, float alpha,
                     const float *bias, int relu, int accumulate, int split_k, pcm_stream_t stream);
/* _ex2: (1) a second A operand A2 (same shape, majorness and batching as A; row pitch lda2) from which the output
 * columns n >= a2_from_col are computed (a2_from_col must be a multiple of the output tile width: any multiple of
 * 256) -- nn.MultiheadAttention's fused in-projection reads bf16(x + pos) for Q, K and bf16(x) for V in ONE launch
 * (transformer.py:238-240 q = k = with_pos_embed(src, pos), value = src); A2 = NULL: off.  (2) head-split output
 * (c_mode 1) in parts: column n goes to part n / hs_part_cols, whose (B, nh, L, 64) tensor starts hs_part_stride
 * elements after the previous part's (the Q | K | V buffers); hs_part_cols = 0: a single part. */
int pcm_gemm_bf16_ex2(int M, int N, int K, int batch, const void *A, int lda, int a_mn,
                      long long a_rows_total, long long a_batch_rows, const void *B, int ldb, int b_mn,
                      long long b_rows_total, long long b_batch_rows, void *C, int ldc, int c_bf16,
                      int c_mode, long long c_batch_rows, int hs_B, int hs_nh, int hs_L, float alpha,
                      const float *bias, int relu, int accumulate, int split_k, const void *A2, int lda2,
                      int a2_from_col, int hs_part_cols, long long hs_part_stride, pcm_stream_t stream);

/* Debug aid for tile-shape sweeps (tools/bench_gemm.py): force the N extent of the output tile
 * of subsequent GEMM launches (64 / 128 / 256; 0 = heuristic). */
int pcm_gemm_debug_force_bn(int bn);

/* Row-wise softmax stages of multi-head attention between the batched GEMMs (replaces the
 * softmax / dropout of nn.MultiheadAttention's math path, transformer.py:246-248).  Buffers are
 * [Z = B*nh, Lp, Sp] with zero padding.  fwd: S fp32 raw scores -> Y = softmax(scale*S + mask)
 * (bf16), Zd = dropout(Y) (bf16; pass Zd == Y when p_drop == 0); kpm (B, Sk) bytes, non-zero =
 * masked key, may be NULL.  bwd (in place on dZ): dS = scale * Y * (dY - <dY, Y>).  The dropout
 * seed is *seed_base (device memory, may be NULL = 0; lets a CUDA-graph replay draw fresh masks
 * every step) + seed_offset (distinguishes call sites). */
int pcm_attn_softmax_fwd(int Z, int L, int Lp, int Sk, int Sp, int nh, const float *S,
                         const unsigned char *kpm, float scale, float p_drop,
                         const unsigned long long *seed_base, unsigned long long seed_offset, void *Y,
                         void *Zd, pcm_stream_t stream);
int pcm_attn_softmax_bwd(int Z, int L, int Lp, int Sk, int Sp, const void *Y, void *dZ, float scale,
                         float p_drop, const unsigned long long *seed_base,
                         unsigned long long seed_offset, pcm_stream_t stream);

/* Fused multi-head attention for head_dim 64 (tcgen05 / TMEM / TMA; csrc/flash_attn.cu): replaces
 * the whole score -> softmax -> dropout -> value chain of nn.MultiheadAttention's math path
 * (transformer.py:246-248 self-attention, :329-340 cross-attention; torch's
 * multi_head_attention_forward) without materialising the (B*nh, L, S) tensors.
 *   Q (B*nh, L, 64), K, V (B*nh, S, 64): bf16 "head-split" (pcm_gemm_bf16_ex c_mode 1);
 *   kpm (B, S) bytes, non-zero = masked key, may be NULL; scale = 1/sqrt(64) at the call sites;
 *   dropout on the probabilities with the counter-based mask of (*seed_base, seed_offset);
 *   O: bf16 token-major (row l*B + b, column h*64 + d), pitch ldo;
 *   lse (B*nh, L) fp32: log2-domain log-sum-exp per query row (saved for backward).
 * Backward recomputes the probabilities from lse: dO (B*nh, L, 64) bf16 head-split in;
 * dQ (pitch ldq), dK, dV (pitch ldkv) bf16 token-major out; delta (B*nh*L floats) and dQacc
 * (B*nh*L*64 floats) are caller-owned workspaces (dQacc is cleared by the call). */
int pcm_flash_attn_fwd(int B, int nh, int L, int S, const void *Q, const void *K, const void *V,
                       const unsigned char *kpm, float scale, float p_drop,
                       const unsigned long long *seed_base, unsigned long long seed_offset, void *O,
                       int ldo, float *lse, pcm_stream_t stream);
int pcm_flash_attn_bwd(int B, int nh, int L, int S, const void *Q, const void *K, const void *V,
                       const void *O, int ldo, const void *dO, const float *lse,
                       const unsigned char *kpm, float scale, float p_drop,
                       const unsigned long long *seed_base, unsigned long long seed_offset,
                       float *delta, float *dQacc, void *dQ, int ldq, void *dK, void *dV, int ldkv,
                       pcm_stream_t stream);

/* Debug aid: the first n_ctas CTAs of subsequent pcm_flash_attn_* launches write 64 clock64()
 * stamps each (role phase boundaries) to buf (device, n_ctas * 64 int64); NULL = off. */
int pcm_flash_attn_debug_trace(long long *buf, int n_ctas);

/* ------------------------------------------------------------------------------------------
 * Fused set-abstraction head.  Replaces, for ACTPCD.pcd_sampling (src/models/components/act/
 * act.py:446-460) and PCDObsEncoder.pcd_sampling (.../vision/pcd_obs_encoder.py:179-193), the
 * chain pointops.grouping(with_xyz=True) (libs/pointops/functions/grouping.py:35-59) ->
 * nn.Linear(3+C -> H, bias=False) -> BatchNorm1d(H) -> ReLU -> MaxPool1d(k).  See
 * csrc/sa_fused.cu for the reformulation.  W is the Linear weight (H, 3+C) with row pitch ldw;
 * Pf = feat * W[:, 3:]^T (n, H) fp32 comes from pcm_gemm_bf16; idx (m, k) int32 (-1 = padding).
 * stats / gstats are zero-initialised (5, H) fp64 accumulators; coef = [a, b, mean, invstd] (4, H).
 * ------------------------------------------------------------------------------------------ */
int pcm_sa_gather_stats(int m, int k, int H, const float *Pf, const float *xyz, const float *new_xyz,
                        const int *idx, const float *W, int ldw, float *ymax, float *ymin,
                        unsigned char *jmax, unsigned char *jmin, double *stats, pcm_stream_t stream);
int pcm_sa_bn_finalize(int H, const double *stats, double n_rows, const float *gamma,
                       const float *beta, float eps, float momentum, int training,
                       float *running_mean, float *running_var, float *coef, pcm_stream_t stream);
int pcm_sa_output(int m, int H, const float *ymax, const float *ymin, const unsigned char *jmax,
                  const unsigned char *jmin, const float *coef, float *out, unsigned char *jsel,
                  pcm_stream_t stream);
int pcm_sa_bwd_scatter(int m, int k, int H, const float *dout, const float *out,
                       const unsigned char *jsel, const int *idx, const float *xyz,
                       const float *new_xyz, const float *coef, float *dPf, double *gstats,
                       pcm_stream_t stream);
/* Token-layout variants (the set-abstraction head feeding the transformer): query q = b * per_cloud + mi is row
 * (head_rows + mi) * batch + b of the seq-first (S, B, H) token tensor that transformer.py:75-92 builds with
 * flatten / permute / cat passes.  pcm_sa_output_tokens also emits the bf16 operand copies bf16(out) and
 * bf16(out + pos) of the first encoder layer's projections (either may be NULL); pcm_sa_bwd_scatter_tokens reads
 * dout (+ optional second gradient dout2, summed on load) and out in that layout. */
int pcm_sa_output_tokens(int m, int H, int per_cloud, int batch, int head_rows, const float *ymax,
                         const float *ymin, const unsigned char *jmax, const unsigned char *jmin,
                         const float *coef, const float *pos, float *out, void *out_bf16,
                         void *out_pos_bf16, unsigned char *jsel, pcm_stream_t stream);
int pcm_sa_bwd_scatter_tokens(int m, int k, int H, int per_cloud, int batch, int head_rows,
                              const float *dout, const float *dout2, const float *out,
                              const unsigned char *jsel, const int *idx, const float *xyz,
                              const float *new_xyz, const float *coef, float *dPf, double *gstats,
                              pcm_stream_t stream);
/* Cloud-slice form of pcm_sa_gather_stats: a CTA stages the Pf rows of ONE cloud, restricted to a slice of channels, in
 * shared memory, so Pf is read from global memory once instead of once per incoming edge.  offset / new_offset:
 * cumulative END offsets (int32, b clouds) of the source points and of the queries, as the reference's pointops take them
 * (libs/pointops/functions/query.py:8-23); n_max: host-known upper bound of the cloud sizes.  Same results as the generic
 * entry point (per-channel sums: summation order only).  Returns PCM_EUNSUPPORTED outside the fast path (k != 16, or a
 * cloud slice larger than shared memory). */
int pcm_sa_gather_stats_clouds(int b, int n_max, int m, int k, int H, const float *Pf, const float *xyz,
                               const float *new_xyz, const int *idx, const int *offset, const int *new_offset,
                               const float *W, int ldw, float *ymax, float *ymin, unsigned char *jmax,
                               unsigned char *jmin, double *stats, pcm_stream_t stream);
/* Single-extreme forms of the gather pass: BatchNorm's per-channel scale a_c = gamma_c * invstd_c has the sign of gamma_c
 * (the nn.BatchNorm1d weight, act.py:446-460), known before the batch statistics, so only the extreme that ReLU(BN(.)) +
 * max-pool will select is tracked: max of y where gamma >= 0, min where gamma < 0.  yext / jext replace ymax / jmax in
 * pcm_sa_output[_tokens], which then take ymin = jmin = NULL.  Results are bit-identical to the two-extreme entry points.
 * pcm_sa_gather_sel: nsample 16 or 32 only (PCM_EUNSUPPORTED otherwise). */
int pcm_sa_gather_sel(int m, int k, int H, const float *Pf, const float *xyz, const float *new_xyz, const int *idx,
                      const float *W, int ldw, const float *gamma, float *yext, unsigned char *jext, double *stats,
                      pcm_stream_t stream);
int pcm_sa_gather_sel_clouds(int b, int n_max, int m, int k, int H, const float *Pf, const float *xyz,
                             const float *new_xyz, const int *idx, const int *offset, const int *new_offset,
                             const float *W, int ldw, const float *gamma, float *yext, unsigned char *jext,
                             double *stats, pcm_stream_t stream);
int pcm_sa_edge_stats(int m, int k, const int *idx, const float *xyz, const float *new_xyz,
                      float *cnt, float *sq, double *sdtot, pcm_stream_t stream);
int pcm_sa_bwd_coef(int H, const double *gstats, const double *fstats, const double *sdtot,
                    const float *coef, double n_rows, int training, float *ab, float *dW, int ldw,
                    float *dgamma, float *dbeta, pcm_stream_t stream);
int pcm_sa_bwd_dense(int n, int H, const float *Pf, const float *xyz, const float *cnt,
                     const float *sq, const float *W, int ldw, const float *ab, const float *dPf,
                     void *dPf_bf16, pcm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused residual + dropout + LayerNorm: y = LayerNorm(res + dropout(x)) -- the epilogue of every
 * transformer sub-block of the reference (`src = self.norm1(src + self.dropout1(src2))`,
 * src/models/components/act/transformer.py:249-253,333-345; ATen dropout / add / layer_norm
 * kernels there).  C in {128, 256, 512, 1024}; x may be NULL (plain LayerNorm).  Optional outputs
 * y_bf16, h = res + dropout(x), mean, rstd.  Backward accumulates dgamma / dbeta (caller
 * zero-fills) and writes dres (and dx = dropout-backward(dres) when dx != dres).
 * pcm_colsum: out[c] += sum_r src[r, c] (bias gradients of nn.Linear / in_proj / out_proj).
 * ------------------------------------------------------------------------------------------ */
int pcm_add_dropout_ln_fwd(long long rows, int C, const float *x, const float *res, const float *gamma,
                           const float *beta, float eps, float p_drop,
                           const unsigned long long *seed_base, unsigned long long seed_offset, float *y,
                           void *y_bf16, float *h, float *mean, float *rstd, pcm_stream_t stream);
int pcm_add_dropout_ln_bwd(long long rows, int C, const float *dy, const float *h, const float *mean,
                           const float *rstd, const float *gamma, float p_drop,
                           const unsigned long long *seed_base, unsigned long long seed_offset,
                           float *dres, float *dx, float *dgamma, float *dbeta, pcm_stream_t stream);
/* Extended forms: the forward can additionally emit ypos_bf16 = bf16(y + pos[r / pos_row_div]) --
 * the `with_pos_embed` operand of the NEXT attention block (transformer.py:235-236), so no separate
 * add + cast pass reads y again; the backward can emit dx_bf16 = bf16(dx), the operand of the
 * sub-block's backward GEMMs.  Backward: dy_b (may be NULL) is a second gradient of y, summed with dy on
 * load -- y feeds both the next sub-block and the next residual connection, and handing the two contributions
 * over separately saves autograd's add kernel. */
int pcm_add_dropout_ln_fwd_ex(long long rows, int C, const float *x, const float *res,
                              const float *gamma, const float *beta, float eps, float p_drop,
                              const unsigned long long *seed_base, unsigned long long seed_offset,
                              float *y, void *y_bf16, float *h, float *mean, float *rstd,
                              const float *pos, int pos_row_div, void *ypos_bf16, pcm_stream_t stream);
int pcm_add_dropout_ln_bwd_ex(long long rows, int C, const float *dy, const float *dy_b, const float *h, const float *mean,
                              const float *rstd, const float *gamma, float p_drop,
                              const unsigned long long *seed_base, unsigned long long seed_offset,
                              float *dres, float *dx, float *dgamma, float *dbeta, void *dx_bf16,
                              pcm_stream_t stream);
/* _ex2: additionally ACCUMULATES the column sums of dx into dx_colsum (C floats; NULL = off): the bias gradient of the
 * linear layer that produced x (out_proj / linear2), formed while dx is in registers. */
int pcm_add_dropout_ln_bwd_ex2(long long rows, int C, const float *dy, const float *dy_b, const float *h,
                               const float *mean, const float *rstd, const float *gamma, float p_drop,
                               const unsigned long long *seed_base, unsigned long long seed_offset,
                               float *dres, float *dx, float *dgamma, float *dbeta, void *dx_bf16,
                               float *dx_colsum, pcm_stream_t stream);
int pcm_colsum(long long rows, int C, const void *src, long long ld, int src_bf16, float *out,
               pcm_stream_t stream);
/* FFN hidden layer (transformer.py:243-247,336-340: `linear2(dropout(relu(linear1(x))))`): dropout of the
 * ReLU'd hidden activation h (rows, Hd) bf16 -- the output of pcm_gemm_bf16 with bias+ReLU epilogue -- and its
 * backward dh = bf16(d(dropped) * keep * scale * [h > 0]) (p_drop = 0: ReLU gate only).  Hd % 8 == 0.  Masks
 * come from the counter-based RNG of the LayerNorm kernels (seed_base in device memory, seed_offset per call). */
int pcm_ffn_dropout_fwd(long long rows, int Hd, const void *h, float p_drop, const unsigned long long *seed_base,
                        unsigned long long seed_offset, void *out, pcm_stream_t stream);
int pcm_ffn_relu_dropout_bwd(long long rows, int Hd, const float *dhd, const void *h, float p_drop,
                             const unsigned long long *seed_base, unsigned long long seed_offset, void *dh,
                             pcm_stream_t stream);
/* _ex: additionally ACCUMULATES the column sums of dh into dh_colsum (Hd floats; NULL = off; Hd <= 256, Hd / 8 a
 * power of two): the gradient of linear1's bias. */
int pcm_ffn_relu_dropout_bwd_ex(long long rows, int Hd, const float *dhd, const void *h, float p_drop,
                                const unsigned long long *seed_base, unsigned long long seed_offset, void *dh,
                                float *dh_colsum, pcm_stream_t stream);
/* Profiling aid (tools/bench_ln.py): launch-shape knobs of the LayerNorm / colsum kernels (maximum CTAs of the
 * forward and of the backward, target CTA count and minimum rows per CTA of colsum); <= 0 keeps a value. */
int pcm_ln_debug_tune(int ln_fwd_max_ctas, int ln_bwd_max_ctas, int colsum_ctas, int colsum_min_rows);
/* out = bf16(a + b): `with_pos_embed` (transformer.py:235-236) fused with the operand cast of the
 * Q/K projections.  b may be NULL; b_row_div > 1 broadcasts b's rows over the batch. */
int pcm_add_cast_bf16(long long rows, int C, const float *a, const float *b, int b_row_div, void *out,
                      pcm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused clip-by-global-norm + AdamW over flat fp32 buffers.
 * Replaces torch.nn.utils.clip_grad_norm_ (Lightning gradient_clip_val, configs/trainer/
 * ddp.yaml:12) + torch.optim.AdamW (src/utils/optimizer.py:33-72; configs/model/
 * maniskill2_act_pcd_model.yaml:11-14).  hyper (device, 9 floats) = [lr, beta1, beta2, eps,
 * weight_decay, bias_correction1, bias_correction2, clip_norm, grad_scale]; grad is scaled by
 * grad_scale (1/world after a SUM all-reduce) and the clip coefficient in place.  n % 4 == 0.
 * ------------------------------------------------------------------------------------------ */
int pcm_clip_adamw_step(long long n, float *param, float *grad, float *exp_avg, float *exp_avg_sq,
                        const float *hyper, double *sumsq, float *norm_out, pcm_stream_t stream);
/* Same, and additionally writes bf16(param) to param_bf16 (n elements, 8-byte aligned; may be
 * NULL): the operand copy the tensor-core GEMMs of the next step read, so no fp32 -> bf16 weight
 * conversion kernels run inside the step. */
int pcm_clip_adamw_step_bf16(long long n, float *param, float *grad, float *exp_avg,
                             float *exp_avg_sq, const float *hyper, double *sumsq, float *norm_out,
                             void *param_bf16, pcm_stream_t stream);
/* _ex: zero_grad = 1 leaves the flat gradient ZEROED (ready for the next step's in-place accumulation) instead of holding
 * the clipped gradient: the separate zero-fill pass of the next step disappears. */
int pcm_clip_adamw_step_ex(long long n, float *param, float *grad, float *exp_avg, float *exp_avg_sq,
                           const float *hyper, double *sumsq, float *norm_out, void *param_bf16, int zero_grad,
                           pcm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Diffusion-Policy denoiser: the non-GEMM kernels of
 * ConditionalUnet1D (src/models/components/diffusion_policy/diffusion/conditional_unet1d.py:17-297,
 * conv1d_components.py:8-45).  Activations are channel-last (B, T, C); nn.Conv1d / nn.ConvTranspose1d
 * (cuDNN in the reference) become pcm_gemm_bf16 over rows = B*T against the weight in its own torch
 * layout viewed (Cout, Cin*k) / (Cin, Cout*k), with these two kernels in front of / behind the GEMM:
 *   unfold: col[(b,r), c*k+tap] = x[b, r*stride+tap-pad, c] (0 outside [0,L)); col is bf16 with row
 *           pitch ldc >= C*k (extra columns zeroed); x is fp32 or bf16 with row pitch ldx.
 *   fold:   y[b,p,c] = bias[c] + sum over (r,tap) with r*stride+tap-pad == p of col[(b,r), c*k+tap];
 *           col fp32; y (fp32) and/or y_bf16 are written; bias may be NULL.
 * Conv1d(k,stride,pad): forward = unfold(R = Tout) -> GEMM; dX = GEMM -> fold(L = Tin).
 * ConvTranspose1d(k,stride,pad): forward = GEMM -> fold(R = Tin, L = Tout); backward = unfold.
 * ------------------------------------------------------------------------------------------ */
int pcm_conv1d_unfold(int B, int L, int C, int k, int stride, int pad, int R, const void *x, int x_bf16,
                      long long ldx, void *col, long long ldc, pcm_stream_t stream);
int pcm_conv1d_fold(int B, int L, int C, int k, int stride, int pad, int R, const float *col, long long ldc,
                    const float *bias, float *y, void *y_bf16, pcm_stream_t stream);
/* y = film_scale * Mish(GroupNorm_G(x)) + film_bias (+ res): nn.GroupNorm + nn.Mish of Conv1dBlock
 * (conv1d_components.py:30-41) fused with the FiLM modulation (conditional_unet1d.py:72-77) and the
 * residual add (:80).  x, res, y (B,T,C) fp32 channel-last; film (B, 2C) = [scale | bias] or NULL;
 * mean / rstd (B*G) are saved for the backward.  Backward: dx (fp32), dgamma / dbeta ACCUMULATED
 * with atomics (C), dfilm (B, 2C) written when film != NULL; d(res) = dy is the caller's. */
int pcm_groupnorm_mish_fwd(int B, int T, int C, int G, const float *x, const float *gamma, const float *beta,
                           float eps, const float *film, const float *res, float *y, void *y_bf16,
                           float *mean, float *rstd, pcm_stream_t stream);
int pcm_groupnorm_mish_bwd(int B, int T, int C, int G, const float *x, const float *gamma, const float *beta,
                           const float *mean, const float *rstd, const float *film, const float *dy,
                           float *dx, float *dgamma, float *dbeta, float *dfilm, pcm_stream_t stream);
/* nn.Mish on a flat fp32 vector (cond_encoder / diffusion_step_encoder, conditional_unet1d.py:44-48,103-108) */
int pcm_mish_fwd(long long n, const float *x, float *y, void *y_bf16, pcm_stream_t stream);
int pcm_mish_bwd(long long n, const float *x, const float *dy, float *dx, pcm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * BatchNorm1d + ReLU over token-major rows (R, C) fp32 (the
 * SubMConv3d(k=1) -> BatchNorm1d -> ReLU chain of src/models/components/pcd_encoder/pointnet.py:29-55,
 * and the projector of pcd_obs_encoder.py:100-121).  C % 4 == 0, C <= 1024.
 *   pcm_bn_stats:      stats (2, C) fp64 (caller zero-fills) += [sum_r y, sum_r y^2]
 *   pcm_sa_bn_finalize (above) turns stats into coef (4, C) = [a, b, mean, invstd] and updates the running buffers
 *   pcm_bn_apply_relu: out = max(a*y + b, 0) (relu = 0: no clamp) as fp32 and / or bf16
 *   pcm_bn_relu_bwd:   gstats (2, C) fp64 (caller zero-fills) = [sum dz, sum dz*xhat], dz = dout * [a*y+b > 0];
 *                      dy = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)) (training) or a*dz (eval), as fp32
 *                      and / or bf16; dgamma / dbeta (may be NULL) are ACCUMULATED into.
 * ------------------------------------------------------------------------------------------ */
int pcm_bn_stats(long long R, int C, const float *y, double *stats, pcm_stream_t stream);
int pcm_bn_apply_relu(long long R, int C, const float *y, const float *coef, int relu, float *out, void *out_bf16,
                      pcm_stream_t stream);
int pcm_bn_relu_bwd(long long R, int C, const float *dout, const float *y, const float *coef, int relu,
                    int training, double *gstats, float *dy, void *dy_bf16, float *dgamma, float *dbeta,
                    pcm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Observation tokens and action heads of ACT.
 *   pcm_coord_embed_sine_tokens: ACTPCD.coord_embedding_sine (src/models/components/act/act.py:467-506,
 *     normalize=False; dim_t (npf) = temperature ** (2 * (j // 2) / npf) supplied by the caller) written straight
 *     into the positional tensor pos (S, B, E), S = head_rows + per_cloud: rows < head_rows = add_pos
 *     (additional_pos_embed, transformer.py:82-88) broadcast over the batch, row (head_rows + mi, b) = embedding of
 *     coord[b * per_cloud + mi]: per axis [sin(x / dim_t[0::2]) | cos(x / dim_t[1::2])] (act.py:494-502; npf even);
 *     channels [3 * npf, E) are zero (act.py:505).
 *   pcm_fill_head_rows: rows 0 .. head_rows-1 of the token tensor = [latent (B, E) ; proprio (head_rows-1, B, E)]
 *     (transformer.py:89-92) + optional bf16(tokens) / bf16(tokens + pos) for those rows.
 *   pcm_act_heads_loss_fwd / _bwd: a_hat = action_head(hs) (outputs d >= sig_start through a sigmoid: RLBench
 *     gripper / collision, act.py:770-795), is_pad_hat = is_pad_head(hs), action_loss = mean(w_d * (a_hat - a)^2 *
 *     ~is_pad) with w_d = w_pos for d < n_pos else 1 (act.py:272-291,800-825), kl = KLDivergence(mu, logvar)
 *     (loss/misc.py:11-26), losses = [action_loss + kl_weight * kl, action_loss, kl].  hs row (b, q) lives at
 *     hs + b * ld_b + q * ld_q (the decoder output is (Q, B, E) in memory).  actions == NULL: heads only.
 *     acc (1 double) / ticket (1 uint32) are a zero-initialised workspace the kernel leaves zeroed.
 *     Backward: upstream scalars g_loss / g_action / g_kl and optional g_a_hat (B, Q, A) / g_pad (B, Q); writes d_hs,
 *     dmu, dlogvar and ACCUMULATES dWa, dba (and dWp, dbp when g_pad is given).
 * ------------------------------------------------------------------------------------------ */
int pcm_coord_embed_sine_tokens(int per_cloud, int batch, int head_rows, int E, int npf, const float *coord,
                                const float *dim_t, const float *add_pos, float *pos, pcm_stream_t stream);
int pcm_fill_head_rows(int batch, int head_rows, int E, const float *latent, const float *proprio,
                       const float *pos, float *tokens, void *tokens_bf16, void *tokens_pos_bf16,
                       pcm_stream_t stream);
int pcm_act_heads_loss_fwd(int B, int Q, int E, int A, int L, int sig_start, int n_pos, float w_pos,
                           float kl_weight, const float *hs, long long ld_b, long long ld_q, const float *Wa,
                           const float *ba, const float *Wp, const float *bp, const float *actions,
                           const unsigned char *is_pad, const float *mu, const float *logvar, float *a_hat,
                           float *is_pad_hat, float *losses, double *acc, unsigned int *ticket,
                           pcm_stream_t stream);
int pcm_act_heads_loss_bwd(int B, int Q, int E, int A, int L, int sig_start, int n_pos, float w_pos,
                           float kl_weight, const float *hs, long long ld_b, long long ld_q, const float *Wa,
                           const float *Wp, const float *actions, const unsigned char *is_pad, const float *mu,
                           const float *logvar, const float *a_hat, const float *g_loss, const float *g_action,
                           const float *g_kl, const float *g_a_hat, const float *g_pad, float *d_hs,
                           long long dld_b, long long dld_q, float *dWa, float *dba, float *dWp, float *dbp,
                           float *dmu, float *dlogvar, pcm_stream_t stream);

/* Slice lists: n equally sized slices at unrelated addresses (ptrs: DEVICE array of n addresses, 16-byte aligned) --
 * the same sub-block of every decoder layer's in_proj_weight / in_proj_bias inside the flat parameter and gradient
 * buffers.  gather: dst[s] = *ptrs[s] (stack the slices into one GEMM operand; the reference recomputes the memory
 * K / V projections layer by layer, transformer.py:317-346); add: *ptrs[s] += src[s] (hand the stacked gradient back). */
int pcm_gather_slices(int n, long long bytes_per_slice, const long long *ptrs, void *dst, pcm_stream_t stream);
int pcm_add_slices(int n, long long floats_per_slice, const long long *ptrs, const float *src, pcm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * GPU data path: voxel-grid subsampling of a packed batch of raw clouds, replacing the
 * per-sample CPU transform GridSamplePCD (src/data/components/transformpcd.py:684-793, hash_type "fnv") and the
 * packing of pcd_collate_fn (src/utils/sparse_tensor_utils.py:65-82).
 *   grid = floor(coord / grid_size) (float64 arithmetic like numpy >= 2; f32_div = 1: float32 like numpy 1.x),
 *   grid -= per-cloud minimum, key = FNV64-1A(grid) (:775-793), one survivor per distinct key, output ordered by
 *   ascending key (:693-704).  Survivor = the member with the smallest prio[i] (uint32; NULL: the point's index inside
 *   its cloud, i.e. the first member = the reference's test-mode part 0 under a stable argsort; a random priority =
 *   its train mode), ties by index.
 * pcm_grid_sample_select: workspace grid (n,3) i32, gmin (b,3) i32 preset to INT_MAX, tkey / tbest (2n) u64 preset to
 *   all-ones, scratch_key (2n) u64 / scratch_val (2n) u32; outputs at RAW offsets: idx_raw (n) i64, grid_raw (n,3) i64,
 *   counts (b) i32 = voxels per cloud.
 * pcm_grid_sample_gather: dense packing once the counts are known: coord_out (m,3), grid_out (m,3) i64, feat_out
 *   (m, fc [+3]) = [feat / feat_scale - feat_shift (NormalizeColorPCD) , coord (CollectPCD feat_keys)], index_out (m).
 * ------------------------------------------------------------------------------------------ */
int pcm_grid_sample_select(int b, long long n, const float *coord, const long long *offset, double gsx, double gsy,
                           double gsz, int f32_div, const unsigned int *prio, int *grid, int *gmin,
                           unsigned long long *tkey, unsigned long long *tbest, unsigned long long *scratch_key,
                           unsigned int *scratch_val, long long *idx_raw, long long *grid_raw, int *counts,
                           pcm_stream_t stream);
int pcm_grid_sample_gather(int b, long long m, const long long *raw_offset, const long long *new_offset,
                           const long long *idx_raw, const long long *grid_raw, const float *coord,
                           const float *feat, int fc, float feat_scale, float feat_shift, int append_coord,
                           float *coord_out, long long *grid_out, float *feat_out, long long *index_out,
                           pcm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Sparse convolutions of the SpUNet encoder (reference
 * src/models/components/pcd_encoder/spunet.py:95-122,159-217,368-372 on the un-vendored `spconv` library): rules
 * (integer work) + gathers; the contraction itself is pcm_gemm_bf16 on the (rows, k^3 * Cin) column matrix with the
 * weight (Cout, k, k, k, Cin) read in place.  coords (n, 4) int32 = [batch, x, y, z], 0 <= value < 65536.
 *   pcm_spconv_build_table: open-addressing table (tkey (cap) u64 preset to all-ones, tval (cap) i32 preset to INT_MAX,
 *     cap a power of two >= 2n) of the voxels (shift 0) or of their stride-2 parents coord >> 1 (shift 1, value = the
 *     smallest child row).
 *   pcm_spconv_subm_rules: nbr (n, k^3): row of the active voxel at coord + (ox, oy, oz) - k/2, offset index
 *     o = (ox k + oy) k + oz, or -1 (SubMConv3d: output set = input set).
 *   pcm_spconv_down_rules (SparseConv3d k = 2, stride 2 / SparseInverseConv3d): parent (n), kidx (n) = ((x&1) 2 + (y&1)) 2
 *     + (z&1), child (n, 8) preset to -1, coarse_coords (n, 4), m_out (1) = number of coarse voxels (numbered by their
 *     smallest child row); scratch leader (n), excl (n).
 *   pcm_spconv_gather: col[i, o Cp + c] = x[nbr[i, o], c] as bf16 (0 where nbr = -1 or c >= C), Cp = C rounded up to 8.
 *   pcm_spconv_gather_bwd: dx[j, c] = sum_o dcol[nbr[j, k^3-1-o], o Cp + c] (mode 0, submanifold) or
 *     dcol[parent[j], kidx[j] Cp + c] (mode 1, stride-2).
 *   pcm_spconv_inverse_pick / _place: out[i, co] = Z[parent[i], co 8 + kidx[i]]; dZ[m, co 8 + kk] = dout[child[m, kk], co].
 * ------------------------------------------------------------------------------------------ */
int pcm_spconv_build_table(long long n, const int *coords, int shift, unsigned long long *tkey, int *tval,
                           long long cap, pcm_stream_t stream);
int pcm_spconv_subm_rules(long long n, int k, const int *coords, const unsigned long long *tkey, const int *tval,
                          long long cap, int *nbr, pcm_stream_t stream);
int pcm_spconv_down_rules(long long n, const int *coords, const unsigned long long *tkey, const int *tval,
                          long long cap, int *leader, int *excl, int *parent, int *kidx, int *child,
                          int *coarse_coords, int *m_out, pcm_stream_t stream);
int pcm_spconv_gather(long long rows, int kvol, int C, int Cp, const void *x, long long ldx, int x_bf16,
                      const int *nbr, void *col, pcm_stream_t stream);
int pcm_spconv_gather_bwd(long long rows, int kvol, int C, int Cp, int mode, const float *dcol, const int *nbr,
                          const int *parent, const int *kidx, float *dx, pcm_stream_t stream);
int pcm_spconv_inverse_pick(long long rows, int Cout, const float *Z, const int *parent, const int *kidx,
                            float *out, pcm_stream_t stream);
int pcm_spconv_inverse_place(long long coarse_rows, int Cout, const float *dout, const int *child, void *dZ,
                             pcm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * SyncBatchNorm support (reference DDP preset: configs/trainer/ddp.yaml:9 `sync_batchnorm: true`): the statistics buffers
 * are all-reduced over the ranks BETWEEN the reduce and the finalize / apply kernels, together with the row count, which
 * the kernels then read from device memory (n_rows_dev / n_total_dev; NULL = the host value, single-rank behaviour).
 * Weight / bias gradients stay local sums (torch.nn.SyncBatchNorm semantics; the gradient all-reduce averages them).
 * ------------------------------------------------------------------------------------------ */
int pcm_sa_bn_finalize_ex(int H, const double *stats, double n_rows, const double *n_rows_dev, const float *gamma,
                          const float *beta, float eps, float momentum, int training, float *running_mean,
                          float *running_var, float *coef, pcm_stream_t stream);
int pcm_sa_bwd_coef_ex(int H, const double *gstats, const double *fstats, const double *sdtot, const float *coef,
                       double n_rows, const double *n_rows_dev, int training, float *ab, float *dW, int ldw,
                       float *dgamma, float *dbeta, pcm_stream_t stream);
int pcm_bn_relu_bwd_reduce(long long R, int C, const float *dout, const float *y, const float *coef, int relu,
                           double *gstats, pcm_stream_t stream);
int pcm_bn_relu_bwd_apply(long long R, int C, const float *dout, const float *y, const float *coef, int relu,
                          int training, const double *gstats, const double *n_total_dev, float *dy, void *dy_bf16,
                          float *dgamma, float *dbeta, pcm_stream_t stream);

/* Grouped weight-gradient GEMM: n independent problems C_p (M_p x N_p fp32, pitch ldc_p) += A_p^T B_p, A_p = (K_p x M_p)
 * bf16 (pitch lda_p), B_p = (K_p x N_p) bf16 (pitch ldb_p) -- dW = dY^T X with both operands read in place -- run as ONE
 * persistent tcgen05 launch per tile-width class (<= 40 problems per launch).  All arrays are HOST arrays of length n.
 * The reference forms these products one nn.Linear backward at a time (~110 per step); here the operator layer queues
 * them during backward and flushes the queue at the gradient-bucket boundaries. */
int pcm_gemm_dw_grouped(int n, const void *const *A, const int *lda, const void *const *B, const int *ldb,
                        float *const *C, const int *ldc, const int *M, const int *N, const int *K,
                        pcm_stream_t stream);

/* n bf16 column-sum problems out_p[c] += sum_r src_p[r, c] in one launch (HOST arrays of length n; C_p % 8 == 0,
 * C_p <= 2048, ld_p % 8 == 0): the in-projection bias gradients, queued with the weight-gradient GEMMs. */
int pcm_colsum_grouped(int n, const void *const *src, const long long *rows, const int *C, const long long *ld,
                       float *const *out, pcm_stream_t stream);

/* Fused feed-forward sub-block for dim_feedforward = 32 (transformer.py:243-247,336-340; maniskill2_act_pcd_model.yaml
 * dim_feedforward: 32), one kernel each way; E % 64 == 0, 64 <= E <= 512, Hd must be 32.
 *   pcm_ffn32_fwd: hd (rows, 32) bf16 = dropout(relu(x W1^T + b1)) (saved for the backward; clipped / dropped units are
 *     exactly 0), y (rows, E) fp32 = hd W2^T + b2.  x (rows, E) bf16 (pitch ldx), W1 (32, E) bf16, W2 (E, 32) bf16.
 *   pcm_ffn32_bwd: dh (rows, 32) bf16 = (dy W2) * keep_scale(p_drop) * [hd > 0], dx (rows, E) fp32 = dh W1; the weight /
 *     bias gradients are ordinary dW = dY^T X products (pcm_gemm_dw_grouped / pcm_colsum_grouped). */
int pcm_ffn32_fwd(long long rows, int E, int Hd, const void *x, long long ldx, const void *w1, const float *b1,
                  const void *w2, const float *b2, float p_drop, const unsigned long long *seed_base,
                  unsigned long long seed_offset, void *hd, float *y, pcm_stream_t stream);
int pcm_ffn32_bwd(long long rows, int E, int Hd, const void *dy, long long lddy, const void *hd, const void *w1,
                  const void *w2, float p_drop, void *dh, float *dx, pcm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Raw camera frames -> packed per-sample point clouds (the stage in front of pcm_grid_sample_*): the boolean-mask
 * selections of the reference's dataset classes, for a whole batch at once, survivors in ascending point order like
 * numpy's a[mask].
 *   mode 0, ManiSkill2 (src/data/components/maniskill2/maniskill2_single_task_pcd_act.py:196-224): xyz = (b, P, 4) xyzw,
 *     keep w > 0 and z > 0.005 (include_ground: x > -0.8 instead); crop = NULL or (b, 2) int32 {first row, first column}
 *     of the crop_size x crop_size pixel window kept in every cam_h x cam_w camera image (rand_crop, :200-208).
 *   mode 1, RLBench (src/data/components/rlbench/rlbench_single_task_act.py:266-295): xyz = (b, P, xyz_stride >= 3), keep the
 *     points strictly inside bounds = {xmin, ymin, zmin, xmax, ymax, zmax} (float64 comparison); seg = optional (b, P)
 *     fp32 instance ids appended to the colours as a {0, 1} channel (ids listed in `invalid` -> 0, other ids > 0 -> 1).
 * pcm_frame_filter_count: chunk_count[(sample, chunk)], chunk = 1024 consecutive points, ceil(P / 1024) chunks per sample.
 * pcm_frame_filter_scatter: chunk_base = exclusive prefix sum of chunk_count (int64, sample-major); out_xyz (N, 3),
 *   out_color (N, color_ch + (seg != NULL)) fp32; color: (b, P, color_ch) uint8 or fp32.  All pointers are device
 *   pointers except `bounds` (host). */
int pcm_frame_filter_count(int b, long long P, int mode, const float *xyz, int xyz_stride, int include_ground,
                           const double *bounds, const int *crop, int cam_h, int cam_w, int crop_size,
                           int *chunk_count, pcm_stream_t stream);
int pcm_frame_filter_scatter(int b, long long P, int mode, const float *xyz, int xyz_stride, int include_ground,
                             const double *bounds, const int *crop, int cam_h, int cam_w, int crop_size,
                             const void *color, int color_is_u8, int color_ch, const float *seg, const float *invalid,
                             int n_invalid, const long long *chunk_base, float *out_xyz, float *out_color,
                             pcm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PCM_B200_H_ */
