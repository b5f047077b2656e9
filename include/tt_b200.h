/*
 * tt_b200.h -- C ABI of the B200-native two-tower DSSM hot path.
 *
 * The reference (juankim834/RecommendSystemProject) has no FFI or operator
 * registry: its hot path is Python calling ATen ops.  Each entry point below
 * replaces one group of those ATen call sites (reference file:line given per
 * function) and is what a
 * ctypes / torch custom-op binding on the reference side would load (see
 * INTEGRATION.md).
 *
 * Conventions
 *  - all pointers are DEVICE pointers unless a name ends in _host;
 *  - every function enqueues work on `stream` (a cudaStream_t passed as
 *    void*) and returns without synchronising; scalars that would force a
 *    host sync (unique-row counts, loss, clip coefficient, NaN flags, Adam
 *    step) live in device memory;
 *  - return value: 0 on success, a positive cudaError_t value for CUDA
 *    failures, a negative TT_E_* code for argument errors; no exception ever
 *    crosses the ABI; tt_last_error() returns a thread-local message;
 *  - caller allocates outputs and workspaces (size queries provided);
 *  - matrices are row-major, ids are int64 at the boundary;
 *  - no global state besides cached function attributes.
 */
#ifndef TT_B200_H
#define TT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TT_ABI_VERSION 2

#if defined(__GNUC__)
#define TT_API __attribute__((visibility("default")))
#else
#define TT_API
#endif

#define TT_E_BADARG (-1)      /* null pointer / non-positive size / bad enum      */
#define TT_E_UNSUPPORTED (-2) /* shape outside what the kernels implement         */
#define TT_E_WORKSPACE (-3)   /* workspace smaller than the size query returned    */
#define TT_E_DEVICE (-4)      /* not an sm_100 device                             */

/* pooling modes (GenericTower.py:155-160, SequenceFeatureProcessor.py:64-68) */
#define TT_POOL_NONE 0 /* L must be 1: plain row gather                           */
#define TT_POOL_SUM 1
#define TT_POOL_MEAN 2 /* divides by L (pads included), as the reference does     */
#define TT_POOL_MAX 3

/* element types of embedding tables / activations */
#define TT_F32 0
#define TT_BF16 1

TT_API int tt_abi_version(void);
TT_API const char *tt_last_error(void);
/* sm count, compute capability; returns TT_E_DEVICE when the current device is not cc 10.x */
TT_API int tt_device_info(int *sm_count_host, int *cc_major_host, int *cc_minor_host);

/* ------------------------------------------------------------------------
 * 1. Embedding gather + pooling forward.
 * Replaces aten::embedding (+ mean/sum/max over dim 1) at
 * GenericTower.py:153-160,182 and SequenceFeatureProcessor.py:60-68.
 *   out[b, :] = pool_{l<L} table[ids[b, l], :]      (never materialises [B,L,D])
 * Padding positions (ids == padding_idx) contribute table[padding_idx] like
 * any other id (the reference pools over all L positions); the kernel counts
 * them and adds the pad row once.  padding_idx < 0: no id is special.
 * out rows are `out_stride` floats apart so a feature can be written straight
 * into its column slice of the tower's concat buffer.
 * argmax (int32 [B, D], TT_POOL_MAX only, may be NULL): position l of the max.
 * ---------------------------------------------------------------------- */
TT_API int tt_emb_gather_pool_fwd(const void *table, int table_dtype, int64_t vocab, int dim,
                           const int64_t *ids, int64_t n_rows, int len, int mode, int64_t padding_idx,
                           float *out, int64_t out_stride, int32_t *argmax, int *oob_flag, void *stream);

/* ------------------------------------------------------------------------
 * 2. Sparse embedding gradient: deterministic sorted-segment scatter-add and
 *    the fused row-wise optimizer.
 * Replaces embedding_dense_backward (autograd of the call sites above,
 * training_utils.py:51) and, for tables, clip_grad_norm_ + Adam.step
 * (training_utils.py:53-56, train_twotower.py:111).
 *
 * tt_emb_segment_grad: positions p = b*len + l, id = ids[p]; source gradient
 * row = grad_out[b] (pooled modes; scaled 1/len for MEAN; routed by argmax
 * for MAX) or grad_out[p] (len == 1).  Positions with id == padding_idx are
 * dropped (nn.Embedding padding_idx semantics).  Outputs, all device:
 *   unique_rows[U] ascending, row_grad[U, dim] fp32, *n_unique = U,
 *   *sq_norm += sum(row_grad^2)   (fixed-order reduction, deterministic).
 * The dense [V, D] gradient is never formed.
 * ---------------------------------------------------------------------- */
TT_API int tt_emb_segment_grad_workspace(int64_t n_pos, int dim, size_t *bytes_host);
TT_API int tt_emb_segment_grad(const int64_t *ids, int64_t n_rows, int len, int mode, int64_t padding_idx,
                        int64_t vocab, const float *grad_out, int64_t grad_stride, const int32_t *argmax,
                        int dim, int64_t *unique_rows, float *row_grad, int32_t *n_unique, float *sq_norm,
                        void *workspace, size_t workspace_bytes, void *stream);

/* List form of tt_emb_segment_grad (owner side of the row-sharded exchange, section 7): position p = piece * piece_len + e
 * holds local row rows[piece * piece_stride + e] (negative or >= vocab: dropped) and reads the gradient row that
 * starts at grad + 4 * pos_src[p] (pos_src: float4 offsets, written by tt_shard_owner_gather) or, when pos_src is NULL,
 * row p of a buffer of pieces: grad[(p / grad_piece_rows) * grad_piece_stride + (p % grad_piece_rows) * dim].
 * Same outputs and the same determinism (stable sort => ascending positions inside a segment). */
TT_API int tt_emb_segment_grad_lists(const int32_t *rows, int64_t n_pieces, int64_t piece_len, int64_t piece_stride,
                              const int32_t *pos_src, int64_t vocab, const float *grad, int64_t grad_piece_rows,
                              int64_t grad_piece_stride, int dim, int64_t *unique_rows, float *row_grad,
                              int32_t *n_unique, float *sq_norm, void *workspace, size_t workspace_bytes, void *stream);

/* Deferred form of the pair (tt_emb_segment_grad_lists, tt_emb_rowwise_adam): "a deterministic sorted-segment scatter-add
 * with a fused row-wise optimizer update".  The global-norm clip (training_utils.py:53-54) needs every gradient before
 * any update, so the work is split differently instead:
 *   phase 1  tt_emb_segment_grad_lists(..., row_grad = NULL, ...): sort, segments, per-row sums -> only their squares
 *            (*sq_norm), unique_rows and *n_unique leave the kernel; `workspace` keeps the sorted lists;
 *   phase 2  tt_emb_segment_adam_lists: forms each row's sum AGAIN (same order, same bits) from the same gradient buffer
 *            and applies Adam to the row in the same kernel.
 * A [U, D] row_grad buffer is never written or read: 7 instead of 9 HBM streams of U x D x 4 bytes per step (the
 * upstream [B, D] gradients are L2-resident).  Between the phases `workspace`, `grad`, `pos_src` and `unique_rows` must
 * stay untouched; n_positions = n_pieces * piece_len of phase 1; dim in {64, 96, 128, 192, 256}, fp32 table. */
TT_API int tt_emb_segment_adam_lists(int64_t n_positions, const int32_t *pos_src, const float *grad, int64_t grad_piece_rows,
                              int64_t grad_piece_stride, int dim, const int64_t *unique_rows, int64_t max_rows,
                              void *workspace, size_t workspace_bytes, float *table, float *exp_avg, float *exp_avg_sq,
                              const float *clip_coef, double lr, double beta1, double beta2, double eps,
                              const int64_t *step_dev, const double *lr_dev, void *stream);

/* Adam on the touched rows only ("lazy" Adam; equals dense Adam the first
 * time a row is touched).  g = row_grad * (*clip_coef) (NULL -> 1).  The step
 * count t is read from *step_dev (so CUDA graphs can replay); lr_dev (nullable,
 * device double) overrides `lr`, so an LR scheduler reaches a captured graph. */
TT_API int tt_emb_rowwise_adam(void *table, int table_dtype, float *exp_avg, float *exp_avg_sq, int dim,
                        const int64_t *unique_rows, const float *row_grad, const int32_t *n_unique,
                        int64_t max_rows, const float *clip_coef, double lr, double beta1, double beta2, double eps,
                        const int64_t *step_dev, const double *lr_dev, void *stream);

/* dense[rows[u], :] += row_grad[u, :]  -- builds the dense .grad the drop-in
 * modules expose to an unmodified torch.optim.Adam. */
TT_API int tt_emb_scatter_rows(float *dense, int dim, const int64_t *unique_rows, const float *row_grad,
                        const int32_t *n_unique, int64_t max_rows, void *stream);

/* ------------------------------------------------------------------------
 * Dense-parameter side of clip_grad_norm_ + Adam (training_utils.py:53-56).
 * ---------------------------------------------------------------------- */
/* *out += sum(x[i]^2), fixed-order two-level reduction */
TT_API int tt_sq_norm_accum(const float *x, int64_t n, float *out, void *workspace, size_t workspace_bytes, void *stream);
/* *coef = min(1, max_norm / (sqrt(sum_k sq_terms[k]) + 1e-6)); also *total_norm if not NULL */
TT_API int tt_clip_coef(const float *sq_terms, int n_terms, float max_norm, float *coef, float *total_norm, void *stream);
/* flat dense Adam over n contiguous floats, g scaled by *clip_coef */
TT_API int tt_adam_flat(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n,
                 const float *clip_coef, double lr, double beta1, double beta2, double eps,
                 const int64_t *step_dev, const double *lr_dev, void *stream);

/* ------------------------------------------------------------------------
 * 3. Fused in-batch (+ hard-negative) softmax cross-entropy.
 * Replaces mm/div/eq/masked_fill/bmm/cat/log_softmax/nll_loss at
 * TwoTowerModel.py:95-140 and their autograd.  The B x (B+H) logit matrix is
 * never written to HBM.
 *   Z = [ mask(U I^T * inv_T) , <U, HN_row> * inv_T , U Pool^T * inv_T ]
 *   loss = mean_b( logsumexp(Z_b) - Z_bb )
 * item_ids (nullable): in-batch logits with item_ids[b]==item_ids[j], b!=j
 * are set to -1e9 after scaling.  hn_rows [B,N,D] is the reference's per-row
 * form; pool [H,D] is the shared-pool form (== hn_rows = pool expanded).
 * nan_flags bits (device int, OR-ed): 1 NaN in user rows, 2 in item rows, 4 in hard negatives; tensor-core entry points also
 * 8 = an item id outside the declared id range (tt_ce_fwd_tc_rect_bits), 16 = logits outside the single-pass form's range
 * (tt_ce_fwd_tc_fused).
 * fp32 variant: exact SIMT path.  bf16 variant: tcgen05/TMA tensor-core path
 * (requires D == 64 or 128 and 16-byte aligned rows).
 * ---------------------------------------------------------------------- */
TT_API int tt_ce_workspace(int64_t batch, int64_t pool, int n_rowneg, int dim, size_t *bytes_host);
TT_API int tt_ce_fwd_f32(const float *user, const float *item, const int64_t *item_ids, const float *hn_rows,
                  int n_rowneg, const float *pool, int64_t pool_rows, int64_t batch, int dim, float inv_temp,
                  float *loss, float *row_lse, float *row_pos, int *nan_flags, void *workspace,
                  size_t workspace_bytes, void *stream);
TT_API int tt_ce_bwd_f32(const float *user, const float *item, const int64_t *item_ids, const float *hn_rows,
                  int n_rowneg, const float *pool, int64_t pool_rows, int64_t batch, int dim, float inv_temp,
                  const float *row_lse, const float *grad_loss, float *d_user, float *d_item, float *d_hn_rows,
                  float *d_pool, void *workspace, size_t workspace_bytes, void *stream);

/* bf16 tensor-core path (tcgen05.mma + TMEM accumulators + TMA operand staging).  Same contract as
 * tt_ce_fwd_f32; inputs are the fp32 activations, converted to bf16 (and permuted into item-id order,
 * which turns the false-negative mask into a contiguous column run per row) inside the call.
 * dim must be 64 or 128.  The workspace must stay alive until the matching backward has run. */
TT_API int tt_ce_tc_workspace(int64_t batch, int64_t pool, int n_rowneg, int dim, size_t *bytes_host);
TT_API int tt_ce_fwd_tc(const float *user, const float *item, const int64_t *item_ids, const float *hn_rows,
                 int n_rowneg, const float *pool, int64_t pool_rows, int64_t batch, int dim, float inv_temp,
                 float *loss, float *row_lse, float *row_pos, int *nan_flags, void *workspace,
                 size_t workspace_bytes, void *stream);

/* Backward of tt_ce_fwd_tc on the tensor cores: two passes of one tcgen05 kernel (dU; dI + dPool) that
 * recompute each 128x128 logit tile into TMEM, turn it into bf16 (P - onehot) in TMEM and feed it straight
 * back to the tensor core as the A operand of the gradient product (the W tile in shared memory is read a
 * second time through an MN-major descriptor).  `fwd_workspace` is the workspace the forward call filled
 * (bf16 operands in item-id order, permutation, collision runs); row_lse is the forward's output.
 * Gradients are fp32, in the caller's (unsorted) row order. */
TT_API int tt_ce_bwd_tc_workspace(int64_t batch, int64_t pool, int n_rowneg, int dim, size_t *bytes_host);
TT_API int tt_ce_bwd_tc(const float *user, const float *hn_rows, int n_rowneg, int64_t pool_rows, int64_t batch, int dim,
                 float inv_temp, const float *row_lse, const float *grad_loss, float *d_user, float *d_item,
                 float *d_hn_rows, float *d_pool, void *fwd_workspace, size_t fwd_workspace_bytes, void *workspace,
                 size_t workspace_bytes, void *stream);

/* Rectangular form of the two calls above for data-parallel towers (SURVEY 8e "towers + loss"): n_user LOCAL user
 * rows against n_item >= n_user item rows -- the all-gathered GLOBAL batch -- where the positive of user b is item row
 * item_offset + b.  item_ids_all [n_item] (nullable) masks every item row that carries the id of the user's positive
 * (false negatives across ALL ranks, TwoTowerModel.py:98-114 applied to the global batch).  loss = mean over the
 * n_user rows; d_item_all [n_item, dim] is this rank's contribution to EVERY item row (sum it over the ranks).
 * n_item == n_user, item_offset == 0 is exactly tt_ce_fwd_tc / tt_ce_bwd_tc. */
TT_API int tt_ce_tc_workspace_rect(int64_t n_user, int64_t n_item, int64_t pool, int n_rowneg, int dim, size_t *bytes_host);
TT_API int tt_ce_fwd_tc_rect(const float *user, const float *item_all, const int64_t *item_ids_all, int64_t item_offset,
                      const float *hn_rows, int n_rowneg, const float *pool, int64_t pool_rows, int64_t n_user,
                      int64_t n_item, int dim, float inv_temp, float *loss, float *row_lse, float *row_pos,
                      int *nan_flags, void *workspace, size_t workspace_bytes, void *stream);
/* tt_ce_fwd_tc_rect with a declared id range: every item id lies in [0, 2^id_bits) (id_bits = bits of vocab_size - 1
 * when the ids index an embedding table, 64 = anything), so the sort that groups equal ids runs ceil(id_bits / 8) radix
 * passes instead of 8.  An id outside the range raises bit 3 (value 8) of *nan_flags. */
TT_API int tt_ce_fwd_tc_rect_bits(const float *user, const float *item_all, const int64_t *item_ids_all, int64_t item_offset,
                           const float *hn_rows, int n_rowneg, const float *pool, int64_t pool_rows, int64_t n_user,
                           int64_t n_item, int dim, float inv_temp, float *loss, float *row_lse, float *row_pos,
                           int *nan_flags, void *workspace, size_t workspace_bytes, int id_bits, void *stream);
TT_API int tt_ce_bwd_tc_workspace_rect(int64_t n_user, int64_t n_item, int64_t pool, int n_rowneg, int dim, size_t *bytes_host);
TT_API int tt_ce_bwd_tc_rect(const float *user, const float *hn_rows, int n_rowneg, int64_t pool_rows, int64_t n_user,
                      int64_t n_item, int dim, float inv_temp, const float *row_lse, const float *grad_loss,
                      float *d_user, float *d_item_all, float *d_hn_rows, float *d_pool, void *fwd_workspace,
                      size_t fwd_workspace_bytes, void *workspace, size_t workspace_bytes, void *stream);

/* Single-pass TRAINING form of the tensor-core CE (TwoTowerModel.py:95-140 forward + the dU half of its autograd in ONE
 * walk over the logit tiles; the logits are evaluated twice per step instead of three times).
 * tt_ce_fwd_tc_fused = tt_ce_fwd_tc_rect_bits without per-row hard negatives, and it ALSO leaves dU's raw partial sums in
 * `bwd_workspace` (tt_ce_bwd_tc_workspace_rect(n_user, n_item, pool, 0, dim) bytes, 256-byte aligned), which must reach
 * tt_ce_bwd_tc_fused untouched together with `workspace`.  tt_ce_bwd_tc_fused then runs only the dI / dPool pass.
 * Precondition: bounded logits, max|u| * max|item or pool row| / T * log2(e) <= 96 (L2-normalised embeddings at any
 * T >= 0.015): the exponentials are taken without a row shift.  The forward checks the bound on the device and raises
 * bit 4 (value 16) of *nan_flags when it is violated -- results are then unusable, call the three-pass entry points. */
TT_API int tt_ce_fwd_tc_fused(const float *user, const float *item_all, const int64_t *item_ids_all, int64_t item_offset,
                       const float *pool, int64_t pool_rows, int64_t n_user, int64_t n_item, int dim, float inv_temp,
                       float *loss, float *row_lse, float *row_pos, int *nan_flags, void *workspace, size_t workspace_bytes,
                       void *bwd_workspace, size_t bwd_workspace_bytes, int id_bits, void *stream);
TT_API int tt_ce_bwd_tc_fused(const float *user, int64_t pool_rows, int64_t n_user, int64_t n_item, int dim, float inv_temp,
                       const float *row_lse, const float *grad_loss, float *d_user, float *d_item_all, float *d_pool,
                       void *fwd_workspace, size_t fwd_workspace_bytes, void *bwd_workspace, size_t bwd_workspace_bytes,
                       void *stream);

/* developer hook (tools/ce_trace.py): SM-clock stamps of CTA 0's pipeline events of the next tcgen05 CE
 * launches are written to dbg[11][256] (device memory); NULL disables.  Not used by the product path. */
TT_API int tt_ce_tc_debug_trace(long long *dbg);

/* ------------------------------------------------------------------------
 * 4. Corpus scoring + top-K for retrieval evaluation.
 * Replaces matmul + per-user -inf masking + topk at training_utils.py:220-258.
 * Scores are never written to HBM.  Selection is two-stage: an fp32 scoring
 * pass keeps the best K+margin candidates per query, which are re-scored in
 * fp64 and ordered by (score descending, corpus row ascending) -- the stated
 * tie-break.  out_idx = local corpus row + row_offset.  mask_offsets/mask_rows
 * (nullable): CSR list of LOCAL corpus rows excluded per query (history mask).
 * ---------------------------------------------------------------------- */
TT_API int tt_score_topk_workspace(int64_t n_query, int64_t n_corpus, int dim, int k, size_t *bytes_host);
TT_API int tt_score_topk_f32(const float *query, int64_t n_query, const float *corpus, int64_t n_corpus, int dim, int k,
                      int64_t row_offset, const int64_t *mask_offsets, const int64_t *mask_rows,
                      double *out_scores, int64_t *out_idx, void *workspace, size_t workspace_bytes,
                      void *stream);
/* Tensor-core scoring path (tcgen05.mma M=128 N=256, bf16 operands, fp32 accumulate in TMEM) with the same
 * contract and the same bit-exact result: the bf16 scores only FILTER candidates (K' = K + margin per list); the
 * survivors are re-scored in fp64 from the fp32 inputs and a per-query proof obligation
 *     exact_score[K-th] > tau + |q - bf16(q)| * max|bf16(e)| + |q| * max|e - bf16(e)| + 2e-5 * |q| * max|e|
 * (tau = best approximate score ever left out; Cauchy-Schwarz on the two rounding-error vectors, the last term
 * covers fp32 accumulation in the tensor core) is evaluated on the device.
 * unverified[q] = 1 marks the (rare) queries for which it does not hold; the caller must re-run those with
 * flags = TT_TOPK_SAMPLING | TT_TOPK_WIDE, then TT_TOPK_WIDE and, if still flagged, through tt_score_topk_f32.
 * flags: TT_TOPK_SAMPLING (corpora of >= 2^17 rows) -- a first pass over every 16th corpus tile gives each query
 *   a starting threshold above which ~2.5 K' items score, which keeps the filter on its fast path; a threshold
 *   that turns out too high (sampling noise, ~1e-4 of the queries) is exactly what the obligation detects.
 *   TT_TOPK_WIDE -- K' = 256 whatever K (the repair pass: tolerates 2.5x more near-ties around the K-th score).
 * dim must be 64 or 128, k <= 224.
 * corpus_bf16 / corpus_bounds (nullable, both or neither): a bf16 copy of the corpus and its two bounds
 * {max row norm of the bf16 copy, max row norm of the rounding error}, prepared once with
 * tt_topk_tc_prepare_corpus (bounds = 2 floats); when NULL the call converts the corpus into its workspace. */
#define TT_TOPK_SAMPLING 1
#define TT_TOPK_WIDE 2
TT_API int tt_topk_tc_prepare_corpus(const float *corpus, int64_t n_corpus, int dim, void *corpus_bf16, float *bounds,
                              void *stream);
TT_API int tt_score_topk_tc_workspace(int64_t n_query, int64_t n_corpus, int dim, int k, int own_corpus, size_t *bytes_host);
TT_API int tt_score_topk_tc(const float *query, int64_t n_query, const float *corpus, const void *corpus_bf16,
                     const float *corpus_bounds, int64_t n_corpus, int dim, int k, int64_t row_offset,
                     const int64_t *mask_offsets, const int64_t *mask_rows, double *out_scores, int64_t *out_idx,
                     int32_t *unverified, int flags, void *workspace, size_t workspace_bytes, void *stream);
/* merge W per-shard lists [W, n_query, k] into the global top-k with the same tie-break */
TT_API int tt_topk_merge(const double *scores, const int64_t *idx, int n_shards, int64_t n_query, int k,
                  double *out_scores, int64_t *out_idx, void *stream);

/* ------------------------------------------------------------------------
 * 5. Small-sequence Transformer behaviour encoder, the non-GEMM parts.
 * Replaces, inside nn.TransformerEncoderLayer as the reference builds it (SequenceEncoder.py:13-21: batch_first,
 * post-norm, ReLU; called at SequenceEncoder.py:60 with src_key_padding_mask): scaled-dot-product attention with
 * key padding mask and attention dropout, and  norm(x + dropout(sublayer(x))).  fp32 throughout.
 * Dropout masks are a counter-based hash of (*seed_dev, call_id, element index); the backward rebuilds them, so the
 * caller passes the same three values to both and changes *seed_dev between steps (it lives in device memory: a
 * CUDA-graph replay sees the new value).  dropout_p = 0 or seed_dev = NULL: no dropout.
 *   tt_attn_small_*: qkv [batch, len, 3*heads*head_dim] packed as the in_proj GEMM leaves it (q | k | v, head h =
 *     columns [h*head_dim, (h+1)*head_dim) of each third); key_pad_mask [batch, len] bytes, 1 = ignore this key
 *     (nullable); out / grad_out [batch, len, heads*head_dim].  len <= 32, head_dim in {8, 16, 32}.
 *   tt_add_dropout_ln_*: y = LayerNorm(x + dropout(z)) over rows of `dim` (= 32k <= 256, biased variance, eps as
 *     given); xhat [rows, dim] and rstd [rows] are saved for the backward, which returns grad_x, grad_z and the full
 *     grad_gamma / grad_beta (per-CTA partials in `workspace`, added in fixed order: deterministic).
 * ---------------------------------------------------------------------- */
TT_API int tt_attn_small_fwd(const float *qkv, const uint8_t *key_pad_mask, int64_t batch, int len, int heads, int head_dim,
                      float dropout_p, const int64_t *seed_dev, int64_t call_id, float *out, void *stream);
TT_API int tt_attn_small_bwd(const float *qkv, const uint8_t *key_pad_mask, const float *grad_out, int64_t batch, int len,
                      int heads, int head_dim, float dropout_p, const int64_t *seed_dev, int64_t call_id,
                      float *grad_qkv, void *stream);
TT_API int tt_add_dropout_ln_fwd(const float *x, const float *z, int64_t rows, int dim, const float *gamma, const float *beta,
                          float eps, float dropout_p, const int64_t *seed_dev, int64_t call_id, float *y, float *xhat,
                          float *rstd, void *stream);
TT_API int tt_add_dropout_ln_bwd_workspace(int64_t rows, int dim, size_t *bytes_host);
TT_API int tt_add_dropout_ln_bwd(const float *grad_y, const float *xhat, const float *rstd, const float *gamma, int64_t rows,
                          int dim, float dropout_p, const int64_t *seed_dev, int64_t call_id, float *grad_x,
                          float *grad_z, float *grad_gamma, float *grad_beta, int accumulate, void *workspace,
                          size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------
 * 6. Linear-layer weight + bias gradient in one pass (backward of every nn.Linear of the towers and the sequence
 * encoder: Tower.py:17-24, SequenceFeatureProcessor.py:30, SequenceEncoder.py:13-21):
 *   grad_weight[n_out, n_in] = grad_out[rows, n_out]^T . input[rows, n_in],  grad_bias[n_out] = column sums of grad_out
 * (grad_bias nullable).  The row dimension is split over the whole chip, partials are added in fixed order.
 * accumulate != 0: the results are ADDED to grad_weight / grad_bias (a parameter's .grad buffer) instead of stored;
 * tt_add_dropout_ln_bwd's grad_gamma / grad_beta take the same flag.
 * ---------------------------------------------------------------------- */
/* out[rows, n_out] = input[rows, n_in] . weight[n_out, n_in]^T + bias (nullable), optional ReLU; and the input gradient
 * grad_input[rows, n_in] = grad_out[rows, n_out] . weight.  One launch each (fp32 FMA, k loop in order: deterministic). */
TT_API int tt_linear_fwd(const float *input, const float *weight, const float *bias, int64_t rows, int n_out, int n_in,
                  int relu, float *out, void *stream);
TT_API int tt_linear_dgrad(const float *grad_out, const float *weight, int64_t rows, int n_out, int n_in, float *grad_input,
                    void *stream);
TT_API int tt_linear_wgrad_workspace(int64_t rows, int n_out, int n_in, size_t *bytes_host);
TT_API int tt_linear_wgrad(const float *grad_out, const float *input, int64_t rows, int n_out, int n_in, float *grad_weight,
                    float *grad_bias, int accumulate, void *workspace, size_t workspace_bytes, void *stream);

/* TF32 tensor-core versions of the three Linear GEMMs above (tcgen05.mma kind::tf32 straight from the fp32 operands,
 * fp32 accumulation in TMEM, TMA staging; the weight is read through an MN-major descriptor for the input gradient and
 * both operands are for the weight gradient, so nothing is transposed or converted in memory).  Same contracts; used
 * when the caller allows TF32 matmuls (torch.backends.cuda.matmul.allow_tf32).  Needs n_out, n_in multiples of 4 and
 * 16-byte aligned pointers (tt_linear_tc_supported).  Weight gradient: split over the rows across the chip, partials
 * added in a fixed order. */
TT_API int tt_linear_tc_supported(int64_t rows, int n_out, int n_in);
TT_API int tt_linear_fwd_tc(const float *input, const float *weight, const float *bias, int64_t rows, int n_out, int n_in,
                     int relu, float *out, void *stream);
TT_API int tt_linear_dgrad_tc(const float *grad_out, const float *weight, int64_t rows, int n_out, int n_in,
                       float *grad_input, void *stream);
TT_API int tt_linear_wgrad_tc_workspace(int64_t rows, int n_out, int n_in, size_t *bytes_host);
TT_API int tt_linear_wgrad_tc(const float *grad_out, const float *input, int64_t rows, int n_out, int n_in,
                       float *grad_weight, float *grad_bias, int accumulate, void *workspace, size_t workspace_bytes,
                       void *stream);

/* Row-wise L2 normalisation of the tower outputs (F.normalize(x, p=2, dim=1), Tower.py:41): y = x / max(||x||, eps),
 * inv_norm[rows] saved for the backward dx = inv_norm (g - y <g, y>) (dx = g / eps for a clamped row).  dim % 4 == 0. */
TT_API int tt_l2_normalize_fwd(const float *x, int64_t rows, int dim, float eps, float *y, float *inv_norm, void *stream);
TT_API int tt_l2_normalize_bwd(const float *grad_y, const float *y, const float *inv_norm, int64_t rows, int dim, float eps,
                        float *grad_x, void *stream);

/* ------------------------------------------------------------------------
 * 8. BatchNorm1d in training mode with the MLP block's ReLU + Dropout fused in (GenericTower.py:234, Tower.py:16-21:
 * Linear -> BatchNorm1d -> ReLU -> Dropout), statistics optionally spanning several ranks (data-parallel towers: the
 * reference normalises over the whole batch).
 *   y = dropout(relu(gamma[c % P] * (x - mean_c) * rstd_c + beta[c % P])),  P = param_period (cols for a plain layer;
 *   the [B, G*C] view of G item slabs that each keep their own batch statistics, TwoTowerModel.py:54-60, uses P = C)
 *   tt_bn_stats      stats[2*cols + 1] = per-channel mean, M2 = sum (x - mean)^2, and the row count (this rank)
 *   tt_bn_apply      merges n_ranks stats blocks (rank order, Chan's formula) -> save_mean / save_rstd [cols]
 *                    (biased variance + eps), batch_var_unbiased [cols] (nullable), running stats (nullable; momentum
 *                    update with the unbiased variance, *num_batches += 1), then normalises
 *   tt_bn_bwd_stats  sums[2*cols] = per-channel sum g, sum g*xhat over this rank's rows (g = dy through dropout and
 *                    ReLU); grad_gamma / grad_beta [P] from these LOCAL sums (stored, or added when accumulate != 0)
 *   tt_bn_bwd_apply  dx = gamma rstd (g - sum_g / N - xhat sum_gx / N); the global sums are the n_ranks blocks
 *                    sums_all[n_ranks][2*cols] added in rank order (deterministic), N = total_rows
 * Dropout masks: counter-based hash of (*seed_dev, call_id, element); the backward rebuilds them and recomputes the
 * ReLU mask from x.  cols, strides and P must be multiples of 4.  Deterministic (fixed reduction order).
 * ---------------------------------------------------------------------- */
TT_API int tt_bn_workspace(int64_t rows, int cols, size_t *bytes_host);
TT_API int tt_bn_stats(const float *x, int64_t rows, int cols, int64_t x_stride, float *stats, void *workspace,
                size_t workspace_bytes, void *stream);
TT_API int tt_bn_apply(const float *x, int64_t rows, int cols, int64_t x_stride, const float *stats_all, int n_ranks,
                const float *gamma, const float *beta, int param_period, float eps, int relu, float dropout_p,
                const int64_t *seed_dev, int64_t call_id, float *y, int64_t y_stride, float *save_mean, float *save_rstd,
                float *batch_var_unbiased, float *running_mean, float *running_var, float momentum, int64_t *num_batches,
                void *stream);
TT_API int tt_bn_bwd_stats(const float *dy, int64_t dy_stride, const float *x, int64_t rows, int cols, int64_t x_stride,
                    const float *save_mean, const float *save_rstd, const float *gamma, const float *beta,
                    int param_period, int relu, float dropout_p, const int64_t *seed_dev, int64_t call_id, float *sums,
                    float *grad_gamma, float *grad_beta, int accumulate, void *workspace, size_t workspace_bytes,
                    void *stream);
TT_API int tt_bn_bwd_apply(const float *dy, int64_t dy_stride, const float *x, int64_t rows, int cols, int64_t x_stride,
                    const float *save_mean, const float *save_rstd, const float *gamma, const float *beta,
                    int param_period, int relu, float dropout_p, const int64_t *seed_dev, int64_t call_id,
                    const float *sums_all, int n_ranks, double total_rows, float *dx, int64_t dx_stride, void *stream);

/* One-shot all-gather of a small fp32 vector over NVLink peer memory (the cross-rank BatchNorm statistics above):
 * every rank stores src[n] into slot [rank] of EVERY peer's symmetric buffer (peer_bufs_host[w] = rank w's buffer as
 * mapped in this process; region at region_off_floats, slots slot_floats apart), publishes a system-scope release flag
 * to each peer (uint32 flags at flag_off_floats, one per source rank) and waits for all W flags of its own buffer.
 * *epoch_dev (device, starts at 0) is bumped by the kernel, so CUDA-graph replays need no host update.  A (region, flag)
 * pair must be used by ONE call site, once per step.  world <= 16. */
TT_API int tt_p2p_allgather_small(const float *src, int n, int rank, int world, const void *const *peer_bufs_host,
                           int64_t region_off_floats, int64_t slot_floats, int64_t flag_off_floats,
                           unsigned int *epoch_dev, void *stream);

/* ------------------------------------------------------------------------
 * 7. Row-sharded embedding tables (owner = row % world, local row = row / world): device side of the exchange.
 * New (the reference is single-process); it is what GenericTower.py:141-183 becomes when a table of BASELINE
 * configs[2] (100M users / 10M items x 128) is spread over the GPUs of one NVSwitch box.
 * Every rank sends every owner one int32 block [block_ints] and receives one fp32 block [block_floats] back; both have
 * fixed sizes (capacities), so no size ever crosses the host and the step can be one CUDA graph.  Per table a block holds
 *   off [n_rows + 1] at off_base : exclusive offsets, the entries of sample b for this owner are [off[b], off[b+1])
 *   rows[cap]        at rows_base: the owner's LOCAL rows in (sample, position) order, unused slots = -1
 * and the float block holds, at vec_base, [n_rows][dim] partial sums (len > 1) or [cap][dim] rows (len == 1).
 *   tt_shard_route        source: buckets ids [n_rows, len] by owner into send[world][block_ints] (count, scan, fill:
 *                         ballots only, order preserved); n_pad[n_rows] (nullable) = pads per sample;
 *                         workspace: 4 * world * ceil(n_rows / 4096) bytes (tile totals of the scan);
 *                         *flags |= 1 (id outside [0, vocab)), |= 2 (an owner's entries exceed cap: dropped)
 *   tt_shard_owner_gather owner: pooled != 0: out[s][vec_base + b*dim] = sum of table rows of (source s, sample b) and
 *                         pos_src[s*cap + e] = (s*block_floats + vec_base + b*dim) / 4: where that entry's gradient row
 *                         sits in the backward's float blocks (nullable; feeds tt_emb_segment_grad_lists);
 *                         pooled == 0: out[s][vec_base + e*dim] = table[rows[s][e]], pos_src likewise with e for b
 *   tt_shard_combine      source: out[b] = sum_w recv_vec[w][b] in rank order + n_pad[b] * pad_row, / len for MEAN
 *                         (len > 1), or recv_vec[owner(b)][slot(b)] (len == 1; pad id -> pad_row)
 *   tt_shard_grad_pack    source, backward: grad_out[b] (x 1/len for MEAN) into every owner's block (len > 1) or into
 *                         the owner's slot (len == 1), the layout tt_shard_owner_gather produced
 * ---------------------------------------------------------------------- */
TT_API int tt_shard_route(const int64_t *ids, int64_t n_rows, int len, int64_t padding_idx, int64_t vocab, int world,
                   int32_t *send, int64_t block_ints, int64_t off_base, int64_t rows_base, int64_t cap, int32_t *n_pad,
                   int *flags, void *workspace, size_t workspace_bytes, void *stream);
TT_API int tt_shard_owner_gather(const void *table, int table_dtype, int64_t local_rows, int dim, int world,
                          const int32_t *recv, int64_t block_ints, int64_t off_base, int64_t rows_base, int64_t cap,
                          int64_t n_rows, int pooled, float *out, int64_t block_floats, int64_t vec_base,
                          int32_t *pos_src, void *stream);
TT_API int tt_shard_combine(const float *recv_vec, int64_t block_floats, int64_t vec_base, int world, const int64_t *ids,
                     int64_t n_rows, int len, int64_t padding_idx, int64_t vocab, int mode, const int32_t *send,
                     int64_t block_ints, int64_t off_base, int64_t cap, const int32_t *n_pad, const float *pad_row,
                     int dim, float *out, int64_t out_stride, void *stream);
TT_API int tt_shard_grad_pack(const float *grad_out, int64_t grad_stride, int64_t n_rows, int len, int mode, int dim,
                       int world, const int64_t *ids, int64_t padding_idx, int64_t vocab, const int32_t *send,
                       int64_t block_ints, int64_t off_base, int64_t cap, float *send_vec, int64_t block_floats,
                       int64_t vec_base, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TT_B200_H */
