/* besskge_b200 — C-ABI of the B200-native BESS-KGE hot path.
 *
 * Every entry point is `extern "C"`, takes plain device pointers, sizes and a
 * CUDA stream (as void*), is stream-ordered, never allocates, never throws and
 * returns 0 on success or a negative bess_status (message: bess_last_error()).
 * All pointers are BORROWED device pointers that must stay alive until the
 * enqueued work has finished.  No torch types appear in any signature.
 *
 * The reference (graphcore-research/bess-kge) has no arithmetic FFI: its only
 * native symbol is a PopART pattern registration loaded with
 * ctypes.cdll.LoadLibrary (besskge/__init__.py:30,
 * custom_ops/remove_all_reduce_pattern.cpp:45-47).  The functions below are
 * what a CUDA back-end of that library binds instead; each one cites the
 * reference torch expression (file:line under besskge/) it replaces.
 */
#ifndef BESSKGE_B200_H
#define BESSKGE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  BESS_OK = 0,
  BESS_ERR_INVALID_ARG = -1,
  BESS_ERR_CUDA = -2,
  BESS_ERR_UNSUPPORTED = -3
} bess_status;

/* BESS_F16X3 is an OPERAND format of the tensor-core path, not a table dtype: fp32 values
 * carried as scaled fp16 (hi, lo) pairs, see bess_dot_gemm. */
typedef enum { BESS_F32 = 0, BESS_F16 = 1, BESS_BF16 = 2, BESS_F16X3 = 3 } bess_dtype;

/* score-function families (scoring.py) */
typedef enum {
  BESS_TRANSE = 0, BESS_ROTATE = 1, BESS_DISTMULT = 2,
  BESS_COMPLEX = 3, BESS_PAIRRE = 4, BESS_BOXE = 5,
  BESS_TRIPLERE = 6, /* scoring.py:596-743: PairRE kernels + a relation offset row */
  /* scoring.py:1418-1572 / 1575-1750: entity rows [main | aux] of 2 d elements,
   * residual h^ (t~^ + o (+ r_bar)) - t^ (h~^ + o (- r_hat)) + r; `rel_u` carries the offset o;
   * relation rows [r] (InterHT) / [r | r_bar | r_hat] (TranS); own kernels in csrc/pair2.cu.
   * For shared negatives `cand_scale` holds 2 * n_cand inverse norms: main halves, then aux. */
  BESS_INTERHT = 7,
  BESS_TRANS = 8
} bess_family;

/* which entity the candidates replace: score_tails / score_heads */
typedef enum { BESS_MODE_TAILS = 0, BESS_MODE_HEADS = 1 } bess_mode;

/* losses (loss.py) */
typedef enum {
  BESS_LOSS_LOGSIGMOID = 0, BESS_LOSS_MARGIN_RANKING = 1, BESS_LOSS_SOFTMAX_CE = 2
} bess_loss_kind;

typedef enum { BESS_OPT_SGD = 0, BESS_OPT_SGDM = 1, BESS_OPT_ADAMW = 2 } bess_opt_kind;

/* Row addressing (up to three nested levels, outer level optional):
 *   x' = x % group1, a = x / group1          (skipped when group1 <= 0)
 *   row(x) = a * stride1 + (x' / group) * stride + x' % group + offset
 * group <= 0: row(x) = x + offset.  Lets kernels walk the reference's
 * [n_shard, p + B*Nn, D] exchange layout (bess.py:333-360) and its
 * transposed views in place. */
typedef struct {
  int32_t group, stride, offset, group1, stride1;
} bess_rowmap_t;

/* A set of rows: storage row = idx ? idx[row(x)] : row(x); address =
 * base + storage_row * pitch (pitch in ELEMENTS of the table dtype). */
typedef struct {
  const void* base;
  const int32_t* idx;
  bess_rowmap_t map;
  int64_t pitch;
} bess_rows_t;

/* score-function configuration (constructor arguments of scoring.py classes) */
typedef struct {
  int32_t family;      /* bess_family */
  int32_t norm_p;      /* scoring_norm, 1 or 2 (distance families) */
  int32_t d;           /* embedding_size as passed to the constructor */
  int32_t normalize;   /* PairRE normalize_entities */
  int32_t apply_tanh;  /* BoxE */
  int32_t per_dim;     /* BoxE dist_func_per_dim */
  float eps;           /* BoxE */
  float rel_u;         /* TripleRE v2 offset u added to both relation projections (0 = v1) */
} bess_score_cfg_t;

#define BESS_MAX_SHARD 16

const char* bess_last_error(void);
int bess_version(void);
/* number of kernels this library has launched in this process (host-side count) */
int64_t bess_launch_count(void);
/* entity / relation row widths (elements) implied by a config */
int bess_entity_width(const bess_score_cfg_t* cfg);
int bess_relation_width(const bess_score_cfg_t* cfg);
/* number of query-side vectors of entity width the negative kernels consume */
int bess_query_nvec(const bess_score_cfg_t* cfg);

/* ---------------------------------------------------------------- gather --
 * replaces `self.entity_embedding[gather_idx]` + torch.split + the data
 * movement of all_to_all (bess.py:331-355, 501-507) for ONE source shard.
 * idx = [n_local | n_dst * per_dst] local rows of `table`.  The first n_local
 * rows (heads) go to local_out row i.  Row y of the second part goes to
 * dst_out[y / per_dst] + (slot * per_dst + y % per_dst) * row_elems, i.e. into
 * destination replica (y / per_dst)'s receive buffer [n_src, per_dst, D] at
 * source slot `slot`.  dst_out[] may be local buffers, peer-mapped pointers
 * (NVLink stores) or slices of one NCCL send buffer. */
int bess_gather_route(const void* table, int64_t table_pitch, int dtype, int row_elems,
                      const int32_t* idx, int n_local, int n_dst, int per_dst, void* local_out,
                      void* const* dst_out /* host array [n_dst] */, int slot, void* stream);

/* plain gather out[i] = table[idx[i]] (bess.py:765,768,787) */
int bess_gather_rows(const void* table, int64_t table_pitch, int dtype, int row_elems,
                     const int32_t* idx, int n_idx, void* out, void* stream);

/* ---------------------------------------------------------- score_triple --
 * BaseScoreFunction.score_triple (scoring.py:45-65 and per family). */
int bess_score_triple_fwd(const bess_score_cfg_t* cfg, int dtype, bess_rows_t head,
                          bess_rows_t tail, const void* rel_table, const int32_t* rel_id,
                          bess_rowmap_t rel_map, int n, float* score, bess_rowmap_t score_map,
                          void* stream);
/* backward of the above given dL/dscore; fp32 gradient rows.  d_head / d_tail
 * rows are addressed like head / tail (base = fp32 buffer, pitch in floats);
 * d_rel is a per-query row buffer [n, rel_width] (reduced by relation later).
 * add_* != 0 accumulates. */
int bess_score_triple_bwd(const bess_score_cfg_t* cfg, int dtype, bess_rows_t head,
                          bess_rows_t tail, const void* rel_table, const int32_t* rel_id,
                          bess_rowmap_t rel_map, int n, const float* score,
                          const float* d_score, bess_rowmap_t score_map, bess_rows_t d_head,
                          bess_rows_t d_tail, float* d_rel, int add_head, int add_tail,
                          int add_rel, void* stream);

/* ---------------------------------------------------- query prologue ------
 * The relation-dependent half of score_heads / score_tails, e.g. h + r,
 * h o r, h (x) r, rotate(h, r) (scoring.py:342,354,446-448,460-462,825,837,
 * 928-932,944-946, 565-573, 1372-1389): fixed entity rows + relation ids ->
 * query vectors qv [n, nvec, W] fp32. */
int bess_query_prologue_fwd(const bess_score_cfg_t* cfg, int dtype, int mode, bess_rows_t fixed,
                            const void* rel_table, const int32_t* rel_id, bess_rowmap_t rel_map,
                            int n, float* qv, void* stream);
int bess_query_prologue_bwd(const bess_score_cfg_t* cfg, int dtype, int mode, bess_rows_t fixed,
                            const void* rel_table, const int32_t* rel_id, bess_rowmap_t rel_map,
                            int n, const float* d_qv, bess_rows_t d_fixed, float* d_rel,
                            int add_fixed, int add_rel, void* stream);
/* BoxE only: turn accumulated pre-chain relation gradients into gradients of
 * the raw relation row, in place (chain rule of scoring.py:1279-1309). */
int bess_boxe_rel_finalize(const bess_score_cfg_t* cfg, int dtype, const void* rel_table,
                           const int32_t* rel_id, int n, float* d_rel, void* stream);
/* PairRE candidate normalisation: inv_norm[i] = 1 / max(||cand_i||, 1e-12) */
int bess_cand_inv_norm(int dtype, bess_rows_t cand, int n, int width, float* inv_norm,
                       void* stream);
/* ... and its backward: d_cand_i = (g_i - c^_i (c^_i . g_i)) * inv_norm_i, in place on the
 * fp32 gradient rows g (addressed by d_cand). */
int bess_cand_norm_bwd(int dtype, bess_rows_t cand, int n, int width, const float* inv_norm,
                       bess_rows_t d_cand, void* stream);

/* -------------------------------------------- shared-negative scoring -----
 * broadcasted_distance / broadcasted_dot_product with negative_sample_sharing
 * (scoring.py:176-200, 231-255; pea.distance_matrix scoring.py:195-197) and
 * the PairRE / BoxE broadcast forms: scores[q, c] for every query q against
 * ONE shared candidate list.  out[(score_map(q)) * ld_out + col0 + c].
 * cand_scale: optional per-candidate factor (PairRE inv norms) or NULL.
 * aux: BoxE p=2 only, per-pair norm of the first box (needed by backward). */
int bess_score_shared_fwd(const bess_score_cfg_t* cfg, int dtype, int mode, const float* qv,
                          int n_query, bess_rows_t cand, const float* cand_scale, int n_cand,
                          float* out, bess_rowmap_t score_map, int64_t ld_out, int col0,
                          float* aux, void* stream);
/* backward w.r.t. the query vectors: d_qv [n_query, nvec, W] (overwritten) */
int bess_score_shared_bwd_query(const bess_score_cfg_t* cfg, int dtype, int mode, const float* qv,
                                int n_query, bess_rows_t cand, const float* cand_scale,
                                int n_cand, const float* score, const float* d_score,
                                bess_rowmap_t score_map, int64_t ld, int col0, const float* aux,
                                float* d_qv, void* stream);
/* backward w.r.t. the candidate rows: fp32 rows addressed by d_cand
 * (overwritten, or accumulated when add_cand != 0).  workspace: >=
 * bess_shared_bwd_cand_workspace() bytes. */
int64_t bess_shared_bwd_cand_workspace(const bess_score_cfg_t* cfg, int n_query, int n_cand);
int bess_score_shared_bwd_cand(const bess_score_cfg_t* cfg, int dtype, int mode, const float* qv,
                               int n_query, bess_rows_t cand, const float* cand_scale,
                               int n_cand, const float* score, const float* d_score,
                               bess_rowmap_t score_map, int64_t ld, int col0, const float* aux,
                               bess_rows_t d_cand, int add_cand, void* workspace, void* stream);

/* ------------------------- tensor-core path of the DOT scorers (tcgen05) ---
 * `torch.matmul(v1, v2.T)` of broadcasted_dot_product (scoring.py:252) and the
 * two contractions of its autograd backward, as one TMA + tcgen05.mma + TMEM
 * kernel:   out[out_map(m) * ld_out + col0 + n] (+)= sum_k A[m, k] * B[n, k].
 * A [M, K] and B [N, K] are dense K-major arrays (leading dimensions lda / ldb
 * in elements, 16-byte multiples) produced by bess_split_operand:
 *   dtype BESS_F32        : fp32 arrays hi / lo (3xTF32: hi*hi + hi*lo + lo*hi,
 *                           fp32-grade products, fp32 accumulate in TMEM)
 *   dtype BESS_F16 / BF16 : half arrays in *_hi (one MMA per k-step), *_lo unused
 *   dtype BESS_F16X3      : fp16 arrays hi / lo of the SCALED operand, x * s = hi + lo with
 *                           s a power of two chosen per operand matrix by bess_operand_scale
 *                           (3xFP16: the same three products as 3xTF32 — fp16 and tf32 share
 *                           the 11-bit significand — at twice the tensor-core rate and half
 *                           the operand bytes); a_scale / b_scale point at the operands'
 *                           {s, 1 / s} pairs on the device and the epilogue multiplies the
 *                           accumulator by the two inverse scales (exact).  NULL otherwise.
 * a_mn_major != 0: A is given transposed, as [K, M] with M contiguous (leading
 * dimension lda) and consumed through MN-major UMMA descriptors — how the
 * dC = dS^T Q contraction reads the [S, N] score gradient without a transposed copy.
 * Small output grids are split over K; `workspace` (>= bess_dot_gemm_workspace
 * bytes, may be NULL when that is 0) holds the partial sums, reduced in a fixed
 * order (deterministic). */
int64_t bess_dot_gemm_workspace(int M, int N, int K);
int bess_dot_gemm(int dtype, const void* a_hi, const void* a_lo, int64_t lda, int a_mn_major,
                  const void* b_hi, const void* b_lo, int64_t ldb, int M, int N, int K, float* out,
                  bess_rowmap_t out_map, int64_t ld_out, int col0, int accumulate, void* workspace,
                  int64_t workspace_bytes, const float* a_scale, const float* b_scale, void* stream);
/* Operand pre-pass for bess_dot_gemm: rows of `src` (dtype src_dtype, addressed
 * through map / idx / pitch, optionally scaled per row) -> dense operand arrays
 * hi / lo [n_rows, ld] and / or their transposes hiT / loT [width, ldT].
 * out_dtype BESS_F32: hi = rna_tf32(x), lo = rna_tf32(x - hi); BESS_F16 / BF16:
 * hi = x rounded to that type (lo / loT ignored); BESS_F16X3: fp16 arrays,
 * hi = fp16(x * s), lo = fp16(x * s - hi) with s = op_scale[0] (device, from
 * bess_operand_scale; NULL for the other formats).  Any output may be NULL. */
int bess_split_operand(int src_dtype, bess_rows_t src, int n_rows, int width,
                       const float* row_scale, int out_dtype, void* hi, void* lo, int64_t ld,
                       void* hiT, void* loT, int64_t ldT, const float* op_scale, void* stream);
/* Power-of-two scale of a BESS_F16X3 operand: scale[0] = s = 2^(14 - e), scale[1] = 1 / s,
 * where factor * max |x| over the rows (times |row_scale|) = m * 2^e, m in [0.5, 1): the
 * largest scaled element lies in [2^13, 2^14].  One kernel, no host sync; `state` = 2 x
 * uint32 owned by the caller, zero-initialised once (the kernel leaves it zero). */
int bess_operand_scale(int src_dtype, bess_rows_t src, int n_rows, int width,
                       const float* row_scale, float factor, float* scale, void* state,
                       void* stream);
/* Cached 3xTF32 operands of a whole fp32 table for inference (TopKQueryBessKGE scores
 * every window of a constant shard on every call, bess.py:771-853): streams the
 * table once to form a 64-bit position-dependent checksum, compares it ON THE
 * DEVICE with the one kept in state[1] and re-builds hi = rna_tf32(x) /
 * lo = rna_tf32(x - hi) [n_rows, ld] only when the bytes changed or force != 0.
 * state: 4 x uint64 owned by the caller, zero-initialised; after the call
 * state[3] == 1 iff the operands were rebuilt.  Stream-ordered, no host sync. */
int bess_table_operand_refresh(const float* table, int64_t n_rows, int width, int64_t pitch,
                               float* hi, float* lo, int64_t ld, uint64_t* state, int force,
                               void* stream);

/* --------------------------- norm-expanded L2 distance on the tensor cores ---
 * `pea.distance_matrix(v1, v2, p=2)` (scoring.py:195-197) for shared negatives as
 * ||q - c||^2 = ||q||^2 + ||c||^2 - 2 q.c : the q.c block is one bess_dot_gemm, these are the
 * kernels around it.  Backward with b = (dL/dscore) / dist: dQ = B C - rb * Q, dC = B^T Q - cb * C
 * (two more bess_dot_gemm + row scalings).  All reductions run in a fixed order.
 *   bess_row_sqnorm : out[i] = sum_k row_i[k]^2
 *   bess_l2_from_dot: score[map(q), col0 + c] <- -sqrt(max(qn[q] + cn[c] - 2 score, 0)) in place
 *   bess_l2_coef    : coef[q, c] = -d_score / score (0 where score == 0), row_sum[q], col_sum[c];
 *                     workspace of bess_l2_coef_workspace(n_query, n_cand) bytes
 *   bess_rows_axpy  : out_i[k] += scale * alpha[i] * src_i[k]  (out rows fp32) */
int bess_row_sqnorm(int dtype, bess_rows_t rows, int n, int width, float* out, void* stream);
int bess_l2_from_dot(float* score, bess_rowmap_t score_map, int64_t ld, int col0, int n_query,
                     int n_cand, const float* q_sqnorm, const float* c_sqnorm, void* stream);
int64_t bess_l2_coef_workspace(int n_query, int n_cand);
int bess_l2_coef(const float* d_score, const float* score, bess_rowmap_t score_map, int64_t ld,
                 int col0, int n_query, int n_cand, float* coef, int64_t ld_coef, float* row_sum,
                 float* col_sum, void* workspace, void* stream);
int bess_rows_axpy(int dtype, const float* alpha, float scale, bess_rows_t src, bess_rows_t out,
                   int n, int width, void* stream);

/* ------------------------------------------ per-triple negative scoring ---
 * negative_sample_sharing == False: reduce_embedding(v1.unsqueeze(1) - v2)
 * (scoring.py:199, 254): query q against its OWN n_per candidates; candidate
 * c of the query at position qpos = score_map(q) is logical row
 * cand.map(c) + qpos * cand_q_stride of `cand` (then cand.idx, if any), so the
 * table can be read in place: fused gather + score.  d_cand is addressed the
 * same way. */
int bess_score_pertriple_fwd(const bess_score_cfg_t* cfg, int dtype, int mode, const float* qv,
                             int n_query, bess_rows_t cand, int64_t cand_q_stride, int n_per,
                             float* out, bess_rowmap_t score_map, int64_t ld_out, int col0,
                             float* aux, void* stream);
int bess_score_pertriple_bwd(const bess_score_cfg_t* cfg, int dtype, int mode, const float* qv,
                             int n_query, bess_rows_t cand, int64_t cand_q_stride, int n_per,
                             const float* score, const float* d_score, bess_rowmap_t score_map,
                             int64_t ld, int col0, const float* aux, float* d_qv,
                             bess_rows_t d_cand, void* stream);

/* ------------------------------------------------------- masks and loss ---
 * score[r, c] += value where mask[r * ld_mask + c] == flag
 * (BAD_NEGATIVE_SCORE handling, bess.py:201-245, 802-806). */
int bess_mask_add(float* score, int n_row, int n_col, int64_t ld, const uint8_t* mask,
                  int64_t ld_mask, int mask_rows, int flag, float value, void* stream);
/* augment_negative diagonal mask (bess.py:201-226): score[r, diag_col(r)] += value */
int bess_mask_diag(float* score, int n_row, int64_t ld, int step, int half_group, int group,
                   float value, void* stream);

/* loss forward + gradient w.r.t. the scores (loss.py:115-134, 179-195,
 * 226-251).  weight: [n] or single value (weight_n == 1).  Outputs: row_loss
 * [n] partials and, via bess_sum_f32, the replica loss; d_pos [n], d_neg
 * [n, n_neg] (may alias neg for in-place).  For SOFTMAX_CE `neg` is adjusted
 * in place first like the reference (loss.py:233-237). */
int bess_loss_fwd_bwd(int kind, float margin, int adversarial, float adv_scale, float loss_scale,
                      int64_t n_entity, const float* pos, float* neg, int n, int n_neg,
                      int64_t ld, const float* weight, int weight_n, float* row_loss,
                      float* d_pos, float* d_neg, void* stream);
/* Same loss, but dL/dneg is written directly in the operand form bess_dot_gemm
 * consumes for the backward contractions: grad_dtype BESS_F32 -> d_neg_hi =
 * rna_tf32(g), d_neg_lo = rna_tf32(g - hi) (fp32 arrays); BESS_BF16 / BESS_F16 ->
 * d_neg_hi = g rounded to that type, d_neg_lo unused; BESS_F16X3 -> fp16 arrays,
 * d_neg_hi = fp16(g * s), d_neg_lo = fp16(g * s - hi) with s = grad_scale[0] (device; the
 * caller derives it from the bound |g| <= loss_scale * max(weight) with bess_operand_scale
 * over the weight vector).  ld_grad in elements; grad_scale NULL for the other formats. */
int bess_loss_fwd_bwd_operand(int kind, float margin, int adversarial, float adv_scale,
                              float loss_scale, int64_t n_entity, const float* pos, float* neg, int n,
                              int n_neg, int64_t ld, const float* weight, int weight_n,
                              float* row_loss, float* d_pos, int grad_dtype, void* d_neg_hi,
                              void* d_neg_lo, int64_t ld_grad, const float* grad_scale,
                              void* stream);
/* deterministic sum of n floats (fixed order) -> out[0] */
int bess_sum_f32(const float* x, int n, float* out, void* stream);

/* ranks_from_scores (metric.py:129-183): mode 0 optimistic, 1 pessimistic,
 * 2 average.  rank[n] fp32 (inf when worst_rank_infty and nothing is beaten). */
int bess_rank_from_scores(const float* pos, const float* neg, int n, int n_neg, int64_t ld,
                          int mode, int worst_rank_infty, float* rank, void* stream);

/* ------------------------------------------------- backward: scatter ------
 * Stable LSD radix sort of n (key, position) pairs; keys in [0, 2^key_bits).
 * workspace >= bess_sort_workspace(n) bytes.  perm_out[i] = original position
 * of the i-th smallest key (ties keep input order -> deterministic sums). */
int64_t bess_sort_workspace(int n);
int bess_sort_keys(const int32_t* keys, int n, int key_bits, int32_t* keys_out,
                   int32_t* perm_out, void* workspace, void* stream);

/* Deterministic segmented scatter-add + optimizer on one shard (implicit in the
 * reference: autograd of bess.py:333-337 followed by the optimizer step).
 * For each run of equal sorted keys, the fp32 gradient rows grad[perm[.]] are
 * summed in sorted (= input) order and applied to table row `key`.
 *   SGD   (no momentum / weight decay): sparse update, exact.
 *   SGDM / ADAMW: the segment sums are stored to seg_grad[first position of the
 *   run] and row_to_seg[key]; bess_opt_dense then updates EVERY row (dense
 *   torch.optim semantics: rows without gradient still decay / coast).
 * grad rows: position x < n_local -> grad_local row x; else y = x - n_local ->
 * grad_dst row (y / per_dst) * dst_stride_rows + y % per_dst  (mirror of
 * bess_gather_route's routing).
 *
 * `hyper` (device pointer, may be NULL): when given, the optimizer hyper-parameters are
 * read from this fp32 array on the device (slots BESS_HYPER_*) instead of the by-value
 * arguments, so a captured CUDA graph follows a learning-rate schedule and AdamW's bias
 * correction: the host rewrites the array before each replay. */
enum {
  BESS_HYPER_LR = 0, BESS_HYPER_MOMENTUM = 1, BESS_HYPER_DAMPENING = 2, BESS_HYPER_BETA1 = 3,
  BESS_HYPER_BETA2 = 4, BESS_HYPER_EPS = 5, BESS_HYPER_WEIGHT_DECAY = 6,
  BESS_HYPER_BC1 = 7 /* 1 - beta1^step */, BESS_HYPER_BC2 = 8 /* 1 - beta2^step */,
  BESS_HYPER_FIRST_STEP = 9 /* 1.0 on the first step (momentum buffer := grad) */,
  BESS_HYPER_COUNT = 16
};
int bess_scatter_sgd(void* table, int64_t table_pitch, int dtype, int row_elems,
                     const int32_t* sorted_keys, const int32_t* perm, int n, int n_local,
                     int per_dst, const float* grad_local, const float* grad_dst,
                     int64_t dst_stride_rows, float lr, const float* hyper, void* stream);
int bess_scatter_collect(int row_elems, const int32_t* sorted_keys, const int32_t* perm, int n,
                         int n_local, int per_dst, const float* grad_local,
                         const float* grad_dst, int64_t dst_stride_rows, float* seg_grad,
                         int32_t* row_to_seg, void* stream);
/* gradient accumulation over micro-batches (the reference's runtime option
 * `options.Training.gradientAccumulation(k)`, notebooks 1 cell 26 / 2 cell 14): add the
 * segment sums of one micro-batch to a dense fp32 accumulator acc [Es, row_elems]
 * (deterministic: one warp owns a key). */
int bess_scatter_accumulate(int row_elems, const int32_t* sorted_keys, const int32_t* perm, int n,
                            int n_local, int per_dst, const float* grad_local,
                            const float* grad_dst, int64_t dst_stride_rows, float* acc,
                            void* stream);
/* dense optimizer pass over [n_rows, row_elems]; row_to_seg[row] < 0: zero grad;
 * row_to_seg == NULL: seg_grad is a dense [n_rows, row_elems] gradient.
 * state0: momentum buffer / Adam m; state1: Adam v (fp32).  step: 1-based.
 * The gradient is multiplied by grad_scale (1/k for a mean over k accumulated
 * micro-batches); zero_grad != 0 (dense gradient only) clears seg_grad after reading it. */
int bess_opt_dense(int kind, void* table, int64_t table_pitch, int dtype, int n_rows,
                   int row_elems, const float* seg_grad, const int32_t* row_to_seg, float* state0,
                   float* state1, float lr, float momentum, float dampening, float beta1,
                   float beta2, float eps, float weight_decay, int step, const float* hyper,
                   float grad_scale, int zero_grad, void* stream);
/* fill one BESS_HYPER_COUNT-float device array for optimizer step `step` (1-based):
 * the by-value hyper-parameters plus bc1 = 1 - beta1^step, bc2 = 1 - beta2^step and the
 * first-step flag.  Kernel arguments are captured at launch, so the host may call this
 * every step without synchronising. */
int bess_set_hyper(float* hyper, float lr, float momentum, float dampening, float beta1,
                   float beta2, float eps, float weight_decay, int step, void* stream);
/* relation table: deterministic reduce of per-query gradient rows by relation
 * id -> d_table [n_rel, width] fp32 (overwritten).  sorted_rel/perm from
 * bess_sort_keys over the relation ids. */
int bess_relation_grad_reduce(const float* d_rel_rows, int width, const int32_t* sorted_rel,
                              const int32_t* perm, int n, int n_rel, float* d_table,
                              void* stream);

/* ------------------------------------------------------------ top-k -------
 * Running top-k over a window of candidates (bess.py:771-822): merges the
 * scores of `n_win` new candidates (ids win_ids or win_id0 + c) into the
 * per-query best lists best_score/best_id [n_query, k] (sorted descending,
 * ties keep the earlier entry). */
int bess_topk_merge(const float* win_score, int64_t ld, int n_query, int n_win,
                    const int32_t* win_ids, int64_t ld_ids, int win_id0, float* best_score,
                    int32_t* best_id, int k, void* stream);

/* Exact re-rank of the running best lists (csrc/exact.cu).  The
 * window scorers only select candidates; this call re-scores the k_in (<= 32) entries of
 * every query's list — local row ids ids_in [n_query, k_in] into `table` — in ONE fixed fp32
 * summation order (coordinate 0..row_elems-1, each product / sum rounded separately, no FMA)
 * and writes the k_out best, ordered by (score descending, id ascending), to score_out /
 * ids_out [n_query, k_out].  `fixed` [n_query, row_elems]: the known entity row of each query
 * (head for BESS_MODE_TAILS, tail for BESS_MODE_HEADS); rel_id [n_query].  Entries with
 * id >= n_table_rows (empty slots) keep score_in.  TransE / DistMult / ComplEx only
 * (bess_topk_exact_supported): their arithmetic is + - x sqrt, reproducible bit for bit by
 * the CPU oracle, so ranks and MRR can be asserted equal with no near-tie tolerance. */
int bess_topk_exact_supported(int family);
int bess_topk_exact_rescore(const bess_score_cfg_t* cfg, int dtype, int mode, const void* fixed,
                            int64_t fixed_pitch, const void* rel_table, int64_t rel_pitch,
                            const int32_t* rel_id, const void* table, int64_t table_pitch,
                            int n_table_rows, int row_elems, const int32_t* ids_in,
                            const float* score_in, int n_query, int k_in, int k_out,
                            float* score_out, int32_t* ids_out, void* stream);

/* Final step of TopKQueryBessKGE (bess.py:866-891) on the shard that owns the
 * queries: score / idx [n_shard, n_query, kb] are the best lists received from
 * every scoring shard.  Adds bad_score to entries whose local id is a padding
 * row of its shard (idx >= shard_counts[j]), maps local -> global ids through
 * shard_idx_to_entity [n_shard, max_entity_per_shard] and writes the k best
 * (score descending, ties by position) to out_score / out_id [n_query, k]. */
int bess_topk_finalize(const float* score, const int32_t* idx, int n_shard, int n_query, int kb,
                       const int32_t* shard_counts, const int32_t* shard_idx_to_entity,
                       int max_entity_per_shard, int k, float bad_score, float* out_score,
                       int32_t* out_id, void* stream);

/* ------------------------------------------------ AllScores post-processing ---
 * Device forms of the host fancy-indexing in AllScoresPipeline.forward
 * (pipeline.py:266-298).
 * bess_select_scores: out[i, e] = col_idx[e] >= 0 ? src[row_idx[i], col_idx[e]] : fill
 *   for i < n_rows, e < n_cols (row_idx NULL = identity): block-column scores ->
 *   global-entity order, padding queries dropped, non-candidates filled (-inf).
 * bess_pairs_get / bess_pairs_set: out[t] = mat[rows[t], cols[t]] and
 *   mat[rows[t], cols[t]] = values ? values[t] : value (rows NULL = t): ground-truth
 *   score read / restore, sparse (query, entity) filters. */
int bess_select_scores(const float* src, int64_t ld_src, const int32_t* row_idx, int n_rows,
                       const int32_t* col_idx, int n_cols, float fill, float* out, int64_t ld_out,
                       void* stream);
int bess_pairs_get(const float* mat, int64_t ld, const int32_t* rows, const int32_t* cols, int n,
                   float* out, void* stream);
int bess_pairs_set(float* mat, int64_t ld, const int32_t* rows, const int32_t* cols, int n,
                   const float* values, float value, void* stream);

/* ------------------------------------------- peer-memory exchange ---------
 * One process per GPU, one entity shard per GPU: the balanced AllToAll of
 * bess.py:348-350 (and its autograd transpose, and the all-reduce of the
 * replicated relation gradient) without a collective library inside the step.
 * bess_gather_route already stores rows through peer-mapped pointers; these
 * four calls add the flag handshake and the pushes.  `peer_flags[j]` points at
 * rank j's flag row (n int32, symmetric memory) of one channel, `counter` is a
 * local int32 sequence counter of the same channel.
 *   signal: seq = ++*counter; flag row of every rank [my_rank] := seq (release.sys)
 *   wait  : until all n local flags >= *counter (acquire.sys).  timeout_ms > 0 bounds the
 *           spin in wall-clock time (%globaltimer): on expiry the launch prints the
 *           missing rank and traps; timeout_ms <= 0 waits for ever
 *   push  : dst[j] <- src + j * src_stride_bytes, bytes_each bytes, for j < n (an SM kernel:
 *           remote stores)
 *   copy  : the same transfer for destination j alone, issued to a copy engine
 *           (cudaMemcpyAsync, a memcpy node under stream capture): no SM is involved, so it
 *           does not slow a GEMM that runs next to it; the caller spreads the destinations
 *           over a few streams to keep several engines busy
 *   reduce: out[i] = scale * sum_j slots[j * count + i], j ascending */
int bess_peer_signal(int32_t* counter, void* const* peer_flags /* host array [n] */, int my_rank,
                     int n, void* stream);
int bess_peer_wait(const int32_t* counter, const int32_t* my_flags, int n, int64_t timeout_ms,
                   void* stream);
int bess_peer_push(const void* src, int64_t src_stride_bytes, void* const* dst /* host array [n] */,
                   int n, int64_t bytes_each, void* stream);
int bess_peer_copy(const void* src, void* dst, int64_t bytes, void* stream);
int bess_peer_reduce(const float* slots, int n, int64_t count, float scale, float* out, void* stream);

/* Peer-mapped receive buffers through CUDA IPC (no framework involved): alloc (zero-filled,
 * synchronous) -> export a BESS_IPC_HANDLE_BYTES-byte handle -> ship it to the peers over any
 * host channel -> import maps the peer's buffer into this process (between GPUs of a node, or
 * between two processes sharing one GPU).  A process must not import its own handle. */
#define BESS_IPC_HANDLE_BYTES 64
int bess_peer_alloc(int64_t bytes, void** ptr);
int bess_peer_free(void* ptr);
int bess_peer_export(void* ptr, void* handle_out /* BESS_IPC_HANDLE_BYTES */);
int bess_peer_import(const void* handle, void** ptr);
int bess_peer_unmap(void* ptr);

/* Python-surface helpers of the reference's utils.py.
 * take_along_rows (utils.py:10-33 `gather_indices`): out[i, j] = x[i or 0, index[i or 0, j]];
 *   x [a, e] of elem_bytes-wide words (1/2/4/8), index int32 [b, k], a == b or one of them 1,
 *   out [max(a, b), k].
 * complex_mul (utils.py:72-112): rows are [re | im] halves of e elements each;
 *   rotate == 0: out = v1 * v2 (both [n, 2e]); rotate != 0: v2 is [n, e] angles and
 *   out = v1 * (cos v2 + i sin v2). */
int bess_take_along_rows(const void* x, int a, int64_t e, int elem_bytes, const int32_t* index,
                         int b, int k, void* out, void* stream);
int bess_complex_mul(int dtype, const void* v1, const void* v2, int n, int e, int rotate, void* out,
                     void* stream);

/* profiling aid: out[0] <- %globaltimer (ns) when the stream reaches this point */
int bess_stamp(uint64_t* out, void* stream);

/* utility */
int bess_fill_f32(float* p, int64_t n, float v, void* stream);
int bess_fill_i32(int32_t* p, int64_t n, int32_t v, void* stream);
int bess_cast_from_f32(const float* src, void* dst, int dtype, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BESSKGE_B200_H */
