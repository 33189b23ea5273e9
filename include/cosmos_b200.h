/*
 * cosmos_b200 - C ABI of the B200-native COSMOS loss head.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry point takes
 * the CUDA device ordinal and stream explicitly, keeps no host-side state between calls, allocates
 * nothing (workspaces are caller-provided device memory) and returns a status code
 * (0 = COSMOS_OK; cosmos_status_string() names the others).  All device pointers are raw
 * CUdeviceptr values; matrices are row-major, contiguous and 16-byte aligned.
 *
 * What each group replaces in the reference (paths relative to the geniusxxx/cosmos checkout):
 *   cosmos_infonce_*   src/open_clip/loss.py:103-142  ClipLoss.get_logits + F.cross_entropy x2 per pair,
 *                      as composed by COSMOSLoss.forward src/open_clip/loss.py:176-207
 *   cosmos_ema_*       src/training/train.py:195-203  per-parameter mul_/add_ loop
 *   cosmos_clamp_scalars  src/training/train.py:237-243  logit-scale clamps of student and teacher
 *   cosmos_retrieval_ranks  src/training/train.py:712-785  eval similarity + argsort + rank search (outside the step)
 *   cosmos_xpool_*     src/open_clip/transformer.py:210-230 AttentionalCrossPooler.forward and its call
 *                      site src/open_clip/model.py:366-387
 * The reference has no FFI of its own (pure PyTorch); INTEGRATION.md shows the ctypes binding and
 * the two import lines a maintainer changes.
 */
#ifndef COSMOS_B200_H_
#define COSMOS_B200_H_

#include <stddef.h>
#include <stdint.h>

/* The library is built with -fvisibility=hidden: only the functions declared below are exported. */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
#ifdef __cplusplus
extern "C" {
#endif

#define COSMOS_B200_ABI_VERSION 3

/* status codes */
#define COSMOS_OK 0
#define COSMOS_ERR_INVALID_ARGUMENT 1 /* bad shape / null pointer / misalignment */
#define COSMOS_ERR_UNSUPPORTED 2      /* dtype or size the kernels are not built for */
#define COSMOS_ERR_CUDA 3             /* a CUDA runtime call or launch failed */
#define COSMOS_ERR_NO_DEVICE 4        /* device is not an sm_100 part */
#define COSMOS_ERR_WORKSPACE 5        /* caller workspace too small */

/* element types */
#define COSMOS_DTYPE_F32 0
#define COSMOS_DTYPE_BF16 1
#define COSMOS_DTYPE_F16 2

int cosmos_abi_version(void);
const char* cosmos_status_string(int status);
/* Diagnostics: the last CUDA error code a cosmos_* call made on THIS host thread ran into, and its name. */
int cosmos_last_cuda_error(void);
const char* cosmos_cuda_error_string(int code);
/* 0 when `device` can run the kernels (compute capability 10.x), else COSMOS_ERR_NO_DEVICE / _CUDA. */
int cosmos_device_check(int device);

/* ------------------------------------------------------------------------------------------------
 * EMA teacher update           (train.py:200-203:  k.mul_(m).add_((1 - m) * q)  for every parameter)
 * ------------------------------------------------------------------------------------------------
 * The parameter set is described once by a chunk table (host-built, then copied to the device by
 * the caller); each step is one kernel launch over that table.
 */
#define COSMOS_EMA_CHUNK 8192 /* elements per table entry */

typedef struct cosmos_ema_chunk {
  uint64_t teacher; /* device address of the first element of this chunk (updated in place) */
  uint64_t student; /* device address of the matching student elements (read only)         */
  uint32_t count;   /* elements in this chunk, <= COSMOS_EMA_CHUNK                          */
  uint32_t aligned; /* 1 when both addresses are 16-byte aligned                            */
} cosmos_ema_chunk;

/* Number of table entries needed for n_tensors tensors with the given element counts. */
int64_t cosmos_ema_table_entries(int64_t n_tensors, const int64_t* numel);
/* Fill `table_host` (cosmos_ema_table_entries() entries). elem_size is 4 (f32) or 2 (bf16/f16). */
int cosmos_ema_table_fill(int64_t n_tensors, const uint64_t* teacher_ptrs, const uint64_t* student_ptrs,
                          const int64_t* numel, int elem_size, cosmos_ema_chunk* table_host);
/* One EMA step over a device-resident table.  `momentum` is the Python double of the reference. */
int cosmos_ema_apply(const cosmos_ema_chunk* table_dev, int64_t n_entries, double momentum, int dtype,
                     int device, void* stream);

/* Logit-scale clamp (train.py:237-243: logit_scale.clamp_(0, ln 100) on student and teacher, and the same for
 * distill_logit_scale - four 1-element launches in the reference).  `ptrs` is a HOST array of n <= COSMOS_CLAMP_MAX
 * device addresses of scalars of type `dtype`; all are clamped in place to [lo, hi] by one launch, with
 * torch.clamp_ semantics (bounds rounded to fp32, NaN kept). */
#define COSMOS_CLAMP_MAX 8
int cosmos_clamp_scalars(const uint64_t* ptrs, int32_t n, double lo, double hi, int dtype, int device, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Block-structured InfoNCE  (loss.py:103-142 for every (row tensor i, column tensor j) pair at once)
 * ------------------------------------------------------------------------------------------------
 * x is a stack of gx row-side tensors [n_rows, dim] (this rank's rows), y a stack of gy column-side
 * tensors [n_cols, dim] (all ranks' rows, gathered by the caller).  For pair p = i*gy + j the logits
 * are S = scale * x_i y_j^T and row r's positive is column label_offset + r.  The N x N logits are
 * never written to memory: the kernels keep them in TMEM tiles.
 *
 * All log-sum-exps are exchanged in log2 units (lse2 = log2(sum_c 2^(S*log2(e)))), diagonals as raw
 * dot products (S / scale).
 */
typedef struct cosmos_infonce_problem {
  uint64_t x;           /* device address, [gx][n_rows][dim], bf16 or fp16            */
  uint64_t y;           /* device address, [gy][n_cols][dim], same dtype              */
  int32_t gx, gy;       /* tensors in each stack                                       */
  int32_t n_rows;       /* rows per row-side tensor (local batch)                      */
  int32_t n_cols;       /* rows per column-side tensor (global batch)                  */
  int32_t dim;          /* embedding dim: multiple of 64, <= 512                       */
  int32_t label_offset; /* rank * n_rows                                               */
  int32_t dtype;        /* COSMOS_DTYPE_BF16 or COSMOS_DTYPE_F16                       */
  int32_t reserved;
  uint64_t scale;       /* device address of one fp32 (logit scale, exp already taken) */
} cosmos_infonce_problem;

/* Bytes of device workspace cosmos_infonce_fwd / _bwd need for this problem (the larger of the two). */
int64_t cosmos_infonce_workspace_bytes(const cosmos_infonce_problem* p);

/* Forward pass.  Outputs (fp32, device):
 *   row_lse2 [gx*gy][n_rows]  log2-sum-exp of every local row over all n_cols columns
 *   diag_raw [gx*gy][n_rows]  x_i[r] . y_j[label_offset + r]
 *   col_lse2 [gx*gy][n_cols]  log2-sum-exp of every column over THIS rank's n_rows rows
 *                             (the caller combines ranks with a log-sum-exp all-reduce)         */
int cosmos_infonce_fwd(const cosmos_infonce_problem* p, float* row_lse2, float* diag_raw, float* col_lse2,
                       void* workspace, int64_t workspace_bytes, int device, void* stream);

/* dst[k] = src[k] * (*num / den) for n 16-bit values (n % 8 == 0, 16-byte aligned; products in fp32, rounded once): the
 * gradients of the stored-exponential route are formed in the forward for a power-of-two stand-in `den` of the upstream
 * gradient `*num` (a device scalar: GradScaler's factor, src/training/train.py:62-66) and scaled when backward() learns it. */
int cosmos_scale16(const void* src, void* dst, const float* num, float den, int32_t dtype, int64_t n, int device, void* stream);

/* Merge of per-rank partial column statistics (row-sharded forward, src/open_clip/loss.py:21-65 gathers the features
 * instead): parts [n_parts][n] fp32 log2-sum-exp values (every rank's col_lse2, all-gathered), out [n] their
 * log2-sum-exp2.  -inf (a rank without rows) is the neutral element.                                              */
int cosmos_lse2_merge(const float* parts, float* out, int32_t n_parts, int64_t n, int device, void* stream);

/* Per-pair loss sums (natural-log units), out[p][0] = sum_r (LSE_row[r] - S[r, label(r)]) over local rows,
 * out[p][1] = sum_c (LSE_col[c] - S[c - label_offset, c]) over this rank's n_rows diagonal columns, with
 * col_lse2 the GLOBAL column log-sum-exp (already combined across ranks).  out: fp32 [gx*gy][2].          */
int cosmos_infonce_loss_sums(const cosmos_infonce_problem* p, const float* row_lse2, const float* diag_raw,
                             const float* col_lse2, int32_t use_rows, int32_t use_cols, float* out,
                             void* workspace, int device, void* stream);

/* Backward pass.  With R = softmax over columns of each row (from row_lse2), C = softmax over rows of each
 * column (from the global col_lse2) and I the positives, per pair
 *     G  = a_row * R + a_col * C - (a_row + a_col) * I
 *     dx_i = (*upstream) * weight * scale * sum_j G_ij y_j             (written in the stack dtype)
 *     dscale = (*upstream) * weight * sum_ij <s_row * R + s_col * C - (s_row + s_col) * I, x_i y_j^T>
 * dx [gx][n_rows][dim] and dscale (one fp32) may be NULL to skip either output.                       */
int cosmos_infonce_bwd(const cosmos_infonce_problem* p, const float* row_lse2, const float* col_lse2,
                       float a_row, float a_col, float s_row, float s_col, float weight, const float* upstream,
                       void* dx, float* dscale, void* workspace, int64_t workspace_bytes, int device, void* stream);

/* Same, and additionally stores G itself (stack dtype, without the upstream * weight * scale factor):
 *     g_out[(i * n_rows + r) * g_ld + j * n_cols + c] = G_ij[r][c]
 * so that the gradient of the COLUMN side, dy_j = (*upstream) * weight * scale * sum_i G_ij^T x_i (loss.py:103-119: both
 * logits_per_image and logits_per_text back-propagate into both feature lists), becomes one plain GEMM per column tensor
 * (cosmos_gemm with a_kmajor = b_kmajor = 0) instead of a second sweep that recomputes every logit.  dim 512 and dx != NULL
 * only (COSMOS_ERR_UNSUPPORTED otherwise); n_cols and g_ld multiples of 8, g_ld >= gy * n_cols, g_out 16-byte aligned. */
int cosmos_infonce_bwd_g(const cosmos_infonce_problem* p, const float* row_lse2, const float* col_lse2,
                         float a_row, float a_col, float s_row, float s_col, float weight, const float* upstream,
                         void* dx, float* dscale, void* g_out, int64_t g_ld, void* workspace, int64_t workspace_bytes,
                         int device, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Cross-attention pooler building blocks
 *   (transformer.py:210-230 AttentionalCrossPooler.forward = LayerNorm x2 + nn.MultiheadAttention with one
 *    query per crop; model.py:378-387 residual add + F.normalize).  cosmos_b200/pooler.py composes them.
 * ------------------------------------------------------------------------------------------------ */

/* D[M,N] (+)= alpha * opA * opB^T (+ bias[N]) on tcgen05 tensor cores, fp32 accumulation.
 *   a_kmajor = 1: A stored [M, K] row-major with row stride lda;  0: stored [K, M] (used for weight gradients)
 *   b_kmajor likewise for B [N, K] / [K, N].  in_dtype bf16/f16; out_dtype f32/bf16/f16.
 *   splits > 1 splits K over gridDim.z and accumulates with fp32 atomics into D, which the caller has zeroed
 *   (out_dtype is then forced to f32).  Strides in elements, multiples of 8; pointers 16-byte aligned.       */
int cosmos_gemm(const void* a, const void* b, void* d, const float* bias, int32_t M, int32_t N, int32_t K,
                int64_t lda, int64_t ldb, int64_t ldd, int32_t a_kmajor, int32_t b_kmajor, int32_t in_dtype,
                int32_t out_dtype, int32_t splits, float alpha, int device, void* stream);

/* The same GEMM for `batch` independent problems of one shape: problem t reads a + t * stride_a, b + t * stride_b (elements;
 * multiples of 8, and they may be SMALLER than the row strides: the per-head column blocks of one matrix are a batch), adds
 * bias + t * stride_bias and writes d + t * stride_d.  Tiles past M, N or K of one problem read zeros, never a neighbour's
 * rows (each problem is its own slice of a 3-D tensor map).  accumulate != 0: D += ... (splits must be 1).
 * K2 > 0: a second operand pair of the same storage orders, D = alpha * (A B^T + A2 B2^T) in one pass over the output (the
 * token gradient of the folded attention is P dZ + dS Q~).
 * Used by the folded attention of the pooler (cosmos_b200/pooler.py): with one query per (sample, crop) the key projection
 * of nn.MultiheadAttention (transformer.py:214, 225-229) folds into the queries, q~ = W_k,h^T q_h, so that per sample
 * scores = LN(x) q~^T and pooled_h = W_v,h (P^T LN(x)): four small GEMMs per sample instead of the [L, d] x [d, 2d] key /
 * value projection, and no key / value tensor in memory.                                                                   */
int cosmos_gemm_batched(const void* a, const void* b, void* d, const float* bias, int32_t M, int32_t N, int32_t K, int64_t lda,
                        int64_t ldb, int64_t ldd, int32_t batch, int64_t stride_a, int64_t stride_b, int64_t stride_d,
                        int64_t stride_bias, int32_t a_kmajor, int32_t b_kmajor, int32_t in_dtype, int32_t out_dtype, int32_t splits,
                        int32_t accumulate, float alpha, const void* a2, const void* b2, int32_t K2, int64_t lda2, int64_t ldb2,
                        int64_t stride_a2, int64_t stride_b2, int device, void* stream);

/* The general form: every field of the batched GEMM plus an INNER batch dimension - problem (t1, t2), t1 < batch, t2 < batch_in,
 * reads a + t1 * stride_a + t2 * stride_a_in, ... (heads inside samples: the attention core of nn.MultiheadAttention as
 * batched GEMMs, scores^T = K_h Q_h^T and O_h = P^T V_h per (sample, head), F.multi_head_attention_forward).  Plain C struct,
 * zero-initialise and fill; batch_in = 0 is read as 1.                                                                      */
typedef struct cosmos_gemm_desc {
  const void* a; const void* b; void* d; const float* bias;
  const void* a2; const void* b2;                 /* optional second operand pair (K2 > 0), same storage orders */
  int32_t M, N, K, K2;
  int64_t lda, ldb, ldd, lda2, ldb2;
  int32_t batch, batch_in;
  int64_t stride_a, stride_b, stride_d, stride_bias, stride_a2, stride_b2;                             /* outer batch */
  int64_t stride_a_in, stride_b_in, stride_d_in, stride_bias_in, stride_a2_in, stride_b2_in;           /* inner batch */
  int32_t a_kmajor, b_kmajor, in_dtype, out_dtype, splits, accumulate;
  float alpha;
  int32_t reserved;
} cosmos_gemm_desc;
int cosmos_gemm_ex(const cosmos_gemm_desc* g, int device, void* stream);

/* Softmax over the KEYS of the folded attention (F.multi_head_attention_forward's softmax(dim=-1) of [queries, keys] scores,
 * stored here keys-major): s fp32 [n_sets][L][n_cols] (row stride lds, set stride s_stride) -> p 16-bit, same indexing with
 * its own strides; p[set][:, c] = softmax over l of s[set][:, c].  zero_key != 0: add_zero_attn (transformer.py:221) - one more
 * key with score 0 and value 0 per column; it only enters the denominator.                                                 */
int cosmos_colsoftmax_fwd(const float* s, int64_t s_stride, int32_t lds, void* p, int64_t p_stride, int32_t ldp, int32_t p_dtype,
                          int32_t n_sets, int32_t L, int32_t n_cols, int32_t zero_key, int device, void* stream);
/* Its backward: ds = p * (dp - sum_l p * dp) per column; dp fp32, p and ds 16-bit (dtype).  (Unchanged by a zero key: its
 * value is zero, so its dp is zero and it drops out of the sum.)                                                           */
int cosmos_colsoftmax_bwd(const void* p, int64_t p_stride, int32_t ldp, const float* dp, int64_t dp_stride, int32_t lddp, void* ds,
                          int64_t ds_stride, int32_t ldds, int32_t dtype, int32_t n_sets, int32_t L, int32_t n_cols, int device,
                          void* stream);

/* LayerNorm over the last dim (eps: the module's own, nn.LayerNorm default 1e-5): y = (x - mean) * rstd * w + b, one row per warp.
 * x: [rows, dim] of x_dtype; y: [rows, dim] bf16/f16 (y_dtype); mean, rstd: fp32 [rows] (saved for backward). */
int cosmos_layernorm_fwd(const void* x, int32_t x_dtype, const float* w, const float* b, void* y, int32_t y_dtype,
                         float* mean, float* rstd, int64_t rows, int32_t dim, float eps, int device, void* stream);
/* dx = LayerNorm backward of dy (dy_dtype) w.r.t. x, written as dx_dtype (added to dx when accumulate != 0);
 * dw, db: fp32 [dim], accumulated with atomics (caller zeroes them).                                          */
int cosmos_layernorm_bwd(const void* dy, int32_t dy_dtype, const void* x, int32_t x_dtype, const float* w,
                         const float* mean, const float* rstd, void* dx, int32_t dx_dtype, int32_t accumulate,
                         float* dw, float* db, int64_t rows, int32_t dim, int device, void* stream);

/* Multi-head attention core with few queries per key/value set.  kv: [n_sets * L, 2 * dim] (keys | values,
 * heads contiguous inside each half), q / o: [n_q, dim]; query c of set s is row s * q_stride_set + c * q_stride_q,
 * c < q_per_set.  Softmax scale 1/sqrt(head_dim).  lse: fp32 [n_q, heads] natural-log sum-exp (saved).          */
int cosmos_attn_core_fwd(const void* q, const void* kv, void* o, float* lse, int32_t dtype, int32_t n_sets, int32_t L,
                         int32_t dim, int32_t heads, int32_t q_per_set, int64_t q_stride_set, int64_t q_stride_q,
                         int device, void* stream);
int cosmos_attn_core_bwd(const void* q, const void* kv, const void* d_o, const float* lse, void* dq, void* dkv,
                         int32_t dtype, int32_t n_sets, int32_t L, int32_t dim, int32_t heads, int32_t q_per_set,
                         int64_t q_stride_set, int64_t q_stride_q, int device, void* stream);

/* out = normalize(f + pooled) row-wise (F.normalize, eps 1e-12): f [rows, dim] f_dtype, pooled fp32, out f_dtype,
 * inv_norm fp32 [rows] saved.  Backward: g_z = (g_out - out * <out, g_out>) * inv_norm, written as fp32 (for the
 * residual branch) and as 16-bit g_dtype (input of the out-projection gradient GEMMs).                          */
int cosmos_addnorm_fwd(const void* f, int32_t f_dtype, const float* pooled, void* out, float* inv_norm, int64_t rows,
                       int32_t dim, int device, void* stream);
int cosmos_addnorm_bwd(const void* g_out, const void* out, int32_t f_dtype, const float* inv_norm, float* g_z32,
                       void* g_z16, int32_t g_dtype, int64_t rows, int32_t dim, int device, void* stream);

/* dst[n] += sum over rows of src[rows, n] (fp32 atomics; bias gradients).                                       */
int cosmos_colsum(const void* src, int32_t dtype, float* dst, int64_t rows, int32_t n, int64_t ld, int device,
                  void* stream);

/* Stored-exponential route (dim 512): the forward keeps, next to its statistics, every 2^(s2 - m) it forms
 * (s2 = logit in log2 units) as bf16 in e_out - cosmos_infonce_e_bytes(p) bytes: one contiguous 32 KB image
 * [4 slabs of 32 rows][16 pieces of 8 columns][32 rows][8] per (pair, 128-row tile, 128-column step) - and the offsets m
 * in off_out (fp32 [gx*gy][ceil(n_cols/32)][n_rows]: the offset row r used in that 32-column chunk; any value works
 * for which the row's significant terms stay inside 2^+-126 - the kernels use a lazily updated per-warp reference).
 * cosmos_infonce_bwd_e then recomputes no logit: dx = G y is its only contraction.  Same outputs and mode scalars
 * as cosmos_infonce_bwd / _bwd_g (the two d(scale) weights must be proportional to the two gradient weights);
 * col_lse2 must be the complete (all ranks) column log-sum-exps; diag_raw is the forward's output of that name: the
 * gradient of a row's positive (softmax terms minus their weights, nearly cancelling for a confident row) is formed
 * from it in fp32, not from the bf16 exponential. */
int64_t cosmos_infonce_e_bytes(const cosmos_infonce_problem* p);
/* workspace of cosmos_infonce_bwd_e: d(scale) partials, plus fp32 partial dX when the column sweep is cut into slices so
 * that a small launch fills whole waves of SM pairs (-1: unsupported problem) */
int64_t cosmos_infonce_bwd_e_workspace_bytes(const cosmos_infonce_problem* p, int device);
int cosmos_infonce_fwd_e(const cosmos_infonce_problem* p, float* row_lse2, float* diag_raw, float* col_lse2, void* e_out,
                         float* off_out, void* workspace, int64_t workspace_bytes, int device, void* stream);
int cosmos_infonce_bwd_e(const cosmos_infonce_problem* p, const void* e, const float* off, const float* diag_raw,
                         const float* row_lse2, const float* col_lse2, float a_row, float a_col, float s_row, float s_col,
                         float weight, const float* upstream, void* dx, float* dscale, void* g_out, int64_t g_ld, void* workspace,
                         int64_t workspace_bytes, int device, void* stream);

/* Column side of the same route (the CLIP term needs gradients on both sides of its logits, src/open_clip/loss.py:206):
 *   dy[s][j][c][:] = sum over row tensors i and the rows r of slice s of  G_ij[r][c] * x_i[r][:]      (fp32, UNIT scale)
 * from the same exponentials - no G tile is written to memory, nothing is transposed (the stored tile image read with
 * its contiguous dimension as M is the MN-major operand G^T).  a_row / a_col as in cosmos_infonce_bwd_e.  dy is
 * [splits][gy][n_cols][512]; the caller sums the slices, all-reduces / reduce-scatters over ranks and multiplies by
 * upstream * scale * weight.  cosmos_infonce_bwd_e_cols_splits proposes the number of slices of the row sweep for which
 * the launch fills whole waves of SM pairs (1 .. 4; -1: unsupported problem).                                            */
int32_t cosmos_infonce_bwd_e_cols_splits(const cosmos_infonce_problem* p, int device);
int cosmos_infonce_bwd_e_cols(const cosmos_infonce_problem* p, const void* e, const float* off, const float* diag_raw,
                              const float* row_lse2, const float* col_lse2, float a_row, float a_col, float* dy, int32_t splits,
                              int device, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Retrieval ranks for the evaluation metrics  (src/training/train.py:766-785 get_clip_metrics and
 * 712-763 compute_retrieval: similarity matrix on the CPU + argsort of every row + position search)
 * ------------------------------------------------------------------------------------------------
 * ranks[r] = the 0-based position of r's best ground-truth item in a STABLE descending sort of row r's fp32
 * dot products with the gallery, in torch.sort's order (NaN above every number, equal scores in index
 * order): the number of items scoring higher plus the tied items with a lower index.  A collapsed model
 * (all scores equal) or NaN features give chance-level ranks, as the reference's argsort does, not 0.  q is [M][D] with row stride ldq, g is [N][D] with row stride ldg (elements; dtype f32,
 * bf16 or f16, products and sums in fp32).  Ground truth in CSR form: items gt_index[gt_offsets[r] ..
 * gt_offsets[r+1]) (device int32 arrays); both null = item r for query r (get_clip_metrics); gt_index
 * null = the contiguous range [gt_offsets[r], gt_offsets[r+1]).  best [M] fp32 / best_col [M] int32 receive
 * score (raw dot product) and gallery index of that ground-truth item, ranks [M] int32 the result.  No similarity matrix is written to memory. */
int cosmos_retrieval_ranks(const void* q, const void* g, int dtype, int32_t M, int32_t N, int32_t D, int64_t ldq, int64_t ldg,
                           const int32_t* gt_offsets, const int32_t* gt_index, float* best, int32_t* best_col, int32_t* ranks,
                           int device, void* stream);

#ifdef __cplusplus
}
#endif
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#endif /* COSMOS_B200_H_ */
