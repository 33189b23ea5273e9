// Declarations shared by the .cu files behind the C ABI (include/cosmos_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cosmos_b200.h"

namespace cb {

typedef cosmos_ema_chunk EmaChunk;

cudaError_t launch_ema(const EmaChunk* table, int n_chunks, double momentum, int dtype, int sm_count, cudaStream_t stream);

struct ClampTable {
  uint64_t ptr[COSMOS_CLAMP_MAX];
  int n;
};
cudaError_t launch_clamp_scalars(const ClampTable& t, double lo, double hi, int dtype, cudaStream_t stream);


// ---- retrieval ranks for the eval metrics (retrieval.cu) ----
cudaError_t launch_retrieval_ranks(const void* q, const void* g, int dtype, int M, int N, int D, long long ldq, long long ldg,
                                   const int* gt_offsets, const int* gt_index, float* best, int* best_col, int* ranks,
                                   cudaStream_t stream);

// ---- generic tcgen05 GEMM (gemm.cu) ----
struct GemmParams {
  int M, N, K, ldd;
  int a_kmajor, b_kmajor, out_dtype, splits;
  uint32_t idesc;
  float alpha;
  const float* bias;
  void* d;
  int batch, accumulate;
  long long sd, sbias;        // batch strides (elements) of D and of the bias
  int K2;                     // contraction length of the optional second operand pair (0: none)
  int batch2;                 // inner batch dimension (problem t = t1 * batch2 + t2)
  long long sd2, sbias2;      // its strides for D and the bias
};
struct GemmArgs {
  const void* a; const void* b; void* d; const float* bias;
  int M, N, K;
  int64_t lda, ldb, ldd;      // row strides (elements) of the stored matrices
  int a_kmajor, b_kmajor;     // 1: stored [rows, K]; 0: stored [K, rows]
  int in_dtype, out_dtype, splits;
  float alpha;
  int batch = 1;              // independent problems of the same shape: the third dimension of the operand tensor maps
  int64_t sa = 0, sb = 0, sd = 0, sbias = 0;    // their strides (elements); any multiple of 8 for a / b, also below the row stride
  int accumulate = 0;         // D += ... instead of D = ... (read-modify-write in the epilogue; splits == 1)
  // optional second operand pair, same storage orders and batch count: D = alpha * (A B^T + A2 B2^T)
  const void* a2 = nullptr; const void* b2 = nullptr;
  int K2 = 0;
  int64_t lda2 = 0, ldb2 = 0, sa2 = 0, sb2 = 0;
  // inner batch dimension: problem (t1, t2), t2 < batch_in, at a + t1 * sa + t2 * sa_in, ... (heads inside samples)
  int batch_in = 1;
  int64_t sa_in = 0, sb_in = 0, sd_in = 0, sbias_in = 0, sa2_in = 0, sb2_in = 0;
};
// returns 0, -1 (CUDA error in *err) or 100000 + CUresult (tensor map)
int launch_gemm(const GemmArgs& a, int sm_count, cudaStream_t stream, cudaError_t* err);


// ---- pooler helpers (xpool.cu) ----
cudaError_t launch_layernorm_fwd(const void* x, int x_dtype, const float* w, const float* b, void* y, int y_dtype, float* mean,
                                 float* rstd, int64_t rows, int dim, float eps, cudaStream_t s);
cudaError_t launch_layernorm_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* w, const float* mean,
                                 const float* rstd, void* dx, int dx_dtype, int accumulate, float* dw, float* db, int64_t rows,
                                 int dim, cudaStream_t s);
cudaError_t launch_attn_core_fwd(const void* q, const void* kv, void* o, float* lse, int dtype, int n_sets, int L, int dim, int heads,
                                 int q_per_set, int64_t qs, int64_t qq, cudaStream_t s);
cudaError_t launch_attn_core_bwd(const void* q, const void* kv, const void* d_o, const float* lse, void* dq, void* dkv, int dtype,
                                 int n_sets, int L, int dim, int heads, int q_per_set, int64_t qs, int64_t qq, cudaStream_t s);
cudaError_t launch_addnorm_fwd(const void* f, int f_dtype, const float* pooled, void* out, float* inv_norm, int64_t rows, int dim,
                               cudaStream_t s);
cudaError_t launch_addnorm_bwd(const void* g_out, const void* out, int f_dtype, const float* inv_norm, float* g_z32, void* g_z16,
                               int g_dtype, int64_t rows, int dim, cudaStream_t s);
cudaError_t launch_colsum(const void* src, int dtype, float* dst, int64_t rows, int n, int64_t ld, cudaStream_t s);
cudaError_t launch_colsoftmax_fwd(const float* S, int64_t s_bs, int lds, void* P, int64_t p_bs, int ldp, int dtype, int n_sets, int L,
                                  int Nc, int zero_key, cudaStream_t s);
cudaError_t launch_colsoftmax_bwd(const void* P, int64_t p_bs, int ldp, const float* dP, int64_t d_bs, int ldd, void* dS, int64_t ds_bs,
                                  int ldds, int dtype, int n_sets, int L, int Nc, cudaStream_t s);

}  // namespace cb
