// Parameter blocks and launchers of the InfoNCE kernels (shared by infonce_*.cu and api.cu).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cb {

constexpr int kFwdBM = 128;  // rows per CTA tile  (= TMEM lanes)
constexpr int kFwdBN = 256;  // columns per forward MMA tile
constexpr int kBwdBN = 128;  // columns per backward step
constexpr int kBwdDP = 256;  // embedding columns of dX one backward CTA accumulates (TMEM budget)

struct FwdParams {
  int gx, gy, n_rows, n_cols, ks, label_offset;
  int n_row_tiles, n_col_tiles, n_slabs;  // n_slabs = n_row_tiles (column partials per 128-row tile)
  uint32_t idesc;
  int dbg;  // diagnostics only (COSMOS_B200_DBG): 1 = skip epilogue math, 2 = skip MMA issue, 1024 = print stall counters
  const float* scale;
  float* row_lse2;
  float* diag_raw;
  float2* col_part;  // [gx*gy][n_slabs][n_cols] (max2, sum)
  // optional (stored-exponential route): every 2^(S2 - m) the statistics loop forms anyway, as bf16, and the offsets m
  uint16_t* e_out;   // [gx*gy][n_row_tiles][n_steps] tiles of [4 slabs of 32 rows][16 column pieces][32 rows][8] bf16 (32 KB each):
                     // 2^(s2[r][c] - off[c / 32][r]), s2 = logit in log2 units
  float* off_out;    // [gx*gy][n_chunks][n_rows]: the row's running maximum (of its 64-column group) when the chunk was processed
  int n_steps;       // ceil(n_cols / 128)
  int n_chunks;      // ceil(n_cols / 32)
};

struct BwdParams {
  int gx, gy, n_rows, n_cols, ks, label_offset;
  int n_row_tiles, n_col_tiles, n_parts;  // n_parts = ceil(dim / 256)
  int dtype;                              // COSMOS_DTYPE_BF16 / _F16
  int dbg;                                // diagnostics only (COSMOS_B200_DBG)
  uint32_t idesc_s, idesc_g;
  float a_row, a_col, s_row, s_col, weight;
  const float* scale;
  const float* upstream;
  const float* row_lse2;  // [gx*gy][n_rows]
  const float* col_lse2;  // [gx*gy][n_cols]
  void* dx;               // [gx][n_rows][dim] stack dtype, may be null
  float* dscale_part;     // [items] partial sums of <dscale-mix, raw logits>, may be null
  int t_splits;           // pair kernel: the column sweep of one row block is split over this many clusters
  float* dx32;            // fp32 [gx][n_rows][dim] accumulation buffer (zeroed), used when t_splits > 1
  void* g_out;            // quad kernel: optional copy of every G tile, [gx * n_rows][g_ld] stack dtype, column j * n_cols + c
  long long g_ld;         // row stride of g_out in elements (multiple of 8)
};

cudaError_t launch_infonce_fwd(const CUtensorMap& tmX, const CUtensorMap& tmY, const FwdParams& p, bool pair, cudaStream_t stream);
cudaError_t launch_infonce_bwd(const CUtensorMap& tmX, const CUtensorMap& tmY, const BwdParams& p, cudaStream_t stream);

cudaError_t launch_infonce_bwd_pair(const CUtensorMap& tmX, const CUtensorMap& tmY64, const CUtensorMap& tmY128, const BwdParams& p,
                                    cudaStream_t stream);

cudaError_t launch_infonce_bwd_quad(const CUtensorMap& tmX, const CUtensorMap& tmY64, const BwdParams& p, cudaStream_t stream);

// Backward from the exponentials the forward stored (infonce_bwd_e2.cu / infonce_bwd_e2t.cu): no logit is recomputed.
struct BwdEParams {
  int gx, gy, n_rows, n_cols, label_offset;
  int n_row_tiles, n_col_tiles;   // 128-row tiles, 128-column steps
  int n_chunks;                   // ceil(n_cols / 32)
  int dtype;                      // dtype of y, dx and g_out (the A operand is converted to it in shared memory)
  int dbg;
  uint32_t idesc_g;
  float a_row, a_col, s_row, s_col, weight;
  const float* scale;
  const float* upstream;
  const void* e;          // the forward's e_out
  const float* off;       // [gx*gy][n_chunks][n_rows]
  const float* row_lse2;  // [gx*gy][n_rows]
  const float* col_lse2;  // [gx*gy][n_cols]
  const float* diag_raw;  // [gx*gy][n_rows]: raw dot product of every row with its positive (the forward's output)
  const void* x;          // [gx][n_rows][512]: only read at the end, for d(scale) = sum_r <x_r, (G y)_r>
  void* dx;               // [gx][n_rows][512]
  float* dscale_part;     // [gx * n_row_tiles] partial sums of <dscale-mix, raw logits>, may be null
  void* g_out;            // optional copy of every G tile, [gx * n_rows][g_ld], column j * n_cols + c
  long long g_ld;
  int t_splits;           // slices of the sweep.  Column-side kernel: of the row sweep, dx = fp32 [t_splits][gy][n_cols][512];
                          // row-side kernel: of the column sweep, partial dX go to dx32 when > 1
  float* dx32;            // row-side kernel, t_splits > 1: fp32 [t_splits][gx][n_rows][512] partial dX (else null)
};
// infonce_bwd_e2.cu: 16 independent scaling warps, E two steps ahead in registers
cudaError_t launch_infonce_bwd_e2(const CUtensorMap& tmY64, const BwdEParams& p, cudaStream_t stream);
// column side of the same route (infonce_bwd_e2t.cu): dY = G^T X from the same exponentials, no G tiles in HBM
cudaError_t launch_infonce_bwd_e2t(const CUtensorMap& tmX64, const BwdEParams& p, cudaStream_t stream);

// infonce_aux.cu
cudaError_t launch_col_combine(const float2* col_part, float* col_lse2, int pairs, int n_slabs, int n_cols, cudaStream_t stream);
cudaError_t launch_loss_sums(const float* row_lse2, const float* diag_raw, const float* col_lse2, const float* scale, int pairs,
                             int n_rows, int n_cols, int label_offset, int use_rows, int use_cols, float* out,
                             cudaStream_t stream);
cudaError_t launch_lse2_merge(const float* parts, float* out, int n_parts, size_t n, cudaStream_t stream);
cudaError_t launch_reduce_dx(const float* parts, int n_parts, void* dst, int dtype, size_t n, cudaStream_t stream);
cudaError_t launch_scale16(const void* src, void* dst, const float* num, float den, int dtype, size_t n, cudaStream_t stream);
cudaError_t launch_convert_dx(const float* src, void* dst, int dtype, size_t n, cudaStream_t stream);
cudaError_t launch_dscale_reduce(const float* part, int n, float weight, const float* upstream, float* dscale,
                                 cudaStream_t stream);

}  // namespace cb
