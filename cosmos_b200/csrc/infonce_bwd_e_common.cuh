// Shared by the two stored-exponential backward kernels (infonce_bwd_e2.cu: dX = G Y over the rows of a tile;
// infonce_bwd_e2t.cu: dY = G^T X over the columns of a tile): shared-memory plan, barriers, and the E -> G scaling helpers.
#pragma once

#include "common.cuh"
#include "infonce.h"
#include "internal.h"

namespace cb {
namespace bwd_e {

constexpr int kSlabG = 128 * 64 * 2;   // 16 KB: 8 pieces of [128 rows][8 columns] = one 64-column (K) slab of a G tile
constexpr int kStageG = 2 * kSlabG;    // one 128-column step
constexpr int kStagesG = 3;            // G tiles (A operand): being written / waiting / being read
constexpr int kSlabB = 64 * 64 * 2;    // 8 KB: 64 columns (K) x 64 embedding elements
constexpr int kUnitB = 2 * kSlabB;     // this CTA's 128 embedding columns of one N half, for one 64-column half step
constexpr int kUnitsB = 8;             // two steps of B slabs
constexpr int kSmemMisc = 3072;
constexpr int kScaleWarps = 16;
constexpr int kScale = 32 * kScaleWarps;
constexpr int kThreads = 128 + kScale;   // warps 0-3: roles; warps 4-19: scaling + dX drain

struct Misc {
  uint64_t g_empty[kStagesG];      // per CTA: tcgen05.commit (multicast) once the step's MMAs have read the G tile of this stage
  uint64_t g_full[kStagesG][2];    // pair leader, per 64-column half: one arrive per scaling warp of that half, both CTAs
  uint64_t b_full[kUnitsB];        // pair leader: TMA bytes of both CTAs
  uint64_t b_empty[kUnitsB];
  uint64_t dx_full;
  uint32_t tmem_slot;
  uint32_t pad[3];
  float red[kScaleWarps];
  alignas(16) float kc[kScaleWarps][32];   // per warp: 2^(o - lse_col[c]) of its chunk's 32 columns (o = lse_col of the first)
};
static_assert(sizeof(Misc) <= kSmemMisc, "misc smem");
static_assert(kStagesG * kStageG + kUnitsB * kUnitB + kSmemMisc <= 232448, "shared memory budget");

// two bf16 products at once (round to nearest even), operands and result as packed pairs
__device__ __forceinline__ uint32_t mul_bf16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
// the factors of two neighbouring columns, f = A2 * kc + A1 in fp32 (ONE FFMA2: both lanes rounded like fmaf), as a bf16 pair
#ifndef COSMOS_BWD_F2
#define COSMOS_BWD_F2 1
#endif
__device__ __forceinline__ uint32_t factor_pair(float kc0, float kc1, float A2, float A1) {
#if COSMOS_BWD_F2
  float f0, f1;
  f2_unpack(f2_fma(f2_pack(kc0, kc1), f2_pack(A2, A2), f2_pack(A1, A1)), f0, f1);
  return pack2(f0, f1, 1);
#else
  return pack2(fmaf(A2, kc0, A1), fmaf(A2, kc1, A1), 1);
#endif
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}


// The gradient of a row's positive: a_row softmax_row + a_col softmax_col - (a_row + a_col).  For a confident row the three
// terms nearly cancel, so it is formed in fp32 from the forward's own dot product of the pair (diag_raw) and the final
// log-sum-exps - not from the stored bf16 exponential, whose 2^-9 rounding would be most of the result.
__device__ __forceinline__ float positive_grad(const BwdEParams& p, int pair, int grow, int label, float k2, float lr) {
  const float s2 = __ldg(p.diag_raw + static_cast<size_t>(pair) * p.n_rows + grow) * k2;
  const float lc = __ldg(p.col_lse2 + static_cast<size_t>(pair) * p.n_cols + label);
  return p.a_row * ex2(s2 - lr) + p.a_col * ex2(s2 - lc) - (p.a_row + p.a_col);
}

// One 16-byte piece (8 columns of one row) outside the common case (fp16 stacks, a factor that may leave fp32's range,
// diagnostics): unpack to fp32, scale, pack.  Every lane of the warp calls it (shuffles inside).
// g_pos: the gradient of the row's positive, formed by the caller in fp32 from the forward's dot product (positive_grad).
__device__ __forceinline__ uint4 scale_piece_generic(uint4 w, int p4, float off, float lcv, float A1, float A2, bool slow, int fmt,
                                                      int label, int c0p, bool row_valid, const float* kc_w, int n_cols, float g_pos,
                                                      int dbg) {
  float kcv[8];
  if (!slow) {
#pragma unroll
    for (int k = 0; k < 8; ++k) kcv[k] = kc_w[p4 * 8 + k];
  } else {
    // the exact exponent of every element (e <= 1, so e * 2^126 stays finite)
#pragma unroll
    for (int k = 0; k < 8; ++k) kcv[k] = ex2(fminf(off - __shfl_sync(0xffffffffu, lcv, p4 * 8 + k), 126.f));
  }
  if (!row_valid || c0p >= n_cols || (dbg & 2048)) return w;   // zeros for rows / columns that do not exist; 2048: no math
  const uint32_t wv[4] = {w.x, w.y, w.z, w.w};
  float g[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    g[2 * k] = __uint_as_float(wv[k] << 16) * fmaf(A2, kcv[2 * k], A1);                  // bf16 -> fp32
    g[2 * k + 1] = __uint_as_float(wv[k] & 0xffff0000u) * fmaf(A2, kcv[2 * k + 1], A1);
  }
  const int idx = label - c0p;                          // 0..7 when this piece holds the row's positive
  if (idx >= 0 && idx < 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = (k == idx) ? g_pos : g[k];   // selects, not an indexed store: g stays in registers
  }
  return make_uint4(pack2(g[0], g[1], fmt), pack2(g[2], g[3], fmt), pack2(g[4], g[5], fmt), pack2(g[6], g[7], fmt));
}


}  // namespace bwd_e
}  // namespace cb
