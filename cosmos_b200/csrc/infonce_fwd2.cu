// InfoNCE forward, second-generation epilogue (CTA pairs only; the TMA / MMA side is the one of infonce_fwd.cu).
//
// What limited the first generation (profiles/README_r02.md): its 16 statistics warps were ISSUE-bound - ~360 instructions per
// thread and 32-column chunk, of which 124 were the 31-shuffle transposing column reduce (lane = row, 32 columns per
// thread), ~30 the per-row running maximum and its warp maximum, 32 the re-basing multiplies for the column sums.  With
// four warps per scheduler that is ~2900 issue slots per 256-column tile against the 4096 cycles its MMAs take: the tensor
// pipe waited for a free accumulator ~30 % of the time, and the 20 extra instructions of the E stores cost 17 %.
// This epilogue needs ~230:
//   * tcgen05.ld.16x256b: a thread holds 4 rows x 8 columns of the chunk (rows g, g+8, g+16, g+24 of the warp's 32, g =
//     lane / 4; columns 8 i + 2 (lane % 4) + {0, 1}) instead of 1 x 32.  Column sums: 24 in-thread adds over the 4 rows,
//     then a transposing reduce over the 8 lanes that share lane % 4 - 7 shuffles instead of 31.  Row sums: 7 in-thread
//     adds per row and two shuffles inside the lane quad.
//   * ONE offset per warp and chunk, updated lazily: e = 2^(s2 - M), M the warp's reference - not a running maximum that
//     every chunk has to re-derive.  M only moves when a row sum outgrows 2^60 (a new maximum more than ~60 above the
//     reference: the chunk is redone from the accumulator, which is still in tensor memory).  No per-chunk max, no rescale,
//     no per-row factor for the column sums: with one offset the row exponentials ARE the column terms.  bf16 and fp32 both
//     hold 2^+-126, so values above the reference are as exact as values below it.
//   * rows whose whole content lies more than 2^-64 below the warp's reference get their own offset M - 64 k (k = 1, 2, ...,
//     exact powers of two; found when a row sum stays below 2^-64, rare: it takes logit spreads > 44 nats inside a 32-row
//     slab, i.e. logit scales >= 30 and rows without any competitive column).  Their column terms are re-based by 2^(-64 k).
//   * a column whose terms all vanish against the warp's reference (sum < 2^-90) redoes the chunk's column statistics with
//     true column maxima (second exponential), as the first generation did.
// Outputs are the first generation's (include/cosmos_b200.h): row log-sum-exp, positives, per-slab column partials, and -
// for the stored-exponential route - e as bf16 tiles with off[pair][chunk][row] the offset each row used.
#include <cstdio>
#include "common.cuh"
#include "infonce.h"

namespace cb {

namespace {

constexpr int BM = kFwdBM, BN = kFwdBN;
constexpr int kSlabX = BM * 64 * 2;    // 16 KB : 128 rows x 64 elements
constexpr int kSmemX = 8 * kSlabX;     // 128 KB
constexpr int kSmemY = 64 * 1024;      // Y ring: 4 x 16 KB (a stage = one 64-element K slab of this CTA's 128 columns)
constexpr int kStages = 4;
constexpr int kStageY = kSmemY / kStages;
constexpr int kSmemE = 32 * 1024;      // E staging: 2 KB per epilogue warp = its 32 rows x 32 columns of a chunk, as stored
constexpr int kSmemMisc = 3072;
constexpr int kEpiWarps = 16;            // 4 TMEM lane quarters x 4 column groups of 64
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr int kEpiThreads = 32 * kEpiWarps;

struct Misc {
  uint64_t x_full;
  uint64_t y_full[kStages];
  uint64_t y_empty[kStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_slot;
  uint32_t pad[5];
};
static_assert(sizeof(Misc) <= kSmemMisc, "misc smem");

// 16 TMEM lanes x 256 bits, four times along the columns: thread t of the warp receives rows t / 4 and t / 4 + 8 (of the 16
// lanes starting at the address' lane), columns 8 i + 2 (t % 4) + {0, 1} for i = 0..3; register 4 i + 2 rr + cc.
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// Transposing reduce over the 8 lanes that share lane % 4 (lane bits 2..4): on entry every lane holds c[0..7]; on return
// the lane whose bits (4, 3, 2) spell k holds op over those 8 lanes of c[k].  7 shuffles.
template <class Op>
__device__ __forceinline__ float reduce8_over_groups(const float (&c)[8], uint32_t lane, Op op) {
  float a[4];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float keep = up ? c[k + 4] : c[k];
      const float send = up ? c[k] : c[k + 4];
      a[k] = op(keep, __shfl_xor_sync(0xffffffffu, send, 16));
    }
  }
  float b[2];
  {
    const bool up = lane & 8;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float keep = up ? a[k + 2] : a[k];
      const float send = up ? a[k] : a[k + 2];
      b[k] = op(keep, __shfl_xor_sync(0xffffffffu, send, 8));
    }
  }
  const bool up = lane & 4;
  const float keep = up ? b[1] : b[0];
  const float send = up ? b[0] : b[1];
  return op(keep, __shfl_xor_sync(0xffffffffu, send, 4));
}

}  // namespace

template <bool kProf>
__global__ void __launch_bounds__(kThreads, 1)
infonce_fwd2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sX = smem;
  uint8_t* sY = smem + kSmemX;
  uint8_t* sE = smem + kSmemX + kSmemY;
  Misc* misc = reinterpret_cast<Misc*>(smem + kSmemX + kSmemY + kSmemE);

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int tiles_padded = 2 * ((p.n_row_tiles + 1) / 2);
  const int per_j = p.gx * tiles_padded;
  const int j = blockIdx.x / per_j;
  const int rem = blockIdx.x - j * per_j;
  const int i = rem / tiles_padded;
  const int tr = rem - i * tiles_padded;   // may be one past the last real tile: fully masked
  const int pair = i * p.gy + j;
  const int ks = p.ks;
  const int n_ct = p.n_col_tiles;

  cluster_sync_all();
  if (tid == 0) {
    mbar_init(&misc->x_full, 2);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&misc->y_full[s], 2);        // one producer arrive per CTA (on the leader's barrier)
      mbar_init(&misc->y_empty[s], 1);       // tcgen05.commit (multicast to both CTAs)
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&misc->acc_full[s], 1);
      mbar_init(&misc->acc_empty[s], 2 * kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 2) tmem_alloc_pair<512>(&misc->tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = misc->tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    auto arm = [&](uint64_t* bar, uint32_t bytes_per_cta) {
      if (leader) mbar_expect_tx(bar, bytes_per_cta * 2);
      else mbar_arrive_cluster(bar, 0);
    };
    if (elect_one()) {
      for (int s = 0; s < ks; ++s) tma_load_3d_pair(sX + s * kSlabX, &tmX, &misc->x_full, s * 64, tr * BM, i);
      arm(&misc->x_full, ks * kSlabX);
    }
    __syncwarp();
    uint32_t stage = 0, phase = 0;
    for (int tc = 0; tc < n_ct; ++tc) {
      for (int s = 0; s < ks; ++s) {
        mbar_wait(&misc->y_empty[stage], phase ^ 1);
        if (elect_one()) {
          tma_load_3d_pair(sY + stage * kStageY, &tmY, &misc->y_full[stage], s * 64, tc * BN + cta_rank * (BN / 2), j);
          arm(&misc->y_full[stage], kStageY);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------- MMA issuer (leader CTA of the pair; whole warp waits, one elected lane issues) ----------------
      mbar_wait(&misc->x_full, 0);
      uint32_t stage = 0, phase = 0;
      for (int tc = 0; tc < n_ct; ++tc) {
        const uint32_t as = tc & 1;
        mbar_wait(&misc->acc_empty[as], ((tc >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem + as * BN;
        for (int s = 0; s < ks; ++s) {
          mbar_wait(&misc->y_full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sX + s * kSlabX);
          const uint32_t b_base = smem_u32(sY + stage * kStageY);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_ss_pair(d_tmem, make_smem_desc(a_base + kk * 32, 0, 1024), make_smem_desc(b_base + kk * 32, 0, 1024), p.idesc,
                           (s | kk) != 0);
            tc_commit_pair(&misc->y_empty[stage], 3);
            if (s == ks - 1) tc_commit_pair(&misc->acc_full[as], 3);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue: row / column softmax statistics, E tiles ----------------
    const uint32_t ew = warp - 4;
    const uint32_t q = warp & 3;          // TMEM lane quarter this warp may access = 32 rows of the tile
    const uint32_t h = ew >> 2;           // 64-column group of the tile
    const uint32_t g = lane >> 2;         // row group: rows g, g + 8, g + 16, g + 24 of the quarter
    const uint32_t ql = lane & 3;         // lane of the quad: columns 8 i + 2 ql + {0, 1}
    const float k2 = __ldg(p.scale) * kLog2e;
    const float NEG_INF = -INFINITY;
    const bool tile_ok = tr < p.n_row_tiles;
    const int row0 = tr * BM + static_cast<int>(q * 32 + g);          // row of R = 0; row of R is row0 + 8 R
    float rowmask[4];                     // 0 for real rows, -inf for rows past the batch (their terms vanish everywhere)
    int label_rel[4];                     // column of the row's positive
#pragma unroll
    for (int R = 0; R < 4; ++R) {
      rowmask[R] = (tile_ok && row0 + 8 * R < p.n_rows) ? 0.f : NEG_INF;
      label_rel[R] = p.label_offset + row0 + 8 * R;
    }
    // the column this lane ends up with after the transposing reduce: idx8 = lane bits (4, 3, 2) = 2 i + cc
    const int cfin = 8 * static_cast<int>(((lane >> 4) & 1) * 2 + ((lane >> 3) & 1)) + 2 * static_cast<int>(ql) +
                     static_cast<int>((lane >> 2) & 1);

    float M = NEG_INF;                    // the warp's reference offset (log2 units), warp-uniform
    float kshift[4] = {0.f, 0.f, 0.f, 0.f};   // row R uses M - 64 kshift[R]; identical in the four lanes of a quad
    bool anyshift = false;                // warp-uniform
    float L[4] = {0.f, 0.f, 0.f, 0.f};    // complete row sums so far, relative to the row's offset
    float diag[4] = {0.f, 0.f, 0.f, 0.f}; // the positive's raw dot product (captured by the lane that owns its column)

    float2* col_part = p.col_part + (static_cast<size_t>(pair) * p.n_slabs + tr * 4 + q) * p.n_cols;
    const bool keep_e = p.e_out != nullptr && tile_ok;
    // E tile images: [pair][row tile][128-column step][4 slabs of 32 rows][16 pieces][32 rows][8 columns] bf16.  A warp's
    // chunk (its 32 rows x 32 columns = 4 pieces) is 2 contiguous KB there: it is assembled in the warp's own 2 KB of shared
    // memory with 4-byte stores and leaves as ONE bulk asynchronous copy - sixteen 4-byte global stores per thread and chunk
    // instead made the load/store queue, which the shuffles share, the bottleneck of the whole epilogue.
    uint8_t* e_slab = keep_e ? reinterpret_cast<uint8_t*>(p.e_out) +
                                   ((static_cast<size_t>(pair) * p.n_row_tiles + tr) * p.n_steps * 4 + q) * 8192
                             : nullptr;                       // + step * 32768 + piece * 512
    const uint32_t e_stage = smem_u32(sE + ew * 2048);
    const uint32_t e_word = e_stage + (g * 4 + ql) * 4;       // + ii * 512 + R * 128
    float* off_rows = keep_e ? p.off_out + static_cast<size_t>(pair) * p.n_chunks * p.n_rows : nullptr;
    const int my_row = row0 + 8 * static_cast<int>(ql);               // the row this lane reports (offsets, final statistics)
    const bool my_row_ok = tile_ok && my_row < p.n_rows;

    // a warp without a single real row (past the batch in the last tile) only reports empty column partials
    const bool warp_has_rows =
        __any_sync(0xffffffffu, rowmask[0] == 0.f || rowmask[1] == 0.f || rowmask[2] == 0.f || rowmask[3] == 0.f) != 0;

    // diagnostics (COSMOS_B200_DBG=1024): where one epilogue warp spends its cycles
    const bool eprof = kProf && (blockIdx.x % 194) == 10 && lane == 0 && (ew == 0 || ew == 7);
    long long pw_wait = 0, pw_ld = 0, pw_exp = 0, pw_store = 0, pw_col = 0;
    const long long p_begin = kProf ? clock64() : 0;

    for (int tc = 0; tc < n_ct; ++tc) {
      const uint32_t as = tc & 1;
      {
        const long long c0 = eprof ? clock64() : 0;
        mbar_wait(&misc->acc_full[as], (tc >> 1) & 1);
        if (eprof) pw_wait += clock64() - c0;
      }
      tc_fence_after();
#pragma unroll 1
      for (int chunk = 0; chunk < 2; ++chunk) {
        const int col0 = tc * BN + static_cast<int>(h) * 64 + chunk * 32;
        if (col0 >= p.n_cols) break;
        if (!warp_has_rows) {
          if (tile_ok && col0 + cfin < p.n_cols) col_part[col0 + cfin] = make_float2(NEG_INF, 0.f);
          continue;
        }
        const bool ragged = col0 + 32 > p.n_cols;            // warp-uniform: the last chunk of a ragged sweep
        const uint32_t taddr = tmem + ((q * 32u) << 16) + as * BN + h * 64 + chunk * 32;
        float t[32];                      // index 16 hh + 4 i + 2 rr + cc: row R = rr + 2 hh, column 8 i + 2 ql + cc
        float srow[4];

        // ---- exponentials against the current offsets; redone (rare) until every row sum is in range ----
#pragma unroll 1
        for (int attempt = 0;; ++attempt) {
          const long long c0 = eprof ? clock64() : 0;
          uint32_t v[32];
          tmem_ld_16x256b_x4(taddr, v);
          tmem_ld_16x256b_x4(taddr + (16u << 16), v + 16);
          tmem_ld_wait();
          const long long c1 = eprof ? clock64() : 0;
          if (eprof) pw_ld += c1 - c0;
          float negm[4];
#pragma unroll
          for (int R = 0; R < 4; ++R) negm[R] = rowmask[R] - (anyshift ? fmaf(-64.f, kshift[R], M) : M);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh)
#pragma unroll
            for (int ii = 0; ii < 4; ++ii)
#pragma unroll
              for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                  const int k = 16 * hh + 4 * ii + 2 * rr + cc;
                  float arg = fmaf(__uint_as_float(v[k]), k2, negm[rr + 2 * hh]);
                  if (ragged && col0 + 8 * ii + 2 * static_cast<int>(ql) + cc >= p.n_cols) arg = NEG_INF;
                  // diagnostics (wrong results): 2048 = no exponential at all, 4096 = every second one (what MUFU costs here)
                  t[k] = ((p.dbg & 2048) || ((p.dbg & 4096) && (k & 1))) ? fminf(fabsf(arg) * 1e-3f, 1.f) : ex2(arg);
                }
#pragma unroll
          for (int R = 0; R < 4; ++R) {
            const int base = 16 * (R >> 1) + 2 * (R & 1);
            float s = (t[base] + t[base + 1]) + (t[base + 4] + t[base + 5]);
            s += (t[base + 8] + t[base + 9]) + (t[base + 12] + t[base + 13]);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            srow[R] = s;
          }
          bool big = false, tiny = false;
#pragma unroll
          for (int R = 0; R < 4; ++R) {
            big = big || !(srow[R] <= 1.152921504606847e18f);                            // 2^60; also catches inf and NaN
            tiny = tiny || (rowmask[R] == 0.f && L[R] + srow[R] < 5.421010862427522e-20f);  // 2^-64 (first chunk: every row)
          }
          const bool again = __any_sync(0xffffffffu, big || tiny) && attempt < 12;
          if (eprof) pw_exp += clock64() - c1;
          if (!again) break;
          // ---- rare: move offsets, then redo the chunk from the accumulator (still in tensor memory) ----
          if (__any_sync(0xffffffffu, big)) {
            // the warp's true maximum of this chunk becomes (at least) the reference
            uint32_t w[32];           // (read again: keeping the first copy alive would cost the common path 32 registers)
            tmem_ld_16x256b_x4(taddr, w);
            tmem_ld_16x256b_x4(taddr + (16u << 16), w + 16);
            tmem_ld_wait();
            float cm = NEG_INF;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const int R = ((k >> 1) & 1) + 2 * (k >> 4);
              float s2 = __uint_as_float(w[k]) * k2 + rowmask[R];
              if (ragged && col0 + 8 * ((k >> 2) & 3) + 2 * static_cast<int>(ql) + (k & 1) >= p.n_cols) s2 = NEG_INF;
              cm = fmaxf(cm, s2);
            }
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, sft));
            if (cm > M) {
              const float down = (M == NEG_INF) ? 0.f : ex2(M - cm);      // the sums so far, re-based to the new reference
#pragma unroll
              for (int R = 0; R < 4; ++R) L[R] *= down;
              M = cm;
            }
            // a row with its own (lower) offset that outgrew it moves one step back towards the reference
#pragma unroll
            for (int R = 0; R < 4; ++R)
              if (!(srow[R] <= 1.152921504606847e18f) && kshift[R] > 0.f) {
                kshift[R] -= 1.f;
                L[R] *= 5.421010862427522e-20f;
              }
          } else {
            // rows whose whole content so far lies 2^-64 below their offset: lower the offset by 64 (exact power of two)
#pragma unroll
            for (int R = 0; R < 4; ++R)
              if (rowmask[R] == 0.f && L[R] + srow[R] < 5.421010862427522e-20f) {
                kshift[R] += 1.f;
                L[R] *= 1.8446744073709552e19f;
              }
          }
          bool sh = false;
#pragma unroll
          for (int R = 0; R < 4; ++R) sh = sh || kshift[R] != 0.f;
          anyshift = __any_sync(0xffffffffu, sh);
        }
#pragma unroll
        for (int R = 0; R < 4; ++R) L[R] += srow[R];

        // ---- positives: the lane that owns the label column keeps the raw dot product ----
        {
          bool mine = false;
#pragma unroll
          for (int R = 0; R < 4; ++R) mine = mine || (rowmask[R] == 0.f && static_cast<uint32_t>(label_rel[R] - col0) < 32u);
          if (__any_sync(0xffffffffu, mine)) {
            uint32_t v[32];
            tmem_ld_16x256b_x4(taddr, v);
            tmem_ld_16x256b_x4(taddr + (16u << 16), v + 16);
            tmem_ld_wait();
#pragma unroll
            for (int R = 0; R < 4; ++R) {
              const int rel = label_rel[R] - col0;
              if (rowmask[R] == 0.f && static_cast<uint32_t>(rel) < 32u && ((rel & 7) >> 1) == static_cast<int>(ql)) {
                const int sel = 4 * (rel >> 3) + (rel & 1);             // 4 i + cc
#pragma unroll
                for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                  for (int cc = 0; cc < 2; ++cc)
                    if (sel == 4 * ii + cc) diag[R] = __uint_as_float(v[16 * (R >> 1) + 4 * ii + 2 * (R & 1) + cc]);
              }
            }
          }
        }

        // ---- E tile words and the offsets they are relative to ----
        const long long c2 = eprof ? clock64() : 0;
        if (keep_e) {
          if (lane == 0) bulk_wait_group_read0();     // the previous chunk's copy has read this warp's buffer
          __syncwarp();
#pragma unroll
          for (int R = 0; R < 4; ++R)
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
              const int k = 16 * (R >> 1) + 4 * ii + 2 * (R & 1);      // rows past the batch hold zeros (their terms vanish)
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(e_word + ii * 512 + R * 128), "r"(pack2(t[k], t[k + 1], 1)) : "memory");
            }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            bulk_store_s2g(e_slab + static_cast<size_t>(col0 >> 7) * 32768 + ((col0 & 127) >> 3) * 512, e_stage, 2048);
            bulk_commit_group();
          }
          if (my_row_ok) {
            float mk = kshift[0];
            if (anyshift) mk = ql == 0 ? kshift[0] : ql == 1 ? kshift[1] : ql == 2 ? kshift[2] : kshift[3];
            off_rows[static_cast<size_t>(col0 >> 5) * p.n_rows + my_row] = anyshift ? fmaf(-64.f, mk, M) : M;
          }
        }

        // ---- column sums over the warp's 32 rows, relative to the reference M ----
        const long long c3 = eprof ? clock64() : 0;
        if (eprof) pw_store += c3 - c2;
        float c8[8];
        if (!anyshift) {
#pragma unroll
          for (int e8 = 0; e8 < 8; ++e8) {
            const int k = 4 * (e8 >> 1) + (e8 & 1);
            c8[e8] = (t[k] + t[k + 2]) + (t[k + 16] + t[k + 18]);
          }
        } else {
          float pw[4];
#pragma unroll
          for (int R = 0; R < 4; ++R) pw[R] = ex2(-64.f * kshift[R]);          // exact; 0 once the row is out of fp32's range
#pragma unroll
          for (int e8 = 0; e8 < 8; ++e8) {
            const int k = 4 * (e8 >> 1) + (e8 & 1);
            c8[e8] = fmaf(t[k], pw[0], fmaf(t[k + 2], pw[1], fmaf(t[k + 16], pw[2], t[k + 18] * pw[3])));
          }
        }
        const float csum = reduce8_over_groups(c8, lane, OpAdd());
        const bool col_ok = col0 + cfin < p.n_cols;
        // Every significant term of a column is a normal fp32 number iff its sum is not tiny relative to 2^M; otherwise
        // the chunk's column statistics are redone with true column maxima (second exponential).
        if (__all_sync(0xffffffffu, !col_ok || csum >= 8.0779e-28f)) {              // 2^-90
          if (tile_ok && col_ok) col_part[col0 + cfin] = make_float2(M, csum);
          if (eprof) pw_col += clock64() - c3;
          continue;
        }
        {
          uint32_t v[32];
          tmem_ld_16x256b_x4(taddr, v);
          tmem_ld_16x256b_x4(taddr + (16u << 16), v + 16);
          tmem_ld_wait();
          float cmx[8];
#pragma unroll
          for (int e8 = 0; e8 < 8; ++e8) {
            const int k = 4 * (e8 >> 1) + (e8 & 1);
            const bool cok = !ragged || col0 + 8 * (e8 >> 1) + 2 * static_cast<int>(ql) + (e8 & 1) < p.n_cols;
            float m = fmaxf(fmaxf(fmaf(__uint_as_float(v[k]), k2, rowmask[0]), fmaf(__uint_as_float(v[k + 2]), k2, rowmask[1])),
                            fmaxf(fmaf(__uint_as_float(v[k + 16]), k2, rowmask[2]), fmaf(__uint_as_float(v[k + 18]), k2, rowmask[3])));
            m = cok ? m : NEG_INF;
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
            cmx[e8] = m;                                     // true maximum of this column over the warp's 32 rows
          }
          float c2[8];
#pragma unroll
          for (int e8 = 0; e8 < 8; ++e8) {
            const int k = 4 * (e8 >> 1) + (e8 & 1);
            const float o = cmx[e8] == NEG_INF ? 0.f : cmx[e8];
            c2[e8] = (ex2(fmaf(__uint_as_float(v[k]), k2, rowmask[0]) - o) + ex2(fmaf(__uint_as_float(v[k + 2]), k2, rowmask[1]) - o)) +
                     (ex2(fmaf(__uint_as_float(v[k + 16]), k2, rowmask[2]) - o) + ex2(fmaf(__uint_as_float(v[k + 18]), k2, rowmask[3]) - o));
            if (cmx[e8] == NEG_INF) c2[e8] = 0.f;
          }
          const float csum2 = reduce8_over_groups(c2, lane, OpAdd());
          const int e8f = static_cast<int>(((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1));
          float cmf = cmx[0];
#pragma unroll
          for (int e8 = 1; e8 < 8; ++e8) cmf = (e8f == e8) ? cmx[e8] : cmf;
          if (tile_ok && col_ok) col_part[col0 + cfin] = make_float2(cmf, csum2);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&misc->acc_empty[as]);
        else mbar_arrive_cluster(&misc->acc_empty[as], 0);
      }
    }

    if (eprof)
      printf("fwd2 prof cta %d warp %u: epilogue total %lld clk: acc_full wait %lld, tmem load %lld, exp + row sums + checks %lld, "
             "E stores %lld, column sums + store %lld (col tiles %d)\n", blockIdx.x, ew, clock64() - p_begin, pw_wait, pw_ld, pw_exp,
             pw_store, pw_col, n_ct);
    if (keep_e && lane == 0) bulk_wait_group0();      // this warp's last E copies are complete before the CTA may retire
    // ---- merge the four column groups of each row; the Y ring is idle now (every MMA has completed) ----
#pragma unroll
    for (int R = 0; R < 4; ++R) {
      diag[R] += __shfl_xor_sync(0xffffffffu, diag[R], 1);
      diag[R] += __shfl_xor_sync(0xffffffffu, diag[R], 2);
    }
    const float myk = ql == 0 ? kshift[0] : ql == 1 ? kshift[1] : ql == 2 ? kshift[2] : kshift[3];
    const float m_run = (M == NEG_INF) ? NEG_INF : fmaf(-64.f, myk, M);
    const float l_run = ql == 0 ? L[0] : ql == 1 ? L[1] : ql == 2 ? L[2] : L[3];
    const float dg_run = ql == 0 ? diag[0] : ql == 1 ? diag[1] : ql == 2 ? diag[2] : diag[3];
    const uint32_t slot = q * 32 + g + 8 * ql;
    float4* exch = reinterpret_cast<float4*>(sY);
    if (h != 0) exch[(h - 1) * 128 + slot] = make_float4(m_run, l_run, dg_run, 0.f);
    named_bar_sync(1, kEpiThreads);
    if (h == 0 && my_row_ok) {
      float m = m_run, dg = dg_run;
      float4 o[3];
#pragma unroll
      for (int gq = 0; gq < 3; ++gq) {
        o[gq] = exch[gq * 128 + slot];
        m = fmaxf(m, o[gq].x);
        dg += o[gq].z;            // the label column lives in exactly one group; the others' diag stayed 0
      }
      float l = 0.f;
      if (m_run != NEG_INF) l += l_run * ex2(m_run - m);
#pragma unroll
      for (int gq = 0; gq < 3; ++gq)
        if (o[gq].x != NEG_INF) l += o[gq].y * ex2(o[gq].x - m);
      const size_t out = static_cast<size_t>(pair) * p.n_rows + my_row;
      p.row_lse2[out] = m + log2f(l);
      p.diag_raw[out] = dg;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair<512>(tmem);
}

cudaError_t launch_infonce_fwd2(const CUtensorMap& tmX, const CUtensorMap& tmY, const FwdParams& p, cudaStream_t stream) {
  const int smem_bytes = kSmemX + kSmemY + kSmemE + kSmemMisc;
  static_assert(kSmemX + kSmemY + kSmemE + kSmemMisc <= 232448, "shared memory budget");
  void (*kern)(CUtensorMap, CUtensorMap, FwdParams) = (p.dbg & 1024) ? infonce_fwd2_kernel<true> : infonce_fwd2_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.gy * p.gx * 2 * ((p.n_row_tiles + 1) / 2));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, tmX, tmY, p);
}

}  // namespace cb
