// InfoNCE forward: flash-style log-sum-exp of S = scale * X Y^T over rows AND columns, for every
// (row tensor i, column tensor j) block of a stack at once.  S never leaves the SM:
//   TMA (128B-swizzled slabs) -> smem -> tcgen05.mma (M128 x N256 x K16, bf16/fp16 in, fp32 out)
//   -> TMEM (2 x 256 columns, double buffered) -> tcgen05.ld -> online softmax statistics.
//
// One CTA = one 128-row tile of (column tensor j, row tensor i): its X tile stays in shared memory
// (<= 8 K-slabs of 16 KB) while 256-column tiles of Y_j stream through a TMA ring.
// kPair = true (the product path): two CTAs on neighbouring SMs form a cluster and run ONE
// tcgen05.mma.cta_group::2 (M = 256) per K step: each CTA holds its own 128 rows of X and only HALF of
// every Y tile (128 columns, 16 KB per stage, 5 stages), which halves the L2->SM operand traffic that
// bounds the single-CTA version.  CTA 0 of the pair issues the MMAs; TMA loads of both CTAs complete on
// its mbarriers; tcgen05.commit multicasts "slot free" / "accumulator ready" to both CTAs.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-19 epilogue
// (warp%4 selects the TMEM lane quarter = 32 rows, warp/4-1 the 64-column group of the tile).
//
// Row statistics are thread-local (one thread = one row): running max / sum in log2 units.
// Column statistics need a reduction over rows: a 31-shuffle warp transpose-reduce per 32x32 block
// gives lane L the (max, sum) of column L over the warp's 32 rows; the four warps that hold the same
// columns of the tile's four 32-row slabs merge them through shared memory (four buffers, an mbarrier each:
// only the chunk's merger waits) and ONE partial per tile goes to the workspace [pair][128-row tile][column];
// col_combine_kernel (infonce_aux.cu) merges the tiles.
//
// Reference semantics: src/open_clip/loss.py:103-142 (get_logits + the two F.cross_entropy calls);
// the positive of local row r is column label_offset + r (loss.py:90-101 with rank offset).
#include <cstdio>
#include "common.cuh"
#include "infonce.h"

namespace cb {

namespace {

constexpr int BM = kFwdBM, BN = kFwdBN;
constexpr int kSlabX = BM * 64 * 2;    // 16 KB : 128 rows x 64 elements
constexpr int kSmemX = 8 * kSlabX;     // 128 KB
constexpr int kSmemY = 80 * 1024;      // Y ring: 2 x 32 KB (single CTA, diagnostics) or 5 x 16 KB (pair)
constexpr int kMaxStages = 5;
constexpr int kSmemMisc = 19456;
constexpr int kEpiWarps = 16;            // 4 TMEM lane quarters x 4 column groups of 64: the statistics loop is latency-bound,
constexpr int kThreads = 128 + 32 * kEpiWarps;   // 4 warps per scheduler hide what 2 could not
constexpr int kEpiThreads = 32 * kEpiWarps;

// E tile images (include/cosmos_b200.h): per 128-column step [4 slabs of 32 rows][16 pieces of 8 columns][32 rows][8] bf16;
// offset (16-byte units) of global piece gp = step * 16 + piece inside a row tile's images, without the slab / row part:
// (gp >> 4) * 2048 + (gp & 15) * 32

struct Misc {
  uint64_t x_full;
  uint64_t y_full[kMaxStages];
  uint64_t y_empty[kMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_slot;
  uint32_t pad[5];
  float bcast[kEpiWarps][32];
  // column partials of a chunk, (max2, sum) per warp and column, four buffers deep: the four warps that hold the four 32-row
  // slabs of the same columns merge them here, so one partial per 128-row tile goes to memory instead of four
  uint64_t xbar[4][4];                   // [column group][buffer]: one arrive per warp of the group
  float2 xch[4][kEpiWarps][32];
};
static_assert(sizeof(Misc) <= kSmemMisc, "misc smem");

}  // namespace

template <bool kPair, bool kProf>
__global__ void __launch_bounds__(kThreads, 1)
infonce_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, FwdParams p) {
  constexpr int kStages = kPair ? 5 : 2;
  constexpr int kLoadCols = kPair ? BN / 2 : BN;                  // Y rows (= S columns) this CTA loads
  constexpr int kStageY = kLoadCols * 64 * 2;                     // bytes of Y this CTA loads per K step
  static_assert(kStages * kStageY <= kSmemY, "Y ring");
  constexpr uint32_t kCtas = kPair ? 2 : 1;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sX = smem;
  uint8_t* sY = smem + kSmemX;
  Misc* misc = reinterpret_cast<Misc*>(smem + kSmemX + kSmemY);

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // work item
  // (pair mode: blockIdx.x = 2 * item + cta rank; the two CTAs take consecutive row tiles)
  const uint32_t cta_rank = kPair ? cluster_ctarank() : 0;
  const bool leader = cta_rank == 0;
  const int tiles_padded = kPair ? 2 * ((p.n_row_tiles + 1) / 2) : p.n_row_tiles;
  const int per_j = p.gx * tiles_padded;
  const int j = blockIdx.x / per_j;
  const int rem = blockIdx.x - j * per_j;
  const int i = rem / tiles_padded;
  const int tr = rem - i * tiles_padded;   // may be one past the last real tile: fully masked
  const int pair = i * p.gy + j;
  const int ks = p.ks;
  const int n_ct = p.n_col_tiles;

  if (kPair) cluster_sync_all();   // both CTAs are resident before any cross-CTA traffic / pair allocation
  if (tid == 0) {
    mbar_init(&misc->x_full, kCtas);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&misc->y_full[s], kCtas);     // one producer arrive per CTA (on the leader's barrier)
      mbar_init(&misc->y_empty[s], 1);        // tcgen05.commit (multicast to both CTAs)
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&misc->acc_full[s], 1);
      mbar_init(&misc->acc_empty[s], (p.dbg & 16) ? kCtas * kEpiThreads : kCtas * kEpiWarps);   // per-warp (or per-thread) arrives of both CTAs
    }
    for (int g = 0; g < 16; ++g) mbar_init(&misc->xbar[g >> 2][g & 3], 4);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 2) {
    if (kPair) tmem_alloc_pair<512>(&misc->tmem_slot);
    else tmem_alloc<512>(&misc->tmem_slot);
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc->tmem_slot;

  // Single-thread roles run with the whole warp converged and an elect.sync predicate around the asynchronous
  // instructions: inside a divergent `lane == 0` branch the compiler wraps every UTMALDG / UTCHMMA / UTCBAR in an
  // elect-and-branch loop, several times the instruction count per MMA.
  if (warp == 0) {
    // ---------------- TMA producer ----------------
    // In pair mode every load is credited to the leader's barrier; the leader arms it with the bytes of
    // BOTH CTAs, the other CTA adds a plain (remote) arrive.
    auto arm = [&](uint64_t* bar, uint32_t bytes_per_cta) {
      if (leader) mbar_expect_tx(bar, bytes_per_cta * kCtas);
      else mbar_arrive_cluster(bar, 0);
    };
    auto load = [&](void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
      if (kPair) tma_load_3d_pair(dst, m, bar, c0, c1, c2);
      else tma_load_3d(dst, m, bar, c0, c1, c2);
    };
    if (elect_one()) {
      for (int s = 0; s < ks; ++s) load(sX + s * kSlabX, &tmX, &misc->x_full, s * 64, tr * BM, i);
      arm(&misc->x_full, ks * kSlabX);
    }
    __syncwarp();
    uint32_t stage = 0, phase = 0;
    for (int tc = 0; tc < n_ct; ++tc) {
      for (int s = 0; s < ks; ++s) {
        mbar_wait(&misc->y_empty[stage], phase ^ 1);
        if (elect_one()) {
          load(sY + stage * kStageY, &tmY, &misc->y_full[stage], s * 64, tc * BN + cta_rank * kLoadCols, j);
          arm(&misc->y_full[stage], kStageY);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------- MMA issuer (leader CTA of the pair only; whole warp waits, one elected lane issues) ----------------
      constexpr bool prof = kProf;                    // diagnostics (COSMOS_B200_DBG=1024): where the issuing warp waits
      long long w_acc = 0, w_y = 0;
      const long long t_begin = kProf ? clock64() : 0;
      auto wait_t = [&](uint64_t* bar, uint32_t ph, long long& acc) {
        if (prof) {
          const long long c0 = clock64();
          mbar_wait(bar, ph);
          acc += clock64() - c0;
        } else {
          mbar_wait(bar, ph);
        }
      };
      mbar_wait(&misc->x_full, 0);
      uint32_t stage = 0, phase = 0;
      for (int tc = 0; tc < n_ct; ++tc) {
        const uint32_t as = tc & 1;
        wait_t(&misc->acc_empty[as], ((tc >> 1) & 1) ^ 1, w_acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem + as * BN;
        for (int s = 0; s < ks; ++s) {
          wait_t(&misc->y_full[stage], phase, w_y);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sX + s * kSlabX);
          const uint32_t b_base = smem_u32(sY + stage * kStageY);
          if (elect_one()) {
            if (!(p.dbg & 2)) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t da = make_smem_desc(a_base + kk * 32, 0, 1024), db = make_smem_desc(b_base + kk * 32, 0, 1024);
                if (kPair) umma_ss_pair(d_tmem, da, db, p.idesc, (s | kk) != 0);
                else umma_ss(d_tmem, da, db, p.idesc, (s | kk) != 0);
              }
            }
            if (kPair) tc_commit_pair(&misc->y_empty[stage], 3); else tc_commit(&misc->y_empty[stage]);
            if (s == ks - 1) {
              if (kPair) tc_commit_pair(&misc->acc_full[as], 3); else tc_commit(&misc->acc_full[as]);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      if (prof && lane == 0 && (blockIdx.x % 194) == 10)
        printf("fwd prof cta %d: issue warp total %lld clk, waits acc_empty %lld y_full %lld (col tiles %d)\n", blockIdx.x,
               clock64() - t_begin, w_acc, w_y, n_ct);
    }
  } else if (warp >= 4) {
    // ---------------- epilogue: online row / column softmax statistics ----------------
    const uint32_t ew = warp - 4;
    const uint32_t q = warp & 3;   // TMEM lane quarter this warp may access
    const uint32_t h = ew >> 2;    // 64-column group of the tile
    const int row = tr * BM + q * 32 + lane;
    const bool row_valid = row < p.n_rows;
    const int label = p.label_offset + row;
    const float k2 = __ldg(p.scale) * kLog2e;
    const float NEG_INF = -INFINITY;
    const bool fast_ok = k2 > 0.f && !(p.dbg & 32);

    float m_run = NEG_INF, l_run = 0.f, diag = 0.f;
    float2* col_part = p.col_part + (static_cast<size_t>(pair) * p.n_slabs + tr) * p.n_cols;       // n_slabs = row tiles
    // Publish this warp's (max2, sum) of the chunk's 32 columns.  The four warps of a column group take turns as the chunk's
    // merger: everybody arrives on the buffer's mbarrier (non-blocking), only the merger waits - a CTA-style barrier per chunk
    // made the forward with E stores 2 % slower (a warp held up by its stores held up three others).  Four buffers: before a
    // warp overwrites buffer b (chunk c + 4) it has been the merger of a chunk after c, for which every warp had already
    // written - i.e. had finished reading chunk c.  Fixed merge order: deterministic.
    uint32_t cc = 0;                // chunks done, the same sequence in the four warps of a group
    auto publish = [&](int col0, float pm, float ps) {
      const uint32_t bsel = cc & 3;
      misc->xch[bsel][ew][lane] = make_float2(pm, ps);
      __syncwarp();
      if (lane == 0) mbar_arrive(&misc->xbar[h][bsel]);
      if (q == bsel) {
        mbar_wait(&misc->xbar[h][bsel], (cc >> 2) & 1);
        float2 a[4];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) a[qq] = misc->xch[bsel][h * 4 + qq][lane];
        const float m = fmaxf(fmaxf(a[0].x, a[1].x), fmaxf(a[2].x, a[3].x));
        float sum = 0.f;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq)
          if (a[qq].x != NEG_INF) sum = fmaf(a[qq].y, ex2(a[qq].x - m), sum);
        if (tr < p.n_row_tiles && col0 + static_cast<int>(lane) < p.n_cols) col_part[col0 + lane] = make_float2(m, sum);
      }
      ++cc;
    };
    // stored-exponential route (infonce_bwd_e2.cu, infonce_bwd_e2t.cu): this row's 2^(s2 - m_run) of every chunk as bf16, and m_run itself
    // Layout of e_out: one contiguous 32 KB image per (pair, 128-row tile, 128-column step): [4 slabs of 32 rows][16 pieces of 8
    // columns][32 rows][8 elements] - a warp's store of one piece is 512 contiguous bytes, its four pieces of a chunk 2 KB.
    const bool keep_e = p.e_out != nullptr && row_valid && tr < p.n_row_tiles;
    uint4* e_tile = keep_e ? reinterpret_cast<uint4*>(p.e_out) +
                                 (static_cast<size_t>(pair) * p.n_row_tiles + tr) * p.n_steps * 2048 + (q * 512 + lane)
                           : nullptr;       // + (step * 16 + piece) * 128 (16-byte units)
    float* off_row = keep_e ? p.off_out + static_cast<size_t>(pair) * p.n_chunks * p.n_rows + row : nullptr;

    const bool eprof = kProf && (blockIdx.x % 194) == 10 && lane == 0 && (ew == 0 || ew == 7);
    long long e_wait = 0;
    const long long e_begin = kProf ? clock64() : 0;
    for (int tc = 0; tc < n_ct; ++tc) {
      const uint32_t as = tc & 1;
      if (eprof) {
        const long long c0 = clock64();
        mbar_wait(&misc->acc_full[as], (tc >> 1) & 1);
        e_wait += clock64() - c0;
      } else {
        mbar_wait(&misc->acc_full[as], (tc >> 1) & 1);
      }
      tc_fence_after();
      // (Measured, no effect: starting the four column groups of a scheduler 400 / 800 cycles apart at the first tile, so that
      // they would not queue for the same pipe phase by phase - forward launch 17.90 / 17.90 / 17.90 ms.  profiles/stagger_r02.txt)
#pragma unroll 1   // measured: unrolling by 2 lowers throughput (register pressure in the 8 epilogue warps)
      for (int chunk = 0; chunk < 2; ++chunk) {
        const int col0 = tc * BN + h * 64 + chunk * 32;
        if (col0 >= p.n_cols) break;
        if ((p.dbg & 1) && chunk > 0) break;
        // the chunk's four 16-byte pieces of this row lie 512 bytes apart (col0 is a multiple of 32: one address per chunk)
        uint4* e_chunk = e_tile + static_cast<size_t>(col0 >> 7) * 2048 + ((col0 >> 5) & 3) * 128;
        const uint32_t t_chunk = tmem + ((q * 32u) << 16) + as * BN + h * 64 + chunk * 32;
        uint32_t v[32];
        tmem_ld32(t_chunk, v);
        tmem_ld_wait();
        if (row_valid && label >= col0 && label < col0 + 32) {
          const int idx = label - col0;
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (k == idx) diag = __uint_as_float(v[k]);
        }
        float t[32];
        bool need_exact = !fast_ok || (col0 + 32 > p.n_cols);
        if (!need_exact) {
          // ---- one exponential per logit: rows use their own running max m_r as offset; the column sums reuse the
          // same exponentials, re-based to the warp's largest running max M_w by one multiply with f_r = 2^(m_r - M_w).
          // The epilogue is bound by the LATENCY of its dependent chains, not by issue slots (ncu, round 2: issue 56 - 65 %
          // busy at 70 % tensor pipe), so: the row maximum is a depth-4 tree of three-input maxima, the warp maximum is ONE
          // instruction (CREDUX.MAX.F32), and the exponentials - which need only m_r - are issued before M_w is consumed.
          // Packed pairs (FFMA2 / FADD2 / FMUL2) cut 80 of the ~420 instructions per chunk on top.
          float m1[11];
#pragma unroll
          for (int k = 0; k < 10; ++k)
            m1[k] = max3(__uint_as_float(v[3 * k]), __uint_as_float(v[3 * k + 1]), __uint_as_float(v[3 * k + 2]));
          m1[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
          const float m2a = max3(m1[0], m1[1], m1[2]), m2b = max3(m1[3], m1[4], m1[5]), m2c = max3(m1[6], m1[7], m1[8]);
          const float cm = fmaxf(max3(m2a, m2b, m2c), fmaxf(m1[9], m1[10])) * k2;          // k2 > 0 on this path
          const float m_before = m_run, l_before = l_run;
          if (cm > m_run) {
            l_run *= ex2(m_run - cm);
            m_run = cm;
          }
          const float mw = warp_max_f32(row_valid ? m_run : NEG_INF);
          // rows past the batch hold zero logits (TMA zero fill): finite everywhere, their results are dropped below
          const float neg_m = -m_run;
          const uint64_t k2p = f2_pack(k2, k2), nmp = f2_pack(neg_m, neg_m);
          uint64_t ss2 = f2_pack(0.f, 0.f);
          auto pair_of = [&](int k) {                  // t[k], t[k + 1] <- 2^(s2 - m_r), and their sum
            float a0, a1;
            f2_unpack(f2_fma(f2_pack_bits(v[k], v[k + 1]), k2p, nmp), a0, a1);
            t[k] = ex2(a0);
            t[k + 1] = ex2(a1);
            ss2 = f2_add(ss2, f2_pack(t[k], t[k + 1]));
          };
          if (keep_e) {
            // 16-byte pieces leave as soon as their 8 exponentials exist
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) {
              uint32_t pk[4];
#pragma unroll
              for (int k2i = 0; k2i < 4; ++k2i) {
                const int k = k8 * 8 + k2i * 2;
                pair_of(k);
                pk[k2i] = pack2(t[k], t[k + 1], 1);
              }
              e_chunk[k8 * 32] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            off_row[static_cast<size_t>(col0 >> 5) * p.n_rows] = m_run;
          } else {
#pragma unroll
            for (int k = 0; k < 32; k += 2) pair_of(k);
          }
          float ss_lo, ss_hi;
          f2_unpack(ss2, ss_lo, ss_hi);
          l_run += ss_lo + ss_hi;
          if (mw == NEG_INF) {                             // warp-uniform: no row of this warp is real
            publish(col0, NEG_INF, 0.f);
            continue;
          }
          const float f = row_valid ? ex2(m_run - mw) : 0.f;
          const uint64_t fp = f2_pack(f, f);
#pragma unroll
          for (int k = 0; k < 32; k += 2) f2_unpack(f2_mul(f2_pack(t[k], t[k + 1]), fp), t[k], t[k + 1]);
          // (Measured and rejected: requesting the NEXT chunk's accumulator - tcgen05.ld into the registers that are free after
          // two rounds of this reduce - so that the tensor-memory read overlaps the rest: the 32 registers stay live across the
          // loop edge, ptxas spills ~20 values, forward with E stores 16.1 -> 20.3 ms.  profiles/fwd_abc_r02.txt)
          const float csum = warp_transpose_sum(t, lane);
          // Every significant term of a column is a normal fp32 number iff the column sum is not tiny relative
          // to 2^M_w (DESIGN.md "one-exp statistics"); otherwise redo this block with true column maxima.
          if (__all_sync(0xffffffffu, csum >= 8.0779e-28f)) {   // 2^-90
            publish(col0, mw, csum);
            continue;
          }
          need_exact = true;
          // undo this block's row update: the exact path below redoes it from scratch - from the accumulator, which is still
          // in tensor memory: reading it again (rare) instead of keeping its 32 registers alive across the column reduce
          // above takes the spills out of the common path (ncu, round 2: ~10 local-memory accesses per chunk, and the
          // instructions behind them held ~15 % of the kernel's stall samples)
          m_run = m_before;
          l_run = l_before;
          tmem_ld32(t_chunk, v);
          tmem_ld_wait();
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) t[k] = __uint_as_float(v[k]) * k2;
        if (col0 + 32 > p.n_cols) {
#pragma unroll
          for (int k = 0; k < 32; ++k) t[k] = (col0 + k < p.n_cols) ? t[k] : NEG_INF;
        }
        // rows: this thread's row, running (max, sum)
        float cm = t[0];
#pragma unroll
        for (int k = 1; k < 32; ++k) cm = fmaxf(cm, t[k]);
        if (cm > m_run) {
          l_run *= ex2(m_run - cm);
          m_run = cm;
        }
        if (m_run != NEG_INF) {
          float s = 0.f;
          if (keep_e) {
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) {
              uint32_t pk[4];
#pragma unroll
              for (int k2i = 0; k2i < 4; ++k2i) {
                const int k = k8 * 8 + k2i * 2;
                const float e0 = ex2(t[k] - m_run), e1 = ex2(t[k + 1] - m_run);    // columns past n_cols: 2^-inf = 0
                s += e0;
                s += e1;
                pk[k2i] = pack2(e0, e1, 1);
              }
              e_chunk[k8 * 32] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            off_row[static_cast<size_t>(col0 >> 5) * p.n_rows] = m_run;
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) s += ex2(t[k] - m_run);
          }
          l_run += s;
        } else if (keep_e) {
          // every logit of this row so far is -inf (scale * x.y = -inf cannot happen with finite inputs; kept for safety)
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8) e_chunk[k8 * 32] = make_uint4(0u, 0u, 0u, 0u);
          off_row[static_cast<size_t>(col0 >> 5) * p.n_rows] = 0.f;
        }
        // columns: reduce over the warp's 32 rows
        if (!row_valid) {
#pragma unroll
          for (int k = 0; k < 32; ++k) t[k] = NEG_INF;
        }
        const float cmx = warp_transpose_reduce(t, lane, OpMax());
        misc->bcast[ew][lane] = cmx;
        __syncwarp();
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 o = *reinterpret_cast<const float4*>(&misc->bcast[ew][k4 * 4]);
          t[k4 * 4 + 0] = ex2(t[k4 * 4 + 0] - (o.x == NEG_INF ? 0.f : o.x));
          t[k4 * 4 + 1] = ex2(t[k4 * 4 + 1] - (o.y == NEG_INF ? 0.f : o.y));
          t[k4 * 4 + 2] = ex2(t[k4 * 4 + 2] - (o.z == NEG_INF ? 0.f : o.z));
          t[k4 * 4 + 3] = ex2(t[k4 * 4 + 3] - (o.w == NEG_INF ? 0.f : o.w));
        }
        __syncwarp();
        const float csum = warp_transpose_reduce(t, lane, OpAdd());
        publish(col0, cmx, csum);
      }
      tc_fence_before();
      if (p.dbg & 16) {
        if (leader) mbar_arrive(&misc->acc_empty[as]);
        else mbar_arrive_cluster(&misc->acc_empty[as], 0);
      } else {
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&misc->acc_empty[as]);
          else mbar_arrive_cluster(&misc->acc_empty[as], 0);
        }
      }
    }

    if (eprof)
      printf("fwd prof cta %d warp %u: epilogue total %lld clk, acc_full wait %lld\n", blockIdx.x, ew, clock64() - e_begin, e_wait);
    // merge the four column groups of each row; the Y ring is idle now (every MMA has completed)
    float4* exch = reinterpret_cast<float4*>(sY);
    if (h != 0) exch[(h - 1) * 128 + q * 32 + lane] = make_float4(m_run, l_run, diag, 0.f);
    named_bar_sync(1, kEpiThreads);
    if (h == 0 && row_valid) {
      float m = m_run, dg = diag;
      float4 o[3];
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        o[g] = exch[g * 128 + q * 32 + lane];
        m = fmaxf(m, o[g].x);
        dg += o[g].z;            // the label column lives in exactly one group; the others' diag stayed 0
      }
      float l = 0.f;
      if (m_run != NEG_INF) l += l_run * ex2(m_run - m);
#pragma unroll
      for (int g = 0; g < 3; ++g)
        if (o[g].x != NEG_INF) l += o[g].y * ex2(o[g].x - m);
      const size_t out = static_cast<size_t>(pair) * p.n_rows + row;
      p.row_lse2[out] = m + log2f(l);
      p.diag_raw[out] = dg;
    }
  }

  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    if (kPair) tmem_dealloc_pair<512>(tmem);
    else tmem_dealloc<512>(tmem);
  }
}

cudaError_t launch_infonce_fwd(const CUtensorMap& tmX, const CUtensorMap& tmY, const FwdParams& p, bool pair, cudaStream_t stream) {
  const int smem_bytes = kSmemX + kSmemY + kSmemMisc;
  static_assert(kSmemX + kSmemY + kSmemMisc <= 232448, "shared memory budget");
  const bool prof = (p.dbg & 1024) != 0;
  void (*kern)(CUtensorMap, CUtensorMap, FwdParams) =
      pair ? (prof ? infonce_fwd_kernel<true, true> : infonce_fwd_kernel<true, false>)
           : (prof ? infonce_fwd_kernel<false, true> : infonce_fwd_kernel<false, false>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = pair ? dim3(p.gy * p.gx * 2 * ((p.n_row_tiles + 1) / 2)) : dim3(p.gy * p.gx * p.n_row_tiles);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, tmX, tmY, p);
}

}  // namespace cb
