// InfoNCE backward from STORED exponentials (dim 512): no logit is recomputed.
//
// The forward's statistics loop forms e = 2^(s2 - m) for every logit anyway (s2 = logit in log2 units, m = the row's
// running maximum at that moment); infonce_fwd_kernel can keep them (bf16, [pair][row][column]) together with the
// offsets m ([pair][32-column chunk][row]).  With the final log-sum-exps of rows and columns the softmax gradient is
//     G[r][c] = e * (a_row 2^(m - lse_row[r]) + a_col 2^(m - lse_col[c])) - (a_row + a_col) [c == label r]
// - a per-(row, chunk) factor plus a per-(row, chunk) x per-column product: one FMA and one multiply per element, no
// exponential, no S = X Y^T.  What is left of the backward is ONE GEMM, dX = G Y, so the executed work of a
// distillation pair drops from fwd 2 + bwd 4 to 2 + 2 (x b N D flop) - the algorithmic minimum - and a CTA can own
// all 512 embedding columns of its 128 rows: the fp32 dX accumulator is the whole tensor memory (512 columns), no
// S / G buffers compete for it, so there are no "parts", no 4-CTA clusters and no DSMEM exchange.
//
// One CTA pair (cluster of 2, tcgen05 cta_group::2, M = 256) = 256 rows of one row tensor; per 128-column step and CTA:
//   8 scaling warps: E tile [128 rows x 128 columns] bf16 - stored by the forward as 16 contiguous [128 rows][8 columns]
//        pieces - straight from global memory into registers (loaded one step ahead; a warp load is 512 contiguous bytes;
//        one warp prefetches the tiles into L2 six steps ahead), e -> G (stack dtype, positives subtracted), one store
//        into a 3-stage shared-memory ring in the unswizzled K-major core-matrix layout [piece][row][16 bytes]
//        (A operand: LBO 2048, SBO 128), optional copy of G to HBM for the column-side GEMM of the CLIP term
//   fence.proxy.async, then  dX[:, 0:256] += G Y[:, 0:256],  dX[:, 256:512] += G Y[:, 256:512]  (A and B from shared memory,
//   B = MN-major view of the Y slabs the other kernels read K-major; each CTA supplies 128 of the 256 N columns)
// At the end d(scale) = sum_r <x_r, (G Y)_r> is read off the fp32 accumulators while they are drained.
// HBM traffic: 2 bytes per logit read here + 2 written by the forward (~2.4 TB/s each while the tensor pipe is busy).
// Elements more than 2^-126 below their row's running maximum are zero in E: their row-softmax weight is below fp32
// resolution (the reference's own softmax flushes them too); a column-softmax weight such an element may still carry
// (a column whose every entry is that far below its row's maximum) is dropped - DESIGN.md section 3.
// COSMOS_B200_DBG (diagnostics): 1024 print stall counters (results unchanged), 2048 no scaling math (wrong results).
#include <cstdio>
#include "common.cuh"
#include "infonce.h"
#include "internal.h"

namespace cb {

namespace {

constexpr int kSlabG = 128 * 64 * 2;   // 16 KB: 8 pieces of [128 rows][8 columns] = one 64-column (K) slab of a G tile
constexpr int kStageG = 2 * kSlabG;    // one 128-column step
constexpr int kStagesG = 3;            // G tiles (A operand): being written / waiting / being read
constexpr int kSlabB = 64 * 64 * 2;    // 8 KB: 64 columns (K) x 64 embedding elements
constexpr int kUnitB = 2 * kSlabB;     // this CTA's 128 embedding columns of one N half, for one 64-column half step
constexpr int kUnitsB = 8;             // two steps of B slabs
constexpr int kSmemMisc = 3072;
constexpr int kThreads = 384;          // warps 0-3: roles; warps 4-11: scaling + dX drain
constexpr int kScale = 256;

struct Misc {
  uint64_t g_empty[kStagesG];   // per CTA: tcgen05.commit (multicast) once the step's MMAs have read the G tile of this stage
  uint64_t g_full[kStagesG];    // pair leader: one arrive per scaling warp of both CTAs
  uint64_t b_full[kUnitsB];     // pair leader: TMA bytes of both CTAs
  uint64_t b_empty[kUnitsB];
  uint64_t dx_full;
  uint32_t tmem_slot;
  uint32_t pad[3];
  float red[8];
  alignas(16) float kc[2][128];   // 2^(o - lse_col[c]) of the step's 128 columns (o = lse_col of the step's first column)
  alignas(16) float lc[2][128];   // lse_col itself (slow path)
};
static_assert(sizeof(Misc) <= kSmemMisc, "misc smem");
static_assert(kStagesG * kStageG + kUnitsB * kUnitB + kSmemMisc <= 232448, "shared memory budget");

__device__ __forceinline__ uint32_t bar_red_or(uint32_t id, uint32_t nthreads, bool pred) {
  uint32_t out;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 q, %3, 0;\n\t"
      "bar.red.or.pred p, %1, %2, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(out)
      : "r"(id), "r"(nthreads), "r"(static_cast<uint32_t>(pred))
      : "memory");
  return out;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
infonce_bwd_e_kernel(const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmY64, BwdEParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t r = cluster_ctarank();
  const bool leader = r == 0;

  const int tiles_padded = 2 * ((p.n_row_tiles + 1) / 2);
  const int rt = (blockIdx.x >> 1) * 2 + static_cast<int>(r);
  const int i = rt / tiles_padded;
  const int tr = rt - i * tiles_padded;
  const bool tile_valid = tr < p.n_row_tiles;
  const int n_ct = p.n_col_tiles;
  const int T = p.gy * n_ct;

  uint8_t* sG = smem;
  uint8_t* sB = sG + kStagesG * kStageG;
  Misc* misc = reinterpret_cast<Misc*>(sB + kUnitsB * kUnitB);

  cluster_sync_all();
  if (tid == 0) {
    for (int s = 0; s < kStagesG; ++s) {
      mbar_init(&misc->g_empty[s], 1);
      mbar_init(&misc->g_full[s], 2 * 8);
    }
    for (int u = 0; u < kUnitsB; ++u) {
      mbar_init(&misc->b_full[u], 2);
      mbar_init(&misc->b_empty[u], 1);
    }
    mbar_init(&misc->dx_full, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmE);
    tma_prefetch_desc(&tmY64);
  }
  if (warp == 2) tmem_alloc_pair<512>(&misc->tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = misc->tmem_slot;

  if (warp == 0) {
    // ---------------- L2 prefetch of this CTA's E tiles ----------------
    // E comes from HBM and is read exactly once, by plain loads of the scaling warps (a TMA landing buffer would cost
    // shared-memory bandwidth the tensor core needs: the A and B operands of every MMA are read from shared memory).
    // Those loads run only one step ahead, so the tiles are pulled into L2 p.e_ahead steps early.
    uint32_t s = 0, ph = 0;
    auto piece_of = [&](int t) {
      const int j = t / n_ct, tc = t - j * n_ct;
      return (((i * p.gy + j) * p.n_row_tiles + (tile_valid ? tr : 0)) * n_ct + tc) * 16;
    };
    const int ahead = p.e_ahead;
    auto prefetch_tile = [&](int t) {
      if (p.e_bulk) {          // the 32 KB image of a tile is contiguous: 16 pieces of 2 KB
        bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(p.e) + static_cast<size_t>(piece_of(t)) * 2048, 32768);
      } else {
        tma_prefetch_3d(&tmE, 0, 0, piece_of(t));
        tma_prefetch_3d(&tmE, 0, 0, piece_of(t) + 8);
      }
    };
    if (elect_one())
      for (int t = 0; t < ahead && t < T; ++t) prefetch_tile(t);
    __syncwarp();
    for (int t = 0; ahead > 0 && t + ahead < T; ++t) {
      mbar_wait(&misc->g_empty[s], ph ^ 1);        // paced by the consumption of the G stages
      if (elect_one()) prefetch_tile(t + ahead);
      __syncwarp();
      if (++s == kStagesG) { s = 0; ph ^= 1; }
    }
  } else if (warp == 3) {
    // ---------------- TMA producer: Y slabs (B operand), in the order the MMA warp consumes them ----------------
    uint32_t u = 0, ph = 0;
    for (int t = 0; t < T; ++t) {
      const int j = t / n_ct, tc = t - j * n_ct;
      for (int half = 0; half < 2; ++half) {
        for (int nh = 0; nh < 2; ++nh) {
          mbar_wait(&misc->b_empty[u], ph ^ 1);
          if (elect_one()) {
            for (int sl = 0; sl < 2; ++sl)
              tma_load_3d_pair(sB + u * kUnitB + sl * kSlabB, &tmY64, &misc->b_full[u], (nh * 4 + static_cast<int>(r) * 2 + sl) * 64,
                               tc * 128 + half * 64, j);
            if (leader) mbar_expect_tx(&misc->b_full[u], 2 * kUnitB);
            else mbar_arrive_cluster(&misc->b_full[u], 0);
          }
          __syncwarp();
          if (++u == kUnitsB) { u = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------- MMA issuer (pair leader; whole warp waits, one elected lane issues) ----------------
      uint32_t u = 0, ph = 0, s = 0, phs = 0;
      const bool prof = (p.dbg & 1024) != 0;          // diagnostics: where the issuing warp waits
      long long w_g = 0, w_b = 0;
      const long long t_begin = clock64();
      auto wait_t = [&](uint64_t* bar, uint32_t parity, long long& acc) {
        if (prof) {
          const long long c0 = clock64();
          mbar_wait(bar, parity);
          acc += clock64() - c0;
        } else {
          mbar_wait(bar, parity);
        }
      };
      for (int t = 0; t < T; ++t) {
        wait_t(&misc->g_full[s], phs, w_g);
        fence_proxy_async_all();     // G was written by ordinary stores of both CTAs (each fenced before its arrive)
        tc_fence_after();
        for (int half = 0; half < 2; ++half) {
          for (int nh = 0; nh < 2; ++nh) {
            wait_t(&misc->b_full[u], ph, w_b);
            tc_fence_after();
            const uint32_t a_base = smem_u32(sG + s * kStageG + half * kSlabG);
            const uint32_t b_base = smem_u32(sB + u * kUnitB);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_ss_pair(tmem + nh * 256, make_smem_desc_noswizzle(a_base + kk * 4096, 2048, 128),
                             make_smem_desc(b_base + kk * 2048, kSlabB, 1024), p.idesc_g, (t | half | kk) != 0);
              tc_commit_pair(&misc->b_empty[u], 3);
              if (half == 1 && nh == 1) tc_commit_pair(&misc->g_empty[s], 3);
            }
            __syncwarp();
            if (++u == kUnitsB) { u = 0; ph ^= 1; }
          }
        }
        if (++s == kStagesG) { s = 0; phs ^= 1; }
      }
      if (elect_one()) tc_commit_pair(&misc->dx_full, 3);
      __syncwarp();
      if (prof && lane == 0 && ((blockIdx.x >> 1) % 97) == 5)
        printf("bwd_e prof cluster %d: issue warp total %lld clk, waits g_full %lld b_full %lld (steps %d)\n", blockIdx.x >> 1,
               clock64() - t_begin, w_g, w_b, T);
    }
  } else if (warp >= 4) {
    // ---------------- scaling warps: E -> G in place ----------------
    const uint32_t ts = tid - 128;                 // 0..255
    const int row_t = static_cast<int>(ts & 127);  // row of the tile this thread scales (lanes = consecutive rows)
    const uint32_t hr = ts >> 7;                   // which 32-column chunk of each 64-column slab
    const int grow = tr * 128 + row_t;
    const bool row_valid = tile_valid && grow < p.n_rows;
    const int label = p.label_offset + grow;
    const float scale = __ldg(p.scale);
    const float a_sum = p.a_row + p.a_col;
    const int fmt = p.dtype == COSMOS_DTYPE_BF16 ? 1 : 0;
    const bool want_ds = p.dscale_part != nullptr;
    // statistics of one step: this row's two chunk offsets and log-sum-exp, one column's log-sum-exp (threads 0..127) and
    // the step's reference o = lse_col of its first column; loaded one step ahead of their use
    struct Stats { float off0, off1, lr, lcv, o; };
    auto load_stats = [&](int t) {
      Stats st;
      const int j = t / n_ct, tc = t - j * n_ct;
      const int pair = i * p.gy + j;
      const int chunk0 = tc * 4 + static_cast<int>(hr);
      const float* offp = p.off + static_cast<size_t>(pair) * p.n_chunks * p.n_rows + grow;
      st.off0 = (row_valid && chunk0 < p.n_chunks) ? __ldg(offp + static_cast<size_t>(chunk0) * p.n_rows) : 0.f;
      st.off1 = (row_valid && chunk0 + 2 < p.n_chunks) ? __ldg(offp + static_cast<size_t>(chunk0 + 2) * p.n_rows) : 0.f;
      st.lr = row_valid ? __ldg(p.row_lse2 + static_cast<size_t>(pair) * p.n_rows + grow) : INFINITY;
      const float* lc_ptr = p.col_lse2 + static_cast<size_t>(pair) * p.n_cols;
      const int c = tc * 128 + static_cast<int>(ts);
      st.lcv = (ts < 128 && c < p.n_cols) ? __ldg(lc_ptr + c) : INFINITY;
      st.o = __ldg(lc_ptr + tc * 128);             // tc * 128 < n_cols for every step
      return st;
    };
    Stats nxt = load_stats(0);
    // this thread's 8 pieces (16 bytes = 8 columns of its row) of a step's E tile: slab sl, piece hr * 4 + p4
    const uint4* e_base = reinterpret_cast<const uint4*>(p.e) + row_t;
    auto load_e = [&](int t, uint4 (&dst)[8]) {
      const int j = t / n_ct, tc = t - j * n_ct;
      const size_t piece0 = (static_cast<size_t>((i * p.gy + j) * p.n_row_tiles + tr) * n_ct + tc) * 16;
#pragma unroll
      for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int p4 = 0; p4 < 4; ++p4) {
          const int lp = static_cast<int>(hr) * 4 + p4;
          // pieces the forward never wrote (rows past the batch, columns past the last chunk) must not reach the tensor core
          const bool ok = row_valid && tc * 128 + sl * 64 + lp * 8 < p.n_cols;
          dst[sl * 4 + p4] = ok ? __ldcs(e_base + (piece0 + sl * 8 + lp) * 128) : make_uint4(0u, 0u, 0u, 0u);
        }
    };
    uint4 e_nxt[8];
    load_e(0, e_nxt);

    const bool eprof = (p.dbg & 1024) != 0 && ((blockIdx.x >> 1) % 97) == 5 && lane == 0 && (warp == 4 || warp == 11);
    long long e_wait = 0, e_work = 0, e_pre = 0;
    uint32_t s = 0, phs = 0;
    for (int t = 0; t < T; ++t) {
      const long long c_top = eprof ? clock64() : 0;
      const int j = t / n_ct, tc = t - j * n_ct;
      const int pair = i * p.gy + j;
      const int col_base = tc * 128;
      const uint32_t par = t & 1;
      const Stats st = nxt;
      uint4 e_cur[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) e_cur[k] = e_nxt[k];
      if (t + 1 < T) {
        nxt = load_stats(t + 1);
        load_e(t + 1, e_nxt);
      }
      const float offv[2] = {st.off0, st.off1};
      const float o = st.o, lr = st.lr;
      bool risky = false;
      if (ts < 128) {
        misc->lc[par][ts] = st.lcv;
        misc->kc[par][ts] = ex2(o - st.lcv);       // 0 for the columns past n_cols
        risky = st.lcv != INFINITY && fabsf(o - st.lcv) > 60.f;
      }
      if (row_valid) risky = risky || fabsf(offv[0] - o) > 60.f || fabsf(offv[1] - o) > 60.f;
      // publishes kc / lc to the 256 scaling threads and tells them whether a factor of the product form
      // 2^(off - o) * 2^(o - lse_col) could leave fp32's range in this step (each stays within 2^+-60 otherwise)
      const bool slow = bar_red_or(1, kScale, risky) != 0;

      long long c0 = 0, c1 = 0;
      if (eprof) c0 = clock64();
      mbar_wait(&misc->g_empty[s], phs ^ 1);        // the MMAs that read this stage three steps ago are done
      if (eprof) c1 = clock64();
      const uint32_t stage = smem_u32(sG + s * kStageG);
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        float A1 = 0.f, A2 = 0.f;
        if (row_valid) {
          const float pr = ex2(offv[sl] - lr);                 // <= 1: the running maximum never exceeds the row's log-sum-exp
          const float qc = slow ? 1.f : ex2(offv[sl] - o);
          A1 = p.a_row * pr;
          A2 = p.a_col * qc;
        }
#pragma unroll
        for (int p4 = 0; p4 < 4; ++p4) {
          const uint32_t lp = hr * 4 + p4;                                    // 8-column piece of the slab
          const uint32_t addr = stage + sl * kSlabG + lp * 2048 + row_t * 16;   // 8 lanes = 128 contiguous bytes: no bank conflicts
          const int c0 = col_base + sl * 64 + static_cast<int>(lp) * 8;
          const uint4 w = e_cur[sl * 4 + p4];
          if (!row_valid || c0 >= p.n_cols || (p.dbg & 2048)) {   // 2048: diagnostics, no scaling math (wrong results)
            sts128(addr, w);
            continue;
          }
          const uint32_t wv[4] = {w.x, w.y, w.z, w.w};
          float e[8], kcv[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            e[2 * k] = __uint_as_float(wv[k] << 16);                          // bf16 -> fp32
            e[2 * k + 1] = __uint_as_float(wv[k] & 0xffff0000u);
          }
          const int cl = sl * 64 + static_cast<int>(lp) * 8;                  // column of e[0] inside the step
          if (!slow) {
            const float4 k0 = *reinterpret_cast<const float4*>(&misc->kc[par][cl]);
            const float4 k1 = *reinterpret_cast<const float4*>(&misc->kc[par][cl + 4]);
            kcv[0] = k0.x; kcv[1] = k0.y; kcv[2] = k0.z; kcv[3] = k0.w; kcv[4] = k1.x; kcv[5] = k1.y; kcv[6] = k1.z; kcv[7] = k1.w;
          } else {
            // the exact exponent of every element instead (e <= 1, so e * 2^126 stays finite)
#pragma unroll
            for (int k = 0; k < 8; ++k) kcv[k] = ex2(fminf(offv[sl] - misc->lc[par][cl + k], 126.f));
          }
          const int idx = label - c0;                 // 0..7 when this piece holds the row's positive
          float g[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) g[k] = e[k] * fmaf(A2, kcv[k], A1);
          if (idx >= 0 && idx < 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) g[k] -= (k == idx) ? a_sum : 0.f;       // selects, not an indexed store: g stays in registers
          }
          const uint4 outv = make_uint4(pack2(g[0], g[1], fmt), pack2(g[2], g[3], fmt), pack2(g[4], g[5], fmt), pack2(g[6], g[7], fmt));
          sts128(addr, outv);
          if (p.g_out != nullptr)
            *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.g_out) +
                                      (static_cast<size_t>(i) * p.n_rows + grow) * static_cast<size_t>(p.g_ld) +
                                      static_cast<size_t>(j) * p.n_cols + c0) = outv;
        }
      }
      // ordinary shared-memory stores -> visible to the tensor core's (async proxy) reads
      fence_proxy_async_smem();
      __syncwarp();
      // Each CTA's tensor core reads its OWN rows of A, so the data never crosses CTAs: the proxy fence above plus a plain
      // remote arrive orders it (a release at cluster scope on this arrive cost ~3000 cycles per step in the second CTA)
      if (lane == 0) {
        if (leader) mbar_arrive(&misc->g_full[s]);
        else mbar_arrive_cluster(&misc->g_full[s], 0);
      }
      if (eprof) {
        e_wait += c1 - c0;
        e_work += clock64() - c1;
        e_pre += c0 - c_top;
      }
      if (++s == kStagesG) { s = 0; phs ^= 1; }
    }
    if (eprof)
      printf("bwd_e prof cluster %d cta %u warp %u: scaling warps: stats+barrier %lld, stage wait %lld, scale+publish %lld (steps %d)\n",
             blockIdx.x >> 1, r, warp, e_pre, e_wait, e_work, T);
    const uint32_t ew = warp - 4;
    // drain dX: lanes = rows of the tile (warp % 4 selects the TMEM lane quarter), 256 columns per warp.
    // d(scale) needs no pass of its own: sum_rc G[r][c] <x_r, y_c> = sum_r <x_r, (G y)_r>, the dot product of every row of X
    // with its fp32 accumulator row (the mode scalars weigh d(scale) like G in every mode this kernel accepts).
    mbar_wait(&misc->dx_full, 0);
    tc_fence_after();
    const uint32_t q = warp & 3, h = ew >> 2;
    const int drow = tr * 128 + static_cast<int>(q) * 32 + static_cast<int>(lane);
    const bool drow_valid = tile_valid && drow < p.n_rows;
    const float coef = __ldg(p.upstream) * p.weight * scale;
    const uint16_t* xrow = reinterpret_cast<const uint16_t*>(p.x) + (static_cast<size_t>(i) * p.n_rows + (drow_valid ? drow : 0)) * 512;
    float ds_acc = 0.f;
    for (int c = static_cast<int>(h) * 256; c < static_cast<int>(h) * 256 + 256; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((q * 32u) << 16) + c, v);
      tmem_ld_wait();
      if (drow_valid) {
        if (want_ds) {
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8) {
            const uint4 xv = __ldg(reinterpret_cast<const uint4*>(xrow + c) + k8);
            const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float x0, x1;
              if (fmt) {
                x0 = __uint_as_float(xw[k] << 16);
                x1 = __uint_as_float(xw[k] & 0xffff0000u);
              } else {
                const __half2 hv = *reinterpret_cast<const __half2*>(&xw[k]);
                x0 = __low2float(hv);
                x1 = __high2float(hv);
              }
              ds_acc = fmaf(x0, __uint_as_float(v[k8 * 8 + 2 * k]), ds_acc);
              ds_acc = fmaf(x1, __uint_as_float(v[k8 * 8 + 2 * k + 1]), ds_acc);
            }
          }
        }
        uint32_t ow[16];
#pragma unroll
        for (int kk = 0; kk < 16; ++kk)
          ow[kk] = pack2(__uint_as_float(v[2 * kk]) * coef, __uint_as_float(v[2 * kk + 1]) * coef, fmt);
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.dx) + (static_cast<size_t>(i) * p.n_rows + drow) * 512 + c);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) dst[kk] = make_uint4(ow[4 * kk], ow[4 * kk + 1], ow[4 * kk + 2], ow[4 * kk + 3]);
      }
    }
    if (want_ds) {
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) ds_acc += __shfl_xor_sync(0xffffffffu, ds_acc, sft);
      if (lane == 0) misc->red[ew] = ds_acc;
      named_bar_sync(2, kScale);
      if (ts == 0 && tile_valid) {
        float sum = 0.f;
        for (int w = 0; w < 8; ++w) sum += misc->red[w];
        p.dscale_part[i * p.n_row_tiles + tr] = sum;      // <G, raw dot products> of this tile's rows
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair<512>(tmem);
}

cudaError_t launch_infonce_bwd_e(const CUtensorMap& tmE, const CUtensorMap& tmY64, const BwdEParams& p, cudaStream_t stream) {
  const int smem_bytes = kStagesG * kStageG + kUnitsB * kUnitB + kSmemMisc;
  cudaError_t e = cudaFuncSetAttribute(infonce_bwd_e_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.gx * ((p.n_row_tiles + 1) / 2) * 2);
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
  return cudaLaunchKernelEx(&cfg, infonce_bwd_e_kernel, tmE, tmY64, p);
}

}  // namespace cb
