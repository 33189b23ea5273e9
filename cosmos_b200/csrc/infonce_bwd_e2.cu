// InfoNCE backward from STORED exponentials (dim 512): 16 independent scaling warps, dX = G Y.
//
// The forward (infonce_fwd.cu, cosmos_infonce_fwd_e) kept, for every logit s2 (log2 units) of a row block,
//     e = 2^(s2 - m)  (bf16; one contiguous 32 KB image per (pair, 128-row tile, 128-column step):
//                      [4 slabs of 32 rows][16 pieces of 8 columns][32 rows][8])   and
//     m              (fp32, [pair][32-column chunk][row]: the offset the chunk's exponentials are relative to).
// With the final log-sum-exps the gradient of a logit is
//     G = e * (a_row 2^(m - lse_row[r]) + a_col 2^(m - lse_col[c])) - (a_row + a_col)[c == label r]
// - no exponential per element, no X Y^T - and the only contraction left is dX = G Y, accumulated for 128 rows x 512
// embedding columns in the whole tensor memory of the SM (CTA pair, tcgen05 cta_group::2, M = 256).  The positive's own
// gradient is formed in fp32 from the forward's diag_raw (positive_grad), not from its bf16 exponential, and
// d(scale) = sum_r <x_r, (G Y)_r> is read off the fp32 accumulators at the end.
//
// What the in-kernel counters of the first generation of this kernel (8 scaling warps, a CTA-wide bar.red per step, L2
// prefetch role; removed) showed (profiles/README_r02.md): its 8 scaling warps needed ~4000
// cycles per 128-column step - the tensor core needs 2048 - because every step was one serial chain per warp: wait for
// the E registers loaded a step earlier (L2 / HBM latency under load exceeded a step), a 256-thread bar.red that publishes
// the step's column factors, the dependent unpack - FMA - pack - store chain of 64 elements with two warps per scheduler,
// the arrive.  The L2 prefetches (tensor-map or bulk) bought nothing: with none at all the kernel was fastest.  Here:
//   * 16 scaling warps (640-thread CTAs, <= 102 registers): thread = (row, 32-column chunk), 4 pieces of 16 bytes per step,
//     four warps per scheduler to hide the chains;
//   * E is loaded into registers TWO steps ahead (16-byte loads, a warp load is 512 contiguous bytes, read exactly once,
//     ld.global.cs), no L2 prefetch role;
//   * no CTA-wide barrier in the loop: a 32-column chunk has ONE row offset, and each warp forms the 32 column factors of
//     its own chunk itself (lane = column, one ex2 per lane and step) and hands them to its lanes through 128 bytes of its
//     own shared memory; the "a factor may leave fp32's range" decision is per warp (__any_sync);
//   * the two 64-column halves of a step have their own full barriers, so the MMAs of the first half start while the
//     second is still being scaled.
// COSMOS_B200_DBG (diagnostics): 1024 print stall counters (results unchanged), 2048 no scaling math (wrong results).
#include <cstdio>
#include "infonce_bwd_e_common.cuh"

namespace cb {

using namespace bwd_e;

template <bool kBf16, bool kProf>
__global__ void __launch_bounds__(kThreads, 1)
infonce_bwd_e2_kernel(const __grid_constant__ CUtensorMap tmY64, BwdEParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t r = cluster_ctarank();
  const bool leader = r == 0;

  // work item: (slice of the column sweep, row tensor i, row tile tr); the two CTAs of a pair take consecutive row tiles.
  // t_splits > 1 cuts the sweep over the gy * n_col_tiles steps into slices so that small launches (a rank of an 8-GPU job
  // has 256 pairs = 3.46 waves of 74) fill whole waves; the slices' fp32 partial dX are summed by reduce_dx_kernel.
  const int tiles_padded = 2 * ((p.n_row_tiles + 1) / 2);
  const int per_split = p.gx * tiles_padded;
  const int item = (blockIdx.x >> 1) * 2 + static_cast<int>(r);
  const int split = item / per_split;
  const int rt = item - split * per_split;
  const int i = rt / tiles_padded;
  const int tr = rt - i * tiles_padded;
  const bool tile_valid = tr < p.n_row_tiles;
  const int n_ct = p.n_col_tiles;
  const int T_all = p.gy * n_ct;
  const int t_first = static_cast<int>(static_cast<long long>(T_all) * split / p.t_splits);
  const int T = static_cast<int>(static_cast<long long>(T_all) * (split + 1) / p.t_splits) - t_first;
  const int j_first = t_first / n_ct, tc_first = t_first - j_first * n_ct;

  uint8_t* sG = smem;
  uint8_t* sB = sG + kStagesG * kStageG;
  Misc* misc = reinterpret_cast<Misc*>(sB + kUnitsB * kUnitB);

  cluster_sync_all();
  if (tid == 0) {
    for (int s = 0; s < kStagesG; ++s) {
      mbar_init(&misc->g_empty[s], 1);
      mbar_init(&misc->g_full[s][0], 2 * (kScaleWarps / 2));
      mbar_init(&misc->g_full[s][1], 2 * (kScaleWarps / 2));
    }
    for (int u = 0; u < kUnitsB; ++u) {
      mbar_init(&misc->b_full[u], 2);
      mbar_init(&misc->b_empty[u], 1);
    }
    mbar_init(&misc->dx_full, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmY64);
  if (warp == 2) tmem_alloc_pair<512>(&misc->tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = misc->tmem_slot;

  if (warp == 3) {
    // ---------------- TMA producer: Y slabs (B operand), in the order the MMA warp consumes them ----------------
    uint32_t u = 0, ph = 0;
    int j = j_first, tc = tc_first;
    for (int t = 0; t < T; ++t) {
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
      if (++tc == n_ct) { tc = 0; ++j; }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------- MMA issuer (pair leader; whole warp waits, one elected lane issues) ----------------
      uint32_t u = 0, ph = 0, s = 0, phs = 0;
      constexpr bool prof = kProf;                    // diagnostics (COSMOS_B200_DBG=1024): where the issuing warp waits
      long long w_g = 0, w_b = 0, w_f = 0, w_i = 0;
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
        for (int half = 0; half < 2; ++half) {
          wait_t(&misc->g_full[s][half], phs, w_g);
          {
            const long long f0 = prof ? clock64() : 0;
            // G was written by ordinary stores of both CTAs, each fenced (fence.proxy.async.shared::cta) before its arrive.
            // The unqualified fence.proxy.async that stood here in the first generation cost the issuing warp ~2200 cycles per
            // step (1.13 M of its 2.4 M cycles): 17.7 instead of 20.7 ms for 64 pairs at b = N = 16384.
            fence_proxy_async_smem();
            if (prof) w_f += clock64() - f0;
          }
          tc_fence_after();
          for (int nh = 0; nh < 2; ++nh) {
            wait_t(&misc->b_full[u], ph, w_b);
            tc_fence_after();
            const uint32_t a_base = smem_u32(sG + s * kStageG + half * kSlabG);
            const uint32_t b_base = smem_u32(sB + u * kUnitB);
            const long long i0 = prof ? clock64() : 0;
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_ss_pair(tmem + nh * 256, make_smem_desc_noswizzle(a_base + kk * 4096, 2048, 128),
                             make_smem_desc(b_base + kk * 2048, kSlabB, 1024), p.idesc_g, (t | half | kk) != 0);
              tc_commit_pair(&misc->b_empty[u], 3);
              if (half == 1 && nh == 1) tc_commit_pair(&misc->g_empty[s], 3);
            }
            __syncwarp();
            if (prof) w_i += clock64() - i0;
            if (++u == kUnitsB) { u = 0; ph ^= 1; }
          }
        }
        if (++s == kStagesG) { s = 0; phs ^= 1; }
      }
      if (elect_one()) tc_commit_pair(&misc->dx_full, 3);
      __syncwarp();
      if (prof && lane == 0 && ((blockIdx.x >> 1) % 97) == 5)
        printf("bwd_e2 prof cluster %d: issue warp total %lld clk, waits g_full %lld b_full %lld, proxy fence %lld, mma issue + commit %lld (steps %d)\n",
               blockIdx.x >> 1, clock64() - t_begin, w_g, w_b, w_f, w_i, T);
    }
  } else if (warp >= 4) {
    // ---------------- scaling warps: E (registers) -> G (shared memory, A operand) ----------------
    const uint32_t ts = tid - 128;                 // 0..511
    const uint32_t sw = ts >> 5;                   // scaling warp
    const int row_t = static_cast<int>(ts & 127);  // row of the tile this thread scales (lanes = consecutive rows)
    const uint32_t ch = ts >> 7;                   // 32-column chunk of the step this warp scales (warp-uniform)
    const uint32_t sl = ch >> 1;                   // its 64-column slab (K half)
    const uint32_t lp0 = (ch & 1) * 4;             // first of its four 8-column pieces inside the slab
    const int grow = tr * 128 + row_t;
    const bool row_valid = tile_valid && grow < p.n_rows;
    const int label = p.label_offset + grow;
    const float scale = __ldg(p.scale);
    const float a_sum = p.a_row + p.a_col;
    constexpr int fmt = kBf16 ? 1 : 0;
    const bool want_ds = p.dscale_part != nullptr;
    float* kc_w = misc->kc[sw];

    // Prefetch position (two steps ahead of the step being scaled): column tensor j_pf, 128-column step tc_pf.  Kept as a
    // pair of counters: no integer division in the loop.
    int j_pf = j_first, tc_pf = tc_first;
    // statistics of one step: this row's chunk offset and (lane = column) one column's log-sum-exp; the row's own
    // log-sum-exp changes only with the column tensor
    struct Stats { float off, lcv; };
    auto load_stats = [&]() {
      Stats st;
      const int pair = i * p.gy + j_pf;
      const int chunk = tc_pf * 4 + static_cast<int>(ch);
      st.off = (row_valid && chunk < p.n_chunks)
                   ? __ldg(p.off + (static_cast<size_t>(pair) * p.n_chunks + chunk) * p.n_rows + grow) : 0.f;
      const int c = chunk * 32 + static_cast<int>(lane);
      st.lcv = c < p.n_cols ? __ldg(p.col_lse2 + static_cast<size_t>(pair) * p.n_cols + c) : INFINITY;
      return st;
    };
    // this thread's 4 pieces (16 bytes = 8 columns of its row) of a step's E tile: pieces ch * 4 .. + 3 of the tile's 16
    const uint4* e_base = reinterpret_cast<const uint4*>(p.e) + (row_t >> 5) * 512 + (row_t & 31) + ch * 4 * 32;
    auto load_e = [&](uint4 (&dst)[4]) {
      const uint4* src = e_base + (static_cast<size_t>((i * p.gy + j_pf) * p.n_row_tiles + tr) * n_ct + tc_pf) * 2048;
      const int c_first = tc_pf * 128 + static_cast<int>(ch) * 32;
#pragma unroll
      for (int p4 = 0; p4 < 4; ++p4) {
        // pieces the forward never wrote (rows past the batch, columns past the last chunk) must not reach the tensor core
        const bool ok = row_valid && c_first + p4 * 8 < p.n_cols;
        dst[p4] = ok ? __ldcs(src + p4 * 32) : make_uint4(0u, 0u, 0u, 0u);
      }
    };
    auto advance_pf = [&]() {
      if (++tc_pf == n_ct) { tc_pf = 0; ++j_pf; }
    };
    // Three register sets (statistics + the four E pieces of a step) used round-robin by a loop unrolled three times: the
    // loads of step t + 2 go into the set that was consumed at step t - 1, and NO register is moved from one set to another.
    // (The first version rotated two sets through a third with moves; a move waits for its source, so the loads were
    // effectively consumed one step after their issue, and the warps spent 40 % of their time - ncu: 30 % of all samples
    // at the first use of the step's offset - waiting for them.)
    Stats sA, sB, sC;
    uint4 eA[4], eB[4], eC[4];
    auto issue = [&](Stats& st, uint4 (&e)[4]) {
      st = load_stats();
      load_e(e);
      advance_pf();
    };
    sA = sB = sC = Stats{0.f, INFINITY};
#pragma unroll
    for (int k = 0; k < 4; ++k) eA[k] = eB[k] = eC[k] = make_uint4(0u, 0u, 0u, 0u);
    if (T > 0) issue(sA, eA);
    if (T > 1) issue(sB, eB);

    const bool eprof = kProf && ((blockIdx.x >> 1) % 97) == 5 && lane == 0 && (sw == 0 || sw == 15);
    long long e_wait = 0, e_work = 0, e_pre = 0, e_fence = 0, e_arrive = 0;
    uint32_t s = 0, phs = 0;
    int j = j_first, tc = tc_first;
    float lr = row_valid ? __ldg(p.row_lse2 + static_cast<size_t>(i * p.gy + j_first) * p.n_rows + grow) : INFINITY;
    int t = 0;
    auto step = [&](const Stats& st, const uint4 (&e_cur)[4]) {
      const long long c_top = eprof ? clock64() : 0;
      const int col0 = tc * 128 + static_cast<int>(ch) * 32;      // first column of this warp's chunk
      const bool chunk_valid = col0 < p.n_cols;                  // warp-uniform
      // column factors of this chunk, relative to o = lse_col of its first column (valid whenever the chunk is)
      const float o = __shfl_sync(0xffffffffu, st.lcv, 0);
      bool risky = false;
      if (chunk_valid) {
        risky = st.lcv != INFINITY && fabsf(o - st.lcv) > 60.f;
        if (row_valid) risky = risky || fabsf(st.off - o) > 60.f;
      }
      // can a factor of the product form 2^(off - o) * 2^(o - lse_col) leave fp32's range in this chunk? (each stays within
      // 2^+-60 otherwise)  Then the chunk uses the exact exponent of every element instead: one ex2 per element.
      const bool slow = __any_sync(0xffffffffu, risky);
      __syncwarp();                                 // every lane has read the previous step's factors
      kc_w[lane] = chunk_valid ? ex2(o - st.lcv) : 0.f;         // 0 for the columns past n_cols
      __syncwarp();

      long long c0 = 0, c1 = 0;
      if (eprof) c0 = clock64();
      mbar_wait(&misc->g_empty[s], phs ^ 1);        // the MMAs that read this stage three steps ago are done
      if (eprof) c1 = clock64();
      const uint32_t stage = smem_u32(sG + s * kStageG) + sl * kSlabG + lp0 * 2048 + row_t * 16;
      float A1 = 0.f, A2 = 0.f;
      if (row_valid && chunk_valid) {
        const float pr = ex2(st.off - lr);                    // <= 1: the running maximum never exceeds the row's log-sum-exp
        const float qc = slow ? 1.f : ex2(st.off - o);
        A1 = p.a_row * pr;
        A2 = p.a_col * qc;
      }
      uint16_t* g_row = p.g_out == nullptr ? nullptr
                            : reinterpret_cast<uint16_t*>(p.g_out) + (static_cast<size_t>(i) * p.n_rows + grow) * static_cast<size_t>(p.g_ld) +
                                  static_cast<size_t>(j) * p.n_cols + col0;
      if (kBf16 && !slow && !(p.dbg & 2048)) {      // warp-uniform; the common case: bf16 stack, every factor in range
        // G = e * f, f = A2 * kc + A1 in fp32, rounded once to bf16; the product is ONE packed bf16 multiply per two elements:
        // e is bf16 already (no unpack) and the result is the operand format (no pack) - 16 instructions per 8 elements instead
        // of 28.  Rows past the batch have e = 0 and A1 = A2 = 0, columns past n_cols e = 0 and kc = 0: their G is 0 without a test.
#pragma unroll
        for (int p4 = 0; p4 < 4; ++p4) {
          const uint4 w = e_cur[p4];
          const float4 k0 = *reinterpret_cast<const float4*>(&kc_w[p4 * 8]);
          const float4 k1 = *reinterpret_cast<const float4*>(&kc_w[p4 * 8 + 4]);
          const uint32_t f01 = factor_pair(k0.x, k0.y, A2, A1);
          const uint32_t f23 = factor_pair(k0.z, k0.w, A2, A1);
          const uint32_t f45 = factor_pair(k1.x, k1.y, A2, A1);
          const uint32_t f67 = factor_pair(k1.z, k1.w, A2, A1);
          const uint4 outv = make_uint4(mul_bf16x2(w.x, f01), mul_bf16x2(w.y, f23), mul_bf16x2(w.z, f45), mul_bf16x2(w.w, f67));
          sts128(stage + p4 * 2048, outv);          // a warp's store of one piece: 512 contiguous bytes, no bank conflicts
          if (g_row != nullptr && row_valid && col0 + p4 * 8 < p.n_cols) *reinterpret_cast<uint4*>(g_row + p4 * 8) = outv;
        }
        // the row's positive (once per row and column tensor): R + C - (a_row + a_col) nearly cancels for a confident row, so
        // it is formed in fp32 from the forward's own dot product of the pair, not from the bf16 exponential
        const int lrel = label - col0;
        if (row_valid && static_cast<uint32_t>(lrel) < 32u) {
          const int pi = lrel >> 3, k = lrel & 7;
          const float g = positive_grad(p, i * p.gy + j, grow, label, scale * kLog2e, lr);
          const uint16_t gb = static_cast<uint16_t>(pack2(g, 0.f, 1) & 0xffffu);
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(stage + pi * 2048 + k * 2), "h"(gb) : "memory");
          if (g_row != nullptr) g_row[lrel] = gb;
        }
      } else if (kBf16) {
        // rare in a bf16 launch (a factor out of range): not unrolled, so that its registers do not weigh on the loop above
        const float g_pos = (row_valid && static_cast<uint32_t>(label - col0) < 32u)
                                ? positive_grad(p, i * p.gy + j, grow, label, scale * kLog2e, lr) : 0.f;
#pragma unroll 1
        for (int p4 = 0; p4 < 4; ++p4) {
          const uint4 w = p4 == 0 ? e_cur[0] : p4 == 1 ? e_cur[1] : p4 == 2 ? e_cur[2] : e_cur[3];
          const int c0p = col0 + p4 * 8;
          const uint4 outv = scale_piece_generic(w, p4, st.off, st.lcv, A1, A2, slow, fmt, label, c0p, row_valid, kc_w, p.n_cols, g_pos, p.dbg);
          sts128(stage + p4 * 2048, outv);
          if (g_row != nullptr && row_valid && c0p < p.n_cols) *reinterpret_cast<uint4*>(g_row + p4 * 8) = outv;
        }
      } else {
        const float g_pos = (row_valid && static_cast<uint32_t>(label - col0) < 32u)
                                ? positive_grad(p, i * p.gy + j, grow, label, scale * kLog2e, lr) : 0.f;
#pragma unroll
        for (int p4 = 0; p4 < 4; ++p4) {
          const int c0p = col0 + p4 * 8;
          const uint4 outv = scale_piece_generic(e_cur[p4], p4, st.off, st.lcv, A1, A2, slow, fmt, label, c0p, row_valid, kc_w, p.n_cols,
                                                 g_pos, p.dbg);
          sts128(stage + p4 * 2048, outv);
          if (g_row != nullptr && row_valid && c0p < p.n_cols) *reinterpret_cast<uint4*>(g_row + p4 * 8) = outv;
        }
      }
      const long long c2 = eprof ? clock64() : 0;
      // ordinary shared-memory stores -> visible to the tensor core's (async proxy) reads
      fence_proxy_async_smem();
      const long long c3 = eprof ? clock64() : 0;
      __syncwarp();
      // Each CTA's tensor core reads its OWN rows of A, so the data never crosses CTAs: the proxy fence above plus a plain
      // remote arrive orders it (a release at cluster scope on this arrive cost ~3000 cycles per step in the second CTA)
      if (lane == 0) {
        if (leader) mbar_arrive(&misc->g_full[s][sl]);
        else mbar_arrive_cluster(&misc->g_full[s][sl], 0);
      }
      if (eprof) {
        e_wait += c1 - c0;
        e_work += c2 - c1;
        e_fence += c3 - c2;
        e_arrive += clock64() - c3;
        e_pre += c0 - c_top;
      }
      if (++s == kStagesG) { s = 0; phs ^= 1; }
      if (++tc == n_ct) {
        tc = 0;
        ++j;
        if (j < p.gy && row_valid && t + 1 < T) lr = __ldg(p.row_lse2 + static_cast<size_t>(i * p.gy + j) * p.n_rows + grow);
      }
      ++t;
    };
    while (t < T) {
      if (t + 2 < T) issue(sC, eC);
      step(sA, eA);
      if (t >= T) break;
      if (t + 2 < T) issue(sA, eA);
      step(sB, eB);
      if (t >= T) break;
      if (t + 2 < T) issue(sB, eB);
      step(sC, eC);
    }
    if (eprof)
      printf("bwd_e2 prof cluster %d cta %u warp %u: scaling warps: loads+factors %lld, stage wait %lld, scale+store %lld, proxy fence %lld, "
             "syncwarp+arrive %lld (steps %d)\n", blockIdx.x >> 1, r, sw, e_pre, e_wait, e_work, e_fence, e_arrive, T);
    // drain dX: lanes = rows of the tile (warp % 4 selects the TMEM lane quarter), 128 columns per warp.
    // d(scale) needs no pass of its own: sum_rc G[r][c] <x_r, y_c> = sum_r <x_r, (G y)_r>, the dot product of every row of X
    // with its fp32 accumulator row (the mode scalars weigh d(scale) like G in every mode this kernel accepts).
    mbar_wait(&misc->dx_full, 0);
    tc_fence_after();
    const uint32_t q = warp & 3, h = sw >> 2;
    const int drow = tr * 128 + static_cast<int>(q) * 32 + static_cast<int>(lane);
    const bool drow_valid = tile_valid && drow < p.n_rows;
    const float coef = __ldg(p.upstream) * p.weight * scale;
    const uint16_t* xrow = reinterpret_cast<const uint16_t*>(p.x) + (static_cast<size_t>(i) * p.n_rows + (drow_valid ? drow : 0)) * 512;
    float ds_acc = 0.f;
    for (int c = static_cast<int>(h) * 128; c < static_cast<int>(h) * 128 + 128; c += 32) {
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
        if (p.dx32 != nullptr) {       // sliced sweep: this slice's fp32 partial, [slice][gx][n_rows][512]
          float4* dst = reinterpret_cast<float4*>(p.dx32 + ((static_cast<size_t>(split) * p.gx + i) * p.n_rows + drow) * 512 + c);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            dst[kk] = make_float4(__uint_as_float(v[4 * kk]) * coef, __uint_as_float(v[4 * kk + 1]) * coef,
                                  __uint_as_float(v[4 * kk + 2]) * coef, __uint_as_float(v[4 * kk + 3]) * coef);
        } else {
          uint32_t ow[16];
#pragma unroll
          for (int kk = 0; kk < 16; ++kk)
            ow[kk] = pack2(__uint_as_float(v[2 * kk]) * coef, __uint_as_float(v[2 * kk + 1]) * coef, fmt);
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.dx) + (static_cast<size_t>(i) * p.n_rows + drow) * 512 + c);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) dst[kk] = make_uint4(ow[4 * kk], ow[4 * kk + 1], ow[4 * kk + 2], ow[4 * kk + 3]);
        }
      }
    }
    if (want_ds) {
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) ds_acc += __shfl_xor_sync(0xffffffffu, ds_acc, sft);
      if (lane == 0) misc->red[sw] = ds_acc;
      named_bar_sync(2, kScale);
      if (ts == 0 && tile_valid) {
        float sum = 0.f;
        for (int w = 0; w < kScaleWarps; ++w) sum += misc->red[w];
        p.dscale_part[(split * p.gx + i) * p.n_row_tiles + tr] = sum;      // <G, raw dot products> of this tile's rows and slice
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair<512>(tmem);
}

cudaError_t launch_infonce_bwd_e2(const CUtensorMap& tmY64, const BwdEParams& p, cudaStream_t stream) {
  const int smem_bytes = kStagesG * kStageG + kUnitsB * kUnitB + kSmemMisc;
  const bool bf16 = p.dtype == COSMOS_DTYPE_BF16, prof = (p.dbg & 1024) != 0;
  void (*kern)(CUtensorMap, BwdEParams) =
      bf16 ? (prof ? infonce_bwd_e2_kernel<true, true> : infonce_bwd_e2_kernel<true, false>)
           : (prof ? infonce_bwd_e2_kernel<false, true> : infonce_bwd_e2_kernel<false, false>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.t_splits * p.gx * ((p.n_row_tiles + 1) / 2) * 2);
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
  return cudaLaunchKernelEx(&cfg, kern, tmY64, p);
}

}  // namespace cb
