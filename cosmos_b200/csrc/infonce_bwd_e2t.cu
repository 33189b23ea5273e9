// Column-side gradient of the stored-exponential route: dY = G^T X straight from the exponentials the forward kept.
//
// The CLIP term (src/open_clip/loss.py:206) needs gradients on BOTH sides of its logits.  infonce_bwd_e2.cu forms the row
// side, dX = G Y.  The column side was a round trip so far: that kernel also wrote every G tile to HBM (2 bytes per logit,
// 16-byte pieces at a 64 KB stride), and a GEMM read them back for G^T X.  This kernel needs neither: it is the same
// pipeline turned by 90 degrees.  One CTA pair (cluster of 2, tcgen05 cta_group::2, M = 256) owns 256 COLUMNS of one column
// tensor j and all 512 embedding columns (fp32 accumulators = the whole tensor memory) and sweeps over the local rows of every
// row tensor i, 128 at a step:
//   16 scaling warps: the E tile of (i, j, row tile, column tile) from global memory into registers two steps ahead, E -> G
//        with the same factors as the row pass (G = e * (a_row 2^(m - lse_row[r]) + a_col 2^(m - lse_col[c])) - positives),
//        stored into a [16 pieces of 8 columns][128 rows][8] image in shared memory (the G stage of the row pass).  Read with the contiguous
//        dimension as M, that image IS the MN-major operand A = G^T without swizzle: core matrix = 8 rows (K) x 16 bytes
//        (8 columns, M), next 8 rows + 128 B (LBO), next 8 columns + 2048 B (SBO).  No transpose is ever executed.
//   B = the X rows of the step (TMA, 128-byte swizzle, MN-major like the Y slabs of the row pass), each CTA of the pair
//        supplying 128 of the 256 N columns;   dY[:, 0:256] += G^T X[:, 0:256],  dY[:, 256:512] += G^T X[:, 256:512].
// The two 64-ROW halves of a step have their own full barriers (a scaling warp owns 32 rows), so the MMAs of the first
// half start while the second is still being scaled.
// Output: fp32 [splits][gy][n_cols][512], the sum over THIS rank's rows only (ranks are combined by a reduce-scatter); the
// row sweep can be cut into `splits` slices so that the launch fills whole waves of CTA pairs (gy * n_cols / 256 pairs are
// few: 256 at 2 x 32768 columns = 3.46 waves of 74).  Unit scale: the caller multiplies by upstream * scale * weight.
#include "infonce_bwd_e_common.cuh"

namespace cb {

using namespace bwd_e;

template <bool kBf16>
__global__ void __launch_bounds__(kThreads, 1)
infonce_bwd_e2t_kernel(const __grid_constant__ CUtensorMap tmX64, BwdEParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t r = cluster_ctarank();
  const bool leader = r == 0;

  // work item: (split, column tensor j, column tile tc); the two CTAs of a pair take consecutive column tiles
  const int n_ct = p.n_col_tiles;
  const int ct_padded = 2 * ((n_ct + 1) / 2);
  const int item = (blockIdx.x >> 1) * 2 + static_cast<int>(r);
  const int split = item / (p.gy * ct_padded);
  const int rem = item - split * (p.gy * ct_padded);
  const int j = rem / ct_padded;
  const int tc = rem - j * ct_padded;
  const bool tile_valid = tc < n_ct;
  const int n_rt = p.n_row_tiles;
  const int T_all = p.gx * n_rt;                               // steps of the whole row sweep
  const int t_begin = static_cast<int>(static_cast<long long>(T_all) * split / p.t_splits);
  const int t_end = static_cast<int>(static_cast<long long>(T_all) * (split + 1) / p.t_splits);
  const int T = t_end - t_begin;                               // the same for both CTAs of a pair (same split)
  const int i_first = t_begin / n_rt, tr_first = t_begin - i_first * n_rt;

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
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmX64);
  if (warp == 2) tmem_alloc_pair<512>(&misc->tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = misc->tmem_slot;

  if (warp == 0) {
    // ---------------- L2 prefetch of this CTA's E tiles ----------------
    // Unlike the row pass, whose tiles follow one another in memory, this sweep jumps n_col_tiles * 32 KB from step to step:
    // every tile is a cold DRAM access (measured: 21.6 ms with the E loads, 10.4 ms without them, for 16 pairs at N = 32768),
    // and the two steps the registers run ahead do not cover it.  One bulk prefetch per contiguous 32 KB tile, kAhead steps
    // early, paced by the consumption of the G stages.
    const int ahead = (p.dbg & 4096) ? 0 : 8;          // 4096: diagnostics, no prefetch
    int i = i_first, tr = tr_first;
    auto prefetch = [&]() {
      if (tile_valid && elect_one())
        bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(p.e) + (static_cast<size_t>((i * p.gy + j) * n_rt + tr) * n_ct + tc) * 32768, 32768);
      __syncwarp();
      if (++tr == n_rt) { tr = 0; ++i; }
    };
    for (int t = 0; t < ahead && t < T; ++t) prefetch();
    uint32_t s = 0, ph = 0;
    for (int t = 0; ahead > 0 && t + ahead < T; ++t) {
      mbar_wait(&misc->g_empty[s], ph ^ 1);
      prefetch();
      if (++s == kStagesG) { s = 0; ph ^= 1; }
    }
  } else if (warp == 3) {
    // ---------------- TMA producer: X slabs (B operand), in the order the MMA warp consumes them ----------------
    uint32_t u = 0, ph = 0;
    int i = i_first, tr = tr_first;
    for (int t = 0; t < T; ++t) {
      for (int half = 0; half < 2; ++half) {
        for (int nh = 0; nh < 2; ++nh) {
          mbar_wait(&misc->b_empty[u], ph ^ 1);
          if (elect_one()) {
            for (int sl = 0; sl < 2; ++sl)
              tma_load_3d_pair(sB + u * kUnitB + sl * kSlabB, &tmX64, &misc->b_full[u], (nh * 4 + static_cast<int>(r) * 2 + sl) * 64,
                               tr * 128 + half * 64, i);         // rows past the batch read as zero
            if (leader) mbar_expect_tx(&misc->b_full[u], 2 * kUnitB);
            else mbar_arrive_cluster(&misc->b_full[u], 0);
          }
          __syncwarp();
          if (++u == kUnitsB) { u = 0; ph ^= 1; }
        }
      }
      if (++tr == n_rt) { tr = 0; ++i; }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------- MMA issuer (pair leader; whole warp waits, one elected lane issues) ----------------
      uint32_t u = 0, ph = 0, s = 0, phs = 0;
      for (int t = 0; t < T; ++t) {
        for (int half = 0; half < 2; ++half) {
          mbar_wait(&misc->g_full[s][half], phs);
          fence_proxy_async_smem();    // G was written by ordinary stores of both CTAs, each fenced before its arrive
          tc_fence_after();
          for (int nh = 0; nh < 2; ++nh) {
            mbar_wait(&misc->b_full[u], ph);
            tc_fence_after();
            // A = G^T: rows 64 * half .. + 63 of the tile image as K, its 16 pieces (128 columns) as M
            const uint32_t a_base = smem_u32(sG + s * kStageG) + half * 1024;
            const uint32_t b_base = smem_u32(sB + u * kUnitB);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_ss_pair(tmem + nh * 256, make_smem_desc_noswizzle(a_base + kk * 256, 128, 2048),
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
    }
  } else if (warp >= 4) {
    // ---------------- scaling warps: E (registers) -> G (shared memory) ----------------
    const uint32_t ts = tid - 128;                 // 0..511
    const uint32_t sw = ts >> 5;                   // scaling warp
    const int row_t = static_cast<int>(ts & 127);  // row of the step's tile this thread scales (lanes = consecutive rows)
    const uint32_t ch = ts >> 7;                   // 32-column chunk of this CTA's column tile (warp-uniform)
    const uint32_t khalf = (sw >> 1) & 1;          // which 64-row half of the step (K half) this warp's rows belong to
    const float k2 = __ldg(p.scale) * kLog2e;
    constexpr int fmt = kBf16 ? 1 : 0;
    float* kc_w = misc->kc[sw];
    const int col0 = tc * 128 + static_cast<int>(ch) * 32;            // first column of this warp's chunk: fixed for the CTA
    const int chunk = tc * 4 + static_cast<int>(ch);
    const bool chunk_valid = tile_valid && col0 < p.n_cols;           // warp-uniform
    const int ccol = col0 + static_cast<int>(lane);                   // lane = column for the column factors

    // Prefetch position (two steps ahead of the step being scaled): row tensor i_pf, row tile tr_pf (no division in the loop)
    int i_pf = i_first, tr_pf = tr_first;
    // statistics of one step: this row's chunk offset and log-sum-exp, and (lane = column) one column's log-sum-exp
    struct Stats { float off, lr, lcv; int grow, pair; };
    auto load_stats = [&]() {
      Stats st;
      const int pair = i_pf * p.gy + j;
      st.pair = pair;
      st.grow = tr_pf * 128 + row_t;
      const bool rv = chunk_valid && st.grow < p.n_rows;
      st.off = rv ? __ldg(p.off + (static_cast<size_t>(pair) * p.n_chunks + chunk) * p.n_rows + st.grow) : 0.f;
      st.lr = rv ? __ldg(p.row_lse2 + static_cast<size_t>(pair) * p.n_rows + st.grow) : INFINITY;
      st.lcv = (chunk_valid && ccol < p.n_cols) ? __ldg(p.col_lse2 + static_cast<size_t>(pair) * p.n_cols + ccol) : INFINITY;
      if (!rv) st.grow = -1;                       // marks a row that does not exist
      return st;
    };
    // this thread's 4 pieces (16 bytes = 8 columns of its row) of a step's E tile: pieces ch * 4 .. + 3 of the tile's 16
    const uint4* e_base = reinterpret_cast<const uint4*>(p.e) + (row_t >> 5) * 512 + (row_t & 31) + ch * 4 * 32;
    auto load_e = [&](uint4 (&dst)[4]) {
      const uint4* src = e_base + (static_cast<size_t>((i_pf * p.gy + j) * n_rt + tr_pf) * n_ct + (tile_valid ? tc : 0)) * 2048;
      const bool rv = chunk_valid && tr_pf * 128 + row_t < p.n_rows;
#pragma unroll
      for (int p4 = 0; p4 < 4; ++p4) {
        // pieces the forward never wrote (rows past the batch, columns past the last chunk) must not reach the tensor core
        const bool ok = rv && col0 + p4 * 8 < p.n_cols && !(p.dbg & 2048);     // 2048: diagnostics, no E traffic (wrong results)
        dst[p4] = ok ? __ldcs(src + p4 * 32) : make_uint4(0u, 0u, 0u, 0u);
      }
    };
    auto advance_pf = [&]() {
      if (++tr_pf == n_rt) { tr_pf = 0; ++i_pf; }
    };
    uint4 e1[4], e2[4];
    Stats s1 = load_stats(), s2 = s1;
    load_e(e1);
    advance_pf();
    if (T > 1) {
      s2 = load_stats();
      load_e(e2);
      advance_pf();
    }

    uint32_t s = 0, phs = 0;
    for (int t = 0; t < T; ++t) {
      const Stats st = s1;
      uint4 e_cur[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) e_cur[k] = e1[k];
      s1 = s2;
#pragma unroll
      for (int k = 0; k < 4; ++k) e1[k] = e2[k];
      if (t + 2 < T) {
        s2 = load_stats();
        load_e(e2);
        advance_pf();
      }
      const bool row_valid = st.grow >= 0;
      const int label = p.label_offset + st.grow;
      const int pair_cur = st.pair;
      // column factors of this chunk, relative to o = lse_col of its first column (valid whenever the chunk is)
      const float o = __shfl_sync(0xffffffffu, st.lcv, 0);
      bool risky = false;
      if (chunk_valid) {
        risky = st.lcv != INFINITY && fabsf(o - st.lcv) > 60.f;
        if (row_valid) risky = risky || fabsf(st.off - o) > 60.f;
      }
      const bool slow = __any_sync(0xffffffffu, risky);    // a factor of the product form may leave fp32's range: exact exponents
      __syncwarp();                                 // every lane has read the previous step's factors
      kc_w[lane] = chunk_valid ? ex2(o - st.lcv) : 0.f;         // 0 for the columns past n_cols
      __syncwarp();

      mbar_wait(&misc->g_empty[s], phs ^ 1);        // the MMAs that read this stage three steps ago are done
      const uint32_t stage = smem_u32(sG + s * kStageG) + ch * 4 * 2048 + row_t * 16;
      float A1 = 0.f, A2 = 0.f;
      if (row_valid) {
        const float pr = ex2(st.off - st.lr);                 // <= 1: the running maximum never exceeds the row's log-sum-exp
        const float qc = slow ? 1.f : ex2(st.off - o);
        A1 = p.a_row * pr;
        A2 = p.a_col * qc;
      }
      if (kBf16 && !slow) {                        // warp-uniform; the common case (see infonce_bwd_e2.cu)
#pragma unroll
        for (int p4 = 0; p4 < 4; ++p4) {
          const uint4 w = e_cur[p4];
          const float4 k0 = *reinterpret_cast<const float4*>(&kc_w[p4 * 8]);
          const float4 k1 = *reinterpret_cast<const float4*>(&kc_w[p4 * 8 + 4]);
          const uint32_t f01 = factor_pair(k0.x, k0.y, A2, A1);
          const uint32_t f23 = factor_pair(k0.z, k0.w, A2, A1);
          const uint32_t f45 = factor_pair(k1.x, k1.y, A2, A1);
          const uint32_t f67 = factor_pair(k1.z, k1.w, A2, A1);
          sts128(stage + p4 * 2048,
                 make_uint4(mul_bf16x2(w.x, f01), mul_bf16x2(w.y, f23), mul_bf16x2(w.z, f45), mul_bf16x2(w.w, f67)));
        }
        const int lrel = label - col0;              // the row's positive, in fp32 from the forward's dot product (cancellation)
        if (row_valid && static_cast<uint32_t>(lrel) < 32u) {
          const float g = positive_grad(p, pair_cur, st.grow, label, k2, st.lr);
          const uint16_t gb = static_cast<uint16_t>(pack2(g, 0.f, 1) & 0xffffu);
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(stage + (lrel >> 3) * 2048 + (lrel & 7) * 2), "h"(gb) : "memory");
        }
      } else {
        const float g_pos = (row_valid && static_cast<uint32_t>(label - col0) < 32u) ? positive_grad(p, pair_cur, st.grow, label, k2, st.lr)
                                                                                     : 0.f;
#pragma unroll 1
        for (int p4 = 0; p4 < 4; ++p4) {
          const uint4 w = p4 == 0 ? e_cur[0] : p4 == 1 ? e_cur[1] : p4 == 2 ? e_cur[2] : e_cur[3];
          sts128(stage + p4 * 2048, scale_piece_generic(w, p4, st.off, st.lcv, A1, A2, slow, fmt, label, col0 + p4 * 8, row_valid,
                                                        kc_w, p.n_cols, g_pos, 0));
        }
      }
      fence_proxy_async_smem();      // ordinary shared-memory stores -> visible to the tensor core's (async proxy) reads
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&misc->g_full[s][khalf]);
        else mbar_arrive_cluster(&misc->g_full[s][khalf], 0);
      }
      if (++s == kStagesG) { s = 0; phs ^= 1; }
    }
    // drain dY: lanes = columns of the tile (warp % 4 selects the TMEM lane quarter), 128 embedding columns per warp
    mbar_wait(&misc->dx_full, 0);
    tc_fence_after();
    const uint32_t q = warp & 3, h = sw >> 2;
    const int dcol = tc * 128 + static_cast<int>(q) * 32 + static_cast<int>(lane);
    const bool dcol_valid = tile_valid && dcol < p.n_cols;
    float* out = reinterpret_cast<float*>(p.dx) +
                 ((static_cast<size_t>(split) * p.gy + j) * p.n_cols + (dcol_valid ? dcol : 0)) * 512;
    for (int c = static_cast<int>(h) * 128; c < static_cast<int>(h) * 128 + 128; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((q * 32u) << 16) + c, v);
      tmem_ld_wait();
      if (dcol_valid) {
        float4* dst = reinterpret_cast<float4*>(out + c);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          dst[kk] = make_float4(__uint_as_float(v[4 * kk]), __uint_as_float(v[4 * kk + 1]), __uint_as_float(v[4 * kk + 2]),
                                __uint_as_float(v[4 * kk + 3]));
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair<512>(tmem);
}

cudaError_t launch_infonce_bwd_e2t(const CUtensorMap& tmX64, const BwdEParams& p, cudaStream_t stream) {
  const int smem_bytes = kStagesG * kStageG + kUnitsB * kUnitB + kSmemMisc;
  const bool bf16 = p.dtype == COSMOS_DTYPE_BF16;
  void (*kern)(CUtensorMap, BwdEParams) = bf16 ? infonce_bwd_e2t_kernel<true> : infonce_bwd_e2t_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.t_splits * p.gy * ((p.n_col_tiles + 1) / 2) * 2);
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
  return cudaLaunchKernelEx(&cfg, kern, tmX64, p);
}

}  // namespace cb
