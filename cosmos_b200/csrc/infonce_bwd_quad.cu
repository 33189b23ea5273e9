// InfoNCE backward, 4-CTA cluster version for dim = 512 (the product path of the headline configuration).
//
// Why: a [128 x 512] fp32 dX accumulator is all 512 tensor-memory columns of an SM, so one CTA can only
// accumulate a 256-wide part of the embedding next to the S tile it recomputes; the pair kernel
// (infonce_bwd_pair.cu) therefore recomputes every S tile twice, once per part.  Here the two CTA pairs that own
// the two parts of the SAME 256 rows sit in one cluster and SHARE the softmax-gradient tiles:
//   pair p (cluster ranks 2p, 2p+1; tcgen05 cta_group::2, M = 256) computes S_t and G_t only for steps t = p mod 2,
//   keeps G_t in its own tensor memory (A operand of its GEMM2, tcgen05.mma [d],[a_tmem],b_desc) and ALSO stores it,
//   128B-swizzled, into the shared memory of the sibling CTA of the other pair through DSMEM (st.async, completion
//   counted in bytes on an mbarrier of the destination CTA), which uses it as the shared-memory A operand of ITS GEMM2
//   for that step.  The stores are shaped for the DSMEM path (8 rows x 64 contiguous bytes per warp store, see the
//   lane-quad transpose in the epilogue): 2x the bandwidth of the natural lane = row shape (tools/dsmem_bw.cu).
// Executed work per pair of steps and SM: one S tile (2048 MMA cycles) + two dX updates (2 x 1024) instead of
// two S tiles + two updates: 1.5x fewer tensor-core cycles, half the softmax work, 2/3 of the operand traffic.
//
// Issue order of each pair (o_k own steps, q_k remote steps):
//   G1(o_0) | G1a(o_{k+1})  G2(o_k)  G1b(o_{k+1})  G2(q_k) | ...
// so the softmax of an own step has two MMA slots (2048 cycles) of cover, every GEMM2 is followed by half a GEMM1
// while its next operand half streams in, and the in-order tensor pipe keeps G1(o_{k+1}) from overwriting the
// buffer G2(o_{k-1}) reads.  (Consuming the remote tile earlier in the round serialises the two pairs: G2(q_k) then
// blocks the in-order issue of G1b(o_{k+1}), whose softmax produces the tile the other pair is waiting for.)
// Barriers: *_full on the pair leader (TMA bytes of both CTAs / arrivals of both epilogues), *_empty by multicast
// tcgen05.commit; gr_full lives in every CONSUMING CTA (complete_tx bytes of the sibling's st.async), a helper thread
// relays it to its pair leader's gr_ready; gr_empty is committed back to the producing CTAs.
// COSMOS_B200_DBG (diagnostics, wrong results unless noted): 512 two-slot operand ring, 1024 print stall counters
// (results unchanged), 2048 no softmax math, 4096 one eighth of the exchange.
#include <cstdio>
#include "common.cuh"
#include "infonce.h"
#include "internal.h"

namespace cb {

namespace {

constexpr int BM = kFwdBM, BN = kBwdBN;
constexpr int kSlabX = 128 * 64 * 2;    // 16 KB
constexpr int kSlotA = 64 * 64 * 2;     // 8 KB : this CTA's 64 columns of a GEMM1 B slab
constexpr int kSlotsA = 4;
constexpr int kSlabB = 64 * 64 * 2;     // 8 KB : 64 columns (half a step) x 64 embedding elements
constexpr int kStageB = 2 * kSlabB;     // this CTA's 128 embedding columns of one half step
constexpr int kSmemG = 2 * kSlabX;      // 32 KB: G tile received from the sibling pair, two [128 x 64] K-major slabs
constexpr int kSmemMisc = 3072;
constexpr int kThreads = 384;
constexpr int kKs = 8;                  // dim 512

struct Misc {
  uint64_t x_full;
  uint64_t a_full[kSlotsA];
  uint64_t a_empty[kSlotsA];
  uint64_t b_full[2];
  uint64_t b_empty[2];
  uint64_t s_full[2];
  uint64_t g_full[2];
  uint64_t gr_full;    // per CTA: bytes of the G tile the sibling CTA sends us (st.async complete_tx)
  uint64_t gr_ready;   // pair leader: both CTAs of the pair hold the tile
  uint64_t gr_empty;
  uint64_t dx_full;
  uint32_t tmem_slot;
  uint32_t pad[3];
  float red[8];
  alignas(16) float kappa[8][64];
};
static_assert(sizeof(Misc) <= kSmemMisc, "misc smem");
static_assert(kKs * kSlabX + kSlotsA * kSlotA + 2 * kStageB + kSmemG + kSmemMisc <= 232448, "shared memory budget");

}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
infonce_bwd_quad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY64, BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();           // 0..3
  const uint32_t pr = rank >> 1;                     // pair index == embedding part == parity of the steps this pair computes
  const uint32_t r = rank & 1;                       // row tile inside the 256-row block
  const uint32_t L = rank & ~1u;                     // leader of my pair
  const uint32_t Lo = L ^ 2u;                        // leader of the other pair
  const uint32_t sibling = rank ^ 2u;                // same row tile, other part
  const bool leader = r == 0;
  const uint16_t my_pair_mask = static_cast<uint16_t>(3u << L);
  const uint16_t other_pair_mask = static_cast<uint16_t>(3u << Lo);

  const int tiles_padded = 2 * ((p.n_row_tiles + 1) / 2);
  const int rt = (blockIdx.x >> 2) * 2 + r;
  const int i = rt / tiles_padded;
  const int tr = rt - i * tiles_padded;
  const bool tile_valid = tr < p.n_row_tiles;
  const int slab0 = pr * 4;
  const int n_ct = p.n_col_tiles;
  const int T = p.gy * n_ct;
  const int n_own = (T - static_cast<int>(pr) + 1) / 2;          // steps t = pr, pr + 2, ...
  const int n_rem = (T - static_cast<int>(1 - pr) + 1) / 2;      // steps of the other pair
  const int n_round = n_own > n_rem ? n_own : n_rem;
  const uint32_t n_slots_a = (p.dbg & 512) ? 2u : static_cast<uint32_t>(kSlotsA);   // diagnostics: ring-depth sensitivity

  uint8_t* sX = smem;
  uint8_t* sA = sX + kKs * kSlabX;
  uint8_t* sB = sA + kSlotsA * kSlotA;
  uint8_t* sG = sB + 2 * kStageB;
  Misc* misc = reinterpret_cast<Misc*>(sG + kSmemG);

  cluster_sync_all();
  if (tid == 0) {
    mbar_init(&misc->x_full, 2);
    for (int s = 0; s < kSlotsA; ++s) {
      mbar_init(&misc->a_full[s], 2);
      mbar_init(&misc->a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&misc->b_full[s], 2);
      mbar_init(&misc->b_empty[s], 1);
      mbar_init(&misc->s_full[s], 1);
      mbar_init(&misc->g_full[s], 2 * 8);
    }
    mbar_init(&misc->gr_full, 1);
    mbar_init(&misc->gr_ready, 2);
    mbar_init(&misc->gr_empty, 1);
    mbar_init(&misc->dx_full, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY64);
  }
  if (warp == 2) tmem_alloc_pair<512>(&misc->tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = misc->tmem_slot;

  auto arm = [&](uint64_t* bar, uint32_t bytes_per_cta) {
    if (leader) mbar_expect_tx(bar, 2 * bytes_per_cta);
    else mbar_arrive_cluster(bar, L);
  };

  // Single-thread roles run with the WHOLE warp converged and an elect.sync predicate around the asynchronous instructions:
  // inside a divergent `lane == 0` branch the compiler wraps every UTMALDG / UTCHMMA / UTCBAR in an elect-and-branch
  // loop (~14 instructions and a branch per MMA), which made the issuing thread, not the tensor pipe, the bottleneck.
  if (warp == 0) {
    // ---------------- TMA producer, ring A: this CTA's 64-column halves of the S-tile operands of OWN steps ----------------
    if (elect_one()) {
      for (int s = 0; s < kKs; ++s) tma_load_3d_pair(sX + s * kSlabX, &tmX, &misc->x_full, s * 64, tr * BM, i);
      arm(&misc->x_full, kKs * kSlabX);
    }
    __syncwarp();
    uint32_t sa = 0, pa = 0;
    for (int k = 0; k < n_own; ++k) {
      const int t = 2 * k + pr;
      const int j = t / n_ct, tc = t - j * n_ct;
      for (int s = 0; s < kKs; ++s) {
        mbar_wait(&misc->a_empty[sa], pa ^ 1);
        if (elect_one()) {
          tma_load_3d_pair(sA + sa * kSlotA, &tmY64, &misc->a_full[sa], s * 64, tc * BN + r * 64, j);
          arm(&misc->a_full[sa], kSlotA);
        }
        __syncwarp();
        if (++sa == n_slots_a) { sa = 0; pa ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ---------------- TMA producer, ring B: GEMM2 operands of EVERY step, in the order the MMA warp uses them ----------------
    uint32_t sb = 0, pb = 0;
    auto load_b = [&](int t) {
      const int j = t / n_ct, tc = t - j * n_ct;
      for (int half = 0; half < 2; ++half) {
        mbar_wait(&misc->b_empty[sb], pb ^ 1);
        if (elect_one()) {
          for (int s = 0; s < 2; ++s)
            tma_load_3d_pair(sB + sb * kStageB + s * kSlabB, &tmY64, &misc->b_full[sb], (slab0 + r * 2 + s) * 64,
                             tc * BN + half * 64, j);
          arm(&misc->b_full[sb], kStageB);
        }
        __syncwarp();
        sb ^= 1;
        if (sb == 0) pb ^= 1;
      }
    };
    for (int k = 0; k < n_round; ++k) {
      if (k < n_own) load_b(2 * k + pr);
      if (k < n_rem) load_b(2 * k + (1 - pr));
    }
  } else if (warp == 2) {
    if (lane == 0) {
      // ---------------- receive side of the G exchange: arm the byte count, wait for the tile, tell the pair leader ----------------
      for (int k = 0; k < n_rem; ++k) {
        mbar_expect_tx(&misc->gr_full, (p.dbg & 4096) ? kSmemG / 8 : kSmemG);
        // plain (CTA-scope) wait: a cluster-scope acquire compiles to an L1 invalidate (CCTL.IVALL) per poll, which this
        // spinning thread would inflict on the epilogue warps' global loads; the consumer fences once, below
        mbar_wait(&misc->gr_full, k & 1);
        if (leader) mbar_arrive(&misc->gr_ready);
        else mbar_arrive_remote_release(mapa_u32(smem_u32(&misc->gr_ready), L));
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------- MMA issuer (leader of each pair; whole warp waits, one elected lane issues) ----------------
      const bool prof = (p.dbg & 1024) != 0;
      long long w_a = 0, w_g = 0, w_gr = 0, w_b = 0;
      const long long t_begin = clock64();
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
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      bool dx_started = false;
      auto gemm1_half = [&](int k, int hlf) {          // own step index k, K slabs [4*hlf, 4*hlf + 4)
        const uint32_t buf = k & 1;
        for (int s = 4 * hlf; s < 4 * hlf + 4; ++s) {
          wait_t(&misc->a_full[sa], pa, w_a);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sX + s * kSlabX);
          const uint32_t b_base = smem_u32(sA + sa * kSlotA);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_ss_pair(tmem + buf * BN, make_smem_desc(a_base + kk * 32, 0, 1024), make_smem_desc(b_base + kk * 32, 0, 1024),
                           p.idesc_s, (s | kk) != 0);
            tc_commit_pair(&misc->a_empty[sa], my_pair_mask);
            if (hlf == 1 && s == 7) tc_commit_pair(&misc->s_full[buf], my_pair_mask);
          }
          __syncwarp();
          if (++sa == n_slots_a) { sa = 0; pa ^= 1; }
        }
      };
      auto gemm2 = [&](bool own, int k) {
        const uint32_t buf = k & 1;
        if (own) {
          wait_t(&misc->g_full[buf], (k >> 1) & 1, w_g);
        } else {
          wait_t(&misc->gr_ready, k & 1, w_gr);           // G tile written by the sibling pair through DSMEM, in both CTAs
          fence_proxy_async_all();                     // st.async data (complete_tx observed through the barrier chain) -> UMMA
        }
        for (int half = 0; half < 2; ++half) {
          wait_t(&misc->b_full[sb], pb, w_b);
          tc_fence_after();
          const uint32_t b_base = smem_u32(sB + sb * kStageB);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t db = make_smem_desc(b_base + kk * 2048, kSlabB, 1024);
              const uint32_t acc = (dx_started || kk > 0) ? 1u : 0u;
              if (own) {
                umma_ts_pair(tmem + 256, tmem + buf * BN + half * 64 + kk * 8, db, p.idesc_g, acc);
              } else {
                const uint64_t da = make_smem_desc(smem_u32(sG) + half * kSlabX + kk * 32, 0, 1024);
                umma_ss_pair(tmem + 256, da, db, p.idesc_g, acc);
              }
            }
            tc_commit_pair(&misc->b_empty[sb], my_pair_mask);
            if (!own && half == 1) tc_commit_pair(&misc->gr_empty, other_pair_mask);   // the producers may overwrite our G buffer
          }
          __syncwarp();
          dx_started = true;
          sb ^= 1;
          if (sb == 0) pb ^= 1;
        }
      };
      if (n_own > 0) {
        gemm1_half(0, 0);
        gemm1_half(0, 1);
      }
      for (int k = 0; k < n_round; ++k) {
        if (k + 1 < n_own) gemm1_half(k + 1, 0);
        if (k < n_own) gemm2(true, k);
        if (k + 1 < n_own) gemm1_half(k + 1, 1);
        if (k < n_rem) gemm2(false, k);
      }
      if (elect_one()) tc_commit_pair(&misc->dx_full, my_pair_mask);
      __syncwarp();
      if (prof && lane == 0 && ((blockIdx.x >> 2) % 97) == 5)
        printf("quad prof cluster %d pair %u: issue thread total %lld clk, waits a_full %lld g_full %lld gr_ready %lld b_full %lld (rounds %d)\n",
               blockIdx.x >> 2, pr, clock64() - t_begin, w_a, w_g, w_gr, w_b, n_round);
    }
  } else if (warp >= 4) {
    // ---------------- epilogue: softmax gradient of OWN steps, kept in TMEM and sent to the sibling CTA ----------------
    const uint32_t ew = warp - 4;
    const uint32_t q = warp & 3;
    const uint32_t h = ew >> 2;
    const int r_t = q * 32 + lane;
    const int row = tr * BM + r_t;
    const bool row_valid = tile_valid && row < p.n_rows;
    const int label = p.label_offset + row;
    const float scale = __ldg(p.scale);
    const float k2 = scale * kLog2e;
    const float a_sum = p.a_row + p.a_col, s_sum = p.s_row + p.s_col;
    const int fmt = p.dtype == COSMOS_DTYPE_BF16 ? 1 : 0;
    const uint32_t lane_base = (q * 32u) << 16;
    float ds_acc = 0.f;
    const bool ds_prop = fabsf(p.a_row * p.s_col - p.a_col * p.s_row) <= 1e-12f && (a_sum != 0.f);
    const float ds_ratio = ds_prop ? s_sum / a_sum : 0.f;
    const bool ds_rows = !ds_prop && p.s_col == 0.f;
    const bool fast_ok = (ds_prop || ds_rows) && !(p.dbg & 32);
    float* kbuf = misc->kappa[ew];
    const uint32_t g_remote_slab = mapa_u32(smem_u32(sG) + h * kSlabX, sibling);
    const uint32_t gr_full_remote = mapa_u32(smem_u32(&misc->gr_full), sibling);

    const bool eprof = (p.dbg & 1024) != 0 && ((blockIdx.x >> 2) % 97) == 5 && lane == 0 && (ew == 0 || ew == 7);
    long long e_wait = 0, e_work = 0, e_gre = 0, e_send = 0;
    for (int k = 0; k < n_own; ++k) {
      const int t = 2 * k + pr;
      const int j = t / n_ct, tc = t - j * n_ct;
      const int pair = i * p.gy + j;
      const uint32_t buf = k & 1;
      const float lr = row_valid ? __ldg(p.row_lse2 + static_cast<size_t>(pair) * p.n_rows + row) : INFINITY;
      const float* lc_ptr = p.col_lse2 + static_cast<size_t>(pair) * p.n_cols;
      const int colw = tc * BN + h * 64;
      const bool full = colw + 64 <= p.n_cols;

      bool fast = fast_ok && full;
      float o = 0.f, rho = 0.f, rho_s = 0.f;
      if (fast) {
        const float lc0 = __ldg(lc_ptr + colw + lane), lc1 = __ldg(lc_ptr + colw + 32 + lane);
        float lo = fminf(lc0, lc1), hi = fmaxf(lc0, lc1);
        if (row_valid) { lo = fminf(lo, lr); hi = fmaxf(hi, lr); }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) {
          lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, sft));
          hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, sft));
        }
        fast = (hi - lo) <= 200.f;
        if (fast) {
          o = 0.5f * (hi + lo);
          rho = p.a_row * ex2(o - lr);
          rho_s = p.s_row * ex2(o - lr);
          __syncwarp();
          kbuf[lane] = p.a_col * ex2(o - lc0);
          kbuf[32 + lane] = p.a_col * ex2(o - lc1);
          __syncwarp();
        }
      }

      long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
      if (eprof) c0 = clock64();
      mbar_wait(&misc->s_full[buf], (k >> 1) & 1);
      tc_fence_after();
      if (eprof) c1 = clock64();
      uint32_t packed_all[2][16];
#pragma unroll
      for (int chunk = 0; chunk < 2; ++chunk) {
        const int col0 = colw + chunk * 32;
        uint32_t (&packed)[16] = packed_all[chunk];
        uint32_t v[32];
        tmem_ld32(tmem + lane_base + buf * BN + h * 64 + chunk * 32, v);
        if (p.dbg & 2048) {          // diagnostics: no softmax math (wrong results)
          tmem_ld_wait();
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) packed[kk] = v[2 * kk] ^ v[2 * kk + 1];
        } else if (fast) {
          float kap[32];
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 f = *reinterpret_cast<const float4*>(kbuf + chunk * 32 + k4 * 4);
            kap[4 * k4 + 0] = f.x; kap[4 * k4 + 1] = f.y; kap[4 * k4 + 2] = f.z; kap[4 * k4 + 3] = f.w;
          }
          tmem_ld_wait();
          float g[32];
          float acc = 0.f;
          const float neg_o = -o;
          if (ds_prop) {
#pragma unroll
            for (int kk = 0; kk < 32; ++kk) {
              const float raw = __uint_as_float(v[kk]);
              g[kk] = ex2(fmaf(raw, k2, neg_o)) * (rho + kap[kk]);
              acc = fmaf(g[kk], raw, acc);
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < 32; ++kk) {
              const float raw = __uint_as_float(v[kk]);
              const float e = ex2(fmaf(raw, k2, neg_o));
              g[kk] = e * (rho + kap[kk]);
              acc = fmaf(e, raw, acc);
            }
          }
          if (label >= col0 && label < col0 + 32) {
            const int idx = label - col0;
#pragma unroll
            for (int kk = 0; kk < 32; ++kk)
              if (kk == idx) {
                g[kk] -= a_sum;
                acc -= (ds_prop ? a_sum : (rho_s != 0.f ? p.s_row / rho_s : 0.f)) * __uint_as_float(v[kk]);
              }
          }
          if (row_valid) ds_acc += ds_prop ? ds_ratio * acc : rho_s * acc;
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) packed[kk] = pack2(g[2 * kk], g[2 * kk + 1], fmt);
        } else {
          float lcv[32];
#pragma unroll
          for (int kk = 0; kk < 32; ++kk) lcv[kk] = (col0 + kk < p.n_cols) ? __ldg(lc_ptr + col0 + kk) : INFINITY;
          tmem_ld_wait();
          float g[32];
#pragma unroll
          for (int kk = 0; kk < 32; ++kk) {
            const int col = col0 + kk;
            const float raw = __uint_as_float(v[kk]);
            const float tt = raw * k2;
            const float prw = ex2(tt - lr);
            const float pc = ex2(tt - lcv[kk]);
            float gg = p.a_row * prw + p.a_col * pc;
            float dd = p.s_row * prw + p.s_col * pc;
            if (col == label) {
              gg -= a_sum;
              dd -= s_sum;
            }
            const bool ok = (col < p.n_cols) && row_valid;
            g[kk] = ok ? gg : 0.f;
            ds_acc += ok ? dd * raw : 0.f;
          }
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) packed[kk] = pack2(g[2 * kk], g[2 * kk + 1], fmt);
        }
        // own pair: back into tensor memory, over the S columns this warp has already read
        tmem_st16(tmem + lane_base + buf * BN + h * 64 + chunk * 16, packed);
      }
      // publish the tile to our own GEMM2 first: it is on the critical path, the copy to the sibling pair is not
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&misc->g_full[buf]);
        else mbar_arrive_cluster(&misc->g_full[buf], L);
      }
      // other pair: K-major, 128B-swizzled rows in the sibling CTA's shared memory, once it has consumed the tile
      // we sent one own-step ago
      if (eprof) c2 = clock64();
      mbar_wait(&misc->gr_empty, (k & 1) ^ 1);      // write-after-read only: no data is acquired, a CTA-scope wait suffices
      if (eprof) c3 = clock64();
      // A warp store that touches 32 different rows moves ~10 B/clk per SM through DSMEM, one whose lanes cover 64 contiguous
      // bytes of 8 rows ~19.7 B/clk (tools/dsmem_bw.cu).  So the four lanes that own rows 4g..4g+3 first transpose their
      // 4 x 4 blocks of 16-byte pieces (16 shuffles per block): afterwards lane j holds piece j of each of the four rows.
#pragma unroll
      for (int chunk = 0; chunk < 2; ++chunk) {
        uint4 a[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          a[c] = make_uint4(packed_all[chunk][c * 4 + 0], packed_all[chunk][c * 4 + 1], packed_all[chunk][c * 4 + 2],
                            packed_all[chunk][c * 4 + 3]);
        quad_transpose_u4(a, lane);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if ((p.dbg & 4096) && (chunk | c) != 0) continue;   // diagnostics: 1/8 of the exchange (wrong results)
          const uint32_t rr = (static_cast<uint32_t>(r_t) & ~3u) + c;             // row of the tile this piece belongs to
          const uint32_t pos = (chunk * 4 + (lane & 3)) ^ (rr & 7);                 // 128B swizzle of the K-major slab
          st_async_cluster_v4(g_remote_slab + rr * 128 + (pos << 4), a[c], gr_full_remote);
          if (p.g_out != nullptr) {
            // optional copy of the tile for the column-side gradient GEMM (api.cu: cosmos_infonce_bwd_g); the transposed
            // layout also makes these 8 rows x 64 contiguous bytes per warp store
            const int grow = tr * BM + static_cast<int>(rr);
            const int gcol = tc * BN + static_cast<int>(h) * 64 + (chunk * 4 + static_cast<int>(lane & 3)) * 8;
            if (tile_valid && grow < p.n_rows && gcol < p.n_cols)
              *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.g_out) +
                                        (static_cast<size_t>(i) * p.n_rows + grow) * static_cast<size_t>(p.g_ld) +
                                        static_cast<size_t>(j) * p.n_cols + gcol) = a[c];
          }
        }
      }
      if (eprof) {
        e_wait += c1 - c0;
        e_work += c2 - c1;
        e_gre += c3 - c2;
        e_send += clock64() - c3;
      }
    }
    if (eprof)
      printf("quad prof cluster %d rank %u warp %u: epilogue s_full wait %lld, softmax+publish %lld, gr_empty wait %lld, send %lld (own steps %d)\n",
             blockIdx.x >> 2, rank, ew, e_wait, e_work, e_gre, e_send, n_own);

    if (p.dscale_part != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ds_acc += __shfl_xor_sync(0xffffffffu, ds_acc, o);
      if (lane == 0) misc->red[ew] = ds_acc;
      named_bar_sync(1, 256);
      if (ew == 0 && lane == 0 && tile_valid) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += misc->red[w];
        p.dscale_part[(i * p.n_row_tiles + tr) * 2 + pr] = s;      // each pair reports the steps it computed
      }
    }

    // drain dX: this CTA's 128 rows x its 256-wide part
    mbar_wait(&misc->dx_full, 0);
    tc_fence_after();
    const float coef = __ldg(p.upstream) * p.weight * scale;
    const int dim = kKs * 64;
    const int c_begin = h * 128, c_end = c_begin + 128;
    for (int c = c_begin; c < c_end; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + lane_base + 256 + c, v);
      tmem_ld_wait();
      if (row_valid) {
        uint32_t ow[16];
#pragma unroll
        for (int kk = 0; kk < 16; ++kk)
          ow[kk] = pack2(__uint_as_float(v[2 * kk]) * coef, __uint_as_float(v[2 * kk + 1]) * coef, fmt);
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.dx) +
                                              (static_cast<size_t>(i) * p.n_rows + row) * dim + slab0 * 64 + c);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) dst[kk] = make_uint4(ow[4 * kk], ow[4 * kk + 1], ow[4 * kk + 2], ow[4 * kk + 3]);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair<512>(tmem);
}

cudaError_t launch_infonce_bwd_quad(const CUtensorMap& tmX, const CUtensorMap& tmY64, const BwdParams& p, cudaStream_t stream) {
  const int smem_bytes = kKs * kSlabX + kSlotsA * kSlotA + 2 * kStageB + kSmemG + kSmemMisc;
  cudaError_t e = cudaFuncSetAttribute(infonce_bwd_quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.gx * ((p.n_row_tiles + 1) / 2) * 4);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, infonce_bwd_quad_kernel, tmX, tmY64, p);
}

}  // namespace cb
