// InfoNCE backward, CTA-pair version (the product path for dim in {128, 256, 512}).
//
// Same math as infonce_bwd.cu (dX_i = coef * sum_j G_ij Y_j, G rebuilt from the forward's
// log-sum-exps), restructured around what bounds it on B200 - L2->SM operand traffic:
//   * two CTAs on neighbouring SMs form a cluster and issue tcgen05.mma.cta_group::2 (M = 256): each CTA
//     keeps its own 128-row X tile resident and loads only HALF of every Y operand
//       GEMM1  S_t  [256 x 128] : B = Y tile, K-major,  each CTA loads 64 of the 128 columns  (8 KB slabs)
//       GEMM2  dX  [256 x 256] : B = Y tile, MN-major, each CTA loads 128 of the 256 embedding columns
//   * G never touches shared memory: the epilogue writes it back to TENSOR MEMORY (tcgen05.st, two 16-bit
//     values per column, over the S tile it was computed from) and GEMM2 reads A from TMEM
//     (tcgen05.mma [d], [a_tmem], b_desc) - no smem G buffer, no smem-bandwidth limit on GEMM2, and the
//     32 KB it frees deepen the TMA rings.
// TMEM per CTA: S/G double buffer in columns [0,256), dX accumulator [256, 256 + part width).
// Shared memory per CTA: X tile (<= 128 KB) | ring A: 8 x 8 KB GEMM1 slabs | ring B: 1-2 x <= 32 KB GEMM2 operand.
// Issue order G1(0) G1(1) G2(0) G1(2) G2(1) ...: the softmax epilogue of step t overlaps GEMM1 of t+1, and
// because the tensor pipe is in order, GEMM1(t+2) cannot overwrite the buffer GEMM2(t) still reads.
#include "common.cuh"
#include "infonce.h"
#include "internal.h"

namespace cb {

namespace {

constexpr int BM = kFwdBM, BN = kBwdBN;
constexpr int kSlabX = 128 * 64 * 2;    // 16 KB: 128 rows x 64 elements
constexpr int kSlotA = 64 * 64 * 2;     // 8 KB : 64 columns x 64 elements (this CTA's half of a GEMM1 B slab)
constexpr int kSlotsA = 8;
constexpr int kSlabB = 64 * 64 * 2;     // 8 KB : 64 columns (half a step) x 64 embedding elements
constexpr int kSmemMisc = 3072;
constexpr int kThreads = 384;
constexpr int kEpiThreads = 256;

struct Misc {
  uint64_t x_full;
  uint64_t a_full[kSlotsA];
  uint64_t a_empty[kSlotsA];
  uint64_t b_full[2];
  uint64_t b_empty[2];
  uint64_t s_full[2];
  uint64_t g_full[2];
  uint64_t dx_full;
  uint32_t tmem_slot;
  uint32_t pad[3];
  float red[8];
  alignas(16) float kappa[8][64];   // per epilogue warp: a_col * 2^(o - lse_col) of its 64 columns
};
static_assert(sizeof(Misc) <= kSmemMisc, "misc smem");

}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
infonce_bwd_pair_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY64,
                        const __grid_constant__ CUtensorMap tmY128, BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];   // no static shared memory in this kernel: offset 0, 1 KB aligned
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;

  const int ks = p.ks;
  // work item: (row tensor i, pair of row tiles, embedding part); the two CTAs take consecutive row tiles
  const int tiles_padded = 2 * ((p.n_row_tiles + 1) / 2);
  const int split = (blockIdx.x >> 1) % p.t_splits;                    // which slice of the column sweep
  const int part = ((blockIdx.x >> 1) / p.t_splits) % p.n_parts;
  const int rt = ((blockIdx.x >> 1) / (p.t_splits * p.n_parts)) * 2 + cta_rank;   // padded row-tile index over all tensors
  const int i = rt / tiles_padded;
  const int tr = rt - i * tiles_padded;
  const bool tile_valid = tr < p.n_row_tiles;
  const int slab0 = part * 4;
  const int nh = (ks - slab0) < 4 ? (ks - slab0) : 4;                  // 64-wide embedding slabs in this part (even)
  const int nb = nh >> 1;                                              // slabs this CTA loads for GEMM2
  const int b_stages = 2;                                              // one per 64-column half of a step
  const int b_stage_bytes = nb * kSlabB;
  const int n_ct = p.n_col_tiles;
  const int T_all = p.gy * n_ct;
  const int T_per = (T_all + p.t_splits - 1) / p.t_splits;
  const int t_begin = split * T_per;
  const int T = min(T_all, t_begin + T_per) - t_begin;                 // >= 1: the host keeps t_splits <= T_all / 2
  const bool want_dx = p.dx != nullptr;

  uint8_t* sX = smem;
  uint8_t* sA = smem + ks * kSlabX;
  uint8_t* sB = sA + kSlotsA * kSlotA;
  Misc* misc = reinterpret_cast<Misc*>(sB + b_stages * b_stage_bytes);

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
      mbar_init(&misc->g_full[s], 2 * 8);   // one arrive per epilogue warp of both CTAs
    }
    mbar_init(&misc->dx_full, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY64);
    tma_prefetch_desc(&tmY128);
  }
  if (warp == 2) tmem_alloc_pair<512>(&misc->tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = misc->tmem_slot;

  // Single-thread roles: whole warp converged, elect.sync around the asynchronous instructions (a divergent `lane == 0`
  // branch makes the compiler wrap every UTMALDG / UTCHMMA / UTCBAR in an elect-and-branch loop).
  if (warp == 0) {
    // ---------------- TMA producer ----------------
    auto arm = [&](uint64_t* bar, uint32_t bytes_per_cta) {
      if (leader) mbar_expect_tx(bar, 2 * bytes_per_cta);
      else mbar_arrive_cluster(bar, 0);
    };
    if (elect_one()) {
      for (int s = 0; s < ks; ++s) tma_load_3d_pair(sX + s * kSlabX, &tmX, &misc->x_full, s * 64, tr * BM, i);
      arm(&misc->x_full, ks * kSlabX);
    }
    __syncwarp();
    uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
    auto load_g1 = [&](int t) {
      const int j = (t_begin + t) / n_ct, tc = (t_begin + t) - j * n_ct;
      for (int s = 0; s < ks; ++s) {
        mbar_wait(&misc->a_empty[sa], pa ^ 1);
        if (elect_one()) {
          tma_load_3d_pair(sA + sa * kSlotA, &tmY64, &misc->a_full[sa], s * 64, tc * BN + cta_rank * 64, j);
          arm(&misc->a_full[sa], kSlotA);
        }
        __syncwarp();
        if (++sa == kSlotsA) { sa = 0; pa ^= 1; }
      }
    };
    auto load_g2 = [&](int t) {     // GEMM2 operand in two 64-column halves: the first half's stage refills while the second runs
      const int j = (t_begin + t) / n_ct, tc = (t_begin + t) - j * n_ct;
      for (int half = 0; half < 2; ++half) {
        mbar_wait(&misc->b_empty[sb], pb ^ 1);
        if (elect_one()) {
          for (int s = 0; s < nb; ++s)
            tma_load_3d_pair(sB + sb * b_stage_bytes + s * kSlabB, &tmY64, &misc->b_full[sb],
                             (slab0 + cta_rank * nb + s) * 64, tc * BN + half * 64, j);
          arm(&misc->b_full[sb], b_stage_bytes);
        }
        __syncwarp();
        sb ^= 1;
        if (sb == 0) pb ^= 1;
      }
    };
    load_g1(0);
    for (int t = 0; t < T; ++t) {
      if (t + 1 < T) load_g1(t + 1);
      if (want_dx) load_g2(t);
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------- MMA issuer (leader CTA only; whole warp waits, one elected lane issues) ----------------
      mbar_wait(&misc->x_full, 0);
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      auto gemm1 = [&](int t) {
        const uint32_t buf = t & 1;
        for (int s = 0; s < ks; ++s) {
          mbar_wait(&misc->a_full[sa], pa);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sX + s * kSlabX);
          const uint32_t b_base = smem_u32(sA + sa * kSlotA);
          if (elect_one()) {
            if (!(p.dbg & 2)) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_ss_pair(tmem + buf * BN, make_smem_desc(a_base + kk * 32, 0, 1024), make_smem_desc(b_base + kk * 32, 0, 1024),
                             p.idesc_s, (s | kk) != 0);
            }
            tc_commit_pair(&misc->a_empty[sa], 3);
            if (s == ks - 1) tc_commit_pair(&misc->s_full[buf], 3);
          }
          __syncwarp();
          if (++sa == kSlotsA) { sa = 0; pa ^= 1; }
        }
      };
      gemm1(0);
      for (int t = 0; t < T; ++t) {
        if (t + 1 < T) gemm1(t + 1);
        // the epilogue of step t has consumed S_t (and, when dX is wanted, written G_t over it)
        const uint32_t buf = t & 1;
        mbar_wait(&misc->g_full[buf], (t >> 1) & 1);
        if (want_dx) {
          for (int half = 0; half < 2; ++half) {
            mbar_wait(&misc->b_full[sb], pb);
            tc_fence_after();
            const uint32_t b_base = smem_u32(sB + sb * b_stage_bytes);
            if (elect_one()) {
              if (!(p.dbg & 2)) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {   // K = 64 columns of this half, 16 per MMA = 8 packed TMEM columns of G
                  const uint32_t a_tmem = tmem + buf * BN + half * 64 + k * 8;
                  umma_ts_pair(tmem + 256, a_tmem, make_smem_desc(b_base + k * 2048, kSlabB, 1024), p.idesc_g, (t | half | k) != 0);
                }
              }
              tc_commit_pair(&misc->b_empty[sb], 3);
            }
            __syncwarp();
            sb ^= 1;
            if (sb == 0) pb ^= 1;
          }
        }
      }
      if (elect_one()) tc_commit_pair(&misc->dx_full, 3);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ---------------- epilogue ----------------
    const uint32_t ew = warp - 4;
    const uint32_t q = warp & 3;
    const uint32_t h = ew >> 2;           // which 64-column half of the step this warp converts
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

    // dscale mix relative to the G mix: proportional (ds = ratio * <G, raw>, every shipped mode but local_loss with
    // gather_with_grad) or rows-only (s_col == 0); anything else takes the two-exp path.
    const bool ds_prop = fabsf(p.a_row * p.s_col - p.a_col * p.s_row) <= 1e-12f && (a_sum != 0.f);
    const float ds_ratio = ds_prop ? s_sum / a_sum : 0.f;
    const bool ds_rows = !ds_prop && p.s_col == 0.f;
    const bool fast_ok = (ds_prop || ds_rows) && !(p.dbg & 32);
    float* kbuf = misc->kappa[ew];

    for (int t = 0; t < T; ++t) {
      const int j = (t_begin + t) / n_ct, tc = (t_begin + t) - j * n_ct;
      const int pair = i * p.gy + j;
      const uint32_t buf = t & 1;
      const float lr = row_valid ? __ldg(p.row_lse2 + static_cast<size_t>(pair) * p.n_rows + row) : INFINITY;
      const float* lc_ptr = p.col_lse2 + static_cast<size_t>(pair) * p.n_cols;
      const int colw = tc * BN + h * 64;                  // first of this warp's 64 columns
      const bool full = colw + 64 <= p.n_cols;

      // ---- one-exp path: G = 2^(S2 - o) * (a_row 2^(o - lse_row) + a_col 2^(o - lse_col)) with a per-warp offset o.
      // Valid when the log-sum-exps this warp touches span < 200 (log2 units): then every factor stays a normal
      // fp32 number wherever the product matters.  Otherwise (or in a ragged last tile) the two-exp path runs.
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
        fast = (hi - lo) <= 200.f;       // false for NaN / inf as well
        if (fast) {
          o = 0.5f * (hi + lo);
          rho = p.a_row * ex2(o - lr);   // 0 for rows past the batch (lr = +inf)
          rho_s = p.s_row * ex2(o - lr);
          __syncwarp();
          kbuf[lane] = p.a_col * ex2(o - lc0);
          kbuf[32 + lane] = p.a_col * ex2(o - lc1);
          __syncwarp();
        }
      }

      mbar_wait(&misc->s_full[buf], (t >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int chunk = 0; chunk < 2; ++chunk) {
        const int col0 = colw + chunk * 32;
        uint32_t packed[16];
        if (p.dbg & 1) {
#pragma unroll
          for (int k = 0; k < 16; ++k) packed[k] = 0;
        } else if (fast) {
          uint32_t v[32];
          tmem_ld32(tmem + lane_base + buf * BN + h * 64 + chunk * 32, v);
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
            for (int k = 0; k < 32; ++k) {
              const float raw = __uint_as_float(v[k]);
              g[k] = ex2(fmaf(raw, k2, neg_o)) * (rho + kap[k]);
              acc = fmaf(g[k], raw, acc);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const float raw = __uint_as_float(v[k]);
              const float e = ex2(fmaf(raw, k2, neg_o));
              g[k] = e * (rho + kap[k]);
              acc = fmaf(e, raw, acc);
            }
          }
          if (label >= col0 && label < col0 + 32) {         // the positive of this row sits in this chunk
            const int idx = label - col0;
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (k == idx) {
                g[k] -= a_sum;
                acc -= (ds_prop ? a_sum : (rho_s != 0.f ? p.s_row / rho_s : 0.f)) * __uint_as_float(v[k]);
              }
          }
          if (row_valid) ds_acc += ds_prop ? ds_ratio * acc : rho_s * acc;
#pragma unroll
          for (int k = 0; k < 16; ++k) packed[k] = pack2(g[2 * k], g[2 * k + 1], fmt);
        } else {
          uint32_t v[32];
          tmem_ld32(tmem + lane_base + buf * BN + h * 64 + chunk * 32, v);
          float lcv[32];
          if (col0 + 32 <= p.n_cols && (p.n_cols & 3) == 0) {
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
              const float4 f = __ldg(reinterpret_cast<const float4*>(lc_ptr + col0) + k4);
              lcv[4 * k4 + 0] = f.x; lcv[4 * k4 + 1] = f.y; lcv[4 * k4 + 2] = f.z; lcv[4 * k4 + 3] = f.w;
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) lcv[k] = (col0 + k < p.n_cols) ? __ldg(lc_ptr + col0 + k) : INFINITY;
          }
          tmem_ld_wait();
          float g[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const int col = col0 + k;
            const float raw = __uint_as_float(v[k]);
            const float tt = raw * k2;
            const float pr = ex2(tt - lr);
            const float pc = ex2(tt - lcv[k]);
            float gg = p.a_row * pr + p.a_col * pc;
            float dd = p.s_row * pr + p.s_col * pc;
            if (col == label) {
              gg -= a_sum;
              dd -= s_sum;
            }
            const bool ok = (col < p.n_cols) && row_valid;
            g[k] = ok ? gg : 0.f;
            ds_acc += ok ? dd * raw : 0.f;
          }
#pragma unroll
          for (int k = 0; k < 16; ++k) packed[k] = pack2(g[2 * k], g[2 * k + 1], fmt);
        }
        // G goes back into this warp's own (already read) S columns: 32 S columns -> 16 packed columns
        if (want_dx) tmem_st16(tmem + lane_base + buf * BN + h * 64 + chunk * 16, packed);
      }
      if (want_dx) tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&misc->g_full[buf]);
        else mbar_arrive_cluster(&misc->g_full[buf], 0);
      }
    }

    if (p.dscale_part != nullptr && part == 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ds_acc += __shfl_xor_sync(0xffffffffu, ds_acc, o);
      if (lane == 0) misc->red[ew] = ds_acc;
      named_bar_sync(1, kEpiThreads);
      if (ew == 0 && lane == 0 && tile_valid) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += misc->red[w];
        p.dscale_part[(i * p.n_row_tiles + tr) * p.t_splits + split] = s;
      }
    }

    if (want_dx) {
      mbar_wait(&misc->dx_full, 0);
      tc_fence_after();
      const float coef = __ldg(p.upstream) * p.weight * scale;
      const int dim = ks * 64;
      const int width = nh * 64;
      const int c_begin = h * (width / 2), c_end = c_begin + width / 2;
      for (int c = c_begin; c < c_end; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_base + 256 + c, v);
        tmem_ld_wait();
        if (row_valid) {
          const size_t off = (static_cast<size_t>(i) * p.n_rows + row) * dim + slab0 * 64 + c;
          if (p.t_splits > 1) {
            float* dst = p.dx32 + off;       // several clusters sweep different column ranges of this row block
#pragma unroll
            for (int k = 0; k < 32; ++k) atomicAdd(dst + k, __uint_as_float(v[k]) * coef);
          } else {
            uint32_t o[16];
#pragma unroll
            for (int k = 0; k < 16; ++k)
              o[k] = pack2(__uint_as_float(v[2 * k]) * coef, __uint_as_float(v[2 * k + 1]) * coef, fmt);
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.dx) + off);
#pragma unroll
            for (int k = 0; k < 4; ++k) dst[k] = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
          }
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair<512>(tmem);
}

cudaError_t launch_infonce_bwd_pair(const CUtensorMap& tmX, const CUtensorMap& tmY64, const CUtensorMap& tmY128, const BwdParams& p,
                                    cudaStream_t stream) {
  const int nh_max = p.ks < 4 ? p.ks : 4;
  const int b_stages = 2;
  const int smem_bytes = p.ks * kSlabX + kSlotsA * kSlotA + b_stages * (nh_max / 2) * kSlabB + kSmemMisc;
  if (smem_bytes > 232448) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(infonce_bwd_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.gx * ((p.n_row_tiles + 1) / 2) * p.n_parts * p.t_splits * 2);
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
  return cudaLaunchKernelEx(&cfg, infonce_bwd_pair_kernel, tmX, tmY64, tmY128, p);
}

}  // namespace cb
