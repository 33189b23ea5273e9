// InfoNCE backward for the row side of a stack:  dX_i = coef * sum_j G_ij Y_j,  dscale = coef' * <G, X Y^T>,
// with G = a_row * softmax_rows(S) + a_col * softmax_cols(S) - (a_row + a_col) * I rebuilt on the fly
// from the forward's log-sum-exps (flash-style recompute: S and G never reach HBM).
//
// One CTA = (row tensor i, 128-row tile, 256-wide part of the embedding dim).  Per 128-column step t:
//   GEMM1  S_t   = X_tile (smem, resident) . Y_tile^T        tcgen05.mma M128 N128 K16, K-major B
//   epi    G_t   = mix(exp2(S - lse_row), exp2(S - lse_col)) -> 16-bit -> smem (128B-swizzled, K-major)
//   GEMM2  dX   += G_t (smem A) . Y_tile[:, part]            tcgen05.mma M128 N64 K16 x4 slabs, MN-major B
// TMEM: S double-buffered in columns [0,256), dX accumulator in [256, 256 + part width).
// The same TMA-written Y slab bytes serve as K-major B (GEMM1) and MN-major B (GEMM2).
// MMA issue order G1(0) G1(1) G2(0) G1(2) G2(1) ... so the softmax epilogue of step t overlaps GEMM1 of t+1.
//
// Reference: autograd of src/open_clip/loss.py:110-138 (logits, two cross-entropies); the mode
// scalings of loss.py:21-65 enter through a_row/a_col/s_row/s_col/weight (see cosmos_b200/infonce.py).
#include "common.cuh"
#include "infonce.h"
#include "internal.h"

namespace cb {

namespace {

constexpr int BM = kFwdBM, BN = kBwdBN;
constexpr int kStages = 4;
constexpr int kSlab = 128 * 64 * 2;   // 16 KB: 128 rows (or columns) x 64 elements
constexpr int kSmemX = 8 * kSlab;     // 128 KB resident X tile
constexpr int kSmemG = 2 * kSlab;     // 32 KB  G tile: 128 rows x 128 cols
constexpr int kSmemY = kStages * kSlab;
constexpr int kSmemMisc = 2048;
constexpr int kThreads = 384;
constexpr int kEpiThreads = 256;

struct Misc {
  uint64_t x_full;
  uint64_t y_full[kStages];
  uint64_t y_empty[kStages];
  uint64_t s_full[2];
  uint64_t s_empty[2];
  uint64_t g_full;
  uint64_t g_empty;
  uint64_t dx_full;
  uint32_t tmem_slot;
  uint32_t pad[3];
  float red[8];
};
static_assert(sizeof(Misc) <= kSmemMisc, "misc smem");

}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
infonce_bwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sG = smem + kSmemX;
  uint8_t* sY = smem + kSmemX + kSmemG;
  Misc* misc = reinterpret_cast<Misc*>(smem + kSmemX + kSmemG + kSmemY);

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // work item: part fastest so the CTAs sharing an X tile run side by side
  const int part = blockIdx.x % p.n_parts;
  const int rt = blockIdx.x / p.n_parts;
  const int i = rt / p.n_row_tiles;
  const int tr = rt - i * p.n_row_tiles;
  const int ks = p.ks;
  const int slab0 = part * 4;                              // first embedding slab of this part
  const int nh = (ks - slab0) < 4 ? (ks - slab0) : 4;      // slabs in this part
  const int n_ct = p.n_col_tiles;
  const int T = p.gy * n_ct;                               // steps: all column tensors, all tiles
  const bool want_dx = p.dx != nullptr;

  if (tid == 0) {
    mbar_init(&misc->x_full, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&misc->y_full[s], 1);
      mbar_init(&misc->y_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&misc->s_full[s], 1);
      mbar_init(&misc->s_empty[s], kEpiThreads);
    }
    mbar_init(&misc->g_full, kEpiThreads);
    mbar_init(&misc->g_empty, 1);
    mbar_init(&misc->dx_full, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 2) tmem_alloc<512>(&misc->tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc->tmem_slot;

  // Single-thread roles: whole warp converged, elect.sync around the asynchronous instructions (a divergent `lane == 0`
  // branch makes the compiler wrap every UTMALDG / UTCHMMA / UTCBAR in an elect-and-branch loop).
  if (warp == 0) {
    // ---------------- TMA producer: slabs in the order the MMA warp consumes them ----------------
    if (elect_one()) {
      mbar_expect_tx(&misc->x_full, ks * kSlab);
      for (int s = 0; s < ks; ++s) tma_load_3d(sX + s * kSlab, &tmX, &misc->x_full, s * 64, tr * BM, i);
    }
    __syncwarp();
    uint32_t stage = 0, phase = 0;
    auto load = [&](int t, int s) {
      const int j = t / n_ct, tc = t - j * n_ct;
      mbar_wait(&misc->y_empty[stage], phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&misc->y_full[stage], kSlab);
        tma_load_3d(sY + stage * kSlab, &tmY, &misc->y_full[stage], s * 64, tc * BN, j);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    };
    for (int s = 0; s < ks; ++s) load(0, s);
    for (int t = 0; t < T; ++t) {
      if (t + 1 < T)
        for (int s = 0; s < ks; ++s) load(t + 1, s);
      if (want_dx)
        for (int s = 0; s < nh; ++s) load(t, slab0 + s);
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (whole warp waits, one elected lane issues) ----------------
    mbar_wait(&misc->x_full, 0);
    uint32_t stage = 0, phase = 0;
    auto gemm1 = [&](int t) {
      const uint32_t sb = t & 1;
      mbar_wait(&misc->s_empty[sb], ((t >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int s = 0; s < ks; ++s) {
        mbar_wait(&misc->y_full[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sX + s * kSlab);
        const uint32_t b_base = smem_u32(sY + stage * kSlab);
        if (elect_one()) {
          if (!(p.dbg & 2)) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_ss(tmem + sb * BN, make_smem_desc(a_base + kk * 32, 0, 1024), make_smem_desc(b_base + kk * 32, 0, 1024),
                      p.idesc_s, (s | kk) != 0);
          }
          tc_commit(&misc->y_empty[stage]);
          if (s == ks - 1) tc_commit(&misc->s_full[sb]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    };
    gemm1(0);
    for (int t = 0; t < T; ++t) {
      if (t + 1 < T) gemm1(t + 1);
      if (want_dx) {
        mbar_wait(&misc->g_full, t & 1);
        tc_fence_after();
        for (int s = 0; s < nh; ++s) {
          mbar_wait(&misc->y_full[stage], phase);
          tc_fence_after();
          const uint32_t b_base = smem_u32(sY + stage * kSlab);
          if (elect_one()) {
            if (!(p.dbg & 2)) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {   // 128 columns of this step = K of GEMM2, 16 per MMA
                const uint32_t a_addr = smem_u32(sG) + (k >> 2) * kSlab + (k & 3) * 32;
                umma_ss(tmem + 256 + s * 64, make_smem_desc(a_addr, 0, 1024), make_smem_desc(b_base + k * 2048, kSlab, 1024),
                        p.idesc_g, (t | k) != 0);
              }
            }
            tc_commit(&misc->y_empty[stage]);
            if (s == nh - 1) tc_commit(&misc->g_empty);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    if (elect_one()) tc_commit(&misc->dx_full);
    __syncwarp();
  } else if (warp >= 4) {
    // ---------------- epilogue ----------------
    const uint32_t ew = warp - 4;
    const uint32_t q = warp & 3;
    const uint32_t h = ew >> 2;           // which 64-column half of the step this warp converts
    const int r_t = q * 32 + lane;        // row inside the tile == TMEM lane
    const int row = tr * BM + r_t;
    const bool row_valid = row < p.n_rows;
    const int label = p.label_offset + row;
    const float scale = __ldg(p.scale);
    const float k2 = scale * kLog2e;
    const float a_sum = p.a_row + p.a_col, s_sum = p.s_row + p.s_col;
    const int fmt = p.dtype == COSMOS_DTYPE_BF16 ? 1 : 0;
    float ds_acc = 0.f;

    for (int t = 0; t < T; ++t) {
      const int j = t / n_ct, tc = t - j * n_ct;
      const int pair = i * p.gy + j;
      const uint32_t sb = t & 1;
      const float lr = row_valid ? __ldg(p.row_lse2 + static_cast<size_t>(pair) * p.n_rows + row) : INFINITY;
      const float* lc_ptr = p.col_lse2 + static_cast<size_t>(pair) * p.n_cols;

      mbar_wait(&misc->s_full[sb], (t >> 1) & 1);
      tc_fence_after();
      uint32_t packed[2][16];
      if (p.dbg & 1) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int k = 0; k < 16; ++k) packed[c][k] = 0;
      }
#pragma unroll
      for (int chunk = 0; chunk < 2; ++chunk) {
        if (p.dbg & 1) break;
        const int col0 = tc * BN + h * 64 + chunk * 32;
        uint32_t v[32];
        tmem_ld32(tmem + ((q * 32u) << 16) + sb * BN + h * 64 + chunk * 32, v);
        tmem_ld_wait();
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
        float g[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const int col = col0 + k;
          const bool cv = col < p.n_cols;
          const float lc = lcv[k];
          const float raw = __uint_as_float(v[k]);
          const float tt = raw * k2;
          const float pr = ex2(tt - lr);
          const float pc = ex2(tt - lc);
          float gg = p.a_row * pr + p.a_col * pc;
          float dd = p.s_row * pr + p.s_col * pc;
          if (col == label) {
            gg -= a_sum;
            dd -= s_sum;
          }
          const bool ok = cv && row_valid;
          g[k] = ok ? gg : 0.f;
          ds_acc += ok ? dd * raw : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) packed[chunk][k] = pack2(g[2 * k], g[2 * k + 1], fmt);
      }
      // S buffer may be overwritten by GEMM1 of step t+2
      tc_fence_before();
      mbar_arrive(&misc->s_empty[sb]);

      if (want_dx) {
        // G tile: wait until GEMM2 of the previous step has read it, then publish this step's
        mbar_wait(&misc->g_empty, (t & 1) ^ 1);
        uint8_t* g_row = sG + h * kSlab + r_t * 128;
#pragma unroll
        for (int chunk = 0; chunk < 2; ++chunk) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int jj = chunk * 4 + c;  // 16-byte chunk index inside the 128-byte row
            *reinterpret_cast<uint4*>(g_row + ((jj ^ (r_t & 7)) << 4)) =
                make_uint4(packed[chunk][c * 4 + 0], packed[chunk][c * 4 + 1], packed[chunk][c * 4 + 2], packed[chunk][c * 4 + 3]);
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(&misc->g_full);
      }
    }

    // dscale partial of this CTA (only the first embedding part reports it)
    if (p.dscale_part != nullptr && part == 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ds_acc += __shfl_xor_sync(0xffffffffu, ds_acc, o);
      if (lane == 0) misc->red[ew] = ds_acc;
      named_bar_sync(1, kEpiThreads);
      if (ew == 0 && lane == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += misc->red[w];
        p.dscale_part[rt] = s;
      }
    }

    if (want_dx) {
      // drain dX: this warp's 32 rows, its half of the part's columns
      mbar_wait(&misc->dx_full, 0);
      tc_fence_after();
      const float coef = __ldg(p.upstream) * p.weight * scale;
      const int dim = ks * 64;
      const int width = nh * 64;
      const int c_begin = h * (width / 2), c_end = c_begin + width / 2;
      for (int c = c_begin; c < c_end; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((q * 32u) << 16) + 256 + c, v);
        tmem_ld_wait();
        if (row_valid) {
          uint32_t o[16];
#pragma unroll
          for (int k = 0; k < 16; ++k)
            o[k] = pack2(__uint_as_float(v[2 * k]) * coef, __uint_as_float(v[2 * k + 1]) * coef, fmt);
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.dx) +
                                                (static_cast<size_t>(i) * p.n_rows + row) * dim + slab0 * 64 + c);
#pragma unroll
          for (int k = 0; k < 4; ++k) dst[k] = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

cudaError_t launch_infonce_bwd(const CUtensorMap& tmX, const CUtensorMap& tmY, const BwdParams& p, cudaStream_t stream) {
  const int smem_bytes = kSmemX + kSmemG + kSmemY + kSmemMisc + 1024;
  static_assert(kSmemX + kSmemG + kSmemY + kSmemMisc + 1024 <= 232448, "shared memory budget");
  cudaError_t e = cudaFuncSetAttribute(infonce_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  const int grid = p.gx * p.n_row_tiles * p.n_parts;
  infonce_bwd_kernel<<<grid, kThreads, smem_bytes, stream>>>(tmX, tmY, p);
  return cudaGetLastError();
}

}  // namespace cb
