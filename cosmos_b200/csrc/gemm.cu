// Generic tcgen05 GEMM used by the cross-attention pooler (linear layers, their input and weight
// gradients, and - batched over samples or heads - the folded attention products):
//   D[M,N] (+)= alpha * opA[M,K] * opB[N,K]^T (+ bias[N]),   optionally for `batch` independent problems of one shape
//   opA stored [M,K] row-major (K-major operand) or [K,M] row-major (MN-major operand); same for opB.
//   16-bit inputs (bf16 / fp16), fp32 accumulation in TMEM, fp32 or 16-bit output.
// Tile 128 x 128 x 64, 6-stage TMA ring, one CTA per output tile and K split (gridDim.z); K splits
// accumulate with fp32 red.global.add into a zero-initialised D (weight gradients contract over the very
// long row dimension, so the output has few tiles and needs the split for parallelism).
// Warp roles as in the InfoNCE kernels: warp 0 TMA, warp 1 MMA issue, warp 2 TMEM, warps 4-11 epilogue.
#include "common.cuh"
#include "internal.h"
#include "tma_host.h"

namespace cb {

namespace {

constexpr int BM = 128, BK = 64;
constexpr int kABytes = 128 * 64 * 2;    // 16 KB of A per stage
constexpr int kEpiWarps = 8;             // two per TMEM lane quarter, each draining half of the tile's columns
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr int kMaxStages = 6;

struct Misc {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_slot;
};

}  // namespace

// Persistent: CTA c works on items c, c + gridDim.x, ...; item = (output tile, K split).  Two accumulator stages in
// TMEM let the epilogue of one tile overlap the MMAs of the next.
template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmA2,
            const __grid_constant__ CUtensorMap tmB2, GemmParams p) {
  constexpr int kStages = BN == 256 ? 4 : 6;
  constexpr int kBBytes = BN * 64 * 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kABytes;
  Misc* misc = reinterpret_cast<Misc*>(smem + kStages * (kABytes + kBBytes));

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_m = (p.M + BM - 1) / BM;
  // D = A B^T + A2 B2^T: the K slabs of an optional second operand pair (K2 > 0; same storage orders) follow those of the first
  const int k_slabs1 = (p.K + BK - 1) / BK;
  const int k_slabs = k_slabs1 + (p.K2 + BK - 1) / BK;
  const int per_split = (k_slabs + p.splits - 1) / p.splits;
  const int per_batch = tiles_m * tiles_n * p.splits;
  const int n_items = per_batch * p.batch * p.batch2;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&misc->full[s], 1);
      mbar_init(&misc->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&misc->acc_full[s], 1);
      mbar_init(&misc->acc_empty[s], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.K2 > 0) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
  }
  if (warp == 2) tmem_alloc<2 * BN>(&misc->tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc->tmem_slot;

  auto decode = [&](int item, int& bt, int& bi, int& m0, int& n0, int& ks0, int& ks1) {
    const int t = item / per_batch;          // batch slowest: CTAs running together work on neighbouring problems
    item -= t * per_batch;
    bt = t / p.batch2;                       // outer (samples), inner (heads)
    bi = t - bt * p.batch2;
    const int split = item % p.splits;
    const int tile = item / p.splits;
    n0 = (tile % tiles_n) * BN;     // n fastest: CTAs running together share the A rows
    m0 = (tile / tiles_n) * BM;
    ks0 = split * per_split;
    ks1 = min(k_slabs, ks0 + per_split);
  };

  // Single-thread roles: whole warp converged, elect.sync around the asynchronous instructions (a divergent `lane == 0`
  // branch makes the compiler wrap every UTMALDG / UTCHMMA / UTCBAR in an elect-and-branch loop).
  if (warp == 0) {
    uint32_t stage = 0, phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int bt, bi, m0, n0, ks0, ks1;
      decode(item, bt, bi, m0, n0, ks0, ks1);
      for (int s = ks0; s < ks1; ++s) {
        mbar_wait(&misc->empty[stage], phase ^ 1);
        uint8_t* a = sA + stage * kABytes;
        uint8_t* b = sB + stage * kBBytes;
        if (elect_one()) {
        const bool second = s >= k_slabs1;
        const CUtensorMap* mA = second ? &tmA2 : &tmA;
        const CUtensorMap* mB = second ? &tmB2 : &tmB;
        const int k0 = (second ? s - k_slabs1 : s) * BK;
        mbar_expect_tx(&misc->full[stage], kABytes + kBBytes);
        if (p.a_kmajor) {
          tma_load_4d(a, mA, &misc->full[stage], k0, m0, bi, bt);           // box 64 k x 128 rows
        } else {
          tma_load_4d(a, mA, &misc->full[stage], m0, k0, bi, bt);           // box 64 m x 64 k-rows per 64-wide chunk
          tma_load_4d(a + 8192, mA, &misc->full[stage], m0 + 64, k0, bi, bt);
        }
        if (p.b_kmajor) {
#pragma unroll
          for (int c = 0; c < BN / 128; ++c) tma_load_4d(b + c * 16384, mB, &misc->full[stage], k0, n0 + c * 128, bi, bt);
        } else {
#pragma unroll
          for (int c = 0; c < BN / 64; ++c) tma_load_4d(b + c * 8192, mB, &misc->full[stage], n0 + c * 64, k0, bi, bt);
        }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    uint32_t stage = 0, phase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      int bt, bi, m0, n0, ks0, ks1;
      decode(item, bt, bi, m0, n0, ks0, ks1);
      const uint32_t as = it & 1;
      mbar_wait(&misc->acc_empty[as], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int s = ks0; s < ks1; ++s) {
        mbar_wait(&misc->full[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + stage * kABytes);
        const uint32_t b_base = smem_u32(sB + stage * kBBytes);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t da = p.a_kmajor ? make_smem_desc(a_base + kk * 32, 0, 1024) : make_smem_desc(a_base + kk * 2048, 8192, 1024);
            const uint64_t db = p.b_kmajor ? make_smem_desc(b_base + kk * 32, 0, 1024) : make_smem_desc(b_base + kk * 2048, 8192, 1024);
            umma_ss(tmem + as * BN, da, db, p.idesc, (s > ks0 || kk > 0) ? 1u : 0u);
          }
          tc_commit(&misc->empty[stage]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) tc_commit(&misc->acc_full[as]);     // also fires for an empty K range: the epilogue then sees no valid data, see below
      __syncwarp();
    }
  } else if (warp >= 4) {
    const uint32_t q = warp & 3;
    // The folded attention's GEMMs have one or two K slabs per tile: four epilogue warps then set the pace (a 128 x 256 tile
    // is 64 KB of output per 4 MMAs).  Eight warps: warp / 4 selects the half of the tile's columns.
    const int c_begin = static_cast<int>((warp - 4) >> 2) * (BN / 2), c_end = c_begin + BN / 2;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      int bt, bi, m0, n0, ks0, ks1;
      decode(item, bt, bi, m0, n0, ks0, ks1);
      const uint32_t as = it & 1;
      const int row = m0 + q * 32 + lane;
      mbar_wait(&misc->acc_full[as], (it >> 1) & 1);
      tc_fence_after();
      const bool add_bias = p.bias != nullptr && (item % p.splits) == 0;
      const float* bias = p.bias != nullptr ? p.bias + static_cast<size_t>(bt) * p.sbias + static_cast<size_t>(bi) * p.sbias2 : nullptr;
      const bool has_k = ks1 > ks0;
      for (int c = c_begin; c < c_end; c += 32) {
        if (n0 + c >= p.N) break;
        uint32_t v[32];
        tmem_ld32(tmem + ((q * 32u) << 16) + as * BN + c, v);
        tmem_ld_wait();
        if (row < p.M) {
          float o[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const int col = n0 + c + k;
            o[k] = (has_k ? __uint_as_float(v[k]) * p.alpha : 0.f) + ((add_bias && col < p.N) ? __ldg(bias + col) : 0.f);
          }
          const size_t off = static_cast<size_t>(bt) * p.sd + static_cast<size_t>(bi) * p.sd2 + static_cast<size_t>(row) * p.ldd + n0 + c;
          if (p.accumulate) {             // D += ...: the caller's second product into the same output (splits == 1)
            if (p.out_dtype == COSMOS_DTYPE_F32) {
              const float* src = reinterpret_cast<const float*>(p.d) + off;
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if (n0 + c + k < p.N) o[k] += src[k];
            } else if (n0 + c + 32 <= p.N && (p.ldd & 7) == 0 && (p.sd & 7) == 0 && (p.sd2 & 7) == 0) {
              const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.d) + off);
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const uint4 u = src[k4];
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (p.out_dtype == COSMOS_DTYPE_BF16) {
                    o[k4 * 8 + 2 * j] += __uint_as_float(w[j] << 16);
                    o[k4 * 8 + 2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
                  } else {
                    o[k4 * 8 + 2 * j] += __half2float(__ushort_as_half(static_cast<unsigned short>(w[j] & 0xffff)));
                    o[k4 * 8 + 2 * j + 1] += __half2float(__ushort_as_half(static_cast<unsigned short>(w[j] >> 16)));
                  }
                }
              }
            } else {
              const uint16_t* src = reinterpret_cast<const uint16_t*>(p.d) + off;
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if (n0 + c + k < p.N)
                  o[k] += p.out_dtype == COSMOS_DTYPE_BF16 ? __uint_as_float(static_cast<uint32_t>(src[k]) << 16)
                                                           : __half2float(__ushort_as_half(src[k]));
            }
          }
          if (p.splits > 1) {
            float* dst = reinterpret_cast<float*>(p.d) + off;
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (n0 + c + k < p.N) atomicAdd(dst + k, o[k]);
          } else if (p.out_dtype == COSMOS_DTYPE_F32) {
            float* dst = reinterpret_cast<float*>(p.d) + off;
            if (n0 + c + 32 <= p.N && (p.ldd & 3) == 0 && (p.sd & 3) == 0 && (p.sd2 & 3) == 0) {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                reinterpret_cast<float4*>(dst)[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
            } else {
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if (n0 + c + k < p.N) dst[k] = o[k];
            }
          } else {
            const int fmt = p.out_dtype == COSMOS_DTYPE_BF16 ? 1 : 0;
            uint16_t* dst = reinterpret_cast<uint16_t*>(p.d) + off;
            if (n0 + c + 32 <= p.N && (p.ldd & 7) == 0 && (p.sd & 7) == 0 && (p.sd2 & 7) == 0) {
              uint32_t w[16];
#pragma unroll
              for (int k = 0; k < 16; ++k) w[k] = pack2(o[2 * k], o[2 * k + 1], fmt);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                reinterpret_cast<uint4*>(dst)[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
            } else {
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if (n0 + c + k < p.N) {
                  const uint32_t w = pack2(o[k], 0.f, fmt);
                  dst[k] = static_cast<uint16_t>(w & 0xffff);
                }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&misc->acc_empty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<2 * BN>(tmem);
}

// 2-D operand map: K-major [rows, K] -> box {64 k, 128 rows}; MN-major [K, rows] -> box {64 rows, 64 k}.
static int make_operand_map(CUtensorMap* map, const void* ptr, int is_bf16, int kmajor, int64_t rows, int64_t K, int64_t ld,
                            int batch, int64_t bstride, int batch_in, int64_t bstride_in) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return -1;
  // rank 4: {contiguous dim, rows, inner batch, outer batch}
  cuuint64_t gdim[4], gstride[3];
  cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
  if (kmajor) {
    gdim[0] = K; gdim[1] = rows;
    box[0] = 64; box[1] = 128;
  } else {
    gdim[0] = rows; gdim[1] = K;
    box[0] = 64; box[1] = 64;
  }
  gdim[2] = batch_in; gdim[3] = batch;
  box[2] = 1; box[3] = 1;
  // (the batch stride may be smaller than the row stride - per-head column blocks of one matrix: strides only have to be
  //  multiples of 16 bytes; every dimension is bounded on its own, so tiles past M, N or K of one problem read zeros and
  //  never its neighbour's rows)
  gstride[0] = static_cast<cuuint64_t>(ld) * 2;
  const cuuint64_t whole = gstride[0] * gdim[1];
  gstride[1] = batch_in > 1 ? static_cast<cuuint64_t>(bstride_in) * 2 : whole;
  gstride[2] = batch > 1 ? static_cast<cuuint64_t>(bstride) * 2 : whole;
  CUresult r = fn(map, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), gdim,
                  gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

template <int BN>
static cudaError_t launch_gemm_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmA2, const CUtensorMap& tmB2,
                                  GemmParams p, int bf, int sm_count, cudaStream_t stream) {
  constexpr int kStages = BN == 256 ? 4 : 6;
  p.idesc = make_idesc(bf, p.a_kmajor ? 0 : 1, p.b_kmajor ? 0 : 1, BM, BN);
  const int smem_bytes = kStages * (kABytes + BN * 64 * 2) + 1024;
  cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  const int items = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN) * p.splits * p.batch * p.batch2;
  const int grid = items < sm_count ? items : sm_count;
  gemm_kernel<BN><<<grid, kThreads, smem_bytes, stream>>>(tmA, tmB, tmA2, tmB2, p);
  return cudaGetLastError();
}

int launch_gemm(const GemmArgs& a, int sm_count, cudaStream_t stream, cudaError_t* err) {
  *err = cudaSuccess;
  CUtensorMap tmA, tmB;
  const int bf = a.in_dtype == COSMOS_DTYPE_BF16;
  const bool wide = a.N >= 256;       // 128 x 256 tiles when the output is wide enough
  int r1 = make_operand_map(&tmA, a.a, bf, a.a_kmajor, a.M, a.K, a.lda, a.batch, a.sa, a.batch_in, a.sa_in);
  int r2 = make_operand_map(&tmB, a.b, bf, a.b_kmajor, a.N, a.K, a.ldb, a.batch, a.sb, a.batch_in, a.sb_in);
  if (r1 != 0 || r2 != 0) return 100000 + (r1 != 0 ? r1 : r2);
  CUtensorMap tmA2 = tmA, tmB2 = tmB;
  if (a.K2 > 0) {
    r1 = make_operand_map(&tmA2, a.a2, bf, a.a_kmajor, a.M, a.K2, a.lda2, a.batch, a.sa2, a.batch_in, a.sa2_in);
    r2 = make_operand_map(&tmB2, a.b2, bf, a.b_kmajor, a.N, a.K2, a.ldb2, a.batch, a.sb2, a.batch_in, a.sb2_in);
    if (r1 != 0 || r2 != 0) return 100000 + (r1 != 0 ? r1 : r2);
  }
  GemmParams p;
  p.M = a.M; p.N = a.N; p.K = a.K; p.ldd = static_cast<int>(a.ldd);
  p.a_kmajor = a.a_kmajor; p.b_kmajor = a.b_kmajor; p.splits = a.splits;
  p.out_dtype = a.splits > 1 ? COSMOS_DTYPE_F32 : a.out_dtype;
  p.alpha = a.alpha; p.bias = a.bias; p.d = a.d;
  p.batch = a.batch; p.accumulate = a.accumulate; p.sd = a.sd; p.sbias = a.sbias; p.K2 = a.K2;
  p.batch2 = a.batch_in; p.sd2 = a.sd_in; p.sbias2 = a.sbias_in;
  p.idesc = 0;
  *err = wide ? launch_gemm_bn<256>(tmA, tmB, tmA2, tmB2, p, bf, sm_count, stream)
              : launch_gemm_bn<128>(tmA, tmB, tmA2, tmB2, p, bf, sm_count, stream);
  return *err == cudaSuccess ? 0 : -1;
}

}  // namespace cb
