// HBM-bound pieces of the cross-attention pooler (everything that is not a GEMM):
// LayerNorm fwd/bwd, the few-queries attention core fwd/bwd, residual add + L2 normalise fwd/bwd,
// column sums for bias gradients.  Reference: src/open_clip/transformer.py:24-30, 210-230 and
// src/open_clip/model.py:378-387.  The contractions (in-proj, out-proj, their input and weight gradients)
// run on tcgen05 through gemm.cu.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "internal.h"

namespace cb {

namespace {

__device__ __forceinline__ float ld_elem(const void* p, int dtype, size_t i) {
  if (dtype == COSMOS_DTYPE_F32) return reinterpret_cast<const float*>(p)[i];
  if (dtype == COSMOS_DTYPE_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return __half2float(reinterpret_cast<const __half*>(p)[i]);
}
__device__ __forceinline__ void st_elem(void* p, int dtype, size_t i, float v) {
  if (dtype == COSMOS_DTYPE_F32) reinterpret_cast<float*>(p)[i] = v;
  else if (dtype == COSMOS_DTYPE_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}


// One row in registers.  Two element-to-lane maps.  Scalar: column c = k * 32 + lane.  Vector (kVec, dim a multiple of 256):
// register k = 8 j + e holds column (32 j + lane) * 8 + e - groups of 8 consecutive elements, so a 16-bit row moves as one
// 16-byte vector per lane and group and every warp access is 512 contiguous bytes.  Measured on the [200704, 512] token
// LayerNorm (tools/ln_ab.py, same box, L2 flushed): forward 171 -> 149 us with the vector map, backward 297 us against an
// erratic 240 - 680 us with the scalar map; four rows per warp in flight made both slower (197 / 587 us: fewer warps per SM
// cost more than the longer bursts gained).  row_col() gives the column of register k (or -1).  The dtype switch
// sits OUTSIDE the unrolled loops so that the loads of a row are independent instructions the memory system can overlap.
template <bool kVec>
__device__ __forceinline__ int row_per(int dim) { return (kVec && (dim & 255) == 0) ? dim >> 5 : 0; }
template <bool kVec>
__device__ __forceinline__ int row_col(int k, int lane, int dim) {
  const int per = row_per<kVec>(dim);
  if (per) return k < per ? (((k >> 3) * 32 + lane) << 3) + (k & 7) : -1;
  const int c = k * 32 + lane;
  return c < dim ? c : -1;
}
template <class T>
__device__ __forceinline__ float cvt_in(T v);
template <>
__device__ __forceinline__ float cvt_in<float>(float v) { return v; }
template <>
__device__ __forceinline__ float cvt_in<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float cvt_in<__half>(__half v) { return __half2float(v); }
template <class T>
__device__ __forceinline__ T cvt_out(float v);
template <>
__device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half cvt_out<__half>(float v) { return __float2half_rn(v); }

template <bool kVec = false, class T, int V>
__device__ __forceinline__ void load_row_t(const T* __restrict__ p, int dim, int lane, float (&v)[V]) {
  const int per = row_per<kVec>(dim);
  if (kVec && per && per <= V && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    constexpr int E = 16 / sizeof(T);        // elements per 16-byte vector: 8 (one per group) or 4 (two per group)
#pragma unroll
    for (int k0 = 0; k0 < V; k0 += E) {
      if (k0 < per) {
        const uint4 u = *reinterpret_cast<const uint4*>(p + ((((k0 >> 3) * 32 + lane) << 3) + (k0 & 7)));
        const T* e = reinterpret_cast<const T*>(&u);
#pragma unroll
        for (int j = 0; j < E; ++j) v[k0 + j] = cvt_in<T>(e[j]);
      } else {
#pragma unroll
        for (int j = 0; j < E; ++j) v[k0 + j] = 0.f;
      }
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int c = row_col<kVec>(k, lane, dim);
    v[k] = c >= 0 ? cvt_in<T>(p[c]) : 0.f;
  }
}
template <bool kVec = false, int V>
__device__ __forceinline__ void load_row(const void* base, int dtype, int64_t row, int dim, int lane, float (&v)[V]) {
  if (dtype == COSMOS_DTYPE_F32) load_row_t<kVec>(reinterpret_cast<const float*>(base) + row * dim, dim, lane, v);
  else if (dtype == COSMOS_DTYPE_BF16) load_row_t<kVec>(reinterpret_cast<const __nv_bfloat16*>(base) + row * dim, dim, lane, v);
  else load_row_t<kVec>(reinterpret_cast<const __half*>(base) + row * dim, dim, lane, v);
}
template <bool kVec = false, class T, int V>
__device__ __forceinline__ void store_row_t(T* __restrict__ p, int dim, int lane, const float (&v)[V]) {
  const int per = row_per<kVec>(dim);
  if (kVec && per && per <= V && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    constexpr int E = 16 / sizeof(T);
#pragma unroll
    for (int k0 = 0; k0 < V; k0 += E) {
      if (k0 < per) {
        uint4 u;
        T* e = reinterpret_cast<T*>(&u);
#pragma unroll
        for (int j = 0; j < E; ++j) e[j] = cvt_out<T>(v[k0 + j]);
        *reinterpret_cast<uint4*>(p + ((((k0 >> 3) * 32 + lane) << 3) + (k0 & 7))) = u;
      }
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int c = row_col<kVec>(k, lane, dim);
    if (c >= 0) p[c] = cvt_out<T>(v[k]);
  }
}
template <bool kVec = false, int V>
__device__ __forceinline__ void store_row(void* base, int dtype, int64_t row, int dim, int lane, const float (&v)[V]) {
  if (dtype == COSMOS_DTYPE_F32) store_row_t<kVec>(reinterpret_cast<float*>(base) + row * dim, dim, lane, v);
  else if (dtype == COSMOS_DTYPE_BF16) store_row_t<kVec>(reinterpret_cast<__nv_bfloat16*>(base) + row * dim, dim, lane, v);
  else store_row_t<kVec>(reinterpret_cast<__half*>(base) + row * dim, dim, lane, v);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row lives in registers (dim <= 1024)
// ------------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const void* __restrict__ x, int x_dtype, const float* __restrict__ w, const float* __restrict__ b,
                     void* __restrict__ y, int y_dtype, float* __restrict__ mean, float* __restrict__ rstd, int64_t rows, int dim, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[V], wv[V], bv[V];
  load_row<true>(x, x_dtype, row, dim, lane, v);
  load_row_t<true>(w, dim, lane, wv);
  load_row_t<true>(b, dim, lane, bv);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) s += v[k];
  const float mu = warp_sum(s) / dim;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const float d = row_col<true>(k, lane, dim) >= 0 ? v[k] - mu : 0.f;
    q += d * d;
  }
  const float rs = rsqrtf(warp_sum(q) / dim + eps);
#pragma unroll
  for (int k = 0; k < V; ++k) v[k] = (v[k] - mu) * rs * wv[k] + bv[k];
  store_row<true>(y, y_dtype, row, dim, lane, v);
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
}

#ifndef COSMOS_LN_BWD_VEC
#define COSMOS_LN_BWD_VEC 1
#endif
constexpr bool kLnBwdVec = COSMOS_LN_BWD_VEC != 0;      // which row map the backward uses (A/B builds flip it)

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * w;  dw += dy * xhat, db += dy (block partials -> atomics)
template <int V>
__global__ void __launch_bounds__(256, V <= 16 ? 2 : 1)
layernorm_bwd_kernel(const void* __restrict__ dy, int dy_dtype, const void* __restrict__ x, int x_dtype,
                     const float* __restrict__ w, const float* __restrict__ mean, const float* __restrict__ rstd,
                     void* __restrict__ dx, int dx_dtype, int accumulate, float* __restrict__ dw, float* __restrict__ db,
                     int64_t rows, int dim, int rows_per_block) {
  extern __shared__ float red[];   // [2][dim]
  for (int c = threadIdx.x; c < 2 * dim; c += blockDim.x) red[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float aw[V], ab[V], wv[V];
#pragma unroll
  for (int k = 0; k < V; ++k) aw[k] = ab[k] = 0.f;
  load_row_t<kLnBwdVec>(w, dim, lane, wv);
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  for (int64_t row = r0 + wid; row < min(rows, r0 + rows_per_block); row += 8) {
    const float mu = mean[row], rs = rstd[row];
    float g[V], xh[V];
    load_row<kLnBwdVec>(dy, dy_dtype, row, dim, lane, g);
    load_row<kLnBwdVec>(x, x_dtype, row, dim, lane, xh);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      xh[k] = row_col<kLnBwdVec>(k, lane, dim) >= 0 ? (xh[k] - mu) * rs : 0.f;
      aw[k] = fmaf(g[k], xh[k], aw[k]);
      ab[k] += g[k];
      g[k] *= wv[k];
      s1 += g[k];
      s2 = fmaf(g[k], xh[k], s2);
    }
    s1 = warp_sum(s1) / dim;
    s2 = warp_sum(s2) / dim;
    // (in place: the result overwrites g, a previous dx to add to is loaded into xh - one row array less alive)
#pragma unroll
    for (int k = 0; k < V; ++k) g[k] = rs * (g[k] - s1 - xh[k] * s2);
    if (accumulate) {
      load_row<kLnBwdVec>(dx, dx_dtype, row, dim, lane, xh);
#pragma unroll
      for (int k = 0; k < V; ++k) g[k] += xh[k];
    }
    store_row<kLnBwdVec>(dx, dx_dtype, row, dim, lane, g);
  }
  if (dw != nullptr) {
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int c = row_col<kLnBwdVec>(k, lane, dim);
      if (c >= 0) {
        atomicAdd(&red[c], aw[k]);
        atomicAdd(&red[dim + c], ab[k]);
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < dim; c += blockDim.x) {
      atomicAdd(dw + c, red[c]);
      atomicAdd(db + c, red[dim + c]);
    }
  }
}

// ---- forward, second generation for dims that are multiples of 256 (the vector row map): CTAs stride over the rows, a warp
// has TWO rows in flight, and the LayerNorm weights live in shared memory (register order: a lane's four consecutive values
// are one conflict-free LDS.128) so that the second row does not cost occupancy.  tools/ln_ab.py, [200704, 512] bf16, same
// box: 149.5 -> 135 us (3.0 TB/s).  Measured and rejected: loading the NEXT row before reducing the current one (178 us: the
// second register set halves the resident warps), four rows per warp (197 us); for the backward, two rows per warp (389 us)
// and next-row prefetch (299 us, the same as without) - it stays on the first-generation kernel with the vector map.
__device__ __forceinline__ int vec_slot(int k, int lane) { return (((k >> 2) * 32 + lane) << 2) + (k & 3); }   // of register k

template <int V>
__global__ void __launch_bounds__(256)
layernorm_fwd2_kernel(const void* __restrict__ x, int x_dtype, const float* __restrict__ w, const float* __restrict__ b,
                      void* __restrict__ y, int y_dtype, float* __restrict__ mean, float* __restrict__ rstd, int64_t rows, int dim, float eps) {
  extern __shared__ __align__(16) float swb[];            // [V * 32] w, [V * 32] b in register order
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int per = dim >> 5;
  for (int i = threadIdx.x; i < V * 32; i += blockDim.x) {
    const int k = ((i >> 7) << 2) + (i & 3), ln = (i >> 2) & 31;       // inverse of vec_slot
    const int c = row_col<true>(k, ln, dim);
    swb[i] = c >= 0 ? w[c] : 0.f;
    swb[V * 32 + i] = c >= 0 ? b[c] : 0.f;
  }
  __syncthreads();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * 16;
  for (int64_t r0 = (static_cast<int64_t>(blockIdx.x) * 8 + wid) * 2; r0 < rows; r0 += stride) {
    const bool two = r0 + 1 < rows;
    float v[2][V];
    load_row<true>(x, x_dtype, r0, dim, lane, v[0]);
    if (two) {
      load_row<true>(x, x_dtype, r0 + 1, dim, lane, v[1]);
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) v[1][k] = 0.f;
    }
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      s0 += v[0][k];
      s1 += v[1][k];
    }
    const float mu0 = warp_sum(s0) / dim, mu1 = warp_sum(s1) / dim;
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float d0 = k < per ? v[0][k] - mu0 : 0.f, d1 = k < per ? v[1][k] - mu1 : 0.f;
      q0 += d0 * d0;
      q1 += d1 * d1;
    }
    const float rs0 = rsqrtf(warp_sum(q0) / dim + eps), rs1 = rsqrtf(warp_sum(q1) / dim + eps);
#pragma unroll
    for (int k4 = 0; k4 < V; k4 += 4) {
      const float4 ww = *reinterpret_cast<const float4*>(swb + vec_slot(k4, lane));
      const float4 bb = *reinterpret_cast<const float4*>(swb + V * 32 + vec_slot(k4, lane));
      const float wv[4] = {ww.x, ww.y, ww.z, ww.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[0][k4 + j] = (v[0][k4 + j] - mu0) * rs0 * wv[j] + bv[j];
        v[1][k4 + j] = (v[1][k4 + j] - mu1) * rs1 * wv[j] + bv[j];
      }
    }
    store_row<true>(y, y_dtype, r0, dim, lane, v[0]);
    if (two) store_row<true>(y, y_dtype, r0 + 1, dim, lane, v[1]);
    if (lane == 0) {
      mean[r0] = mu0;
      rstd[r0] = rs0;
      if (two) {
        mean[r0 + 1] = mu1;
        rstd[r0 + 1] = rs1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// attention core: one CTA per (key/value set, head); K_h and V_h of the set stay in shared memory while the
// set's few queries are processed.  Rows padded to hd + 2 elements so that "one thread = one key" reads do
// not collide on a bank.
// ------------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ float tof(T v);
template <>
__device__ __forceinline__ float tof<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float tof<__half>(__half v) { return __half2float(v); }
template <class T>
__device__ __forceinline__ T fromf(float v);
template <>
__device__ __forceinline__ __nv_bfloat16 fromf<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half fromf<__half>(float v) { return __float2half_rn(v); }

struct AttnArgs {
  int n_sets, L, dim, heads, hd, q_per_set;
  int64_t q_stride_set, q_stride_q;
};

template <class T>
__device__ __forceinline__ void load_kv(const T* __restrict__ kv, T* sK, T* sV, const AttnArgs& a, int set, int h) {
  // 16-byte global loads (hd is a power of two >= 16), 4-byte stores into the padded rows
  const int hd = a.hd, ldk = hd + 2;
  const int per_row = hd >> 3;                       // uint4 per row
  const int shift = 31 - __clz(per_row);
  for (int idx = threadIdx.x; idx < a.L * per_row; idx += blockDim.x) {
    const int l = idx >> shift, j = idx & (per_row - 1);
    const T* rowp = kv + (static_cast<size_t>(set) * a.L + l) * (2 * a.dim) + h * hd + j * 8;
    const uint4 k4 = __ldg(reinterpret_cast<const uint4*>(rowp));
    const uint4 v4 = __ldg(reinterpret_cast<const uint4*>(rowp + a.dim));
    uint32_t* dk = reinterpret_cast<uint32_t*>(sK + l * ldk + j * 8);
    uint32_t* dv = reinterpret_cast<uint32_t*>(sV + l * ldk + j * 8);
    dk[0] = k4.x; dk[1] = k4.y; dk[2] = k4.z; dk[3] = k4.w;
    dv[0] = v4.x; dv[1] = v4.y; dv[2] = v4.z; dv[3] = v4.w;
  }
}

// block-wide (128 threads) max / sum through shared scratch
__device__ __forceinline__ float block_reduce(float v, float* scratch, bool is_max) {
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = scratch[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) r = is_max ? fmaxf(r, scratch[w]) : r + scratch[w];
  return r;
}

constexpr int kAttnThreads = 256;
constexpr int kQG = 8;   // queries processed together (one register accumulator each)

// scores of up to kQG queries against key row l: acc[c] += sum_k A[c][k] * row[k]   (A fp32 [kQG][hd] in smem)
template <class T>
__device__ __forceinline__ void dot_rows(const float* __restrict__ A, const T* __restrict__ row, int hd, float (&acc)[kQG]) {
  for (int k0 = 0; k0 < hd; k0 += 8) {
    float kv[8];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const uint32_t u = *reinterpret_cast<const uint32_t*>(row + k0 + 2 * w);
      T lo, hi;
      *reinterpret_cast<uint16_t*>(&lo) = static_cast<uint16_t>(u & 0xffff);
      *reinterpret_cast<uint16_t*>(&hi) = static_cast<uint16_t>(u >> 16);
      kv[2 * w] = tof(lo);
      kv[2 * w + 1] = tof(hi);
    }
#pragma unroll
    for (int c = 0; c < kQG; ++c) {
      const float4 a0 = *reinterpret_cast<const float4*>(A + c * hd + k0);
      const float4 a1 = *reinterpret_cast<const float4*>(A + c * hd + k0 + 4);
      acc[c] = fmaf(a0.x, kv[0], acc[c]); acc[c] = fmaf(a0.y, kv[1], acc[c]);
      acc[c] = fmaf(a0.z, kv[2], acc[c]); acc[c] = fmaf(a0.w, kv[3], acc[c]);
      acc[c] = fmaf(a1.x, kv[4], acc[c]); acc[c] = fmaf(a1.y, kv[5], acc[c]);
      acc[c] = fmaf(a1.z, kv[6], acc[c]); acc[c] = fmaf(a1.w, kv[7], acc[c]);
    }
  }
}

// out[c][k] = sum_l W[c][l] * M[l][k] for this thread's k and its share of the keys: groups of 4 consecutive keys
// (one LDS.128 of W per query and group).  W rows are padded to a multiple of 4 with zeros.
template <class T>
__device__ __forceinline__ void weighted_rows(const float* __restrict__ W, int ldw, const T* __restrict__ M, int ldk, int L, int k,
                                              int part, int parts, float (&acc)[kQG]) {
  for (int l0 = part * 4; l0 < L; l0 += parts * 4) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (l0 + u < L) ? tof(M[(l0 + u) * ldk + k]) : 0.f;
#pragma unroll
    for (int c = 0; c < kQG; ++c) {
      const float4 w = *reinterpret_cast<const float4*>(W + c * ldw + l0);
      acc[c] = fmaf(w.x, v[0], acc[c]);
      acc[c] = fmaf(w.y, v[1], acc[c]);
      acc[c] = fmaf(w.z, v[2], acc[c]);
      acc[c] = fmaf(w.w, v[3], acc[c]);
    }
  }
}

template <class T>
__global__ void __launch_bounds__(kAttnThreads)
attn_core_fwd_kernel(const T* __restrict__ q, const T* __restrict__ kv, T* __restrict__ o, float* __restrict__ lse, AttnArgs a) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int hd = a.hd, ldk = hd + 2, L = a.L;
  const int Lp = (L + 3) & ~3;
  T* sK = reinterpret_cast<T*>(sm);
  T* sV = sK + L * ldk;
  float* sQ = reinterpret_cast<float*>(sm + ((2 * static_cast<size_t>(L) * ldk * 2 + 15) & ~size_t(15)));   // [kQG][hd]
  float* sP = sQ + kQG * hd;                              // [kQG][Lp]
  float* sRed = sP + kQG * Lp;                            // [8 warps][kQG]
  float* sStat = sRed + 8 * kQG;                          // [kQG] max, [kQG] 1/sum
  float* sO = sStat + 2 * kQG;                            // [parts][kQG][hd]
  const int set = blockIdx.x / a.heads, h = blockIdx.x - set * a.heads;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  load_kv(kv, sK, sV, a, set, h);
  const float kappa = rsqrtf(static_cast<float>(hd));
  for (int c0 = 0; c0 < a.q_per_set; c0 += kQG) {
    const int nq = min(kQG, a.q_per_set - c0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < kQG * hd; idx += blockDim.x) {
      const int c = idx / hd, k = idx - c * hd;
      const size_t qrow = static_cast<size_t>(set) * a.q_stride_set + static_cast<size_t>(c0 + c) * a.q_stride_q;
      sQ[idx] = c < nq ? tof(q[qrow * a.dim + h * hd + k]) * kappa : 0.f;
    }
    __syncthreads();
    float mx[kQG];
#pragma unroll
    for (int c = 0; c < kQG; ++c) mx[c] = -INFINITY;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
      float s[kQG];
#pragma unroll
      for (int c = 0; c < kQG; ++c) s[c] = 0.f;
      dot_rows(sQ, sK + l * ldk, hd, s);
#pragma unroll
      for (int c = 0; c < kQG; ++c) {
        sP[c * Lp + l] = s[c];
        mx[c] = fmaxf(mx[c], s[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < kQG; ++c) mx[c] = warp_max(mx[c]);
    if (lane == 0)
#pragma unroll
      for (int c = 0; c < kQG; ++c) sRed[wid * kQG + c] = mx[c];
    __syncthreads();
    if (threadIdx.x < kQG) {
      float m = sRed[threadIdx.x];
      for (int w = 1; w < 8; ++w) m = fmaxf(m, sRed[w * kQG + threadIdx.x]);
      sStat[threadIdx.x] = m;
    }
    __syncthreads();
    float sum[kQG];
#pragma unroll
    for (int c = 0; c < kQG; ++c) sum[c] = 0.f;
    for (int l = threadIdx.x; l < Lp; l += blockDim.x) {
#pragma unroll
      for (int c = 0; c < kQG; ++c) {
        const float e = l < L ? __expf(sP[c * Lp + l] - sStat[c]) : 0.f;      // pad columns stay zero for the P V product
        sP[c * Lp + l] = e;
        sum[c] += e;
      }
    }
#pragma unroll
    for (int c = 0; c < kQG; ++c) sum[c] = warp_sum(sum[c]);
    __syncthreads();
    if (lane == 0)
#pragma unroll
      for (int c = 0; c < kQG; ++c) sRed[wid * kQG + c] = sum[c];
    __syncthreads();
    if (threadIdx.x < kQG) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += sRed[w * kQG + threadIdx.x];
      sStat[kQG + threadIdx.x] = 1.f / t;
      if (threadIdx.x < nq) {
        const size_t qrow = static_cast<size_t>(set) * a.q_stride_set + static_cast<size_t>(c0 + threadIdx.x) * a.q_stride_q;
        lse[qrow * a.heads + h] = sStat[threadIdx.x] + __logf(t);
      }
    }
    __syncthreads();
    // O = P V
    const int parts = blockDim.x / hd, k = threadIdx.x % hd, part = threadIdx.x / hd;
    float acc[kQG];
#pragma unroll
    for (int c = 0; c < kQG; ++c) acc[c] = 0.f;
    weighted_rows(sP, Lp, sV, ldk, L, k, part, parts, acc);
#pragma unroll
    for (int c = 0; c < kQG; ++c) sO[(part * kQG + c) * hd + k] = acc[c];
    __syncthreads();
    for (int idx = threadIdx.x; idx < nq * hd; idx += blockDim.x) {
      const int c = idx / hd, kk = idx - c * hd;
      float r = 0.f;
      for (int pp = 0; pp < parts; ++pp) r += sO[(pp * kQG + c) * hd + kk];
      const size_t qrow = static_cast<size_t>(set) * a.q_stride_set + static_cast<size_t>(c0 + c) * a.q_stride_q;
      o[qrow * a.dim + h * hd + kk] = fromf<T>(r * sStat[kQG + c]);
    }
  }
}

template <class T>
__global__ void __launch_bounds__(kAttnThreads)
attn_core_bwd_kernel(const T* __restrict__ q, const T* __restrict__ kv, const T* __restrict__ d_o, const float* __restrict__ lse,
                     T* __restrict__ dq, T* __restrict__ dkv, AttnArgs a) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int hd = a.hd, ldk = hd + 2, L = a.L;
  const int Lp = (L + 3) & ~3;
  T* sK = reinterpret_cast<T*>(sm);
  T* sV = sK + L * ldk;
  float* sQ = reinterpret_cast<float*>(sm + ((2 * static_cast<size_t>(L) * ldk * 2 + 15) & ~size_t(15)));   // [kQG][hd] kappa * q
  float* sDO = sQ + kQG * hd;                             // [kQG][hd]
  float* sP = sDO + kQG * hd;                             // [kQG][Lp]
  float* sDS = sP + kQG * Lp;                             // [kQG][Lp]
  float* sRed = sDS + kQG * Lp;                           // [8][kQG]
  float* sStat = sRed + 8 * kQG;                          // [kQG] lse, [kQG] dot
  float* sO = sStat + 2 * kQG;                            // [parts][kQG][hd]
  const int set = blockIdx.x / a.heads, h = blockIdx.x - set * a.heads;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  load_kv(kv, sK, sV, a, set, h);
  const float kappa = rsqrtf(static_cast<float>(hd));
  const int n_groups = (a.q_per_set + kQG - 1) / kQG;
  for (int g = 0; g < n_groups; ++g) {
    const int c0 = g * kQG, nq = min(kQG, a.q_per_set - c0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < kQG * hd; idx += blockDim.x) {
      const int c = idx / hd, k = idx - c * hd;
      const size_t qrow = static_cast<size_t>(set) * a.q_stride_set + static_cast<size_t>(c0 + c) * a.q_stride_q;
      sQ[idx] = c < nq ? tof(q[qrow * a.dim + h * hd + k]) * kappa : 0.f;
      sDO[idx] = c < nq ? tof(d_o[qrow * a.dim + h * hd + k]) : 0.f;
    }
    if (threadIdx.x < kQG) {
      const size_t qrow = static_cast<size_t>(set) * a.q_stride_set + static_cast<size_t>(c0 + threadIdx.x) * a.q_stride_q;
      sStat[threadIdx.x] = threadIdx.x < nq ? lse[qrow * a.heads + h] : INFINITY;
    }
    __syncthreads();
    float dot[kQG];
#pragma unroll
    for (int c = 0; c < kQG; ++c) dot[c] = 0.f;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
      float s[kQG], dp[kQG];
#pragma unroll
      for (int c = 0; c < kQG; ++c) s[c] = dp[c] = 0.f;
      dot_rows(sQ, sK + l * ldk, hd, s);
      dot_rows(sDO, sV + l * ldk, hd, dp);
#pragma unroll
      for (int c = 0; c < kQG; ++c) {
        const float pr = __expf(s[c] - sStat[c]);     // 0 for padded queries (lse = +inf)
        sP[c * Lp + l] = pr;
        sDS[c * Lp + l] = dp[c];
        dot[c] = fmaf(pr, dp[c], dot[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < kQG; ++c) dot[c] = warp_sum(dot[c]);
    if (lane == 0)
#pragma unroll
      for (int c = 0; c < kQG; ++c) sRed[wid * kQG + c] = dot[c];
    __syncthreads();
    if (threadIdx.x < kQG) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += sRed[w * kQG + threadIdx.x];
      sStat[kQG + threadIdx.x] = t;
    }
    __syncthreads();
    for (int l = threadIdx.x; l < Lp; l += blockDim.x)
#pragma unroll
      for (int c = 0; c < kQG; ++c) sDS[c * Lp + l] = l < L ? sP[c * Lp + l] * (sDS[c * Lp + l] - sStat[kQG + c]) : 0.f;
    __syncthreads();
    // dQ = kappa * dS K
    {
      const int parts = blockDim.x / hd, k = threadIdx.x % hd, part = threadIdx.x / hd;
      float acc[kQG];
#pragma unroll
      for (int c = 0; c < kQG; ++c) acc[c] = 0.f;
      weighted_rows(sDS, Lp, sK, ldk, L, k, part, parts, acc);
#pragma unroll
      for (int c = 0; c < kQG; ++c) sO[(part * kQG + c) * hd + k] = acc[c];
      __syncthreads();
      for (int idx = threadIdx.x; idx < nq * hd; idx += blockDim.x) {
        const int c = idx / hd, kk = idx - c * hd;
        float r = 0.f;
        for (int pp = 0; pp < parts; ++pp) r += sO[(pp * kQG + c) * hd + kk];
        const size_t qrow = static_cast<size_t>(set) * a.q_stride_set + static_cast<size_t>(c0 + c) * a.q_stride_q;
        dq[qrow * a.dim + h * hd + kk] = fromf<T>(r * kappa);
      }
    }
    // dK[l][k] (+)= sum_c dS[c][l] * (kappa q[c][k]);  dV[l][k] (+)= sum_c P[c][l] dO[c][k]   (two k per thread)
    for (int idx = threadIdx.x; idx < L * (hd / 2); idx += blockDim.x) {
      const int l = idx / (hd / 2), k = (idx - l * (hd / 2)) * 2;
      float dk0 = 0.f, dk1 = 0.f, dv0 = 0.f, dv1 = 0.f;
#pragma unroll
      for (int c = 0; c < kQG; ++c) {
        const float ds = sDS[c * Lp + l], pr = sP[c * Lp + l];
        const float2 qq = *reinterpret_cast<const float2*>(sQ + c * hd + k);
        const float2 oo = *reinterpret_cast<const float2*>(sDO + c * hd + k);
        dk0 = fmaf(ds, qq.x, dk0); dk1 = fmaf(ds, qq.y, dk1);
        dv0 = fmaf(pr, oo.x, dv0); dv1 = fmaf(pr, oo.y, dv1);
      }
      T* rowp = dkv + (static_cast<size_t>(set) * L + l) * (2 * a.dim) + h * hd + k;
      if (g > 0) {   // more than kQG queries per set: accumulate over the groups
        dk0 += tof(rowp[0]); dk1 += tof(rowp[1]); dv0 += tof(rowp[a.dim]); dv1 += tof(rowp[a.dim + 1]);
      }
      T pk[2] = {fromf<T>(dk0), fromf<T>(dk1)}, pv[2] = {fromf<T>(dv0), fromf<T>(dv1)};
      *reinterpret_cast<uint32_t*>(rowp) = *reinterpret_cast<const uint32_t*>(pk);            // k is even: 4-byte aligned
      *reinterpret_cast<uint32_t*>(rowp + a.dim) = *reinterpret_cast<const uint32_t*>(pv);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// residual add + L2 normalise (one warp per row)
// ------------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256)
addnorm_fwd_kernel(const void* __restrict__ f, int f_dtype, const float* __restrict__ pooled, void* __restrict__ out,
                   float* __restrict__ inv_norm, int64_t rows, int dim) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[V], pv[V];
  load_row(f, f_dtype, row, dim, lane, v);
  load_row_t(pooled + row * dim, dim, lane, pv);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    // the reference adds in the feature dtype (model.py:379): round the sum to it before normalising
    float z = v[k] + pv[k];
    if (f_dtype == COSMOS_DTYPE_BF16) z = __bfloat162float(__float2bfloat16_rn(z));
    else if (f_dtype == COSMOS_DTYPE_F16) z = __half2float(__float2half_rn(z));
    v[k] = z;
    s += z * z;
  }
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(s)), 1e-12f);
#pragma unroll
  for (int k = 0; k < V; ++k) v[k] *= inv;
  store_row(out, f_dtype, row, dim, lane, v);
  if (lane == 0) inv_norm[row] = inv;
}

template <int V>
__global__ void __launch_bounds__(256)
addnorm_bwd_kernel(const void* __restrict__ g_out, const void* __restrict__ out, int f_dtype, const float* __restrict__ inv_norm,
                   float* __restrict__ g_z32, void* __restrict__ g_z16, int g_dtype, int64_t rows, int dim) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float g[V], y[V];
  load_row(g_out, f_dtype, row, dim, lane, g);
  load_row(out, f_dtype, row, dim, lane, y);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) s = fmaf(g[k], y[k], s);
  s = warp_sum(s);
  const float inv = inv_norm[row];
#pragma unroll
  for (int k = 0; k < V; ++k) g[k] = (g[k] - y[k] * s) * inv;
  store_row_t(g_z32 + row * dim, dim, lane, g);
  store_row(g_z16, g_dtype, row, dim, lane, g);
}

__global__ void __launch_bounds__(256)
colsum_kernel(const void* __restrict__ src, int dtype, float* __restrict__ dst, int64_t rows, int n, int64_t ld, int rows_per_block) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_block;
  float acc = 0.f;
  for (int64_t r = r0; r < min(rows, r0 + rows_per_block); ++r) acc += ld_elem(src, dtype, r * ld + c);
  atomicAdd(dst + c, acc);
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
cudaError_t launch_layernorm_fwd(const void* x, int x_dtype, const float* w, const float* b, void* y, int y_dtype, float* mean,
                                 float* rstd, int64_t rows, int dim, float eps, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  if ((dim & 255) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    // strided CTAs, two rows per warp in flight, weights in shared memory (dims 256 / 512 / 768 / 1024)
    const int64_t want = (rows + 15) / 16;
    const unsigned grid2 = static_cast<unsigned>(want < 148 * 8 ? want : 148 * 8);
    if (dim <= 256) layernorm_fwd2_kernel<8><<<grid2, 256, 2 * 8 * 32 * 4, s>>>(x, x_dtype, w, b, y, y_dtype, mean, rstd, rows, dim, eps);
    else if (dim <= 512) layernorm_fwd2_kernel<16><<<grid2, 256, 2 * 16 * 32 * 4, s>>>(x, x_dtype, w, b, y, y_dtype, mean, rstd, rows, dim, eps);
    else layernorm_fwd2_kernel<32><<<grid2, 256, 2 * 32 * 32 * 4, s>>>(x, x_dtype, w, b, y, y_dtype, mean, rstd, rows, dim, eps);
    return cudaGetLastError();
  }
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  if (dim <= 256) layernorm_fwd_kernel<8><<<grid, 256, 0, s>>>(x, x_dtype, w, b, y, y_dtype, mean, rstd, rows, dim, eps);
  else if (dim <= 512) layernorm_fwd_kernel<16><<<grid, 256, 0, s>>>(x, x_dtype, w, b, y, y_dtype, mean, rstd, rows, dim, eps);
  else layernorm_fwd_kernel<32><<<grid, 256, 0, s>>>(x, x_dtype, w, b, y, y_dtype, mean, rstd, rows, dim, eps);
  return cudaGetLastError();
}

cudaError_t launch_layernorm_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* w, const float* mean,
                                 const float* rstd, void* dx, int dx_dtype, int accumulate, float* dw, float* db, int64_t rows,
                                 int dim, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  // few, long-running blocks: the dw/db partials of a block end in 2*dim global atomics
  int64_t blocks = (rows + 63) / 64;
  if (blocks > 592) blocks = 592;
  const int rows_per_block = static_cast<int>(((rows + blocks - 1) / blocks + 7) / 8 * 8);
  const unsigned grid = static_cast<unsigned>((rows + rows_per_block - 1) / rows_per_block);
  const size_t sm = 2 * dim * sizeof(float);
  if (dim <= 256)
    layernorm_bwd_kernel<8><<<grid, 256, sm, s>>>(dy, dy_dtype, x, x_dtype, w, mean, rstd, dx, dx_dtype, accumulate, dw, db, rows, dim, rows_per_block);
  else if (dim <= 512)
    layernorm_bwd_kernel<16><<<grid, 256, sm, s>>>(dy, dy_dtype, x, x_dtype, w, mean, rstd, dx, dx_dtype, accumulate, dw, db, rows, dim, rows_per_block);
  else
    layernorm_bwd_kernel<32><<<grid, 256, sm, s>>>(dy, dy_dtype, x, x_dtype, w, mean, rstd, dx, dx_dtype, accumulate, dw, db, rows, dim, rows_per_block);
  return cudaGetLastError();
}

static AttnArgs make_attn_args(int n_sets, int L, int dim, int heads, int q_per_set, int64_t qs, int64_t qq) {
  AttnArgs a;
  a.n_sets = n_sets; a.L = L; a.dim = dim; a.heads = heads; a.hd = dim / heads; a.q_per_set = q_per_set;
  a.q_stride_set = qs; a.q_stride_q = qq;
  return a;
}

cudaError_t launch_attn_core_fwd(const void* q, const void* kv, void* o, float* lse, int dtype, int n_sets, int L, int dim, int heads,
                                 int q_per_set, int64_t qs, int64_t qq, cudaStream_t s) {
  const AttnArgs a = make_attn_args(n_sets, L, dim, heads, q_per_set, qs, qq);
  const size_t Lp = (L + 3) & ~3;
  const size_t smem = ((2 * static_cast<size_t>(L) * (a.hd + 2) * 2 + 15) & ~size_t(15)) +
                      (kQG * a.hd + kQG * Lp + 8 * kQG + 2 * kQG + kAttnThreads * kQG) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e;
  if (dtype == COSMOS_DTYPE_BF16) {
    e = cudaFuncSetAttribute(attn_core_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attn_core_fwd_kernel<__nv_bfloat16><<<n_sets * heads, kAttnThreads, smem, s>>>(
        static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(kv), static_cast<__nv_bfloat16*>(o), lse, a);
  } else {
    e = cudaFuncSetAttribute(attn_core_fwd_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attn_core_fwd_kernel<__half><<<n_sets * heads, kAttnThreads, smem, s>>>(static_cast<const __half*>(q), static_cast<const __half*>(kv),
                                                                     static_cast<__half*>(o), lse, a);
  }
  return cudaGetLastError();
}

cudaError_t launch_attn_core_bwd(const void* q, const void* kv, const void* d_o, const float* lse, void* dq, void* dkv, int dtype,
                                 int n_sets, int L, int dim, int heads, int q_per_set, int64_t qs, int64_t qq, cudaStream_t s) {
  const AttnArgs a = make_attn_args(n_sets, L, dim, heads, q_per_set, qs, qq);
  const size_t Lp = (L + 3) & ~3;
  const size_t smem = ((2 * static_cast<size_t>(L) * (a.hd + 2) * 2 + 15) & ~size_t(15)) +
                      (2 * kQG * a.hd + 2 * kQG * Lp + 8 * kQG + 2 * kQG + kAttnThreads * kQG) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e;
  if (dtype == COSMOS_DTYPE_BF16) {
    e = cudaFuncSetAttribute(attn_core_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attn_core_bwd_kernel<__nv_bfloat16><<<n_sets * heads, kAttnThreads, smem, s>>>(
        static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(kv), static_cast<const __nv_bfloat16*>(d_o), lse,
        static_cast<__nv_bfloat16*>(dq), static_cast<__nv_bfloat16*>(dkv), a);
  } else {
    e = cudaFuncSetAttribute(attn_core_bwd_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attn_core_bwd_kernel<__half><<<n_sets * heads, kAttnThreads, smem, s>>>(static_cast<const __half*>(q), static_cast<const __half*>(kv),
                                                                     static_cast<const __half*>(d_o), lse, static_cast<__half*>(dq),
                                                                     static_cast<__half*>(dkv), a);
  }
  return cudaGetLastError();
}

cudaError_t launch_addnorm_fwd(const void* f, int f_dtype, const float* pooled, void* out, float* inv_norm, int64_t rows, int dim,
                               cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  if (dim <= 256) addnorm_fwd_kernel<8><<<grid, 256, 0, s>>>(f, f_dtype, pooled, out, inv_norm, rows, dim);
  else if (dim <= 512) addnorm_fwd_kernel<16><<<grid, 256, 0, s>>>(f, f_dtype, pooled, out, inv_norm, rows, dim);
  else addnorm_fwd_kernel<32><<<grid, 256, 0, s>>>(f, f_dtype, pooled, out, inv_norm, rows, dim);
  return cudaGetLastError();
}

cudaError_t launch_addnorm_bwd(const void* g_out, const void* out, int f_dtype, const float* inv_norm, float* g_z32, void* g_z16,
                               int g_dtype, int64_t rows, int dim, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  if (dim <= 256) addnorm_bwd_kernel<8><<<grid, 256, 0, s>>>(g_out, out, f_dtype, inv_norm, g_z32, g_z16, g_dtype, rows, dim);
  else if (dim <= 512) addnorm_bwd_kernel<16><<<grid, 256, 0, s>>>(g_out, out, f_dtype, inv_norm, g_z32, g_z16, g_dtype, rows, dim);
  else addnorm_bwd_kernel<32><<<grid, 256, 0, s>>>(g_out, out, f_dtype, inv_norm, g_z32, g_z16, g_dtype, rows, dim);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// folded attention: softmax over the KEYS (rows) of a [L, Nc] score block per token set, Nc = queries x heads columns
// ------------------------------------------------------------------------------------------------
// One CTA per set; thread = (column within a pass of 64, one of 4 row groups): a warp reads 32 consecutive columns of a row.
// The block (L x Nc fp32, 50 KB at 196 x 64) is read twice; the second read hits L2.
template <class T>
__global__ void __launch_bounds__(256)
colsoftmax_fwd_kernel(const float* __restrict__ S, long long s_bs, int lds, T* __restrict__ P, long long p_bs, int ldp, int L, int Nc,
                      int zero_key) {
  __shared__ float sm_m[4][64], sm_s[4][64];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const float* s0 = S + static_cast<size_t>(blockIdx.x) * s_bs;
  T* p0 = P + static_cast<size_t>(blockIdx.x) * p_bs;
  for (int c0 = 0; c0 < Nc; c0 += 64) {
    const int c = c0 + tx;
    float m = -INFINITY, sum = 0.f;
    if (c < Nc) {
      for (int l = ty; l < L; l += 4) {
        const float v = s0[static_cast<size_t>(l) * lds + c];
        if (v > m) {
          sum *= __expf(m - v);
          m = v;
        }
        sum += __expf(v - m);
      }
    }
    __syncthreads();
    sm_m[ty][tx] = m;
    sm_s[ty][tx] = sum;
    __syncthreads();
    float M = fmaxf(fmaxf(sm_m[0][tx], sm_m[1][tx]), fmaxf(sm_m[2][tx], sm_m[3][tx]));
    if (zero_key) M = fmaxf(M, 0.f);       // add_zero_attn: one more key with score 0 (and value 0: it only enters the denominator)
    float tot = zero_key ? __expf(-M) : 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (sm_m[g][tx] != -INFINITY) tot += sm_s[g][tx] * __expf(sm_m[g][tx] - M);
    const float inv = 1.f / tot;
    if (c < Nc) {
      for (int l = ty; l < L; l += 4)
        p0[static_cast<size_t>(l) * ldp + c] = fromf<T>(__expf(s0[static_cast<size_t>(l) * lds + c] - M) * inv);
    }
  }
}

// dS = P * (dP - sum_l P dP) per column
template <class T>
__global__ void __launch_bounds__(256)
colsoftmax_bwd_kernel(const T* __restrict__ P, long long p_bs, int ldp, const float* __restrict__ dP, long long d_bs, int ldd,
                      T* __restrict__ dS, long long ds_bs, int ldds, int L, int Nc) {
  __shared__ float sm_d[4][64];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const T* p0 = P + static_cast<size_t>(blockIdx.x) * p_bs;
  const float* d0 = dP + static_cast<size_t>(blockIdx.x) * d_bs;
  T* o0 = dS + static_cast<size_t>(blockIdx.x) * ds_bs;
  for (int c0 = 0; c0 < Nc; c0 += 64) {
    const int c = c0 + tx;
    float dot = 0.f;
    if (c < Nc)
      for (int l = ty; l < L; l += 4) dot = fmaf(tof(p0[static_cast<size_t>(l) * ldp + c]), d0[static_cast<size_t>(l) * ldd + c], dot);
    __syncthreads();
    sm_d[ty][tx] = dot;
    __syncthreads();
    const float tot = (sm_d[0][tx] + sm_d[1][tx]) + (sm_d[2][tx] + sm_d[3][tx]);
    if (c < Nc)
      for (int l = ty; l < L; l += 4)
        o0[static_cast<size_t>(l) * ldds + c] =
            fromf<T>(tof(p0[static_cast<size_t>(l) * ldp + c]) * (d0[static_cast<size_t>(l) * ldd + c] - tot));
  }
}

cudaError_t launch_colsoftmax_fwd(const float* S, int64_t s_bs, int lds, void* P, int64_t p_bs, int ldp, int dtype, int n_sets, int L,
                                  int Nc, int zero_key, cudaStream_t s) {
  if (n_sets == 0) return cudaSuccess;
  if (dtype == COSMOS_DTYPE_BF16)
    colsoftmax_fwd_kernel<__nv_bfloat16><<<n_sets, 256, 0, s>>>(S, s_bs, lds, static_cast<__nv_bfloat16*>(P), p_bs, ldp, L, Nc, zero_key);
  else
    colsoftmax_fwd_kernel<__half><<<n_sets, 256, 0, s>>>(S, s_bs, lds, static_cast<__half*>(P), p_bs, ldp, L, Nc, zero_key);
  return cudaGetLastError();
}

cudaError_t launch_colsoftmax_bwd(const void* P, int64_t p_bs, int ldp, const float* dP, int64_t d_bs, int ldd, void* dS, int64_t ds_bs,
                                  int ldds, int dtype, int n_sets, int L, int Nc, cudaStream_t s) {
  if (n_sets == 0) return cudaSuccess;
  if (dtype == COSMOS_DTYPE_BF16)
    colsoftmax_bwd_kernel<__nv_bfloat16><<<n_sets, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(P), p_bs, ldp, dP, d_bs, ldd,
                                                            static_cast<__nv_bfloat16*>(dS), ds_bs, ldds, L, Nc);
  else
    colsoftmax_bwd_kernel<__half><<<n_sets, 256, 0, s>>>(static_cast<const __half*>(P), p_bs, ldp, dP, d_bs, ldd,
                                                     static_cast<__half*>(dS), ds_bs, ldds, L, Nc);
  return cudaGetLastError();
}

cudaError_t launch_colsum(const void* src, int dtype, float* dst, int64_t rows, int n, int64_t ld, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  const int rows_per_block = 256;
  dim3 grid((n + 255) / 256, static_cast<unsigned>((rows + rows_per_block - 1) / rows_per_block));
  colsum_kernel<<<grid, 256, 0, s>>>(src, dtype, dst, rows, n, ld, rows_per_block);
  return cudaGetLastError();
}

}  // namespace cb
