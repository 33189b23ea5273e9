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

constexpr int kLnMaxPerLane = 32;   // dim <= 1024

}  // namespace

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row lives in registers (dim <= 1024)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const void* __restrict__ x, int x_dtype, const float* __restrict__ w, const float* __restrict__ b,
                     void* __restrict__ y, int y_dtype, float* __restrict__ mean, float* __restrict__ rstd, int64_t rows, int dim) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[kLnMaxPerLane];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kLnMaxPerLane; ++k) {
    const int c = k * 32 + lane;
    v[k] = c < dim ? ld_elem(x, x_dtype, row * dim + c) : 0.f;
    s += v[k];
  }
  const float mu = warp_sum(s) / dim;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < kLnMaxPerLane; ++k) {
    const int c = k * 32 + lane;
    const float d = c < dim ? v[k] - mu : 0.f;
    q += d * d;
  }
  const float rs = rsqrtf(warp_sum(q) / dim + 1e-5f);
#pragma unroll
  for (int k = 0; k < kLnMaxPerLane; ++k) {
    const int c = k * 32 + lane;
    if (c < dim) st_elem(y, y_dtype, row * dim + c, (v[k] - mu) * rs * w[c] + b[c]);
  }
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * w;  dw += dy * xhat, db += dy (block partials -> atomics)
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const void* __restrict__ dy, int dy_dtype, const void* __restrict__ x, int x_dtype,
                     const float* __restrict__ w, const float* __restrict__ mean, const float* __restrict__ rstd,
                     void* __restrict__ dx, int dx_dtype, int accumulate, float* __restrict__ dw, float* __restrict__ db,
                     int64_t rows, int dim, int rows_per_block) {
  extern __shared__ float red[];   // [2][dim]
  for (int c = threadIdx.x; c < 2 * dim; c += blockDim.x) red[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float aw[kLnMaxPerLane], ab[kLnMaxPerLane];
#pragma unroll
  for (int k = 0; k < kLnMaxPerLane; ++k) aw[k] = ab[k] = 0.f;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  for (int64_t row = r0 + wid; row < min(rows, r0 + rows_per_block); row += 8) {
    const float mu = mean[row], rs = rstd[row];
    float g[kLnMaxPerLane], xh[kLnMaxPerLane];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < kLnMaxPerLane; ++k) {
      const int c = k * 32 + lane;
      if (c < dim) {
        const float d = ld_elem(dy, dy_dtype, row * dim + c);
        xh[k] = (ld_elem(x, x_dtype, row * dim + c) - mu) * rs;
        g[k] = d * w[c];
        aw[k] += d * xh[k];
        ab[k] += d;
        s1 += g[k];
        s2 += g[k] * xh[k];
      } else {
        g[k] = xh[k] = 0.f;
      }
    }
    s1 = warp_sum(s1) / dim;
    s2 = warp_sum(s2) / dim;
#pragma unroll
    for (int k = 0; k < kLnMaxPerLane; ++k) {
      const int c = k * 32 + lane;
      if (c < dim) {
        float o = rs * (g[k] - s1 - xh[k] * s2);
        if (accumulate) o += ld_elem(dx, dx_dtype, row * dim + c);
        st_elem(dx, dx_dtype, row * dim + c, o);
      }
    }
  }
  if (dw != nullptr) {
#pragma unroll
    for (int k = 0; k < kLnMaxPerLane; ++k) {
      const int c = k * 32 + lane;
      if (c < dim) {
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
  const int hd = a.hd, ldk = hd + 2;
  const int per_row = hd / 2;      // 32-bit words per row
  for (int idx = threadIdx.x; idx < a.L * per_row; idx += blockDim.x) {
    const int l = idx / per_row, w = idx - l * per_row;
    const T* rowp = kv + (static_cast<size_t>(set) * a.L + l) * (2 * a.dim) + h * hd;
    reinterpret_cast<uint32_t*>(sK + l * ldk)[w] = reinterpret_cast<const uint32_t*>(rowp)[w];
    reinterpret_cast<uint32_t*>(sV + l * ldk)[w] = reinterpret_cast<const uint32_t*>(rowp + a.dim)[w];
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

template <class T>
__global__ void __launch_bounds__(128)
attn_core_fwd_kernel(const T* __restrict__ q, const T* __restrict__ kv, T* __restrict__ o, float* __restrict__ lse, AttnArgs a) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int hd = a.hd, ldk = hd + 2, L = a.L;
  T* sK = reinterpret_cast<T*>(sm);
  T* sV = sK + L * ldk;
  float* sQ = reinterpret_cast<float*>(sV + L * ldk);     // [hd]
  float* sP = sQ + hd;                                    // [L]
  float* scratch = sP + L;                                // [8]
  float* sO = scratch + 8;                                // [128]: (128 / hd) partial sums per output
  const int set = blockIdx.x / a.heads, h = blockIdx.x - set * a.heads;
  load_kv(kv, sK, sV, a, set, h);
  const float kappa = rsqrtf(static_cast<float>(hd));
  for (int c = 0; c < a.q_per_set; ++c) {
    const size_t qrow = static_cast<size_t>(set) * a.q_stride_set + static_cast<size_t>(c) * a.q_stride_q;
    __syncthreads();
    for (int k = threadIdx.x; k < hd; k += blockDim.x) sQ[k] = tof(q[qrow * a.dim + h * hd + k]) * kappa;
    __syncthreads();
    float mx = -INFINITY;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
      float s = 0.f;
      for (int k = 0; k < hd; k += 2) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(sK + l * ldk + k);
        T lo, hi;
        *reinterpret_cast<uint16_t*>(&lo) = static_cast<uint16_t>(w & 0xffff);
        *reinterpret_cast<uint16_t*>(&hi) = static_cast<uint16_t>(w >> 16);
        s = fmaf(sQ[k], tof(lo), s);
        s = fmaf(sQ[k + 1], tof(hi), s);
      }
      sP[l] = s;
      mx = fmaxf(mx, s);
    }
    mx = block_reduce(mx, scratch, true);
    float sum = 0.f;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
      const float e = __expf(sP[l] - mx);
      sP[l] = e;
      sum += e;
    }
    sum = block_reduce(sum, scratch, false);   // also makes sP visible
    const float inv = 1.f / sum;
    // O[k] = sum_l p_l V[l][k]: thread = (k, part), part strides over the keys; 128 / hd parts
    {
      const int parts = blockDim.x / hd;
      const int k = threadIdx.x % hd, part = threadIdx.x / hd;
      float acc = 0.f;
      for (int l = part; l < L; l += parts) acc = fmaf(sP[l], tof(sV[l * ldk + k]), acc);
      sO[part * hd + k] = acc;
      __syncthreads();
      if (threadIdx.x < hd) {
        float r = 0.f;
        for (int pp = 0; pp < parts; ++pp) r += sO[pp * hd + threadIdx.x];
        o[qrow * a.dim + h * hd + threadIdx.x] = fromf<T>(r * inv);
      }
    }
    if (threadIdx.x == 0) lse[qrow * a.heads + h] = mx + __logf(sum);
  }
}

template <class T>
__global__ void __launch_bounds__(128)
attn_core_bwd_kernel(const T* __restrict__ q, const T* __restrict__ kv, const T* __restrict__ d_o, const float* __restrict__ lse,
                     T* __restrict__ dq, T* __restrict__ dkv, AttnArgs a) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int hd = a.hd, ldk = hd + 2, L = a.L, nq = a.q_per_set;
  T* sK = reinterpret_cast<T*>(sm);
  T* sV = sK + L * ldk;
  float* sQ = reinterpret_cast<float*>(sV + L * ldk);     // [nq][hd]  (scaled by kappa)
  float* sDO = sQ + nq * hd;                              // [nq][hd]
  float* sP = sDO + nq * hd;                              // [nq][L]  probabilities
  float* sDS = sP + nq * L;                               // [nq][L]  dS
  float* scratch = sDS + nq * L;                          // [8]
  const int set = blockIdx.x / a.heads, h = blockIdx.x - set * a.heads;
  load_kv(kv, sK, sV, a, set, h);
  const float kappa = rsqrtf(static_cast<float>(hd));
  for (int idx = threadIdx.x; idx < nq * hd; idx += blockDim.x) {
    const int c = idx / hd, k = idx - c * hd;
    const size_t qrow = static_cast<size_t>(set) * a.q_stride_set + static_cast<size_t>(c) * a.q_stride_q;
    sQ[idx] = tof(q[qrow * a.dim + h * hd + k]) * kappa;
    sDO[idx] = tof(d_o[qrow * a.dim + h * hd + k]);
  }
  __syncthreads();
  for (int c = 0; c < nq; ++c) {
    const size_t qrow = static_cast<size_t>(set) * a.q_stride_set + static_cast<size_t>(c) * a.q_stride_q;
    const float l_c = lse[qrow * a.heads + h];
    float dot = 0.f;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
      float s = 0.f, dp = 0.f;
      for (int k = 0; k < hd; ++k) {
        s = fmaf(sQ[c * hd + k], tof(sK[l * ldk + k]), s);
        dp = fmaf(sDO[c * hd + k], tof(sV[l * ldk + k]), dp);
      }
      const float pr = __expf(s - l_c);
      sP[c * L + l] = pr;
      sDS[c * L + l] = dp;          // dP for now
      dot = fmaf(pr, dp, dot);
    }
    dot = block_reduce(dot, scratch, false);
    for (int l = threadIdx.x; l < L; l += blockDim.x) sDS[c * L + l] = sP[c * L + l] * (sDS[c * L + l] - dot);
    __syncthreads();
    // dQ[c][k] = kappa * sum_l dS[c][l] K[l][k]
    for (int k = threadIdx.x; k < hd; k += blockDim.x) {
      float acc = 0.f;
      for (int l = 0; l < L; ++l) acc = fmaf(sDS[c * L + l], tof(sK[l * ldk + k]), acc);
      dq[qrow * a.dim + h * hd + k] = fromf<T>(acc * kappa);
    }
  }
  __syncthreads();
  // dK[l][k] = sum_c dS[c][l] * (kappa q[c][k]);  dV[l][k] = sum_c P[c][l] dO[c][k]
  for (int idx = threadIdx.x; idx < L * hd; idx += blockDim.x) {
    const int l = idx / hd, k = idx - l * hd;
    float dk = 0.f, dv = 0.f;
    for (int c = 0; c < nq; ++c) {
      dk = fmaf(sDS[c * L + l], sQ[c * hd + k], dk);
      dv = fmaf(sP[c * L + l], sDO[c * hd + k], dv);
    }
    T* rowp = dkv + (static_cast<size_t>(set) * L + l) * (2 * a.dim) + h * hd + k;
    rowp[0] = fromf<T>(dk);
    rowp[a.dim] = fromf<T>(dv);
  }
}

// ------------------------------------------------------------------------------------------------
// residual add + L2 normalise (one warp per row)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
addnorm_fwd_kernel(const void* __restrict__ f, int f_dtype, const float* __restrict__ pooled, void* __restrict__ out,
                   float* __restrict__ inv_norm, int64_t rows, int dim) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[kLnMaxPerLane];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kLnMaxPerLane; ++k) {
    const int c = k * 32 + lane;
    // the reference adds in the feature dtype (model.py:379): round the sum to it before normalising
    float z = c < dim ? ld_elem(f, f_dtype, row * dim + c) + pooled[row * dim + c] : 0.f;
    if (f_dtype == COSMOS_DTYPE_BF16) z = __bfloat162float(__float2bfloat16_rn(z));
    else if (f_dtype == COSMOS_DTYPE_F16) z = __half2float(__float2half_rn(z));
    v[k] = z;
    s += z * z;
  }
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(s)), 1e-12f);
#pragma unroll
  for (int k = 0; k < kLnMaxPerLane; ++k) {
    const int c = k * 32 + lane;
    if (c < dim) st_elem(out, f_dtype, row * dim + c, v[k] * inv);
  }
  if (lane == 0) inv_norm[row] = inv;
}

__global__ void __launch_bounds__(256)
addnorm_bwd_kernel(const void* __restrict__ g_out, const void* __restrict__ out, int f_dtype, const float* __restrict__ inv_norm,
                   float* __restrict__ g_z32, void* __restrict__ g_z16, int g_dtype, int64_t rows, int dim) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float g[kLnMaxPerLane], y[kLnMaxPerLane];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kLnMaxPerLane; ++k) {
    const int c = k * 32 + lane;
    g[k] = c < dim ? ld_elem(g_out, f_dtype, row * dim + c) : 0.f;
    y[k] = c < dim ? ld_elem(out, f_dtype, row * dim + c) : 0.f;
    s += g[k] * y[k];
  }
  s = warp_sum(s);
  const float inv = inv_norm[row];
#pragma unroll
  for (int k = 0; k < kLnMaxPerLane; ++k) {
    const int c = k * 32 + lane;
    if (c < dim) {
      const float gz = (g[k] - y[k] * s) * inv;
      g_z32[row * dim + c] = gz;
      st_elem(g_z16, g_dtype, row * dim + c, gz);
    }
  }
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
                                 float* rstd, int64_t rows, int dim, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  layernorm_fwd_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, s>>>(x, x_dtype, w, b, y, y_dtype, mean, rstd, rows, dim);
  return cudaGetLastError();
}

cudaError_t launch_layernorm_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* w, const float* mean,
                                 const float* rstd, void* dx, int dx_dtype, int accumulate, float* dw, float* db, int64_t rows,
                                 int dim, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  const int rows_per_block = 64;
  layernorm_bwd_kernel<<<static_cast<unsigned>((rows + rows_per_block - 1) / rows_per_block), 256, 2 * dim * sizeof(float), s>>>(
      dy, dy_dtype, x, x_dtype, w, mean, rstd, dx, dx_dtype, accumulate, dw, db, rows, dim, rows_per_block);
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
  const size_t smem = 2 * static_cast<size_t>(L) * (a.hd + 2) * 2 + (a.hd + L + 8 + 128) * sizeof(float);
  cudaError_t e;
  if (dtype == COSMOS_DTYPE_BF16) {
    e = cudaFuncSetAttribute(attn_core_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attn_core_fwd_kernel<__nv_bfloat16><<<n_sets * heads, 128, smem, s>>>(
        static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(kv), static_cast<__nv_bfloat16*>(o), lse, a);
  } else {
    e = cudaFuncSetAttribute(attn_core_fwd_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attn_core_fwd_kernel<__half><<<n_sets * heads, 128, smem, s>>>(static_cast<const __half*>(q), static_cast<const __half*>(kv),
                                                                     static_cast<__half*>(o), lse, a);
  }
  return cudaGetLastError();
}

cudaError_t launch_attn_core_bwd(const void* q, const void* kv, const void* d_o, const float* lse, void* dq, void* dkv, int dtype,
                                 int n_sets, int L, int dim, int heads, int q_per_set, int64_t qs, int64_t qq, cudaStream_t s) {
  const AttnArgs a = make_attn_args(n_sets, L, dim, heads, q_per_set, qs, qq);
  const size_t smem = 2 * static_cast<size_t>(L) * (a.hd + 2) * 2 +
                      (2 * static_cast<size_t>(q_per_set) * a.hd + 2 * static_cast<size_t>(q_per_set) * L + 8) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e;
  if (dtype == COSMOS_DTYPE_BF16) {
    e = cudaFuncSetAttribute(attn_core_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attn_core_bwd_kernel<__nv_bfloat16><<<n_sets * heads, 128, smem, s>>>(
        static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(kv), static_cast<const __nv_bfloat16*>(d_o), lse,
        static_cast<__nv_bfloat16*>(dq), static_cast<__nv_bfloat16*>(dkv), a);
  } else {
    e = cudaFuncSetAttribute(attn_core_bwd_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attn_core_bwd_kernel<__half><<<n_sets * heads, 128, smem, s>>>(static_cast<const __half*>(q), static_cast<const __half*>(kv),
                                                                     static_cast<const __half*>(d_o), lse, static_cast<__half*>(dq),
                                                                     static_cast<__half*>(dkv), a);
  }
  return cudaGetLastError();
}

cudaError_t launch_addnorm_fwd(const void* f, int f_dtype, const float* pooled, void* out, float* inv_norm, int64_t rows, int dim,
                               cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  addnorm_fwd_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, s>>>(f, f_dtype, pooled, out, inv_norm, rows, dim);
  return cudaGetLastError();
}

cudaError_t launch_addnorm_bwd(const void* g_out, const void* out, int f_dtype, const float* inv_norm, float* g_z32, void* g_z16,
                               int g_dtype, int64_t rows, int dim, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  addnorm_bwd_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, s>>>(g_out, out, f_dtype, inv_norm, g_z32, g_z16, g_dtype, rows,
                                                                          dim);
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
