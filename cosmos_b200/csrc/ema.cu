// EMA teacher update: k <- k*m + (1-m)*q over every parameter tensor in ONE launch.
// Replaces the per-parameter loop of the reference (src/training/train.py:195-203: 3 elementwise
// kernels + 1 temporary per tensor, 969 launches, 28 B/param) by a chunk-table kernel that moves
// the algorithmic 12 B/param (read k, read q, write k) with 128-bit accesses.
//
// Rounding follows the reference expression exactly: fl(fl(k*m) + fl((1-m)*q)), with m and (1-m)
// rounded to fp32 from Python doubles - no FMA contraction - so fp32 results are bit-identical.
#include "common.cuh"
#include "internal.h"

namespace cb {

constexpr int kEmaThreads = 256;
constexpr int kEmaChunk = COSMOS_EMA_CHUNK;          // elements per table entry
constexpr int kEmaVecPerThread = kEmaChunk / (kEmaThreads * 4);  // float4 per thread per chunk

__device__ __forceinline__ float ema1(float k, float q, float m, float om) {
  return __fadd_rn(__fmul_rn(k, m), __fmul_rn(q, om));
}

__global__ void __launch_bounds__(kEmaThreads)
ema_f32_kernel(const EmaChunk* __restrict__ table, int n_chunks, float m, float om) {
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const EmaChunk ch = table[c];
    float* __restrict__ k = reinterpret_cast<float*>(ch.teacher);
    const float* __restrict__ q = reinterpret_cast<const float*>(ch.student);
    const uint32_t n = ch.count;
    if (n == kEmaChunk && ch.aligned) {
      float4 kv[kEmaVecPerThread], qv[kEmaVecPerThread];
#pragma unroll
      for (int i = 0; i < kEmaVecPerThread; ++i) {
        const int idx = (i * kEmaThreads + threadIdx.x);
        kv[i] = __ldcs(reinterpret_cast<const float4*>(k) + idx);
        qv[i] = __ldcs(reinterpret_cast<const float4*>(q) + idx);
      }
#pragma unroll
      for (int i = 0; i < kEmaVecPerThread; ++i) {
        const int idx = (i * kEmaThreads + threadIdx.x);
        float4 r;
        r.x = ema1(kv[i].x, qv[i].x, m, om);
        r.y = ema1(kv[i].y, qv[i].y, m, om);
        r.z = ema1(kv[i].z, qv[i].z, m, om);
        r.w = ema1(kv[i].w, qv[i].w, m, om);
        __stcs(reinterpret_cast<float4*>(k) + idx, r);
      }
    } else if (ch.aligned) {
      const uint32_t n4 = n >> 2;
      for (uint32_t i = threadIdx.x; i < n4; i += kEmaThreads) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(k) + i);
        const float4 b = __ldcs(reinterpret_cast<const float4*>(q) + i);
        float4 r;
        r.x = ema1(a.x, b.x, m, om);
        r.y = ema1(a.y, b.y, m, om);
        r.z = ema1(a.z, b.z, m, om);
        r.w = ema1(a.w, b.w, m, om);
        __stcs(reinterpret_cast<float4*>(k) + i, r);
      }
      for (uint32_t i = (n4 << 2) + threadIdx.x; i < n; i += kEmaThreads) k[i] = ema1(k[i], q[i], m, om);
    } else {
      for (uint32_t i = threadIdx.x; i < n; i += kEmaThreads) k[i] = ema1(k[i], q[i], m, om);
    }
  }
}

// 16-bit parameters (pure_bf16 / pure_fp16 precisions): every intermediate is rounded to the
// storage type, as the three reference kernels do.
template <class T>
__device__ __forceinline__ float rnd16(float x);
template <>
__device__ __forceinline__ float rnd16<__nv_bfloat16>(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
template <>
__device__ __forceinline__ float rnd16<__half>(float x) { return __half2float(__float2half_rn(x)); }
template <class T>
__device__ __forceinline__ float to_f(T x);
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <>
__device__ __forceinline__ float to_f<__half>(__half x) { return __half2float(x); }
template <class T>
__device__ __forceinline__ T from_f(float x);
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
template <>
__device__ __forceinline__ __half from_f<__half>(float x) { return __float2half_rn(x); }

template <class T>
__global__ void __launch_bounds__(kEmaThreads)
ema_16_kernel(const EmaChunk* __restrict__ table, int n_chunks, float m, float om) {
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const EmaChunk ch = table[c];
    T* __restrict__ k = reinterpret_cast<T*>(ch.teacher);
    const T* __restrict__ q = reinterpret_cast<const T*>(ch.student);
    const uint32_t n = ch.count;
    uint32_t done = 0;
    if (ch.aligned) {
      const uint32_t n8 = n >> 3;
      for (uint32_t i = threadIdx.x; i < n8; i += kEmaThreads) {
        uint4 a = __ldcs(reinterpret_cast<const uint4*>(k) + i);
        const uint4 b = __ldcs(reinterpret_cast<const uint4*>(q) + i);
        T* ae = reinterpret_cast<T*>(&a);
        const T* be = reinterpret_cast<const T*>(&b);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float t1 = rnd16<T>(__fmul_rn(to_f<T>(ae[e]), m));
          const float t2 = rnd16<T>(__fmul_rn(to_f<T>(be[e]), om));
          ae[e] = from_f<T>(__fadd_rn(t1, t2));
        }
        __stcs(reinterpret_cast<uint4*>(k) + i, a);
      }
      done = n8 << 3;
    }
    for (uint32_t i = done + threadIdx.x; i < n; i += kEmaThreads) {
      const float t1 = rnd16<T>(__fmul_rn(to_f<T>(k[i]), m));
      const float t2 = rnd16<T>(__fmul_rn(to_f<T>(q[i]), om));
      k[i] = from_f<T>(__fadd_rn(t1, t2));
    }
  }
}

// Clamp of the logit-scale scalars (src/training/train.py:237-243: four 1-element clamp_ launches) in one launch.
// torch.clamp_ semantics: min(max(x, lo), hi) with the bounds rounded to fp32, NaN kept.
template <class T>
__global__ void clamp_scalars_kernel(ClampTable t, float lo, float hi) {
  const int i = threadIdx.x;
  if (i >= t.n) return;
  T* ptr = reinterpret_cast<T*>(t.ptr[i]);
  const float x = static_cast<float>(*ptr);
  if (x != x) return;
  *ptr = static_cast<T>(fminf(fmaxf(x, lo), hi));   // the expression ATen's CUDA clamp evaluates (in fp32 for 16-bit types)
}

cudaError_t launch_clamp_scalars(const ClampTable& t, double lo, double hi, int dtype, cudaStream_t stream) {
  if (t.n <= 0) return cudaSuccess;
  const float flo = static_cast<float>(lo), fhi = static_cast<float>(hi);
  if (dtype == COSMOS_DTYPE_F32) clamp_scalars_kernel<float><<<1, 32, 0, stream>>>(t, flo, fhi);
  else if (dtype == COSMOS_DTYPE_BF16) clamp_scalars_kernel<__nv_bfloat16><<<1, 32, 0, stream>>>(t, flo, fhi);
  else clamp_scalars_kernel<__half><<<1, 32, 0, stream>>>(t, flo, fhi);
  return cudaGetLastError();
}

cudaError_t launch_ema(const EmaChunk* table, int n_chunks, double momentum, int dtype, int sm_count, cudaStream_t stream) {
  if (n_chunks <= 0) return cudaSuccess;
  const float m = static_cast<float>(momentum);
  const float om = static_cast<float>(1.0 - momentum);
  // 8 resident CTAs of 256 threads per SM; a multiple of the SM count
  int grid = sm_count * 8;
  if (grid > n_chunks) grid = n_chunks;
  if (dtype == COSMOS_DTYPE_F32) {
    ema_f32_kernel<<<grid, kEmaThreads, 0, stream>>>(table, n_chunks, m, om);
  } else if (dtype == COSMOS_DTYPE_BF16) {
    ema_16_kernel<__nv_bfloat16><<<grid, kEmaThreads, 0, stream>>>(table, n_chunks, m, om);
  } else {
    ema_16_kernel<__half><<<grid, kEmaThreads, 0, stream>>>(table, n_chunks, m, om);
  }
  return cudaGetLastError();
}

}  // namespace cb
