// Small HBM-bound helpers around the InfoNCE tile kernels: merge of per-row-tile column statistics,
// per-pair loss sums, and the deterministic reduction of the per-CTA dscale partials.
#include "common.cuh"
#include "infonce.h"

namespace cb {

// col_part [pairs][n_slabs][n_cols] (max2, sum) -> col_lse2 [pairs][n_cols]; coalesced over columns.
__global__ void __launch_bounds__(256)
col_combine_kernel(const float2* __restrict__ col_part, float* __restrict__ col_lse2, int n_slabs, int n_cols) {
  const int pair = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  const float2* src = col_part + static_cast<size_t>(pair) * n_slabs * n_cols + c;
  float m = -INFINITY, l = 0.f;
  for (int s = 0; s < n_slabs; ++s) {
    const float2 v = __ldcs(src + static_cast<size_t>(s) * n_cols);
    if (v.x > m) {
      l *= ex2(m - v.x);
      m = v.x;
    }
    if (v.x != -INFINITY) l += v.y * ex2(v.x - m);
  }
  col_lse2[static_cast<size_t>(pair) * n_cols + c] = m + log2f(l);
}

cudaError_t launch_col_combine(const float2* col_part, float* col_lse2, int pairs, int n_slabs, int n_cols, cudaStream_t stream) {
  dim3 grid((n_cols + 255) / 256, pairs);
  col_combine_kernel<<<grid, 256, 0, stream>>>(col_part, col_lse2, n_slabs, n_cols);
  return cudaGetLastError();
}

// parts [n_parts][n] -> out [n]: log2-sum-exp2 over the parts (the per-rank partial column statistics of a sharded forward,
// all-gathered: ONE collective and one pass instead of a MAX all-reduce, an exp, a SUM all-reduce and a log).
__global__ void __launch_bounds__(256)
lse2_merge_kernel(const float* __restrict__ parts, float* __restrict__ out, int n_parts, size_t n) {
  const size_t k = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (k >= n) return;
  float m = -INFINITY;
  for (int q = 0; q < n_parts; ++q) m = fmaxf(m, __ldcs(parts + static_cast<size_t>(q) * n + k));
  if (m == -INFINITY || m == INFINITY) {
    out[k] = m;
    return;
  }
  float l = 0.f;
  for (int q = 0; q < n_parts; ++q) l += ex2(__ldcs(parts + static_cast<size_t>(q) * n + k) - m);
  out[k] = m + log2f(l);
}

cudaError_t launch_lse2_merge(const float* parts, float* out, int n_parts, size_t n, cudaStream_t stream) {
  lse2_merge_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(parts, out, n_parts, n);
  return cudaGetLastError();
}

// out[pair][0] = sum_r (ln2 * row_lse2[r] - scale * diag[r])
// out[pair][1] = sum_r (ln2 * col_lse2[label_offset + r] - scale * diag[r])      (this rank's diagonal columns)
__global__ void __launch_bounds__(256)
loss_sums_kernel(const float* __restrict__ row_lse2, const float* __restrict__ diag_raw, const float* __restrict__ col_lse2,
                 const float* __restrict__ scale_ptr, int n_rows, int n_cols, int label_offset, int use_rows, int use_cols,
                 float* __restrict__ out) {
  const int pair = blockIdx.x;
  const float scale = __ldg(scale_ptr);
  double sr = 0.0, sc = 0.0;
  for (int r = threadIdx.x; r < n_rows; r += blockDim.x) {
    const float d = scale * diag_raw[static_cast<size_t>(pair) * n_rows + r];
    if (use_rows) sr += static_cast<double>(kLn2 * row_lse2[static_cast<size_t>(pair) * n_rows + r] - d);
    if (use_cols) sc += static_cast<double>(kLn2 * col_lse2[static_cast<size_t>(pair) * n_cols + label_offset + r] - d);
  }
  __shared__ double red[2][8];
  for (int o = 16; o > 0; o >>= 1) {
    sr += __shfl_xor_sync(0xffffffffu, sr, o);
    sc += __shfl_xor_sync(0xffffffffu, sc, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = sr;
    red[1][threadIdx.x >> 5] = sc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < 8; ++w) {
      a += red[0][w];
      b += red[1][w];
    }
    out[2 * pair + 0] = static_cast<float>(a);
    out[2 * pair + 1] = static_cast<float>(b);
  }
}

cudaError_t launch_loss_sums(const float* row_lse2, const float* diag_raw, const float* col_lse2, const float* scale, int pairs,
                             int n_rows, int n_cols, int label_offset, int use_rows, int use_cols, float* out,
                             cudaStream_t stream) {
  loss_sums_kernel<<<pairs, 256, 0, stream>>>(row_lse2, diag_raw, col_lse2, scale, n_rows, n_cols, label_offset, use_rows,
                                              use_cols, out);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
dscale_reduce_kernel(const float* __restrict__ part, int n, float weight, const float* __restrict__ upstream,
                     float* __restrict__ dscale) {
  double s = 0.0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) s += static_cast<double>(part[k]);
  __shared__ double red[8];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0;
    for (int w = 0; w < 8; ++w) a += red[w];
    *dscale = static_cast<float>(a * static_cast<double>(weight) * static_cast<double>(__ldg(upstream)));
  }
}

cudaError_t launch_dscale_reduce(const float* part, int n, float weight, const float* upstream, float* dscale,
                                 cudaStream_t stream) {
  dscale_reduce_kernel<<<1, 256, 0, stream>>>(part, n, weight, upstream, dscale);
  return cudaGetLastError();
}


__global__ void __launch_bounds__(256)
convert_dx_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, int fmt, size_t n4) {
  for (size_t k = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; k < n4; k += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 v = __ldcs(src + k);
    dst[k] = make_uint2(pack2(v.x, v.y, fmt), pack2(v.z, v.w, fmt));
  }
}

// dst[k] = sum over the n_parts slices of parts[q][k], rounded once to the stack dtype (partial dX of a sliced column sweep)
__global__ void __launch_bounds__(256)
reduce_dx_kernel(const float4* __restrict__ parts, int n_parts, uint2* __restrict__ dst, int fmt, size_t n4) {
  for (size_t k = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; k < n4; k += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 a = __ldcs(parts + k);
    for (int q = 1; q < n_parts; ++q) {
      const float4 b = __ldcs(parts + static_cast<size_t>(q) * n4 + k);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    dst[k] = make_uint2(pack2(a.x, a.y, fmt), pack2(a.z, a.w, fmt));
  }
}

cudaError_t launch_reduce_dx(const float* parts, int n_parts, void* dst, int dtype, size_t n, cudaStream_t stream) {
  reduce_dx_kernel<<<148 * 8, 256, 0, stream>>>(reinterpret_cast<const float4*>(parts), n_parts, reinterpret_cast<uint2*>(dst),
                                                dtype == 1 ? 1 : 0, n / 4);      // n is a multiple of 512
  return cudaGetLastError();
}

// dst[k] = src[k] * (*num / den), 16-bit in and out, product formed in fp32 and rounded once (the unit gradients of the
// stored-exponential route times upstream / stand-in: one vectorised pass instead of ATen's mixed-dtype elementwise kernel)
__global__ void __launch_bounds__(256)
scale16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, const float* __restrict__ num, float den, int fmt, size_t n8) {
  const float f = __ldg(num) / den;
  for (size_t k = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; k < n8; k += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 v = __ldcs(src + k);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float a, b;
      if (fmt) {
        a = __uint_as_float(w[q] << 16);
        b = __uint_as_float(w[q] & 0xffff0000u);
      } else {
        const __half2 h = *reinterpret_cast<const __half2*>(&w[q]);
        a = __low2float(h);
        b = __high2float(h);
      }
      o[q] = pack2(a * f, b * f, fmt);
    }
    dst[k] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

cudaError_t launch_scale16(const void* src, void* dst, const float* num, float den, int dtype, size_t n, cudaStream_t stream) {
  scale16_kernel<<<148 * 8, 256, 0, stream>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), num, den,
                                              dtype == 1 ? 1 : 0, n / 8);
  return cudaGetLastError();
}

cudaError_t launch_convert_dx(const float* src, void* dst, int dtype, size_t n, cudaStream_t stream) {
  const size_t n4 = n / 4;   // n is a multiple of 64
  convert_dx_kernel<<<148 * 8, 256, 0, stream>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(dst),
                                                 dtype == 1 ? 1 : 0, n4);
  return cudaGetLastError();
}

}  // namespace cb
