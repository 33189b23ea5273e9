// Retrieval ranks for the evaluation metrics.
//
// The reference (src/training/train.py:766-785 get_clip_metrics, 712-763 compute_retrieval) materialises the full
// similarity matrix on the CPU, argsorts every row and searches the position of the ground-truth item(s) in a Python
// loop.  The position of an item in a descending sort is the number of entries with a larger score, so
//     rank[r] = #{ c : <q_r, g_c>  >  max_{t in gt(r)} <q_r, g_t> }
// needs no sort and no matrix in memory: one thread per row forms the threshold, a tiled fp32 sweep counts.
//
// Arithmetic: fp32 on the CUDA cores (the reference's eval similarity is an fp32 CPU matmul, train.py:683,769; bf16
// tensor-core products would reorder near-ties).  Every dot product - in the threshold kernel and in the sweep - is the
// SAME chain acc = fmaf(q[k], g[k], acc) for k = 0 .. D-1, so the sweep reproduces the threshold bit for bit at the
// ground-truth column and the item never counts against itself.  logit_scale > 0 does not change the order and is not applied.
//
// Ties and NaN follow torch's sort order, the one the reference's `argsort(descending=True)` produces when it is stable:
// NaN sorts above every number (torch.sort's documented convention) and equal scores keep their index order, so
//     rank[r] = #{ c : s_c above s_t } + #{ c < t : s_c ties with s_t },   t = the best ground-truth column.
// A collapsed model (every score equal) or a diverged one (NaN features) therefore gets chance-level ranks, as it does
// in the reference - not rank 0 for every query, which a bare `>` count would report.
#include "common.cuh"
#include "internal.h"

namespace cb {

namespace {

constexpr int kRB = 128;        // rows and columns per CTA tile
constexpr int kRK = 16;         // k elements per shared-memory stage
constexpr int kRThreads = 256;  // 16 x 16 threads, 8 x 8 scores each
constexpr int kRLd = kRB + 4;   // padded leading dimension of the transposed stages

template <class T>
__device__ __forceinline__ float ld_f32(const T* p);
template <>
__device__ __forceinline__ float ld_f32<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld_f32<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <>
__device__ __forceinline__ float ld_f32<__half>(const __half* p) { return __half2float(*p); }

// Does (score a, column ca) come before (score b, column cb) in a stable descending sort where NaN is the largest value?
__device__ __forceinline__ bool sorts_before(float a, int ca, float b, int cb) {
  const bool na = a != a, nb = b != b;
  if (na != nb) return na;
  if (!na && a != b) return a > b;
  return ca < cb;
}

// best[r], best_col[r] = score and column of the ground-truth item of row r that sorts first (sequential fmaf chain, see the
// header)
template <class T>
__global__ void __launch_bounds__(128)
retrieval_best_kernel(const T* __restrict__ q, const T* __restrict__ g, int M, int N, int D, long long ldq, long long ldg,
                      const int* __restrict__ gt_offsets, const int* __restrict__ gt_index, float* __restrict__ best,
                      int* __restrict__ best_col) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  const int beg = gt_offsets != nullptr ? gt_offsets[r] : r;
  const int end = gt_offsets != nullptr ? gt_offsets[r + 1] : r + 1;
  const T* qr = q + static_cast<size_t>(r) * ldq;
  float b = -INFINITY;
  int bt = N;      // no ground truth: threshold -inf at a column past the gallery, so every score counts and rank = N
  for (int e = beg; e < end; ++e) {
    const int t = gt_index != nullptr ? gt_index[e] : e;
    if (t < 0 || t >= N) continue;          // validated on the host side of the Python wrapper; never trusted here
    const T* gr = g + static_cast<size_t>(t) * ldg;
    float acc = 0.f;
    for (int k = 0; k < D; ++k) acc = fmaf(ld_f32(qr + k), ld_f32(gr + k), acc);
    // the ground-truth item that sorts first: higher score (NaN highest), then lower index
    if (bt == N || sorts_before(acc, t, b, bt)) { b = acc; bt = t; }
  }
  best[r] = b;
  best_col[r] = bt;
}

// One stage of a [128 x 16] operand block: thread `tid` owns row tid / 2 and the 8 consecutive k of half tid % 2.
template <class T>
__device__ __forceinline__ void load_stage(const T* __restrict__ base, long long ld, int row0, int n_rows, int k0, int D, int tid,
                                           float (&v)[8]) {
  const int row = row0 + (tid >> 1);
  const int k = k0 + (tid & 1) * 8;
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = 0.f;
  if (row >= n_rows) return;
  const T* p = base + static_cast<size_t>(row) * ld + k;
#pragma unroll
  for (int e = 0; e < 8; ++e)
    if (k + e < D) v[e] = ld_f32(p + e);
}

__device__ __forceinline__ void store_stage(float* __restrict__ s, int tid, const float (&v)[8]) {
  const int row = tid >> 1;
  const int k = (tid & 1) * 8;
#pragma unroll
  for (int e = 0; e < 8; ++e) s[(k + e) * kRLd + row] = v[e];
}

// counts[r] += #{ c in this CTA's 128 columns : (<q_r, g_c>, c) sorts before (best[r], best_col[r]) }
template <class T>
__global__ void __launch_bounds__(kRThreads, 2)
retrieval_count_kernel(const T* __restrict__ q, const T* __restrict__ g, int M, int N, int D, long long ldq, long long ldg,
                       const float* __restrict__ best, const int* __restrict__ best_col, int* __restrict__ counts) {
  __shared__ __align__(16) float sQ[kRK * kRLd];
  __shared__ __align__(16) float sG[kRK * kRLd];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.y * kRB, col0 = blockIdx.x * kRB;

  // two neighbouring columns per accumulator register pair: ONE FFMA2 (sm_100's packed fp32 pair, each half rounded like fmaf)
  // per two scores - the plain three-register FFMA issues every other cycle, which held this kernel at half of the fp32 peak
  uint64_t acc2[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc2[i][j] = f2_pack(0.f, 0.f);

  float vq[8], vg[8];
  load_stage(q, ldq, row0, M, 0, D, tid, vq);
  load_stage(g, ldg, col0, N, 0, D, tid, vg);
  for (int k0 = 0; k0 < D; k0 += kRK) {
    store_stage(sQ, tid, vq);
    store_stage(sG, tid, vg);
    __syncthreads();
    if (k0 + kRK < D) {      // next stage's global loads fly while this one is consumed
      load_stage(q, ldq, row0, M, k0 + kRK, D, tid, vq);
      load_stage(g, ldg, col0, N, k0 + kRK, D, tid, vg);
    }
    // k past D is zero-filled: fmaf(0, 0, acc) leaves the chain's value unchanged
#pragma unroll
    for (int kk = 0; kk < kRK; ++kk) {
      // rows ty*4 .. +3 and 64 + ty*4 .. +3; columns tx*4 .. +3 and 64 + tx*4 .. +3 (conflict-free 16-byte reads)
      const float4 a0 = *reinterpret_cast<const float4*>(sQ + kk * kRLd + ty * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(sQ + kk * kRLd + 64 + ty * 4);
      const float4 b0 = *reinterpret_cast<const float4*>(sG + kk * kRLd + tx * 4);
      const float4 b1 = *reinterpret_cast<const float4*>(sG + kk * kRLd + 64 + tx * 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const uint64_t b2[4] = {f2_pack(b0.x, b0.y), f2_pack(b0.z, b0.w), f2_pack(b1.x, b1.y), f2_pack(b1.z, b1.w)};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint64_t ai = f2_pack(a[i], a[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc2[i][j] = f2_fma(ai, b2[j], acc2[i][j]);
      }
    }
    __syncthreads();
  }
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) f2_unpack(acc2[i][j], acc[i][2 * j], acc[i][2 * j + 1]);

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    const float thr = row < M ? __ldg(best + row) : INFINITY;
    const int tcol = row < M ? __ldg(best_col + row) : -1;
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = col0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      cnt += (col < N && sorts_before(acc[i][j], col, thr, tcol)) ? 1 : 0;
    }
    // the 16 threads that share this row are the 16 lanes of one half warp
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 8);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 4);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
    if (tx == 0 && row < M && cnt != 0) atomicAdd(counts + row, cnt);   // integer: the result does not depend on the order
  }
}

template <class T>
cudaError_t launch_t(const void* q, const void* g, int M, int N, int D, long long ldq, long long ldg, const int* gt_offsets,
                     const int* gt_index, float* best, int* best_col, int* ranks, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(ranks, 0, static_cast<size_t>(M) * sizeof(int), stream);
  if (e != cudaSuccess) return e;
  retrieval_best_kernel<T><<<(M + 127) / 128, 128, 0, stream>>>(static_cast<const T*>(q), static_cast<const T*>(g), M, N, D, ldq, ldg,
                                                                gt_offsets, gt_index, best, best_col);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  dim3 grid((N + kRB - 1) / kRB, (M + kRB - 1) / kRB);
  retrieval_count_kernel<T><<<grid, kRThreads, 0, stream>>>(static_cast<const T*>(q), static_cast<const T*>(g), M, N, D, ldq, ldg,
                                                            best, best_col, ranks);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_retrieval_ranks(const void* q, const void* g, int dtype, int M, int N, int D, long long ldq, long long ldg,
                                   const int* gt_offsets, const int* gt_index, float* best, int* best_col, int* ranks,
                                   cudaStream_t stream) {
  if (dtype == COSMOS_DTYPE_F32) return launch_t<float>(q, g, M, N, D, ldq, ldg, gt_offsets, gt_index, best, best_col, ranks, stream);
  if (dtype == COSMOS_DTYPE_BF16)
    return launch_t<__nv_bfloat16>(q, g, M, N, D, ldq, ldg, gt_offsets, gt_index, best, best_col, ranks, stream);
  return launch_t<__half>(q, g, M, N, D, ldq, ldg, gt_offsets, gt_index, best, best_col, ranks, stream);
}

}  // namespace cb
