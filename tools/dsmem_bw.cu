// Microbenchmark: sustained st.async (DSMEM) bandwidth when every SM of the GPU streams 32 KB tiles to a cluster sibling,
// in the access pattern of the backward kernel's G exchange (lane = row, 16-byte swizzled chunks).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/dsmem_bw tools/dsmem_bw.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../cosmos_b200/csrc/common.cuh"
using namespace cb;

__global__ void __launch_bounds__(288, 1) k(int iters, int contiguous, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];     // two 32 KB receive slots
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 200 * 1024);   // [2] bytes of a tile have landed in slot s
  uint64_t* credit = full + 2;                                         // [2] (sender side) the sibling has consumed slot s
  const uint32_t rank = cluster_ctarank();
  const uint32_t sibling = rank ^ 2u;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  cluster_sync_all();
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&full[s], 1); mbar_init(&credit[s], 1); }
    fence_mbar_init();
  }
  cluster_sync_all();
  const long long t0 = clock64();
  if (warp < 8) {                    // senders
    const uint32_t q = warp & 3, h = warp >> 2;
    const int r_t = q * 32 + lane;
    const uint32_t dst_row = mapa_u32(smem_u32(smem) + h * 16384 + r_t * 128, sibling);
    const uint32_t dst_lin = mapa_u32(smem_u32(smem) + warp * 4096 + lane * 16, sibling);
    const uint32_t rbar = mapa_u32(smem_u32(full), sibling);
    const uint32_t dst_base = mapa_u32(smem_u32(smem), sibling);
    for (int it = 0; it < iters; ++it) {
      const uint32_t s = it & 1;
      mbar_wait(&credit[s], ((it >> 1) & 1) ^ 1);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        // pattern 0: lane = row, 16 B per row (32 rows per warp store); 1: 512 contiguous bytes per warp store;
        // 2: two lanes per row (16 rows x 32 B); 3: four lanes per row (8 rows x 64 B)
        uint32_t a;
        if (contiguous == 0) a = dst_row + ((jj ^ (r_t & 7)) << 4);
        else if (contiguous == 1) a = dst_lin + jj * 512;
        else if (contiguous == 2) {
          const int row = q * 32 + (jj >> 2) * 16 + (lane >> 1), c = (jj & 3) * 2 + (lane & 1);
          a = dst_base + h * 16384 + row * 128 + ((c ^ (row & 7)) << 4);
        } else {
          const int row = q * 32 + (jj >> 1) * 8 + (lane >> 2), c = (jj & 1) * 4 + (lane & 3);
          a = dst_base + h * 16384 + row * 128 + ((c ^ (row & 7)) << 4);
        }
        a += s * 32768;
        st_async_cluster_v4(a, make_uint4(it, jj, tid, 0), rbar + s * 8);
      }
    }
  } else if (lane == 0) {            // receiver
    const uint32_t rcredit = mapa_u32(smem_u32(credit), sibling);
    mbar_expect_tx(&full[0], 32768);
    mbar_expect_tx(&full[1], 32768);
    for (int it = 0; it < iters; ++it) {
      const uint32_t s = it & 1;
      mbar_wait(&full[s], (it >> 1) & 1);
      if (it + 2 < iters) mbar_expect_tx(&full[s], 32768);
      mbar_arrive_remote_release(rcredit + s * 8);
    }
  }
  const long long t1 = clock64();
  cluster_sync_all();
  if (tid == 32 && blockIdx.x == 8) out[0] = t1 - t0;
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  for (int contiguous = 0; contiguous < 4; ++contiguous) {
    const int iters = 2000, smem = 201 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(288); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaLaunchKernelEx(&cfg, k, 10, contiguous, d);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, k, iters, contiguous, d);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long clk; cudaMemcpy(&clk, d, 8, cudaMemcpyDeviceToHost);
    printf("%s: launch %s, %d tiles of 32 KB per SM: %.3f ms, sender clocks %lld -> %.2f B/clk/SM, %.2f TB/s chip (148 SMs)\n",
           contiguous == 0 ? "32 rows x 16 B per warp store (kernel pattern)" : contiguous == 1 ? "512 contiguous bytes per warp store" : contiguous == 2 ? "16 rows x 32 B per warp store" : "8 rows x 64 B per warp store", cudaGetErrorString(err), iters, ms,
           clk, double(iters) * 32768 / double(clk), 148.0 * iters * 32768 / (ms * 1e-3) * 1e-12);
  }
  return 0;
}
