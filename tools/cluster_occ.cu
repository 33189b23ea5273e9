// How many clusters of 1/2/4/8 CTAs (232,448 B of dynamic shared memory, 384 threads: the footprint of
// infonce_bwd_quad_kernel) can be resident at once?  Cluster members must share a GPC, and the GPCs of a B200 do not all
// hold a multiple of 4 SMs, so cluster-4 launches may leave SMs idle.  Build: nvcc -arch=sm_100a -o tools/cluster_occ tools/cluster_occ.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void probe_kernel(int* out) {
  extern __shared__ int smem[];
  if (out != nullptr && threadIdx.x == 0) out[blockIdx.x] = smem[0];
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 232448;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  printf("SMs %d\n", sms);
  for (int cs = 1; cs <= 16; cs *= 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 1024);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
    printf("cluster size %2d: max active clusters %3d -> %3d CTAs resident of %d SMs (%s)\n", cs, n, n * cs, sms, cudaGetErrorString(e));
  }
  return 0;
}
