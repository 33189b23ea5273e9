// Stand-alone primitive self-test for the tcgen05 / TMA building blocks the loss-head kernels use.
// Each case runs ONE CTA on one tile and is compared with a double-precision CPU product:
//   case 0  SS  A K-major (TMA)            x B K-major (TMA)            M128 N256 K512   (logits GEMM)
//   case 1  SS  A K-major (thread-written) x B MN-major (TMA), 4 x N64  M128 N256 K128   (grad GEMM, per slab)
//   case 2  SS  same, one N=256 chain with LBO = slab stride
//   case 3  TS  A in TMEM (tcgen05.st)     x B MN-major (TMA), N=256                      (grad GEMM, A from TMEM)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o selftest selftest.cu   (tests/ runs it on the GPU box)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "common.cuh"
#include "tma_host.h"

using namespace cb;

struct Params {
  int mode;
  float* out;                 // [128][256] fp32
  const __nv_bfloat16* gmat;  // [128][128] row-major (modes 1-3)
};

__global__ void __launch_bounds__(128, 1)
selftest_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;               // 32 KB : A slab(s)
  uint8_t* sB = smem + 32768;       // 64 KB : B
  __shared__ uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_slot;

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = (warp * 32u) << 16;

  if (p.mode == 0) {
    // D[128x256] = A[128x512] * B[256x512]^T, one 64-wide K slab at a time through a single stage
    const uint32_t idesc = make_idesc(1, 0, 0, 128, 256);
    uint32_t ph = 0;
    for (int s = 0; s < 8; ++s) {
      if (tid == 0) {
        mbar_expect_tx(&full_bar, 16384 + 32768);
        tma_load_3d(sA, &tmA, &full_bar, s * 64, 0, 0);
        tma_load_3d(sB, &tmB, &full_bar, s * 64, 0, 0);
        mbar_wait(&full_bar, ph);
        tc_fence_after();
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t da = make_smem_desc(smem_u32(sA) + kk * 32, 0, 1024);
          const uint64_t db = make_smem_desc(smem_u32(sB) + kk * 32, 0, 1024);
          umma_ss(tmem, da, db, idesc, (s | kk) != 0);
        }
        tc_commit(&done_bar);
        mbar_wait(&done_bar, ph);
      }
      ph ^= 1;
      __syncthreads();
    }
  } else {
    // B = Y tile [128 cols (K)] x [256 d (N)], 4 TMA slabs of [128 x 64] at 16 KB stride
    if (tid == 0) {
      mbar_expect_tx(&full_bar, 65536);
      for (int s = 0; s < 4; ++s) tma_load_3d(sB + s * 16384, &tmB, &full_bar, s * 64, 0, 0);
    }
    if (p.mode == 1 || p.mode == 2) {
      // A = G [128 rows][128 cols] bf16 -> two K-major SW128 slabs, written by its row's thread
      const int r = tid;
      for (int c8 = 0; c8 < 16; ++c8) {  // 16-byte chunks of 8 columns
        const uint4 v = *reinterpret_cast<const uint4*>(p.gmat + r * 128 + c8 * 8);
        const int slab = c8 >> 3, j = c8 & 7;
        *reinterpret_cast<uint4*>(sA + slab * 16384 + r * 128 + ((j ^ (r & 7)) << 4)) = v;
      }
      fence_proxy_async_smem();
    } else {
      // A -> TMEM columns [256, 320): lane = row, column k holds (G[r][2k], G[r][2k+1])
      const int r = tid;
      for (int blk = 0; blk < 4; ++blk) {
        uint32_t regs[16];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.gmat + r * 128 + blk * 32);
#pragma unroll
        for (int k = 0; k < 16; ++k) regs[k] = src[k];
        tmem_st16(tmem + lane_base + 256 + blk * 16, regs);
      }
      tmem_st_wait();
      tc_fence_before();
    }
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(&full_bar, 0);
      tc_fence_after();
      if (p.mode == 1) {
        const uint32_t idesc = make_idesc(1, 0, 1, 128, 64);
        for (int s = 0; s < 4; ++s)
          for (int k = 0; k < 8; ++k) {  // K = 128 cols, 16 per MMA
            const uint64_t da = make_smem_desc(smem_u32(sA) + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024);
            const uint64_t db = make_smem_desc(smem_u32(sB) + s * 16384 + k * 2048, 16384, 1024);
            umma_ss(tmem + s * 64, da, db, idesc, k != 0);
          }
      } else if (p.mode == 2) {
        const uint32_t idesc = make_idesc(1, 0, 1, 128, 256);
        for (int k = 0; k < 8; ++k) {
          const uint64_t da = make_smem_desc(smem_u32(sA) + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024);
          const uint64_t db = make_smem_desc(smem_u32(sB) + k * 2048, 16384, 1024);
          umma_ss(tmem, da, db, idesc, k != 0);
        }
      } else {
        const uint32_t idesc = make_idesc(1, 0, 1, 128, 256);
        for (int k = 0; k < 8; ++k) {
          const uint64_t db = make_smem_desc(smem_u32(sB) + k * 2048, 16384, 1024);
          umma_ts(tmem, tmem + 256 + k * 8, db, idesc, k != 0);
        }
      }
      tc_commit(&done_bar);
      mbar_wait(&done_bar, 0);
    }
    __syncthreads();
  }
  tc_fence_after();
  // epilogue: TMEM -> global
  for (int c = 0; c < 256; c += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + lane_base + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 32; ++k) p.out[tid * 256 + c + k] = __uint_as_float(v[k]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  int dev = 0;
  cudaSetDevice(dev);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, dev);
  printf("device: %s sm_%d%d\n", prop.name, prop.major, prop.minor);
  const int smem_bytes = 32768 + 65536 + 1024;
  cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);

  srand(7);
  auto rnd = [] { return (rand() / (float)RAND_MAX) * 2.f - 1.f; };
  std::vector<float> A(128 * 512), B(256 * 512), G(128 * 128), Y(128 * 256);
  for (auto& x : A) x = bf(rnd());
  for (auto& x : B) x = bf(rnd());
  for (auto& x : G) x = bf(rnd());
  for (auto& x : Y) x = bf(rnd());
  auto to_dev = [](const std::vector<float>& h) {
    std::vector<__nv_bfloat16> t(h.size());
    for (size_t i = 0; i < h.size(); ++i) t[i] = __float2bfloat16(h[i]);
    __nv_bfloat16* d;
    cudaMalloc(&d, t.size() * 2);
    cudaMemcpy(d, t.data(), t.size() * 2, cudaMemcpyHostToDevice);
    return d;
  };
  __nv_bfloat16 *dA = to_dev(A), *dB = to_dev(B), *dG = to_dev(G), *dY = to_dev(Y);
  float* dOut;
  cudaMalloc(&dOut, 128 * 256 * 4);

  CUtensorMap mA, mB, mY;
  int e1 = make_stack_map(&mA, dA, 1, 512, 128, 1, 128);
  int e2 = make_stack_map(&mB, dB, 1, 512, 256, 1, 256);
  int e3 = make_stack_map(&mY, dY, 1, 256, 128, 1, 128);
  if (e1 || e2 || e3) {
    printf("tensor map encode failed %d %d %d\n", e1, e2, e3);
    return 2;
  }

  std::vector<double> ref0(128 * 256), ref1(128 * 256);
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 256; ++j) {
      double s = 0;
      for (int k = 0; k < 512; ++k) s += (double)A[i * 512 + k] * B[j * 512 + k];
      ref0[i * 256 + j] = s;
      double t = 0;
      for (int k = 0; k < 128; ++k) t += (double)G[i * 128 + k] * Y[k * 256 + j];
      ref1[i * 256 + j] = t;
    }

  int fails = 0;
  std::vector<float> out(128 * 256);
  for (int mode = 0; mode < 4; ++mode) {
    cudaMemset(dOut, 0xff, 128 * 256 * 4);
    Params p{mode, dOut, dG};
    selftest_kernel<<<1, 128, smem_bytes>>>(mode == 0 ? mA : mA, mode == 0 ? mB : mY, p);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
      printf("case %d: CUDA error %s\n", mode, cudaGetErrorString(err));
      return 3;
    }
    cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
    const std::vector<double>& ref = mode == 0 ? ref0 : ref1;
    double maxerr = 0;
    int bad = 0;
    for (size_t i = 0; i < out.size(); ++i) {
      double e = fabs(out[i] - ref[i]);
      if (!(e <= 1e-3 * (1 + fabs(ref[i])))) ++bad;
      if (e > maxerr || e != e) maxerr = e;
    }
    printf("case %d: %s max_abs_err %.3e bad %d  (out[0]=%f ref[0]=%f out[last]=%f ref[last]=%f)\n", mode, bad ? "FAIL" : "PASS",
           maxerr, bad, out[0], ref[0], out.back(), ref.back());
    fails += bad != 0;
  }
  printf(fails ? "SELFTEST FAILED\n" : "SELFTEST OK\n");
  return fails ? 1 : 0;
}
