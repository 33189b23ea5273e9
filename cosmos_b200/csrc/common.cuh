// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM
// alloc / ld / st / commit / fences) and the shared-memory + instruction descriptors they take.
// Nothing here is a library call: every wrapper is one PTX instruction (or a spin on one).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cb {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ----------------------------------------------------------------------------------------------
// misc
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(r));
  return r;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float ex2(float x) {  // MUFU.EX2, ftz; ex2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {  // generic-proxy smem writes -> visible to async proxy (UMMA/TMA)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Spin until the phase with the given parity has completed.  A bounded spin: a kernel that would
// otherwise hang (wrong descriptor, lost arrive) traps instead, so a bad build cannot wedge the GPU.
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t addr, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(addr), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_try_wait(addr, parity)) return;
  const uint64_t t0 = global_timer_ns();
#pragma unroll 1
  for (uint32_t spin = 1;; ++spin) {
    if (mbar_try_wait(addr, parity)) return;
    if ((spin & 0x3FF) == 0 && global_timer_ns() - t0 > 4000000000ull) __trap();  // 4 s: something is lost
  }
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// Start moving a box towards L2 only (no shared memory, no barrier): hides DRAM latency behind a short smem ring.
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// The same for a contiguous range of global memory (16-byte aligned, a multiple of 16 bytes): one request per range instead of
// one per box row - a tensor prefetch whose box rows are 16 bytes makes the L2 tag stage look up every 128-byte line 8 times.
__device__ __forceinline__ void bulk_prefetch_l2(const void* gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(gptr)), "r"(bytes) : "memory");
}

// Bulk asynchronous copy shared memory -> global memory (contiguous, 16-byte aligned, a multiple of 16 bytes), tracked by
// the issuing thread's bulk async-groups: commit, then wait until the group's reads of shared memory are done (the buffer
// may be rewritten) or until it has completed.
__device__ __forceinline__ void bulk_store_s2g(void* gptr, uint32_t smem_addr, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(reinterpret_cast<uint64_t>(gptr)), "r"(smem_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, fences, commit
// ----------------------------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed.
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}


// ----------------------------------------------------------------------------------------------
// CTA pairs (cluster of 2, tcgen05 cta_group::2): one MMA spans both SMs (M = 256), each CTA keeps
// its own 128 rows of A, half of B and its half of the accumulator.  CTA 0 of the pair is the leader:
// only it issues MMAs; TMA loads of both CTAs complete on the leader's mbarrier.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of both CTAs
  asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot) {  // same warp index in both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> leader CTA
// TMA load into THIS CTA's smem; the transaction bytes are credited to the LEADER CTA's mbarrier at the same offset.
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
// ---- distributed shared memory (cluster) ----
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t cta) {   // same offset in CTA `cta`'s smem
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t remote_addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(remote_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
// asynchronous 16-byte store into another CTA's shared memory; its bytes are credited (complete_tx) to an mbarrier that
// lives in the SAME destination CTA, so the sender never waits for the transfer
__device__ __forceinline__ void st_async_cluster_v4(uint32_t remote_addr, uint4 v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(remote_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(remote_bar)
               : "memory");
}
// 4 x 4 transpose of 16-byte elements inside each aligned group of four lanes: on entry lane j of a group holds
// a[c] = element (row j, column c); on return a[i] = element (row i, column j).  Two butterfly stages (lane ^ 1, lane ^ 2),
// 16 shuffles; all register indices are compile-time (the lane only picks between two registers).
__device__ __forceinline__ void quad_transpose_u4(uint4 (&a)[4], uint32_t lane) {
  const bool b0 = (lane & 1) != 0, b1 = (lane & 2) != 0;
  auto xchg = [](uint4& keep_if_set, uint4& keep_if_clear, bool set, int mask) {
    // a lane with `set` sends keep_if_clear and overwrites it with what arrives; a lane without sends / overwrites the other
    uint4 x = set ? keep_if_clear : keep_if_set, y;
    y.x = __shfl_xor_sync(0xffffffffu, x.x, mask);
    y.y = __shfl_xor_sync(0xffffffffu, x.y, mask);
    y.z = __shfl_xor_sync(0xffffffffu, x.z, mask);
    y.w = __shfl_xor_sync(0xffffffffu, x.w, mask);
    if (set) keep_if_clear = y; else keep_if_set = y;
  };
  xchg(a[1], a[0], b0, 1);
  xchg(a[3], a[2], b0, 1);
  xchg(a[2], a[0], b1, 2);
  xchg(a[3], a[1], b1, 2);
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// arrive (release at cluster scope) on an mbarrier given by its shared::cluster address
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t remote_bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t addr, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(addr), "r"(parity)
      : "memory");
  return done != 0;
}
// wait for a phase whose arrivals come from other CTAs of the cluster (acquire at cluster scope)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_try_wait_cluster(addr, parity)) return;
  const uint64_t t0 = global_timer_ns();
#pragma unroll 1
  for (uint32_t spin = 1;; ++spin) {
    if (mbar_try_wait_cluster(addr, parity)) return;
    if ((spin & 0x3FF) == 0 && global_timer_ns() - t0 > 4000000000ull) __trap();
  }
}
__device__ __forceinline__ void umma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit: arrive (once) on the mbarrier at this smem offset in every CTA of `cta_mask`
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05.mma descriptors
// ----------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, 128-byte swizzle (what TMA SWIZZLE_128B writes):
//   bits [0,14)  start address >> 4          bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4     bits [46,48) version = 1 (sm_100)
//   bits [61,64) layout type: 2 = SWIZZLE_128B
// K-major operand ([rows][64 x bf16], 128 B per row, 8-row atoms of 1024 B): SBO = 1024, LBO unused.
// MN-major operand (same bytes read with the contiguous dim as M/N): SBO = 1024 (next 8 k),
//   LBO = byte distance between consecutive 64-element chunks of the M/N dim.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// K-major operand WITHOUT swizzle: core matrices of 8 rows x 16 bytes stored as 128 contiguous bytes;
// LBO = byte distance to the core matrix of the next 8 k, SBO = to the one of the next 8 rows.
__device__ __forceinline__ uint64_t make_smem_desc_noswizzle(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
// Instruction descriptor for kind::f16, fp32 accumulate.
//   ab_fmt: 0 = fp16, 1 = bf16;  *_mn_major: 0 = K-major, 1 = MN-major;  M in {64,128}, N % 16 == 0 (M=128).
__host__ __device__ constexpr uint32_t make_idesc(uint32_t ab_fmt, uint32_t a_mn_major, uint32_t b_mn_major, uint32_t M,
                                                  uint32_t N) {
  return (1u << 4) | (ab_fmt << 7) | (ab_fmt << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]   (A: lane = row, each 32-bit column = two consecutive K elements)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// TMEM <-> registers (32 lanes x 32 bit, this thread = one lane, N consecutive columns)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// two floats -> packed 16-bit pair (lo = a, hi = b); fmt 1 = bf16, 0 = fp16
__device__ __forceinline__ uint32_t pack2(float a, float b, int fmt) {
  uint32_t r;
  if (fmt) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  } else {
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  }
  return r;
}

// warp-wide maximum in ONE instruction (sm_100a: CREDUX.MAX.F32; -inf for a warp of -inf) instead of five shuffle + max rounds
__device__ __forceinline__ float warp_max_f32(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }   // one FMNMX3

// ----------------------------------------------------------------------------------------------
// packed fp32 pairs (sm_100: FFMA2 / FADD2 / FMUL2 - two fp32 operations per lane and issue slot, each rounded exactly like the
// scalar instruction; a pair of equal scalars becomes the instruction's broadcast operand, no packing is executed)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t f2_pack_bits(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// ----------------------------------------------------------------------------------------------
// warp-level transpose-reduce: every lane holds v[0..31] (one row, 32 columns); on return lane L
// holds op over the 32 lanes of column L.  31 shuffles instead of 32 * 5.
// ----------------------------------------------------------------------------------------------
struct OpMax {
  __device__ __forceinline__ float operator()(float a, float b) const { return fmaxf(a, b); }
};
struct OpAdd {
  __device__ __forceinline__ float operator()(float a, float b) const { return a + b; }
};
template <class Op>
__device__ __forceinline__ float warp_transpose_reduce(const float (&v)[32], uint32_t lane, Op op) {
  float a[16];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float keep = up ? v[j + 16] : v[j];
      const float send = up ? v[j] : v[j + 16];
      a[j] = op(keep, __shfl_xor_sync(0xffffffffu, send, 16));
    }
  }
  float b[8];
  {
    const bool up = lane & 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float keep = up ? a[j + 8] : a[j];
      const float send = up ? a[j] : a[j + 8];
      b[j] = op(keep, __shfl_xor_sync(0xffffffffu, send, 8));
    }
  }
  float c[4];
  {
    const bool up = lane & 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float keep = up ? b[j + 4] : b[j];
      const float send = up ? b[j] : b[j + 4];
      c[j] = op(keep, __shfl_xor_sync(0xffffffffu, send, 4));
    }
  }
  float d[2];
  {
    const bool up = lane & 2;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float keep = up ? c[j + 2] : c[j];
      const float send = up ? c[j] : c[j + 2];
      d[j] = op(keep, __shfl_xor_sync(0xffffffffu, send, 2));
    }
  }
  const bool up = lane & 1;
  const float keep = up ? d[1] : d[0];
  const float send = up ? d[0] : d[1];
  return op(keep, __shfl_xor_sync(0xffffffffu, send, 1));
}

// The same reduction for sums with the additions of two columns packed into one FADD2 (the identical additions in the identical
// order: bit-equal to warp_transpose_reduce(v, lane, OpAdd()), 16 instead of 31 add instructions).
template <int kN>
__device__ __forceinline__ void transpose_add_stage(const float (&in)[2 * kN], float (&out)[kN], bool up, int sft) {
#pragma unroll
  for (int j = 0; j < kN; j += 2) {
    const float k0 = up ? in[j + kN] : in[j], k1 = up ? in[j + 1 + kN] : in[j + 1];
    const float s0 = up ? in[j] : in[j + kN], s1 = up ? in[j + 1] : in[j + 1 + kN];
    const float r0 = __shfl_xor_sync(0xffffffffu, s0, sft), r1 = __shfl_xor_sync(0xffffffffu, s1, sft);
    f2_unpack(f2_add(f2_pack(k0, k1), f2_pack(r0, r1)), out[j], out[j + 1]);
  }
}
__device__ __forceinline__ float warp_transpose_sum(const float (&v)[32], uint32_t lane) {
  float a[16], b[8], c[4], d[2];
  transpose_add_stage<16>(v, a, lane & 16, 16);
  transpose_add_stage<8>(a, b, lane & 8, 8);
  transpose_add_stage<4>(b, c, lane & 4, 4);
  transpose_add_stage<2>(c, d, lane & 2, 2);
  const bool up = lane & 1;
  const float keep = up ? d[1] : d[0];
  const float send = up ? d[0] : d[1];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

}  // namespace cb
