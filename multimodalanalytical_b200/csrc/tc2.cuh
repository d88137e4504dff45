// Constants and PTX wrappers shared by the CTA-pair (cta_group::2) tcgen05 kernels: gemm_tc2.cu, gemm_glu2.cu.
#pragma once
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tma.cuh"

// long waits of the producer / MMA-issuer threads: -DMMA_TC2_BACKOFF=1 backs off (nanosleep) between polls.  Measured on
// B200, same box A/B: C2 6.94 vs 6.89 ms/step, paper variant 8.20 vs 8.21, C4 13.97 vs 13.99 - nothing; the plain spin stays.
#ifndef MMA_TC2_BACKOFF
#define MMA_TC2_BACKOFF 0
#endif
#if MMA_TC2_BACKOFF
#define MBAR_WAIT_LONG(bar, parity) mbar_wait_backoff(bar, parity)
#else
#define MBAR_WAIT_LONG(bar, parity) mbar_wait(bar, parity)
#endif

namespace tc2 {
using namespace tma;

constexpr int BM = 128;      // rows per CTA; the pair tile is 2 * BM x BN
constexpr int BN = 256;      // pair-tile columns (each CTA loads BN / 2 rows of B)
constexpr int BK = 64;       // 64 bf16 = one 128-byte swizzle row
constexpr int STAGES = 4;
constexpr int NUM_EPI_WARPS = 16;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr uint32_t A_BYTES = BM * BK * 2;
constexpr uint32_t B_BYTES = (BN / 2) * BK * 2;
constexpr uint32_t BOX_BYTES = 32 * 64;                 // one epilogue box: 32 rows x 64 bytes
constexpr uint32_t EPI_WARP_BYTES = 3 * BOX_BYTES;
constexpr uint32_t BAR_BYTES = 512;
constexpr uint32_t SMEM = 1024 /*align slack*/ + STAGES * (A_BYTES + B_BYTES) + NUM_EPI_WARPS * EPI_WARP_BYTES + BAR_BYTES;
static_assert(SMEM <= 232448, "shared memory budget");

// ---- PTX wrappers that only this kernel needs -----------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Arrive on a peer CTA's mbarrier.  CTA-scope release (the PTX default, what CUTLASS' ClusterBarrier::arrive issues):
// the only work this hand-off orders is the warp's TMEM reads, which tcgen05.wait::ld + tcgen05.fence::before_thread_sync
// have already retired.  A cluster-scope release made ptxas emit MEMBAR.ALL.GPU + CGAERRBAR in front of every arrive
// (10 % of the epilogue warps' stall samples in the GELU kernel).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of one CTA of a pair: data lands in this CTA's shared memory, the bytes are counted on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_mma2_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs of the pair arrives on the mbarrier at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void tc_commit2_mc(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, 128B swizzle (same layouts as gemm_tc.cu)
template <bool MN_MAJOR>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t tile_addr, int k16) {
  uint32_t addr, lbo, sbo;
  if (!MN_MAJOR) {
    addr = tile_addr + (uint32_t)k16 * 32u;
    lbo = 16;
    sbo = 1024;
  } else {
    addr = tile_addr + (uint32_t)k16 * 2048u;
    lbo = BK * 128;
    sbo = 1024;
  }
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// dropout key / threshold of a launch, hoisted out of the per-chunk epilogue math
struct DropCtx {
  bool on;
  uint32_t thr, key;
  float inv_keep;
};

// 16-byte piece `k` (0..3) of row `lane` inside a [32][64 B] box; swz = 1: CU_TENSOR_MAP_SWIZZLE_64B
__device__ __forceinline__ uint32_t box_off(int lane, int k, int swz) {
  return (uint32_t)(lane * 64 + ((k ^ (swz ? ((lane >> 1) & 3) : 0)) << 4));
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

}  // namespace tc2
