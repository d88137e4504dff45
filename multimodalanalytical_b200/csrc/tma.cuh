// TMA / mbarrier / tcgen05 PTX wrappers and the host-side tensor-map cache shared by the tcgen05 kernels.
#pragma once
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must trap (error returned to the host), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
// Same, for waits that are long by construction (a producer waiting for a free ring slot, the MMA issuer waiting for the
// epilogue to drain an accumulator): back off between polls so the spinning thread does not take issue slots from the
// epilogue warps that share its scheduler (ncu: BRA + SYNCS + YIELD were 7 % of the GELU-epilogue kernel's instructions).
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(64);
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- host side: cached 2-D bf16 tensor maps (128B swizzle) -----------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  unsigned long long d0, d1, ld;
  unsigned int b0, b1;
  unsigned int fmt;  // element type | swizzle mode << 8
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && ld == o.ld && b0 == o.b0 && b1 == o.b1 && fmt == o.fmt;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h = h * 1000003u ^ k.d0;
    h = h * 1000003u ^ k.d1;
    h = h * 1000003u ^ k.ld;
    h = h * 1000003u ^ ((size_t)k.b0 << 16 | k.b1);
    h = h * 1000003u ^ k.fmt;
    return h;
  }
};

// 2-D tensor map: inner extent d0 (contiguous), outer extent d1, row pitch ld elements, box b0 x b1 elements.
// is_f32: element type float instead of bf16.  swizzle: 0 none, 1 32B, 2 64B, 3 128B (CUtensorMapSwizzle values).
static int make_map_ex(CUtensorMap* out, const void* ptr, unsigned long long d0, unsigned long long d1,
                       unsigned long long ld, unsigned int b0, unsigned int b1, int is_f32, int swizzle) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const unsigned long long es_bytes = is_f32 ? 4 : 2;
  MapKey key{ptr, d0, d1, ld, b0, b1, (unsigned int)(is_f32 ? 1 : 0) | ((unsigned int)swizzle << 8)};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return MMA_OK;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return MMA_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * es_bytes) & 15)) return MMA_ERR_ARG;
  cuuint64_t dims[2] = {d0, d1};
  cuuint64_t strides[1] = {ld * es_bytes};
  cuuint32_t box[2] = {b0, b1};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(out, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  (CUtensorMapSwizzle)swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return MMA_ERR_DRIVER;
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 65536) cache.clear();
  cache.emplace(key, *out);
  return MMA_OK;
}

// 2-D bf16 tensor map with the 128-byte swizzle the tcgen05 operand tiles use.
static int make_map(CUtensorMap* out, const void* ptr, unsigned long long d0, unsigned long long d1,
                    unsigned long long ld, unsigned int b0, unsigned int b1) {
  return make_map_ex(out, ptr, d0, d1, ld, b0, b1, 0, (int)CU_TENSOR_MAP_SWIZZLE_128B);
}


}  // namespace tma
