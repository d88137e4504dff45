// Data-parallel optimiser step over NVLink peer memory (one process per GPU, buffers in symmetric allocations).
//
// The reference trains under Lightning DDP (trainer/trainer.py:58-71): a bucketed NCCL all-reduce of the 44 M fp32
// gradients, then every rank runs the same Adam step over all parameters.  Here the collective and the optimiser are one
// pass over peer memory (reduce-scatter + rank-sharded Adam + all-gather of the bf16 weights, with no intermediate copy
// and no NCCL kernel competing for SMs with the persistent tcgen05 CTAs of backward):
//
//   barrier                      every rank's backward is done, its gradient buffer is final
//   reduce_shard   rank r sums ITS shard [lo_r, hi_r) of the gradient over all peers (P2P loads over NVLink), keeps the sum
//                  in its own gradient buffer and the shard's sum of squares in its symmetric `sumsq` slot
//   barrier                      all partial norms are visible; nobody reads remote gradients any more
//   adam_shard     global norm = sqrt(sum of the peers' partials, fixed order: identical on every rank) -> clip ->
//                  Adam / AdamW on the shard's fp32 master weights and moments (1/world of the optimiser work per GPU)
//                  -> the updated weights go out as bf16 straight into EVERY peer's GEMM-operand mirror (P2P stores);
//                  the rank's whole gradient buffer is zeroed for the next step
//   barrier                      every mirror is complete before the next forward reads it
//
// fp32 master weights and Adam moments are therefore sharded (ZeRO-1 style); `ParamStore.gather_master()` reassembles the
// master for checkpoints.  Barriers are single-block kernels over flag words in symmetric memory (release / acquire at
// system scope, epoch counter on the device so a captured CUDA graph replays them), bounded spins that trap instead of
// hanging the GPU.
#include "common.cuh"

namespace p2p {

constexpr int MAX_WORLD = 16;
struct Peers {
  void* p[MAX_WORLD];
};

__device__ __forceinline__ void st_release_sys(int* addr, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* addr) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_sys_f4(const float* addr) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float ld_sys_f(const float* addr) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(addr) : "memory");
  return v;
}

// NVLS (NVSwitch multicast objects): one load returns the SUM over all ranks' copies of the address, reduced inside the
// switch; one store is replicated by the switch into every rank's copy
__device__ __forceinline__ float4 multimem_ld_reduce_add_f4(const float* mc_addr) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc_addr) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_b128(void* mc_addr, uint4 u) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               ::"l"(mc_addr), "f"(__uint_as_float(u.x)), "f"(__uint_as_float(u.y)), "f"(__uint_as_float(u.z)),
                 "f"(__uint_as_float(u.w)) : "memory");
}

// flags: per rank an int[MAX_WORLD] array in symmetric memory; flags_r[s] = last epoch rank s has reached
__global__ void barrier_kernel(Peers flags, int* epoch_ctr, int world, int rank) {
  __shared__ int e;
  if (threadIdx.x == 0) {
    e = *epoch_ctr + 1;
    *epoch_ctr = e;
  }
  __syncthreads();
  const int p = threadIdx.x;
  if (p < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<int*>(flags.p[p]) + rank, e);
    const int* mine = reinterpret_cast<const int*>(flags.p[rank]) + p;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) - e < 0) {
      if (clock64() - t0 > 20000000000LL) __trap();  // ~10 s: a peer is gone - fail instead of hanging
    }
  }
}

// g_rank[lo:hi) = sum_p g_p[lo:hi);  partial[block] = sum of squares of the block's part
__global__ void __launch_bounds__(256) reduce_shard_kernel(Peers g, const float* mc_g, int world, int rank, long long lo,
                                                           long long hi, float* __restrict__ partial) {
  __shared__ float red[8];
  float ss = 0.f;
  float* mine = reinterpret_cast<float*>(g.p[rank]);
  const long long n4 = (hi - lo) >> 2;  // shard bounds are multiples of 128 elements
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long e = lo + (i << 2);
    float4 s;
    if (mc_g) {
      s = multimem_ld_reduce_add_f4(mc_g + e);  // in-switch sum over every rank's copy (this rank's included)
    } else {
      s = *reinterpret_cast<const float4*>(mine + e);
      for (int p = 0; p < world; ++p) {
        if (p == rank) continue;
        const float4 v = ld_sys_f4(reinterpret_cast<const float*>(g.p[p]) + e);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
    *reinterpret_cast<float4*>(mine + e) = s;
    ss += s.x * s.x + s.y * s.y + s.z * s.z + s.w * s.w;
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    partial[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) sumsq_final_kernel(const float* __restrict__ partial, int nblocks, float* out) {
  __shared__ float red[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < nblocks; i += 256) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

// hyper: as adam_kernel (trainops.cu); hyper[8] = 1 / (world * accumulation) scales the SUMMED gradient
__global__ void __launch_bounds__(256) adam_shard_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, Peers pb, void* mc_pb, Peers sumsq, int world, int rank,
                                                         long long n, long long lo, long long hi,
                                                         const float* __restrict__ hyper, int decoupled) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const float bc1 = hyper[5], bc2 = hyper[6], max_norm = hyper[7], gs = hyper[8];
  float clip = gs;
  if (max_norm > 0.f) {
    float tot = 0.f;
    for (int q = 0; q < world; ++q) tot += q == rank ? reinterpret_cast<const float*>(sumsq.p[q])[0]
                                                     : ld_sys_f(reinterpret_cast<const float*>(sumsq.p[q]));
    const float c = max_norm / (sqrtf(tot) * gs + 1e-6f);
    if (c < 1.f) clip *= c;
  }
  const float step = lr / bc1;
  const float rs2 = rsqrtf(bc2);
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  const long long n8 = (hi - lo) >> 3;
  for (long long i = tid; i < n8; i += nth) {
    const long long e = lo + (i << 3);
    float pv[8], gv[8], mv[8], vv[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 a = *reinterpret_cast<const float4*>(p + e + 4 * h), b = *reinterpret_cast<const float4*>(g + e + 4 * h);
      const float4 c = *reinterpret_cast<const float4*>(m + e + 4 * h), d = *reinterpret_cast<const float4*>(v + e + 4 * h);
      pv[4 * h] = a.x; pv[4 * h + 1] = a.y; pv[4 * h + 2] = a.z; pv[4 * h + 3] = a.w;
      gv[4 * h] = b.x; gv[4 * h + 1] = b.y; gv[4 * h + 2] = b.z; gv[4 * h + 3] = b.w;
      mv[4 * h] = c.x; mv[4 * h + 1] = c.y; mv[4 * h + 2] = c.z; mv[4 * h + 3] = c.w;
      vv[4 * h] = d.x; vv[4 * h + 1] = d.y; vv[4 * h + 2] = d.z; vv[4 * h + 3] = d.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float pi = pv[j];
      float gi = gv[j] * clip;
      if (!decoupled && wd != 0.f) gi += wd * pi;
      const float mi = b1 * mv[j] + (1.f - b1) * gi;
      const float vi = b2 * vv[j] + (1.f - b2) * gi * gi;
      if (decoupled && wd != 0.f) pi *= 1.f - lr * wd;
      pi -= step * mi / (sqrtf(vi) * rs2 + eps);
      pv[j] = pi; mv[j] = mi; vv[j] = vi;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      *reinterpret_cast<float4*>(p + e + 4 * h) = make_float4(pv[4 * h], pv[4 * h + 1], pv[4 * h + 2], pv[4 * h + 3]);
      *reinterpret_cast<float4*>(m + e + 4 * h) = make_float4(mv[4 * h], mv[4 * h + 1], mv[4 * h + 2], mv[4 * h + 3]);
      *reinterpret_cast<float4*>(v + e + 4 * h) = make_float4(vv[4 * h], vv[4 * h + 1], vv[4 * h + 2], vv[4 * h + 3]);
    }
    uint4 u;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(pv[0], pv[1]), h1 = __floats2bfloat162_rn(pv[2], pv[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(pv[4], pv[5]), h3 = __floats2bfloat162_rn(pv[6], pv[7]);
    u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
    if (mc_pb) {
      multimem_st_b128(reinterpret_cast<bf16*>(mc_pb) + e, u);  // replicated into every rank's mirror by the switch
    } else {
      for (int q = 0; q < world; ++q) *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(pb.p[q]) + e) = u;
    }
    // consumed: the next step accumulates into a clean gradient buffer
    *reinterpret_cast<float4*>(g + e) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(g + e + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // the rest of this rank's gradient buffer (the peers' shards; their remote reads ended at the barrier before this
  // kernel).  Only OUTSIDE [lo, hi): other threads of this grid may still be reading shard elements.
  const long long n4 = n >> 2, lo4 = lo >> 2, hi4 = hi >> 2;
  for (long long i = tid; i < n4; i += nth)
    if (i < lo4 || i >= hi4) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

static int fill(Peers& P, const void* const* ptrs, int world) {
  if (!ptrs || world < 1 || world > MAX_WORLD) return MMA_ERR_ARG;
  for (int i = 0; i < MAX_WORLD; ++i) P.p[i] = i < world ? const_cast<void*>(ptrs[i]) : nullptr;
  for (int i = 0; i < world; ++i)
    if (!P.p[i]) return MMA_ERR_ARG;
  return MMA_OK;
}

}  // namespace p2p

using namespace p2p;

// Cross-GPU barrier of `world` ranks.  peer_flags: HOST array of the `world` device pointers of the ranks' flag arrays
// (int[16] each, symmetric memory, zero-initialised); epoch_ctr: this rank's device counter (zero-initialised).
extern "C" int mma_p2p_barrier(const void* const* peer_flags, int* epoch_ctr, int world, int rank, cudaStream_t stream) {
  Peers F;
  if (int rc = fill(F, peer_flags, world)) return rc;
  if (!epoch_ctr || rank < 0 || rank >= world) return MMA_ERR_ARG;
  barrier_kernel<<<1, 32, 0, stream>>>(F, epoch_ctr, world, rank);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

// g_rank[lo:hi) = sum over the peers' g[lo:hi); sumsq_out[0] = sum of squares of the result.  peer_g: HOST array of the
// `world` device pointers of the ranks' flat fp32 gradient buffers; mc_g: NVLS multicast address of the same buffer (the
// sum is then formed inside the NVSwitch by multimem.ld_reduce) or NULL (explicit P2P loads); lo, hi multiples of 4;
// workspace >= 1024 floats.
extern "C" int mma_p2p_reduce_shard(const void* const* peer_g, const void* mc_g, int world, int rank, long long lo,
                                    long long hi, float* workspace, float* sumsq_out, cudaStream_t stream) {
  Peers G;
  if (int rc = fill(G, peer_g, world)) return rc;
  if (rank < 0 || rank >= world || lo < 0 || hi < lo || ((lo | hi) & 3) || !workspace || !sumsq_out) return MMA_ERR_ARG;
  long long want = ((hi - lo) / 4 + 255) / 256;
  const int blocks = (int)(want < 1 ? 1 : (want > 1024 ? 1024 : want));
  reduce_shard_kernel<<<blocks, 256, 0, stream>>>(G, reinterpret_cast<const float*>(mc_g), world, rank, lo, hi, workspace);
  MMA_CHECK_LAUNCH();
  p2p::sumsq_final_kernel<<<1, 256, 0, stream>>>(workspace, blocks, sumsq_out);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

// Adam / AdamW on the shard [lo, hi) of this rank's master weights with the clipped, world-averaged gradient; the new
// weights are written as bf16 into every peer's mirror (peer_pb: HOST array of `world` device pointers; mc_pb: NVLS
// multicast address of the mirror - one multimem.st replicated by the switch - or NULL: one P2P store per peer); this rank's
// whole gradient buffer [0, n) is zeroed.  peer_sumsq: HOST array of the ranks' sumsq slots (mma_p2p_reduce_shard).
// lo, hi multiples of 8, n multiple of 4.  hyper as mma_adam_step.
extern "C" int mma_p2p_adam_shard(float* p, float* g, float* m, float* v, const void* const* peer_pb, void* mc_pb,
                                  const void* const* peer_sumsq, int world, int rank, long long n, long long lo,
                                  long long hi, const float* hyper, int decoupled, cudaStream_t stream) {
  Peers PB, SS;
  if (int rc = fill(PB, peer_pb, world)) return rc;
  if (int rc = fill(SS, peer_sumsq, world)) return rc;
  if (!p || !g || !m || !v || !hyper || rank < 0 || rank >= world || lo < 0 || hi < lo || hi > n || ((lo | hi) & 7) || (n & 3))
    return MMA_ERR_ARG;
  long long want = (n / 4 + 255) / 256;
  const int blocks = (int)(want > 148 * 8 ? 148 * 8 : (want < 1 ? 1 : want));
  adam_shard_kernel<<<blocks, 256, 0, stream>>>(p, g, m, v, PB, mc_pb, SS, world, rank, n, lo, hi, hyper, decoupled);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}
