// tcgen05 / TMEM attention for short sequences (bf16, head dim 64): the shapes of this path (S ~ 27-128 encoder
// positions, T <= 128 target positions) fit ONE 128-row score tile, so forward and backward are single-shot
// kernels with every matrix product on the 5th-generation tensor cores:
//
//   forward   S = Q K^T -> row softmax in registers (one thread per row) -> P (bf16, smem) -> O = P V
//   backward  S = Q K^T, dP = dO V^T -> P, dS (bf16, smem) -> dV = P^T dO, dK = dS^T Q, dQ = dS K
//
// Operands are TMA-loaded 128B-swizzled tiles used directly as K-major or MN-major UMMA operands (the P / dS
// tiles serve as K-major A for "x K / x V" and as MN-major A for the transposed products - no transposes).
// PACK = 2 packs two (batch, head) problems of <= 64 queries and <= 64 keys into one 128x128 tile (rows 0-63 /
// keys 0-63 = problem A, rows 64-127 / keys 64-127 = problem B; the off-diagonal blocks of P / dS are written as
// zeros so the packed products stay block-diagonal).  PACK = 1 is one problem with <= 128 queries and keys.
// Longer sequences use the streaming mma.sync kernels (attention_mma.cu).  Masks / dropout stream: as everywhere.
#include <cstdlib>

#include "common.cuh"
#include "tma.cuh"

namespace at5 {
using namespace tma;

constexpr float LOG2E = 1.4426950408889634f;
constexpr uint32_t TILE = 128 * 128;  // bytes of a [128 rows x 64 bf16] tile

struct Args {
  int B, H, Lq, Lk, causal, nprob;
  float scale, p_drop;
  unsigned long long seed;
  unsigned int site;
  const unsigned char* kmask;
  bf16* o; long long ldo;
  float* lse;
  // backward
  const bf16* o_in; const bf16* dout; long long lddo;
  bf16* dq; bf16* dk; bf16* dv;
  long long lddq, lddk, lddv;
};

__device__ __forceinline__ void tmem_ld32f(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptors (128B swizzle), see gemm_tc.cu
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) {  // rows of 128 B, 8-row groups 1024 B apart
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t addr, uint32_t atom_stride) {  // k rows of 128 B
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((atom_stride >> 4) & 0x3FFF) << 16) |
         ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ constexpr uint32_t idesc(int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// write 8 bf16 (one 16-byte chunk) of row r, logical column c (multiple of 8) of a [128 x 64] swizzled tile
__device__ __forceinline__ void st_chunk(uint8_t* tile, int r, int c, uint4 v) {
  *reinterpret_cast<uint4*>(tile + r * 128 + ((((c >> 3) ^ (r & 7))) << 4)) = v;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}

// which (batch, head, local row) a tile row belongs to
template <int PACK>
struct RowMap {
  int prob, b, h, i;   // problem index, batch, head, local row index
  bool exists;
  int key_col0;        // first score column of this row's problem
  __device__ RowMap(const Args& a, int row, int tile = blockIdx.x) {
    const int sub = PACK == 2 ? (row >> 6) : 0;
    prob = tile * PACK + sub;
    exists = prob < a.nprob;
    const int pp = exists ? prob : 0;
    b = pp / a.H;
    h = pp % a.H;
    i = PACK == 2 ? (row & 63) : row;
    key_col0 = sub * 64;
  }
};

template <int PACK>
__device__ __forceinline__ void issue_tile_loads(const Args& a, const CUtensorMap* tm, uint32_t dst, uint32_t bar, int L,
                                                 int tile = blockIdx.x) {
  // two 64-row boxes: rows [0,64) and [64,128) of the tile
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    int prob = PACK == 2 ? tile * 2 + s : tile;
    if (prob >= a.nprob) prob = a.nprob - 1;  // dummy (masked) when the pair is incomplete
    const int b = prob / a.H, h = prob % a.H;
    const int row0 = b * L + (PACK == 2 ? 0 : s * 64);
    tma_load_2d(dst + s * 8192, tm, bar, h * 64, row0);
  }
}


// Validity bits of the keys [ch*32, ch*32+32) of this row's problem: key < Lk, key-padding mask (warp ballot: all rows
// of a warp belong to the same problem), causal j <= i.
template <int NCH>
__device__ __forceinline__ void key_valid_bits(uint32_t (&bits)[NCH], const unsigned char* km, int Lk, int causal, int i,
                                               int lane) {
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int j = ch * 32 + lane;
    const bool ok = j < Lk && (!km || km[j] != 0);
    uint32_t m = __ballot_sync(0xffffffffu, ok);
    if (causal) {
      const int hi = i - ch * 32;  // keys 0..hi of this chunk are visible
      m &= hi >= 31 ? 0xffffffffu : hi < 0 ? 0u : ((2u << hi) - 1u);
    }
    bits[ch] = m;
  }
}

// ------------------------------------------------------------------------------------------------ forward
template <int PACK>
__global__ void __launch_bounds__(128) fwd_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                                                  const __grid_constant__ CUtensorMap tv, Args a) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];  // no static smem in this kernel: the window starts 1024-aligned
  uint8_t* sQ = smem;               // after S is formed, sQ|sK (32 KB) is reused for P (two 64-key atoms)
  uint8_t* sK = smem + TILE;
  uint8_t* sV = smem + 2 * TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * TILE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  pdl_wait();
  if (threadIdx.x == 0) {
    const uint32_t bar = smem_u32(&bars[0]);
    mbar_expect_tx(bar, 3 * TILE);
    issue_tile_loads<PACK>(a, &tq, smem_u32(sQ), bar, a.Lq);
    issue_tile_loads<PACK>(a, &tk, smem_u32(sK), bar, a.Lk);
    issue_tile_loads<PACK>(a, &tv, smem_u32(sV), bar, a.Lk);
    mbar_wait(bar, 0);
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < 4; ++k)  // S[128 x 128] = Q K^T
      tc_mma_bf16(tm, desc_kmajor(smem_u32(sQ) + k * 32), desc_kmajor(smem_u32(sK) + k * 32), idesc(128, false, false), k > 0);
    tc_commit(smem_u32(&bars[1]));
  }
  const int row = warp * 32 + lane;
  const RowMap<PACK> rm(a, row);
  const unsigned char* km = a.kmask ? a.kmask + (long long)rm.b * a.Lk : nullptr;
  const float sl2 = a.scale * LOG2E;
  const bool drop = a.p_drop > 0.f;
  const uint32_t thr = drop ? drop_threshold(a.p_drop) : 0u;
  const float inv_keep = drop ? 1.f / (1.f - a.p_drop) : 1.f;
  const uint32_t dkey = drop_key(a.seed, a.site);
  const unsigned long long ebase = (((unsigned long long)rm.b * a.H + rm.h) * a.Lq + rm.i) * (unsigned long long)((a.Lk + 1) & ~1);
  constexpr int NCH = PACK == 2 ? 2 : 4;  // 32-column chunks of this row's problem
  const uint32_t t_row = tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)rm.key_col0;

  uint32_t vbits[NCH];
  key_valid_bits<NCH>(vbits, km, a.Lk, a.causal, rm.i, lane);
  const uint32_t ebase32 = (uint32_t)ebase;  // even: the dropout row pitch is padded to a multiple of 2

  if (lane == 0) mbar_wait(smem_u32(&bars[1]), 0);  // one poller per warp
  __syncwarp();
  tc_fence_after();
  float mx = -INFINITY;
#pragma unroll 1
  for (int ch = 0; ch < NCH; ++ch) {
    float s[32];
    tmem_ld32f(t_row + ch * 32, s);
    const uint32_t vb = vbits[ch];
#pragma unroll
    for (int c = 0; c < 32; ++c) mx = fmaxf(mx, (vb >> c) & 1u ? s[c] * sl2 : -INFINITY);
  }
  // every thread has read what it needs for the max; P overwrites sQ|sK only after the S product has retired (bars[1])
  float l = 0.f;
  uint8_t* sP = sQ;
#pragma unroll 1
  for (int ch = 0; ch < NCH; ++ch) {
    float s[32];
    tmem_ld32f(t_row + ch * 32, s);
    const uint32_t vb = vbits[ch];
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
      float p0 = (vb >> c) & 1u ? ex2_approx(s[c] * sl2 - mx) : 0.f;
      float p1 = (vb >> (c + 1)) & 1u ? ex2_approx(s[c + 1] * sl2 - mx) : 0.f;
      l += p0 + p1;
      if (drop) {
        const uint32_t r = drop_pair(dkey, (ebase32 + (uint32_t)(ch * 32 + c)) >> 1);
        p0 *= (r & 0xFFFFu) >= thr ? inv_keep : 0.f;
        p1 *= (r >> 16) >= thr ? inv_keep : 0.f;
      }
      s[c] = p0;
      s[c + 1] = p1;
    }
    const int kcol = rm.key_col0 + ch * 32;  // column in the packed 128-wide K dimension
    uint8_t* atom = sP + (kcol >> 6) * TILE;
#pragma unroll
    for (int c = 0; c < 32; c += 8) {
      uint4 u;
      u.x = pack2(s[c], s[c + 1]); u.y = pack2(s[c + 2], s[c + 3]);
      u.z = pack2(s[c + 4], s[c + 5]); u.w = pack2(s[c + 6], s[c + 7]);
      st_chunk(atom, row, (kcol & 63) + c, u);
    }
  }
  if (PACK == 2) {  // zero the other problem's key block of this row
    uint8_t* atom = sP + ((rm.key_col0 >> 6) ^ 1) * TILE;
#pragma unroll
    for (int c = 0; c < 64; c += 8) st_chunk(atom, row, c, make_uint4(0, 0, 0, 0));
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < 8; ++k)  // O[128 x 64] = P[128 x 128] V[128 x 64]
      tc_mma_bf16(tm, desc_kmajor(smem_u32(sP) + (k >> 2) * TILE + (k & 3) * 32),
                  desc_mnmajor(smem_u32(sV) + k * 2048, TILE), idesc(64, false, true), k > 0);
    tc_commit(smem_u32(&bars[2]));
  }
  if (lane == 0) mbar_wait(smem_u32(&bars[2]), 0);  // one poller per warp
  __syncwarp();
  tc_fence_after();
  {
    const float inv = l > 0.f ? 1.f / l : 0.f;
    const bool wr = rm.exists && rm.i < a.Lq;
    bf16* orow = a.o + ((long long)rm.b * a.Lq + rm.i) * a.ldo + rm.h * 64;
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      float v[32];
      tmem_ld32f(tm + ((uint32_t)(warp * 32) << 16) + ch * 32, v);
      if (wr) {
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          uint4 u;
          u.x = pack2(v[c] * inv, v[c + 1] * inv); u.y = pack2(v[c + 2] * inv, v[c + 3] * inv);
          u.z = pack2(v[c + 4] * inv, v[c + 5] * inv); u.w = pack2(v[c + 6] * inv, v[c + 7] * inv);
          *reinterpret_cast<uint4*>(orow + ch * 32 + c) = u;
        }
      }
    }
    if (wr && a.lse) a.lse[((long long)rm.b * a.H + rm.h) * a.Lq + rm.i] = l > 0.f ? mx / LOG2E + logf(l) : -INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ backward
template <int PACK>
__global__ void __launch_bounds__(128) bwd_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                                                  const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo,
                                                  const __grid_constant__ CUtensorMap to, Args a) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];  // no static smem in this kernel: the window starts 1024-aligned
  // layout (ascending): Pd atom 0 | dS atom 0 | dS atom 1 | Q | K | dO | V (= Pd atom 1 once dP has retired)
  uint8_t* sPd0 = smem;
  uint8_t* sdS = smem + TILE;
  uint8_t* sQ = smem + 3 * TILE;
  uint8_t* sK = smem + 4 * TILE;
  uint8_t* sDO = smem + 5 * TILE;
  uint8_t* sV = smem + 6 * TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * TILE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  pdl_wait();
  if (threadIdx.x == 0) {
    const uint32_t bar = smem_u32(&bars[0]);
    mbar_expect_tx(bar, 5 * TILE);
    issue_tile_loads<PACK>(a, &tq, smem_u32(sQ), bar, a.Lq);
    issue_tile_loads<PACK>(a, &tk, smem_u32(sK), bar, a.Lk);
    issue_tile_loads<PACK>(a, &tv, smem_u32(sV), bar, a.Lk);
    issue_tile_loads<PACK>(a, &tdo, smem_u32(sDO), bar, a.Lq);
    issue_tile_loads<PACK>(a, &to, smem_u32(sdS), bar, a.Lq);  // O parks in dS atom 0 until D_i is formed
    mbar_wait(bar, 0);
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < 4; ++k)  // S = Q K^T -> cols [0,128)
      tc_mma_bf16(tm, desc_kmajor(smem_u32(sQ) + k * 32), desc_kmajor(smem_u32(sK) + k * 32), idesc(128, false, false), k > 0);
#pragma unroll
    for (int k = 0; k < 4; ++k)  // dP = dO V^T -> cols [128,256)
      tc_mma_bf16(tm + 128, desc_kmajor(smem_u32(sDO) + k * 32), desc_kmajor(smem_u32(sV) + k * 32), idesc(128, false, false), k > 0);
    tc_commit(smem_u32(&bars[1]));
  }
  const int row = warp * 32 + lane;
  const RowMap<PACK> rm(a, row);
  const bool qvalid = rm.exists && rm.i < a.Lq;
  const unsigned char* km = a.kmask ? a.kmask + (long long)rm.b * a.Lk : nullptr;
  // D_i = sum_c dO[i,c] O[i,c] and lse_i while the tensor core works.  Both rows come from the TMA-loaded shared
  // tiles (a thread owns one row: read from global that is 16 scattered 16-byte loads per thread)
  float Di = 0.f, lse2 = 0.f;
  if (qvalid) lse2 = a.lse[((long long)rm.b * a.H + rm.h) * a.Lq + rm.i] * LOG2E;
  if (lane == 0) mbar_wait(smem_u32(&bars[0]), 0);
  __syncwarp();
  {
    const uint8_t* orow = sdS + row * 128;
    const uint8_t* drow = sDO + row * 128;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int sw = (u ^ (row & 7)) << 4;
      const uint4 x = *reinterpret_cast<const uint4*>(orow + sw), y = *reinterpret_cast<const uint4*>(drow + sw);
      const __nv_bfloat162* xb = reinterpret_cast<const __nv_bfloat162*>(&x);
      const __nv_bfloat162* yb = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float2 xf = __bfloat1622float2(xb[w]), yf = __bfloat1622float2(yb[w]);
        Di += xf.x * yf.x + xf.y * yf.y;
      }
    }
    if (!qvalid) Di = 0.f;
  }
  const float sl2 = a.scale * LOG2E;
  const bool drop = a.p_drop > 0.f;
  const uint32_t thr = drop ? drop_threshold(a.p_drop) : 0u;
  const float inv_keep = drop ? 1.f / (1.f - a.p_drop) : 1.f;
  const uint32_t dkey = drop_key(a.seed, a.site);
  const unsigned long long ebase = (((unsigned long long)rm.b * a.H + rm.h) * a.Lq + rm.i) * (unsigned long long)((a.Lk + 1) & ~1);
  constexpr int NCH = PACK == 2 ? 2 : 4;
  const uint32_t t_row = tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)rm.key_col0;

  uint32_t vbits[NCH];
  key_valid_bits<NCH>(vbits, km, a.Lk, a.causal, rm.i, lane);
  const uint32_t ebase32 = (uint32_t)ebase;

  if (lane == 0) mbar_wait(smem_u32(&bars[1]), 0);  // one poller per warp
  __syncwarp();
  tc_fence_after();
#pragma unroll 1
  for (int ch = 0; ch < NCH; ++ch) {
    float s[32], dp[32];
    tmem_ld32f(t_row + ch * 32, s);
    tmem_ld32f(t_row + 128 + ch * 32, dp);
    const uint32_t vb = qvalid ? vbits[ch] : 0u;
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
      const float p0 = (vb >> c) & 1u ? ex2_approx(s[c] * sl2 - lse2) : 0.f;
      const float p1 = (vb >> (c + 1)) & 1u ? ex2_approx(s[c + 1] * sl2 - lse2) : 0.f;
      float k0 = 1.f, k1 = 1.f;
      if (drop) {
        const uint32_t r = drop_pair(dkey, (ebase32 + (uint32_t)(ch * 32 + c)) >> 1);
        k0 = (r & 0xFFFFu) >= thr ? inv_keep : 0.f;
        k1 = (r >> 16) >= thr ? inv_keep : 0.f;
      }
      s[c] = p0 * k0;                        // P_drop
      s[c + 1] = p1 * k1;
      dp[c] = p0 * (dp[c] * k0 - Di);        // dS
      dp[c + 1] = p1 * (dp[c + 1] * k1 - Di);
    }
    const int kcol = rm.key_col0 + ch * 32;
    uint8_t* pa = (kcol >> 6) ? sV : sPd0;
    uint8_t* da = sdS + (kcol >> 6) * TILE;
#pragma unroll
    for (int c = 0; c < 32; c += 8) {
      uint4 u, w;
      u.x = pack2(s[c], s[c + 1]); u.y = pack2(s[c + 2], s[c + 3]);
      u.z = pack2(s[c + 4], s[c + 5]); u.w = pack2(s[c + 6], s[c + 7]);
      w.x = pack2(dp[c], dp[c + 1]); w.y = pack2(dp[c + 2], dp[c + 3]);
      w.z = pack2(dp[c + 4], dp[c + 5]); w.w = pack2(dp[c + 6], dp[c + 7]);
      st_chunk(pa, row, (kcol & 63) + c, u);
      st_chunk(da, row, (kcol & 63) + c, w);
    }
  }
  if (PACK == 2) {
    const int other = (rm.key_col0 >> 6) ^ 1;
    uint8_t* pa = other ? sV : sPd0;
    uint8_t* da = sdS + other * TILE;
#pragma unroll
    for (int c = 0; c < 64; c += 8) {
      st_chunk(pa, row, c, make_uint4(0, 0, 0, 0));
      st_chunk(da, row, c, make_uint4(0, 0, 0, 0));
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    const uint32_t pd_stride = smem_u32(sV) - smem_u32(sPd0);  // distance between the two key atoms of Pd
#pragma unroll
    for (int k = 0; k < 8; ++k)  // dV[keys x 64] = Pd^T dO     -> cols [0,64)
      tc_mma_bf16(tm, desc_mnmajor(smem_u32(sPd0) + k * 2048, pd_stride), desc_mnmajor(smem_u32(sDO) + k * 2048, TILE),
                  idesc(64, true, true), k > 0);
#pragma unroll
    for (int k = 0; k < 8; ++k)  // dK[keys x 64] = dS^T Q      -> cols [64,128)
      tc_mma_bf16(tm + 64, desc_mnmajor(smem_u32(sdS) + k * 2048, TILE), desc_mnmajor(smem_u32(sQ) + k * 2048, TILE),
                  idesc(64, true, true), k > 0);
#pragma unroll
    for (int k = 0; k < 8; ++k)  // dQ[q x 64] = dS K           -> cols [128,192)
      tc_mma_bf16(tm + 128, desc_kmajor(smem_u32(sdS) + (k >> 2) * TILE + (k & 3) * 32),
                  desc_mnmajor(smem_u32(sK) + k * 2048, TILE), idesc(64, false, true), k > 0);
    tc_commit(smem_u32(&bars[2]));
  }
  if (lane == 0) mbar_wait(smem_u32(&bars[2]), 0);  // one poller per warp
  __syncwarp();
  tc_fence_after();
  {
    // row r is query r of its problem for dQ and key r of its problem for dK / dV
    const bool kvalid = rm.exists && rm.i < a.Lk;
    bf16* dqrow = a.dq + ((long long)rm.b * a.Lq + rm.i) * a.lddq + rm.h * 64;
    bf16* dkrow = a.dk + ((long long)rm.b * a.Lk + rm.i) * a.lddk + rm.h * 64;
    bf16* dvrow = a.dv + ((long long)rm.b * a.Lk + rm.i) * a.lddv + rm.h * 64;
    const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int part = 0; part < 6; ++part) {  // dV lo/hi, dK lo/hi, dQ lo/hi
      float v[32];
      tmem_ld32f(tl + part * 32, v);
      const int which = part >> 1, half = part & 1;
      const bool wr = which == 2 ? qvalid : kvalid;
      const float sc = which == 0 ? 1.f : a.scale;
      bf16* dst = (which == 0 ? dvrow : which == 1 ? dkrow : dqrow) + half * 32;
      if (wr) {
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          uint4 u;
          u.x = pack2(v[c] * sc, v[c + 1] * sc); u.y = pack2(v[c + 2] * sc, v[c + 3] * sc);
          u.z = pack2(v[c + 4] * sc, v[c + 5] * sc); u.w = pack2(v[c + 6] * sc, v[c + 7] * sc);
          *reinterpret_cast<uint4*>(dst + c) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ forward, pipelined
// Persistent, warp-specialised version of fwd_kernel (same structure as bwd_pipe_kernel below): TMA warp with a 2-stage
// (Q K V) ring, one MMA thread, an 8-warp softmax group (warps w and w+4 share a row quarter and split its key
// columns; the row maximum is exchanged through shared memory), a 4-warp epilogue group.  S is double-buffered in
// TMEM, so S of tile i+1 is computed while tile i is still in the softmax.
constexpr int FPIPE_THREADS = 448;
constexpr uint32_t FPIPE_STAGE = 3 * TILE;
constexpr int FST = 3;  // operand ring depth
constexpr uint32_t FPIPE_SMEM = FST * FPIPE_STAGE + 2 * TILE + 4 * 128 * 4 /*row max*/ + 2 * 2 * 128 * 4 /*sums*/ + 2 * 128 * 4 + 256;

template <int PACK>
__global__ void __launch_bounds__(FPIPE_THREADS, 1)
fwd_pipe_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                const __grid_constant__ CUtensorMap tv, Args a, int ntiles) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sIn = smem;                     // [FST stages][Q | K | V]
  uint8_t* sP = smem + FST * FPIPE_STAGE;  // P (dropout applied), two 64-key atoms
  float* s_hmax = reinterpret_cast<float*>(sP + 2 * TILE);  // [2 tile parities][2 halves][128] partial row maxima
  float* s_sum = s_hmax + 4 * 128;                          // [2 tile parities][2 halves][128] partial row sums
  float* s_max = s_sum + 4 * 128;                           // [2 tile parities][128] row maxima (log2 domain)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_max + 2 * 128);
  uint64_t* in_full = bars;       // [FST]
  uint64_t* in_empty = bars + 4;  // [FST]
  uint64_t* s_full = bars + 8;    // [2]  S buffer b holds the scores of a tile
  uint64_t* s_free = bars + 10;   // [2]  the softmax group has read it
  uint64_t* p_full = bars + 12;
  uint64_t* p_free = bars + 13;
  uint64_t* o_full = bars + 14;
  uint64_t* o_free = bars + 15;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < FST; ++i) {
      mbar_init(smem_u32(&in_full[i]), 1);
      mbar_init(smem_u32(&in_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&s_free[i]), 8);
    }
    mbar_init(smem_u32(p_full), 8);
    mbar_init(smem_u32(p_free), 1);
    mbar_init(smem_u32(o_full), 1);
    mbar_init(smem_u32(o_free), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;  // S0 [0,128)  S1 [128,256)  O [256,320)
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int i = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
        const int s = i % FST;
        mbar_wait(smem_u32(&in_empty[s]), ((i / FST) & 1) ^ 1);
        const uint32_t bar = smem_u32(&in_full[s]);
        const uint32_t base = smem_u32(sIn + s * FPIPE_STAGE);
        mbar_expect_tx(bar, 3 * TILE);
        issue_tile_loads<PACK>(a, &tq, base, bar, a.Lq, tile);
        issue_tile_loads<PACK>(a, &tk, base + TILE, bar, a.Lk, tile);
        issue_tile_loads<PACK>(a, &tv, base + 2 * TILE, bar, a.Lk, tile);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // Both products are issued as soon as their inputs are ready (non-blocking polls): S of tile t+1 must not wait
      // for the softmax of tile t, and O of tile t must not wait for the operands of tile t+1.
      int n_my = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) ++n_my;
      int t1 = 0, t2 = 0;  // next tile for MMA-1 (S = Q K^T) / MMA-2 (O = P V)
      const long long t0 = clock64();
      while (t2 < n_my) {
        if (t1 < n_my && t1 < t2 + 2) {
          const int st = t1 % FST, sb = t1 & 1;
          if (mbar_test(smem_u32(&in_full[st]), (t1 / FST) & 1) && mbar_test(smem_u32(&s_free[sb]), ((t1 >> 1) & 1) ^ 1)) {
            tc_fence_after();
            const uint32_t base = smem_u32(sIn + st * FPIPE_STAGE);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16(tm + sb * 128, desc_kmajor(base + k * 32), desc_kmajor(base + TILE + k * 32), idesc(128, false, false), k > 0);
            tc_commit(smem_u32(&s_full[sb]));
            ++t1;
          }
        }
        if (t2 < t1 && mbar_test(smem_u32(p_full), t2 & 1) && mbar_test(smem_u32(o_free), (t2 & 1) ^ 1)) {
          tc_fence_after();
          const int st = t2 % FST;
          const uint32_t v_a = smem_u32(sIn + st * FPIPE_STAGE + 2 * TILE);
#pragma unroll
          for (int k = 0; k < 8; ++k)  // O[128 x 64] = P[128 x 128] V[128 x 64]
            tc_mma_bf16(tm + 256, desc_kmajor(smem_u32(sP) + (k >> 2) * TILE + (k & 3) * 32),
                        desc_mnmajor(v_a + k * 2048, TILE), idesc(64, false, true), k > 0);
          tc_commit(smem_u32(o_full));
          tc_commit(smem_u32(p_free));
          tc_commit(smem_u32(&in_empty[st]));
          ++t2;
        }
        if (clock64() - t0 > 8000000000LL) __trap();  // a protocol bug must trap, never hang
      }
    }
  } else if (warp < 10) {
    // ================= softmax group =================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const float sl2 = a.scale * LOG2E;
    const bool drop = a.p_drop > 0.f;
    const uint32_t thr = drop ? drop_threshold(a.p_drop) : 0u;
    const float inv_keep = drop ? 1.f / (1.f - a.p_drop) : 1.f;
    const uint32_t dkey = drop_key(a.seed, a.site);
    constexpr int NCH = PACK == 2 ? 2 : 4;
    constexpr int CPW = NCH / 2;
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      const int sb = i & 1;
      const RowMap<PACK> rm(a, row, tile);
      const unsigned char* km = a.kmask ? a.kmask + (long long)rm.b * a.Lk : nullptr;
      const unsigned long long ebase =
          (((unsigned long long)rm.b * a.H + rm.h) * a.Lq + rm.i) * (unsigned long long)((a.Lk + 1) & ~1);
      const uint32_t ebase32 = (uint32_t)ebase;
      uint32_t vbits[NCH];
      key_valid_bits<NCH>(vbits, km, a.Lk, a.causal, rm.i, lane);
      const uint32_t t_row = tm + ((uint32_t)(q * 32) << 16) + (uint32_t)(sb * 128 + rm.key_col0);
      mbar_wait(smem_u32(&s_full[sb]), (i >> 1) & 1);
      tc_fence_after();
      float sv[CPW][32];
      float mx = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < CPW; ++cc) {
        const int ch = half * CPW + cc;
        tmem_ld32f(t_row + ch * 32, sv[cc]);
        const uint32_t vb = vbits[ch];
#pragma unroll
        for (int c = 0; c < 32; ++c) mx = fmaxf(mx, (vb >> c) & 1u ? sv[cc][c] * sl2 : -INFINITY);
      }
      // the scores are in registers: the S buffer may be overwritten by the tile after next
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_free[sb]));
      // exchange the partial maxima with the warp that owns the other half of this row quarter's columns
      s_hmax[(sb * 2 + half) * 128 + row] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      mx = fmaxf(mx, s_hmax[(sb * 2 + (half ^ 1)) * 128 + row]);
      float l = 0.f;
      mbar_wait(smem_u32(p_free), (i & 1) ^ 1);  // MMA-2 of the previous tile has read P
#pragma unroll
      for (int cc = 0; cc < CPW; ++cc) {
        const int ch = half * CPW + cc;
        const uint32_t vb = vbits[ch];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float p0 = (vb >> c) & 1u ? ex2_approx(sv[cc][c] * sl2 - mx) : 0.f;
          float p1 = (vb >> (c + 1)) & 1u ? ex2_approx(sv[cc][c + 1] * sl2 - mx) : 0.f;
          l += p0 + p1;
          if (drop) {
            const uint32_t r = drop_pair(dkey, (ebase32 + (uint32_t)(ch * 32 + c)) >> 1);
            p0 *= (r & 0xFFFFu) >= thr ? inv_keep : 0.f;
            p1 *= (r >> 16) >= thr ? inv_keep : 0.f;
          }
          sv[cc][c] = p0;
          sv[cc][c + 1] = p1;
        }
        const int kcol = rm.key_col0 + ch * 32;
        uint8_t* atom = sP + (kcol >> 6) * TILE;
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          uint4 u;
          u.x = pack2(sv[cc][c], sv[cc][c + 1]); u.y = pack2(sv[cc][c + 2], sv[cc][c + 3]);
          u.z = pack2(sv[cc][c + 4], sv[cc][c + 5]); u.w = pack2(sv[cc][c + 6], sv[cc][c + 7]);
          st_chunk(atom, row, (kcol & 63) + c, u);
        }
      }
      if (PACK == 2) {  // zero this warp's share of the other problem's key block of this row
        uint8_t* atom = sP + ((rm.key_col0 >> 6) ^ 1) * TILE;
#pragma unroll
        for (int c = half * 32; c < half * 32 + 32; c += 8) st_chunk(atom, row, c, make_uint4(0, 0, 0, 0));
      }
      s_sum[(sb * 2 + half) * 128 + row] = l;
      if (half == 0) s_max[sb * 128 + row] = mx;
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(p_full));
    }
  } else {
    // ================= epilogue group =================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      const int sb = i & 1;
      const RowMap<PACK> rm(a, row, tile);
      const bool wr = rm.exists && rm.i < a.Lq;
      bf16* orow = a.o + ((long long)rm.b * a.Lq + rm.i) * a.ldo + rm.h * 64;
      mbar_wait(smem_u32(o_full), i & 1);
      tc_fence_after();
      const float l = s_sum[(sb * 2) * 128 + row] + s_sum[(sb * 2 + 1) * 128 + row];
      const float mx = s_max[sb * 128 + row];
      const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        float v[32];
        tmem_ld32f(tm + ((uint32_t)(q * 32) << 16) + 256 + ch * 32, v);
        if (ch == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(o_free));
        }
        if (wr) {
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            uint4 u;
            u.x = pack2(v[c] * inv, v[c + 1] * inv); u.y = pack2(v[c + 2] * inv, v[c + 3] * inv);
            u.z = pack2(v[c + 4] * inv, v[c + 5] * inv); u.w = pack2(v[c + 6] * inv, v[c + 7] * inv);
            *reinterpret_cast<uint4*>(orow + ch * 32 + c) = u;
          }
        }
      }
      if (wr && a.lse) a.lse[((long long)rm.b * a.H + rm.h) * a.Lq + rm.i] = l > 0.f ? mx / LOG2E + logf(l) : -INFINITY;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ backward, pipelined
// Persistent, warp-specialised version of bwd_kernel: one CTA per SM walks over tiles; a TMA warp prefetches the next
// tile's five operand tiles (Q K V dO O) into a 2-stage ring, one thread issues the MMAs, a SOFTMAX group (4 warps)
// turns S / dP into P~ / dS, an EPILOGUE group (4 warps) drains dV / dK / dQ.  S | dP and the three outputs live in
// disjoint TMEM columns, so MMA-1 of tile i+1, the softmax of tile i+1 and the epilogue of tile i overlap; the
// single-shot kernel paid the whole load -> MMA -> softmax -> MMA -> store latency chain per CTA with only two CTAs
// per SM to hide it.  Math identical to bwd_kernel.
constexpr int PIPE_THREADS = 448;  // warp 0 TMA, warp 1 MMA, warps 2-9 softmax (two per row quarter), warps 10-13 epilogue
constexpr uint32_t PIPE_STAGE = 5 * TILE;
constexpr uint32_t PIPE_SMEM = 2 * PIPE_STAGE + 4 * TILE + 256;

template <int PACK>
__global__ void __launch_bounds__(PIPE_THREADS, 1)
bwd_pipe_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo,
                const __grid_constant__ CUtensorMap to, Args a, int ntiles) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];  // no static smem: the window starts 1024-aligned
  uint8_t* sIn = smem;                    // [2 stages][Q | K | V | dO | O]
  uint8_t* sPd = smem + 2 * PIPE_STAGE;   // P~ (dropout applied), two 64-key atoms
  uint8_t* sdS = sPd + 2 * TILE;          // dS, two 64-key atoms
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + 2 * TILE);
  uint64_t* in_full = bars;        // [2]
  uint64_t* in_empty = bars + 2;   // [2]
  uint64_t* sdp_full = bars + 4;   // S | dP of the current tile are in TMEM
  uint64_t* sdp_free = bars + 5;   // the softmax group has read them
  uint64_t* pds_full = bars + 6;   // P~ | dS are in shared memory
  uint64_t* pd_free = bars + 7;    // the MMAs reading them have retired
  uint64_t* out_full = bars + 8;   // dV | dK | dQ are in TMEM
  uint64_t* out_free = bars + 9;   // the epilogue group has read them
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&in_full[i]), 1);
      mbar_init(smem_u32(&in_empty[i]), 1);
    }
    mbar_init(smem_u32(sdp_full), 1);
    mbar_init(smem_u32(sdp_free), 8);
    mbar_init(smem_u32(pds_full), 8);
    mbar_init(smem_u32(pd_free), 1);
    mbar_init(smem_u32(out_full), 1);
    mbar_init(smem_u32(out_free), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA loader =================
      int i = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
        const int s = i & 1;
        mbar_wait(smem_u32(&in_empty[s]), ((i >> 1) & 1) ^ 1);
        const uint32_t bar = smem_u32(&in_full[s]);
        const uint32_t base = smem_u32(sIn + s * PIPE_STAGE);
        mbar_expect_tx(bar, 5 * TILE);
        issue_tile_loads<PACK>(a, &tq, base, bar, a.Lq, tile);
        issue_tile_loads<PACK>(a, &tk, base + TILE, bar, a.Lk, tile);
        issue_tile_loads<PACK>(a, &tv, base + 2 * TILE, bar, a.Lk, tile);
        issue_tile_loads<PACK>(a, &tdo, base + 3 * TILE, bar, a.Lq, tile);
        issue_tile_loads<PACK>(a, &to, base + 4 * TILE, bar, a.Lq, tile);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================= MMA issuer =================
      auto mma1 = [&](int s) {  // S = Q K^T -> cols [0,128);  dP = dO V^T -> cols [128,256)
        const uint32_t base = smem_u32(sIn + s * PIPE_STAGE);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_bf16(tm, desc_kmajor(base + k * 32), desc_kmajor(base + TILE + k * 32), idesc(128, false, false), k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_bf16(tm + 128, desc_kmajor(base + 3 * TILE + k * 32), desc_kmajor(base + 2 * TILE + k * 32),
                      idesc(128, false, false), k > 0);
        tc_commit(smem_u32(sdp_full));
      };
      // Both product groups are issued as soon as their inputs are ready (non-blocking polls): MMA-2 of tile t must
      // not wait for the operands of tile t+1, MMA-1 of tile t+1 must not wait for the softmax of tile t to finish.
      int n_my = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) ++n_my;
      int t1 = 0, t2 = 0;
      const long long t0 = clock64();
      while (t2 < n_my) {
        if (t1 < n_my && t1 < t2 + 2 && mbar_test(smem_u32(&in_full[t1 & 1]), (t1 >> 1) & 1) &&
            mbar_test(smem_u32(sdp_free), (t1 & 1) ^ 1)) {
          tc_fence_after();
          mma1(t1 & 1);
          ++t1;
        }
        if (t2 < t1 && mbar_test(smem_u32(pds_full), t2 & 1) && mbar_test(smem_u32(out_free), (t2 & 1) ^ 1)) {
          tc_fence_after();
          const int s = t2 & 1;
          const uint32_t base = smem_u32(sIn + s * PIPE_STAGE);
          const uint32_t q_a = base, k_a = base + TILE, do_a = base + 3 * TILE;
#pragma unroll
          for (int k = 0; k < 8; ++k)  // dV[keys x 64] = P~^T dO     -> cols [256,320)
            tc_mma_bf16(tm + 256, desc_mnmajor(smem_u32(sPd) + k * 2048, TILE), desc_mnmajor(do_a + k * 2048, TILE),
                        idesc(64, true, true), k > 0);
#pragma unroll
          for (int k = 0; k < 8; ++k)  // dK[keys x 64] = dS^T Q      -> cols [320,384)
            tc_mma_bf16(tm + 320, desc_mnmajor(smem_u32(sdS) + k * 2048, TILE), desc_mnmajor(q_a + k * 2048, TILE),
                        idesc(64, true, true), k > 0);
#pragma unroll
          for (int k = 0; k < 8; ++k)  // dQ[q x 64] = dS K           -> cols [384,448)
            tc_mma_bf16(tm + 384, desc_kmajor(smem_u32(sdS) + (k >> 2) * TILE + (k & 3) * 32),
                        desc_mnmajor(k_a + k * 2048, TILE), idesc(64, false, true), k > 0);
          tc_commit(smem_u32(out_full));
          tc_commit(smem_u32(pd_free));
          tc_commit(smem_u32(&in_empty[s]));
          ++t2;
        }
        if (clock64() - t0 > 8000000000LL) __trap();  // a protocol bug must trap, never hang
      }
    }
  } else if (warp < 10) {
    // ================= softmax group: S, dP -> P~, dS  (warps w and w+4 share a row quarter and split its columns) ====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const float sl2 = a.scale * LOG2E;
    const bool drop = a.p_drop > 0.f;
    const uint32_t thr = drop ? drop_threshold(a.p_drop) : 0u;
    const float inv_keep = drop ? 1.f / (1.f - a.p_drop) : 1.f;
    const uint32_t dkey = drop_key(a.seed, a.site);
    constexpr int NCH = PACK == 2 ? 2 : 4;
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      const int s = i & 1;
      const RowMap<PACK> rm(a, row, tile);
      const bool qvalid = rm.exists && rm.i < a.Lq;
      const unsigned char* km = a.kmask ? a.kmask + (long long)rm.b * a.Lk : nullptr;
      const float lse2 = qvalid ? a.lse[((long long)rm.b * a.H + rm.h) * a.Lq + rm.i] * LOG2E : 0.f;
      const unsigned long long ebase =
          (((unsigned long long)rm.b * a.H + rm.h) * a.Lq + rm.i) * (unsigned long long)((a.Lk + 1) & ~1);
      const uint32_t ebase32 = (uint32_t)ebase;
      uint32_t vbits[NCH];
      key_valid_bits<NCH>(vbits, km, a.Lk, a.causal, rm.i, lane);
      // D_i = sum_c dO[i,c] O[i,c] from the staged tiles
      mbar_wait(smem_u32(&in_full[s]), (i >> 1) & 1);
      float Di = 0.f;
      {
        const uint8_t* orow = sIn + s * PIPE_STAGE + 4 * TILE + row * 128;
        const uint8_t* drow = sIn + s * PIPE_STAGE + 3 * TILE + row * 128;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int sw = (u ^ (row & 7)) << 4;
          const uint4 x = *reinterpret_cast<const uint4*>(orow + sw), y = *reinterpret_cast<const uint4*>(drow + sw);
          const __nv_bfloat162* xb = reinterpret_cast<const __nv_bfloat162*>(&x);
          const __nv_bfloat162* yb = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const float2 xf = __bfloat1622float2(xb[w]), yf = __bfloat1622float2(yb[w]);
            Di += xf.x * yf.x + xf.y * yf.y;
          }
        }
        if (!qvalid) Di = 0.f;
      }
      const uint32_t t_row = tm + ((uint32_t)(q * 32) << 16) + (uint32_t)rm.key_col0;
      mbar_wait(smem_u32(sdp_full), i & 1);
      tc_fence_after();
      constexpr int CPW = NCH / 2;  // chunks per warp
#pragma unroll 1
      for (int ch = half * CPW; ch < (half + 1) * CPW; ++ch) {
        float sv[32], dp[32];
        tmem_ld32f(t_row + ch * 32, sv);
        tmem_ld32f(t_row + 128 + ch * 32, dp);
        if (ch == (half + 1) * CPW - 1) {  // S | dP of this tile are in registers: MMA-1 of the next tile may overwrite them
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(sdp_free));
        }
        const uint32_t vb = qvalid ? vbits[ch] : 0u;
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float p0 = (vb >> c) & 1u ? ex2_approx(sv[c] * sl2 - lse2) : 0.f;
          const float p1 = (vb >> (c + 1)) & 1u ? ex2_approx(sv[c + 1] * sl2 - lse2) : 0.f;
          float k0 = 1.f, k1 = 1.f;
          if (drop) {
            const uint32_t r = drop_pair(dkey, (ebase32 + (uint32_t)(ch * 32 + c)) >> 1);
            k0 = (r & 0xFFFFu) >= thr ? inv_keep : 0.f;
            k1 = (r >> 16) >= thr ? inv_keep : 0.f;
          }
          sv[c] = p0 * k0;
          sv[c + 1] = p1 * k1;
          dp[c] = p0 * (dp[c] * k0 - Di);
          dp[c + 1] = p1 * (dp[c + 1] * k1 - Di);
        }
        if (ch == half * CPW) mbar_wait(smem_u32(pd_free), (i & 1) ^ 1);  // MMA-2 of the previous tile has read P~ | dS
        const int kcol = rm.key_col0 + ch * 32;
        uint8_t* pa = sPd + (kcol >> 6) * TILE;
        uint8_t* da = sdS + (kcol >> 6) * TILE;
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          uint4 u, w;
          u.x = pack2(sv[c], sv[c + 1]); u.y = pack2(sv[c + 2], sv[c + 3]);
          u.z = pack2(sv[c + 4], sv[c + 5]); u.w = pack2(sv[c + 6], sv[c + 7]);
          w.x = pack2(dp[c], dp[c + 1]); w.y = pack2(dp[c + 2], dp[c + 3]);
          w.z = pack2(dp[c + 4], dp[c + 5]); w.w = pack2(dp[c + 6], dp[c + 7]);
          st_chunk(pa, row, (kcol & 63) + c, u);
          st_chunk(da, row, (kcol & 63) + c, w);
        }
      }
      if (PACK == 2) {
        const int other = (rm.key_col0 >> 6) ^ 1;
#pragma unroll
        for (int c = half * 32; c < half * 32 + 32; c += 8) {
          st_chunk(sPd + other * TILE, row, c, make_uint4(0, 0, 0, 0));
          st_chunk(sdS + other * TILE, row, c, make_uint4(0, 0, 0, 0));
        }
      }
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(pds_full));
    }
  } else {
    // ================= epilogue group: dV, dK, dQ -> global =================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      const RowMap<PACK> rm(a, row, tile);
      const bool qvalid = rm.exists && rm.i < a.Lq;
      const bool kvalid = rm.exists && rm.i < a.Lk;
      bf16* dqrow = a.dq + ((long long)rm.b * a.Lq + rm.i) * a.lddq + rm.h * 64;
      bf16* dkrow = a.dk + ((long long)rm.b * a.Lk + rm.i) * a.lddk + rm.h * 64;
      bf16* dvrow = a.dv + ((long long)rm.b * a.Lk + rm.i) * a.lddv + rm.h * 64;
      const uint32_t tl = tm + ((uint32_t)(q * 32) << 16) + 256;
      mbar_wait(smem_u32(out_full), i & 1);
      tc_fence_after();
#pragma unroll 1
      for (int part = 0; part < 6; ++part) {  // dV lo/hi, dK lo/hi, dQ lo/hi
        float v[32];
        tmem_ld32f(tl + part * 32, v);
        if (part == 5) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(out_free));
        }
        const int which = part >> 1, half = part & 1;
        const bool wr = which == 2 ? qvalid : kvalid;
        const float sc = which == 0 ? 1.f : a.scale;
        bf16* dst = (which == 0 ? dvrow : which == 1 ? dkrow : dqrow) + half * 32;
        if (wr) {
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            uint4 u;
            u.x = pack2(v[c] * sc, v[c + 1] * sc); u.y = pack2(v[c + 2] * sc, v[c + 3] * sc);
            u.z = pack2(v[c + 4] * sc, v[c + 5] * sc); u.w = pack2(v[c + 6] * sc, v[c + 7] * sc);
            *reinterpret_cast<uint4*>(dst + c) = u;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Sequences longer than one tile (128 < L <= 512: the multimodal encoder, S ~ 200-300): BLOCKED attention.
// The (Lq x Lk) score matrix of a (batch, head) is cut into 128 x 128 blocks; every block is one single-shot tcgen05
// problem exactly like fwd_kernel<1> / bwd_kernel<1> above (same TMEM / shared-memory plan), addressed by block offsets:
//   forward   block (qb, kb) -> O_kb = softmax_block(S) V_kb (normalised inside the block, fp32) and the block's
//             log-sum-exp; merge_fwd combines the key blocks: lse = log sum_kb exp(lse_kb), O = sum_kb exp(lse_kb - lse) O_kb
//   backward  with the MERGED lse and D_i = rowsum(dO o O) every block is independent: P = exp(S - lse), dS = P o (dP - D);
//             block (qb, kb) contributes dQ_qb += dS K_kb, dK_kb += dS^T Q_qb, dV_kb += P^T dO_qb - written as bf16
//             partials indexed by the OTHER block coordinate and summed by merge_bwd (no atomics: deterministic).
// Masks (key padding, causal) and the dropout stream use GLOBAL (i, j), so results equal the single-tile / streaming
// kernels' for the same seed.  The elementwise softmax / dropout work per score - not the MMA - bounds attention at
// these sizes (ncu: the streaming mma.sync kernel issues ~40 instructions per score); the single-shot tile kernels do
// it with one thread per row out of TMEM at ~3.6x the streaming kernel's rate per score.
// ------------------------------------------------------------------------------------------------------------------
struct BlkArgs {
  int B, H, Lq, Lk, causal, nqb, nkb;
  float scale, p_drop;
  unsigned long long seed;
  unsigned int site;
  const unsigned char* kmask;
  bf16* opart;      // fwd out: [nkb][B*Lq][H*64]  (partials travel as bf16: half the bytes, merged in fp32)
  float* lsepart;   // fwd out: [nkb][B*H*Lq]
  const float* lse; // bwd in: merged [B*H*Lq]
  bf16* dqpart;     // bwd out: [nkb][B*Lq][H*64]
  bf16* dkpart;     // bwd out: [nqb][B*Lk][H*64]
  bf16* dvpart;     // bwd out: [nqb][B*Lk][H*64]
};

struct BlkMap {
  int b, h, qb, kb;
  __device__ BlkMap(const BlkArgs& a, int prob) {
    kb = prob % a.nkb;
    prob /= a.nkb;
    qb = prob % a.nqb;
    prob /= a.nqb;
    h = prob % a.H;
    b = prob / a.H;
  }
};

// two 64-row boxes of the 128-row block `blk` of batch b (rows past the batch's length are masked by the callers)
__device__ __forceinline__ void issue_blk_loads(const CUtensorMap* tm, uint32_t dst, uint32_t bar, int b, int h, int L, int blk) {
#pragma unroll
  for (int s = 0; s < 2; ++s) tma_load_2d(dst + s * 8192, tm, bar, h * 64, b * L + blk * 128 + s * 64);
}

// validity bits of the keys [k0 + ch*32, +32): global key < Lk, key-padding mask, causal j <= i
__device__ __forceinline__ void key_valid_bits_blk(uint32_t (&bits)[4], const unsigned char* km, int Lk, int causal, int i,
                                                   int k0, int lane) {
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    const int j = k0 + ch * 32 + lane;
    const bool ok = j < Lk && (!km || km[j] != 0);
    uint32_t m = __ballot_sync(0xffffffffu, ok);
    if (causal) {
      const int hi = i - (k0 + ch * 32);
      m &= hi >= 31 ? 0xffffffffu : hi < 0 ? 0u : ((2u << hi) - 1u);
    }
    bits[ch] = m;
  }
}

__global__ void __launch_bounds__(128) fwd_blk_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                                                      const __grid_constant__ CUtensorMap tv, BlkArgs a) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;  // after S is formed, sQ|sK is reused for P (two 64-key atoms)
  uint8_t* sK = smem + TILE;
  uint8_t* sV = smem + 2 * TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * TILE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const BlkMap bm(a, blockIdx.x);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  pdl_wait();
  if (threadIdx.x == 0) {
    const uint32_t bar = smem_u32(&bars[0]);
    mbar_expect_tx(bar, 3 * TILE);
    issue_blk_loads(&tq, smem_u32(sQ), bar, bm.b, bm.h, a.Lq, bm.qb);
    issue_blk_loads(&tk, smem_u32(sK), bar, bm.b, bm.h, a.Lk, bm.kb);
    issue_blk_loads(&tv, smem_u32(sV), bar, bm.b, bm.h, a.Lk, bm.kb);
    mbar_wait(bar, 0);
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < 4; ++k)  // S[128 x 128] = Q K^T
      tc_mma_bf16(tm, desc_kmajor(smem_u32(sQ) + k * 32), desc_kmajor(smem_u32(sK) + k * 32), idesc(128, false, false), k > 0);
    tc_commit(smem_u32(&bars[1]));
  }
  const int row = warp * 32 + lane;
  const int i = bm.qb * 128 + row;  // global query index
  const int k0 = bm.kb * 128;
  const unsigned char* km = a.kmask ? a.kmask + (long long)bm.b * a.Lk : nullptr;
  const float sl2 = a.scale * LOG2E;
  const bool drop = a.p_drop > 0.f;
  const uint32_t thr = drop ? drop_threshold(a.p_drop) : 0u;
  const float inv_keep = drop ? 1.f / (1.f - a.p_drop) : 1.f;
  const uint32_t dkey = drop_key(a.seed, a.site);
  const unsigned long long ebase = (((unsigned long long)bm.b * a.H + bm.h) * a.Lq + i) * (unsigned long long)((a.Lk + 1) & ~1) + k0;
  const uint32_t ebase32 = (uint32_t)ebase;  // even
  const uint32_t t_row = tm + ((uint32_t)(warp * 32) << 16);
  uint32_t vbits[4];
  key_valid_bits_blk(vbits, km, a.Lk, a.causal, i, k0, lane);

  if (lane == 0) mbar_wait(smem_u32(&bars[1]), 0);
  __syncwarp();
  tc_fence_after();
  float mx = -INFINITY;
#pragma unroll 1
  for (int ch = 0; ch < 4; ++ch) {
    float s[32];
    tmem_ld32f(t_row + ch * 32, s);
    const uint32_t vb = vbits[ch];
#pragma unroll
    for (int c = 0; c < 32; ++c) mx = fmaxf(mx, (vb >> c) & 1u ? s[c] * sl2 : -INFINITY);
  }
  float l = 0.f;
  uint8_t* sP = sQ;
#pragma unroll 1
  for (int ch = 0; ch < 4; ++ch) {
    float s[32];
    tmem_ld32f(t_row + ch * 32, s);
    const uint32_t vb = vbits[ch];
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
      float p0 = (vb >> c) & 1u ? ex2_approx(s[c] * sl2 - mx) : 0.f;
      float p1 = (vb >> (c + 1)) & 1u ? ex2_approx(s[c + 1] * sl2 - mx) : 0.f;
      l += p0 + p1;
      if (drop) {
        const uint32_t r = drop_pair(dkey, (ebase32 + (uint32_t)(ch * 32 + c)) >> 1);
        p0 *= (r & 0xFFFFu) >= thr ? inv_keep : 0.f;
        p1 *= (r >> 16) >= thr ? inv_keep : 0.f;
      }
      s[c] = p0;
      s[c + 1] = p1;
    }
    const int kcol = ch * 32;
    uint8_t* atom = sP + (kcol >> 6) * TILE;
#pragma unroll
    for (int c = 0; c < 32; c += 8) {
      uint4 u;
      u.x = pack2(s[c], s[c + 1]); u.y = pack2(s[c + 2], s[c + 3]);
      u.z = pack2(s[c + 4], s[c + 5]); u.w = pack2(s[c + 6], s[c + 7]);
      st_chunk(atom, row, (kcol & 63) + c, u);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < 8; ++k)  // O[128 x 64] = P[128 x 128] V[128 x 64]
      tc_mma_bf16(tm, desc_kmajor(smem_u32(sP) + (k >> 2) * TILE + (k & 3) * 32),
                  desc_mnmajor(smem_u32(sV) + k * 2048, TILE), idesc(64, false, true), k > 0);
    tc_commit(smem_u32(&bars[2]));
  }
  if (lane == 0) mbar_wait(smem_u32(&bars[2]), 0);
  __syncwarp();
  tc_fence_after();
  {
    const float inv = l > 0.f ? 1.f / l : 0.f;
    const bool wr = i < a.Lq;
    const long long part = (long long)bm.kb * a.B * a.Lq;
    bf16* orow = a.opart + (part + (long long)bm.b * a.Lq + i) * (a.H * 64) + bm.h * 64;
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      float v[32];
      tmem_ld32f(tm + ((uint32_t)(warp * 32) << 16) + ch * 32, v);
      if (wr) {
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          uint4 u;
          u.x = pack2(v[c] * inv, v[c + 1] * inv); u.y = pack2(v[c + 2] * inv, v[c + 3] * inv);
          u.z = pack2(v[c + 4] * inv, v[c + 5] * inv); u.w = pack2(v[c + 6] * inv, v[c + 7] * inv);
          *reinterpret_cast<uint4*>(orow + ch * 32 + c) = u;
        }
      }
    }
    if (wr)
      a.lsepart[(long long)bm.kb * a.B * a.H * a.Lq + ((long long)bm.b * a.H + bm.h) * a.Lq + i] =
          l > 0.f ? mx / LOG2E + logf(l) : -INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128) : "memory");
  }
}

// one thread per (row, head, 8 columns): O = sum_kb w_kb O_kb, lse = m + log sum_kb exp(lse_kb - m)
__global__ void __launch_bounds__(256) merge_fwd_kernel(const bf16* __restrict__ opart, const float* __restrict__ lsepart, bf16* o,
                                                        long long ldo, float* __restrict__ lse, int B, int H, int Lq, int nkb) {
  pdl_trigger();
  const long long rows = (long long)B * Lq;
  const long long total = rows * H * 8;
  const int W = H * 64;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx & 7);
    const long long rh = idx >> 3;
    const long long r = rh / H;
    const int h = (int)(rh % H);
    const int b = (int)(r / Lq), i = (int)(r % Lq);
    const long long li = ((long long)b * H + h) * Lq + i;
    float lp[4], m = -INFINITY;
    for (int kb = 0; kb < nkb; ++kb) {
      lp[kb] = lsepart[(long long)kb * B * H * Lq + li];
      m = fmaxf(m, lp[kb]);
    }
    float L = 0.f, acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    for (int kb = 0; kb < nkb; ++kb) {
      const float w = lp[kb] == -INFINITY ? 0.f : __expf(lp[kb] - m);
      L += w;
      const uint4 raw = *reinterpret_cast<const uint4*>(opart + ((long long)kb * rows + r) * W + h * 64 + c8 * 8);
      const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int q2 = 0; q2 < 4; ++q2) {
        const float2 f = __bfloat1622float2(hp[q2]);
        acc[2 * q2] = fmaf(w, f.x, acc[2 * q2]);
        acc[2 * q2 + 1] = fmaf(w, f.y, acc[2 * q2 + 1]);
      }
    }
    const float inv = L > 0.f ? 1.f / L : 0.f;
    uint4 u;
    u.x = pack2(acc[0] * inv, acc[1] * inv); u.y = pack2(acc[2] * inv, acc[3] * inv);
    u.z = pack2(acc[4] * inv, acc[5] * inv); u.w = pack2(acc[6] * inv, acc[7] * inv);
    *reinterpret_cast<uint4*>(o + r * ldo + h * 64 + c8 * 8) = u;
    if (c8 == 0 && lse) lse[li] = L > 0.f ? m + logf(L) : -INFINITY;
  }
}

__global__ void __launch_bounds__(128) bwd_blk_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                                                      const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo,
                                                      const __grid_constant__ CUtensorMap to, BlkArgs a) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  // layout (ascending): Pd atom 0 | dS atom 0 | dS atom 1 | Q | K | dO | V (= Pd atom 1 once dP has retired)
  uint8_t* sPd0 = smem;
  uint8_t* sdS = smem + TILE;
  uint8_t* sQ = smem + 3 * TILE;
  uint8_t* sK = smem + 4 * TILE;
  uint8_t* sDO = smem + 5 * TILE;
  uint8_t* sV = smem + 6 * TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * TILE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const BlkMap bm(a, blockIdx.x);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  pdl_wait();
  if (threadIdx.x == 0) {
    const uint32_t bar = smem_u32(&bars[0]);
    mbar_expect_tx(bar, 5 * TILE);
    issue_blk_loads(&tq, smem_u32(sQ), bar, bm.b, bm.h, a.Lq, bm.qb);
    issue_blk_loads(&tk, smem_u32(sK), bar, bm.b, bm.h, a.Lk, bm.kb);
    issue_blk_loads(&tv, smem_u32(sV), bar, bm.b, bm.h, a.Lk, bm.kb);
    issue_blk_loads(&tdo, smem_u32(sDO), bar, bm.b, bm.h, a.Lq, bm.qb);
    issue_blk_loads(&to, smem_u32(sdS), bar, bm.b, bm.h, a.Lq, bm.qb);  // O parks in dS atom 0 until D_i is formed
    mbar_wait(bar, 0);
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < 4; ++k)  // S = Q K^T -> cols [0,128)
      tc_mma_bf16(tm, desc_kmajor(smem_u32(sQ) + k * 32), desc_kmajor(smem_u32(sK) + k * 32), idesc(128, false, false), k > 0);
#pragma unroll
    for (int k = 0; k < 4; ++k)  // dP = dO V^T -> cols [128,256)
      tc_mma_bf16(tm + 128, desc_kmajor(smem_u32(sDO) + k * 32), desc_kmajor(smem_u32(sV) + k * 32), idesc(128, false, false), k > 0);
    tc_commit(smem_u32(&bars[1]));
  }
  const int row = warp * 32 + lane;
  const int i = bm.qb * 128 + row;   // global query index of this row (dQ, softmax)
  const int jk = bm.kb * 128 + row;  // global key index of this row (dK, dV)
  const int k0 = bm.kb * 128;
  const bool qvalid = i < a.Lq;
  const unsigned char* km = a.kmask ? a.kmask + (long long)bm.b * a.Lk : nullptr;
  float Di = 0.f, lse2 = 0.f;
  if (qvalid) lse2 = a.lse[((long long)bm.b * a.H + bm.h) * a.Lq + i] * LOG2E;
  if (lane == 0) mbar_wait(smem_u32(&bars[0]), 0);
  __syncwarp();
  {
    const uint8_t* orow = sdS + row * 128;
    const uint8_t* drow = sDO + row * 128;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int sw = (u ^ (row & 7)) << 4;
      const uint4 x = *reinterpret_cast<const uint4*>(orow + sw), y = *reinterpret_cast<const uint4*>(drow + sw);
      const __nv_bfloat162* xb = reinterpret_cast<const __nv_bfloat162*>(&x);
      const __nv_bfloat162* yb = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float2 xf = __bfloat1622float2(xb[w]), yf = __bfloat1622float2(yb[w]);
        Di += xf.x * yf.x + xf.y * yf.y;
      }
    }
    if (!qvalid) Di = 0.f;
  }
  const float sl2 = a.scale * LOG2E;
  const bool drop = a.p_drop > 0.f;
  const uint32_t thr = drop ? drop_threshold(a.p_drop) : 0u;
  const float inv_keep = drop ? 1.f / (1.f - a.p_drop) : 1.f;
  const uint32_t dkey = drop_key(a.seed, a.site);
  const unsigned long long ebase = (((unsigned long long)bm.b * a.H + bm.h) * a.Lq + i) * (unsigned long long)((a.Lk + 1) & ~1) + k0;
  const uint32_t ebase32 = (uint32_t)ebase;
  const uint32_t t_row = tm + ((uint32_t)(warp * 32) << 16);
  uint32_t vbits[4];
  key_valid_bits_blk(vbits, km, a.Lk, a.causal, i, k0, lane);
  // a fully masked row has lse = -inf: exp2(s - (-inf)) would be inf; its probabilities are all zero
  const bool row_live = qvalid && lse2 > -INFINITY;

  if (lane == 0) mbar_wait(smem_u32(&bars[1]), 0);
  __syncwarp();
  tc_fence_after();
#pragma unroll 1
  for (int ch = 0; ch < 4; ++ch) {
    float s[32], dp[32];
    tmem_ld32f(t_row + ch * 32, s);
    tmem_ld32f(t_row + 128 + ch * 32, dp);
    const uint32_t vb = row_live ? vbits[ch] : 0u;
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
      const float p0 = (vb >> c) & 1u ? ex2_approx(s[c] * sl2 - lse2) : 0.f;
      const float p1 = (vb >> (c + 1)) & 1u ? ex2_approx(s[c + 1] * sl2 - lse2) : 0.f;
      float kk0 = 1.f, kk1 = 1.f;
      if (drop) {
        const uint32_t r = drop_pair(dkey, (ebase32 + (uint32_t)(ch * 32 + c)) >> 1);
        kk0 = (r & 0xFFFFu) >= thr ? inv_keep : 0.f;
        kk1 = (r >> 16) >= thr ? inv_keep : 0.f;
      }
      s[c] = p0 * kk0;                        // P_drop
      s[c + 1] = p1 * kk1;
      dp[c] = p0 * (dp[c] * kk0 - Di);        // dS
      dp[c + 1] = p1 * (dp[c + 1] * kk1 - Di);
    }
    const int kcol = ch * 32;
    uint8_t* pa = (kcol >> 6) ? sV : sPd0;
    uint8_t* da = sdS + (kcol >> 6) * TILE;
#pragma unroll
    for (int c = 0; c < 32; c += 8) {
      uint4 u, w;
      u.x = pack2(s[c], s[c + 1]); u.y = pack2(s[c + 2], s[c + 3]);
      u.z = pack2(s[c + 4], s[c + 5]); u.w = pack2(s[c + 6], s[c + 7]);
      w.x = pack2(dp[c], dp[c + 1]); w.y = pack2(dp[c + 2], dp[c + 3]);
      w.z = pack2(dp[c + 4], dp[c + 5]); w.w = pack2(dp[c + 6], dp[c + 7]);
      st_chunk(pa, row, (kcol & 63) + c, u);
      st_chunk(da, row, (kcol & 63) + c, w);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    const uint32_t pd_stride = smem_u32(sV) - smem_u32(sPd0);
#pragma unroll
    for (int k = 0; k < 8; ++k)  // dV[keys x 64] = Pd^T dO     -> cols [0,64)
      tc_mma_bf16(tm, desc_mnmajor(smem_u32(sPd0) + k * 2048, pd_stride), desc_mnmajor(smem_u32(sDO) + k * 2048, TILE),
                  idesc(64, true, true), k > 0);
#pragma unroll
    for (int k = 0; k < 8; ++k)  // dK[keys x 64] = dS^T Q      -> cols [64,128)
      tc_mma_bf16(tm + 64, desc_mnmajor(smem_u32(sdS) + k * 2048, TILE), desc_mnmajor(smem_u32(sQ) + k * 2048, TILE),
                  idesc(64, true, true), k > 0);
#pragma unroll
    for (int k = 0; k < 8; ++k)  // dQ[q x 64] = dS K           -> cols [128,192)
      tc_mma_bf16(tm + 128, desc_kmajor(smem_u32(sdS) + (k >> 2) * TILE + (k & 3) * 32),
                  desc_mnmajor(smem_u32(sK) + k * 2048, TILE), idesc(64, false, true), k > 0);
    tc_commit(smem_u32(&bars[2]));
  }
  if (lane == 0) mbar_wait(smem_u32(&bars[2]), 0);
  __syncwarp();
  tc_fence_after();
  {
    const bool kvalid = jk < a.Lk;
    const int W = a.H * 64;
    bf16* dqrow = a.dqpart + ((long long)bm.kb * a.B * a.Lq + (long long)bm.b * a.Lq + i) * W + bm.h * 64;
    bf16* dkrow = a.dkpart + ((long long)bm.qb * a.B * a.Lk + (long long)bm.b * a.Lk + jk) * W + bm.h * 64;
    bf16* dvrow = a.dvpart + ((long long)bm.qb * a.B * a.Lk + (long long)bm.b * a.Lk + jk) * W + bm.h * 64;
    const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int part = 0; part < 6; ++part) {  // dV lo/hi, dK lo/hi, dQ lo/hi
      float v[32];
      tmem_ld32f(tl + part * 32, v);
      const int which = part >> 1, half = part & 1;
      const bool wr = which == 2 ? qvalid : kvalid;
      const float sc = which == 0 ? 1.f : a.scale;
      bf16* dst = (which == 0 ? dvrow : which == 1 ? dkrow : dqrow) + half * 32;
      if (wr) {
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          uint4 u;
          u.x = pack2(v[c] * sc, v[c + 1] * sc); u.y = pack2(v[c + 2] * sc, v[c + 3] * sc);
          u.z = pack2(v[c + 4] * sc, v[c + 5] * sc); u.w = pack2(v[c + 6] * sc, v[c + 7] * sc);
          *reinterpret_cast<uint4*>(dst + c) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256) : "memory");
  }
}

// out[r, c] (bf16, pitch ld) = sum_p part[p][r][c]  for the three gradients (dQ over key blocks, dK / dV over query blocks)
__global__ void __launch_bounds__(256) merge_bwd_kernel(const bf16* __restrict__ dqp, const bf16* __restrict__ dkp,
                                                        const bf16* __restrict__ dvp, bf16* dq, long long lddq, bf16* dk,
                                                        long long lddk, bf16* dv, long long lddv, long long rows_q,
                                                        long long rows_k, int W, int nkb, int nqb) {
  pdl_trigger();
  const long long per_q = rows_q * (W / 8), per_k = rows_k * (W / 8);
  const long long total = per_q + 2 * per_k;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const bf16* src;
    bf16* dst;
    long long ld, rows, e = idx;
    int np;
    if (e < per_q) { src = dqp; dst = dq; ld = lddq; rows = rows_q; np = nkb; }
    else if (e < per_q + per_k) { e -= per_q; src = dkp; dst = dk; ld = lddk; rows = rows_k; np = nqb; }
    else { e -= per_q + per_k; src = dvp; dst = dv; ld = lddv; rows = rows_k; np = nqb; }
    const long long r = e / (W / 8);
    const int c = (int)(e % (W / 8)) * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int p = 0; p < np; ++p) {
      const uint4 raw = *reinterpret_cast<const uint4*>(src + ((long long)p * rows + r) * W + c);
      const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int q2 = 0; q2 < 4; ++q2) {
        const float2 f = __bfloat1622float2(hp[q2]);
        acc[2 * q2] += f.x;
        acc[2 * q2 + 1] += f.y;
      }
    }
    uint4 u;
    u.x = pack2(acc[0], acc[1]); u.y = pack2(acc[2], acc[3]); u.z = pack2(acc[4], acc[5]); u.w = pack2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(dst + r * ld + c) = u;
  }
}

static int map64(CUtensorMap* m, const void* p, int H, long long rows, long long ld) {
  return make_map(m, p, (unsigned long long)H * 64, (unsigned long long)rows, (unsigned long long)ld, 64, 64);
}

}  // namespace at5

// Single-tile tcgen05 attention.  Requirements (checked): bf16, head dim 64, Lk <= 128, Lq <= 128, 16-byte aligned
// views with pitches that are multiples of 8 elements.  q/k/v/o/dout are [B*L, ld] views with head h at columns
// [64h, 64h+64).
extern "C" int mma_attn_fwd_t5(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                               long long ldv, const unsigned char* kmask, void* o, long long ldo, float* lse, int B,
                               int H, int Lq, int Lk, int causal, float scale, float p_drop, unsigned long long seed,
                               unsigned int site, cudaStream_t stream) {
  using namespace at5;
  if (B <= 0 || Lq <= 0 || Lk <= 0) return MMA_OK;
  if (Lq > 128 || Lk > 128 || ((ldq | ldk | ldv | ldo) & 7)) return MMA_ERR_UNSUPPORTED;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = map64(&tq, q, H, (long long)B * Lq, ldq))) return rc;
  if ((rc = map64(&tk, k, H, (long long)B * Lk, ldk))) return rc;
  if ((rc = map64(&tv, v, H, (long long)B * Lk, ldv))) return rc;
  Args a{};
  a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk; a.causal = causal; a.nprob = B * H; a.scale = scale; a.p_drop = p_drop;
  a.seed = seed; a.site = site; a.kmask = kmask; a.o = (bf16*)o; a.ldo = ldo; a.lse = lse;
  static int pipe = -1;
  if (pipe < 0) {
    // measured on B200: the 4-CTA/SM single-shot forward kernel already overlaps its latency chain (and wins
    // clearly when there are few tiles, e.g. the decode cross-attention); the persistent variant is opt-in
    const char* e = getenv("MMA_ATTN_FWD_PIPE");
    pipe = e ? atoi(e) : 0;
  }
  if (pipe) {
    static int sms = 0;
    if (!sms) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const bool pack2 = Lq <= 64 && Lk <= 64;
    const int ntiles = pack2 ? (a.nprob + 1) / 2 : a.nprob;
    const int grid = ntiles < sms ? ntiles : sms;
    static bool set2 = false, set1 = false;
    if (pack2) {
      if (!set2) { cudaFuncSetAttribute(fwd_pipe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FPIPE_SMEM); set2 = true; }
      if (launch_pdl(fwd_pipe_kernel<2>, dim3(grid), dim3(FPIPE_THREADS), FPIPE_SMEM, stream, tq, tk, tv, a, ntiles) != cudaSuccess) return MMA_ERR_LAUNCH;
    } else {
      if (!set1) { cudaFuncSetAttribute(fwd_pipe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FPIPE_SMEM); set1 = true; }
      if (launch_pdl(fwd_pipe_kernel<1>, dim3(grid), dim3(FPIPE_THREADS), FPIPE_SMEM, stream, tq, tk, tv, a, ntiles) != cudaSuccess) return MMA_ERR_LAUNCH;
    }
    MMA_CHECK_LAUNCH();
    return MMA_OK;
  }
  const int smem = 3 * TILE + 64;
  if (Lq <= 64 && Lk <= 64) {
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); set = true; }
    if (launch_pdl(fwd_kernel<2>, dim3((a.nprob + 1) / 2), dim3(128), smem, stream, tq, tk, tv, a) != cudaSuccess) return MMA_ERR_LAUNCH;
  } else {
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); set = true; }
    if (launch_pdl(fwd_kernel<1>, dim3(a.nprob), dim3(128), smem, stream, tq, tk, tv, a) != cudaSuccess) return MMA_ERR_LAUNCH;
  }
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

extern "C" int mma_attn_bwd_t5(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                               long long ldv, const unsigned char* kmask, const void* o, long long ldo,
                               const float* lse, const void* dout, long long lddo, void* dq, long long lddq, void* dk,
                               long long lddk, void* dv, long long lddv, int B, int H, int Lq, int Lk, int causal,
                               float scale, float p_drop, unsigned long long seed, unsigned int site,
                               cudaStream_t stream) {
  using namespace at5;
  if (B <= 0 || Lq <= 0 || Lk <= 0) return MMA_OK;
  if (Lq > 128 || Lk > 128 || ((ldq | ldk | ldv | ldo | lddo | lddq | lddk | lddv) & 7)) return MMA_ERR_UNSUPPORTED;
  CUtensorMap tq, tk, tv, tdo, to;
  int rc;
  if ((rc = map64(&tq, q, H, (long long)B * Lq, ldq))) return rc;
  if ((rc = map64(&tk, k, H, (long long)B * Lk, ldk))) return rc;
  if ((rc = map64(&tv, v, H, (long long)B * Lk, ldv))) return rc;
  if ((rc = map64(&tdo, dout, H, (long long)B * Lq, lddo))) return rc;
  if ((rc = map64(&to, o, H, (long long)B * Lq, ldo))) return rc;
  Args a{};
  a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk; a.causal = causal; a.nprob = B * H; a.scale = scale; a.p_drop = p_drop;
  a.seed = seed; a.site = site; a.kmask = kmask; a.lse = const_cast<float*>(lse);
  a.o_in = (const bf16*)o; a.ldo = ldo; a.dout = (const bf16*)dout; a.lddo = lddo;
  a.dq = (bf16*)dq; a.dk = (bf16*)dk; a.dv = (bf16*)dv; a.lddq = lddq; a.lddk = lddk; a.lddv = lddv;
  static int pipe = -1;
  if (pipe < 0) {
    const char* e = getenv("MMA_ATTN_PIPE");
    pipe = e ? atoi(e) : 1;
  }
  if (pipe) {
    static int sms = 0;
    if (!sms) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const bool pack2 = Lq <= 64 && Lk <= 64;
    const int ntiles = pack2 ? (a.nprob + 1) / 2 : a.nprob;
    const int grid = ntiles < sms ? ntiles : sms;
    static bool set2 = false, set1 = false;
    if (pack2) {
      if (!set2) { cudaFuncSetAttribute(bwd_pipe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM); set2 = true; }
      if (launch_pdl(bwd_pipe_kernel<2>, dim3(grid), dim3(PIPE_THREADS), PIPE_SMEM, stream, tq, tk, tv, tdo, to, a, ntiles) != cudaSuccess) return MMA_ERR_LAUNCH;
    } else {
      if (!set1) { cudaFuncSetAttribute(bwd_pipe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM); set1 = true; }
      if (launch_pdl(bwd_pipe_kernel<1>, dim3(grid), dim3(PIPE_THREADS), PIPE_SMEM, stream, tq, tk, tv, tdo, to, a, ntiles) != cudaSuccess) return MMA_ERR_LAUNCH;
    }
    MMA_CHECK_LAUNCH();
    return MMA_OK;
  }
  const int smem = 7 * TILE + 64;  // 2 CTAs per SM
  if (Lq <= 64 && Lk <= 64) {
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); set = true; }
    if (launch_pdl(bwd_kernel<2>, dim3((a.nprob + 1) / 2), dim3(128), smem, stream, tq, tk, tv, tdo, to, a) != cudaSuccess) return MMA_ERR_LAUNCH;
  } else {
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); set = true; }
    if (launch_pdl(bwd_kernel<1>, dim3(a.nprob), dim3(128), smem, stream, tq, tk, tv, tdo, to, a) != cudaSuccess) return MMA_ERR_LAUNCH;
  }
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

// Blocked tcgen05 attention for sequences longer than one 128-row tile (Lq, Lk <= 512; bf16, head dim 64).
// ws_o: bf16 [nkb][B*Lq][H*64], ws_lse: fp32 [nkb][B*H*Lq] with nkb = ceil(Lk / 128) (caller-owned workspaces).
extern "C" int mma_attn_fwd_t5b(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                                const unsigned char* kmask, void* o, long long ldo, float* lse, void* ws_o, float* ws_lse,
                                int B, int H, int Lq, int Lk, int causal, float scale, float p_drop, unsigned long long seed,
                                unsigned int site, cudaStream_t stream) {
  using namespace at5;
  if (B <= 0 || Lq <= 0 || Lk <= 0) return MMA_OK;
  if (Lq > 512 || Lk > 512 || ((ldq | ldk | ldv | ldo) & 7) || !ws_o || !ws_lse) return MMA_ERR_UNSUPPORTED;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = map64(&tq, q, H, (long long)B * Lq, ldq))) return rc;
  if ((rc = map64(&tk, k, H, (long long)B * Lk, ldk))) return rc;
  if ((rc = map64(&tv, v, H, (long long)B * Lk, ldv))) return rc;
  BlkArgs a{};
  a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk; a.causal = causal; a.nqb = (Lq + 127) / 128; a.nkb = (Lk + 127) / 128;
  a.scale = scale; a.p_drop = p_drop; a.seed = seed; a.site = site; a.kmask = kmask; a.opart = (bf16*)ws_o; a.lsepart = ws_lse;
  const int smem = 3 * TILE + 64;
  static bool set = false;
  if (!set) { cudaFuncSetAttribute(fwd_blk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); set = true; }
  const long long nprob = (long long)B * H * a.nqb * a.nkb;
  if (launch_pdl(fwd_blk_kernel, dim3((unsigned)nprob), dim3(128), smem, stream, tq, tk, tv, a) != cudaSuccess) return MMA_ERR_LAUNCH;
  merge_fwd_kernel<<<148 * 8, 256, 0, stream>>>((const bf16*)ws_o, ws_lse, (bf16*)o, ldo, lse, B, H, Lq, a.nkb);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

// ws_dq: bf16 [nkb][B*Lq][H*64]; ws_dk, ws_dv: bf16 [nqb][B*Lk][H*64]
extern "C" int mma_attn_bwd_t5b(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                                const unsigned char* kmask, const void* o, long long ldo, const float* lse, const void* dout,
                                long long lddo, void* dq, long long lddq, void* dk, long long lddk, void* dv, long long lddv,
                                void* ws_dq, void* ws_dk, void* ws_dv, int B, int H, int Lq, int Lk, int causal,
                                float scale, float p_drop, unsigned long long seed, unsigned int site, cudaStream_t stream) {
  using namespace at5;
  if (B <= 0 || Lq <= 0 || Lk <= 0) return MMA_OK;
  if (Lq > 512 || Lk > 512 || ((ldq | ldk | ldv | ldo | lddo | lddq | lddk | lddv) & 7) || !ws_dq || !ws_dk || !ws_dv)
    return MMA_ERR_UNSUPPORTED;
  CUtensorMap tq, tk, tv, tdo, to;
  int rc;
  if ((rc = map64(&tq, q, H, (long long)B * Lq, ldq))) return rc;
  if ((rc = map64(&tk, k, H, (long long)B * Lk, ldk))) return rc;
  if ((rc = map64(&tv, v, H, (long long)B * Lk, ldv))) return rc;
  if ((rc = map64(&tdo, dout, H, (long long)B * Lq, lddo))) return rc;
  if ((rc = map64(&to, o, H, (long long)B * Lq, ldo))) return rc;
  BlkArgs a{};
  a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk; a.causal = causal; a.nqb = (Lq + 127) / 128; a.nkb = (Lk + 127) / 128;
  a.scale = scale; a.p_drop = p_drop; a.seed = seed; a.site = site; a.kmask = kmask; a.lse = lse;
  a.dqpart = (bf16*)ws_dq; a.dkpart = (bf16*)ws_dk; a.dvpart = (bf16*)ws_dv;
  const int smem = 7 * TILE + 64;
  static bool set = false;
  if (!set) { cudaFuncSetAttribute(bwd_blk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); set = true; }
  const long long nprob = (long long)B * H * a.nqb * a.nkb;
  if (launch_pdl(bwd_blk_kernel, dim3((unsigned)nprob), dim3(128), smem, stream, tq, tk, tv, tdo, to, a) != cudaSuccess)
    return MMA_ERR_LAUNCH;
  merge_bwd_kernel<<<148 * 8, 256, 0, stream>>>((const bf16*)ws_dq, (const bf16*)ws_dk, (const bf16*)ws_dv, (bf16*)dq, lddq, (bf16*)dk, lddk, (bf16*)dv, lddv,
                                                (long long)B * Lq, (long long)B * Lk, H * 64, a.nkb, a.nqb);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}
