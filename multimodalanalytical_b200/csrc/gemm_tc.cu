// tcgen05 / TMEM / TMA bf16 GEMM for sm_100a.
//
//   C[M,N] = epilogue( A_op[M,K] * B_op[N,K]^T )      fp32 accumulate in TMEM
//
// Each operand is either K-major (memory [rows, K] row-major: activations, nn.Linear weights) or
// MN-major (memory [K, rows] row-major).  That covers the three GEMMs of a linear layer with no
// transposed copies:   fwd  y = x W^T        (A K-major, B K-major)
//                      dgrad dx = dy W       (A K-major, B MN-major)
//                      wgrad dW = dy^T x     (A MN-major, B MN-major; split over the row dimension)
//
// Structure: persistent CTAs (one per SM), 192 threads = TMA producer warp, MMA issuer warp (one
// elected thread issues tcgen05.mma), four epilogue warps.  STAGES-deep smem ring fed by TMA
// (128B swizzle), two TMEM accumulator stages so the epilogue of tile i overlaps the mainloop of
// tile i+1.  The epilogue (bias / GELU / GLU / residual / dropout / split-K accumulate) is shared
// with the SIMT fp32 GEMM (common.cuh).
#include <cstdlib>

#include "common.cuh"
#include "tma.cuh"

namespace tc {
using namespace tma;

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int NUM_EPI_WARPS = 16;  // four warps per TMEM lane quarter, each drains a quarter of the tile's columns
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;

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
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory"); }

// LayerNorm fused behind the residual epilogue (LNF variant of gemm_tc_kernel, N = 512 = four 128-column tiles of a 4-CTA
// cluster): the rows' statistics meet in every CTA's shared memory through st.shared::cluster
struct LnArgs {
  const float* gamma;
  const float* beta;
  float eps;
  bf16* h;
  long long ldh;
};
constexpr int LNF_N = 512, LNF_CS = 4, LNF_PARTS = LNF_CS * 4;  // 16 partial statistics per row (32 columns each)
constexpr uint32_t LNF_EXTRA_SMEM = 2 * 128 * 4 + BM * LNF_PARTS * 8;
__device__ __forceinline__ uint32_t cl_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cl_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f32x2(uint32_t local_addr, uint32_t rank, float a, float b) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(r), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
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

// Shared-memory matrix descriptor, 128B swizzle (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2 (SW128)
template <bool MN_MAJOR>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t tile_addr, int k16) {
  uint32_t addr, lbo, sbo;
  if (!MN_MAJOR) {
    // rows of 128 B (64 bf16 of K), 8-row swizzle atoms 1024 B apart; K advances 32 B inside the row
    addr = tile_addr + (uint32_t)k16 * 32u;
    lbo = 16;
    sbo = 1024;
  } else {
    // 64-element MN atoms: BK rows (k) of 128 B each; atoms BK*128 B apart; 8-k groups 1024 B apart
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


// ---- coalesced epilogue stores -------------------------------------------------------------------------------
// tcgen05.ld hands every thread one ROW of the accumulator; storing that directly makes each warp-wide store touch
// 32 different 128-byte lines with 16 bytes each.  Instead each epilogue warp parks its rows in a private,
// XOR-swizzled [32 rows][128 B] shared-memory tile and writes it out with 8 lanes per row: full, contiguous lines.
//   mode 0: one bf16 output, 64 columns per staged row     mode 1: one fp32 output, 32 columns per staged row
//   mode 2: two bf16 outputs, 32 + 32 columns per staged row
__device__ __forceinline__ uint4 pack8_bf16(const float* x) {
  __nv_bfloat162 a = __floats2bfloat162_rn(x[0], x[1]), b = __floats2bfloat162_rn(x[2], x[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(x[4], x[5]), d = __floats2bfloat162_rn(x[6], x[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  return u;
}
__device__ __forceinline__ void stage_put(uint8_t* st, int lane, int chunk, uint4 v) {
  *reinterpret_cast<uint4*>(st + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = v;
}
__device__ __forceinline__ void stage_chunk(uint8_t* st, int lane, int mode, int cs, const float (&o1)[16],
                                            const float (&o2)[16]) {
  if (mode == 0) {
    stage_put(st, lane, cs * 2, pack8_bf16(o1));
    stage_put(st, lane, cs * 2 + 1, pack8_bf16(o1 + 8));
  } else if (mode == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      stage_put(st, lane, cs * 4 + i, make_uint4(__float_as_uint(o1[4 * i]), __float_as_uint(o1[4 * i + 1]),
                                                 __float_as_uint(o1[4 * i + 2]), __float_as_uint(o1[4 * i + 3])));
  } else {
    stage_put(st, lane, cs * 2, pack8_bf16(o1));
    stage_put(st, lane, cs * 2 + 1, pack8_bf16(o1 + 8));
    stage_put(st, lane, 4 + cs * 2, pack8_bf16(o2));
    stage_put(st, lane, 4 + cs * 2 + 1, pack8_bf16(o2 + 8));
  }
}
// write the staged rows out: row0 = first tile row of this warp, colbase = first staged column, ncols = columns staged
__device__ __forceinline__ void stage_flush(const uint8_t* st, int lane, int mode, const Epi& ep, long long row0,
                                            int colbase, int ncols, int M, int N) {
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int r = p * 4 + (lane >> 3), ch = lane & 7;
    const long long row = row0 + r;
    const uint4 val = *reinterpret_cast<const uint4*>(st + r * 128 + ((ch ^ (r & 7)) << 4));
    if (row >= M) continue;
    if (mode == 1) {
      const int c = ch * 4;
      if (c >= ncols) continue;
      const int col = colbase + c;
      float* dst = reinterpret_cast<float*>(ep.out) + row * ep.ldo + col;
      if (col + 4 <= N && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        *reinterpret_cast<uint4*>(dst) = val;
      } else {
        const uint32_t w[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (col + j < N) dst[j] = __uint_as_float(w[j]);
      }
    } else {
      const int which = mode == 2 ? (ch >> 2) : 0;
      const int c = (mode == 2 ? (ch & 3) : ch) * 8;
      if (c >= ncols) continue;
      const int col = colbase + c;
      bf16* dst = reinterpret_cast<bf16*>(which ? ep.out2 : ep.out) + row * (which ? ep.ldo2 : ep.ldo) + col;
      if (col + 8 <= N && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        *reinterpret_cast<uint4*>(dst) = val;
      } else {
        const bf16* e = reinterpret_cast<const bf16*>(&val);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (col + j < N) dst[j] = e[j];
      }
    }
  }
}

// Staged (coalesced) epilogue stores cost one pipeline stage of shared memory.  Measured on B200 (scripts/gemm_bench.py):
// they help the plain-store K=512 products (58 -> 49 us) but lose on the arithmetic-heavy epilogues and on every
// large-K product (one stage less), and the C2 training step as a whole got slower - so they are off.
constexpr bool kStagedStores = false;

template <int BN>
struct Cfg {
  static constexpr int STAGES = kStagedStores ? (BN <= 128 ? 4 : 3) : (BN <= 128 ? 6 : 4);
  static_assert(BN == 128 || BN == 256, "tile N must be 128 or 256");
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = BN * BK * 2;
  static constexpr uint32_t STAGE_BYTES = 32 * 128;  // per epilogue warp: [32 rows][128 B]
  static constexpr uint32_t SMEM = STAGES * (A_BYTES + B_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/ +
                                   4 * BN /*bias*/ + (kStagedStores ? NUM_EPI_WARPS * STAGE_BYTES : 0);
  static constexpr uint32_t TMEM_COLS = 2 * BN;  // power of two for BN in {64,128,256}
};

template <int BN, bool A_MN, bool B_MN, int KIND, bool LNF = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
               int splits_dbg, Epi ep, LnArgs ln) {
  pdl_trigger();
  // diagnostics (MMA_GEMM_DBG, scripts/gemm_bench.py): bit0 epilogue skips its global loads / stores, bit1 no TMA
  // loads and no MMAs (epilogue alone), bit2 TMA loads but no MMAs.  Zero in production.
  const int splits = splits_dbg & 0xFFFF;
  const int dbg = splits_dbg >> 16;
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + STAGES * C::B_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* sBias = reinterpret_cast<float*>(bars) + 64;  // [BN], 256 bytes past the barrier block start
  uint8_t* sStage = reinterpret_cast<uint8_t*>(sBias + BN);  // [NUM_EPI_WARPS][32 rows][128 B]
  float* sGam = sBias + BN;                                  // LNF: gamma / beta of this tile's 128 columns,
  float* sBet = sGam + BN;
  float2* sStats = reinterpret_cast<float2*>(sBet + BN);     //      [BM rows][16 partial (mean, M2)]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(smem_u32(&full[i]), 1);
      mbar_init(smem_u32(&empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tfull[i]), 1);
      mbar_init(smem_u32(&tempty[i]), NUM_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; its results are visible from here on

  const int tiles_m = (M + BM - 1) / BM;
  const int tiles_n = (N + BN - 1) / BN;
  const int num_kb = (K + BK - 1) / BK;
  const int total = tiles_m * tiles_n * splits;

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer =================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total && !(dbg & 2); tile += gridDim.x) {
        const int split = tile % splits;
        const int t2 = tile / splits;
        const int n0 = (t2 % tiles_n) * BN;
        const int m0 = (t2 / tiles_n) * BM;
        const int kb0 = (int)((long long)split * num_kb / splits);
        const int kb1 = (int)((long long)(split + 1) * num_kb / splits);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full[stage]);
          mbar_expect_tx(fb, C::A_BYTES + C::B_BYTES);
          const uint32_t a_dst = smem_u32(sA + stage * C::A_BYTES);
          const uint32_t b_dst = smem_u32(sB + stage * C::B_BYTES);
          if (!A_MN) {
            tma_load_2d(a_dst, &tmA, fb, kb * BK, m0);
          } else {
#pragma unroll
            for (int a = 0; a < BM / 64; ++a) tma_load_2d(a_dst + a * (BK * 128), &tmA, fb, m0 + a * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(b_dst, &tmB, fb, kb * BK, n0);
          } else {
#pragma unroll
            for (int a = 0; a < BN / 64; ++a) tma_load_2d(b_dst + a * (BK * 128), &tmB, fb, n0 + a * 64, kb * BK);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    if constexpr (LNF) {  // every thread of the cluster meets at the statistics exchange
      __syncwarp();
      cl_sync_all();
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================= MMA issuer (single thread) =================
      // instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6), a=bf16 [7,10), b=bf16 [10,13),
      // a_major bit15, b_major bit16, N>>3 [17,23), M>>4 [24,29)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                                 ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int split = tile % splits;
        const int kb0 = (int)((long long)split * num_kb / splits);
        const int kb1 = (int)((long long)(split + 1) * num_kb / splits);
        mbar_wait(smem_u32(&tempty[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1 && !(dbg & 2); ++kb) {
          mbar_wait(smem_u32(&full[stage]), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * C::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if (!(dbg & 4)) tc_mma_bf16(d_tmem, make_smem_desc<A_MN>(a_addr, k), make_smem_desc<B_MN>(b_addr, k), idesc,
                        (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit(smem_u32(&empty[stage]));  // smem slot is free once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(smem_u32(&tfull[acc]));  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
    if constexpr (LNF) {
      __syncwarp();
      cl_sync_all();
    }
  } else {
    // ================= epilogue warps (TMEM -> registers -> global) =================
    const int q = warp & 3;           // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int half = (warp - 2) >> 2;  // which slice of the tile's columns this warp drains
    constexpr int COLS = BN / (NUM_EPI_WARPS / 4);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const int t2 = tile / splits;
      const int n0 = (t2 % tiles_n) * BN;
      const int m0 = (t2 / tiles_n) * BM;
      const long long row = (long long)m0 + q * 32 + lane;
      const bool row_ok = row < M && !(dbg & 1);
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + half * COLS);
      const int col0 = n0 + half * COLS;
      if constexpr (LNF) {
        // out = resid + acc + bias (fp32, stored) and h = LayerNorm(out) (bf16): one tile per CTA, the four CTAs of the
        // cluster hold the four 128-column tiles of the same 128 rows.  Each thread keeps its 32 columns of its row in
        // registers, publishes (mean, M2) of them to all four CTAs, and after the cluster barrier merges the row's 16
        // partials (Chan's formula: equal counts) - one exchange, no cancellation.
        static_assert(!LNF || (BN == 128 && KIND == EPI_RESID), "LNF: 128-column tiles, residual epilogue");
        epi_bar_sync();
        const int et = threadIdx.x - 64;
        if (et < BN) {
          sBias[et] = ep.bias ? ep.bias[n0 + et] : 0.f;
          sGam[et] = ln.gamma[n0 + et];
          sBet[et] = ln.beta[n0 + et];
        }
        EpiIn<16, KIND> in[2];
        epi_prefetch<16, KIND>(ep, row, col0, N, row_ok, in[0]);
        epi_prefetch<16, KIND>(ep, row, col0 + 16, N, row_ok, in[1]);
        epi_bar_sync();
        if (lane == 0) mbar_wait(smem_u32(&tfull[acc]), acc_phase);
        __syncwarp();
        tc_fence_after();
        float xv[2][16];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tmem_ld16(t_row + (uint32_t)(c * 16), xv[c]);
#pragma unroll
          for (int j = 0; j < 16; ++j) xv[c][j] = in[c].a[j] + xv[c][j] + sBias[half * COLS + c * 16 + j];
          if (row_ok) {
            float* o = reinterpret_cast<float*>(ep.out) + row * ep.ldo + col0 + c * 16;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              *reinterpret_cast<float4*>(o + 4 * k) = make_float4(xv[c][4 * k], xv[c][4 * k + 1], xv[c][4 * k + 2], xv[c][4 * k + 3]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty[acc]));
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int j = 0; j < 16; ++j) sum += xv[c][j];
        const float m_i = sum * (1.0f / 32.0f);
        float m2_i = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int j = 0; j < 16; ++j) m2_i = fmaf(xv[c][j] - m_i, xv[c][j] - m_i, m2_i);
        const int rloc = q * 32 + lane;
        const uint32_t slot = smem_u32(&sStats[rloc * LNF_PARTS + (int)cl_ctarank() * 4 + half]);
#pragma unroll
        for (uint32_t rk = 0; rk < (uint32_t)LNF_CS; ++rk) st_cluster_f32x2(slot, rk, m_i, m2_i);
        __syncwarp();
        cl_sync_all();
        float mean = 0.f;
#pragma unroll
        for (int i = 0; i < LNF_PARTS; ++i) mean += sStats[rloc * LNF_PARTS + i].x;
        mean *= 1.0f / LNF_PARTS;
        float m2 = 0.f;
#pragma unroll
        for (int i = 0; i < LNF_PARTS; ++i) {
          const float2 pt = sStats[rloc * LNF_PARTS + i];
          m2 += pt.y + 32.0f * (pt.x - mean) * (pt.x - mean);
        }
        const float rstd = rsqrtf(m2 * (1.0f / LNF_N) + ln.eps);
        if (row_ok) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int cc = half * COLS + c * 16 + 2 * j;
              const float y0 = (xv[c][2 * j] - mean) * rstd * sGam[cc] + sBet[cc];
              const float y1 = (xv[c][2 * j + 1] - mean) * rstd * sGam[cc + 1] + sBet[cc + 1];
              __nv_bfloat162 p2 = __floats2bfloat162_rn(y0, y1);
              pk[j] = *reinterpret_cast<uint32_t*>(&p2);
            }
            bf16* hp = ln.h + row * ln.ldh + col0 + c * 16;
            *reinterpret_cast<uint4*>(hp) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(hp + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
        continue;
      }
      if (KIND >= 0) {
        // software-pipelined epilogue: bias staged in smem and the first chunk's side inputs fetched while the MMAs
        // of this tile are still running; chunk c+1's TMEM read and global loads overlap chunk c's math and stores
        constexpr int NCH = COLS / 16;
        const bool use_bias = ep.bias != nullptr && KIND != EPI_ACCUM && KIND != EPI_DGELU && KIND != EPI_DGLU && KIND != EPI_DRELU;
        epi_bar_sync();  // every epilogue warp is done with the previous tile's bias
        const int et = threadIdx.x - 64;
        if (use_bias && et < BN) sBias[et] = (n0 + et < N) ? ep.bias[n0 + et] : 0.f;
        EpiIn<16, KIND> in[2];
        epi_prefetch<16, KIND>(ep, row, col0, N, row_ok, in[0]);
        epi_bar_sync();
        if (lane == 0) mbar_wait(smem_u32(&tfull[acc]), acc_phase);  // one poller per warp
        __syncwarp();
        tc_fence_after();
        // store mode: 0 one bf16 output, 1 one fp32 output, 2 two bf16 outputs, 3 direct (atomics / fp32 pairs)
        const bool dual = (KIND == EPI_DGLU) || ((KIND == EPI_GELU || KIND == EPI_GLU_MUL) && ep.out2 != nullptr);
        const bool f32out = KIND == EPI_ACCUM ? true : ep.out_f32 != 0;
        const int mode = (!kStagedStores || (KIND == EPI_ACCUM && ep.accumulate == 2)) ? 3
                         : dual ? (f32out ? 3 : 2) : (f32out ? 1 : 0);
        const int group = mode == 0 ? (COLS / 16 < 4 ? COLS / 16 : 4) : 2;  // chunks per staged row
        uint8_t* st = sStage + (warp - 2) * C::STAGE_BYTES;
        const long long row0 = (long long)m0 + q * 32;
        uint32_t raw[2][16];
        tmem_ld16_nowait(t_row, raw[0]);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          tmem_wait_ld();
          if (c + 1 < NCH) {
            tmem_ld16_nowait(t_row + (uint32_t)((c + 1) * 16), raw[(c + 1) & 1]);
            epi_prefetch<16, KIND>(ep, row, col0 + (c + 1) * 16, N, row_ok, in[(c + 1) & 1]);
          }
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[c & 1][j]);
          const float* bs = use_bias ? sBias + half * COLS + c * 16 : nullptr;
          if (mode == 3) {
            epi_finish<16, KIND, true>(ep, row, col0 + c * 16, N, row_ok, v, in[c & 1], bs);
          } else {
            float o2[16];
            epi_math<16, KIND, true>(ep, row, col0 + c * 16, v, o2, in[c & 1], bs);
            const int cs = c % group;
            stage_chunk(st, lane, mode, cs, v, o2);
            if (cs == group - 1) {
              __syncwarp();
              stage_flush(st, lane, mode, ep, row0, col0 + (c - cs) * 16, group * 16, M, N);
              __syncwarp();
            }
          }
        }
      } else {
        if (lane == 0) mbar_wait(smem_u32(&tfull[acc]), acc_phase);
        __syncwarp();
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < COLS; c0 += 16) {
          float v[16];
          tmem_ld16(t_row + (uint32_t)c0, v);
          const int col = col0 + c0;
          if (row_ok && col < N) epilogue_store<16, KIND, true>(ep, row, col, N, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty[acc]));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side: launch (tensor maps: tma.cuh)
// ------------------------------------------------------------------------------------------------
static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

template <int BN, bool A_MN, bool B_MN, int KIND>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, int splits, const Epi& ep,
                  int max_ctas, cudaStream_t stream) {
  using C = Cfg<BN>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, KIND>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) != cudaSuccess)
      return MMA_ERR_LAUNCH;
    attr_set = true;
  }
  const int total = ((M + BM - 1) / BM) * ((N + BN - 1) / BN) * splits;
  int grid = total < num_sms() ? total : num_sms();
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("MMA_GEMM_DBG");
    dbg = e ? atoi(e) : 0;
  }
  if (launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), C::SMEM, stream, tmA, tmB, M, N, K, splits | (dbg << 16), ep, LnArgs{}) != cudaSuccess)
    return MMA_ERR_LAUNCH;
  return MMA_OK;
}

// out = resid + A W^T + bias (fp32) and h = LayerNorm(out) (bf16) for N = 512 and a few thousand rows: one 128 x 128 tile
// per CTA, clusters of four CTAs along N (every tile of the launch resident or in later waves - clusters are independent).
// The decode step's three residual products per layer at 640 ... 4736 rows, where the CTA-pair fused kernel
// (gemm2_ln_kernel: 256 rows x all 512 columns per pair) leaves most of the machine idle.
int gemm_resid_ln_c4(const void* A, long long lda, const void* W, long long ldw, int M, int K, const Epi& ep,
                     const float* gamma, const float* beta, float eps, void* h, long long ldh, cudaStream_t stream) {
  using C = Cfg<128>;
  constexpr uint32_t SMEM = C::SMEM + LNF_EXTRA_SMEM;
  static_assert(SMEM <= 232448, "shared memory budget");
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, A, (unsigned long long)K, (unsigned long long)M, lda, BK, BM);
  if (rc) return rc;
  if ((rc = make_map(&tmB, W, (unsigned long long)K, (unsigned long long)LNF_N, ldw, BK, 128))) return rc;
  auto kern = gemm_tc_kernel<128, false, false, EPI_RESID, true>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return MMA_ERR_LAUNCH;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(((M + BM - 1) / BM) * LNF_CS));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = LNF_CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // prologue overlaps the previous kernel's tail
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static int pdl = -1;
  if (pdl < 0) {
    const char* e = getenv("MMA_RESID_LN_C4_PDL");
    pdl = e ? atoi(e) : 1;
  }
  cfg.numAttrs = pdl ? 2 : 1;
  LnArgs ln{gamma, beta, eps, reinterpret_cast<bf16*>(h), ldh};
  if (cudaLaunchKernelEx(&cfg, kern, tmA, tmB, M, LNF_N, K, 1, ep, ln) != cudaSuccess) {
    cudaGetLastError();
    return MMA_ERR_LAUNCH;
  }
  return MMA_OK;
}

template <int BN>
static int dispatch(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, int M, int N, int K,
                    const Epi& ep, int splits, int max_ctas, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = make_map(&tmA, A, (unsigned long long)K, (unsigned long long)M, lda, BK, BM);
  else rc = make_map(&tmA, A, (unsigned long long)M, (unsigned long long)K, lda, 64, BK);
  if (rc) return rc;
  if (!b_mn) rc = make_map(&tmB, B, (unsigned long long)K, (unsigned long long)N, ldb, BK, BN);
  else rc = make_map(&tmB, B, (unsigned long long)N, (unsigned long long)K, ldb, 64, BK);
  if (rc) return rc;
#define MMA_LAUNCH(AM, BM_, KD) return launch<BN, AM, BM_, KD>(tmA, tmB, M, N, K, splits, ep, max_ctas, stream)
  if (!a_mn && !b_mn) {  // forward products
    switch (ep.kind) {
      case EPI_STORE: MMA_LAUNCH(false, false, EPI_STORE);
      case EPI_GELU: MMA_LAUNCH(false, false, EPI_GELU);
      case EPI_RESID: MMA_LAUNCH(false, false, EPI_RESID);
      case EPI_GLU_MUL: MMA_LAUNCH(false, false, EPI_GLU_MUL);
      default: MMA_LAUNCH(false, false, -1);
    }
  }
  if (!a_mn && b_mn) {  // dgrad
    switch (ep.kind) {
      case EPI_STORE: MMA_LAUNCH(false, true, EPI_STORE);
      case EPI_DGELU: MMA_LAUNCH(false, true, EPI_DGELU);
      case EPI_DGLU: MMA_LAUNCH(false, true, EPI_DGLU);
      case EPI_ACCUM: MMA_LAUNCH(false, true, EPI_ACCUM);
      default: MMA_LAUNCH(false, true, -1);
    }
  }
  if (a_mn && b_mn) {  // wgrad
    if (ep.kind == EPI_ACCUM) MMA_LAUNCH(true, true, EPI_ACCUM);
    MMA_LAUNCH(true, true, -1);
  }
  MMA_LAUNCH(true, false, -1);
#undef MMA_LAUNCH
}

// ------------------------------------------------------------------------------------------------
// Grouped weight-gradient kernel: up to 8 independent  dW_g[Nout_g, Kin_g] += dy_g^T x_g  products (one backward
// layer's worth) in ONE persistent launch, so that the ~100-130 output tiles of a layer fill the machine without
// split-K / atomics.  Tiles are 128 x 256, both operands MN-major (rows of dy / x are the reduction dimension).
// The bias gradient db_g[n] = sum_r dy_g[r, n] rides along for free on the first tile column of every tile row: one
// extra N=16 MMA per k-step multiplies the same dy tile with a constant all-ones B tile into 32 spare TMEM columns.
// ------------------------------------------------------------------------------------------------
struct alignas(64) WgProblem {
  CUtensorMap tmA;  // dy [R, Nout]: dims {Nout, R}, box {64, 64}
  CUtensorMap tmB;  // x  [R, Kin] : dims {Kin, R},  box {64, 64}
  float* out;       // [Nout, Kin] fp32, accumulated (+=)
  float* dbias;     // [Nout] fp32, accumulated (+=), may be null
  long long ldo;
  int M, N, R;      // Nout, Kin, rows
  int tiles_n, tile_begin;
};
struct WgGroup {
  WgProblem p[8];
  int count, total_tiles;
};

constexpr int WG_BN = 256;
constexpr int WG_STAGES = 4;
constexpr uint32_t WG_A_BYTES = BM * BK * 2, WG_B_BYTES = WG_BN * BK * 2;
constexpr uint32_t WG_ONES_BYTES = 16 * 128;
constexpr uint32_t WG_SMEM = WG_STAGES * (WG_A_BYTES + WG_B_BYTES) + WG_ONES_BYTES + 1024 + 256;

__global__ void __launch_bounds__(NUM_THREADS, 1) wgrad_group_kernel(const __grid_constant__ WgGroup grp) {
  pdl_trigger();
  constexpr int STAGES = WG_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * WG_A_BYTES;
  uint8_t* sOnes = sB + STAGES * WG_B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + WG_ONES_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < (int)(WG_ONES_BYTES / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;  // bf16 1.0 pairs (swizzle-invariant: every element equal)
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(smem_u32(&full[i]), 1);
      mbar_init(smem_u32(&empty[i]), 1);
    }
    mbar_init(smem_u32(tfull), 1);
    mbar_init(smem_u32(tempty), NUM_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the ones tile is read by the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  auto locate = [&](int tile, int& g, int& m0, int& n0) {
    g = 0;
#pragma unroll 1
    for (int i = 1; i < grp.count; ++i)
      if (tile >= grp.p[i].tile_begin) g = i;
    const int t = tile - grp.p[g].tile_begin;
    n0 = (t % grp.p[g].tiles_n) * WG_BN;
    m0 = (t / grp.p[g].tiles_n) * BM;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < grp.total_tiles; tile += gridDim.x) {
        int g, m0, n0;
        locate(tile, g, m0, n0);
        const WgProblem& P = grp.p[g];
        const int num_kb = (P.R + BK - 1) / BK;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full[stage]);
          mbar_expect_tx(fb, WG_A_BYTES + WG_B_BYTES);
          const uint32_t a_dst = smem_u32(sA + stage * WG_A_BYTES);
          const uint32_t b_dst = smem_u32(sB + stage * WG_B_BYTES);
#pragma unroll
          for (int a = 0; a < BM / 64; ++a) tma_load_2d(a_dst + a * (BK * 128), &P.tmA, fb, m0 + a * 64, kb * BK);
#pragma unroll
          for (int a = 0; a < WG_BN / 64; ++a) tma_load_2d(b_dst + a * (BK * 128), &P.tmB, fb, n0 + a * 64, kb * BK);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_main = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                      ((uint32_t)(WG_BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      constexpr uint32_t idesc_bias = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (0u << 16) |
                                      ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, tphase = 0;
      for (int tile = blockIdx.x; tile < grp.total_tiles; tile += gridDim.x) {
        int g, m0, n0;
        locate(tile, g, m0, n0);
        const WgProblem& P = grp.p[g];
        const int num_kb = (P.R + BK - 1) / BK;
        const bool with_bias = n0 == 0 && P.dbias != nullptr;
        mbar_wait(smem_u32(tempty), tphase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&full[stage]), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * WG_A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * WG_B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = make_smem_desc<true>(a_addr, k);
            tc_mma_bf16(tmem_base, ad, make_smem_desc<true>(b_addr, k), idesc_main, (kb > 0 || k > 0) ? 1u : 0u);
            if (with_bias)
              tc_mma_bf16(tmem_base + WG_BN, ad, make_smem_desc<false>(smem_u32(sOnes), k), idesc_bias,
                          (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(smem_u32(&empty[stage]));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(smem_u32(tfull));
        tphase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    const int slice = (warp - 2) >> 2;
    constexpr int COLS = WG_BN / (NUM_EPI_WARPS / 4);
    uint32_t tphase = 0;
    for (int tile = blockIdx.x; tile < grp.total_tiles; tile += gridDim.x) {
      int g, m0, n0;
      locate(tile, g, m0, n0);
      const WgProblem& P = grp.p[g];
      Epi ep{};
      ep.kind = EPI_ACCUM;
      ep.out_f32 = 1;
      ep.out = P.out;
      ep.ldo = P.ldo;
      ep.alpha = 1.0f;
      ep.accumulate = 1;
      mbar_wait(smem_u32(tfull), tphase);
      tc_fence_after();
      const long long row = (long long)m0 + q * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < COLS; c0 += 16) {
        float v[16];
        tmem_ld16(t_row + (uint32_t)(slice * COLS + c0), v);
        const int col = n0 + slice * COLS + c0;
        if (row < P.M && col < P.N) epilogue_store<16, EPI_ACCUM, true>(ep, row, col, P.N, v);
      }
      if (slice == 0 && n0 == 0 && P.dbias != nullptr) {
        float v[16];
        tmem_ld16(t_row + (uint32_t)WG_BN, v);
        if (row < P.M) P.dbias[row] += v[0];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(tempty));
      tphase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace tc

// Grouped wgrad (+ bias grad): for g < count:  out[g][Nout, Kin] += dy[g]^T x[g];  dbias[g][Nout] += colsum(dy[g]).
// dy[g]: bf16 [R, Nout] pitch lddy;  x[g]: bf16 [R, Kin] pitch ldx.  Every output tile has one writer: no atomics.
extern "C" int mma_wgrad2_group(int count, const void* const* dy, const long long* lddy, const void* const* x,
                                const long long* ldx, float* const* out, const long long* ldo, float* const* dbias,
                                const int* Nout, const int* Kin, const int* R, cudaStream_t stream);

extern "C" int mma_wgrad_group(int count, const void* const* dy, const long long* lddy, const void* const* x,
                               const long long* ldx, float* const* out, const long long* ldo, float* const* dbias,
                               const int* Nout, const int* Kin, const int* R, cudaStream_t stream) {
  using namespace tc;
  if (count < 1 || count > 8) return MMA_ERR_ARG;
  {
    // CTA-pair kernel (gemm_tc2.cu) unless MMA_WGRAD2=0; it needs 16-byte aligned fp32 gradient rows
    static int use2 = -1;
    if (use2 < 0) {
      const char* e = getenv("MMA_WGRAD2");
      use2 = e ? atoi(e) : 1;
    }
    bool ok = use2 != 0;
    for (int g = 0; g < count && ok; ++g)
      ok = (reinterpret_cast<uintptr_t>(out[g]) & 15) == 0 && ((ldo[g] * 4) & 15) == 0;
    if (ok) return mma_wgrad2_group(count, dy, lddy, x, ldx, out, ldo, dbias, Nout, Kin, R, stream);
  }
  WgGroup grp{};
  int tiles = 0;
  for (int g = 0; g < count; ++g) {
    WgProblem& P = grp.p[g];
    if (Nout[g] <= 0 || Kin[g] <= 0 || R[g] <= 0) return MMA_ERR_ARG;
    int rc = make_map(&P.tmA, dy[g], (unsigned long long)Nout[g], (unsigned long long)R[g], lddy[g], 64, BK);
    if (rc) return rc;
    rc = make_map(&P.tmB, x[g], (unsigned long long)Kin[g], (unsigned long long)R[g], ldx[g], 64, BK);
    if (rc) return rc;
    P.out = out[g];
    P.dbias = dbias ? dbias[g] : nullptr;
    P.ldo = ldo[g];
    P.M = Nout[g];
    P.N = Kin[g];
    P.R = R[g];
    P.tiles_n = (Kin[g] + WG_BN - 1) / WG_BN;
    P.tile_begin = tiles;
    tiles += ((Nout[g] + BM - 1) / BM) * P.tiles_n;
  }
  grp.count = count;
  grp.total_tiles = tiles;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(wgrad_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM) != cudaSuccess)
      return MMA_ERR_LAUNCH;
    attr_set = true;
  }
  const int grid = tiles < num_sms() ? tiles : num_sms();
  if (launch_pdl(wgrad_group_kernel, dim3(grid), dim3(NUM_THREADS), WG_SMEM, stream, grp) != cudaSuccess)
    return MMA_ERR_LAUNCH;
  return MMA_OK;
}

extern "C" int mma_gemm2_eligible(int a_mn, int b_mn, int M, int N, int K, const Epi* ep, int splits);
extern "C" int mma_gemm2_bf16(const void* A, long long lda, const void* B, long long ldb, int b_mn, int M, int N, int K,
                              const Epi* ep, cudaStream_t stream);

// C[M,N] = epi(A_op * B_op^T).  a_mn / b_mn: 0 = operand memory is [rows, K] (K-major), 1 = [K, rows].
// lda / ldb: row pitch of the operand's memory in elements.  splits > 1 splits the K loop across CTAs
// (requires ep->kind == EPI_ACCUM with accumulate == 2).  max_ctas <= 0: one CTA per SM.
extern "C" int mma_gemm_bf16(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, int M,
                             int N, int K, const Epi* ep, int splits, int max_ctas, cudaStream_t stream) {
  using namespace tc;
  if (M <= 0 || N <= 0 || K <= 0 || !ep) return MMA_ERR_ARG;
  const int num_kb = (K + BK - 1) / BK;
  if (splits < 1) splits = 1;
  if (splits > num_kb) splits = num_kb;
  if (splits > 1 && !(ep->kind == EPI_ACCUM && ep->accumulate == 2)) return MMA_ERR_ARG;
  // large products: CTA-pair kernel (256 x 256 tiles, TMA epilogue), gemm_tc2.cu
  if (max_ctas <= 0 && mma_gemm2_eligible(a_mn, b_mn, M, N, K, ep, splits))
    return mma_gemm2_bf16(A, lda, B, ldb, b_mn, M, N, K, ep, stream);
  // MMA_GEMM_BN=128|256 forces the tile width (experiments)
  static int forced_bn = -1;
  if (forced_bn < 0) {
    const char* e = getenv("MMA_GEMM_BN");
    forced_bn = e ? atoi(e) : 0;
  }
  if (forced_bn == 128) return dispatch<128>(A, lda, a_mn, B, ldb, b_mn, M, N, K, *ep, splits, max_ctas, stream);
  if (forced_bn == 256 && N >= 256) return dispatch<256>(A, lda, a_mn, B, ldb, b_mn, M, N, K, *ep, splits, max_ctas, stream);
  // 128x256 tiles halve the A re-reads from L2; use them when the 256-wide tiling still fills the machine
  const long long tiles256 = (long long)((M + BM - 1) / BM) * ((N + 255) / 256) * splits;
  if (N >= 256 && tiles256 >= 120) return dispatch<256>(A, lda, a_mn, B, ldb, b_mn, M, N, K, *ep, splits, max_ctas, stream);
  return dispatch<128>(A, lda, a_mn, B, ldb, b_mn, M, N, K, *ep, splits, max_ctas, stream);
}
