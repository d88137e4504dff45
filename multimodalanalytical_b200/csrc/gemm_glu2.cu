// Gated FFN (GLU) on the CTA-pair tcgen05 path: the two products of the reference's
//     FFN(h) = W2 drop( gelu(W1 h + b1) * (Wg h + bg) ) + b2          (custom_modeling.py:137-152,184-199)
// that carry a gate epilogue, forward and backward.
//
//   glu_fwd_kernel   a = drop(gelu(z1) * z2),  z1 = h W1^T + b1,  z2 = h Wg^T + bg      ONE launch for both products:
//                    a CTA pair owns 256 rows x 128 output columns; the leader CTA stages 128 rows of W1, its peer the
//                    matching 128 rows of Wg, so the M = 256 / N = 256 cta_group::2 MMA leaves z1 in accumulator columns
//                    [0, 128) and z2 in [128, 256).  h is read once, z1 never round-trips HBM as a side input, and the
//                    epilogue sees both pre-activations of an output element in the same thread.
//   dglu_kernel      da = (dy W2) * dropmask;  dz1 = da * z2 * gelu'(z1);  dz2 = da * gelu(z1)
//                    256 x 256 pair tiles as in gemm2_kernel; z1 / z2 arrive by TMA into a pair of shared boxes per
//                    32-column box of the output, the results overwrite them in place and leave by TMA stores.
//
// Roles, barriers and the TMEM accumulator double-buffer are those of gemm2_kernel (gemm_tc2.cu).
#include "tc2.cuh"

namespace tc2 {

struct GluArgs {
  const float* b1;
  const float* bg;
  float p_drop;
  unsigned long long seed;
  unsigned int site;
  long long drop_ld;
};

// keep-scales of 16 consecutive elements starting at element index e0 (same stream as common.cuh epilogue_store)
__device__ __forceinline__ void drop_scales16(const DropCtx& dc, uint32_t e0, float (&ds)[16]) {
  if ((e0 & 1u) == 0) {
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const uint32_t r = drop_pair(dc.key, (e0 + j) >> 1);
      ds[j] = (r & 0xFFFFu) >= dc.thr ? dc.inv_keep : 0.f;
      ds[j + 1] = (r >> 16) >= dc.thr ? dc.inv_keep : 0.f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) ds[j] = drop_scale1(dc.key, e0 + j, dc.thr, dc.inv_keep);
  }
}

__device__ __forceinline__ void st_box16(uint8_t* box, int lane, int half, int swz, const float (&v)[16]) {
#pragma unroll
  for (int k = 0; k < 2; ++k)
    *reinterpret_cast<uint4*>(box + box_off(lane, 2 * half + k, swz)) =
        make_uint4(pack2_bf16(v[8 * k], v[8 * k + 1]), pack2_bf16(v[8 * k + 2], v[8 * k + 3]),
                   pack2_bf16(v[8 * k + 4], v[8 * k + 5]), pack2_bf16(v[8 * k + 6], v[8 * k + 7]));
}
__device__ __forceinline__ void ld_box16(const uint8_t* box, int lane, int half, int swz, float (&v)[16]) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const uint4 u = *reinterpret_cast<const uint4*>(box + box_off(lane, 2 * half + k, swz));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float2 f = __bfloat1622float2(h[w]);
      v[8 * k + 2 * w] = f.x;
      v[8 * k + 2 * w + 1] = f.y;
    }
  }
}
// 16 consecutive fp32 values, the same for every lane of the warp (a broadcast read served by L1)
__device__ __forceinline__ void ld_bias16(const float* p, float (&b)[16]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(p) + k);
    b[4 * k] = f.x; b[4 * k + 1] = f.y; b[4 * k + 2] = f.z; b[4 * k + 3] = f.w;
  }
}

constexpr int GLU_BN = 128;  // output columns of a forward pair tile (z1 | z2 fill the 256 accumulator columns)

// ------------------------------------------------------------------------------------------------------------------
// forward: a (and, when TRAIN, the saved pre-activations z1, z2) = GLU(h)
// ------------------------------------------------------------------------------------------------------------------
template <bool TRAIN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
glu_fwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB1,
               const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmOut,
               const __grid_constant__ CUtensorMap tmZ1, const __grid_constant__ CUtensorMap tmZ2, int M, int N, int K,
               int swz, GluArgs ga) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* sEpi = sB + STAGES * B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + NUM_EPI_WARPS * EPI_WARP_BYTES);
  uint64_t* full = bars;                     // [STAGES]  (the leader's are used)
  uint64_t* empty = bars + STAGES;           // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;       // [2]
  uint64_t* tempty = bars + 2 * STAGES + 2;  // [2]       (the leader's are used)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(rank ? &tmB2 : &tmB1)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmOut)) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(smem_u32(&full[i]), 1);
      mbar_init(smem_u32(&empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tfull[i]), 1);
      mbar_init(smem_u32(&tempty[i]), 2 * NUM_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int tiles_n = (N + GLU_BN - 1) / GLU_BN;
  const int tiles_m = (M + 2 * BM - 1) / (2 * BM);
  const int total = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer: this CTA's 128 rows of h + 128 rows of W1 (leader) / Wg (peer) =================
      const CUtensorMap* tb = rank ? &tmB2 : &tmB1;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < total; tile += npairs) {
        const int m0 = (tile / tiles_n) * (2 * BM) + (int)rank * BM;
        const int n0 = (tile % tiles_n) * GLU_BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          MBAR_WAIT_LONG(smem_u32(&empty[stage]), phase ^ 1);
          const uint32_t fb_local = smem_u32(&full[stage]);
          if (rank == 0) mbar_expect_tx(fb_local, 2 * (A_BYTES + B_BYTES));
          const uint32_t fb = mapa(fb_local, 0);
          tma_load_2d_2sm(smem_u32(sA + stage * A_BYTES), &tmA, fb, kb * BK, m0);
          tma_load_2d_2sm(smem_u32(sB + stage * B_BYTES), tb, fb, kb * BK, n0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ================= MMA issuer =================
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)((2 * BM) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < total; tile += npairs) {
        MBAR_WAIT_LONG(smem_u32(&tempty[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&full[stage]), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc_mma2_bf16(d_tmem, make_smem_desc<false>(a_addr, k), make_smem_desc<false>(b_addr, k), idesc,
                         (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit2_mc(smem_u32(&empty[stage]));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit2_mc(smem_u32(&tfull[acc]));
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ================= epilogue: warp (q, s) owns rows [32 q, +32) x output columns [32 s, +32) of the CTA tile ========
    const int ew = warp - 2;
    const int q = warp & 3;
    const int s = ew >> 2;
    uint8_t* myb = sEpi + ew * EPI_WARP_BYTES;  // X0 = a, X1 = z1, X2 = z2
    const uint32_t myb_a = smem_u32(myb);
    const uint32_t tempty_leader0 = mapa(smem_u32(&tempty[0]), 0);
    DropCtx dc;
    dc.on = ga.p_drop > 0.0f;
    dc.thr = dc.on ? drop_threshold(ga.p_drop) : 0u;
    dc.inv_keep = dc.on ? 1.0f / (1.0f - ga.p_drop) : 1.0f;
    dc.key = dc.on ? drop_key(ga.seed, ga.site) : 0u;
    DropCtx dch = dc;  // keep-scales carrying the 1/2 of gelu_tanh_scaled
    if (MMA_GELU_TANH) dch.inv_keep *= 0.5f;

    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < total; tile += npairs) {
      const int r0 = (tile / tiles_n) * (2 * BM) + (int)rank * BM + q * 32;
      const int c0 = (tile % tiles_n) * GLU_BN + s * 32;
      const long long row = (long long)r0 + lane;
      const bool live = r0 < M && c0 < N;  // warp-uniform
      mbar_wait(smem_u32(&tfull[acc]), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + s * 32);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r1[16], r2[16];
        tmem_ld16_nowait(t_row + (uint32_t)(c * 16), r1);
        tmem_ld16_nowait(t_row + (uint32_t)(GLU_BN + c * 16), r2);
        tmem_wait_ld();
        if (c == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader0 + (uint32_t)acc * 8u);
        }
        const int col = c0 + c * 16;
        float z1[16], z2[16], av[16];
        if (live && col < N) {  // N % 16 == 0 (checked on the host): a live chunk is a full chunk
          ld_bias16(ga.b1 + col, z1);
          ld_bias16(ga.bg + col, z2);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) z1[j] = z2[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          z1[j] += __uint_as_float(r1[j]);
          z2[j] += __uint_as_float(r2[j]);
        }
        if (dc.on) {
          float ds[16];
          drop_scales16(dch, (uint32_t)((unsigned long long)row * (unsigned long long)ga.drop_ld + (unsigned long long)col), ds);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            av[j] = MMA_GELU_TANH ? gelu_tanh_scaled(z1[j], ds[j]) * z2[j] : gelu_t<true>(z1[j]) * z2[j] * ds[j];
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) av[j] = gelu_t<true>(z1[j]) * z2[j];
        }
        if (c == 0) {
          // the previous tile's stores must have finished reading the boxes (they had a whole mainloop to do so)
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
        }
        st_box16(myb, lane, c, swz, av);
        if (TRAIN) {
          st_box16(myb + BOX_BYTES, lane, c, swz, z1);
          st_box16(myb + 2 * BOX_BYTES, lane, c, swz, z2);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (live) {
          tma_store_2d(&tmOut, myb_a, c0, r0);
          if (TRAIN) {
            tma_store_2d(&tmZ1, myb_a + BOX_BYTES, c0, r0);
            tma_store_2d(&tmZ2, myb_a + 2 * BOX_BYTES, c0, r0);
          }
        }
        bulk_commit();
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward: dz1, dz2 from dy W2 and the saved pre-activations
// ------------------------------------------------------------------------------------------------------------------
constexpr int DG_STAGES = 3;
constexpr uint32_t DG_WARP_BYTES = 4 * BOX_BYTES;  // two (z1, z2) box pairs per warp
constexpr uint32_t DG_SMEM = 1024 + DG_STAGES * (A_BYTES + B_BYTES) + NUM_EPI_WARPS * DG_WARP_BYTES + BAR_BYTES;
static_assert(DG_SMEM <= 232448, "shared memory budget");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
dglu_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmZ1, const __grid_constant__ CUtensorMap tmZ2,
            const __grid_constant__ CUtensorMap tmD1, const __grid_constant__ CUtensorMap tmD2, int M, int N, int K,
            int swz, GluArgs ga) {
  pdl_trigger();
  constexpr int ST = DG_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + ST * A_BYTES;
  uint8_t* sEpi = sB + ST * B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + NUM_EPI_WARPS * DG_WARP_BYTES);
  uint64_t* full = bars;                 // [ST]
  uint64_t* empty = bars + ST;           // [ST]
  uint64_t* tfull = bars + 2 * ST;       // [2]
  uint64_t* tempty = bars + 2 * ST + 2;  // [2]
  uint64_t* inbar = bars + 2 * ST + 4;   // [NUM_EPI_WARPS][2]: one per box pair
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(inbar + 2 * NUM_EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int i = 0; i < ST; ++i) {
      mbar_init(smem_u32(&full[i]), 1);
      mbar_init(smem_u32(&empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tfull[i]), 1);
      mbar_init(smem_u32(&tempty[i]), 2 * NUM_EPI_WARPS);
    }
    for (int i = 0; i < 2 * NUM_EPI_WARPS; ++i) mbar_init(smem_u32(&inbar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int tiles_n = (N + BN - 1) / BN;
  const int tiles_m = (M + 2 * BM - 1) / (2 * BM);
  const int total = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer: dy rows (K-major) and W2 [K, N] (MN-major) =================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < total; tile += npairs) {
        const int m0 = (tile / tiles_n) * (2 * BM) + (int)rank * BM;
        const int nb0 = (tile % tiles_n) * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          MBAR_WAIT_LONG(smem_u32(&empty[stage]), phase ^ 1);
          const uint32_t fb_local = smem_u32(&full[stage]);
          if (rank == 0) mbar_expect_tx(fb_local, 2 * (A_BYTES + B_BYTES));
          const uint32_t fb = mapa(fb_local, 0);
          const uint32_t b_dst = smem_u32(sB + stage * B_BYTES);
          tma_load_2d_2sm(smem_u32(sA + stage * A_BYTES), &tmA, fb, kb * BK, m0);
#pragma unroll
          for (int a = 0; a < (BN / 2) / 64; ++a) tma_load_2d_2sm(b_dst + a * (BK * 128), &tmB, fb, nb0 + a * 64, kb * BK);
          if (++stage == ST) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)((2 * BM) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < total; tile += npairs) {
        MBAR_WAIT_LONG(smem_u32(&tempty[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&full[stage]), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc_mma2_bf16(d_tmem, make_smem_desc<false>(a_addr, k), make_smem_desc<true>(b_addr, k), idesc,
                         (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit2_mc(smem_u32(&empty[stage]));
          if (++stage == ST) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit2_mc(smem_u32(&tfull[acc]));
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ================= epilogue: warp (q, s) owns rows [32 q, +32) x columns [64 s, +64) = two 32-column boxes =========
    const int ew = warp - 2;
    const int q = warp & 3;
    const int s = ew >> 2;
    uint8_t* myb = sEpi + ew * DG_WARP_BYTES;  // pair p: z1 / dz1 in X[2p], z2 / dz2 in X[2p + 1]
    const uint32_t myb_a = smem_u32(myb);
    const uint32_t ibar0 = smem_u32(&inbar[2 * ew]);
    const uint32_t tempty_leader0 = mapa(smem_u32(&tempty[0]), 0);
    DropCtx dc;
    dc.on = ga.p_drop > 0.0f;
    dc.thr = dc.on ? drop_threshold(ga.p_drop) : 0u;
    dc.inv_keep = dc.on ? 1.0f / (1.0f - ga.p_drop) : 1.0f;
    dc.key = dc.on ? drop_key(ga.seed, ga.site) : 0u;

    auto rows_of = [&](int tile) { return (tile / tiles_n) * (2 * BM) + (int)rank * BM + q * 32; };
    auto cols_of = [&](int tile) { return (tile % tiles_n) * BN + s * 64; };
    // z1 / z2 of the box `ahead` boxes after (tile, box) into pair p; boxes outside the output are neither loaded nor
    // waited for
    auto issue_in = [&](int tile, int box, int ahead, int p) {
      box += ahead;
      while (box >= 2) {
        box -= 2;
        tile += npairs;
      }
      if (tile >= total) return;
      const int r0 = rows_of(tile), c = cols_of(tile) + box * 32;
      if (r0 >= M || c >= N) return;
      if (lane == 0) {
        const uint32_t bar = ibar0 + (uint32_t)p * 8u;
        mbar_expect_tx(bar, 2 * BOX_BYTES);
        tma_load_2d(myb_a + (uint32_t)(2 * p) * BOX_BYTES, &tmZ1, bar, c, r0);
        tma_load_2d(myb_a + (uint32_t)(2 * p + 1) * BOX_BYTES, &tmZ2, bar, c, r0);
      }
    };
    if (pair < total) {
      issue_in(pair, 0, 0, 0);
      issue_in(pair, 0, 1, 1);
    }
    uint32_t ph = 0;  // bit p: parity of the next completion of pair p's load barrier
    int p = 0;        // running box index mod 2
    bool first_box = true;

    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < total; tile += npairs) {
      const int r0 = rows_of(tile);
      const int c0 = cols_of(tile);
      const long long row = (long long)r0 + lane;
      mbar_wait(smem_u32(&tfull[acc]), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + s * 64);
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int bcol = c0 + b * 32;
        const bool live = r0 < M && bcol < N;  // warp-uniform
        uint8_t* x1 = myb + (2 * p) * BOX_BYTES;
        uint8_t* x2 = x1 + BOX_BYTES;
        if (live) {
          mbar_wait(ibar0 + (uint32_t)p * 8u, (ph >> p) & 1u);
          ph ^= 1u << p;
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t raw[16];
          tmem_ld16_nowait(t_row + (uint32_t)(b * 32 + c * 16), raw);
          tmem_wait_ld();
          if (b == 1 && c == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_leader0 + (uint32_t)acc * 8u);
          }
          float z1[16], z2[16], d1[16], d2[16];
          if (live) {
            ld_box16(x1, lane, c, swz, z1);
            ld_box16(x2, lane, c, swz, z2);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) z1[j] = z2[j] = 0.f;
          }
          float ds[16];
          if (dc.on)
            drop_scales16(dc, (uint32_t)((unsigned long long)row * (unsigned long long)ga.drop_ld +
                                         (unsigned long long)(bcol + c * 16)), ds);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float da = __uint_as_float(raw[j]);
            if (dc.on) da *= ds[j];
            float g, dg;
            gelu_both<true>(z1[j], g, dg);
            d1[j] = da * z2[j] * dg;
            d2[j] = da * g;
          }
          if (c == 0 && !first_box) {
            // the store of the previous box (the other pair) has had this chunk's math to drain: its boxes become the
            // landing zone of the box after this one
            if (lane == 0) bulk_wait_read0();
            issue_in(tile, b, 1, p ^ 1);
            __syncwarp();
          }
          st_box16(x1, lane, c, swz, d1);
          st_box16(x2, lane, c, swz, d2);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (live) {
            tma_store_2d(&tmD1, myb_a + (uint32_t)(2 * p) * BOX_BYTES, bcol, r0);
            tma_store_2d(&tmD2, myb_a + (uint32_t)(2 * p + 1) * BOX_BYTES, bcol, r0);
          }
          bulk_commit();
        }
        p ^= 1;
        first_box = false;
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

static int max_pairs_of(const void* kern, size_t smem) {
  cudaLaunchConfig_t q{};
  q.gridDim = dim3(2 * (num_sms() / 2));
  q.blockDim = dim3(NUM_THREADS);
  q.dynamicSmemBytes = smem;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess || n <= 0) n = num_sms() / 2;
  return n < num_sms() / 2 ? n : num_sms() / 2;
}

static bool ok16(const void* p, long long ld) {
  return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ((ld * 2) & 15) == 0;
}

static int box_swizzle(int* swz_flag) {
  static int swz = -1;
  if (swz < 0) {
    const char* e = getenv("MMA_GEMM2_SWZ");
    swz = e ? atoi(e) : 1;
  }
  *swz_flag = swz ? 1 : 0;
  return swz ? (int)CU_TENSOR_MAP_SWIZZLE_64B : (int)CU_TENSOR_MAP_SWIZZLE_NONE;
}

}  // namespace tc2

// a[M,N] = drop(gelu(h W1^T + b1) * (h Wg^T + bg)); z1 / z2 (both or neither): the bf16 pre-activations the backward
// needs.  h: bf16 [M, K]; W1, Wg: bf16 [N, K]; b1, bg: fp32 [N] (16-byte aligned); a, z1, z2: bf16 [M, N].
// Dropout (p_drop > 0): element (row, col) uses stream index row * N + col of (seed, site), as mma_gemm_bf16's
// EPI_GLU_MUL does.  MMA_ERR_UNSUPPORTED outside the kernel's envelope (alignment, N % 16, too few tiles): the caller
// then runs the two products separately (EPI_STORE + EPI_GLU_MUL).
extern "C" int mma_ffn_glu_fwd(const void* h, long long ldh, const void* W1, long long ldw1, const void* Wg,
                               long long ldwg, const float* b1, const float* bg, int M, int N, int K, void* a,
                               long long lda, void* z1, long long ldz1, void* z2, long long ldz2, float p_drop,
                               unsigned long long seed, unsigned int site, cudaStream_t stream) {
  using namespace tc2;
  if (M <= 0 || N <= 0 || K <= 0 || !h || !W1 || !Wg || !b1 || !bg || !a || ((z1 == nullptr) != (z2 == nullptr)))
    return MMA_ERR_ARG;
  if (!ok16(h, ldh) || !ok16(W1, ldw1) || !ok16(Wg, ldwg) || !ok16(a, lda) || (z1 && (!ok16(z1, ldz1) || !ok16(z2, ldz2))) ||
      (reinterpret_cast<uintptr_t>(b1) & 15) || (reinterpret_cast<uintptr_t>(bg) & 15) || (N & 15) || (p_drop > 0.f && (N & 1)))
    return MMA_ERR_UNSUPPORTED;
  static int min_tiles = -1;
  if (min_tiles < 0) {
    const char* e = getenv("MMA_GEMM2_MIN_TILES");
    min_tiles = e ? atoi(e) : 48;
    const char* g = getenv("MMA_GEMM2");
    if (g && atoi(g) == 0) min_tiles = 1 << 30;
  }
  const long long tiles = (long long)((M + 2 * BM - 1) / (2 * BM)) * ((N + GLU_BN - 1) / GLU_BN);
  if (M < 256 || tiles < min_tiles) return MMA_ERR_UNSUPPORTED;
  int swz;
  const int box_swz = box_swizzle(&swz);
  CUtensorMap tmA, tmB1, tmB2, tmOut, tmZ1, tmZ2;
  int rc = make_map(&tmA, h, (unsigned long long)K, (unsigned long long)M, ldh, BK, BM);
  if (rc) return rc;
  if ((rc = make_map(&tmB1, W1, (unsigned long long)K, (unsigned long long)N, ldw1, BK, GLU_BN))) return rc;
  if ((rc = make_map(&tmB2, Wg, (unsigned long long)K, (unsigned long long)N, ldwg, BK, GLU_BN))) return rc;
  if ((rc = make_map_ex(&tmOut, a, (unsigned long long)N, (unsigned long long)M, lda, 32, 32, 0, box_swz))) return rc;
  tmZ1 = tmOut;
  tmZ2 = tmOut;
  if (z1) {
    if ((rc = make_map_ex(&tmZ1, z1, (unsigned long long)N, (unsigned long long)M, ldz1, 32, 32, 0, box_swz))) return rc;
    if ((rc = make_map_ex(&tmZ2, z2, (unsigned long long)N, (unsigned long long)M, ldz2, 32, 32, 0, box_swz))) return rc;
  }
  GluArgs ga{b1, bg, p_drop, seed, site, (long long)N};
  auto kern = z1 ? glu_fwd_kernel<true> : glu_fwd_kernel<false>;
  static int max_pairs[2] = {0, 0};
  int& mp = max_pairs[z1 ? 1 : 0];
  if (!mp) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return MMA_ERR_LAUNCH;
    mp = max_pairs_of(reinterpret_cast<const void*>(kern), SMEM);
  }
  const int pairs = tiles < mp ? (int)tiles : mp;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, kern, tmA, tmB1, tmB2, tmOut, tmZ1, tmZ2, M, N, K, swz, ga) != cudaSuccess)
    return MMA_ERR_LAUNCH;
  return MMA_OK;
}

// Backward of the gate: with da = (dy W2) * dropmask  (dy: bf16 [M, K], W2: bf16 [K, N] = linear2.weight as stored),
// dz1 = da * z2 * gelu'(z1) and dz2 = da * gelu(z1); z1, z2, dz1, dz2: bf16 [M, N].  The dropout stream is the forward's
// (seed, site, row * drop_ld + col).  MMA_ERR_UNSUPPORTED outside the envelope: run mma_gemm_bf16 with EPI_DGLU.
extern "C" int mma_ffn_dglu(const void* dy, long long lddy, const void* W2, long long ldw2, int M, int N, int K,
                            const void* z1, long long ldz1, const void* z2, long long ldz2, void* dz1, long long lddz1,
                            void* dz2, long long lddz2, float p_drop, unsigned long long seed, unsigned int site,
                            long long drop_ld, cudaStream_t stream) {
  using namespace tc2;
  if (M <= 0 || N <= 0 || K <= 0 || !dy || !W2 || !z1 || !z2 || !dz1 || !dz2) return MMA_ERR_ARG;
  if (!ok16(dy, lddy) || !ok16(W2, ldw2) || !ok16(z1, ldz1) || !ok16(z2, ldz2) || !ok16(dz1, lddz1) || !ok16(dz2, lddz2))
    return MMA_ERR_UNSUPPORTED;
  static int min_tiles = -1;
  if (min_tiles < 0) {
    const char* e = getenv("MMA_GEMM2_MIN_TILES");
    min_tiles = e ? atoi(e) : 48;
    const char* g = getenv("MMA_GEMM2");
    if (g && atoi(g) == 0) min_tiles = 1 << 30;
  }
  const long long tiles = (long long)((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  if (M < 512 || N < 256 || tiles < min_tiles) return MMA_ERR_UNSUPPORTED;
  int swz;
  const int box_swz = box_swizzle(&swz);
  CUtensorMap tmA, tmB, tmZ1, tmZ2, tmD1, tmD2;
  int rc = make_map(&tmA, dy, (unsigned long long)K, (unsigned long long)M, lddy, BK, BM);
  if (rc) return rc;
  if ((rc = make_map(&tmB, W2, (unsigned long long)N, (unsigned long long)K, ldw2, 64, BK))) return rc;
  if ((rc = make_map_ex(&tmZ1, z1, (unsigned long long)N, (unsigned long long)M, ldz1, 32, 32, 0, box_swz))) return rc;
  if ((rc = make_map_ex(&tmZ2, z2, (unsigned long long)N, (unsigned long long)M, ldz2, 32, 32, 0, box_swz))) return rc;
  if ((rc = make_map_ex(&tmD1, dz1, (unsigned long long)N, (unsigned long long)M, lddz1, 32, 32, 0, box_swz))) return rc;
  if ((rc = make_map_ex(&tmD2, dz2, (unsigned long long)N, (unsigned long long)M, lddz2, 32, 32, 0, box_swz))) return rc;
  GluArgs ga{nullptr, nullptr, p_drop, seed, site, drop_ld > 0 ? drop_ld : (long long)N};
  static int mp = 0;
  if (!mp) {
    if (cudaFuncSetAttribute(dglu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_SMEM) != cudaSuccess)
      return MMA_ERR_LAUNCH;
    mp = max_pairs_of(reinterpret_cast<const void*>(dglu_kernel), DG_SMEM);
  }
  const int pairs = tiles < mp ? (int)tiles : mp;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = DG_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, dglu_kernel, tmA, tmB, tmZ1, tmZ2, tmD1, tmD2, M, N, K, swz, ga) != cudaSuccess)
    return MMA_ERR_LAUNCH;
  return MMA_OK;
}
