// 2-CTA (cta_group::2) tcgen05 GEMM for the large products of the training step.
//
//   C[M,N] = epilogue( A[M,K] * B_op[N,K]^T )      bf16 operands, fp32 accumulate in TMEM
//
// A CTA pair (a 2x1 cluster on one TPC) owns a 256 x 256 output tile: each CTA stages its own 128 rows of A and HALF
// of the B tile (128 of the 256 columns) per k-step, the leader CTA issues tcgen05.mma.cta_group::2 (M = 256) which
// reads both halves, and every CTA drains its own 128 x 256 accumulator.  Against the single-CTA 128 x 256 kernel
// (gemm_tc.cu) this cuts the L2 -> shared-memory operand traffic by a third (32 KB instead of 48 KB per CTA and
// k-step) - the K = 512 products of the d = 512 model are bound by exactly that traffic - and frees ~96 KB of shared
// memory for the epilogue.
//
// Epilogue: no per-thread global loads / stores.  tcgen05.ld hands each thread one accumulator ROW, so direct global
// accesses touch 32 different lines per warp instruction (measured: the epilogue, not the MMA, bounded every K = 512
// product).  Here each epilogue warp owns three private 2 KB shared-memory boxes ([32 rows][64 bytes], 64-byte
// swizzle): side inputs (residual / saved pre-activation / running gradient) arrive by TMA loads prefetched one box
// ahead, results leave by TMA stores (cp.async.bulk.tensor ... bulk_group).  The 16 epilogue warps never synchronise
// with each other.
//
// Roles per CTA (576 threads): warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA only; both CTAs allocate
// TMEM), warps 2..17 = epilogue.  Two TMEM accumulator stages (2 x 256 columns) overlap epilogue and mainloop.
#include "tc2.cuh"

namespace tc2 {

// ---- epilogue math on one 16-column chunk of one row ------------------------------------------------------------
// Semantics identical to epilogue_store<16, KIND, true> (common.cuh); the dropout key / threshold are hoisted.
template <int KIND>
__device__ __forceinline__ void chunk_math(const Epi& ep, const DropCtx& dc, long long row, int col, float (&v)[16],
                                           float (&o2)[16], const float (&in)[16], const float (&bv)[16], bool has_bias) {
  if (KIND == EPI_ACCUM) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = ep.accumulate == 1 ? fmaf(v[j], ep.alpha, in[j]) : v[j] * ep.alpha;
    return;
  }
  if (KIND == EPI_STORE) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = has_bias ? fmaf(v[j], ep.alpha, bv[j]) : v[j] * ep.alpha;
    return;
  }
  if (has_bias && (KIND == EPI_GELU || KIND == EPI_RESID)) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] += bv[j];
  }
  // keep-scales; for the GELU kinds the 1/2 of Phi = (1 + erf) / 2 rides along (gelu_tanh_scaled)
  constexpr bool HALF = MMA_GELU_TANH && (KIND == EPI_GELU || KIND == EPI_DGELU);
  const float sc = HALF ? 0.5f * dc.inv_keep : dc.inv_keep;
  float ds[16];
  if (dc.on) {
    const uint32_t e0 = (uint32_t)((unsigned long long)row * (unsigned long long)ep.drop_ld + (unsigned long long)col);
    if ((e0 & 1u) == 0) {
      const uint32_t thr_hi = dc.thr << 16;  // (r >> 16) >= thr  <=>  r >= thr << 16 (thr <= 65535)
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const uint32_t r = drop_pair(dc.key, (e0 + j) >> 1);
        ds[j] = (r & 0xFFFFu) >= dc.thr ? sc : 0.f;
        ds[j + 1] = r >= thr_hi ? sc : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) ds[j] = drop_scale1(dc.key, e0 + j, dc.thr, sc);
    }
  } else if (HALF) {
#pragma unroll
    for (int j = 0; j < 16; ++j) ds[j] = 0.5f;
  }
  if (KIND == EPI_GELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      o2[j] = v[j];
      if (MMA_GELU_TANH) {
        v[j] = gelu_tanh_scaled(v[j], ds[j]);
      } else {
        float y = gelu_t<true>(v[j]);
        if (dc.on) y *= ds[j];
        v[j] = y;
      }
    }
  } else if (KIND == EPI_RESID) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = in[j] + (dc.on ? v[j] * ds[j] : v[j]);
  } else if (KIND == EPI_DGELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (MMA_GELU_TANH) {
        v[j] *= dgelu_tanh_scaled(in[j], ds[j]);
      } else {
        float y = v[j] * dgelu_t<true>(in[j]);
        if (dc.on) y *= ds[j];
        v[j] = y;
      }
    }
  }
}


// IO32: side input and outputs are fp32 (16 columns per box); otherwise bf16 (32 columns per box)
template <bool B_MN, int KIND, bool IO32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut2,
             const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmA2,
             const __grid_constant__ CUtensorMap tmB2, int M, int N, int K, int K2, int flags, Epi ep) {
  pdl_trigger();
  // flags: bit0 epilogue boxes use the 64-byte swizzle; diagnostics (MMA_GEMM_DBG): bit8 epilogue skips TMA loads /
  // stores, bit9 no operand loads and no MMAs, bit10 operand loads but no MMAs
  const int swz = flags & 1;
  const int dbg = flags >> 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* sEpi = sB + STAGES * B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + NUM_EPI_WARPS * EPI_WARP_BYTES);
  uint64_t* full = bars;                       // [STAGES]  (only the leader's are used)
  uint64_t* empty = bars + STAGES;             // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;         // [2]
  uint64_t* tempty = bars + 2 * STAGES + 2;    // [2]       (only the leader's are used)
  uint64_t* inbar = bars + 2 * STAGES + 4;     // [NUM_EPI_WARPS][3]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(inbar + 3 * NUM_EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmOut)) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(smem_u32(&full[i]), 1);
      mbar_init(smem_u32(&empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tfull[i]), 1);
      mbar_init(smem_u32(&tempty[i]), 2 * NUM_EPI_WARPS);  // the epilogue warps of BOTH CTAs
    }
    for (int i = 0; i < 3 * NUM_EPI_WARPS; ++i) mbar_init(smem_u32(&inbar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated, before any cross-CTA arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail

  const int tiles_n = (N + BN - 1) / BN;
  const int tiles_m = (M + 2 * BM - 1) / (2 * BM);
  const int total = tiles_m * tiles_n;
  // K2 > 0: the reduction continues over a second operand pair, C = A B^T + A2 B2^T (dh = dz1 W1 + dz2 Wg of the gated FFN)
  const int num_kb1 = (K + BK - 1) / BK;
  const int num_kb = num_kb1 + (K2 + BK - 1) / BK;
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0 && !(dbg & 2)) {
      // ================= TMA producer (both CTAs) =================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < total; tile += npairs) {
        const int m0 = (tile / tiles_n) * (2 * BM) + (int)rank * BM;
        const int nb0 = (tile % tiles_n) * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          MBAR_WAIT_LONG(smem_u32(&empty[stage]), phase ^ 1);
          const uint32_t fb_local = smem_u32(&full[stage]);
          if (rank == 0) mbar_expect_tx(fb_local, 2 * (A_BYTES + B_BYTES));  // both CTAs' bytes land on this barrier
          const uint32_t fb = mapa(fb_local, 0);
          const uint32_t a_dst = smem_u32(sA + stage * A_BYTES);
          const uint32_t b_dst = smem_u32(sB + stage * B_BYTES);
          const bool second = kb >= num_kb1;
          const CUtensorMap* ta = second ? &tmA2 : &tmA;
          const CUtensorMap* tb = second ? &tmB2 : &tmB;
          const int k0 = (second ? kb - num_kb1 : kb) * BK;
          tma_load_2d_2sm(a_dst, ta, fb, k0, m0);
          if (!B_MN) {
            tma_load_2d_2sm(b_dst, tb, fb, k0, nb0);
          } else {
#pragma unroll
            for (int a = 0; a < (BN / 2) / 64; ++a) tma_load_2d_2sm(b_dst + a * (BK * 128), tb, fb, nb0 + a * 64, k0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ================= MMA issuer (one thread of the leader CTA) =================
      // instruction descriptor: c=f32 [4,6), a=bf16 [7,10), b=bf16 [10,13), b_major bit16, N>>3 [17,23), M>>4 [24,29)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((B_MN ? 1u : 0u) << 16) |
                                 ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < total; tile += npairs) {
        MBAR_WAIT_LONG(smem_u32(&tempty[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb && !(dbg & 2); ++kb) {
          mbar_wait(smem_u32(&full[stage]), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if (!(dbg & 4))
              tc_mma2_bf16(d_tmem, make_smem_desc<false>(a_addr, k), make_smem_desc<B_MN>(b_addr, k), idesc,
                           (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit2_mc(smem_u32(&empty[stage]));  // frees the slot in both CTAs once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit2_mc(smem_u32(&tfull[acc]));  // accumulator complete -> epilogue warps of both CTAs
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ================= epilogue warps (TMEM -> registers -> shared boxes -> TMA store) =================
    const int ew = warp - 2;
    const int q = warp & 3;   // TMEM lane quarter this warp may read: rows [32 q, 32 q + 32) of the CTA tile
    const int s = ew >> 2;    // 64-column slice of the tile
    constexpr int CPB = IO32 ? 1 : 2;        // 16-column chunks per box
    constexpr int NBOX = 4 / CPB;            // boxes per tile and warp
    constexpr int BOXCOLS = 16 * CPB;
    constexpr bool IN_KIND = KIND == EPI_RESID || KIND == EPI_DGELU || KIND == EPI_ACCUM;
    const bool has_in = IN_KIND && (KIND != EPI_ACCUM || ep.accumulate == 1) && !(dbg & 1);
    const bool has_out2 = KIND == EPI_GELU && ep.out2 != nullptr;
    const bool has_bias = ep.bias != nullptr && (KIND == EPI_STORE || KIND == EPI_GELU || KIND == EPI_RESID);
    // Three 2 KB boxes X[0..2] per warp, used in rotation by the running box index g:
    //   kinds with a side input: box g's input is TMA-loaded into X[g % 3] two boxes ahead; its results are written
    //     IN PLACE over the input and stored from there, so a load never waits for the store that precedes it by
    //     one box and a store's shared-memory read has a whole box of math to drain;
    //   kinds without: outputs rotate through the three boxes (two per box for GELU + pre-activation copy).
    uint8_t* myb = sEpi + ew * EPI_WARP_BYTES;
    const uint32_t myb_a = smem_u32(myb);
    const uint32_t ibar0 = smem_u32(&inbar[3 * ew]);
    const uint32_t tempty_leader0 = mapa(smem_u32(&tempty[0]), 0);

    DropCtx dc;
    dc.on = ep.p_drop > 0.0f && (KIND == EPI_GELU || KIND == EPI_RESID || KIND == EPI_DGELU);
    dc.thr = dc.on ? drop_threshold(ep.p_drop) : 0u;
    dc.inv_keep = dc.on ? 1.0f / (1.0f - ep.p_drop) : 1.0f;
    dc.key = dc.on ? drop_key(ep.seed, ep.site) : 0u;

    // coordinates of this warp's rows / columns in tile `tile`
    auto rows_of = [&](int tile) { return (tile / tiles_n) * (2 * BM) + (int)rank * BM + q * 32; };
    auto cols_of = [&](int tile) { return (tile % tiles_n) * BN + s * 64; };
    // side-input load of the box `ahead` boxes after (tile, box) into X[k]; boxes entirely outside the output (tail
    // rows / columns) are neither loaded nor waited for
    auto issue_in = [&](int tile, int box, int ahead, int k) {
      box += ahead;
      while (box >= NBOX) {
        box -= NBOX;
        tile += npairs;
      }
      if (tile >= total) return;
      const int r0 = rows_of(tile), c = cols_of(tile) + box * BOXCOLS;
      if (r0 >= M || c >= N) return;
      if (lane == 0) {
        const uint32_t bar = ibar0 + (uint32_t)k * 8u;
        mbar_expect_tx(bar, BOX_BYTES);
        tma_load_2d(myb_a + (uint32_t)k * BOX_BYTES, &tmIn, bar, c, r0);
      }
    };
    if (has_in && pair < total) {
      issue_in(pair, 0, 0, 0);
      issue_in(pair, 0, 1, 1);
    }
    uint32_t ph = 0;  // bit k: parity of the next completion of X[k]'s load barrier
    int gk = 0;       // g % 3 (kinds with a side input, STORE, ACCUM) or (2 g) % 3 (GELU with two outputs)

    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < total; tile += npairs) {
      const int r0 = rows_of(tile);
      const int c0 = cols_of(tile);
      const long long row = (long long)r0 + lane;
      // bias: four warp-uniform 16-byte loads per 16-column chunk (L1 broadcast; the shuffles this replaces were one
      // MIO instruction per element in an issue-bound loop); index clamped for tiles past N (those columns are clipped
      // by the TMA store).  The host checks bias % 16 B == 0 and N % 4 == 0.
      const float4* bias4 = reinterpret_cast<const float4*>(ep.bias);
      const int last4 = (N >> 2) - 1;
      mbar_wait(smem_u32(&tfull[acc]), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + s * 64);
      uint32_t raw[2][16];
      tmem_ld16_nowait(t_row, raw[0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int box = c / CPB;
        const bool first = (c % CPB) == 0, last = (c % CPB) == CPB - 1;
        const int bcol = c0 + box * BOXCOLS;            // first column of this box
        const bool live = r0 < M && bcol < N && !(dbg & 1);  // warp-uniform: the box touches the output at all
        const int k1 = gk;                               // box of output 1 (and of the side input)
        const int k2 = gk == 2 ? 0 : gk + 1;             // box of output 2
        if (first && has_in && live) {
          mbar_wait(ibar0 + (uint32_t)k1 * 8u, (ph >> k1) & 1u);
          ph ^= 1u << k1;
        }
        float v[16], o2[16], in[16], bv[16];
        if (has_bias) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 f = __ldg(bias4 + min(((c0 + c * 16) >> 2) + k, last4));
            bv[4 * k] = f.x; bv[4 * k + 1] = f.y; bv[4 * k + 2] = f.z; bv[4 * k + 3] = f.w;
          }
        }
        tmem_wait_ld();
        if (c + 1 < 4) {
          tmem_ld16_nowait(t_row + (uint32_t)((c + 1) * 16), raw[(c + 1) & 1]);
        } else {
          // the accumulator stage is drained: hand it back to the MMA issuer (leader CTA)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader0 + (uint32_t)acc * 8u);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[c & 1][j]);
        uint8_t* x1 = myb + k1 * BOX_BYTES;
        if (has_in && live) {
          if (IO32) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float4 f = *reinterpret_cast<const float4*>(x1 + box_off(lane, k, swz));
              in[4 * k] = f.x; in[4 * k + 1] = f.y; in[4 * k + 2] = f.z; in[4 * k + 3] = f.w;
            }
          } else {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint4 u = *reinterpret_cast<const uint4*>(x1 + box_off(lane, 2 * (c % CPB) + k, swz));
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                const float2 f = __bfloat1622float2(h[w]);
                in[8 * k + 2 * w] = f.x;
                in[8 * k + 2 * w + 1] = f.y;
              }
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) in[j] = 0.f;
        }
        chunk_math<KIND>(ep, dc, row, c0 + c * 16, v, o2, in, bv, has_bias);
        if (first) {
          if (IN_KIND) {
            // the store issued one box ago has had this box's math to drain: its box becomes the landing zone of the
            // input two boxes ahead
            if (lane == 0) bulk_wait_read0();
            if (has_in) issue_in(tile, box, 2, gk == 0 ? 2 : gk - 1);
            __syncwarp();
          } else {
            // the stores that last read the box(es) written next must have finished reading shared memory
            if (lane == 0) {
              if (has_out2) bulk_wait_read0();
              else asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
            }
            __syncwarp();
          }
        }
        if (IO32) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<float4*>(x1 + box_off(lane, k, swz)) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        } else {
#pragma unroll
          for (int k = 0; k < 2; ++k)
            *reinterpret_cast<uint4*>(x1 + box_off(lane, 2 * (c % CPB) + k, swz)) =
                make_uint4(pack2_bf16(v[8 * k], v[8 * k + 1]), pack2_bf16(v[8 * k + 2], v[8 * k + 3]),
                           pack2_bf16(v[8 * k + 4], v[8 * k + 5]), pack2_bf16(v[8 * k + 6], v[8 * k + 7]));
          if (has_out2) {
            uint8_t* x2 = myb + k2 * BOX_BYTES;
#pragma unroll
            for (int k = 0; k < 2; ++k)
              *reinterpret_cast<uint4*>(x2 + box_off(lane, 2 * (c % CPB) + k, swz)) =
                  make_uint4(pack2_bf16(o2[8 * k], o2[8 * k + 1]), pack2_bf16(o2[8 * k + 2], o2[8 * k + 3]),
                             pack2_bf16(o2[8 * k + 4], o2[8 * k + 5]), pack2_bf16(o2[8 * k + 6], o2[8 * k + 7]));
          }
        }
        if (last) {
          fence_async_smem();  // generic-proxy writes -> visible to the TMA engine
          __syncwarp();
          if (lane == 0) {
            if (live) {
              tma_store_2d(&tmOut, myb_a + (uint32_t)k1 * BOX_BYTES, bcol, r0);
              if (has_out2) tma_store_2d(&tmOut2, myb_a + (uint32_t)k2 * BOX_BYTES, bcol, r0);
            }
            bulk_commit();  // one group per box, live or not: wait_group counts stay aligned with the rotation
          }
          gk += has_out2 ? 2 : 1;
          if (gk >= 3) gk -= 3;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();  // nobody exits while its peer may still signal its barriers or read its operand tiles
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

template <bool B_MN, int KIND, bool IO32>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmOut2,
                  const CUtensorMap& tmIn, const CUtensorMap& tmA2, const CUtensorMap& tmB2, int M, int N, int K, int K2,
                  int flags, const Epi& ep, cudaStream_t stream) {
  auto kern = gemm2_kernel<B_MN, KIND, IO32>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return MMA_ERR_LAUNCH;
    attr_set = true;
  }
  const int total = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  // persistent grid = the CTA pairs that can be co-resident (a GPC with an odd SM count strands one SM)
  static int max_pairs = 0;
  if (!max_pairs) {
    cudaLaunchConfig_t q{};
    q.gridDim = dim3(2 * (num_sms() / 2));
    q.blockDim = dim3(NUM_THREADS);
    q.dynamicSmemBytes = SMEM;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess || n <= 0) n = num_sms() / 2;
    max_pairs = n < num_sms() / 2 ? n : num_sms() / 2;
    if (getenv("MMA_GEMM2_VERBOSE")) fprintf(stderr, "[gemm2] co-resident CTA pairs: %d\n", n);
  }
  int pairs = max_pairs;
  if (total < pairs) pairs = total;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;  // the cluster shape is a compile-time attribute of the kernel (__cluster_dims__)
  if (cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmOut, tmOut2, tmIn, tmA2, tmB2, M, N, K, K2, flags, ep) != cudaSuccess)
    return MMA_ERR_LAUNCH;
  return MMA_OK;
}


// ------------------------------------------------------------------------------------------------------------------
// Residual product fused with the LayerNorm that follows it (d_model = 512 only):
//   x_new[M,512] = resid + drop(A[M,K] W[512,K]^T + bias)            fp32 residual stream
//   h[M,512]     = LayerNorm(x_new) * gamma + beta                   bf16 operand of the next product
// Every out-projection / FFN-2 of the pre-LN layers is followed by exactly one LayerNorm of its result (the next
// sub-layer's, or the stack's final norm), which used to be a separate pass over x_new (read 4 B + write 2 B per
// element and a launch).  Here a CTA pair owns 256 rows x ALL 512 columns: each CTA's 128 x 512 fp32 accumulator
// fills its TMEM (two N = 256 MMAs per k-step); a thread owns one row of its warp's 128-column slice, so the row
// statistics are thread-local sums plus one exchange between the four warps of a row quarter.  Pass 1 forms x_new,
// stores it (TMA boxes, as in gemm2_kernel) and parks it back in TMEM; pass 2 re-reads it, normalises and stores h.
// ------------------------------------------------------------------------------------------------------------------
constexpr int LN_N = 512;
constexpr int LN_STAGES = 2;
constexpr uint32_t LNB_BYTES = 256 * BK * 2;  // per-CTA B tile: 256 of the 512 weight rows
constexpr uint32_t LN_STATS_BYTES = 4 * 128 * 8;
constexpr uint32_t LN_SMEM = 1024 + LN_STAGES * (A_BYTES + LNB_BYTES) + NUM_EPI_WARPS * EPI_WARP_BYTES + LN_STATS_BYTES + BAR_BYTES;
static_assert(LN_SMEM <= 232448, "shared memory budget");

struct LnExtra {
  const float* gamma;
  const float* beta;
  float eps;
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm2_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmIn,
                const __grid_constant__ CUtensorMap tmH, int M, int K, int swz, Epi ep, LnExtra ln) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + LN_STAGES * A_BYTES;
  uint8_t* sEpi = sB + LN_STAGES * LNB_BYTES;
  float2* sStats = reinterpret_cast<float2*>(sEpi + NUM_EPI_WARPS * EPI_WARP_BYTES);  // [4 slices][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sStats) + LN_STATS_BYTES);
  uint64_t* full = bars;                      // [LN_STAGES] (leader's are used)
  uint64_t* empty = bars + LN_STAGES;         // [LN_STAGES]
  uint64_t* tfull = bars + 2 * LN_STAGES;     // [1]
  uint64_t* tempty = tfull + 1;               // [1] (leader's is used)
  uint64_t* inbar = tempty + 1;               // [NUM_EPI_WARPS][3]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(inbar + 3 * NUM_EPI_WARPS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < LN_STAGES; ++i) {
      mbar_init(smem_u32(&full[i]), 1);
      mbar_init(smem_u32(&empty[i]), 1);
    }
    mbar_init(smem_u32(tfull), 1);
    mbar_init(smem_u32(tempty), 2 * NUM_EPI_WARPS);
    for (int i = 0; i < 3 * NUM_EPI_WARPS; ++i) mbar_init(smem_u32(&inbar[i]), 1);
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

  const int total = (M + 2 * BM - 1) / (2 * BM);
  const int num_kb = (K + BK - 1) / BK;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < total; tile += npairs) {
        const int m0 = tile * (2 * BM) + (int)rank * BM;
        for (int kb = 0; kb < num_kb; ++kb) {
          MBAR_WAIT_LONG(smem_u32(&empty[stage]), phase ^ 1);
          const uint32_t fb_local = smem_u32(&full[stage]);
          if (rank == 0) mbar_expect_tx(fb_local, 2 * (A_BYTES + LNB_BYTES));
          const uint32_t fb = mapa(fb_local, 0);
          tma_load_2d_2sm(smem_u32(sA + stage * A_BYTES), &tmA, fb, kb * BK, m0);
          // this CTA's weight rows: [128 r, 128 r + 128) feed output columns [0, 256), [256 + 128 r, ...) feed [256, 512)
          const uint32_t b_dst = smem_u32(sB + stage * LNB_BYTES);
          tma_load_2d_2sm(b_dst, &tmB, fb, kb * BK, (int)rank * 128);
          tma_load_2d_2sm(b_dst + 128 * 128, &tmB, fb, kb * BK, 256 + (int)rank * 128);
          if (++stage == LN_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, tphase = 0;
      for (int tile = pair; tile < total; tile += npairs) {
        MBAR_WAIT_LONG(smem_u32(tempty), tphase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&full[stage]), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * LNB_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = make_smem_desc<false>(a_addr, k);
            tc_mma2_bf16(tmem_base, ad, make_smem_desc<false>(b_addr, k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            tc_mma2_bf16(tmem_base + 256, ad, make_smem_desc<false>(b_addr + 128 * 128, k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit2_mc(smem_u32(&empty[stage]));
          if (++stage == LN_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit2_mc(smem_u32(tfull));
        tphase ^= 1;
      }
    }
  } else {
    const int ew = warp - 2, q = warp & 3, s = ew >> 2;  // rows [32 q, +32) of the CTA tile, columns [128 s, +128)
    uint8_t* myb = sEpi + ew * EPI_WARP_BYTES;
    const uint32_t myb_a = smem_u32(myb);
    const uint32_t ibar0 = smem_u32(&inbar[3 * ew]);
    const uint32_t tempty_leader = mapa(smem_u32(tempty), 0);
    const int c0 = s * 128;
    // per-lane copies of this warp's 128 bias / gamma / beta values (column c0 + 32 j + lane), read back with shuffles
    float bias_r[4], gam_r[4], bet_r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bias_r[j] = ep.bias ? ep.bias[c0 + 32 * j + lane] : 0.f;
      gam_r[j] = ln.gamma[c0 + 32 * j + lane];
      bet_r[j] = ln.beta[c0 + 32 * j + lane];
    }
    DropCtx dc;
    dc.on = ep.p_drop > 0.0f;
    dc.thr = dc.on ? drop_threshold(ep.p_drop) : 0u;
    dc.inv_keep = dc.on ? 1.0f / (1.0f - ep.p_drop) : 1.0f;
    dc.key = dc.on ? drop_key(ep.seed, ep.site) : 0u;

    auto rows_of = [&](int tile) { return tile * (2 * BM) + (int)rank * BM + q * 32; };
    // residual box `ahead` chunks after (tile, chunk) into X[k]
    auto issue_in = [&](int tile, int chunk, int ahead, int k) {
      chunk += ahead;
      while (chunk >= 8) {
        chunk -= 8;
        tile += npairs;
      }
      if (tile >= total) return;
      const int r0 = rows_of(tile);
      if (r0 >= M) return;
      if (lane == 0) {
        const uint32_t bar = ibar0 + (uint32_t)k * 8u;
        mbar_expect_tx(bar, BOX_BYTES);
        tma_load_2d(myb_a + (uint32_t)k * BOX_BYTES, &tmIn, bar, c0 + chunk * 16, r0);
      }
    };
    if (pair < total) {
      issue_in(pair, 0, 0, 0);
      issue_in(pair, 0, 1, 1);
    }
    uint32_t ph = 0;
    int gk = 0;
    uint32_t tphase = 0;
    for (int tile = pair; tile < total; tile += npairs) {
      const int r0 = rows_of(tile);
      const long long row = (long long)r0 + lane;
      const bool live = r0 < M;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      mbar_wait(smem_u32(tfull), tphase);
      tc_fence_after();
      // ---------------- pass 1: x_new = resid + drop(acc + bias); row statistics; x_new -> global and back to TMEM
      float sum = 0.f, sumsq = 0.f;
      uint32_t raw[2][16];
      tmem_ld16_nowait(t_row, raw[0]);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int k1 = gk;
        if (live) {
          mbar_wait(ibar0 + (uint32_t)k1 * 8u, (ph >> k1) & 1u);
          ph ^= 1u << k1;
        }
        tmem_wait_ld();
        if (c + 1 < 8) tmem_ld16_nowait(t_row + (uint32_t)((c + 1) * 16), raw[(c + 1) & 1]);
        uint8_t* x1 = myb + k1 * BOX_BYTES;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          v[j] = __uint_as_float(raw[c & 1][j]) + __shfl_sync(0xffffffffu, bias_r[c >> 1], (16 * (c & 1) + j) & 31);
        if (dc.on) {
          const uint32_t e0 = (uint32_t)((unsigned long long)row * (unsigned long long)ep.drop_ld + (unsigned long long)(c0 + c * 16));
#pragma unroll
          for (int j = 0; j < 16; j += 2) {  // e0 is even: drop_ld and the column are
            const uint32_t r = drop_pair(dc.key, (e0 + j) >> 1);
            v[j] *= (r & 0xFFFFu) >= dc.thr ? dc.inv_keep : 0.f;
            v[j + 1] *= (r >> 16) >= dc.thr ? dc.inv_keep : 0.f;
          }
        }
        if (live) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 f = *reinterpret_cast<const float4*>(x1 + box_off(lane, k, swz));
            v[4 * k] += f.x; v[4 * k + 1] += f.y; v[4 * k + 2] += f.z; v[4 * k + 3] += f.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          sum += v[j];
          sumsq = fmaf(v[j], v[j], sumsq);
        }
        tmem_st16(t_row + (uint32_t)(c * 16), v);
        // the store issued one chunk ago has drained behind this chunk's math: reuse its box for the residual two ahead
        if (lane == 0) bulk_wait_read0();
        issue_in(tile, c, 2, gk == 0 ? 2 : gk - 1);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          *reinterpret_cast<float4*>(x1 + box_off(lane, k, swz)) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (live) tma_store_2d(&tmOut, myb_a + (uint32_t)k1 * BOX_BYTES, c0 + c * 16, r0);
          bulk_commit();
        }
        gk = gk == 2 ? 0 : gk + 1;
      }
      tmem_wait_st();
      // ---------------- row statistics across the four column slices of this row quarter
      sStats[s * 128 + q * 32 + lane] = make_float2(sum, sumsq);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
      float tsum = 0.f, tsq = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 p = sStats[j * 128 + q * 32 + lane];
        tsum += p.x;
        tsq += p.y;
      }
      const float mean = tsum * (1.0f / LN_N);
      const float rstd = rsqrtf(fmaxf(tsq * (1.0f / LN_N) - mean * mean, 0.f) + ln.eps);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");  // sStats may be rewritten by the next tile
      // ---------------- pass 2: h = (x_new - mean) rstd gamma + beta  -> bf16, 32 columns per box
      tmem_ld16_nowait(t_row, raw[0]);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        tmem_wait_ld();
        if (c + 1 < 8) {
          tmem_ld16_nowait(t_row + (uint32_t)((c + 1) * 16), raw[(c + 1) & 1]);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader);
        }
        float hreg[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float g = __shfl_sync(0xffffffffu, gam_r[c >> 1], (16 * (c & 1) + j) & 31);
          const float b = __shfl_sync(0xffffffffu, bet_r[c >> 1], (16 * (c & 1) + j) & 31);
          hreg[j] = fmaf((__uint_as_float(raw[c & 1][j]) - mean) * rstd, g, b);
        }
        uint8_t* xb = myb + gk * BOX_BYTES;
        if ((c & 1) == 0) {
          // the box written next was last read by the store three boxes ago (pass-1 stores included)
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
          __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < 2; ++k)
          *reinterpret_cast<uint4*>(xb + box_off(lane, 2 * (c & 1) + k, swz)) =
              make_uint4(pack2_bf16(hreg[8 * k], hreg[8 * k + 1]), pack2_bf16(hreg[8 * k + 2], hreg[8 * k + 3]),
                         pack2_bf16(hreg[8 * k + 4], hreg[8 * k + 5]), pack2_bf16(hreg[8 * k + 6], hreg[8 * k + 7]));
        if (c & 1) {
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (live) tma_store_2d(&tmH, myb_a + (uint32_t)gk * BOX_BYTES, c0 + (c - 1) * 16, r0);
            bulk_commit();
          }
          gk = gk == 2 ? 0 : gk + 1;
        }
      }
      tphase ^= 1;
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
// Grouped weight-gradient kernel, CTA-pair version:  dW_g[Nout_g, Kin_g] += dy_g^T x_g  (and db_g += colsum(dy_g)) for
// up to 8 products (one backward layer) in ONE persistent launch.  Pair tiles are 256 x 256, both operands MN-major
// (the rows of dy / x are the reduction dimension), every output tile has exactly one writer (no split-K, no atomics,
// bit-reproducible).  The epilogue adds into the fp32 gradient buffer with TMA reduce-add stores
// (cp.reduce.async.bulk.tensor ... add): no read-modify-write through the SM.  The bias gradient rides along on the
// first tile column of every tile row: one extra N = 16 MMA per k-step multiplies the dy tile with an all-ones B tile
// into 16 spare TMEM columns.
// ------------------------------------------------------------------------------------------------------------------
struct alignas(64) Wg2Problem {
  CUtensorMap tmA;    // dy [R, Nout]: dims {Nout, R}, box {64, 64}
  CUtensorMap tmB;    // x  [R, Kin] : dims {Kin, R},  box {64, 64}
  CUtensorMap tmOut;  // dW [Nout, Kin] fp32: dims {Kin, Nout}, box {16, 32}, 64-byte swizzle
  float* dbias;       // [Nout] fp32, accumulated (+=), may be null
  int M, N, R;        // Nout, Kin, rows
  int tiles_n, tile_begin;
};
struct Wg2Group {
  Wg2Problem p[8];
  int count, total_tiles, swz;
  int segs;  // every tile's reduction range is cut into `segs` work items (all land with TMA reduce-add / atomics)
};

constexpr int WG2_STAGES = 5;
constexpr uint32_t WG2_ONES_BYTES = 2048;
constexpr uint32_t WG2_SMEM = 1024 + WG2_STAGES * (A_BYTES + B_BYTES) + WG2_ONES_BYTES + NUM_EPI_WARPS * BOX_BYTES + BAR_BYTES;
static_assert(WG2_SMEM <= 232448, "shared memory budget");

__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
wgrad2_group_kernel(const __grid_constant__ Wg2Group grp) {
  pdl_trigger();
  constexpr int ST = WG2_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + ST * A_BYTES;
  uint8_t* sOnes = sB + ST * B_BYTES;
  uint8_t* sEpi = sOnes + WG2_ONES_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + NUM_EPI_WARPS * BOX_BYTES);
  uint64_t* full = bars;            // [ST]  leader's are used
  uint64_t* empty = bars + ST;      // [ST]
  uint64_t* tfull = bars + 2 * ST;  // [1]
  uint64_t* tempty = tfull + 1;     // [1]  leader's is used
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  for (int i = threadIdx.x; i < (int)(WG2_ONES_BYTES / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;  // bf16 1.0 pairs (swizzle-invariant: every element equal)
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < ST; ++i) {
      mbar_init(smem_u32(&full[i]), 1);
      mbar_init(smem_u32(&empty[i]), 1);
    }
    mbar_init(smem_u32(tfull), 1);
    mbar_init(smem_u32(tempty), 2 * NUM_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_async_smem();  // the ones tile is read by the tensor core
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int segs = grp.segs;
  const int total_items = grp.total_tiles * segs;
  // work item -> problem, tile origin, k-block range [kb0, kb1) of the rows-of-dy reduction
  auto locate = [&](int item, int& g, int& m0, int& n0, int& kb0, int& kb1) {
    const int tile = item / segs, seg = item % segs;
    g = 0;
#pragma unroll 1
    for (int i = 1; i < grp.count; ++i)
      if (tile >= grp.p[i].tile_begin) g = i;
    const int t = tile - grp.p[g].tile_begin;
    n0 = (t % grp.p[g].tiles_n) * BN;
    m0 = (t / grp.p[g].tiles_n) * (2 * BM);
    const int num_kb = (grp.p[g].R + BK - 1) / BK;
    kb0 = (int)((long long)seg * num_kb / segs);
    kb1 = (int)((long long)(seg + 1) * num_kb / segs);
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = pair; item < total_items; item += npairs) {
        int g, m0, n0, kb0, kb1;
        locate(item, g, m0, n0, kb0, kb1);
        const Wg2Problem& P = grp.p[g];
        const int ma = m0 + (int)rank * BM, nb = n0 + (int)rank * (BN / 2);
        for (int kb = kb0; kb < kb1; ++kb) {
          MBAR_WAIT_LONG(smem_u32(&empty[stage]), phase ^ 1);
          const uint32_t fb_local = smem_u32(&full[stage]);
          if (rank == 0) mbar_expect_tx(fb_local, 2 * (A_BYTES + B_BYTES));
          const uint32_t fb = mapa(fb_local, 0);
          const uint32_t a_dst = smem_u32(sA + stage * A_BYTES);
          const uint32_t b_dst = smem_u32(sB + stage * B_BYTES);
#pragma unroll
          for (int a = 0; a < BM / 64; ++a) tma_load_2d_2sm(a_dst + a * (BK * 128), &P.tmA, fb, ma + a * 64, kb * BK);
#pragma unroll
          for (int a = 0; a < (BN / 2) / 64; ++a) tma_load_2d_2sm(b_dst + a * (BK * 128), &P.tmB, fb, nb + a * 64, kb * BK);
          if (++stage == ST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc_main = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                      ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
      constexpr uint32_t idesc_bias = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (0u << 16) |
                                      ((uint32_t)(16 >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, tphase = 0;
      for (int item = pair; item < total_items; item += npairs) {
        int g, m0, n0, kb0, kb1;
        locate(item, g, m0, n0, kb0, kb1);
        const Wg2Problem& P = grp.p[g];
        const bool with_bias = n0 == 0 && P.dbias != nullptr;
        MBAR_WAIT_LONG(smem_u32(tempty), tphase ^ 1);
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(&full[stage]), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = make_smem_desc<true>(a_addr, k);
            tc_mma2_bf16(tmem_base, ad, make_smem_desc<true>(b_addr, k), idesc_main, (kb > kb0 || k > 0) ? 1u : 0u);
            if (with_bias)
              tc_mma2_bf16(tmem_base + BN, ad, make_smem_desc<false>(smem_u32(sOnes), k), idesc_bias,
                           (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit2_mc(smem_u32(&empty[stage]));
          if (++stage == ST) { stage = 0; phase ^= 1; }
        }
        tc_commit2_mc(smem_u32(tfull));
        tphase ^= 1;
      }
    }
  } else {
    const int ew = warp - 2, q = warp & 3, s = ew >> 2;
    uint8_t* ob = sEpi + ew * BOX_BYTES;
    const uint32_t ob_addr = smem_u32(ob);
    const uint32_t tempty_leader = mapa(smem_u32(tempty), 0);
    const int swz = grp.swz;
    uint32_t tphase = 0;
    for (int item = pair; item < total_items; item += npairs) {
      int g, m0, n0, kb0, kb1;
      locate(item, g, m0, n0, kb0, kb1);
      const Wg2Problem& P = grp.p[g];
      const int r0 = m0 + (int)rank * BM + q * 32;
      const int c0 = n0 + s * 64;
      mbar_wait(smem_u32(tfull), tphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
      uint32_t raw[2][16];
      tmem_ld16_nowait(t_row + (uint32_t)(s * 64), raw[0]);
      float bias_v = 0.f;
      const bool do_bias = s == 0 && n0 == 0 && P.dbias != nullptr;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_wait_ld();
        if (c + 1 < 4) {
          tmem_ld16_nowait(t_row + (uint32_t)(s * 64 + (c + 1) * 16), raw[(c + 1) & 1]);
        } else if (do_bias) {
          uint32_t rb[16];
          tmem_ld16_nowait(t_row + (uint32_t)BN, rb);
          tmem_wait_ld();
          bias_v = __uint_as_float(rb[0]);
        }
        if (c == 3) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader);
        }
        const bool live = r0 < P.M && c0 + c * 16 < P.N;
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          *reinterpret_cast<uint4*>(ob + box_off(lane, k, swz)) =
              make_uint4(raw[c & 1][4 * k], raw[c & 1][4 * k + 1], raw[c & 1][4 * k + 2], raw[c & 1][4 * k + 3]);
        fence_async_smem();
        __syncwarp();
        if (lane == 0 && live) {
          tma_reduce_add_2d(&P.tmOut, ob_addr, c0 + c * 16, r0);
          bulk_commit();
        }
      }
      if (do_bias && r0 + lane < P.M) {
        if (segs > 1) atomicAdd(P.dbias + r0 + lane, bias_v);  // several k-segments add into the same entry
        else P.dbias[r0 + lane] += bias_v;
      }
      tphase ^= 1;
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

}  // namespace tc2

// Which products go to the pair kernel: A K-major, a supported epilogue with homogeneous I/O types, TMA-compatible
// output / side-input pitches, and enough 256 x 256 tiles to fill the 74 CTA pairs reasonably.
extern "C" int mma_gemm2_eligible(int a_mn, int b_mn, int M, int N, int K, const Epi* ep, int splits) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("MMA_GEMM2");
    enabled = e ? atoi(e) : 1;
  }
  if (!enabled || a_mn || splits > 1 || !ep) return 0;
  auto ok16 = [](const void* p, long long ld, int f32) {
    return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ((ld * (f32 ? 4 : 2)) & 15) == 0;
  };
  const int k = ep->kind;
  // instantiated: forward STORE (bf16 / fp32), GELU (bf16), RESID (bf16 / fp32); dgrad STORE (bf16), DGELU (bf16), ACCUM
  if (b_mn ? !(k == EPI_STORE || k == EPI_DGELU || k == EPI_ACCUM) : !(k == EPI_STORE || k == EPI_GELU || k == EPI_RESID))
    return 0;
  if (b_mn && (k == EPI_STORE || k == EPI_DGELU) && ep->out_f32) return 0;
  if (k == EPI_STORE) {
    if (!ok16(ep->out, ep->ldo, ep->out_f32)) return 0;
  } else if (k == EPI_GELU) {
    if (ep->out_f32 || !ok16(ep->out, ep->ldo, 0)) return 0;
    if (ep->out2 && !ok16(ep->out2, ep->ldo2, 0)) return 0;
  } else if (k == EPI_RESID) {
    if (ep->out_f32 != ep->resid_f32 || !ok16(ep->out, ep->ldo, ep->out_f32) || !ok16(ep->resid, ep->ldr, ep->resid_f32))
      return 0;
  } else if (k == EPI_DGELU) {
    if (ep->out_f32 != ep->aux_f32 || !ok16(ep->out, ep->ldo, ep->out_f32) || !ok16(ep->aux, ep->lda, ep->aux_f32)) return 0;
  } else if (k == EPI_ACCUM) {
    if (ep->accumulate == 2 || !ok16(ep->out, ep->ldo, 1)) return 0;
  } else {
    return 0;
  }
  if (N < 256 || M < 512) return 0;
  if (ep->bias && ((reinterpret_cast<uintptr_t>(ep->bias) & 15) || (N & 3))) return 0;  // vector bias loads
  static int min_tiles = -1;
  if (min_tiles < 0) {
    const char* e = getenv("MMA_GEMM2_MIN_TILES");
    min_tiles = e ? atoi(e) : 48;
  }
  const long long tiles = (long long)((M + 255) / 256) * ((N + 255) / 256);
  return tiles >= min_tiles;
}

static int gemm2_run(const void* A, long long lda, const void* B, long long ldb, const void* A2, long long lda2,
                     const void* B2, long long ldb2, int b_mn, int M, int N, int K, int K2, const Epi* ep,
                     cudaStream_t stream) {
  using namespace tc2;
  if (M <= 0 || N <= 0 || K <= 0 || K2 < 0 || !ep) return MMA_ERR_ARG;
  static int swz = -1, dbg = 0;
  if (swz < 0) {
    const char* e = getenv("MMA_GEMM2_SWZ");
    swz = e ? atoi(e) : 1;
    const char* d = getenv("MMA_GEMM_DBG");
    dbg = d ? atoi(d) : 0;
  }
  const int flags = (swz ? 1 : 0) | (dbg << 8);
  const int box_swz = swz ? (int)CU_TENSOR_MAP_SWIZZLE_64B : (int)CU_TENSOR_MAP_SWIZZLE_NONE;
  CUtensorMap tmA, tmB, tmOut, tmOut2, tmIn;
  int rc = make_map(&tmA, A, (unsigned long long)K, (unsigned long long)M, lda, BK, BM);
  if (rc) return rc;
  if (!b_mn) rc = make_map(&tmB, B, (unsigned long long)K, (unsigned long long)N, ldb, BK, BN / 2);
  else rc = make_map(&tmB, B, (unsigned long long)N, (unsigned long long)K, ldb, 64, BK);
  if (rc) return rc;
  CUtensorMap tmA2 = tmA, tmB2 = tmB;
  if (K2 > 0) {
    rc = make_map(&tmA2, A2, (unsigned long long)K2, (unsigned long long)M, lda2, BK, BM);
    if (rc) return rc;
    if (!b_mn) rc = make_map(&tmB2, B2, (unsigned long long)K2, (unsigned long long)N, ldb2, BK, BN / 2);
    else rc = make_map(&tmB2, B2, (unsigned long long)N, (unsigned long long)K2, ldb2, 64, BK);
    if (rc) return rc;
  }
  const int kind = ep->kind;
  const int io32 = kind == EPI_ACCUM ? 1 : ep->out_f32;
  const unsigned bc = io32 ? 16 : 32;
  rc = make_map_ex(&tmOut, ep->out, (unsigned long long)N, (unsigned long long)M, ep->ldo, bc, 32, io32, box_swz);
  if (rc) return rc;
  tmOut2 = tmOut;
  tmIn = tmOut;
  if (kind == EPI_GELU && ep->out2) {
    rc = make_map_ex(&tmOut2, ep->out2, (unsigned long long)N, (unsigned long long)M, ep->ldo2, bc, 32, io32, box_swz);
    if (rc) return rc;
  }
  if (kind == EPI_RESID) rc = make_map_ex(&tmIn, ep->resid, (unsigned long long)N, (unsigned long long)M, ep->ldr, bc, 32, io32, box_swz);
  if (kind == EPI_DGELU) rc = make_map_ex(&tmIn, ep->aux, (unsigned long long)N, (unsigned long long)M, ep->lda, bc, 32, io32, box_swz);
  if (rc) return rc;
#define MMA_L2(BMN, KD, IO) \
  return launch<BMN, KD, IO>(tmA, tmB, tmOut, tmOut2, tmIn, tmA2, tmB2, M, N, K, K2, flags, *ep, stream)
  if (!b_mn) {
    if (kind == EPI_STORE && !io32) MMA_L2(false, EPI_STORE, false);
    if (kind == EPI_STORE && io32) MMA_L2(false, EPI_STORE, true);
    if (kind == EPI_GELU && !io32) MMA_L2(false, EPI_GELU, false);
    if (kind == EPI_RESID && io32) MMA_L2(false, EPI_RESID, true);
    if (kind == EPI_RESID && !io32) MMA_L2(false, EPI_RESID, false);
  } else {
    if (kind == EPI_STORE && !io32) MMA_L2(true, EPI_STORE, false);
    if (kind == EPI_DGELU && !io32) MMA_L2(true, EPI_DGELU, false);
    if (kind == EPI_ACCUM) MMA_L2(true, EPI_ACCUM, true);
  }
#undef MMA_L2
  return MMA_ERR_UNSUPPORTED;
}

extern "C" int mma_gemm2_bf16(const void* A, long long lda, const void* B, long long ldb, int b_mn, int M, int N, int K,
                              const Epi* ep, cudaStream_t stream) {
  return gemm2_run(A, lda, B, ldb, nullptr, 0, nullptr, 0, b_mn, M, N, K, 0, ep, stream);
}

// C[M,N] = epi(A1 B1_op^T + A2 B2_op^T): one accumulation over two operand pairs (reduction lengths K1, K2), e.g. the
// gated FFN's dh = dz1 W1 + dz2 Wg.  CTA-pair kernel only: MMA_ERR_UNSUPPORTED when the product is outside its envelope
// (mma_gemm2_eligible) - the caller then runs two products with the accumulate epilogue.
extern "C" int mma_gemm2_dual(const void* A1, long long lda1, const void* B1, long long ldb1, const void* A2,
                              long long lda2, const void* B2, long long ldb2, int b_mn, int M, int N, int K1, int K2,
                              const Epi* ep, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K1 <= 0 || K2 <= 0 || !ep || !A2 || !B2) return MMA_ERR_ARG;
  if (!mma_gemm2_eligible(0, b_mn, M, N, K1 + K2, ep, 1)) return MMA_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(A2) & 15) || ((lda2 * 2) & 15) || (reinterpret_cast<uintptr_t>(B2) & 15) || ((ldb2 * 2) & 15))
    return MMA_ERR_UNSUPPORTED;
  return gemm2_run(A1, lda1, B1, ldb1, A2, lda2, B2, ldb2, b_mn, M, N, K1, K2, ep, stream);
}

// Grouped wgrad (+ bias grad), CTA-pair kernel: same contract as mma_wgrad_group (gemm_tc.cu), which forwards here.
extern "C" int mma_wgrad2_group(int count, const void* const* dy, const long long* lddy, const void* const* x,
                                const long long* ldx, float* const* out, const long long* ldo, float* const* dbias,
                                const int* Nout, const int* Kin, const int* R, cudaStream_t stream) {
  using namespace tc2;
  if (count < 1 || count > 8) return MMA_ERR_ARG;
  static int swz = -1;
  if (swz < 0) {
    const char* e = getenv("MMA_GEMM2_SWZ");
    swz = e ? atoi(e) : 1;
  }
  Wg2Group grp{};
  int tiles = 0;
  for (int g = 0; g < count; ++g) {
    Wg2Problem& P = grp.p[g];
    if (Nout[g] <= 0 || Kin[g] <= 0 || R[g] <= 0) return MMA_ERR_ARG;
    int rc = make_map(&P.tmA, dy[g], (unsigned long long)Nout[g], (unsigned long long)R[g], lddy[g], 64, BK);
    if (rc) return rc;
    rc = make_map(&P.tmB, x[g], (unsigned long long)Kin[g], (unsigned long long)R[g], ldx[g], 64, BK);
    if (rc) return rc;
    rc = make_map_ex(&P.tmOut, out[g], (unsigned long long)Kin[g], (unsigned long long)Nout[g], ldo[g], 16, 32, 1,
                     swz ? (int)CU_TENSOR_MAP_SWIZZLE_64B : (int)CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    P.dbias = dbias ? dbias[g] : nullptr;
    P.M = Nout[g];
    P.N = Kin[g];
    P.R = R[g];
    P.tiles_n = (Kin[g] + BN - 1) / BN;
    P.tile_begin = tiles;
    tiles += ((Nout[g] + 2 * BM - 1) / (2 * BM)) * P.tiles_n;
  }
  grp.count = count;
  grp.total_tiles = tiles;
  grp.swz = swz ? 1 : 0;
  grp.segs = 1;
  static int max_pairs = 0;
  if (!max_pairs) {
    if (cudaFuncSetAttribute(wgrad2_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG2_SMEM) != cudaSuccess)
      return MMA_ERR_LAUNCH;
    cudaLaunchConfig_t q{};
    q.gridDim = dim3(2 * (num_sms() / 2));
    q.blockDim = dim3(NUM_THREADS);
    q.dynamicSmemBytes = WG2_SMEM;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, wgrad2_group_kernel, &q) != cudaSuccess || n <= 0) n = num_sms() / 2;
    max_pairs = n < num_sms() / 2 ? n : num_sms() / 2;
  }
  // Split the reduction (rows of dy) of every tile into `segs` work items when that fills the CTA pairs better: the
  // encoder layers have 48 tiles for 74 pairs, the LM head 2.  Each item adds into the gradient with TMA reduce-add
  // (fp32 add order then depends on the schedule; the unsplit case keeps one writer per tile).  MMA_WGRAD_SPLIT=0: off.
  static int allow_split = -1;
  if (allow_split < 0) {
    const char* e = getenv("MMA_WGRAD_SPLIT");
    allow_split = e ? atoi(e) : 1;
  }
  if (allow_split) {
    int min_kb = 1 << 30;
    for (int g = 0; g < count; ++g) {
      const int kb = (R[g] + BK - 1) / BK;
      if (kb < min_kb) min_kb = kb;
    }
    double best = (double)((tiles + max_pairs - 1) / max_pairs);  // rounds of full-length tiles
    const int cand[] = {2, 3, 4, 6, 8, 12, 16};
    for (int c : cand) {
      if (min_kb / c < 16) break;  // keep every item at least 16 k-blocks long
      const double cost = (double)((tiles * c + max_pairs - 1) / max_pairs) / c + 0.02 * c;  // + epilogue overhead
      if (cost < best - 1e-9) {
        best = cost;
        grp.segs = c;
      }
    }
  }
  const int items = tiles * grp.segs;
  const int pairs = items < max_pairs ? items : max_pairs;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = WG2_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, wgrad2_group_kernel, grp) != cudaSuccess) return MMA_ERR_LAUNCH;
  return MMA_OK;
}

// x_new = resid + drop(A W^T + bias)  (fp32, ep->out)  and  h = LayerNorm(x_new) gamma + beta  (bf16), d_model = 512.
// ep: kind EPI_RESID with fp32 out / resid; everything else as in mma_gemm_bf16.  Returns MMA_ERR_UNSUPPORTED for
// shapes / layouts outside that envelope (the caller then runs the product and the LayerNorm separately).
namespace tc {
int gemm_resid_ln_c4(const void* A, long long lda, const void* W, long long ldw, int M, int K, const Epi& ep,
                     const float* gamma, const float* beta, float eps, void* h, long long ldh, cudaStream_t stream);
}

extern "C" int mma_gemm2_resid_ln(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
                                  const Epi* ep, const float* gamma, const float* beta, float eps, void* h,
                                  long long ldh, cudaStream_t stream) {
  using namespace tc2;
  if (M <= 0 || K <= 0 || !ep || !gamma || !beta || !h) return MMA_ERR_ARG;
  auto ok16 = [](const void* p, long long ld, int esz) {
    return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ((ld * esz) & 15) == 0;
  };
  if (N != LN_N || ep->kind != EPI_RESID || !ep->out_f32 || !ep->resid_f32 || !ok16(ep->out, ep->ldo, 4) ||
      !ok16(ep->resid, ep->ldr, 4) || !ok16(h, ldh, 2) || (ep->p_drop > 0.f && (ep->drop_ld & 1)))
    return MMA_ERR_UNSUPPORTED;
  // a few thousand rows without dropout (the decode step): 128 x 128 tiles in clusters of four along N (gemm_tc.cu) fill
  // the machine where 256-row pairs that own all 512 columns do not
  static int c4_max = -1;
  if (c4_max < 0) {
    const char* e = getenv("MMA_RESID_LN_C4_MAX_ROWS");
    c4_max = e ? atoi(e) : 4736;  // 37 row tiles x 4 = 148 CTAs: one wave
  }
  if (M <= c4_max && ep->p_drop <= 0.f && (lda & 7) == 0 && (ldw & 7) == 0 && (K & 7) == 0 &&
      (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0)
    return tc::gemm_resid_ln_c4(A, lda, W, ldw, M, K, *ep, gamma, beta, eps, h, ldh, stream);
  static int swz = -1;
  if (swz < 0) {
    const char* e = getenv("MMA_GEMM2_SWZ");
    swz = e ? atoi(e) : 1;
  }
  const int box_swz = swz ? (int)CU_TENSOR_MAP_SWIZZLE_64B : (int)CU_TENSOR_MAP_SWIZZLE_NONE;
  CUtensorMap tmA, tmB, tmOut, tmIn, tmH;
  int rc = make_map(&tmA, A, (unsigned long long)K, (unsigned long long)M, lda, BK, BM);
  if (rc) return rc;
  if ((rc = make_map(&tmB, W, (unsigned long long)K, (unsigned long long)N, ldw, BK, 128))) return rc;
  if ((rc = make_map_ex(&tmOut, ep->out, (unsigned long long)N, (unsigned long long)M, ep->ldo, 16, 32, 1, box_swz))) return rc;
  if ((rc = make_map_ex(&tmIn, ep->resid, (unsigned long long)N, (unsigned long long)M, ep->ldr, 16, 32, 1, box_swz))) return rc;
  if ((rc = make_map_ex(&tmH, h, (unsigned long long)N, (unsigned long long)M, ldh, 32, 32, 0, box_swz))) return rc;
  static int max_pairs = 0;
  if (!max_pairs) {
    if (cudaFuncSetAttribute(gemm2_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LN_SMEM) != cudaSuccess)
      return MMA_ERR_LAUNCH;
    cudaLaunchConfig_t q{};
    q.gridDim = dim3(2 * (num_sms() / 2));
    q.blockDim = dim3(NUM_THREADS);
    q.dynamicSmemBytes = LN_SMEM;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm2_ln_kernel, &q) != cudaSuccess || n <= 0) n = num_sms() / 2;
    max_pairs = n < num_sms() / 2 ? n : num_sms() / 2;
  }
  const int total = (M + 2 * BM - 1) / (2 * BM);
  const int pairs = total < max_pairs ? total : max_pairs;
  LnExtra ln{gamma, beta, eps};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = LN_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, gemm2_ln_kernel, tmA, tmB, tmOut, tmIn, tmH, M, K, swz ? 1 : 0, *ep, ln) != cudaSuccess)
    return MMA_ERR_LAUNCH;
  return MMA_OK;
}
