// The whole decoder step of a few spectra in ONE launch.
//
// A KV-cached decode step (custom_modeling.py:155-199 per layer, :418-486 around it) is ~50 strictly dependent small
// operations; with <= 64 rows (spectra x beams) their cost is latency, not math: 52 launches ~ 0.35 ms per step even as
// fused LayerNorm + product kernels (decode_small.cu) replayed from a CUDA graph.  Here a thread-block CLUSTER owns a group
// of <= 16 rows (one spectrum's beams, or a few greedy spectra) and walks the entire step - embedding, 6 x (LayerNorm + QKV,
// cached self-attention, out-projection + residual, LayerNorm + cross-query, cross-attention, out-projection + residual,
// LayerNorm + FFN-1 (+ gate), FFN-2 + residual), final LayerNorm + LM head - with a hardware cluster barrier between
// dependent phases instead of a kernel boundary.  Clusters are independent (rows of different spectra never interact inside
// a step), so B spectra run as B clusters side by side and there is no grid-wide synchronisation at all.
//
// What a phase boundary costs decides everything at this size (measured with %globaltimer stamps: a first version that
// exchanged activations through L2 spent 5-18 us per phase, 470 us per step - every dependent L2 round trip is ~1.4 us):
//   exchange   the activations never leave the SMs: every CTA keeps the cluster's residual stream (fp32 [16][512]) and the
//              bf16 operand of the next product in its OWN shared memory, and a producer writes each output element
//              straight into the shared memory of the CTAs that consume it (st.shared::cluster - all of them for the
//              residual stream / attention output / FFN activation, the one attention CTA for q, k, v);
//              barrier.cluster (release / acquire) orders those writes.
//   weights    16-byte loads straight from L2 into registers (the decoder's 44-56 MB of bf16 weights stay L2-resident),
//              ALL loads of a round in flight at once, and the loads of the NEXT phase (or round) are issued right after
//              the MMAs of the current one - they fly during the reduction, the epilogue and the barrier.
//   products   the CTAs of a cluster split the output features in tiles of 16 (tile = rank, rank + size, ...), the 8 warps
//              of a CTA split the reduction; mma.sync.m16n8k16 with the WEIGHT tile as the 16-row operand and the rows as
//              the 8-column operand (every weight element is read once per cluster); LayerNorm is the prologue (every CTA
//              normalises its copy of the residual stream), bias / GELU / gate / residual the epilogue.
//   attention  (row, head) item i lives on warp i / 16 of CTA i % 16: the K / V rows of up to 64 positions (ancestor-indexed cache
//              rows for self-attention, the spectrum's memory for cross-attention) are pulled into shared memory with
//              cp.async BEFORE the barrier that delivers q (they do not depend on it); scores lane = position, weighted sum
//              lane = 2 head dims, online softmax across chunks of 64 positions.
#include "common.cuh"

namespace dstep {

constexpr int MAX_LAYERS = 12;
constexpr int MAXR = 16;  // rows per cluster
constexpr int CS = 16;    // CTAs per cluster (compile-time: every product phase has a fixed number of rounds per CTA)
constexpr int THREADS = 256;
constexpr int WARPS = 8;
constexpr int D = 512, F = 2048, DH = 64, H = 8;
constexpr int ATT_CHUNK = 64;  // key positions staged per pass and warp
constexpr int KS_PITCH = 144;  // bytes per staged key row: 128 + 16, so that lane = row 16-byte reads are conflict-free
constexpr int STG_WARP = ATT_CHUNK * (KS_PITCH + 128) + ATT_CHUNK * 4;  // K rows, V rows, probabilities
constexpr int P_D = D + 32, P_F = F + 32;  // bf16 pitches of the staged operands (rows 64 bytes apart in bank space)
// shared-memory plan (bytes)
constexpr int OFF_X = 0;                             // fp32 [16][512]   residual stream (replicated in every CTA)
constexpr int OFF_ATT = OFF_X + MAXR * D * 4;        // bf16 [16][544]   attention output (replicated)
constexpr int OFF_ACT0 = OFF_ATT + MAXR * P_D * 2;   // bf16 [16][544]   LayerNorm output (local)
constexpr int OFF_SLOT = OFF_ACT0 + MAXR * P_D * 2;  // [8 warps]{q fp32[64], k bf16[64], v bf16[64]}  this CTA's (row, head) items
constexpr int SLOT_BYTES = 512;
constexpr int OFF_ACT1 = OFF_SLOT + WARPS * SLOT_BYTES;  // bf16 [16][2080]  FFN activation (replicated)
constexpr int OFF_RED = OFF_ACT1 + MAXR * P_F * 2;       // fp32 [8 tiles][8 warps][16][16] partial sums
constexpr int RED_BYTES = 8 * WARPS * MAXR * 16 * 4;
constexpr int OFF_STG = OFF_ACT1;  // attention staging aliases ACT1 + RED (both dead during an attention phase)
constexpr int UNION_END =
    (OFF_RED + RED_BYTES) > (OFF_STG + WARPS * STG_WARP) ? (OFF_RED + RED_BYTES) : (OFF_STG + WARPS * STG_WARP);
constexpr int OFF_LN = UNION_END;  // fp32 [3 norms]{gamma[512], beta[512]} of the current layer (cp.async at layer start)
constexpr int SMEM_BYTES = OFF_LN + 3 * 2 * D * 4;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

// mirrors of MmaDecodeLayer / MmaDecodeStep (include/mma_b200.h)
struct Layer {
  const bf16 *w_qkv, *w_so, *w_cq, *w_co, *w_f1, *w_fg, *w_f2;
  const float *b_qkv, *b_so, *b_cq, *b_co, *b_f1, *b_fg, *b_f2;
  const float *n1g, *n1b, *n2g, *n2b, *n3g, *n3b;
  bf16 *kc, *vc;      // self-attention cache of this layer [R][Lmax][d]
  const bf16* kvmem;  // cross-attention K | V of this layer [B * S][2 d]
};
struct Args {
  Layer layer[MAX_LAYERS];
  const int* tok; const float* emb; const float* emb_g; const float* emb_b; const float* pos; const int* cur_len;
  const float *fin_g, *fin_b; const bf16* w_lm; const float* b_lm;
  float *x, *xa, *xb; bf16 *qkv, *att, *q, *a; float* logits;  // only `logits` is used (the rest: per-op path workspaces)
  const int* anc; const unsigned char* enc_mask;
  unsigned long long* dbg_times;  // optional [64]: %globaltimer of cluster 0 at every phase boundary
  long long ldv;
  int layers, R, rows_per_cluster, beams, d, f, H, Lmax, S, V, gated;
  float eps, scale;
};

__device__ __forceinline__ uint32_t ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
// phase boundary, split so that work which does not depend on the peers' results sits between the two halves
__device__ __forceinline__ void c_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void c_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_b32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_b16(uint32_t addr, unsigned short v) {
  asm volatile("st.shared::cluster.b16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
// weight / bias loads as volatile asm: ptxas otherwise sinks a plain __ldg down to its first use (register pressure), which
// turns the prefetch back into an exposed L2 round trip
// (plain .nc: with L1::no_allocate the same step measured 0.407 vs 0.345 ms - the two 64-byte halves of a weight line
// are requested by consecutive loads and only merge into one L2 request when the line may allocate in L1)
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_f32_early(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct Ctx {
  uint8_t* smem;
  int crank, csize, r0, Rc, NT, warp, lane, t;
  int an, ah;  // this warp's attention item: row an of the cluster (-1: none), head ah
  unsigned long long* fine; int* fn;  // DSTEP_FINE: fine-grained stamps
  int anc0[2];  // cache rows holding positions lane, lane + 32 of this warp's row (the same table for every layer)
};

// ---- weights of one round of a product, in registers --------------------------------------------------------------
// TB feature tiles x (NIT x 32) reduction elements of this warp's slice: 2 (4 with the gate) 16-byte loads per tile and
// 32-element block; the k order inside a block is permuted identically for both MMA operands (lane q owns 8 consecutive k)
template <int TB, int NIT, bool GLU>
struct WR {
  uint4 a0[TB][NIT], a1[TB][NIT], g0[GLU ? TB : 1][NIT], g1[GLU ? TB : 1][NIT];
  float bias[TB], bias2[GLU ? TB : 1];
};
template <int TB, int NIT, bool GLU>
__device__ __forceinline__ void w_load(WR<TB, NIT, GLU>& w, const Ctx& c, const bf16* wt, const bf16* wt2, const float* bias,
                                       const float* bias2, int ldw, int N, int tile0) {
  const int g = c.lane >> 2, q = c.lane & 3;
  const int ntiles = (N + 15) >> 4;
  const int koff = c.warp * (NIT * 32) + 8 * q;
  const int fcol = threadIdx.x & 15;  // the output column (inside a tile) this thread finishes in the epilogue
#pragma unroll
  for (int b = 0; b < TB; ++b) {
    const int f0 = min(tile0 + b * CS, ntiles - 1) * 16;  // tiles past the last one re-read it; nothing is stored
    const int fa = min(f0 + g, N - 1), fb = min(f0 + g + 8, N - 1);
    const bf16* wa = wt + (long long)fa * ldw + koff;
    const bf16* wb = wt + (long long)fb * ldw + koff;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      w.a0[b][it] = ldg_stream16(wa + it * 32);
      w.a1[b][it] = ldg_stream16(wb + it * 32);
    }
    w.bias[b] = bias ? ldg_f32_early(bias + min(f0 + fcol, N - 1)) : 0.f;
    if (GLU) {
      const bf16* ga = wt2 + (long long)fa * ldw + koff;
      const bf16* gb = wt2 + (long long)fb * ldw + koff;
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        w.g0[b][it] = ldg_stream16(ga + it * 32);
        w.g1[b][it] = ldg_stream16(gb + it * 32);
      }
      w.bias2[b] = bias2 ? ldg_f32_early(bias2 + min(f0 + fcol, N - 1)) : 0.f;
    }
  }
}

// LayerNorm weights travel to shared memory with cp.async long before they are used (slot j = norm j + 1 of the layer):
// no registers held across phases, no L2 round trip between the barrier and the normalisation
__device__ __forceinline__ void ln_fetch(const Ctx& c, int slot, const float* gamma, const float* beta) {
  if (threadIdx.x < 2 * (D / 4)) {
    const int half = threadIdx.x >= D / 4, i = threadIdx.x - half * (D / 4);
    cp_async16(smem_addr(c.smem + OFF_LN + (slot * 2 + half) * D * 4 + i * 16), (half ? beta : gamma) + i * 4);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
// ACT0 = bf16(LayerNorm(X)) for the cluster's rows (every CTA, from its own copy of the residual stream)
__device__ __forceinline__ void ln_stage(const Ctx& c, int slot, float eps) {
  const float* X = reinterpret_cast<const float*>(c.smem + OFF_X);
  bf16* A0 = reinterpret_cast<bf16*>(c.smem + OFF_ACT0);
  const float* gam = reinterpret_cast<const float*>(c.smem + OFF_LN + slot * 2 * D * 4);
  const float* bet = gam + D;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();  // every thread's share of the LayerNorm weights has landed
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int r = c.warp + rr * WARPS;
    if (r < c.Rc) {
      float4 v[4];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[i] = *reinterpret_cast<const float4*>(X + r * D + c.lane * 4 + i * 128);
        s += v[i].x + v[i].y + v[i].z + v[i].w;
      }
      const float mean = warp_sum(s) * (1.0f / D);
      float qd = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float e0 = v[i].x - mean, e1 = v[i].y - mean, e2 = v[i].z - mean, e3 = v[i].w - mean;
        qd += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
      }
      const float rstd = rsqrtf(warp_sum(qd) * (1.0f / D) + eps);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 gm = *reinterpret_cast<const float4*>(gam + c.lane * 4 + i * 128);
        const float4 bt = *reinterpret_cast<const float4*>(bet + c.lane * 4 + i * 128);
        const float y0 = (v[i].x - mean) * rstd * gm.x + bt.x, y1 = (v[i].y - mean) * rstd * gm.y + bt.y;
        const float y2 = (v[i].z - mean) * rstd * gm.z + bt.z, y3 = (v[i].w - mean) * rstd * gm.w + bt.w;
        __nv_bfloat162 p0 = __floats2bfloat162_rn(y0, y1), p1 = __floats2bfloat162_rn(y2, y3);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&p0);
        u.y = *reinterpret_cast<uint32_t*>(&p1);
        *reinterpret_cast<uint2*>(A0 + r * P_D + c.lane * 4 + i * 128) = u;
      }
    }
  }
  __syncthreads();
}

enum { O_QKV = 0, O_Q = 1, O_X = 2, O_ACT1 = 3, O_LOGITS = 4 };
#ifdef DSTEP_FINE
#define FINE(c) do { if ((c).fine && threadIdx.x == 0 && blockIdx.x == 0 && *(c).fn < 100) (c).fine[(*(c).fn)++] = gtime(); } while (0)
#else
#define FINE(c) do {} while (0)
#endif

// the product a phase hands its weight registers to: loaded right after the last MMAs of the current phase
struct Next {
  const bf16 *wt, *wt2;
  const float *bias, *bias2;
  int ldw, N;
};

// One product phase: rounds of TB feature tiles.  `w` holds the first round's weights on entry (loaded by the previous
// phase); right after the last round's MMAs the weights of the NEXT product phase are requested into `nw`.
//   xs / pitch: the bf16 operand rows in this CTA's shared memory.   OUT selects where an output element goes.
template <int TB, int NIT, bool GLU, int OUT, bool GELU, int ROUNDS, bool NEXT, int TB2, int NIT2, bool GLU2>
__device__ __forceinline__ void lin_rounds(WR<TB, NIT, GLU>& w, const Ctx& c, const Args& a, const bf16* wt, const bf16* wt2,
                                           const float* bias, const float* bias2, int ldw, int N, const bf16* xs, int pitch,
                                           WR<TB2, NIT2, GLU2>& nw, const Next& nx) {
  constexpr int SETS = GLU ? 2 : 1;
  const int g = c.lane >> 2, q = c.lane & 3;
  float* red = reinterpret_cast<float*>(c.smem + OFF_RED);  // [TB][8 warps][SETS][16][16]
  const int ntiles = (N + 15) >> 4;
  const int koff = c.warp * (NIT * 32) + 8 * q;
  const uint32_t sbase = smem_addr(c.smem);
#pragma unroll
  for (int rd = 0; rd < ROUNDS; ++rd) {
    const int tile0 = c.crank + rd * CS * TB;
#pragma unroll
    for (int b = 0; b < TB; ++b) {
      float acc[2][4], acc2[2][4];
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[t][e] = acc2[t][e] = 0.f;
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (t < c.NT) {
            const uint4 Bv = *reinterpret_cast<const uint4*>(xs + (size_t)(t * 8 + g) * pitch + koff + it * 32);
            mma16816(acc[t], w.a0[b][it].x, w.a1[b][it].x, w.a0[b][it].y, w.a1[b][it].y, Bv.x, Bv.y);
            mma16816(acc[t], w.a0[b][it].z, w.a1[b][it].z, w.a0[b][it].w, w.a1[b][it].w, Bv.z, Bv.w);
            if (GLU) {
              mma16816(acc2[t], w.g0[b][it].x, w.g1[b][it].x, w.g0[b][it].y, w.g1[b][it].y, Bv.x, Bv.y);
              mma16816(acc2[t], w.g0[b][it].z, w.g1[b][it].z, w.g0[b][it].w, w.g1[b][it].w, Bv.z, Bv.w);
            }
          }
        }
      }
      float* mine = red + (size_t)((b * WARPS + c.warp) * SETS) * 256;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (t < c.NT) {
          const int n = t * 8 + 2 * q;
          mine[n * 16 + g] = acc[t][0];
          mine[(n + 1) * 16 + g] = acc[t][1];
          mine[n * 16 + g + 8] = acc[t][2];
          mine[(n + 1) * 16 + g + 8] = acc[t][3];
          if (GLU) {
            mine[256 + n * 16 + g] = acc2[t][0];
            mine[256 + (n + 1) * 16 + g] = acc2[t][1];
            mine[256 + n * 16 + g + 8] = acc2[t][2];
            mine[256 + (n + 1) * 16 + g + 8] = acc2[t][3];
          }
        }
      }
    }
    FINE(c);  // after mma + partial sums
    float bs[TB], bs2[GLU ? TB : 1];
#pragma unroll
    for (int b = 0; b < TB; ++b) {
      bs[b] = w.bias[b];
      if (GLU) bs2[b] = w.bias2[b];
    }
    // the weight registers are free: the next round's (or the next phase's) loads fly during reduction + epilogue
    if (rd + 1 < ROUNDS) w_load<TB, NIT, GLU>(w, c, wt, wt2, bias, bias2, ldw, N, tile0 + CS * TB);
    else if (NEXT) w_load<TB2, NIT2, GLU2>(nw, c, nx.wt, nx.wt2, nx.bias, nx.bias2, nx.ldw, nx.N, c.crank);
    FINE(c);  // after issuing the next loads
    __syncthreads();
    FINE(c);  // after sync
    const int n = threadIdx.x >> 4, f = threadIdx.x & 15;  // this thread finishes element (row n, column f) of every tile
    if (n < c.Rc) {
#pragma unroll
      for (int b = 0; b < TB; ++b) {
        const int tile = tile0 + b * CS;
        const int col = tile * 16 + f;
        if (tile < ntiles && col < N) {
          float v = bs[b], v2 = GLU ? bs2[b] : 0.f;
#pragma unroll
          for (int ww = 0; ww < WARPS; ++ww) {
            const float* src = red + (size_t)((b * WARPS + ww) * SETS) * 256 + n * 16 + f;
            v += src[0];
            if (GLU) v2 += src[256];
          }
          if (GLU) v = gelu_t<true>(v) * v2;
          else if (GELU) v = gelu_t<true>(v);
          if (OUT == O_X) {
            // residual stream, own copy first (this thread is the only reader of the old value); the peers' copies below
            float* xp = reinterpret_cast<float*>(c.smem + OFF_X) + n * D + col;
            *xp = v + *xp;
          } else if (OUT == O_ACT1) {
            reinterpret_cast<bf16*>(c.smem + OFF_ACT1)[n * P_F + col] = __float2bfloat16_rn(v);
          } else if (OUT == O_QKV || OUT == O_Q) {
            // item (row n, head h) = n * 8 + h is attended by warp item / 16 of CTA item % 16: the (row, head) pairs of
            // the cluster are spread over ALL its SMs (with 10 beams: 5 warps on each of 16 SMs, not 8 warps on 10)
            const int which = OUT == O_Q ? 0 : col >> 9, cc = col & (D - 1), h = cc >> 6, e = cc & 63;
            const int item = n * H + h;
            const uint32_t slot = (uint32_t)(OFF_SLOT + (item / CS) * SLOT_BYTES);
            const uint32_t dst = mapa(sbase + slot, (uint32_t)(item % CS));
            if (which == 0) st_cluster_f32(dst + (uint32_t)(e * 4), v * a.scale);
            else st_cluster_b16(dst + 256u + (uint32_t)((which - 1) * 128 + e * 2), __bfloat16_as_ushort(__float2bfloat16_rn(v)));
          } else {
            a.logits[(long long)(c.r0 + n) * a.ldv + col] = v;
          }
        }
      }
    }
    __syncthreads();  // `red` is rewritten by the next round / phase
    FINE(c);  // after epilogue
    if (OUT == O_X || OUT == O_ACT1) {
      // the round's [rows x 16-column] blocks go to the 15 peers as 16-byte stores (element-wise remote stores were the
      // single largest cost of a phase: one cluster-network transaction per 2 / 4 bytes)
      constexpr int CPB = OUT == O_X ? 4 : 2;  // 16-byte chunks per (tile, row)
      const int items = TB * c.Rc * CPB * (CS - 1);
      for (int i = threadIdx.x; i < items; i += THREADS) {
        const int per_peer = TB * c.Rc * CPB;  // lanes of a warp: consecutive chunks of one peer's copy
        const int peer = i / per_peer;
        int k = i - peer * per_peer;
        const int ch = k % CPB;
        k /= CPB;
        const int nrow = k % c.Rc, b = k / c.Rc;
        const int tile = tile0 + b * CS;
        if (tile >= ntiles) continue;
        const uint32_t off = OUT == O_X ? (uint32_t)(OFF_X + (nrow * D + tile * 16) * 4 + ch * 16)
                                        : (uint32_t)(OFF_ACT1 + (nrow * P_F + tile * 16) * 2 + ch * 16);
        const uint4 val = *reinterpret_cast<const uint4*>(c.smem + off);
        const uint32_t rk = (uint32_t)(c.crank + 1 + peer) & (CS - 1);
        st_cluster_v4(mapa(sbase + off, rk), val);
      }
    }
    FINE(c);  // after broadcast
  }
}

// ---- attention ---------------------------------------------------------------------------------------------------
// One (row, head) item per warp (Ctx::an / ah).  att_stage: issue the cp.async of one chunk of K / V rows (self: cache rows
// of positions < t through the ancestor table; cross: memory rows) - no wait.
template <bool CROSS>
__device__ __forceinline__ void att_stage(const Ctx& c, const Args& a, const Layer& L, int c0) {
  if (c.an < 0) return;
  const int r = c.r0 + c.an, h = c.ah;
  const int nkeys = CROSS ? a.S : c.t;  // self: position t itself arrives through the slot
  uint8_t* Ks = c.smem + OFF_STG + c.warp * STG_WARP;
  uint8_t* Vs = Ks + ATT_CHUNK * KS_PITCH;
  const int* anc = (!CROSS && a.anc) ? a.anc + ((long long)((c.t + 1) & 1) * a.R + r) * a.Lmax : nullptr;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int p = c.lane + 32 * i, j = c0 + p;
    if (j < nkeys) {
      const bf16 *ksrc, *vsrc;
      if (CROSS) {
        ksrc = L.kvmem + ((long long)(r / a.beams) * a.S + j) * (2 * D) + h * DH;
        vsrc = ksrc + D;
      } else {
        const int src = c0 == 0 ? c.anc0[i] : (anc ? anc[j] : r);
        const long long off = ((long long)src * a.Lmax + j) * D + h * DH;
        ksrc = L.kc + off;
        vsrc = L.vc + off;
      }
      const uint32_t kd = smem_addr(Ks + p * KS_PITCH), vd = smem_addr(Vs + p * 128);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        cp_async16(kd + 16 * k, ksrc + 8 * k);
        cp_async16(vd + 16 * k, vsrc + 8 * k);
      }
    }
  }
}

// chunk 0 is already in flight (att_stage before the barrier)
template <bool CROSS>
__device__ __forceinline__ void att_compute(const Ctx& c, const Args& a, const Layer& L) {
  if (c.an < 0) return;
  const int lane = c.lane, h = c.ah, n = c.an;
  uint8_t* Ks = c.smem + OFF_STG + c.warp * STG_WARP;
  uint8_t* Vs = Ks + ATT_CHUNK * KS_PITCH;
  float* ps = reinterpret_cast<float*>(Vs + ATT_CHUNK * 128);
  const uint32_t sbase = smem_addr(c.smem);
  const int nkeys = CROSS ? a.S : c.t + 1;
  const int r = c.r0 + n;
  const uint8_t* sl = c.smem + OFF_SLOT + c.warp * SLOT_BYTES;
  const float* qs = reinterpret_cast<const float*>(sl);
  const unsigned char* km = (CROSS && a.enc_mask) ? a.enc_mask + (long long)(r / a.beams) * a.S : nullptr;
  if (!CROSS && lane < 16) {  // append this step's K / V to the cache: later steps (and this row's descendants) read them
    const uint4 u = *reinterpret_cast<const uint4*>(sl + 256 + (lane >> 3) * 128 + (lane & 7) * 16);
    bf16* dst = (lane < 8 ? L.kc : L.vc) + ((long long)r * a.Lmax + c.t) * D + h * DH + (lane & 7) * 8;
    *reinterpret_cast<uint4*>(dst) = u;
  }
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  for (int c0 = 0; c0 < nkeys; c0 += ATT_CHUNK) {
    const int nn = min(ATT_CHUNK, nkeys - c0);
    if (c0 > 0) att_stage<CROSS>(c, a, L, c0);
    if (!CROSS && c.t >= c0 && c.t < c0 + ATT_CHUNK && lane < 16) {  // position t: from the slot
      const uint4 u = *reinterpret_cast<const uint4*>(sl + 256 + (lane >> 3) * 128 + (lane & 7) * 16);
      uint8_t* dst = lane < 8 ? Ks + (c.t - c0) * KS_PITCH + lane * 16 : Vs + (c.t - c0) * 128 + (lane & 7) * 16;
      *reinterpret_cast<uint4*>(dst) = u;
    }
    cp_async_wait_all();
    __syncwarp();
    float s[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int p = lane + 32 * i, j = c0 + p;
      const bool valid = p < nn && (!km || km[j]);
      float acc = 0.f;
      if (valid) {
        const uint8_t* kr = Ks + p * KS_PITCH;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint4 u = *reinterpret_cast<const uint4*>(kr + 16 * k);
          const float4 qa = *reinterpret_cast<const float4*>(qs + 8 * k), qb = *reinterpret_cast<const float4*>(qs + 8 * k + 4);
          const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
          const float2 f0 = __bfloat1622float2(hh[0]), f1 = __bfloat1622float2(hh[1]);
          const float2 f2 = __bfloat1622float2(hh[2]), f3 = __bfloat1622float2(hh[3]);
          acc = fmaf(qa.x, f0.x, acc); acc = fmaf(qa.y, f0.y, acc); acc = fmaf(qa.z, f1.x, acc); acc = fmaf(qa.w, f1.y, acc);
          acc = fmaf(qb.x, f2.x, acc); acc = fmaf(qb.y, f2.y, acc); acc = fmaf(qb.z, f3.x, acc); acc = fmaf(qb.w, f3.y, acc);
        }
      }
      s[i] = valid ? acc : -INFINITY;
    }
    const float mc = warp_max(fmaxf(s[0], s[1]));
    const float mn = fmaxf(m, mc);
    float p0 = 0.f, p1 = 0.f, corr = 1.f;
    if (mn != -INFINITY) {
      p0 = s[0] == -INFINITY ? 0.f : __expf(s[0] - mn);
      p1 = s[1] == -INFINITY ? 0.f : __expf(s[1] - mn);
      corr = m == -INFINITY ? 0.f : __expf(m - mn);
    }
    l = l * corr + warp_sum(p0 + p1);
    m = mn;
    ps[lane] = p0;
    ps[lane + 32] = p1;
    __syncwarp();
    o0 *= corr;
    o1 *= corr;
#pragma unroll 8
    for (int p = 0; p < nn; ++p) {
      const float pv = ps[p];
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(Vs + p * 128 + 4 * lane));
      o0 = fmaf(pv, f.x, o0);
      o1 = fmaf(pv, f.y, o1);
    }
    __syncwarp();  // the staging buffers are refilled by the next chunk
  }
  const float inv = l > 0.f ? 1.f / l : 0.f;
  __nv_bfloat162 o = __floats2bfloat162_rn(o0 * inv, o1 * inv);
  const uint32_t off = (uint32_t)(OFF_ATT + (n * P_D + h * DH + 2 * lane) * 2);
  const uint32_t ov = *reinterpret_cast<uint32_t*>(&o);
  for (int rk = 0; rk < CS; ++rk) st_cluster_b32(mapa(sbase + off, rk), ov);
}

__global__ void __launch_bounds__(THREADS, 1) decode_step_kernel(const __grid_constant__ Args a) {
  extern __shared__ __align__(16) uint8_t smem[];
  Ctx c;
  c.smem = smem;
  c.crank = (int)ctarank();
  c.csize = (int)nctarank();
  c.r0 = (int)(blockIdx.x / CS) * a.rows_per_cluster;
  c.Rc = min(a.rows_per_cluster, a.R - c.r0);
  c.NT = (c.Rc + 7) >> 3;
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  if (c.Rc <= 0 || c.csize != CS) return;  // the whole cluster leaves together
  c.t = *a.cur_len - 1;
  {
    const int item = c.warp * CS + c.crank;
    c.an = item < c.Rc * H ? item / H : -1;
    c.ah = item % H;
    const int r = c.r0 + max(c.an, 0);
    const int* anc = (a.anc && c.an >= 0) ? a.anc + ((long long)((c.t + 1) & 1) * a.R + r) * a.Lmax : nullptr;
#pragma unroll
    for (int i = 0; i < 2; ++i) c.anc0[i] = (anc && c.lane + 32 * i < c.t) ? anc[c.lane + 32 * i] : r;
  }
  int fcount = 0;
  c.fine = nullptr;
  c.fn = &fcount;
  int ts = 0;
  const bool stamp = a.dbg_times != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
#define DSTEP_STAMP() do { if (stamp) a.dbg_times[ts++] = gtime(); } while (0)
  DSTEP_STAMP();
  const bf16* ACT0 = reinterpret_cast<const bf16*>(smem + OFF_ACT0);
  const bf16* ATT = reinterpret_cast<const bf16*>(smem + OFF_ATT);
  const bf16* ACT1 = reinterpret_cast<const bf16*>(smem + OFF_ACT1);
  // the first phase's weights fly while the embedding is gathered
  WR<6, 2, false> w_qkv;
  w_load<6, 2, false>(w_qkv, c, a.layer[0].w_qkv, nullptr, a.layer[0].b_qkv, nullptr, D, 3 * D, c.crank);

  // token embedding (+ per-modality LayerNorm) + positional row t -> X, in every CTA (no exchange needed)
  {
    float* X = reinterpret_cast<float*>(smem + OFF_X);
    const float* prow = a.pos + (long long)c.t * D;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = c.warp + rr * WARPS;
      if (r < c.Rc) {
        const float* src = a.emb + (long long)a.tok[c.r0 + r] * D;
        float4 v[4];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[i] = __ldg(reinterpret_cast<const float4*>(src + c.lane * 4 + i * 128));
          s += v[i].x + v[i].y + v[i].z + v[i].w;
        }
        float mean = 0.f, rstd = 1.f;
        if (a.emb_g) {
          mean = warp_sum(s) * (1.0f / D);
          float qd = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float e0 = v[i].x - mean, e1 = v[i].y - mean, e2 = v[i].z - mean, e3 = v[i].w - mean;
            qd += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
          }
          rstd = rsqrtf(warp_sum(qd) * (1.0f / D) + a.eps);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int col = c.lane * 4 + i * 128;
          float4 y = v[i];
          if (a.emb_g) {
            const float4 gm = __ldg(reinterpret_cast<const float4*>(a.emb_g + col));
            const float4 bt = __ldg(reinterpret_cast<const float4*>(a.emb_b + col));
            y.x = (y.x - mean) * rstd * gm.x + bt.x; y.y = (y.y - mean) * rstd * gm.y + bt.y;
            y.z = (y.z - mean) * rstd * gm.z + bt.z; y.w = (y.w - mean) * rstd * gm.w + bt.w;
          }
          const float4 pp = __ldg(reinterpret_cast<const float4*>(prow + col));
          *reinterpret_cast<float4*>(X + r * D + col) = make_float4(y.x + pp.x, y.y + pp.y, y.z + pp.z, y.w + pp.w);
        }
      }
    }
    __syncthreads();
  }
  // nobody may write into a peer's shared memory before that peer has started
  c_arrive();
  c_wait();
  DSTEP_STAMP();

  const Next none{nullptr, nullptr, nullptr, nullptr, 0, 0};
  ln_fetch(c, 0, a.layer[0].n1g, a.layer[0].n1b);
  // Weight loads of the NEXT product phase are issued right after the MMAs of the current one (they fly during its
  // reduction, epilogue and broadcast).  They cannot be moved further ahead: a cluster barrier drains the thread's
  // outstanding loads (measured: loads placed between arrive and wait, or two phases ahead, lengthen that phase by their
  // full latency; requesting the N = 512 products' weights at the START of the preceding phase instead - two register sets
  // in alternation - measured 270 vs 266 us per step, no gain), and the per-SM rate of these 16-byte loads is ~25 GB/s, so a
  // step is bound by weight streaming.
  for (int li = 0; li < a.layers; ++li) {
    const Layer& L = a.layer[li];
    const bool last = li + 1 == a.layers;
#ifdef DSTEP_FINE
    c.fine = (li == 2 && a.dbg_times) ? a.dbg_times + 64 : nullptr;
#endif
    WR<2, 2, false> w_o;  // N = 512 products: 32 tiles, 2 per CTA
    // ---- h = LN1(x); q, k, v = h Wqkv^T + b  ->  the attention CTAs' slots
    ln_stage(c, 0, a.eps); FINE(c);
    ln_fetch(c, 1, L.n2g, L.n2b);
    ln_fetch(c, 2, L.n3g, L.n3b);
    lin_rounds<6, 2, false, O_QKV, false, 1, true>(w_qkv, c, a, L.w_qkv, nullptr, L.b_qkv, nullptr, D, 3 * D, ACT0, P_D,
                                                   w_o, Next{L.w_so, nullptr, L.b_so, nullptr, D, D});
    FINE(c); c_arrive(); FINE(c);
    att_stage<false>(c, a, L, 0);  // the cache rows of the history do not depend on this step's q / k / v
    // slot 0 is free again: the next layer's norm1 (decoder.norm after the last layer) arrives a whole layer early
    ln_fetch(c, 0, last ? a.fin_g : a.layer[li + 1].n1g, last ? a.fin_b : a.layer[li + 1].n1b);
    c_wait(); FINE(c);
    DSTEP_STAMP();
    att_compute<false>(c, a, L);
    FINE(c); c_arrive(); FINE(c);
    c_wait(); FINE(c);
    DSTEP_STAMP();
    // ---- x += att Wo^T + b
    lin_rounds<2, 2, false, O_X, false, 1, true>(w_o, c, a, L.w_so, nullptr, L.b_so, nullptr, D, D, ATT, P_D, w_o,
                                                 Next{L.w_cq, nullptr, L.b_cq, nullptr, D, D});
    FINE(c); c_arrive(); FINE(c);
    c_wait(); FINE(c);
    DSTEP_STAMP();
    // ---- q = LN2(x) Wq^T + b  ->  slots
    ln_stage(c, 1, a.eps); FINE(c);
    lin_rounds<2, 2, false, O_Q, false, 1, true>(w_o, c, a, L.w_cq, nullptr, L.b_cq, nullptr, D, D, ACT0, P_D, w_o,
                                                 Next{L.w_co, nullptr, L.b_co, nullptr, D, D});
    FINE(c); c_arrive(); FINE(c);
    att_stage<true>(c, a, L, 0);
    c_wait(); FINE(c);
    DSTEP_STAMP();
    att_compute<true>(c, a, L);
    FINE(c); c_arrive(); FINE(c);
    c_wait(); FINE(c);
    DSTEP_STAMP();
    // ---- x += att Wo^T + b ;  a = gelu(LN3(x) W1^T + b1) [* (LN3(x) Wg^T + bg)] ;  x += a W2^T + b2
    WR<2, 8, false> w_f2;
    if (a.gated) {
      WR<4, 2, true> w_f1;
      lin_rounds<2, 2, false, O_X, false, 1, true>(w_o, c, a, L.w_co, nullptr, L.b_co, nullptr, D, D, ATT, P_D, w_f1,
                                                   Next{L.w_f1, L.w_fg, L.b_f1, L.b_fg, D, F});
      FINE(c); c_arrive(); FINE(c);
      c_wait(); FINE(c);
      DSTEP_STAMP();
      ln_stage(c, 2, a.eps); FINE(c);
      lin_rounds<4, 2, true, O_ACT1, false, 2, true>(w_f1, c, a, L.w_f1, L.w_fg, L.b_f1, L.b_fg, D, F, ACT0, P_D, w_f2,
                                                     Next{L.w_f2, nullptr, L.b_f2, nullptr, F, D});
    } else {
      WR<8, 2, false> w_f1;
      lin_rounds<2, 2, false, O_X, false, 1, true>(w_o, c, a, L.w_co, nullptr, L.b_co, nullptr, D, D, ATT, P_D, w_f1,
                                                   Next{L.w_f1, nullptr, L.b_f1, nullptr, D, F});
      FINE(c); c_arrive(); FINE(c);
      c_wait(); FINE(c);
      DSTEP_STAMP();
      ln_stage(c, 2, a.eps); FINE(c);
      lin_rounds<8, 2, false, O_ACT1, true, 1, true>(w_f1, c, a, L.w_f1, nullptr, L.b_f1, nullptr, D, F, ACT0, P_D, w_f2,
                                                     Next{L.w_f2, nullptr, L.b_f2, nullptr, F, D});
    }
    FINE(c); c_arrive(); FINE(c);
    c_wait(); FINE(c);
    DSTEP_STAMP();
    // (after the last layer the QKV registers are fetched once more for nothing: an unconditional load keeps them from
    // being live across the whole layer body)
    const Layer& Nx = a.layer[last ? 0 : li + 1];
    lin_rounds<2, 8, false, O_X, false, 1, true>(w_f2, c, a, L.w_f2, nullptr, L.b_f2, nullptr, F, D, ACT1, P_F, w_qkv,
                                                 Next{Nx.w_qkv, nullptr, Nx.b_qkv, nullptr, D, 3 * D});
    FINE(c); c_arrive(); FINE(c);
    c_wait(); FINE(c);
    DSTEP_STAMP();
  }
  // ---- logits = LN(x) Wlm^T + b  (global: the selection kernel reads them).  After the last barrier above no CTA writes
  // into a peer's shared memory any more, so CTAs may finish independently.
  WR<2, 2, false> w_lm;
  w_load<2, 2, false>(w_lm, c, a.w_lm, nullptr, a.b_lm, nullptr, D, a.V, c.crank);
  ln_stage(c, 0, a.eps); FINE(c);
  lin_rounds<2, 2, false, O_LOGITS, false, 1, false>(w_lm, c, a, a.w_lm, nullptr, a.b_lm, nullptr, D, a.V, ACT0, P_D, w_lm, none);
  DSTEP_STAMP();
#undef DSTEP_STAMP
}

}  // namespace dstep

// One decoder step (logits of every row) in one launch; see include/mma_b200.h.
extern "C" int mma_decode_step(const void* args, int cluster_size, cudaStream_t stream) {
  using namespace dstep;
  if (!args) return MMA_ERR_ARG;
  const Args& a = *reinterpret_cast<const Args*>(args);
  if (a.layers < 1 || a.R < 1 || a.rows_per_cluster < 1 || a.beams < 1 || a.V < 1 || a.S < 1) return MMA_ERR_ARG;
  if (a.layers > MAX_LAYERS || a.rows_per_cluster > MAXR || a.rows_per_cluster > 2 * cluster_size || a.d != D || a.f != F ||
      a.H != H || (a.ldv & 3) || cluster_size != CS || a.V > 2 * CS * 16 || (a.rows_per_cluster % a.beams) != 0)
    return MMA_ERR_UNSUPPORTED;
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(decode_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
      return MMA_ERR_LAUNCH;
    if (cudaFuncSetAttribute(decode_step_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
      return MMA_ERR_LAUNCH;
    set = true;
  }
  const int nclusters = (a.R + a.rows_per_cluster - 1) / a.rows_per_cluster;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(nclusters * cluster_size));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster_size;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, decode_step_kernel, a) != cudaSuccess) {
    cudaGetLastError();
    return MMA_ERR_LAUNCH;
  }
  return MMA_OK;
}

// how many clusters of `cluster_size` CTAs of the step kernel can be resident at once (0: that size cannot be scheduled)
extern "C" int mma_decode_step_max_clusters(int cluster_size) {
  using namespace dstep;
  cudaFuncSetAttribute(decode_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  cudaFuncSetAttribute(decode_step_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(cluster_size * 64));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster_size;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, decode_step_kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
