// Tensor-core attention for the bf16 path (head dim 64): flash-style forward and a two-pass backward on
// mma.sync.m16n8k16 (bf16 in, fp32 accumulate), ldmatrix-fed from XOR-swizzled shared-memory tiles.
//
// The sequences on this path are short (S ~ 27-330, T <= 128), so one CTA = 64 query rows (4 warps x 16 rows)
// streaming 64-key tiles; scores, probabilities and their gradients never leave registers.
//   fwd      : S = QK^T -> online softmax -> (dropout) -> O += P V ; saves O and the per-row log-sum-exp
//   bwd dQ   : recompute P; dP = dO V^T; dS = P o (dP o keep - D); dQ = scale * dS K        (CTA per query tile)
//   bwd dKdV : same in the transposed frame (S^T = K Q^T): dV = P_drop^T dO; dK = scale * dS^T Q (CTA per key tile)
// Two passes recompute S twice but need no atomics and are deterministic.  Mask semantics and the dropout
// stream are identical to the SIMT kernels in attention.cu (which remain the fp32 / other-head-dim path).
#include "common.cuh"

namespace amma {

constexpr int DH = 64;
constexpr int BQ = 64;  // rows per CTA (4 warps x 16)
constexpr int BKV = 64; // rows of the streamed tile
constexpr int NT = 128;
constexpr float LOG2E = 1.4426950408889634f;

struct Args {
  const bf16* q; const bf16* k; const bf16* v;
  long long ldq, ldk, ldv;
  const unsigned char* kmask;
  bf16* o; long long ldo;
  float* lse;   // [B, H, Lq]  natural-log LSE of the scaled scores
  float* dsum;  // [B, H, Lq]  D_i = rowsum(dO o O)   (written by the dQ pass, read by the dK/dV pass)
  int B, H, Lq, Lk, causal;
  float scale, p_drop;
  unsigned long long seed; unsigned int site;
  const bf16* dout; long long lddo;
  bf16* dq; bf16* dk; bf16* dv;
  long long lddq, lddk, lddv;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
// smem tile [64 rows][64 bf16]: 128-byte rows, 16-byte chunks XOR-swizzled with the row index (conflict-free ldmatrix)
__device__ __forceinline__ uint32_t tile_off(int r, int c) { return (uint32_t)(r * 128 + ((((c >> 3) ^ (r & 7))) << 4) + ((c & 7) << 1)); }

__device__ __forceinline__ void load_tile(uint8_t* tile, const bf16* g, long long ld, int r0, int nrows) {
  for (int idx = threadIdx.x; idx < 64 * 8; idx += NT) {
    const int r = idx >> 3, ch = idx & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r0 + r < nrows) v = *reinterpret_cast<const uint4*>(g + (long long)(r0 + r) * ld + ch * 8);
    *reinterpret_cast<uint4*>(tile + r * 128 + ((ch ^ (r & 7)) << 4)) = v;
  }
}
// A fragments (16 rows starting at wr0, all 4 k-steps of 16) of a tile
__device__ __forceinline__ void load_afrag(uint32_t (&f)[4][4], const uint8_t* tile, int wr0, int lane) {
  const int i = lane >> 3;
  const int row = wr0 + (i & 1) * 8 + (lane & 7);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(f[ks], smem_u32(tile) + tile_off(row, ks * 16 + (i >> 1) * 8));
}
// acc[16 x 64] += A(16 x 64) * T^T  where the tile T is [64 n][64 k] (k contiguous): "NT" product
__device__ __forceinline__ void gemm_nt(float (&acc)[8][4], const uint32_t (&a)[4][4], const uint8_t* tile, int lane) {
  const int i = lane >> 3;
  const uint32_t base = smem_u32(tile);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4(b, base + tile_off(np * 16 + (i >> 1) * 8 + (lane & 7), ks * 16 + (i & 1) * 8));
      mma16816(acc[2 * np], a[ks], b[0], b[1]);
      mma16816(acc[2 * np + 1], a[ks], b[2], b[3]);
    }
  }
}
// acc[16 x 64] += A(16 x 64) * T  where the tile T is [64 k][64 n] (n contiguous): "NN" product
__device__ __forceinline__ void gemm_nn(float (&acc)[8][4], const uint32_t (&a)[4][4], const uint8_t* tile, int lane) {
  const int i = lane >> 3;
  const uint32_t base = smem_u32(tile);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, base + tile_off(ks * 16 + (i & 1) * 8 + (lane & 7), np * 16 + (i >> 1) * 8));
      mma16816(acc[2 * np], a[ks], b[0], b[1]);
      mma16816(acc[2 * np + 1], a[ks], b[2], b[3]);
    }
  }
}
// accumulator tile (16 x 64 fp32) -> A fragments (bf16) of the next product
__device__ __forceinline__ void acc_to_afrag(uint32_t (&f)[4][4], const float (&x)[8][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    f[kk][0] = pack_bf16(x[2 * kk][0], x[2 * kk][1]);
    f[kk][1] = pack_bf16(x[2 * kk][2], x[2 * kk][3]);
    f[kk][2] = pack_bf16(x[2 * kk + 1][0], x[2 * kk + 1][1]);
    f[kk][3] = pack_bf16(x[2 * kk + 1][2], x[2 * kk + 1][3]);
  }
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(NT) fwd_kernel(Args a) {
  pdl_trigger();
  __shared__ __align__(128) uint8_t sQ[64 * 128];
  __shared__ __align__(128) uint8_t sK[64 * 128];
  __shared__ __align__(128) uint8_t sV[64 * 128];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const bf16* Q = a.q + (long long)b * a.Lq * a.ldq + h * DH;
  const bf16* K = a.k + (long long)b * a.Lk * a.ldk + h * DH;
  const bf16* V = a.v + (long long)b * a.Lk * a.ldv + h * DH;
  const unsigned char* km = a.kmask ? a.kmask + (long long)b * a.Lk : nullptr;
  load_tile(sQ, Q, a.ldq, q0, a.Lq);
  __syncthreads();
  uint32_t qf[4][4];
  load_afrag(qf, sQ, warp * 16, lane);
  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[n][e] = 0.f;
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  const float sl2 = a.scale * LOG2E;
  const bool drop = a.p_drop > 0.f;
  const uint32_t thr = drop ? drop_threshold(a.p_drop) : 0u;
  const float inv_keep = drop ? 1.f / (1.f - a.p_drop) : 1.f;
  const uint32_t dkey = drop_key(a.seed, a.site);
  const int row0 = q0 + warp * 16 + g;  // this thread's rows: row0 and row0 + 8
  const int k_end = a.causal ? min(a.Lk, q0 + BQ) : a.Lk;

  for (int j0 = 0; j0 < k_end; j0 += BKV) {
    __syncthreads();
    load_tile(sK, K, a.ldk, j0, a.Lk);
    load_tile(sV, V, a.ldv, j0, a.Lk);
    __syncthreads();
    float s[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[n][e] = 0.f;
    gemm_nt(s, qf, sK, lane);
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = j0 + n * 8 + 2 * t + (e & 1);
        const int row = row0 + (e >> 1) * 8;
        const bool ok = col < a.Lk && (!km || km[col]) && (!a.causal || col <= row);
        s[n][e] = ok ? s[n][e] * sl2 : -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], s[n][e]);
      }
    }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float mt = quad_max(mx[r]);
      const float mn = fmaxf(m[r], mt);
      corr[r] = (m[r] == -INFINITY) ? 0.f : exp2f(m[r] - mn);
      if (mn == -INFINITY) corr[r] = 1.f;
      m[r] = mn;
      l[r] *= corr[r];
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        const float p = (m[r] == -INFINITY) ? 0.f : exp2f(s[n][e] - m[r]);
        l[r] += p;
        float pd = p;
        if (drop) {
          const int col = j0 + n * 8 + 2 * t + (e & 1);
          const int row = row0 + r * 8;
          const unsigned long long ei =
              (((unsigned long long)b * a.H + h) * a.Lq + row) * (unsigned long long)((a.Lk + 1) & ~1) + col;
          pd *= drop_scale1(dkey, ei, thr, inv_keep);
        }
        s[n][e] = pd;
        o[n][e] *= corr[r];
      }
    }
    uint32_t pf[4][4];
    acc_to_afrag(pf, s);
    gemm_nn(o, pf, sV, lane);
  }
  bf16* O = a.o + (long long)b * a.Lq * a.ldo + h * DH;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const float lt = quad_sum(l[r]);
    const float inv = lt > 0.f ? 1.f / lt : 0.f;
    const int row = row0 + r * 8;
    if (row < a.Lq) {
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        *reinterpret_cast<uint32_t*>(O + (long long)row * a.ldo + n * 8 + 2 * t) =
            pack_bf16(o[n][2 * r] * inv, o[n][2 * r + 1] * inv);
      }
      if (t == 0 && a.lse)
        a.lse[((long long)b * a.H + h) * a.Lq + row] = lt > 0.f ? m[r] / LOG2E + logf(lt) : -INFINITY;
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward: dQ
__global__ void __launch_bounds__(NT) bwd_dq_kernel(Args a) {
  pdl_trigger();
  __shared__ __align__(128) uint8_t sQ[64 * 128];
  __shared__ __align__(128) uint8_t sDO[64 * 128];
  __shared__ __align__(128) uint8_t sK[64 * 128];
  __shared__ __align__(128) uint8_t sV[64 * 128];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const bf16* Q = a.q + (long long)b * a.Lq * a.ldq + h * DH;
  const bf16* K = a.k + (long long)b * a.Lk * a.ldk + h * DH;
  const bf16* V = a.v + (long long)b * a.Lk * a.ldv + h * DH;
  const bf16* O = a.o + (long long)b * a.Lq * a.ldo + h * DH;
  const bf16* DO = a.dout + (long long)b * a.Lq * a.lddo + h * DH;
  const unsigned char* km = a.kmask ? a.kmask + (long long)b * a.Lk : nullptr;
  load_tile(sQ, Q, a.ldq, q0, a.Lq);
  load_tile(sDO, DO, a.lddo, q0, a.Lq);
  __syncthreads();
  uint32_t qf[4][4], dof[4][4];
  load_afrag(qf, sQ, warp * 16, lane);
  load_afrag(dof, sDO, warp * 16, lane);
  const int row0 = q0 + warp * 16 + g;
  // D_i = sum_c dO[i,c] * O[i,c]; each lane of a quad covers 16 of the 64 columns
  float Di[2], lse2[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row0 + r * 8;
    float sacc = 0.f;
    if (row < a.Lq) {
      const uint4* po = reinterpret_cast<const uint4*>(O + (long long)row * a.ldo + t * 16);
      const uint4* pd = reinterpret_cast<const uint4*>(DO + (long long)row * a.lddo + t * 16);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const uint4 x = po[u], y = pd[u];
        const __nv_bfloat162* xb = reinterpret_cast<const __nv_bfloat162*>(&x);
        const __nv_bfloat162* yb = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const float2 xf = __bfloat1622float2(xb[w]), yf = __bfloat1622float2(yb[w]);
          sacc += xf.x * yf.x + xf.y * yf.y;
        }
      }
    }
    Di[r] = quad_sum(sacc);
    lse2[r] = row < a.Lq ? a.lse[((long long)b * a.H + h) * a.Lq + row] * LOG2E : 0.f;
    if (t == 0 && row < a.Lq) a.dsum[((long long)b * a.H + h) * a.Lq + row] = Di[r];
  }
  float dq[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) dq[n][e] = 0.f;
  const float sl2 = a.scale * LOG2E;
  const bool drop = a.p_drop > 0.f;
  const uint32_t thr = drop ? drop_threshold(a.p_drop) : 0u;
  const float inv_keep = drop ? 1.f / (1.f - a.p_drop) : 1.f;
  const uint32_t dkey = drop_key(a.seed, a.site);
  const int k_end = a.causal ? min(a.Lk, q0 + BQ) : a.Lk;

  for (int j0 = 0; j0 < k_end; j0 += BKV) {
    __syncthreads();
    load_tile(sK, K, a.ldk, j0, a.Lk);
    load_tile(sV, V, a.ldv, j0, a.Lk);
    __syncthreads();
    float s[8][4], dp[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) { s[n][e] = 0.f; dp[n][e] = 0.f; }
    gemm_nt(s, qf, sK, lane);
    gemm_nt(dp, dof, sV, lane);
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        const int col = j0 + n * 8 + 2 * t + (e & 1);
        const int row = row0 + r * 8;
        const bool ok = col < a.Lk && (!km || km[col]) && (!a.causal || col <= row) && row < a.Lq;
        const float p = ok ? exp2f(s[n][e] * sl2 - lse2[r]) : 0.f;
        float dpe = dp[n][e];
        if (drop) {
          const unsigned long long ei =
              (((unsigned long long)b * a.H + h) * a.Lq + row) * (unsigned long long)((a.Lk + 1) & ~1) + col;
          dpe *= drop_scale1(dkey, ei, thr, inv_keep);
        }
        s[n][e] = p * (dpe - Di[r]);
      }
    }
    uint32_t dsf[4][4];
    acc_to_afrag(dsf, s);
    gemm_nn(dq, dsf, sK, lane);
  }
  bf16* DQ = a.dq + (long long)b * a.Lq * a.lddq + h * DH;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row0 + r * 8;
    if (row < a.Lq) {
#pragma unroll
      for (int n = 0; n < 8; ++n)
        *reinterpret_cast<uint32_t*>(DQ + (long long)row * a.lddq + n * 8 + 2 * t) =
            pack_bf16(dq[n][2 * r] * a.scale, dq[n][2 * r + 1] * a.scale);
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward: dK, dV
__global__ void __launch_bounds__(NT) bwd_dkv_kernel(Args a) {
  pdl_trigger();
  __shared__ __align__(128) uint8_t sK[64 * 128];
  __shared__ __align__(128) uint8_t sV[64 * 128];
  __shared__ __align__(128) uint8_t sQ[64 * 128];
  __shared__ __align__(128) uint8_t sDO[64 * 128];
  __shared__ float sLse[64];
  __shared__ float sD[64];
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * BQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const bf16* Q = a.q + (long long)b * a.Lq * a.ldq + h * DH;
  const bf16* K = a.k + (long long)b * a.Lk * a.ldk + h * DH;
  const bf16* V = a.v + (long long)b * a.Lk * a.ldv + h * DH;
  const bf16* DO = a.dout + (long long)b * a.Lq * a.lddo + h * DH;
  const unsigned char* km = a.kmask ? a.kmask + (long long)b * a.Lk : nullptr;
  load_tile(sK, K, a.ldk, k0, a.Lk);
  load_tile(sV, V, a.ldv, k0, a.Lk);
  __syncthreads();
  uint32_t kf[4][4], vf[4][4];
  load_afrag(kf, sK, warp * 16, lane);
  load_afrag(vf, sV, warp * 16, lane);
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) { dk[n][e] = 0.f; dv[n][e] = 0.f; }
  const int krow0 = k0 + warp * 16 + g;  // this thread's key rows: krow0, krow0 + 8
  bool kok[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int kr = krow0 + r * 8;
    kok[r] = kr < a.Lk && (!km || km[kr]);
  }
  const float sl2 = a.scale * LOG2E;
  const bool drop = a.p_drop > 0.f;
  const uint32_t thr = drop ? drop_threshold(a.p_drop) : 0u;
  const float inv_keep = drop ? 1.f / (1.f - a.p_drop) : 1.f;
  const uint32_t dkey = drop_key(a.seed, a.site);
  const int i_begin = a.causal ? (k0 / BKV) * BKV : 0;  // earlier queries never see these keys

  for (int i0 = i_begin; i0 < a.Lq; i0 += BKV) {
    __syncthreads();
    load_tile(sQ, Q, a.ldq, i0, a.Lq);
    load_tile(sDO, DO, a.lddo, i0, a.Lq);
    if (threadIdx.x < 64) {
      const int i = i0 + threadIdx.x;
      const long long idx = ((long long)b * a.H + h) * a.Lq + i;
      sLse[threadIdx.x] = i < a.Lq ? a.lse[idx] * LOG2E : 0.f;
      sD[threadIdx.x] = i < a.Lq ? a.dsum[idx] : 0.f;
    }
    __syncthreads();
    float st[8][4], dpt[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) { st[n][e] = 0.f; dpt[n][e] = 0.f; }
    gemm_nt(st, kf, sQ, lane);    // S^T[key, query]
    gemm_nt(dpt, vf, sDO, lane);  // dP^T[key, query]
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        const int qc = n * 8 + 2 * t + (e & 1);  // query column inside the tile
        const int qi = i0 + qc;
        const int kr = krow0 + r * 8;
        const bool ok = kok[r] && qi < a.Lq && (!a.causal || kr <= qi);
        const float p = ok ? exp2f(st[n][e] * sl2 - sLse[qc]) : 0.f;
        float keep = 1.f;
        if (drop) {
          const unsigned long long ei =
              (((unsigned long long)b * a.H + h) * a.Lq + qi) * (unsigned long long)((a.Lk + 1) & ~1) + kr;
          keep = drop_scale1(dkey, ei, thr, inv_keep);
        }
        st[n][e] = p * keep;                              // P_drop^T
        dpt[n][e] = p * (dpt[n][e] * keep - sD[qc]);      // dS^T
      }
    }
    uint32_t pf[4][4];
    acc_to_afrag(pf, st);
    gemm_nn(dv, pf, sDO, lane);
    acc_to_afrag(pf, dpt);
    gemm_nn(dk, pf, sQ, lane);
  }
  bf16* DK = a.dk + (long long)b * a.Lk * a.lddk + h * DH;
  bf16* DV = a.dv + (long long)b * a.Lk * a.lddv + h * DH;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int kr = krow0 + r * 8;
    if (kr < a.Lk) {
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        *reinterpret_cast<uint32_t*>(DK + (long long)kr * a.lddk + n * 8 + 2 * t) =
            pack_bf16(dk[n][2 * r] * a.scale, dk[n][2 * r + 1] * a.scale);
        *reinterpret_cast<uint32_t*>(DV + (long long)kr * a.lddv + n * 8 + 2 * t) =
            pack_bf16(dv[n][2 * r], dv[n][2 * r + 1]);
      }
    }
  }
}

}  // namespace amma

// bf16, head dim 64 only.  Pointers / pitches as in mma_attn_fwd / mma_attn_bwd (attention.cu).
extern "C" int mma_attn_fwd_tc(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                               long long ldv, const unsigned char* kmask, void* o, long long ldo, float* lse, int B,
                               int H, int Lq, int Lk, int causal, float scale, float p_drop, unsigned long long seed,
                               unsigned int site, cudaStream_t stream) {
  if (B <= 0 || Lq <= 0 || Lk <= 0) return MMA_OK;
  if ((ldq | ldk | ldv | ldo) & 7) return MMA_ERR_UNSUPPORTED;
  amma::Args a{};
  a.q = (const bf16*)q; a.k = (const bf16*)k; a.v = (const bf16*)v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.kmask = kmask; a.o = (bf16*)o; a.ldo = ldo; a.lse = lse; a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk;
  a.causal = causal; a.scale = scale; a.p_drop = p_drop; a.seed = seed; a.site = site;
  dim3 grid((Lq + amma::BQ - 1) / amma::BQ, H, B);
  amma::fwd_kernel<<<grid, amma::NT, 0, stream>>>(a);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

// dsum: fp32 workspace [B*H*Lq]
extern "C" int mma_attn_bwd_tc(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                               long long ldv, const unsigned char* kmask, const void* o, long long ldo,
                               const float* lse, float* dsum, const void* dout, long long lddo, void* dq,
                               long long lddq, void* dk, long long lddk, void* dv, long long lddv, int B, int H, int Lq,
                               int Lk, int causal, float scale, float p_drop, unsigned long long seed,
                               unsigned int site, cudaStream_t stream) {
  if (B <= 0 || Lq <= 0 || Lk <= 0) return MMA_OK;
  if ((ldq | ldk | ldv | ldo | lddo | lddq | lddk | lddv) & 7) return MMA_ERR_UNSUPPORTED;
  amma::Args a{};
  a.q = (const bf16*)q; a.k = (const bf16*)k; a.v = (const bf16*)v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.kmask = kmask; a.o = (bf16*)const_cast<void*>(o); a.ldo = ldo; a.lse = const_cast<float*>(lse); a.dsum = dsum;
  a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk; a.causal = causal; a.scale = scale; a.p_drop = p_drop; a.seed = seed;
  a.site = site; a.dout = (const bf16*)dout; a.lddo = lddo; a.dq = (bf16*)dq; a.dk = (bf16*)dk; a.dv = (bf16*)dv;
  a.lddq = lddq; a.lddk = lddk; a.lddv = lddv;
  dim3 g1((Lq + amma::BQ - 1) / amma::BQ, H, B);
  amma::bwd_dq_kernel<<<g1, amma::NT, 0, stream>>>(a);
  MMA_CHECK_LAUNCH();
  dim3 g2((Lk + amma::BQ - 1) / amma::BQ, H, B);
  amma::bwd_dkv_kernel<<<g2, amma::NT, 0, stream>>>(a);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}
