// Memory-bound row kernels: embedding gather / scatter-add, LayerNorm forward / backward (fused with the
// positional-encoding add + multimodal concat on the way out, and with the residual-gradient add +
// dropout-masked low-precision copy on the way back), column sums (bias gradients), casts.
// One warp per row, float4-vectorised lanes, warp-shuffle reductions; fp32 statistics throughout.
#include "common.cuh"
#include "tma.cuh"

namespace rowops {

constexpr int MAXI = 8;  // a lane owns float4 columns lane*4 + 128*i, i < NI <= MAXI  ->  d <= 1024

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ float4 ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void st(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Vec4<bf16> {
  static __device__ __forceinline__ float4 ld(const bf16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
  }
  static __device__ __forceinline__ void st(bf16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

__device__ __forceinline__ float4 ld4_any(const void* p, long long i, int f32) {
  return f32 ? Vec4<float>::ld(reinterpret_cast<const float*>(p) + i) : Vec4<bf16>::ld(reinterpret_cast<const bf16*>(p) + i);
}
__device__ __forceinline__ void st4_any(void* p, long long i, int f32, float4 v) {
  if (f32) Vec4<float>::st(reinterpret_cast<float*>(p) + i, v);
  else Vec4<bf16>::st(reinterpret_cast<bf16*>(p) + i, v);
}

// ---------------------------------------------------------------------------------------------
// gather: out[r,:] = table[ids[r],:] * (scale ? scale[r] : 1)          (nn.Embedding, XVal scaling)
// ---------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const long long* __restrict__ ids, const float* __restrict__ scale,
                                   const float* __restrict__ table, float* __restrict__ out, int rows, int d) {
  pdl_trigger();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* src = table + ids[warp] * (long long)d;
  const float s = scale ? scale[warp] : 1.f;
  float* dst = out + (long long)warp * d;
  if ((d & 3) == 0) {
    for (int c = lane * 4; c < d; c += 128) {
      float4 v = *reinterpret_cast<const float4*>(src + c);
      v.x *= s; v.y *= s; v.z *= s; v.w *= s;
      *reinterpret_cast<float4*>(dst + c) = v;
    }
  } else {
    for (int c = lane; c < d; c += 32) dst[c] = src[c] * s;
  }
}

// dtable[ids[r],:] += g[r,:] * scale[r]   (skipping the padding row, as nn.Embedding(padding_idx) does)
__global__ void scatter_add_rows_kernel(const long long* __restrict__ ids, const float* __restrict__ scale,
                                        const float* __restrict__ g, float* __restrict__ dtable, int rows, int d,
                                        long long pad_idx) {
  pdl_trigger();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const long long id = ids[warp];
  if (id == pad_idx) return;
  const float s = scale ? scale[warp] : 1.f;
  const float* src = g + (long long)warp * d;
  float* dst = dtable + id * (long long)d;
  for (int c = lane; c < d; c += 32) atomicAdd(dst + c, src[c] * s);
}

// ---------------------------------------------------------------------------------------------
// LayerNorm forward.  Output row r lands at (r / group) * out_group_stride + out_offset + r % group, which
// is how each modality writes its slice of the concatenated [B, S_total, d] sequence; `add` (pos-enc
// table, row = out_offset + r % group) is added after the affine.  gamma == nullptr: no norm.
// ---------------------------------------------------------------------------------------------
struct LnFwdArgs {
  const void* x; int x_f32; long long ldx;
  const float* gamma; const float* beta; float eps;
  void* y; int y_f32; long long ldy;
  void* y2; int y2_f32; long long ldy2;
  const float* add; long long ld_add;
  int rows, d, group, out_group_stride, out_offset;
};

template <int NI>
__global__ void __launch_bounds__(256) ln_fwd_kernel(LnFwdArgs a) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.rows) return;
  const long long r = warp;
  const long long orow = (r / a.group) * (long long)a.out_group_stride + a.out_offset + (r % a.group);
  const int d = a.d;
  float4 xv[NI];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < d) {
      xv[i] = ld4_any(a.x, r * a.ldx + c, a.x_f32);
      sum += xv[i].x + xv[i].y + xv[i].z + xv[i].w;
    }
  }
  float mean = 0.f, rstd = 1.f;
  if (a.gamma) {
    mean = warp_sum(sum) / d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int c = lane * 4 + i * 128;
      if (c < d) {
        const float dx = xv[i].x - mean, dy = xv[i].y - mean, dz = xv[i].z - mean, dw = xv[i].w - mean;
        sq += dx * dx + dy * dy + dz * dz + dw * dw;
      }
    }
    rstd = rsqrtf(warp_sum(sq) / d + a.eps);
  }
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < d) {
      float4 o = xv[i];
      if (a.gamma) {
        const float4 g = *reinterpret_cast<const float4*>(a.gamma + c);
        const float4 b = *reinterpret_cast<const float4*>(a.beta + c);
        o.x = (o.x - mean) * rstd * g.x + b.x;
        o.y = (o.y - mean) * rstd * g.y + b.y;
        o.z = (o.z - mean) * rstd * g.z + b.z;
        o.w = (o.w - mean) * rstd * g.w + b.w;
      }
      if (a.add) {
        const float4 p = *reinterpret_cast<const float4*>(a.add + (long long)(a.out_offset + (r % a.group)) * a.ld_add + c);
        o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
      }
      st4_any(a.y, orow * a.ldy + c, a.y_f32, o);
      if (a.y2) st4_any(a.y2, orow * a.ldy2 + c, a.y2_f32, o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward (statistics recomputed from x).
//   g   = dy * gamma;  xhat = (x - mean) * rstd
//   dx  = [dres +] rstd * (g - mean(g) - xhat * mean(g * xhat))
//   dgamma += sum_rows dy * xhat;  dbeta += sum_rows dy        (block-reduced, then atomics)
// dy rows may be remapped exactly like ln_fwd's output rows (embedding backward reads its slice of the
// [B, S_total, d] gradient).  dxb (optional) = low-precision copy of dx with the dropout mask of the
// residual branch that produced x applied: it is the dy operand of that branch's dgrad/wgrad GEMMs.
// ---------------------------------------------------------------------------------------------
struct LnBwdArgs {
  const void* dy; int dy_f32; long long lddy;
  int group, in_group_stride, in_offset;
  const void* x; int x_f32; long long ldx;
  const float* gamma; float eps;
  const float* dres; long long lddres;
  float* dx; long long lddx;
  void* dxb; int dxb_f32; long long lddxb;
  float p_drop; unsigned long long seed; unsigned int site;
  float* dgamma; float* dbeta;
  int rows, d;
};

template <int NI>
__global__ void __launch_bounds__(256, 2) ln_bwd_kernel(LnBwdArgs a) {
  pdl_trigger();
  // warp-private dgamma / dbeta accumulators (plain read-modify-write, no atomics): [warp][2][NI*128]
  extern __shared__ float s_acc[];
  const int d = a.d;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* my_dg = s_acc + (size_t)wib * 2 * NI * 128;
  float* my_db = my_dg + NI * 128;
  for (int c = threadIdx.x; c < wpb * 2 * NI * 128; c += blockDim.x) s_acc[c] = 0.f;
  __syncthreads();
  const bool drop = a.p_drop > 0.f;
  const uint32_t thr = drop ? drop_threshold(a.p_drop) : 0u;
  const float inv_keep = drop ? 1.f / (1.f - a.p_drop) : 1.f;
  const uint32_t dkey = drop_key(a.seed, a.site);
  const bool acc_params = a.gamma && a.dgamma;

  for (long long r = (long long)blockIdx.x * wpb + wib; r < a.rows; r += (long long)gridDim.x * wpb) {
    const long long irow = (r / a.group) * (long long)a.in_group_stride + a.in_offset + (r % a.group);
    // every global load of the row is issued up front (dres may alias dx, so the compiler cannot hoist it itself)
    float4 xv[NI], gv[NI], rv[NI];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int c = lane * 4 + i * 128;
      if (c < d) {
        xv[i] = ld4_any(a.x, r * a.ldx + c, a.x_f32);
        gv[i] = ld4_any(a.dy, irow * a.lddy + c, a.dy_f32);
        rv[i] = a.dres ? *reinterpret_cast<const float4*>(a.dres + r * a.lddres + c) : make_float4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int c = lane * 4 + i * 128;
      if (c < d) sum += xv[i].x + xv[i].y + xv[i].z + xv[i].w;
    }
    float mean = 0.f, rstd = 1.f;
    if (a.gamma) {
      mean = warp_sum(sum) / d;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int c = lane * 4 + i * 128;
        if (c < d) {
          const float e0 = xv[i].x - mean, e1 = xv[i].y - mean, e2 = xv[i].z - mean, e3 = xv[i].w - mean;
          sq += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
        }
      }
      rstd = rsqrtf(warp_sum(sq) / d + a.eps);
    }
    float sg = 0.f, sgx = 0.f;
    if (a.gamma) {
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int c = lane * 4 + i * 128;
        if (c < d) {
          const float4 gm = *reinterpret_cast<const float4*>(a.gamma + c);
          float4 xh;
          xh.x = (xv[i].x - mean) * rstd; xh.y = (xv[i].y - mean) * rstd;
          xh.z = (xv[i].z - mean) * rstd; xh.w = (xv[i].w - mean) * rstd;
          const float4 dy = gv[i];
          if (acc_params) {
            float4 pg = *reinterpret_cast<float4*>(my_dg + c), pb = *reinterpret_cast<float4*>(my_db + c);
            pg.x += dy.x * xh.x; pg.y += dy.y * xh.y; pg.z += dy.z * xh.z; pg.w += dy.w * xh.w;
            pb.x += dy.x; pb.y += dy.y; pb.z += dy.z; pb.w += dy.w;
            *reinterpret_cast<float4*>(my_dg + c) = pg;
            *reinterpret_cast<float4*>(my_db + c) = pb;
          }
          float4 g;
          g.x = dy.x * gm.x; g.y = dy.y * gm.y; g.z = dy.z * gm.z; g.w = dy.w * gm.w;
          sg += g.x + g.y + g.z + g.w;
          sgx += g.x * xh.x + g.y * xh.y + g.z * xh.z + g.w * xh.w;
          gv[i] = g;
          xv[i] = xh;
        }
      }
      sg = warp_sum(sg) / d;
      sgx = warp_sum(sgx) / d;
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int c = lane * 4 + i * 128;
      if (c < d) {
        float4 o = gv[i];
        if (a.gamma) {
          o.x = rstd * (gv[i].x - sg - xv[i].x * sgx);
          o.y = rstd * (gv[i].y - sg - xv[i].y * sgx);
          o.z = rstd * (gv[i].z - sg - xv[i].z * sgx);
          o.w = rstd * (gv[i].w - sg - xv[i].w * sgx);
        }
        o.x += rv[i].x; o.y += rv[i].y; o.z += rv[i].z; o.w += rv[i].w;
        if (a.dx) *reinterpret_cast<float4*>(a.dx + r * a.lddx + c) = o;
        if (a.dxb) {
          if (drop) {
            const unsigned long long e = (unsigned long long)r * d + c;
            if ((e & 1ull) == 0) {  // one hash per pair of neighbouring elements (the usual case: d and c are even)
              const uint32_t e32 = (uint32_t)e;
              const uint32_t r0 = drop_pair(dkey, e32 >> 1), r1 = drop_pair(dkey, (e32 >> 1) + 1u);
              o.x *= (r0 & 0xFFFFu) >= thr ? inv_keep : 0.f;
              o.y *= (r0 >> 16) >= thr ? inv_keep : 0.f;
              o.z *= (r1 & 0xFFFFu) >= thr ? inv_keep : 0.f;
              o.w *= (r1 >> 16) >= thr ? inv_keep : 0.f;
            } else {
              o.x *= drop_scale1(dkey, e, thr, inv_keep);
              o.y *= drop_scale1(dkey, e + 1, thr, inv_keep);
              o.z *= drop_scale1(dkey, e + 2, thr, inv_keep);
              o.w *= drop_scale1(dkey, e + 3, thr, inv_keep);
            }
          }
          st4_any(a.dxb, r * a.lddxb + c, a.dxb_f32, o);
        }
      }
    }
  }
  if (acc_params) {
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
      float sg = 0.f, sb = 0.f;
      for (int w = 0; w < wpb; ++w) {
        sg += s_acc[(size_t)w * 2 * NI * 128 + c];
        sb += s_acc[(size_t)w * 2 * NI * 128 + NI * 128 + c];
      }
      atomicAdd(a.dgamma + c, sg);
      atomicAdd(a.dbeta + c, sb);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward, prefetching variant (same math and outputs as ln_bwd_kernel).
// ncu on the plain kernel: 44 % of the stall samples are long-scoreboard waits on the row's first use - a warp
// issues the loads of ONE row, waits for DRAM, computes, stores, and only then asks for its next row; with 16 warps
// per SM that leaves the memory system idle most of the time (31 % of DRAM peak).  Here every warp owns two
// shared-memory stages: while row i is being reduced, the x / dy / dres rows of row i + stride are already in flight
// as 1-D bulk copies (cp.async.bulk -> mbarrier complete_tx), issued by lane 0.  dgamma / dbeta partial sums live in
// registers instead of a read-modify-write through shared memory.  Requires d == NI * 128 and 16-byte aligned rows.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}

template <int NI>
__global__ void __launch_bounds__(256, 2) ln_bwd_pipe_kernel(LnBwdArgs a) {
  pdl_trigger();
  constexpr int D = NI * 128;
  constexpr int STAGE_FLOATS = 3 * D;           // x | dy | dres, each sized for fp32
  extern __shared__ __align__(128) float s_buf[];  // [warp][2 stages][3 * D] then the mbarriers
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* my = s_buf + (size_t)wib * 2 * STAGE_FLOATS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_buf + (size_t)wpb * 2 * STAGE_FLOATS);
  const uint32_t bar0 = tma::smem_u32(bars + 2 * wib);
  if (lane == 0) {
    tma::mbar_init(bar0, 1);
    tma::mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const bool drop = a.p_drop > 0.f;
  const uint32_t thr = drop ? drop_threshold(a.p_drop) : 0u;
  const float inv_keep = drop ? 1.f / (1.f - a.p_drop) : 1.f;
  const uint32_t dkey = drop_key(a.seed, a.site);
  const bool acc_params = a.gamma && a.dgamma;
  const uint32_t xb = (uint32_t)D * (a.x_f32 ? 4u : 2u), yb = (uint32_t)D * (a.dy_f32 ? 4u : 2u);
  const uint32_t rb = a.dres ? (uint32_t)D * 4u : 0u;
  const long long stride = (long long)gridDim.x * wpb;

  auto issue = [&](long long r, int st) {
    const long long irow = (r / a.group) * (long long)a.in_group_stride + a.in_offset + (r % a.group);
    const uint32_t bar = bar0 + 8u * (uint32_t)st;
    float* base = my + (size_t)st * STAGE_FLOATS;
    tma::mbar_expect_tx(bar, xb + yb + rb);
    bulk_load_1d(tma::smem_u32(base), reinterpret_cast<const char*>(a.x) + (r * a.ldx) * (a.x_f32 ? 4 : 2), xb, bar);
    bulk_load_1d(tma::smem_u32(base + D), reinterpret_cast<const char*>(a.dy) + (irow * a.lddy) * (a.dy_f32 ? 4 : 2), yb, bar);
    if (rb) bulk_load_1d(tma::smem_u32(base + 2 * D), a.dres + r * a.lddres, rb, bar);
  };

  float4 ag[NI], ab[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  long long r = (long long)blockIdx.x * wpb + wib;
  if (r < a.rows && lane == 0) issue(r, 0);
  int k = 0;
  for (; r < a.rows; r += stride, ++k) {
    const int st = k & 1;
    if (lane == 0 && r + stride < a.rows) issue(r + stride, st ^ 1);  // that stage was drained one iteration ago
    tma::mbar_wait(bar0 + 8u * (uint32_t)st, (uint32_t)(k >> 1) & 1u);
    const float* sx = my + (size_t)st * STAGE_FLOATS;
    const float* sy = sx + D;
    const float* sr = sx + 2 * D;
    float4 xv[NI], gv[NI];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int c = lane * 4 + i * 128;
      xv[i] = ld4_any(sx, c, a.x_f32);
      gv[i] = ld4_any(sy, c, a.dy_f32);
      sum += xv[i].x + xv[i].y + xv[i].z + xv[i].w;
    }
    float mean = 0.f, rstd = 1.f;
    if (a.gamma) {
      mean = warp_sum(sum) / D;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const float e0 = xv[i].x - mean, e1 = xv[i].y - mean, e2 = xv[i].z - mean, e3 = xv[i].w - mean;
        sq += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
      }
      rstd = rsqrtf(warp_sum(sq) / D + a.eps);
    }
    float sg = 0.f, sgx = 0.f;
    if (a.gamma) {
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int c = lane * 4 + i * 128;
        const float4 gm = *reinterpret_cast<const float4*>(a.gamma + c);
        float4 xh;
        xh.x = (xv[i].x - mean) * rstd; xh.y = (xv[i].y - mean) * rstd;
        xh.z = (xv[i].z - mean) * rstd; xh.w = (xv[i].w - mean) * rstd;
        const float4 dy = gv[i];
        if (acc_params) {
          ag[i].x += dy.x * xh.x; ag[i].y += dy.y * xh.y; ag[i].z += dy.z * xh.z; ag[i].w += dy.w * xh.w;
          ab[i].x += dy.x; ab[i].y += dy.y; ab[i].z += dy.z; ab[i].w += dy.w;
        }
        float4 g;
        g.x = dy.x * gm.x; g.y = dy.y * gm.y; g.z = dy.z * gm.z; g.w = dy.w * gm.w;
        sg += g.x + g.y + g.z + g.w;
        sgx += g.x * xh.x + g.y * xh.y + g.z * xh.z + g.w * xh.w;
        gv[i] = g;
        xv[i] = xh;
      }
      sg = warp_sum(sg) / D;
      sgx = warp_sum(sgx) / D;
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int c = lane * 4 + i * 128;
      float4 o = gv[i];
      if (a.gamma) {
        o.x = rstd * (gv[i].x - sg - xv[i].x * sgx);
        o.y = rstd * (gv[i].y - sg - xv[i].y * sgx);
        o.z = rstd * (gv[i].z - sg - xv[i].z * sgx);
        o.w = rstd * (gv[i].w - sg - xv[i].w * sgx);
      }
      if (rb) {
        const float4 rv = *reinterpret_cast<const float4*>(sr + c);
        o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
      }
      if (a.dx) *reinterpret_cast<float4*>(a.dx + r * a.lddx + c) = o;
      if (a.dxb) {
        if (drop) {
          const uint32_t e32 = (uint32_t)((unsigned long long)r * D + c);  // D and c are multiples of 4: pair-aligned
          const uint32_t r0 = drop_pair(dkey, e32 >> 1), r1 = drop_pair(dkey, (e32 >> 1) + 1u);
          o.x *= (r0 & 0xFFFFu) >= thr ? inv_keep : 0.f;
          o.y *= (r0 >> 16) >= thr ? inv_keep : 0.f;
          o.z *= (r1 & 0xFFFFu) >= thr ? inv_keep : 0.f;
          o.w *= (r1 >> 16) >= thr ? inv_keep : 0.f;
        }
        st4_any(a.dxb, r * a.lddxb + c, a.dxb_f32, o);
      }
    }
    __syncwarp();  // every lane is done reading this stage before lane 0 re-arms it next iteration
  }
  if (acc_params) {
    // block reduction of the per-warp register partials through the (now idle) stage buffers
    __syncthreads();
    float* red = s_buf + (size_t)wib * 2 * D;  // [warp][dgamma D | dbeta D], inside this warp's own stage memory
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int c = lane * 4 + i * 128;
      *reinterpret_cast<float4*>(my + c) = ag[i];
      *reinterpret_cast<float4*>(my + D + c) = ab[i];
    }
    (void)red;
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float sg = 0.f, sb = 0.f;
      for (int w = 0; w < wpb; ++w) {
        sg += s_buf[(size_t)w * 2 * STAGE_FLOATS + c];
        sb += s_buf[(size_t)w * 2 * STAGE_FLOATS + D + c];
      }
      atomicAdd(a.dgamma + c, sg);
      atomicAdd(a.dbeta + c, sb);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// column sums: out[c] += sum_r in[r, c]      (bias gradients; batch-sum of the learned pos-enc gradient)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ in, long long ld, float* __restrict__ out,
                                                     int rows, int cols, int rows_per_block) {
  pdl_trigger();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  if (c < cols)
    for (int r = r0 + ty; r < r1; r += 8) s += to_f(in[(long long)r * ld + c]);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0) {
#pragma unroll
    for (int j = 1; j < 8; ++j) s += red[j][tx];
    if (c < cols) atomicAdd(out + c, s);
  }
}

// fp32 -> bf16 copy (weights, inputs)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
  pdl_trigger();
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    Vec4<bf16>::st(out + i, *reinterpret_cast<const float4*>(in + i));
  } else {
    for (long long j = i; j < n; ++j) out[j] = __float2bfloat16_rn(in[j]);
  }
}
__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out, long long n) {
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(in[i]);
}

}  // namespace rowops

using namespace rowops;

extern "C" int mma_gather_rows(const long long* ids, const float* scale, const float* table, float* out, int rows,
                               int d, cudaStream_t stream) {
  if (rows <= 0) return MMA_OK;
  gather_rows_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(ids, scale, table, out, rows, d);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

extern "C" int mma_scatter_add_rows(const long long* ids, const float* scale, const float* g, float* dtable, int rows,
                                    int d, long long pad_idx, cudaStream_t stream) {
  if (rows <= 0) return MMA_OK;
  scatter_add_rows_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(ids, scale, g, dtable, rows, d, pad_idx);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

extern "C" int mma_ln_fwd(const void* x, int x_f32, long long ldx, const float* gamma, const float* beta, float eps,
                          void* y, int y_f32, long long ldy, void* y2, int y2_f32, long long ldy2, const float* add,
                          long long ld_add, int rows, int d, int group, int out_group_stride, int out_offset,
                          cudaStream_t stream) {
  if (rows <= 0) return MMA_OK;
  if (d > MAXI * 128 || (d & 3)) return MMA_ERR_UNSUPPORTED;
  LnFwdArgs a{x, x_f32, ldx, gamma, beta, eps, y, y_f32, ldy, y2, y2_f32, ldy2, add, ld_add,
              rows, d, group > 0 ? group : rows, out_group_stride, out_offset};
  const int ni = (d + 127) / 128;
  const int blocks = (rows + 7) / 8;
  if (ni <= 1) launch_rowop(ln_fwd_kernel<1>, dim3(blocks), dim3(256), 0, stream, a);
  else if (ni <= 2) launch_rowop(ln_fwd_kernel<2>, dim3(blocks), dim3(256), 0, stream, a);
  else if (ni <= 4) launch_rowop(ln_fwd_kernel<4>, dim3(blocks), dim3(256), 0, stream, a);
  else if (ni <= 6) launch_rowop(ln_fwd_kernel<6>, dim3(blocks), dim3(256), 0, stream, a);
  else launch_rowop(ln_fwd_kernel<8>, dim3(blocks), dim3(256), 0, stream, a);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

extern "C" int mma_ln_bwd(const void* dy, int dy_f32, long long lddy, int group, int in_group_stride, int in_offset,
                          const void* x, int x_f32, long long ldx, const float* gamma, float eps, const float* dres,
                          long long lddres, float* dx, long long lddx, void* dxb, int dxb_f32, long long lddxb,
                          float p_drop, unsigned long long seed, unsigned int site, float* dgamma, float* dbeta,
                          int rows, int d, cudaStream_t stream) {
  if (rows <= 0) return MMA_OK;
  if (d > MAXI * 128 || (d & 3)) return MMA_ERR_UNSUPPORTED;
  LnBwdArgs a{dy, dy_f32, lddy, group > 0 ? group : rows, in_group_stride, in_offset, x, x_f32, ldx, gamma, eps,
              dres, lddres, dx, lddx, dxb, dxb_f32, lddxb, p_drop, seed, site, dgamma, dbeta, rows, d};
  const int ni = (d + 127) / 128;
  int blocks = (rows + 31) / 32;  // >= 4 rows per warp: the dgamma/dbeta flush is amortised
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (blocks < 1) blocks = 1;
  // prefetching variant: full 128-column groups, every row 16-byte aligned, enough rows to fill the pipeline
  {
    static int use_pipe = -1;
    if (use_pipe < 0) {
      const char* e = getenv("MMA_LN_BWD_PIPE");
      use_pipe = e ? atoi(e) : 1;
    }
    const long long xe = x_f32 ? 4 : 2, ye = dy_f32 ? 4 : 2;
    const bool aligned = ((uintptr_t)x % 16 == 0) && ((ldx * xe) % 16 == 0) && ((uintptr_t)dy % 16 == 0) &&
                         ((lddy * ye) % 16 == 0) && (!dres || (((uintptr_t)dres % 16 == 0) && ((lddres * 4) % 16 == 0)));
    // d = 512 only: two 96 KB blocks per SM; wider rows would drop to one block per SM (and NI = 8 spills)
    if (use_pipe && d == 512 && aligned && rows >= 1024) {
      const size_t smem = sizeof(float) * 8 * 2 * 3 * d + 8 * 2 * sizeof(uint64_t);
#define LN_BWD_PIPE(NI_)                                                                                     \
  do {                                                                                                       \
    static bool attr = false;                                                                                \
    if (!attr) {                                                                                             \
      cudaFuncSetAttribute(ln_bwd_pipe_kernel<NI_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      attr = true;                                                                                           \
    }                                                                                                        \
    ln_bwd_pipe_kernel<NI_><<<blocks, 256, smem, stream>>>(a);                                               \
  } while (0)
      LN_BWD_PIPE(4);
#undef LN_BWD_PIPE
      MMA_CHECK_LAUNCH();
      return MMA_OK;
    }
  }
#define LN_BWD_LAUNCH(NI_)                                                                                   \
  do {                                                                                                       \
    const size_t smem = sizeof(float) * 8 * 2 * (NI_) * 128;                                                 \
    static bool attr = false;                                                                                \
    if (!attr && smem > 48 * 1024) {                                                                         \
      cudaFuncSetAttribute(ln_bwd_kernel<NI_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
      attr = true;                                                                                           \
    }                                                                                                        \
    ln_bwd_kernel<NI_><<<blocks, 256, smem, stream>>>(a);                                                    \
  } while (0)
  if (ni <= 1) LN_BWD_LAUNCH(1);
  else if (ni <= 2) LN_BWD_LAUNCH(2);
  else if (ni <= 4) LN_BWD_LAUNCH(4);
  else if (ni <= 6) LN_BWD_LAUNCH(6);
  else LN_BWD_LAUNCH(8);
#undef LN_BWD_LAUNCH
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

extern "C" int mma_colsum(const void* in, int in_f32, long long ld, float* out, int rows, int cols,
                          cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return MMA_OK;
  int by = (rows + 255) / 256;
  if (by > 64) by = 64;
  const int rpb = (rows + by - 1) / by;
  dim3 grid((cols + 31) / 32, by);
  if (in_f32) colsum_kernel<float><<<grid, 256, 0, stream>>>((const float*)in, ld, out, rows, cols, rpb);
  else colsum_kernel<bf16><<<grid, 256, 0, stream>>>((const bf16*)in, ld, out, rows, cols, rpb);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

extern "C" int mma_cast_f32_bf16(const float* in, void* out, long long n, cudaStream_t stream) {
  if (n <= 0) return MMA_OK;
  const long long thr = (n + 3) / 4;
  cast_f32_bf16_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, stream>>>(in, (bf16*)out, n);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

extern "C" int mma_cast_bf16_f32(const void* in, float* out, long long n, cudaStream_t stream) {
  if (n <= 0) return MMA_OK;
  cast_bf16_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const bf16*)in, out, n);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

// ---------------------------------------------------------------------------------------------
// PatchPreprocessor on the device (reference: data/preprocessing/patches.py:54-107).  The reference's optional
// interpolation maps the 400..3980 cm^-1 grid (step 2) onto 650..3898 (step 2): every new abscissa IS an old knot, so
// the linear interpolation is the slice [125, 125 + 1625) of the input - `offset` / `n_use` express that.
//   out[b, p, k] = (raw[b, offset + p * hop + k] - mean) / std          p < P, k < ps,  hop = ps / overlap
//   pad[b, p]    = (sum_k out[b, p, k] == 0)   when `masking`, else the caller's per-sample "spectrum missing" flag
// Output is batch-first [B, P, ps] - the layout the embedding GEMM consumes - so the seq-first host tensor and its
// transpose never exist.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) patchify_kernel(const float* __restrict__ raw, long long ld, int offset,
                                                       float mean, float std, float* __restrict__ out,
                                                       unsigned char* __restrict__ pad,
                                                       const unsigned char* __restrict__ missing, int masking, int P,
                                                       int ps, int hop, const int* __restrict__ rows, int Ptot, int p0,
                                                       int deriv, int n_use) {
  pdl_trigger();
  const int b = blockIdx.y, p = blockIdx.x;
  const long long r = rows ? rows[b] : b;  // dataset-resident spectra: batch element b is row rows[b] of `raw`
  const float* src = raw + r * ld + offset + (long long)p * hop;
  // this launch fills patches [p0, p0 + P) of a sample's Ptot output patches
  float* dst = out + ((long long)b * Ptot + p0 + p) * ps;
  float s = 0.f;
  if (!deriv) {
    for (int k = threadIdx.x; k < ps; k += blockDim.x) {
      const float v = (src[k] - mean) / std;  // exact division: bit-identical to the reference's fp32 arithmetic
      dst[k] = v;
      s += v;
    }
  } else {
    // torch.gradient of the raw (not standardised) spectrum, patches.py:91-95: central differences inside the n_use
    // points the preprocessor sees, one-sided at their two ends
    for (int k = threadIdx.x; k < ps; k += blockDim.x) {
      const int i = p * hop + k;
      float v;
      if (i == 0) v = src[k + 1] - src[k];
      else if (i == n_use - 1) v = src[k] - src[k - 1];
      else v = (src[k + 1] - src[k - 1]) * 0.5f;
      dst[k] = v;
      s += v;
    }
  }
  if (pad) {
    __shared__ float red[4];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      const float tot = red[0] + red[1] + red[2] + red[3];
      pad[(long long)b * Ptot + p0 + p] = masking ? (tot == 0.f) : (missing ? missing[r] : 0);
    }
  }
}

extern "C" int mma_patchify(const float* raw, long long ld, int offset, float mean, float std, float* out,
                            unsigned char* pad, const unsigned char* missing, int masking, int B, int P, int ps,
                            int hop, cudaStream_t stream) {
  if (B <= 0 || P <= 0 || ps <= 0 || hop <= 0 || std == 0.f) return MMA_ERR_ARG;
  patchify_kernel<<<dim3(P, B), 128, 0, stream>>>(raw, ld, offset, mean, std, out, pad, missing, masking, P, ps, hop,
                                                  nullptr, P, 0, 0, 0);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

extern "C" int mma_patchify_rows(const float* raw, long long ld, const int* rows, int offset, float mean, float std,
                                 float* out, unsigned char* pad, const unsigned char* missing, int masking, int B,
                                 int P, int ps, int hop, cudaStream_t stream) {
  if (B <= 0 || P <= 0 || ps <= 0 || hop <= 0 || std == 0.f || !rows) return MMA_ERR_ARG;
  patchify_kernel<<<dim3(P, B), 128, 0, stream>>>(raw, ld, offset, mean, std, out, pad, missing, masking, P, ps, hop,
                                                  rows, P, 0, 0, 0);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

// PatchPreprocessor(derivative=True), patches.py:91-95: out [B, P + Pd, ps] = the P standardised patches (hop as above)
// followed by the Pd = n_use / ps patches of torch.gradient(raw spectrum) (not standardised, never overlapping); pad
// [B, P + Pd] by the same rule over all of them.  n_use: the points the preprocessor sees from `offset` on (the whole
// spectrum, or the 1625 of the interpolation slice) - the one-sided differences sit at its ends.  rows may be NULL.
extern "C" int mma_patchify_deriv(const float* raw, long long ld, const int* rows, int offset, int n_use, float mean,
                                  float std, float* out, unsigned char* pad, const unsigned char* missing, int masking,
                                  int B, int P, int Pd, int ps, int hop, cudaStream_t stream) {
  if (B <= 0 || P <= 0 || Pd <= 0 || ps <= 0 || hop <= 0 || std == 0.f || n_use < 2 || Pd * ps > n_use) return MMA_ERR_ARG;
  patchify_kernel<<<dim3(P, B), 128, 0, stream>>>(raw, ld, offset, mean, std, out, pad, missing, masking, P, ps, hop,
                                                  rows, P + Pd, 0, 0, n_use);
  patchify_kernel<<<dim3(Pd, B), 128, 0, stream>>>(raw, ld, offset, mean, std, out, pad, missing, masking, Pd, ps, ps,
                                                   rows, P + Pd, P, 1, n_use);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}
