// Shared device helpers for the sm_100a spectra->SMILES kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

typedef __nv_bfloat16 bf16;

// ---- status codes returned across the C ABI (0 = ok) ---------------------------------------
#define MMA_OK 0
#define MMA_ERR_ARG -1
#define MMA_ERR_LAUNCH -2
#define MMA_ERR_UNSUPPORTED -3
#define MMA_ERR_DRIVER -4

#define MMA_CHECK_LAUNCH()                                  \
  do {                                                      \
    cudaError_t e__ = cudaGetLastError();                   \
    if (e__ != cudaSuccess) return MMA_ERR_LAUNCH;          \
  } while (0)

// element type tags used at the C ABI
#define MMA_BF16 0
#define MMA_F32 1

// ---- epilogue kinds shared by the tcgen05 and the SIMT GEMM ----------------------------------
enum EpiKind : int {
  EPI_STORE = 0,    // out = acc*alpha + bias
  EPI_GELU = 1,     // z = acc + bias; out2 = z (optional); out = drop(gelu(z))
  EPI_RESID = 2,    // out = resid + drop(acc + bias)
  EPI_DGELU = 3,    // out = acc * dropmask * gelu'(aux)
  EPI_GLU_MUL = 4,  // z2 = acc + bias; out2 = z2; out = drop(gelu(aux) * z2)
  EPI_DGLU = 5,     // da = acc*dropmask; out = da * aux2 * gelu'(aux); out2 = da * gelu(aux)
  EPI_ACCUM = 6,    // out(f32) (+)= acc*alpha   [accumulate: 0 overwrite, 1 add, 2 atomic add]
  EPI_RELU = 7,     // out = relu(acc + bias)            (patch-embedding MLPs, align head)
  EPI_DRELU = 8,    // out = acc * [aux > 0]
};

struct Epi {
  int kind;
  int out_f32;   // element type of out / out2 (1 = float, 0 = bf16)
  int aux_f32;   // element type of aux / aux2
  int resid_f32; // element type of resid
  void* out;
  void* out2;
  const float* bias;
  const void* resid;
  const void* aux;
  const void* aux2;
  long long ldo, ldo2, ldr, lda, lda2;
  float p_drop;
  float alpha;
  unsigned long long seed;
  unsigned int site;
  int accumulate;
  long long drop_ld;  // logical row width used to index the dropout stream (== N of the fwd GEMM)
};

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// Every kernel of the library calls pdl_trigger() first thing: the next kernel in the stream, if it was launched
// with the programmatic-stream-serialization attribute (the tcgen05 kernels are), may then start its prologue
// (barrier init, TMEM allocation, tensor-map prefetch) while this one is still running.  Such a kernel calls
// pdl_wait() before it touches global memory: it returns once all prerequisite grids have completed and flushed.
// Both are no-ops for plainly launched kernels.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// Row-wise / decode kernels (LayerNorm forward, decode embedding / attention / selection) CAN be launched with the same
// programmatic attribute (MMA_PDL_ROWOPS=1); such a kernel calls pdl_wait() before its first global access.  Measured on
// B200 (beam-10 decode, 256 spectra): 1.003 ms / step with it against 0.958 ms without - the early-resident dependents
// take SM resources from the kernel they wait for - so plain launches are the default.
static inline bool pdl_rowops_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MMA_PDL_ROWOPS");
    on = e ? (atoi(e) != 0) : 0;
  }
  return on != 0;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_rowop(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                       Args... args) {
  if (pdl_rowops_enabled()) return launch_pdl(kern, grid, block, smem, stream, args...);
  kern<<<grid, block, smem, stream>>>(KArgs(args)...);
  return cudaGetLastError();
}

// ---- small numeric helpers ---------------------------------------------------------------------
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ float ld_any(const void* p, long long i, int is_f32) {
  return is_f32 ? reinterpret_cast<const float*>(p)[i] : __bfloat162float(reinterpret_cast<const bf16*>(p)[i]);
}
__device__ __forceinline__ void st_any(void* p, long long i, int is_f32, float v) {
  if (is_f32) reinterpret_cast<float*>(p)[i] = v;
  else reinterpret_cast<bf16*>(p)[i] = __float2bfloat16_rn(v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- dropout masks are a pure function of (seed, site, element index): backward regenerates them instead of
// storing them.  Counter-based 32-bit mixer (lowbias32 finaliser), ~10 integer ops per element - cheap enough
// to live inside GEMM epilogues and attention inner loops. -------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}
// site bit 31 set: `seed` is the DEVICE ADDRESS of the 64-bit seed (a captured CUDA graph advances the seed on the
// device between replays; the launch parameters stay constant).
#define MMA_SITE_SEED_INDIRECT 0x80000000u
__device__ __forceinline__ uint32_t drop_key(unsigned long long seed, unsigned int site) {
  if (site & MMA_SITE_SEED_INDIRECT) {
    seed = *reinterpret_cast<const unsigned long long*>(seed);
    site &= ~MMA_SITE_SEED_INDIRECT;
  }
  return mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x9E3779B9u) ^ (site * 0x85EBCA77u + 0x165667B1u));
}
// one 32-bit hash serves two neighbouring elements (16 random bits each; p is quantised to 1/65536).
// Two rounds of a 32 x 32 -> 64 bit multiply folded by xor (5 instructions; the lowbias32 finaliser this replaces took
// 10, and the hash sits in the ALU-bound GELU / dGELU GEMM epilogues): the first round turns consecutive counters
// into a Weyl-like sequence, the second destroys its regular spacing.
__device__ __forceinline__ uint32_t drop_pair(uint32_t key, uint32_t pair_idx) {
  unsigned long long t = (unsigned long long)(pair_idx ^ key) * 0x9E3779B1ull;
  uint32_t r = (uint32_t)t ^ (uint32_t)(t >> 32);
  t = (unsigned long long)r * 0x85EBCA77ull;
  return (uint32_t)t ^ (uint32_t)(t >> 32);
}
__device__ __forceinline__ uint32_t drop_threshold(float p) {
  const float t = p * 65536.0f + 0.5f;
  return t >= 65535.0f ? 65535u : (uint32_t)t;
}
// keep-scale (0 or 1/(1-p)) of element e (the element index is taken modulo 2^32)
__device__ __forceinline__ float drop_scale1(uint32_t key, unsigned long long e, uint32_t thr, float inv_keep) {
  const uint32_t e32 = (uint32_t)e;
  const uint32_t r = drop_pair(key, e32 >> 1);
  return ((r >> ((e32 & 1u) << 4)) & 0xFFFFu) >= thr ? inv_keep : 0.0f;
}

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// Phi(x) = 0.5 (1 + erf(x / sqrt 2)) and e = exp(-x^2 / 2).  FAST: Abramowitz-Stegun 7.1.26 (|erf error| < 1.5e-7) on
// rcp.approx / ex2.approx - used when the result is rounded to bf16 anyway; the fp32 parity path keeps erff().
template <bool FAST> __device__ __forceinline__ void normal_cdf_exp(float x, float& cdf, float& e) {
  if (FAST) {
    // u = |x| / sqrt(2);  t = 1 / (1 + 0.3275911 u);  erfc(u) ~ t (a1 + t (a2 + ...)) exp(-u^2)   (A&S 7.1.26)
    // constants folded: w = |x| sqrt(log2(e) / 2) so that exp(-u^2) = 2^(-w^2), 0.3275911 u = 0.27279 w, and the
    // polynomial coefficients carry the factor 1/2 of Phi = erfc / 2
    const float w = fabsf(x) * 0.84932180028801904272f;
    const float t = rcp_approx(fmaf(0.27273748f, w, 1.0f));
    float p = fmaf(0.5307027145f, t, -0.7265760135f);
    p = fmaf(p, t, 0.7107068705f);
    p = fmaf(p, t, -0.142248368f);
    p = fmaf(p, t, 0.127414796f);
    e = ex2_approx(-w * w);
    const float half_erfc = p * t * e;
    cdf = x >= 0.f ? 1.0f - half_erfc : half_erfc;
  } else {
    cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    e = expf(-0.5f * x * x);
  }
}
// ---- GELU on one MUFU (bf16 paths) -----------------------------------------------------------------------------
// erf(x / sqrt 2) = tanh(x q(x^2)) with q a minimax quadratic in x^2 (fit against the exact erf GELU over |x| <= 8:
// |gelu error| < 2.6e-5, |gelu' error| < 1.1e-4 before the tanh.approx error of 2^-11 relative) - both an order of
// magnitude under the bf16 rounding of the result.  7 issued instructions for gelu (FMUL FFMA FFMA FMUL MUFU FMUL FFMA)
// against 15 for the Abramowitz-Stegun form above, 11 against 17 for gelu'; the GELU / dGELU / gate epilogues of the pair
// GEMMs are bound by instruction issue (ncu: 62 % issue slots, tensor pipe 29 %).  -DMMA_GELU_TANH=0 restores A&S.
#ifndef MMA_GELU_TANH
#define MMA_GELU_TANH 1
#endif
constexpr float GT_C0 = 0.7975078843613885f, GT_C1 = 0.03700564597780192f, GT_C2 = -0.0003515167826820022f;
__device__ __forceinline__ float tanh_approx(float x) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// hs * x * (1 + erf(x / sqrt 2)): gelu(x) for hs = 0.5; a dropout keep-scale folds into hs (0 or 0.5 / (1 - p))
__device__ __forceinline__ float gelu_tanh_scaled(float x, float hs) {
  const float x2 = x * x;
  const float q = fmaf(fmaf(x2, GT_C2, GT_C1), x2, GT_C0);
  const float t = tanh_approx(x * q);
  const float hx = x * hs;
  return fmaf(hx, t, hx);
}
// hs * (1 + t + x (1 - t^2) u'(x)),  u = x q(x^2): the derivative of the form above (gelu'(x) for hs = 0.5)
__device__ __forceinline__ float dgelu_tanh_scaled(float x, float hs) {
  const float x2 = x * x;
  const float q = fmaf(fmaf(x2, GT_C2, GT_C1), x2, GT_C0);
  const float t = tanh_approx(x * q);
  const float du = fmaf(fmaf(x2, 5.0f * GT_C2, 3.0f * GT_C1), x2, GT_C0);
  const float w = x * fmaf(-t, t, 1.0f);
  const float s = fmaf(w, du, t);
  return fmaf(hs, s, hs);
}
__device__ __forceinline__ void gelu_both_tanh(float x, float& g, float& dg) {
  const float x2 = x * x;
  const float q = fmaf(fmaf(x2, GT_C2, GT_C1), x2, GT_C0);
  const float t = tanh_approx(x * q);
  const float hx = 0.5f * x;
  g = fmaf(hx, t, hx);
  const float du = fmaf(fmaf(x2, 5.0f * GT_C2, 3.0f * GT_C1), x2, GT_C0);
  const float w = x * fmaf(-t, t, 1.0f);
  dg = fmaf(0.5f, fmaf(w, du, t), 0.5f);
}
template <bool FAST> __device__ __forceinline__ float gelu_t(float x) {
  if (FAST && MMA_GELU_TANH) return gelu_tanh_scaled(x, 0.5f);
  float cdf, e;
  normal_cdf_exp<FAST>(x, cdf, e);
  return x * cdf;
}
template <bool FAST> __device__ __forceinline__ float dgelu_t(float x) {
  if (FAST && MMA_GELU_TANH) return dgelu_tanh_scaled(x, 0.5f);
  float cdf, e;
  normal_cdf_exp<FAST>(x, cdf, e);
  return fmaf(x * 0.39894228040143267794f, e, cdf);
}
template <bool FAST> __device__ __forceinline__ void gelu_both(float x, float& g, float& dg) {
  if (FAST && MMA_GELU_TANH) {
    gelu_both_tanh(x, g, dg);
    return;
  }
  float cdf, e;
  normal_cdf_exp<FAST>(x, cdf, e);
  g = x * cdf;
  dg = fmaf(x * 0.39894228040143267794f, e, cdf);
}

// ---- vector row-segment loads / stores used by the epilogue ------------------------------------------
template <int NV>
__device__ __forceinline__ void ld_row_f32(float (&x)[NV], const float* p, int nvalid) {
  if (nvalid == NV && (NV % 4) == 0 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
#pragma unroll
    for (int j = 0; j < NV; j += 4) {
      const float4 v = *reinterpret_cast<const float4*>(p + j);
      x[j] = v.x; x[j + 1] = v.y; x[j + 2] = v.z; x[j + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j) x[j] = j < nvalid ? p[j] : 0.f;
  }
}
template <int NV>
__device__ __forceinline__ void ld_row_bf16(float (&x)[NV], const bf16* p, int nvalid) {
  if (nvalid == NV && (NV % 8) == 0 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
#pragma unroll
    for (int j = 0; j < NV; j += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(p + j);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float2 f = __bfloat1622float2(h[w]);
        x[j + 2 * w] = f.x;
        x[j + 2 * w + 1] = f.y;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j) x[j] = j < nvalid ? __bfloat162float(p[j]) : 0.f;
  }
}
template <int NV>
__device__ __forceinline__ void ld_row_any(float (&x)[NV], const void* base, long long idx, int is_f32, int nvalid) {
  if (is_f32) ld_row_f32<NV>(x, reinterpret_cast<const float*>(base) + idx, nvalid);
  else ld_row_bf16<NV>(x, reinterpret_cast<const bf16*>(base) + idx, nvalid);
}
template <int NV>
__device__ __forceinline__ void st_row_any(void* base, long long idx, int is_f32, const float (&x)[NV], int nvalid) {
  if (is_f32) {
    float* o = reinterpret_cast<float*>(base) + idx;
    if (nvalid == NV && (NV % 4) == 0 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < NV; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (j < nvalid) o[j] = x[j];
    }
  } else {
    bf16* o = reinterpret_cast<bf16*>(base) + idx;
    if (nvalid == NV && (NV % 8) == 0 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < NV; j += 8) {
        __nv_bfloat162 a = __floats2bfloat162_rn(x[j], x[j + 1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(x[j + 2], x[j + 3]);
        __nv_bfloat162 c = __floats2bfloat162_rn(x[j + 4], x[j + 5]);
        __nv_bfloat162 d = __floats2bfloat162_rn(x[j + 6], x[j + 7]);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&a);
        u.y = *reinterpret_cast<uint32_t*>(&b);
        u.z = *reinterpret_cast<uint32_t*>(&c);
        u.w = *reinterpret_cast<uint32_t*>(&d);
        *reinterpret_cast<uint4*>(o + j) = u;
      }
    } else if (nvalid == NV && (NV % 4) == 0 && ((reinterpret_cast<uintptr_t>(o) & 7) == 0)) {
#pragma unroll
      for (int j = 0; j < NV; j += 4) {
        __nv_bfloat162 a = __floats2bfloat162_rn(x[j], x[j + 1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(x[j + 2], x[j + 3]);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&a);
        u.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(o + j) = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (j < nvalid) o[j] = __float2bfloat16_rn(x[j]);
    }
  }
}

// ---- the GEMM epilogue: NV consecutive columns of one output row -------------------------------------
// v[] holds the fp32 accumulators for columns col .. col+NV-1 of row `row`; ncols is the logical N.
// KIND >= 0 fixes the epilogue at compile time (tcgen05 kernels); KIND < 0 dispatches on ep.kind.
// FAST selects the bf16-grade erf.
template <int NV, int KIND, bool FAST>
__device__ __forceinline__ void epilogue_store(const Epi& ep, long long row, int col, int ncols, float (&v)[NV]) {
  const int nvalid = min(NV, ncols - col);
  if (nvalid <= 0) return;
  const int kind = KIND >= 0 ? KIND : ep.kind;
  if (kind == EPI_ACCUM) {
    float* o = reinterpret_cast<float*>(ep.out) + row * ep.ldo + col;
    if (ep.accumulate == 2) {
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (j < nvalid) atomicAdd(o + j, v[j] * ep.alpha);
    } else if (ep.accumulate == 1) {
      float cur[NV];
      ld_row_f32<NV>(cur, o, nvalid);
#pragma unroll
      for (int j = 0; j < NV; ++j) cur[j] += v[j] * ep.alpha;
      st_row_any<NV>(ep.out, row * ep.ldo + col, 1, cur, nvalid);
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] *= ep.alpha;
      st_row_any<NV>(ep.out, row * ep.ldo + col, 1, v, nvalid);
    }
    return;
  }
  // bias
  if (ep.bias && (kind == EPI_STORE || kind == EPI_GELU || kind == EPI_RESID || kind == EPI_GLU_MUL || kind == EPI_RELU)) {
    float bv[NV];
    ld_row_f32<NV>(bv, ep.bias + col, nvalid);
    if (kind == EPI_STORE) {
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] = fmaf(v[j], ep.alpha, bv[j]);
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] += bv[j];
    }
  } else if (kind == EPI_STORE) {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] *= ep.alpha;
  }
  // dropout keep-scales, applied in place where the kind says so
  const bool drop = ep.p_drop > 0.0f &&
                    (kind == EPI_GELU || kind == EPI_RESID || kind == EPI_DGELU || kind == EPI_GLU_MUL || kind == EPI_DGLU);
  const uint32_t thr = drop ? drop_threshold(ep.p_drop) : 0u;
  const float inv_keep = drop ? 1.0f / (1.0f - ep.p_drop) : 1.0f;
  const uint32_t dkey = drop ? drop_key(ep.seed, ep.site) : 0u;
  const uint32_t e0 = (uint32_t)((unsigned long long)row * (unsigned long long)ep.drop_ld + (unsigned long long)col);
  float ds[NV];
  if (drop) {
    if ((e0 & 1u) == 0 && (NV % 2) == 0) {
#pragma unroll
      for (int j = 0; j < NV; j += 2) {
        const uint32_t r = drop_pair(dkey, (e0 + j) >> 1);
        ds[j] = (r & 0xFFFFu) >= thr ? inv_keep : 0.f;
        ds[j + 1] = (r >> 16) >= thr ? inv_keep : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j) ds[j] = drop_scale1(dkey, e0 + j, thr, inv_keep);
    }
  }

  switch (kind) {
    case EPI_STORE:
      st_row_any<NV>(ep.out, row * ep.ldo + col, ep.out_f32, v, nvalid);
      break;
    case EPI_RELU: {
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] = fmaxf(v[j], 0.f);
      st_row_any<NV>(ep.out, row * ep.ldo + col, ep.out_f32, v, nvalid);
    } break;
    case EPI_GELU: {
      if (ep.out2) st_row_any<NV>(ep.out2, row * ep.ldo2 + col, ep.out_f32, v, nvalid);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        float y = gelu_t<FAST>(v[j]);
        if (drop) y *= ds[j];
        v[j] = y;
      }
      st_row_any<NV>(ep.out, row * ep.ldo + col, ep.out_f32, v, nvalid);
    } break;
    case EPI_RESID: {
      float r[NV];
      ld_row_any<NV>(r, ep.resid, row * ep.ldr + col, ep.resid_f32, nvalid);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        float y = v[j];
        if (drop) y *= ds[j];
        v[j] = r[j] + y;
      }
      st_row_any<NV>(ep.out, row * ep.ldo + col, ep.out_f32, v, nvalid);
    } break;
    case EPI_DGELU: {
      float z[NV];
      ld_row_any<NV>(z, ep.aux, row * ep.lda + col, ep.aux_f32, nvalid);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        float y = v[j] * dgelu_t<FAST>(z[j]);
        if (drop) y *= ds[j];
        v[j] = y;
      }
      st_row_any<NV>(ep.out, row * ep.ldo + col, ep.out_f32, v, nvalid);
    } break;
    case EPI_DRELU: {
      float z[NV];
      ld_row_any<NV>(z, ep.aux, row * ep.lda + col, ep.aux_f32, nvalid);
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] = z[j] > 0.f ? v[j] : 0.f;
      st_row_any<NV>(ep.out, row * ep.ldo + col, ep.out_f32, v, nvalid);
    } break;
    case EPI_GLU_MUL: {
      float z1[NV];
      ld_row_any<NV>(z1, ep.aux, row * ep.lda + col, ep.aux_f32, nvalid);
      if (ep.out2) st_row_any<NV>(ep.out2, row * ep.ldo2 + col, ep.out_f32, v, nvalid);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        float y = gelu_t<FAST>(z1[j]) * v[j];
        if (drop) y *= ds[j];
        v[j] = y;
      }
      st_row_any<NV>(ep.out, row * ep.ldo + col, ep.out_f32, v, nvalid);
    } break;
    case EPI_DGLU: {
      float z1[NV], z2[NV];
      ld_row_any<NV>(z1, ep.aux, row * ep.lda + col, ep.aux_f32, nvalid);
      ld_row_any<NV>(z2, ep.aux2, row * ep.lda2 + col, ep.aux_f32, nvalid);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        float da = v[j];
        if (drop) da *= ds[j];
        float g, dg;
        gelu_both<FAST>(z1[j], g, dg);
        v[j] = da * z2[j] * dg;  // d z1
        z2[j] = da * g;          // d z2
      }
      st_row_any<NV>(ep.out, row * ep.ldo + col, ep.out_f32, v, nvalid);
      st_row_any<NV>(ep.out2, row * ep.ldo2 + col, ep.out_f32, z2, nvalid);
    } break;
    default:
      break;
  }
}

// ---- split epilogue for the software-pipelined tcgen05 kernels ---------------------------------------------
// epi_prefetch issues the global loads an epilogue needs besides the accumulator (residual / saved activations /
// the running gradient) so they overlap the TMEM read of the same chunk; epi_math consumes them and leaves the
// primary output in v[] and the secondary one (pre-activation copy, GLU gate, d z2) in o2[].  Storing is the
// caller's business (shared-memory staged, coalesced).  Semantics identical to epilogue_store<NV, KIND, FAST>.
template <int NV, int KIND>
struct EpiIn {
  float a[NV];
  float b[KIND == EPI_DGLU ? NV : 1];
};

template <int NV, int KIND>
__device__ __forceinline__ void epi_prefetch(const Epi& ep, long long row, int col, int ncols, bool row_ok,
                                             EpiIn<NV, KIND>& in) {
  const int nvalid = row_ok ? min(NV, ncols - col) : 0;
  if (nvalid <= 0) return;
  if (KIND == EPI_RESID) ld_row_any<NV>(in.a, ep.resid, row * ep.ldr + col, ep.resid_f32, nvalid);
  if (KIND == EPI_DGELU || KIND == EPI_GLU_MUL || KIND == EPI_DGLU || KIND == EPI_DRELU)
    ld_row_any<NV>(in.a, ep.aux, row * ep.lda + col, ep.aux_f32, nvalid);
  if constexpr (KIND == EPI_DGLU) ld_row_any<NV>(in.b, ep.aux2, row * ep.lda2 + col, ep.aux_f32, nvalid);
  if (KIND == EPI_ACCUM) {
    if (ep.accumulate == 1) ld_row_f32<NV>(in.a, reinterpret_cast<const float*>(ep.out) + row * ep.ldo + col, nvalid);
  }
}

// returns true when o2[] holds a second output
template <int NV, int KIND, bool FAST>
__device__ __forceinline__ bool epi_math(const Epi& ep, long long row, int col, float (&v)[NV], float (&o2)[NV],
                                         EpiIn<NV, KIND>& in, const float* bias_s) {
  if (KIND == EPI_ACCUM) {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = ep.accumulate == 1 ? fmaf(v[j], ep.alpha, in.a[j]) : v[j] * ep.alpha;
    return false;
  }
  if (KIND == EPI_STORE) {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = bias_s ? fmaf(v[j], ep.alpha, bias_s[j]) : v[j] * ep.alpha;
    return false;
  }
  if (bias_s && (KIND == EPI_GELU || KIND == EPI_RESID || KIND == EPI_GLU_MUL || KIND == EPI_RELU)) {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] += bias_s[j];
  }
  const bool drop = ep.p_drop > 0.0f &&
                    (KIND == EPI_GELU || KIND == EPI_RESID || KIND == EPI_DGELU || KIND == EPI_GLU_MUL || KIND == EPI_DGLU);
  float ds[NV];
  if (drop) {
    const uint32_t thr = drop_threshold(ep.p_drop);
    const float inv_keep = 1.0f / (1.0f - ep.p_drop);
    const uint32_t dkey = drop_key(ep.seed, ep.site);
    const uint32_t e0 = (uint32_t)((unsigned long long)row * (unsigned long long)ep.drop_ld + (unsigned long long)col);
    if ((e0 & 1u) == 0 && (NV % 2) == 0) {
#pragma unroll
      for (int j = 0; j < NV; j += 2) {
        const uint32_t r = drop_pair(dkey, (e0 + j) >> 1);
        ds[j] = (r & 0xFFFFu) >= thr ? inv_keep : 0.f;
        ds[j + 1] = (r >> 16) >= thr ? inv_keep : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j) ds[j] = drop_scale1(dkey, e0 + j, thr, inv_keep);
    }
  }
  bool has2 = false;
  if (KIND == EPI_RELU) {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (KIND == EPI_GELU) {
    has2 = ep.out2 != nullptr;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      o2[j] = v[j];
      float y = gelu_t<FAST>(v[j]);
      if (drop) y *= ds[j];
      v[j] = y;
    }
  } else if (KIND == EPI_RESID) {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = in.a[j] + (drop ? v[j] * ds[j] : v[j]);
  } else if (KIND == EPI_DGELU) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float y = v[j] * dgelu_t<FAST>(in.a[j]);
      if (drop) y *= ds[j];
      v[j] = y;
    }
  } else if (KIND == EPI_DRELU) {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = in.a[j] > 0.f ? v[j] : 0.f;
  } else if (KIND == EPI_GLU_MUL) {
    has2 = ep.out2 != nullptr;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      o2[j] = v[j];
      float y = gelu_t<FAST>(in.a[j]) * v[j];
      if (drop) y *= ds[j];
      v[j] = y;
    }
  } else if constexpr (KIND == EPI_DGLU) {
    has2 = true;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float da = drop ? v[j] * ds[j] : v[j];
      float g, dg;
      gelu_both<FAST>(in.a[j], g, dg);
      v[j] = da * in.b[j] * dg;
      o2[j] = da * g;
    }
  }
  return has2;
}

template <int NV, int KIND, bool FAST>
__device__ __forceinline__ void epi_finish(const Epi& ep, long long row, int col, int ncols, bool row_ok, float (&v)[NV],
                                           EpiIn<NV, KIND>& in, const float* bias_s) {
  const int nvalid = row_ok ? min(NV, ncols - col) : 0;
  if (nvalid <= 0) return;
  if (KIND == EPI_ACCUM && ep.accumulate == 2) {
    float* o = reinterpret_cast<float*>(ep.out) + row * ep.ldo + col;
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (j < nvalid) atomicAdd(o + j, v[j] * ep.alpha);
    return;
  }
  float o2[NV];
  const bool has2 = epi_math<NV, KIND, FAST>(ep, row, col, v, o2, in, bias_s);
  if (has2) st_row_any<NV>(ep.out2, row * ep.ldo2 + col, ep.out_f32, o2, nvalid);
  st_row_any<NV>(ep.out, row * ep.ldo + col, KIND == EPI_ACCUM ? 1 : ep.out_f32, v, nvalid);
}
