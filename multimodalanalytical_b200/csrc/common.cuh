// Shared device helpers for the sm_100a spectra->SMILES kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

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

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float dgelu_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
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

// ---- Philox4x32-10 counter RNG: dropout masks are a pure function of (seed, site, element) so the
// backward pass regenerates them instead of storing them -------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}
// random words for elements 4*g .. 4*g+3 of dropout site `site`
__device__ __forceinline__ uint4 drop_words(unsigned long long seed, unsigned int site, unsigned long long g) {
  return philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), site, 0x5eedu),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
__device__ __forceinline__ uint32_t drop_threshold(float p) {
  double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
}
// keep-scale (0 or 1/(1-p)) of a single element e
__device__ __forceinline__ float drop_scale1(unsigned long long seed, unsigned int site, unsigned long long e,
                                             uint32_t thr, float inv_keep) {
  const uint4 w = drop_words(seed, site, e >> 2);
  const uint32_t r = (e & 3) == 0 ? w.x : (e & 3) == 1 ? w.y : (e & 3) == 2 ? w.z : w.w;
  return r >= thr ? inv_keep : 0.0f;
}

// ---- the GEMM epilogue: `n` consecutive columns of one output row --------------------------------
// v[] holds the fp32 accumulators for columns col .. col+n-1 of row `row`; ncols is the logical N.
template <int NV>
__device__ __forceinline__ void epilogue_store(const Epi& ep, long long row, int col, int ncols, float (&v)[NV]) {
  const int nvalid = min(NV, ncols - col);
  if (nvalid <= 0) return;
  const bool drop = ep.p_drop > 0.0f;
  const uint32_t thr = drop ? drop_threshold(ep.p_drop) : 0u;
  const float inv_keep = drop ? 1.0f / (1.0f - ep.p_drop) : 1.0f;
  float ds[NV];
  if (drop) {
    const unsigned long long e0 = (unsigned long long)row * (unsigned long long)ep.drop_ld + (unsigned long long)col;
    if ((e0 & 3) == 0 && (NV % 4) == 0) {
#pragma unroll
      for (int j = 0; j < NV; j += 4) {
        const uint4 w = drop_words(ep.seed, ep.site, (e0 + j) >> 2);
        ds[j] = w.x >= thr ? inv_keep : 0.f;
        ds[j + 1] = w.y >= thr ? inv_keep : 0.f;
        ds[j + 2] = w.z >= thr ? inv_keep : 0.f;
        ds[j + 3] = w.w >= thr ? inv_keep : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j) ds[j] = drop_scale1(ep.seed, ep.site, e0 + j, thr, inv_keep);
    }
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j) ds[j] = 1.0f;
  }
  float o1[NV], o2[NV];
  bool has2 = false;
  switch (ep.kind) {
    case EPI_STORE: {
#pragma unroll
      for (int j = 0; j < NV; ++j) o1[j] = v[j] * ep.alpha + ((ep.bias && j < nvalid) ? ep.bias[col + j] : 0.f);
    } break;
    case EPI_GELU: {
      has2 = ep.out2 != nullptr;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float z = v[j] + ((ep.bias && j < nvalid) ? ep.bias[col + j] : 0.f);
        o2[j] = z;
        o1[j] = gelu_f(z) * ds[j];
      }
    } break;
    case EPI_RESID: {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float y = v[j] + ((ep.bias && j < nvalid) ? ep.bias[col + j] : 0.f);
        const float r = j < nvalid ? ld_any(ep.resid, row * ep.ldr + col + j, ep.resid_f32) : 0.f;
        o1[j] = r + y * ds[j];
      }
    } break;
    case EPI_DGELU: {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float z = j < nvalid ? ld_any(ep.aux, row * ep.lda + col + j, ep.aux_f32) : 0.f;
        o1[j] = v[j] * ds[j] * dgelu_f(z);
      }
    } break;
    case EPI_GLU_MUL: {
      has2 = ep.out2 != nullptr;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float z2 = v[j] + ((ep.bias && j < nvalid) ? ep.bias[col + j] : 0.f);
        const float z1 = j < nvalid ? ld_any(ep.aux, row * ep.lda + col + j, ep.aux_f32) : 0.f;
        o2[j] = z2;
        o1[j] = gelu_f(z1) * z2 * ds[j];
      }
    } break;
    case EPI_DGLU: {
      has2 = true;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float z1 = j < nvalid ? ld_any(ep.aux, row * ep.lda + col + j, ep.aux_f32) : 0.f;
        const float z2 = j < nvalid ? ld_any(ep.aux2, row * ep.lda2 + col + j, ep.aux_f32) : 0.f;
        const float da = v[j] * ds[j];
        o1[j] = da * z2 * dgelu_f(z1);
        o2[j] = da * gelu_f(z1);
      }
    } break;
    case EPI_RELU: {
#pragma unroll
      for (int j = 0; j < NV; ++j) o1[j] = fmaxf(v[j] + ((ep.bias && j < nvalid) ? ep.bias[col + j] : 0.f), 0.f);
    } break;
    case EPI_DRELU: {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float h = j < nvalid ? ld_any(ep.aux, row * ep.lda + col + j, ep.aux_f32) : 0.f;
        o1[j] = h > 0.f ? v[j] : 0.f;
      }
    } break;
    case EPI_ACCUM: {
      float* o = reinterpret_cast<float*>(ep.out) + row * ep.ldo + col;
      if (ep.accumulate == 2) {
#pragma unroll
        for (int j = 0; j < NV; ++j)
          if (j < nvalid) atomicAdd(o + j, v[j] * ep.alpha);
      } else if (ep.accumulate == 1) {
#pragma unroll
        for (int j = 0; j < NV; ++j)
          if (j < nvalid) o[j] += v[j] * ep.alpha;
      } else {
#pragma unroll
        for (int j = 0; j < NV; ++j)
          if (j < nvalid) o[j] = v[j] * ep.alpha;
      }
      return;
    }
    default:
      return;
  }
  // ---- stores (vectorised when the row segment is 16-byte aligned and full) ----
  auto store_vec = [&](void* base, long long ld, float (&x)[NV]) {
    if (ep.out_f32) {
      float* o = reinterpret_cast<float*>(base) + row * ld + col;
      if (nvalid == NV && (NV % 4) == 0 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < NV; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < NV; ++j)
          if (j < nvalid) o[j] = x[j];
      }
    } else {
      bf16* o = reinterpret_cast<bf16*>(base) + row * ld + col;
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
  };
  store_vec(ep.out, ep.ldo, o1);
  if (has2) store_vec(ep.out2, ep.ldo2, o2);
}
