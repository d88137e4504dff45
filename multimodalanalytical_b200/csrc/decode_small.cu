// Small-batch decode products: Y[R, N] = epilogue( LN?(X)[R, K] W[N, K]^T + b ) for a handful of spectra x beams: blocks of
// <= 64 rows (grid.y walks up to 8 such blocks; the weights of the second and later blocks come out of L2).  At that size the step is ~70 strictly dependent launches whose cost is latency, not math: the tcgen05 kernels pay
// TMEM allocation, tensor-map fetch, a TMA pipeline fill and a row-per-thread epilogue for a 128-row tile that is 92 %
// padding, and every LayerNorm is a launch of its own.  Here one CTA owns 16 output features, its 8 warps split the
// reduction, and the product runs on mma.sync.m16n8k16 with the WEIGHT tile as the 16-row A operand and the (few) rows
// of X as the 8-column B operand - so every weight element is read exactly once, by exactly one warp, straight from
// L2 / HBM with 16-byte loads (the k order inside a 32-element block is permuted identically for A and B, which a dot
// product does not see) - and the LayerNorm of the fp32 residual stream is the kernel's prologue (each CTA normalises the
// R rows into shared memory), the bias / GELU / gate / residual its epilogue.
#include "common.cuh"

namespace dsm {

constexpr int MAXR = 64;
constexpr int NT_MAX = MAXR / 8;
constexpr int MAXBLK = 8;  // row blocks per launch
enum { K_STORE = 0, K_GELU = 1, K_RESID = 2, K_GLU = 3 };

struct LinArgs {
  const void* x; int x_f32; long long ldx;  // [R, K]: fp32 (LayerNorm applied when gamma != null) or bf16
  const float* gamma; const float* beta; float eps;
  const bf16* w; const bf16* w2; long long ldw;  // [N, K] (w2: gate weights, K_GLU)
  const float* bias; const float* bias2;
  const float* resid; long long ldr;  // K_RESID: fp32 [R, N]
  void* out; int out_f32; long long ldo;
  int R, N, K, kind;
  int ftiles;  // 16-feature tiles per CTA: more rows -> more tiles per CTA, so the per-CTA LayerNorm prologue is amortised
  int rblk;    // rows per row block (blockIdx.y), a multiple of 8, <= MAXR
};

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <bool GLU>
__global__ void __launch_bounds__(256) small_linear_kernel(LinArgs a) {
  pdl_trigger();
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int r0 = blockIdx.y * a.rblk;          // this CTA's row block
  const int R = min(a.rblk, a.R - r0), K = a.K;
  if (R <= 0) return;  // uniform over the CTA (cannot happen with the host's even split)
  const int NT = (R + 7) >> 3;
  const int pitch = K + 32;  // bf16 elements per staged row: rows 64 bytes apart in bank space -> conflict-free 16-byte reads
  bf16* xs = reinterpret_cast<bf16*>(smem);                                   // [NT * 8][pitch] when staged
  const bool staged = a.x_f32 != 0;
  float* red = reinterpret_cast<float*>(smem + (staged ? (size_t)NT * 8 * pitch * 2 : 0));  // [8 warps][(GLU ? 2 : 1)][NT*8][16]

  if (staged) {
    // prologue: fp32 rows -> (LayerNorm) -> bf16 rows in shared memory; warp w takes rows w, w + 8, ...
    const float* X = reinterpret_cast<const float*>(a.x);
    for (int r = warp; r < NT * 8; r += 8) {
      bf16* dst = xs + (size_t)r * pitch;
      if (r >= R) {
        for (int c = lane * 8; c < K; c += 256) *reinterpret_cast<uint4*>(dst + c) = make_uint4(0u, 0u, 0u, 0u);
        continue;
      }
      const float* src = X + (long long)(r0 + r) * a.ldx;
      // one pass over global memory: the row lives in registers (K <= 1024: 8 float4 per lane)
      float4 v[8];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = lane * 4 + i * 128;
        v[i] = c < K ? *reinterpret_cast<const float4*>(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        s += v[i].x + v[i].y + v[i].z + v[i].w;
      }
      float mean = 0.f, rstd = 1.f;
      if (a.gamma) {
        mean = warp_sum(s) / K;
        float qd = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (lane * 4 + i * 128 < K) {
            const float e0 = v[i].x - mean, e1 = v[i].y - mean, e2 = v[i].z - mean, e3 = v[i].w - mean;
            qd += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
          }
        }
        rstd = rsqrtf(warp_sum(qd) / K + a.eps);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = lane * 4 + i * 128;
        if (c < K) {
          float4 t = v[i];
          if (a.gamma) {
            const float4 gm = *reinterpret_cast<const float4*>(a.gamma + c), bt = *reinterpret_cast<const float4*>(a.beta + c);
            t.x = (t.x - mean) * rstd * gm.x + bt.x; t.y = (t.y - mean) * rstd * gm.y + bt.y;
            t.z = (t.z - mean) * rstd * gm.z + bt.z; t.w = (t.w - mean) * rstd * gm.w + bt.w;
          }
          __nv_bfloat162 p0 = __floats2bfloat162_rn(t.x, t.y), p1 = __floats2bfloat162_rn(t.z, t.w);
          uint2 u;
          u.x = *reinterpret_cast<uint32_t*>(&p0);
          u.y = *reinterpret_cast<uint32_t*>(&p1);
          *reinterpret_cast<uint2*>(dst + c) = u;
        }
      }
    }
    __syncthreads();
  }

  const int ks = K >> 3;  // reduction slice of this warp (a multiple of 32: checked on the host)
  const int kbeg = warp * ks;
  const int rows8 = NT * 8;
  const int set_stride = rows8 * 16;
  const bf16* xg = staged ? nullptr : reinterpret_cast<const bf16*>(a.x) + (long long)r0 * a.ldx;
  for (int ft = 0; ft < a.ftiles; ++ft) {
  const int f0 = (blockIdx.x * a.ftiles + ft) * 16;
  if (f0 >= a.N) break;  // uniform over the CTA
  const int fa = f0 + g, fb = f0 + g + 8;
  const bf16* wa = a.w + (long long)min(fa, a.N - 1) * a.ldw;
  const bf16* wb = a.w + (long long)min(fb, a.N - 1) * a.ldw;
  const bf16* ga = GLU ? a.w2 + (long long)min(fa, a.N - 1) * a.ldw : nullptr;
  const bf16* gb = GLU ? a.w2 + (long long)min(fb, a.N - 1) * a.ldw : nullptr;
  float acc[NT_MAX][4], acc2[GLU ? NT_MAX : 1][4];
#pragma unroll
  for (int t = 0; t < NT_MAX; ++t)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      acc[t][e] = 0.f;
      if (GLU) acc2[t][e] = 0.f;
    }
  for (int k0 = kbeg; k0 < kbeg + ks; k0 += 32) {
    const int kk = k0 + 8 * q;  // this lane's 8 consecutive k of the 32-element block
    const uint4 A0 = __ldg(reinterpret_cast<const uint4*>(wa + kk)), A1 = __ldg(reinterpret_cast<const uint4*>(wb + kk));
    uint4 G0 = make_uint4(0u, 0u, 0u, 0u), G1 = G0;
    if (GLU) {
      G0 = __ldg(reinterpret_cast<const uint4*>(ga + kk));
      G1 = __ldg(reinterpret_cast<const uint4*>(gb + kk));
    }
#pragma unroll
    for (int t = 0; t < NT_MAX; ++t) {
      if (t < NT) {
        const int n = t * 8 + g;
        uint4 Bv;
        if (staged) Bv = *reinterpret_cast<const uint4*>(xs + (size_t)n * pitch + kk);
        else Bv = n < R ? *reinterpret_cast<const uint4*>(xg + (long long)n * a.ldx + kk) : make_uint4(0u, 0u, 0u, 0u);
        mma16816(acc[t], A0.x, A1.x, A0.y, A1.y, Bv.x, Bv.y);
        mma16816(acc[t], A0.z, A1.z, A0.w, A1.w, Bv.z, Bv.w);
        if (GLU) {
          mma16816(acc2[t], G0.x, G1.x, G0.y, G1.y, Bv.x, Bv.y);
          mma16816(acc2[t], G0.z, G1.z, G0.w, G1.w, Bv.z, Bv.w);
        }
      }
    }
  }
  // cross-warp reduction of the 8 reduction slices: red[warp][set][n][f]
  float* mine = red + (size_t)warp * (GLU ? 2 : 1) * set_stride;
#pragma unroll
  for (int t = 0; t < NT_MAX; ++t) {
    if (t < NT) {
      const int n = t * 8 + 2 * q;
      mine[n * 16 + g] = acc[t][0];
      mine[(n + 1) * 16 + g] = acc[t][1];
      mine[n * 16 + g + 8] = acc[t][2];
      mine[(n + 1) * 16 + g + 8] = acc[t][3];
      if (GLU) {
        float* m2 = mine + set_stride;
        m2[n * 16 + g] = acc2[t][0];
        m2[(n + 1) * 16 + g] = acc2[t][1];
        m2[n * 16 + g + 8] = acc2[t][2];
        m2[(n + 1) * 16 + g + 8] = acc2[t][3];
      }
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < R * 16; idx += 256) {
    const int n = idx >> 4, f = idx & 15;
    const int col = f0 + f;
    if (col >= a.N) continue;
    float v = 0.f, v2 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float* src = red + (size_t)w * (GLU ? 2 : 1) * set_stride;
      v += src[n * 16 + f];
      if (GLU) v2 += src[set_stride + n * 16 + f];
    }
    if (a.bias) v += a.bias[col];
    if (GLU) {
      if (a.bias2) v2 += a.bias2[col];
      v = gelu_t<true>(v) * v2;
    } else if (a.kind == K_GELU) {
      v = gelu_t<true>(v);
    } else if (a.kind == K_RESID) {
      v += a.resid[(long long)(r0 + n) * a.ldr + col];
    }
    if (a.out_f32) reinterpret_cast<float*>(a.out)[(long long)(r0 + n) * a.ldo + col] = v;
    else reinterpret_cast<bf16*>(a.out)[(long long)(r0 + n) * a.ldo + col] = __float2bfloat16_rn(v);
  }
  __syncthreads();  // `red` is rewritten by the next feature tile
  }
}

}  // namespace dsm

// out[R, N] = epilogue( LN?(x)[R, K] w[N, K]^T + bias ) for R <= 512 (the decode step of a few spectra), in row blocks of <= 64.
//   x: fp32 [R, K] (x_f32 = 1; LayerNorm(gamma, beta, eps) applied when gamma != NULL) or bf16 [R, K] (x_f32 = 0);
//   w (and, kind 3, the gate weights w2 with bias2): bf16 [N, K], pitch ldw;  kind: 0 store, 1 exact-erf GELU,
//   2 out = resid (fp32 [R, N]) + result, 3 gated: gelu(x w^T + bias) * (x w2^T + bias2);  out: fp32 or bf16 [R, N].
// K must be a multiple of 256 (8 warps x 32-element blocks); returns MMA_ERR_UNSUPPORTED otherwise or when R > 512.
extern "C" int mma_small_linear(const void* x, int x_f32, long long ldx, const float* gamma, const float* beta, float eps,
                                const void* w, const void* w2, long long ldw, const float* bias, const float* bias2,
                                const float* resid, long long ldr, void* out, int out_f32, long long ldo, int R, int N,
                                int K, int kind, cudaStream_t stream) {
  using namespace dsm;
  if (R <= 0 || N <= 0 || K <= 0 || !x || !w || !out || kind < 0 || kind > 3) return MMA_ERR_ARG;
  if (R > MAXR * MAXBLK || (K & 255) || (x_f32 && K > 1024) || (ldw & 7) || (reinterpret_cast<uintptr_t>(w) & 15) || (kind == K_GLU && !w2) ||
      (kind == K_RESID && !resid))
    return MMA_ERR_UNSUPPORTED;
  if (x_f32 ? ((ldx & 3) || (reinterpret_cast<uintptr_t>(x) & 15)) : ((ldx & 7) || (reinterpret_cast<uintptr_t>(x) & 15)))
    return MMA_ERR_UNSUPPORTED;
  if (gamma && !x_f32) return MMA_ERR_ARG;
  LinArgs a{x, x_f32, ldx, gamma, beta, eps, reinterpret_cast<const bf16*>(w), reinterpret_cast<const bf16*>(w2), ldw, bias,
            bias2, resid, ldr, out, out_f32, ldo, R, N, K, kind, 1, 0};
  const int nblk = (R + MAXR - 1) / MAXR;
  a.rblk = (((R + nblk - 1) / nblk) + 7) & ~7;  // even split, whole 8-row MMA tiles
  const int NT = a.rblk / 8;
  const size_t smem = (x_f32 ? (size_t)NT * 8 * (K + 32) * 2 : 0) + (size_t)8 * (kind == K_GLU ? 2 : 1) * NT * 8 * 16 * 4;
  const dim3 grid((N + 16 * a.ftiles - 1) / (16 * a.ftiles), nblk);
  if (kind == K_GLU) {
    static size_t set = 0;
    if (smem > set) {
      if (cudaFuncSetAttribute(small_linear_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return MMA_ERR_LAUNCH;
      set = smem;
    }
    small_linear_kernel<true><<<grid, 256, smem, stream>>>(a);
  } else {
    static size_t set = 0;
    if (smem > set) {
      if (cudaFuncSetAttribute(small_linear_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return MMA_ERR_LAUNCH;
      set = smem;
    }
    small_linear_kernel<false><<<grid, 256, smem, stream>>>(a);
  }
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}
