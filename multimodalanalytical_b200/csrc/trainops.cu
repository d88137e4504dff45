// Loss and optimiser kernels (HBM-bound, fp32 statistics):
//   * cross-entropy over the target vocabulary (ignore_index, mean over non-ignored tokens, optional label
//     smoothing; reference: nn.CrossEntropyLoss() at custom_modeling.py:490-491, labels pad -> -100 at
//     wrapper.py:389) with a deterministic two-stage reduction, and its gradient written straight into the
//     (padded, low-precision) dlogits operand of the LM-head dgrad / wgrad GEMMs;
//   * fused global-norm clip + Adam / AdamW step on flat fp32 master buffers that also refreshes the bf16
//     weight mirror and zeroes the gradient (reference: torch.optim.Adam/AdamW, wrapper.py:329-344;
//     clip_grad 1.0, trainer/trainer.py:65).
#include "common.cuh"

namespace trainops {

// one warp per row
__global__ void __launch_bounds__(256) ce_rows_kernel(const float* __restrict__ logits, long long ld,
                                                      const long long* __restrict__ labels, int rows, int V,
                                                      float smoothing, long long ignore_index,
                                                      float* __restrict__ row_loss, float* __restrict__ row_lse) {
  pdl_trigger();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* x = logits + (long long)r * ld;
  float mx = -INFINITY;
  for (int c = lane; c < V; c += 32) mx = fmaxf(mx, x[c]);
  mx = warp_max(mx);
  float s = 0.f, tot = 0.f;
  for (int c = lane; c < V; c += 32) {
    s += expf(x[c] - mx);
    tot += x[c];
  }
  s = warp_sum(s);
  tot = warp_sum(tot);
  const float lse = mx + logf(s);
  if (lane == 0) {
    const long long y = labels[r];
    float loss = 0.f;
    if (y != ignore_index) {
      loss = (1.f - smoothing) * (lse - x[y]);
      if (smoothing > 0.f) loss += smoothing * (lse - tot / V);
    }
    row_loss[r] = loss;
    row_lse[r] = lse;
  }
}

// single block, fixed summation order: out[0] = mean loss over valid rows, out[1] = number of valid rows
__global__ void __launch_bounds__(1024) ce_reduce_kernel(const float* __restrict__ row_loss,
                                                         const long long* __restrict__ labels, int rows,
                                                         long long ignore_index, float* __restrict__ out) {
  pdl_trigger();
  __shared__ float ssum[1024];
  __shared__ float scnt[1024];
  float s = 0.f, n = 0.f;
  for (int r = threadIdx.x; r < rows; r += 1024) {
    if (labels[r] != ignore_index) {
      s += row_loss[r];
      n += 1.f;
    }
  }
  ssum[threadIdx.x] = s;
  scnt[threadIdx.x] = n;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      ssum[threadIdx.x] += ssum[threadIdx.x + o];
      scnt[threadIdx.x] += scnt[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = ssum[0] / scnt[0];
    out[1] = scnt[0];
  }
}

// dlogits[r, c] = gscale / n_valid * (softmax(x)_c - (1-eps) [c == y] - eps / V);  zero for ignored rows and for
// the padding columns [V, ldd)
template <typename T>
__global__ void __launch_bounds__(256) ce_bwd_kernel(const float* __restrict__ logits, long long ld,
                                                     const long long* __restrict__ labels,
                                                     const float* __restrict__ row_lse,
                                                     const float* __restrict__ stats, float gscale, int rows, int V,
                                                     float smoothing, long long ignore_index, T* __restrict__ dlogits,
                                                     long long ldd) {
  pdl_trigger();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= rows) return;
  const long long y = labels[r];
  T* o = dlogits + (long long)r * ldd;
  if (y == ignore_index) {
    for (int c = lane; c < ldd; c += 32) o[c] = from_f<T>(0.f);
    return;
  }
  const float* x = logits + (long long)r * ld;
  const float lse = row_lse[r];
  const float k = gscale / stats[1];
  for (int c = lane; c < ldd; c += 32) {
    float g = 0.f;
    if (c < V) {
      g = expf(x[c] - lse) - smoothing / V;
      if (c == y) g -= 1.f - smoothing;
      g *= k;
    }
    o[c] = from_f<T>(g);
  }
}

// ---- optimiser --------------------------------------------------------------------------------
// stage 1: per-block partial sums of squares;  stage 2: one block finishes -> norm[0] = ||g||_2
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, long long n,
                                                            float* __restrict__ partial) {
  pdl_trigger();
  __shared__ float red[8];
  float s = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) s += g[i] * g[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    partial[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) sumsq_final_kernel(const float* __restrict__ partial, int nblocks,
                                                          float* __restrict__ norm) {
  pdl_trigger();
  __shared__ float red[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < nblocks; i += 256) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) norm[0] = sqrtf(red[0]);
}

// hyper[0]=lr, [1]=beta1, [2]=beta2, [3]=eps, [4]=weight_decay, [5]=1-beta1^t, [6]=1-beta2^t,
// [7]=max_norm (<=0: no clipping), [8]=grad_scale (e.g. 1/world or 1/accumulation)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, bf16* __restrict__ pb, long long n,
                                                   const float* __restrict__ hyper, const float* __restrict__ norm,
                                                   int decoupled, int zero_grad) {
  pdl_trigger();
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const float bc1 = hyper[5], bc2 = hyper[6], max_norm = hyper[7], gs = hyper[8];
  float clip = gs;
  if (max_norm > 0.f && norm) {
    const float total = norm[0] * gs;
    const float c = max_norm / (total + 1e-6f);
    if (c < 1.f) clip *= c;
  }
  const float step = lr / bc1;
  const float rs2 = rsqrtf(bc2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float pi = p[i];
    float gi = g[i] * clip;
    if (!decoupled && wd != 0.f) gi += wd * pi;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    if (decoupled && wd != 0.f) pi *= 1.f - lr * wd;
    pi -= step * mi / (sqrtf(vi) * rs2 + eps);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
    if (pb) pb[i] = __float2bfloat16_rn(pi);
    if (zero_grad) g[i] = 0.f;
  }
}

__global__ void add_u64_kernel(unsigned long long* p, unsigned long long inc) {
  pdl_trigger(); *p += inc; }

}  // namespace trainops

using namespace trainops;

extern "C" int mma_ce_fwd(const float* logits, long long ld, const long long* labels, int rows, int V,
                          float smoothing, long long ignore_index, float* row_loss, float* row_lse, float* stats,
                          cudaStream_t stream) {
  if (rows <= 0) return MMA_ERR_ARG;
  ce_rows_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(logits, ld, labels, rows, V, smoothing, ignore_index, row_loss,
                                                     row_lse);
  MMA_CHECK_LAUNCH();
  ce_reduce_kernel<<<1, 1024, 0, stream>>>(row_loss, labels, rows, ignore_index, stats);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

extern "C" int mma_ce_bwd(const float* logits, long long ld, const long long* labels, const float* row_lse,
                          const float* stats, float gscale, int rows, int V, float smoothing, long long ignore_index,
                          void* dlogits, int d_f32, long long ldd, cudaStream_t stream) {
  if (rows <= 0) return MMA_ERR_ARG;
  if (d_f32)
    ce_bwd_kernel<float><<<(rows + 7) / 8, 256, 0, stream>>>(logits, ld, labels, row_lse, stats, gscale, rows, V,
                                                            smoothing, ignore_index, (float*)dlogits, ldd);
  else
    ce_bwd_kernel<bf16><<<(rows + 7) / 8, 256, 0, stream>>>(logits, ld, labels, row_lse, stats, gscale, rows, V,
                                                           smoothing, ignore_index, (bf16*)dlogits, ldd);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

// workspace: >= 1024 floats
extern "C" int mma_grad_norm(const float* g, long long n, float* workspace, float* norm, cudaStream_t stream) {
  if (n <= 0) return MMA_ERR_ARG;
  long long want = (n / 4 + 255) / 256;
  int blocks = (int)(want < 1 ? 1 : (want > 1024 ? 1024 : want));
  sumsq_partial_kernel<<<blocks, 256, 0, stream>>>(g, n, workspace);
  MMA_CHECK_LAUNCH();
  sumsq_final_kernel<<<1, 256, 0, stream>>>(workspace, blocks, norm);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

extern "C" int mma_adam_step(float* p, float* g, float* m, float* v, void* p_bf16, long long n, const float* hyper,
                             const float* norm, int decoupled, int zero_grad, cudaStream_t stream) {
  if (n <= 0) return MMA_ERR_ARG;
  long long want = (n + 255) / 256;
  int blocks = (int)(want > 148 * 16 ? 148 * 16 : want);
  adam_kernel<<<blocks, 256, 0, stream>>>(p, g, m, v, (bf16*)p_bf16, n, hyper, norm, decoupled, zero_grad);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

// *p += inc on the device (dropout seed of a replayed CUDA graph)
extern "C" int mma_add_u64(unsigned long long* p, unsigned long long inc, cudaStream_t stream) {
  add_u64_kernel<<<1, 1, 0, stream>>>(p, inc);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}
