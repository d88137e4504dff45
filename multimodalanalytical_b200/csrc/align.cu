// Encoder-alignment head pieces (reference: modeling/custom_modeling.py:363-396 networks, :453-475 loss;
// modeling/utils.py:8-22 kl_div / sid).  The head itself is a 4-layer (convolutional: on a length-1 sequence only the
// centre tap of the Conv1d contributes) or 2-layer MLP on the masked mean of the encoder states - those products run
// on the GEMM kernels; here are the pooling, the sigmoid + loss (+ its gradient) and a strided accumulate.
#include "common.cuh"

namespace alignk {

// pooled[b, :] = sum_s mask[b, s] * mem[b, s, :] / sum_s mask[b, s]
template <typename T>
__global__ void __launch_bounds__(128) masked_mean_fwd_kernel(const T* __restrict__ mem, long long ld,
                                                              const unsigned char* __restrict__ mask,
                                                              float* __restrict__ pooled, int S, int d) {
  pdl_trigger();
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  float acc = 0.f, cnt = 0.f;
  for (int s = 0; s < S; ++s) {
    const float m = mask[b * S + s] ? 1.f : 0.f;
    cnt += m;
    acc += m * to_f(mem[(long long)(b * S + s) * ld + c]);
  }
  pooled[(long long)b * d + c] = acc / cnt;
}

// dmem[b, s, :] += mask[b, s] * dpooled[b, :] / count_b
__global__ void __launch_bounds__(128) masked_mean_bwd_kernel(const float* __restrict__ dpooled,
                                                              const unsigned char* __restrict__ mask,
                                                              float* __restrict__ dmem, long long ld, int S, int d) {
  pdl_trigger();
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  float cnt = 0.f;
  for (int s = 0; s < S; ++s) cnt += mask[b * S + s] ? 1.f : 0.f;
  const float g = dpooled[(long long)b * d + c] / cnt;
  for (int s = 0; s < S; ++s)
    if (mask[b * S + s]) dmem[(long long)(b * S + s) * ld + c] += g;
}

// pred = sigmoid(z); loss = mae | mse | sid (pred, target); out[0] = loss, out[1] = lm_loss + lambda * loss;
// dz = dscale * lambda * dloss/dz (optional).  One block: a fixed reduction order makes the scalar reproducible.
__global__ void __launch_bounds__(1024) align_loss_kernel(const float* __restrict__ z, long long ldz,
                                                          const float* __restrict__ target, long long ldt, int rows,
                                                          int cols, int kind, float lambda,
                                                          const float* __restrict__ lm_loss, float* __restrict__ out,
                                                          float* __restrict__ dz, long long lddz, float dscale) {
  pdl_trigger();
  __shared__ float red[32];
  const long long n = (long long)rows * cols;
  const float inv_n = 1.0f / (float)n, inv_b = 1.0f / (float)rows;
  const float eps = 1e-16f;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const float zz = z[(long long)r * ldz + c];
    const float p = 1.0f / (1.0f + expf(-zz));
    const float t = target[(long long)r * ldt + c];
    float l, dl;  // loss term and d(loss)/d(pred)
    if (kind == 0) {
      const float e = p - t;
      l = fabsf(e) * inv_n;
      dl = (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f)) * inv_n;
    } else if (kind == 1) {
      const float e = p - t;
      l = e * e * inv_n;
      dl = 2.f * e * inv_n;
    } else {
      const float pc = fmaxf(p, eps), qc = fmaxf(t, eps);
      const float lr = logf(pc / qc);
      l = (pc * lr - qc * lr) * inv_b;  // p log(p/q) + q log(q/p)
      dl = p > eps ? (lr + 1.0f - qc / pc) * inv_b : 0.f;
    }
    acc += l;
    if (dz) dz[(long long)r * lddz + c] = dscale * lambda * dl * p * (1.0f - p);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      out[0] = v;
      out[1] = (lm_loss ? *lm_loss : 0.f) + lambda * v;
    }
  }
}

__global__ void add_strided_kernel(float* __restrict__ dst, long long stride, const float* __restrict__ src, long long n) {
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i * stride] += src[i];
}

}  // namespace alignk

extern "C" int mma_masked_mean_fwd(const void* mem, int mem_ty, long long ld, const unsigned char* mask, float* pooled,
                                   int B, int S, int d, cudaStream_t stream) {
  if (B <= 0 || S <= 0 || d <= 0) return MMA_ERR_ARG;
  dim3 grid((d + 127) / 128, B);
  if (mem_ty == MMA_F32) alignk::masked_mean_fwd_kernel<float><<<grid, 128, 0, stream>>>((const float*)mem, ld, mask, pooled, S, d);
  else alignk::masked_mean_fwd_kernel<bf16><<<grid, 128, 0, stream>>>((const bf16*)mem, ld, mask, pooled, S, d);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

extern "C" int mma_masked_mean_bwd(const float* dpooled, const unsigned char* mask, float* dmem, long long ld, int B,
                                   int S, int d, cudaStream_t stream) {
  if (B <= 0 || S <= 0 || d <= 0) return MMA_ERR_ARG;
  dim3 grid((d + 127) / 128, B);
  alignk::masked_mean_bwd_kernel<<<grid, 128, 0, stream>>>(dpooled, mask, dmem, ld, S, d);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

// kind: 0 mae, 1 mse, 2 sid.  out: float[2] = {alignment loss, lm_loss + lambda * alignment loss}.
extern "C" int mma_align_loss(const float* z, long long ldz, const float* target, long long ldt, int rows, int cols,
                              int kind, float lambda, const float* lm_loss, float* out, float* dz, long long lddz,
                              float dscale, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0 || kind < 0 || kind > 2) return MMA_ERR_ARG;
  alignk::align_loss_kernel<<<1, 1024, 0, stream>>>(z, ldz, target, ldt, rows, cols, kind, lambda, lm_loss, out, dz, lddz, dscale);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

// dst[i * stride] += src[i]  (gradient of the centre tap of a Conv1d weight [C_out, C_in, k])
extern "C" int mma_add_strided(float* dst, long long stride, const float* src, long long n, cudaStream_t stream) {
  if (n <= 0) return MMA_OK;
  alignk::add_strided_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(dst, stride, src, n);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}
