// fp32 SIMT GEMM with arbitrary strides: the fp32-parity path (logits/loss to 1e-5, identical beam
// sequences) and the odd-shaped small products (patch embeddings with K = 75/125/1/2) that TMA's
// 16-byte stride rule excludes.  Shares the epilogue with the tcgen05 kernel (common.cuh).
//
//   C[m,n] = epi( sum_k A(m,k) * B(n,k) ),  A(m,k) = A[m*sam + k*sak],  B(n,k) = B[n*sbn + k*sbk]
#include "common.cuh"

namespace simt {
constexpr int TM = 64, TN = 64, TK = 16;

template <typename TA, typename TB>
__global__ void __launch_bounds__(256) sgemm_kernel(const TA* __restrict__ A, long long sam, long long sak,
                                                    const TB* __restrict__ B, long long sbn, long long sbk, int M,
                                                    int N, int K, int k_per_split, Epi ep) {
  pdl_trigger();
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const long long m0 = (long long)blockIdx.y * TM;
  const long long n0 = (long long)blockIdx.x * TN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // split-K over grid.z (host guarantees an atomic EPI_ACCUM epilogue when gridDim.z > 1)
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  for (int k0 = k_begin; k0 < k_end; k0 += TK) {
#pragma unroll
    for (int e = t; e < TM * TK; e += 256) {
      int m, k;
      if (sak == 1) { k = e % TK; m = e / TK; } else { m = e % TM; k = e / TM; }
      const long long gm = m0 + m;
      const int gk = k0 + k;
      As[k][m] = (gm < M && gk < k_end) ? to_f(A[gm * sam + gk * sak]) : 0.f;
    }
#pragma unroll
    for (int e = t; e < TN * TK; e += 256) {
      int n, k;
      if (sbk == 1) { k = e % TK; n = e / TK; } else { n = e % TN; k = e / TN; }
      const long long gn = n0 + n;
      const int gk = k0 + k;
      Bs[k][n] = (gn < N && gk < k_end) ? to_f(B[gn * sbn + gk * sbk]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long row = m0 + ty * 4 + i;
    if (row < M) epilogue_store<4, -1, false>(ep, row, (int)(n0 + tx * 4), N, acc[i]);
  }
}
}  // namespace simt

extern "C" int mma_gemm_simt(const void* A, int a_type, long long sam, long long sak, const void* B, int b_type,
                             long long sbn, long long sbk, int M, int N, int K, const Epi* ep, int splits,
                             cudaStream_t stream) {
  using namespace simt;
  if (M <= 0 || N <= 0 || K <= 0 || !ep) return MMA_ERR_ARG;
  if (splits < 1) splits = 1;
  if (splits > 1 && !(ep->kind == EPI_ACCUM && ep->accumulate == 2)) return MMA_ERR_ARG;
  int kps = ((K + splits - 1) / splits + TK - 1) / TK * TK;
  splits = (K + kps - 1) / kps;
  dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM, splits);
  if (a_type == MMA_F32 && b_type == MMA_F32)
    sgemm_kernel<float, float><<<grid, 256, 0, stream>>>((const float*)A, sam, sak, (const float*)B, sbn, sbk, M, N, K, kps, *ep);
  else if (a_type == MMA_BF16 && b_type == MMA_BF16)
    sgemm_kernel<bf16, bf16><<<grid, 256, 0, stream>>>((const bf16*)A, sam, sak, (const bf16*)B, sbn, sbk, M, N, K, kps, *ep);
  else if (a_type == MMA_F32 && b_type == MMA_BF16)
    sgemm_kernel<float, bf16><<<grid, 256, 0, stream>>>((const float*)A, sam, sak, (const bf16*)B, sbn, sbk, M, N, K, kps, *ep);
  else
    sgemm_kernel<bf16, float><<<grid, 256, 0, stream>>>((const bf16*)A, sam, sak, (const float*)B, sbn, sbk, M, N, K, kps, *ep);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}
