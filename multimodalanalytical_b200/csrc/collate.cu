// Batch assembly on the device from an HBM-resident, pre-tokenised dataset.
//
// The reference builds every batch on the host (data/datamodules.py:140-351): per-batch tokenizer calls, Python lists
// -> tensors, seq-first transposes, bool masks.  Here the dataset is tokenised once, stored ragged in HBM (flat values
// + row offsets), and a batch is a gather by sample index: the host sends B indices, these kernels write the
// batch-first padded tensors the engine consumes.  HBM-bound: one read of the selected rows, one write of the batch.
//   * collate_tokens   ragged int32 rows -> ids int64 [B, L] (pad_id beyond the row) + validity mask u8 [B, L]
//                      (reference mask = position < length, and all-zero for rows flagged empty: carbon.py:52-56)
//   * collate_target   ragged int32 rows -> decoder input [B, T] = tokens[:-1], decoder mask, labels = tokens[1:] with
//                      pad -> -100 (datamodules.py:178-206 + wrapper.py:365,389), T = L - 1
//   * collate_values   ragged fp32 rows of `width` values -> [B, L, width] (pad_value beyond the row): msms_number
//                      peaks (msms_number.py:50-80), XVal numerical values (multiplets.py)
#include "common.cuh"

namespace col {

__global__ void __launch_bounds__(256) collate_tokens_kernel(const int* __restrict__ flat,
                                                             const long long* __restrict__ offsets,
                                                             const unsigned char* __restrict__ row_valid,
                                                             const int* __restrict__ rows, int B, int L, int pad_id,
                                                             int max_len, long long* __restrict__ ids,
                                                             unsigned char* __restrict__ mask) {
  const int b = blockIdx.y;
  const long long r = rows[b];
  const long long lo = offsets[r];
  const int len = min((int)(offsets[r + 1] - lo), max_len);
  const bool ok = row_valid ? row_valid[r] != 0 : true;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < L; j += gridDim.x * blockDim.x) {
    const bool in = j < len;
    ids[(long long)b * L + j] = in ? flat[lo + j] : pad_id;
    if (mask) mask[(long long)b * L + j] = (in && ok) ? 1 : 0;
  }
}

__global__ void __launch_bounds__(256) collate_target_kernel(const int* __restrict__ flat,
                                                             const long long* __restrict__ offsets,
                                                             const int* __restrict__ rows, int B, int T, int pad_id,
                                                             int max_len, long long* __restrict__ dec_in,
                                                             unsigned char* __restrict__ dec_mask,
                                                             long long* __restrict__ labels) {
  const int b = blockIdx.y;
  const long long r = rows[b];
  const long long lo = offsets[r];
  const int len = min((int)(offsets[r + 1] - lo), max_len);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < T; j += gridDim.x * blockDim.x) {
    const long long o = (long long)b * T + j;
    dec_in[o] = j < len ? flat[lo + j] : pad_id;
    dec_mask[o] = j < len ? 1 : 0;
    long long lab = -100;
    if (j + 1 < len) {
      lab = flat[lo + j + 1];
      if (lab == pad_id) lab = -100;
    }
    labels[o] = lab;
  }
}

__global__ void __launch_bounds__(256) collate_values_kernel(const float* __restrict__ flat,
                                                             const long long* __restrict__ offsets,
                                                             const int* __restrict__ rows, int B, int L, int width,
                                                             float pad_value, int max_len, float* __restrict__ out,
                                                             unsigned char* __restrict__ mask) {
  const int b = blockIdx.y;
  const long long r = rows[b];
  const long long lo = offsets[r];
  const int len = min((int)(offsets[r + 1] - lo), max_len);
  const int n = L * width;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int pos = j / width;
    out[(long long)b * n + j] = pos < len ? flat[lo * width + j] : pad_value;
    if (mask && j % width == 0) mask[(long long)b * L + pos] = pos < len ? 1 : 0;
  }
}

}  // namespace col

extern "C" int mma_collate_tokens(const int* flat, const long long* offsets, const unsigned char* row_valid,
                                  const int* rows, int B, int L, int pad_id, int max_len, long long* ids,
                                  unsigned char* mask, cudaStream_t stream) {
  if (B <= 0 || L <= 0 || !flat || !offsets || !rows || !ids) return MMA_ERR_ARG;
  col::collate_tokens_kernel<<<dim3((L + 255) / 256, B), 256, 0, stream>>>(flat, offsets, row_valid, rows, B, L, pad_id,
                                                                          max_len, ids, mask);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

extern "C" int mma_collate_target(const int* flat, const long long* offsets, const int* rows, int B, int T, int pad_id,
                                  int max_len, long long* dec_in, unsigned char* dec_mask, long long* labels,
                                  cudaStream_t stream) {
  if (B <= 0 || T <= 0 || !flat || !offsets || !rows || !dec_in || !dec_mask || !labels) return MMA_ERR_ARG;
  col::collate_target_kernel<<<dim3((T + 255) / 256, B), 256, 0, stream>>>(flat, offsets, rows, B, T, pad_id, max_len,
                                                                          dec_in, dec_mask, labels);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

extern "C" int mma_collate_values(const float* flat, const long long* offsets, const int* rows, int B, int L, int width,
                                  float pad_value, int max_len, float* out, unsigned char* mask, cudaStream_t stream) {
  if (B <= 0 || L <= 0 || width <= 0 || !flat || !offsets || !rows || !out) return MMA_ERR_ARG;
  col::collate_values_kernel<<<dim3((L * width + 255) / 256, B), 256, 0, stream>>>(flat, offsets, rows, B, L, width,
                                                                                  pad_value, max_len, out, mask);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}
