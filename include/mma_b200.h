/* C ABI of libmma_b200.so - the sm_100a kernels behind the spectra -> SMILES hot path.
 *
 * Conventions ("C-ABI op layer"):
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller (PyTorch's caching
 *     allocator in this repo); kernels never allocate or free; outputs are pre-allocated;
 *   - every function launches on the cudaStream_t passed last and returns immediately with an int status:
 *     0 ok, -1 bad argument, -2 launch failure, -3 unsupported shape/type, -4 driver / tensor-map failure;
 *   - no exceptions cross the boundary, no global mutable state apart from a mutex-guarded TMA-descriptor cache;
 *   - element-type tags: MMA_BF16 = 0, MMA_F32 = 1;  `ld*` arguments are row pitches in ELEMENTS.
 *
 * The reference has no native layer: each entry point replaces the ATen / cuBLAS /
 * transformers call the reference makes implicitly at the cited line of /root/reference/src/analytical_fm.
 */
#ifndef MMA_B200_H
#define MMA_B200_H

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMA_BF16 0
#define MMA_F32 1
#define MMA_SITE_SEED_INDIRECT 0x80000000u /* in any `site` argument: `seed` holds a device pointer to the seed */

/* GEMM epilogues, shared by the tcgen05 and the SIMT kernel */
enum {
  EPI_STORE = 0,   /* out = acc*alpha + bias                                   (nn.Linear)                    */
  EPI_GELU = 1,    /* z = acc+bias; out2 = z; out = drop(gelu(z))              (linear1 + exact-erf GELU)     */
  EPI_RESID = 2,   /* out = resid + drop(acc + bias)                           (out_proj / linear2 + residual)*/
  EPI_DGELU = 3,   /* out = acc * dropmask * gelu'(aux)                        (backward of EPI_GELU)         */
  EPI_GLU_MUL = 4, /* z2 = acc+bias; out2 = z2; out = drop(gelu(aux) * z2)     (GLU gate, custom_modeling.py:137-152) */
  EPI_DGLU = 5,    /* backward of EPI_GLU_MUL: out = d z1, out2 = d z2                                          */
  EPI_ACCUM = 6,   /* out(f32) (+)= acc*alpha; accumulate 0 overwrite / 1 add / 2 atomic add (split-K wgrad)   */
  EPI_RELU = 7,    /* out = relu(acc + bias)                                   (patch-embedding MLPs, utils.py:121-134) */
  EPI_DRELU = 8    /* out = acc * [aux > 0]                                                                   */
};

typedef struct Epi {
  int kind;
  int out_f32, aux_f32, resid_f32; /* element types of out/out2, aux/aux2, resid */
  void* out;
  void* out2;
  const float* bias;
  const void* resid;
  const void* aux;
  const void* aux2;
  long long ldo, ldo2, ldr, lda, lda2;
  float p_drop; /* dropout probability of this site (0 = off); mask = counter hash of (seed, site, element index) */
  float alpha;
  unsigned long long seed;
  unsigned int site; /* bit 31 (MMA_SITE_SEED_INDIRECT): `seed` is the device address of the 64-bit seed */
  int accumulate;
  long long drop_ld; /* logical row width indexing the dropout stream */
} Epi;

/* ---- dense contractions ------------------------------------------------------------------------------------
 * C[M,N] = epi(A_op[M,K] * B_op[N,K]^T).  a_mn / b_mn = 0: operand memory is [rows, K] row-major (K-major);
 * = 1: memory is [K, rows] row-major (MN-major), which yields dgrad (dy * W) and wgrad (dy^T * x) without
 * transposed copies.  bf16 operands, fp32 accumulation in TMEM (tcgen05.mma, TMA-fed, persistent CTAs).
 * Replaces: nn.Linear / F.linear inside nn.MultiheadAttention, linear1/linear2/gate, token_ff
 * (custom_modeling.py:108-199, 418, 486) and their autograd backward.                                          */
int mma_gemm_bf16(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, int M, int N,
                  int K, const Epi* ep, int splits, int max_ctas, cudaStream_t stream);
/* Grouped weight gradients of one layer in one persistent launch (count <= 8): out[g][Nout,Kin] += dy[g]^T x[g] and
 * dbias[g][Nout] += colsum(dy[g]) (bias gradient fused as an all-ones MMA column).  dy[g]: bf16 [R, Nout] pitch
 * lddy[g]; x[g]: bf16 [R, Kin] pitch ldx[g]; the arrays are HOST arrays of device pointers / sizes.               */
int mma_wgrad_group(int count, const void* const* dy, const long long* lddy, const void* const* x,
                    const long long* ldx, float* const* out, const long long* ldo, float* const* dbias,
                    const int* Nout, const int* Kin, const int* R, cudaStream_t stream);
/* One accumulation over two operand pairs, C[M,N] = epi(A1 B1_op^T + A2 B2_op^T) (reduction lengths K1, K2; A1/A2
 * K-major, B1/B2 both K-major or both MN-major): the gated FFN's dh = dz1 W1 + dz2 Wg (backward of
 * custom_modeling.py:137-152,184-199) without an fp32 accumulate round trip.  CTA-pair kernel only:
 * MMA_ERR_UNSUPPORTED (-3) outside its envelope - run two mma_gemm_bf16 with EPI_ACCUM instead.                  */
int mma_gemm2_dual(const void* A1, long long lda1, const void* B1, long long ldb1, const void* A2, long long lda2,
                   const void* B2, long long ldb2, int b_mn, int M, int N, int K1, int K2, const Epi* ep,
                   cudaStream_t stream);
/* Gated FFN, forward (custom_modeling.py:137-152,184-199: `gelu(linear1(x)) * gate(x)` + dropout) as ONE CTA-pair
 * tcgen05 launch over W1 and Wg: a[M,N] = drop(gelu(h W1^T + b1) * (h Wg^T + bg)); z1 / z2 (both or neither) receive
 * the bf16 pre-activations for the backward.  h bf16 [M,K]; W1, Wg bf16 [N,K]; b1, bg fp32 [N].  Dropout stream index
 * of element (row, col) = row * N + col.  MMA_ERR_UNSUPPORTED (-3) outside the envelope (alignment, N % 16, too few
 * tiles): run mma_gemm_bf16 twice (EPI_STORE, then EPI_GLU_MUL).                                                  */
int mma_ffn_glu_fwd(const void* h, long long ldh, const void* W1, long long ldw1, const void* Wg, long long ldwg,
                    const float* b1, const float* bg, int M, int N, int K, void* a, long long lda, void* z1,
                    long long ldz1, void* z2, long long ldz2, float p_drop, unsigned long long seed, unsigned int site,
                    cudaStream_t stream);
/* Gated FFN, backward through the gate: da = (dy W2) * dropmask with W2 = linear2.weight [K, N] as stored;
 * dz1 = da * z2 * gelu'(z1), dz2 = da * gelu(z1)  (all bf16 [M,N]).  -3 outside the envelope: mma_gemm_bf16 + EPI_DGLU. */
int mma_ffn_dglu(const void* dy, long long lddy, const void* W2, long long ldw2, int M, int N, int K, const void* z1,
                 long long ldz1, const void* z2, long long ldz2, void* dz1, long long lddz1, void* dz2,
                 long long lddz2, float p_drop, unsigned long long seed, unsigned int site, long long drop_ld,
                 cudaStream_t stream);
/* fp32 SIMT GEMM with arbitrary element strides (fp32 parity mode; patch embeddings with K = 75/125/1/2,
 * modeling/utils.py:119-134).  A(m,k) = A[m*sam + k*sak], B(n,k) = B[n*sbn + k*sbk].                          */
int mma_gemm_simt(const void* A, int a_type, long long sam, long long sak, const void* B, int b_type,
                  long long sbn, long long sbk, int M, int N, int K, const Epi* ep, int splits, cudaStream_t stream);

/* ---- embedding / LayerNorm / reductions (HBM-bound) ----------------------------------------------------------
 * nn.Embedding gather (+ XVal scale), modeling/utils.py:102-106,154-160 */
int mma_gather_rows(const long long* ids, const float* scale, const float* table, float* out, int rows, int d,
                    cudaStream_t stream);
/* embedding backward: dtable[ids[r]] += g[r]*scale[r], padding row skipped (padding_idx) */
int mma_scatter_add_rows(const long long* ids, const float* scale, const float* g, float* dtable, int rows, int d,
                         long long pad_idx, cudaStream_t stream);
/* LayerNorm forward fused with the positional-encoding add and the multimodal concat-by-offset write
 * (modeling/utils.py:165-180; pre-LN norms custom_modeling.py:129,176,350,399).  Output row of input row r is
 * (r / group) * out_group_stride + out_offset + r % group; gamma == NULL skips the norm.                        */
int mma_ln_fwd(const void* x, int x_f32, long long ldx, const float* gamma, const float* beta, float eps, void* y,
               int y_f32, long long ldy, void* y2, int y2_f32, long long ldy2, const float* add, long long ld_add,
               int rows, int d, int group, int out_group_stride, int out_offset, cudaStream_t stream);
/* LayerNorm backward (+ residual-gradient add, + dropout-masked low-precision copy for the previous branch) */
int mma_ln_bwd(const void* dy, int dy_f32, long long lddy, int group, int in_group_stride, int in_offset,
               const void* x, int x_f32, long long ldx, const float* gamma, float eps, const float* dres,
               long long lddres, float* dx, long long lddx, void* dxb, int dxb_f32, long long lddxb, float p_drop,
               unsigned long long seed, unsigned int site, float* dgamma, float* dbeta, int rows, int d,
               cudaStream_t stream);
/* Residual product fused with the LayerNorm that follows it (d_model = N = 512, CTA-pair tcgen05 kernel):
 *   ep->out (fp32) = ep->resid + drop(A W^T + bias);   h (bf16) = LayerNorm(ep->out) * gamma + beta.
 * Replaces nn.Linear + dropout + residual add + the next sub-layer's nn.LayerNorm (custom_modeling.py:129-152,176-199).
 * MMA_ERR_UNSUPPORTED (-3) outside that envelope: run mma_gemm_bf16 and mma_ln_fwd instead.                      */
int mma_gemm2_resid_ln(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, const Epi* ep,
                       const float* gamma, const float* beta, float eps, void* h, long long ldh, cudaStream_t stream);
/* out[c] += sum_r in[r,c]  (bias gradients) */
int mma_colsum(const void* in, int in_f32, long long ld, float* out, int rows, int cols, cudaStream_t stream);
int mma_cast_f32_bf16(const float* in, void* out, long long n, cudaStream_t stream);
int mma_cast_bf16_f32(const void* in, float* out, long long n, cudaStream_t stream);

/* PatchPreprocessor on the device (data/preprocessing/patches.py:54-107): standardise, (interpolation == slice,
 * see rowops.cu), trim, patch.  out [B, P, ps] batch-first; pad [B, P] (optional): patch-sum == 0 when masking,
 * else missing[b].  hop = ps / overlap.                                                                          */
int mma_patchify(const float* raw, long long ld, int offset, float mean, float std, float* out, unsigned char* pad,
                 const unsigned char* missing, int masking, int B, int P, int ps, int hop, cudaStream_t stream);
/* same, for spectra resident in HBM as a dataset table: batch element b reads row rows[b] of raw / missing */
int mma_patchify_rows(const float* raw, long long ld, const int* rows, int offset, float mean, float std, float* out,
                      unsigned char* pad, const unsigned char* missing, int masking, int B, int P, int ps, int hop,
                      cudaStream_t stream);
/* PatchPreprocessor(derivative=True) (patches.py:41-46,91-95): out [B, P + Pd, ps] = the P standardised patches followed
 * by the Pd = n_use / ps patches of torch.gradient over the n_use raw points the preprocessor sees from `offset` on
 * (central differences, one-sided at both ends; not standardised, never overlapping); pad [B, P + Pd].  rows: NULL or
 * the dataset rows of the batch.                                                                                   */
int mma_patchify_deriv(const float* raw, long long ld, const int* rows, int offset, int n_use, float mean, float std,
                       float* out, unsigned char* pad, const unsigned char* missing, int masking, int B, int P, int Pd,
                       int ps, int hop, cudaStream_t stream);

/* ---- batch assembly from an HBM-resident pre-tokenised dataset (replaces the host collator,
 * data/datamodules.py:140-351).  Ragged storage: flat values + int64 row offsets [N+1]; rows [B] are
 * the sample indices of the batch; rows longer than max_len are truncated (tokenizer truncation=True).          */
/* ids int64 [B, L] (pad_id beyond the row), mask u8 [B, L] = position < length and row_valid[row] (optional)    */
int mma_collate_tokens(const int* flat, const long long* offsets, const unsigned char* row_valid, const int* rows,
                       int B, int L, int pad_id, int max_len, long long* ids, unsigned char* mask,
                       cudaStream_t stream);
/* teacher forcing (datamodules.py:178-206, wrapper.py:365,389): dec_in = tokens[:-1], labels = tokens[1:] with
 * pad -> -100, dec_mask = position < length; all [B, T], T = padded length - 1                                   */
int mma_collate_target(const int* flat, const long long* offsets, const int* rows, int B, int T, int pad_id,
                       int max_len, long long* dec_in, unsigned char* dec_mask, long long* labels,
                       cudaStream_t stream);
/* fp32 rows of `width` values per position -> out [B, L, width] (pad_value beyond the row), mask u8 [B, L] opt.  */
int mma_collate_values(const float* flat, const long long* offsets, const int* rows, int B, int L, int width,
                       float pad_value, int max_len, float* out, unsigned char* mask, cudaStream_t stream);

/* ---- encoder-alignment head (custom_modeling.py:363-396 networks, :453-475 masked mean pool + loss; LOSS_FACTORY
 * :15; modeling/utils.py:8-22 kl_div / sid).  The MLP / centre-tap conv products run on the GEMM entry points.  */
/* pooled[b,:] = sum_s mask[b,s] mem[b,s,:] / sum_s mask[b,s]   (mask 1 = real token) */
int mma_masked_mean_fwd(const void* mem, int mem_f32, long long ld, const unsigned char* mask, float* pooled, int B,
                        int S, int d, cudaStream_t stream);
/* dmem[b,s,:] += mask[b,s] dpooled[b,:] / count_b */
int mma_masked_mean_bwd(const float* dpooled, const unsigned char* mask, float* dmem, long long ld, int B, int S,
                        int d, cudaStream_t stream);
/* pred = sigmoid(z); kind 0 mae, 1 mse, 2 sid; out[0] = loss, out[1] = *lm_loss + lambda * loss;
 * dz (optional) = dscale * lambda * dloss/dz */
int mma_align_loss(const float* z, long long ldz, const float* target, long long ldt, int rows, int cols, int kind,
                   float lambda, const float* lm_loss, float* out, float* dz, long long lddz, float dscale,
                   cudaStream_t stream);
/* dst[i * stride] += src[i] */
int mma_add_strided(float* dst, long long stride, const float* src, long long n, cudaStream_t stream);

/* ---- attention (F.scaled_dot_product_attention inside nn.MultiheadAttention; masks custom_modeling.py:233-234,
 * 299-318).  q/k/v/o are [B*L, ld] with head h at columns [h*dh, (h+1)*dh); kmask[B,Lk] 1 = real token.         */
int mma_attn_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                 const unsigned char* kmask, void* o, long long ldo, float* lse, int B, int H, int Lq, int Lk, int dh,
                 int causal, float scale, float p_drop, unsigned long long seed, unsigned int site, int type,
                 cudaStream_t stream);
int mma_attn_bwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                 const unsigned char* kmask, const void* o, long long ldo, const float* lse, const void* dout,
                 long long lddo, void* dq, long long lddq, void* dk, long long lddk, void* dv, long long lddv, int B,
                 int H, int Lq, int Lk, int dh, int causal, float scale, float p_drop, unsigned long long seed,
                 unsigned int site, int type, cudaStream_t stream);

/* tensor-core variants (mma.sync m16n8k16, bf16, head dim 64); dsum = fp32 workspace [B*H*Lq] */
int mma_attn_fwd_tc(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                    const unsigned char* kmask, void* o, long long ldo, float* lse, int B, int H, int Lq, int Lk,
                    int causal, float scale, float p_drop, unsigned long long seed, unsigned int site,
                    cudaStream_t stream);
int mma_attn_bwd_tc(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                    const unsigned char* kmask, const void* o, long long ldo, const float* lse, float* dsum,
                    const void* dout, long long lddo, void* dq, long long lddq, void* dk, long long lddk, void* dv,
                    long long lddv, int B, int H, int Lq, int Lk, int causal, float scale, float p_drop,
                    unsigned long long seed, unsigned int site, cudaStream_t stream);

/* tcgen05 / TMEM single-tile variants (bf16, head dim 64, Lq <= 128 and Lk <= 128; two (batch, head) problems are
 * packed per 128x128 tile when both lengths are <= 64) */
int mma_attn_fwd_t5(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                    const unsigned char* kmask, void* o, long long ldo, float* lse, int B, int H, int Lq, int Lk,
                    int causal, float scale, float p_drop, unsigned long long seed, unsigned int site,
                    cudaStream_t stream);
int mma_attn_bwd_t5(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                    const unsigned char* kmask, const void* o, long long ldo, const float* lse, const void* dout,
                    long long lddo, void* dq, long long lddq, void* dk, long long lddk, void* dv, long long lddv,
                    int B, int H, int Lq, int Lk, int causal, float scale, float p_drop, unsigned long long seed,
                    unsigned int site, cudaStream_t stream);

/* blocked tcgen05 / TMEM variants for 128 < L <= 512 (the multimodal encoder, S ~ 200-300): the score matrix is cut
 * into 128 x 128 single-shot tile problems; forward merges the key blocks' (O, log-sum-exp), backward sums per-block fp32
 * partial gradients (deterministic, no atomics).  Caller-owned workspaces with nqb = ceil(Lq/128), nkb = ceil(Lk/128):
 * ws_o bf16 [nkb][B*Lq][H*64], ws_lse fp32 [nkb][B*H*Lq]; ws_dq bf16 [nkb][B*Lq][H*64], ws_dk / ws_dv bf16 [nqb][B*Lk][H*64]. */
int mma_attn_fwd_t5b(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                     const unsigned char* kmask, void* o, long long ldo, float* lse, void* ws_o, float* ws_lse, int B,
                     int H, int Lq, int Lk, int causal, float scale, float p_drop, unsigned long long seed,
                     unsigned int site, cudaStream_t stream);
int mma_attn_bwd_t5b(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                     const unsigned char* kmask, const void* o, long long ldo, const float* lse, const void* dout,
                     long long lddo, void* dq, long long lddq, void* dk, long long lddk, void* dv, long long lddv,
                     void* ws_dq, void* ws_dk, void* ws_dv, int B, int H, int Lq, int Lk, int causal, float scale,
                     float p_drop, unsigned long long seed, unsigned int site, cudaStream_t stream);

/* ---- loss (nn.CrossEntropyLoss, custom_modeling.py:490-491; ignore_index -100 set at wrapper.py:389) --------- */
int mma_ce_fwd(const float* logits, long long ld, const long long* labels, int rows, int V, float smoothing,
               long long ignore_index, float* row_loss, float* row_lse, float* stats, cudaStream_t stream);
int mma_ce_bwd(const float* logits, long long ld, const long long* labels, const float* row_lse, const float* stats,
               float gscale, int rows, int V, float smoothing, long long ignore_index, void* dlogits, int d_f32,
               long long ldd, cudaStream_t stream);

/* ---- optimiser (torch.optim.Adam/AdamW + clip_grad_norm_, wrapper.py:329-344, trainer/trainer.py:65) --------- */
int mma_grad_norm(const float* g, long long n, float* workspace, float* norm, cudaStream_t stream);
int mma_adam_step(float* p, float* g, float* m, float* v, void* p_bf16, long long n, const float* hyper,
                  const float* norm, int decoupled, int zero_grad, cudaStream_t stream);

int mma_add_u64(unsigned long long* p, unsigned long long inc, cudaStream_t stream);

/* ---- data-parallel optimiser step over NVLink peer memory (replaces Lightning DDP's NCCL all-reduce + the replicated
 * optimiser step, trainer/trainer.py:58-71, wrapper.py:329-344): reduce-scatter of the gradients by P2P loads, Adam on
 * this rank's shard of the fp32 master weights / moments, bf16 weights written into every peer's mirror by P2P stores.
 * The `peer_*` arguments are HOST arrays of `world` device pointers (symmetric allocations, index = rank); the barrier's
 * flag arrays are int[16] per rank, zero-initialised, `epoch_ctr` a zero-initialised device int.  Sequence per step:
 * barrier, reduce_shard, barrier, adam_shard, barrier.  mc_g / mc_pb: NVLS multicast addresses of the gradient buffer /
 * bf16 mirror (multimem.ld_reduce sums inside the NVSwitch, multimem.st is replicated by it) or NULL (explicit P2P).  */
int mma_p2p_barrier(const void* const* peer_flags, int* epoch_ctr, int world, int rank, cudaStream_t stream);
int mma_p2p_reduce_shard(const void* const* peer_g, const void* mc_g, int world, int rank, long long lo, long long hi,
                         float* workspace, float* sumsq_out, cudaStream_t stream);
int mma_p2p_adam_shard(float* p, float* g, float* m, float* v, const void* const* peer_pb, void* mc_pb,
                       const void* const* peer_sumsq, int world, int rank, long long n, long long lo, long long hi,
                       const float* hyper, int decoupled, cudaStream_t stream);

/* ---- KV-cached decoding (replaces transformers generate(use_cache=False), wrapper.py:443-451) ---------------- */
int mma_decode_embed(const int* tok, const float* table, const float* gamma, const float* beta, float eps,
                     const float* pos, const int* cur_len, float* out, int rows, int d, cudaStream_t stream);
/* K/V cache layout [R][Lmax][H*dh]; anc [2][R][Lmax] ancestor rows (buffer cur_len & 1 is live); beams: rows
 * [s*beams, (s+1)*beams) belong to one spectrum (>= 4: they may be grouped in one CTA so that shared ancestors hit in L1) */
int mma_decode_self_attn(const void* q, long long ldq, const void* knew, const void* vnew, long long ldkv,
                         void* kcache, void* vcache, const int* anc, const int* cur_len, void* o, long long ldo, int R,
                         int H, int dh, int Lmax, float scale, int type, int beams, cudaStream_t stream);
int mma_decode_cross_attn(const void* q, long long ldq, const void* kmem, const void* vmem, long long ldm,
                          const unsigned char* kmask, const int* cur_len, void* o, long long ldo, int R, int H, int dh,
                          int S, int beams, float scale, int type, cudaStream_t stream);
/* Small-batch decode product (R <= 512 rows in blocks of <= 64: a few spectra x beams): out[R,N] = epilogue(LN?(x)[R,K] w[N,K]^T + bias) in
 * one launch on mma.sync with the weight tile as the 16-row operand (every weight element read once), LayerNorm of the fp32
 * residual stream as prologue, bias / GELU / gate / residual as epilogue.  Replaces nn.LayerNorm + nn.Linear (+ gelu, gate,
 * residual) of a decoder layer (custom_modeling.py:155-199) when the tcgen05 tiles would be > 90 % padding.
 * x: fp32 (x_f32 = 1; LayerNorm when gamma != NULL) or bf16; kind 0 store, 1 GELU, 2 out = resid + result, 3 gated
 * (gelu(x w^T + bias) * (x w2^T + bias2)).  K % 256 == 0; -3 otherwise / for R > 512.                              */
int mma_small_linear(const void* x, int x_f32, long long ldx, const float* gamma, const float* beta, float eps,
                     const void* w, const void* w2, long long ldw, const float* bias, const float* bias2,
                     const float* resid, long long ldr, void* out, int out_f32, long long ldo, int R, int N, int K,
                     int kind, cudaStream_t stream);
/* The WHOLE decoder step of a few spectra (logits of every row) in ONE launch: token embedding, every decoder layer
 * (LayerNorm + QKV, KV-cached self-attention, out-projection + residual, LayerNorm + cross-query, cross-attention,
 * out-projection + residual, LayerNorm + FFN-1 (+ gate), FFN-2 + residual), final LayerNorm + LM head.  A thread-block
 * cluster of `cluster_size` CTAs (8 or 16) owns `rows_per_cluster` consecutive rows (a multiple of `beams`, <= 16)
 * and separates dependent phases with barrier.cluster, activations travel between the CTAs' shared memories
 * (st.shared::cluster); clusters are independent.  Replaces the ~50 launches per step of
 * the cached decode of HFWrapper.generate (wrapper.py:409-453 -> custom_modeling.py:155-199,418-486) when rows are few.
 * bf16 weights [out, in] contiguous, fp32 biases / norms / residual stream; d = 512, f = 2048, head dim 64
 * (MMA_ERR_UNSUPPORTED otherwise).  Same buffers and semantics as the per-op entry points above.                  */
#define MMA_DECODE_MAX_LAYERS 12
typedef struct MmaDecodeLayer {
  const void *w_qkv, *w_so, *w_cq, *w_co, *w_f1, *w_fg, *w_f2;        /* bf16: self in_proj [3d,d], self out [d,d], cross
                                                                         q [d,d], cross out [d,d], linear1 [f,d], gate
                                                                         [f,d] (NULL when ungated), linear2 [d,f] */
  const float *b_qkv, *b_so, *b_cq, *b_co, *b_f1, *b_fg, *b_f2;
  const float *n1g, *n1b, *n2g, *n2b, *n3g, *n3b;                     /* norm1 / norm2 / norm3 weight, bias */
  void *kc, *vc;                                                      /* bf16 self K / V cache [R][Lmax][d] */
  const void* kvmem;                                                  /* bf16 cross K | V [B * S][2 d] */
} MmaDecodeLayer;
typedef struct MmaDecodeStep {
  MmaDecodeLayer layer[MMA_DECODE_MAX_LAYERS];
  const int* tok;          /* [R] token fed to this step */
  const float* emb;        /* [V, d] target embedding table */
  const float *emb_g, *emb_b; /* per-modality LayerNorm of the embedding (NULL: none) */
  const float* pos;        /* [Lmax, d] positional rows */
  const int* cur_len;      /* device scalar: tokens so far (position of this step = cur_len - 1) */
  const float *fin_g, *fin_b; /* decoder.norm */
  const void* w_lm;        /* bf16 [V, d] */
  const float* b_lm;
  float *x, *xa, *xb;      /* unused (workspaces of the per-op path; kept so both paths fill one struct) */
  void *qkv, *att, *q, *a; /* unused */
  float* logits;           /* fp32 [R, ldv] */
  const int* anc;          /* [2][R][Lmax] ancestor rows (NULL: identity, greedy) */
  const unsigned char* enc_mask; /* [B][S], 1 = real memory position */
  unsigned long long* dbg_times; /* NULL, or [64] device words: %globaltimer of cluster 0 at every phase boundary */
  long long ldv;
  int layers, R, rows_per_cluster, beams, d, f, H, Lmax, S, V, gated;
  float eps, scale;
} MmaDecodeStep;
int mma_decode_step(const void* args /* const MmaDecodeStep* (host memory) */, int cluster_size, cudaStream_t stream);
/* NOT a status: the number of clusters of `cluster_size` CTAs of the step kernel that can be resident at once on the current
 * device (0: that cluster size cannot be scheduled).  More clusters than that run in waves.                        */
int mma_decode_step_max_clusters(int cluster_size);
/* one beam-search step for B spectra x K beams (transformers GenerationMixin._beam_search semantics) */
int mma_beam_step(const float* logits, long long ldl, const float* extra_bias, int B, int K, int V, int L, int pad_id,
                  int eos_id, const int* cur_len, int* run_seq, int* fin_seq, float* run_score, float* fin_score,
                  unsigned char* fin_flag, int* fin_len, unsigned char* improvable, unsigned char* all_hit, int* anc,
                  int* next_tok, int* parent_row, cudaStream_t stream);
/* greedy step (GenerationMixin._sample with do_sample=False) */
int mma_greedy_step(const float* logits, long long ldl, const float* extra_bias, int R, int V, int L, int pad_id,
                    int eos_id, const int* cur_len, int* seq, unsigned char* unfinished, int* next_tok,
                    cudaStream_t stream);
int mma_advance(int* cur_len, cudaStream_t stream);

/* ---- logits processors inside generate (wrapper.py:443-451 `logits_processor=`) ---------------
 * `_ex` steps: prenorm != 0 -> `logits` already holds the processed scores (log_softmax -> ForcedEOS -> processors),
 * the step only adds the running score and selects.  g_cur != NULL fuses GuidedFormulaProcessor.__call__
 * (generation/logit_processors.py:89-152) into the step: g_cur [rows][n_atoms] = atom counts of each running
 * hypothesis (host chemistry), g_tgt [spectra][n_atoms] = target formula, g_tok_atoms [V] = bit e set when the token
 * adds an atom of element e (logit_processors.py:46-62); <eos> := 0 when all n_atoms counts match, := -inf while any
 * is short, and a token := -inf when it would push one of the first n_check elements over its target. */
int mma_beam_step_ex(const float* logits, long long ldl, const float* extra_bias, int B, int K, int V, int L,
                     int pad_id, int eos_id, const int* cur_len, int* run_seq, int* fin_seq, float* run_score,
                     float* fin_score, unsigned char* fin_flag, int* fin_len, unsigned char* improvable,
                     unsigned char* all_hit, int* anc, int* next_tok, int* parent_row, int prenorm, const int* g_cur,
                     const int* g_tgt, const unsigned* g_tok_atoms, int n_atoms, int n_check, cudaStream_t stream);
int mma_greedy_step_ex(const float* logits, long long ldl, const float* extra_bias, int R, int V, int L, int pad_id,
                       int eos_id, const int* cur_len, int* seq, unsigned char* unfinished, int* next_tok, int prenorm,
                       const int* g_cur, const int* g_tgt, const unsigned* g_tok_atoms, int n_atoms, int n_check,
                       cudaStream_t stream);
/* dense scores handed to host-visible processors: log_softmax (beam) or raw logits (greedy), then ForcedEOS */
int mma_score_rows(const float* logits, long long ldl, float* out, long long ldo, int R, int V, int L, int eos_id,
                   const int* cur_len, int log_softmax, cudaStream_t stream);
/* GuidedFormulaProcessor.__call__ on a dense [R, V] score matrix, in place; target row = r / beams */
int mma_guided_mask(float* scores, long long lds, int R, int V, int eos_id, int beams, const int* g_cur,
                    const int* g_tgt, const unsigned* g_tok_atoms, int n_atoms, int n_check, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MMA_B200_H */
