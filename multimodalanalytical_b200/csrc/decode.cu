// KV-cached decoding kernels (prediction path).
//
// The reference decodes with use_cache=False (wrapper.py:450): every step re-runs the decoder over the whole
// prefix.  Here each step processes one token per live row:
//   * decode_embed      target-token embedding + LayerNorm + positional row `cur_len-1`
//   * decode_attn       single-query attention.  self mode: appends this step's K/V to the cache and attends
//                       over the row's history through an ancestor table (beam reordering never copies the
//                       cache); cross mode: attends over the encoder memory K/V computed once per spectrum and
//                       shared by the K beams of that spectrum (no xK expansion).
//   * beam_step         log-softmax -> forced EOS -> + running score -> top-2K of K*V -> live beams / finished
//                       pool / early-stop heuristic, restating transformers' `_beam_search` (see oracle), one CTA
//                       per spectrum; emits next tokens + parent rows and rewrites the ancestor table.
//   * greedy_step       argmax decoding (n_beams == 1), transformers' `_sample` semantics.
// All kernels read the current length from device memory so one CUDA graph replays for every step.
#include <cstdlib>

#include "common.cuh"

namespace dec {

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) decode_embed_kernel(const int* __restrict__ tok, const float* __restrict__ table,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps,
                                                           const float* __restrict__ pos, const int* __restrict__ cur_len,
                                                           float* __restrict__ out, int rows, int d) {
  pdl_trigger();
  pdl_wait();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= rows) return;
  const int t = *cur_len - 1;
  const float* src = table + (long long)tok[r] * d;
  float mean = 0.f, rstd = 1.f;
  if (gamma) {
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += src[c];
    mean = warp_sum(s) / d;
    float q = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float e = src[c] - mean;
      q += e * e;
    }
    rstd = rsqrtf(warp_sum(q) / d + eps);
  }
  const float* p = pos + (long long)t * d;
  float* o = out + (long long)r * d;
  for (int c = lane; c < d; c += 32) {
    float v = src[c];
    if (gamma) v = (v - mean) * rstd * gamma[c] + beta[c];
    o[c] = v + p[c];
  }
}

// ---------------------------------------------------------------------------------------------
struct AttnArgs {
  const void* q; long long ldq;      // [R, ldq], head h at h*DH
  const void* knew; const void* vnew; long long ldkv;   // self: this step's K/V rows [R, ldkv]
  void* kc; void* vc;                // self: cache [R][Lmax][d];  cross: memory K / V [B][S][ldm] (head offset applied)
  long long ldm;                     // cross: row pitch of memory K/V
  const int* anc;                    // self: [2][R][Lmax] ancestor rows, buffer cur_len&1 is live (nullptr: identity)
  const unsigned char* kmask;        // cross: [B][S]
  const int* cur_len;
  void* o; long long ldo;
  int R, H, d, Lmax, S, beams, cross;
  float scale;
};

__device__ __forceinline__ void ld8(const float* p, float (&x)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float (&x)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const float2 f = __bfloat1622float2(h[w]);
    x[2 * w] = f.x;
    x[2 * w + 1] = f.y;
  }
}
__device__ __forceinline__ void st8(float* p, const float (&x)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(x[4], x[5], x[6], x[7]);
}
__device__ __forceinline__ void st8(bf16* p, const float (&x)[8]) {
  uint4 u;
  __nv_bfloat162 a = __floats2bfloat162_rn(x[0], x[1]), b = __floats2bfloat162_rn(x[2], x[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(x[4], x[5]), d = __floats2bfloat162_rn(x[6], x[7]);
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  *reinterpret_cast<uint4*>(p) = u;
}

// One warp per (row, head).  G = DH/8 lanes share a key (each owns one 8-element chunk -> 16-byte loads, a key row
// is one coalesced 128-byte line for DH = 64); KPP = 32/G keys are scored per pass; online softmax across passes.
template <typename T, int DH>
__global__ void __launch_bounds__(128) decode_attn_kernel(AttnArgs a) {
  pdl_trigger();
  constexpr int G = DH / 8;
  constexpr int KPP = 32 / G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 4 + warp;
  if (gw >= a.R * a.H) return;
  const int r = gw / a.H, h = gw % a.H;
  const int t = *a.cur_len - 1;
  const int sub = lane % G, kslot = lane / G;
  float qreg[8];
  ld8(reinterpret_cast<const T*>(a.q) + (long long)r * a.ldq + h * DH + sub * 8, qreg);
#pragma unroll
  for (int c = 0; c < 8; ++c) qreg[c] *= a.scale;
  int nkeys;
  const T *Kb, *Vb;
  long long pitch;
  const int* anc = nullptr;
  const unsigned char* km = nullptr;
  if (!a.cross) {
    // append this step's K/V for (r, h) at position t, then attend over t+1 positions
    if (lane < G) {
      float tmp[8];
      const long long dst = ((long long)r * a.Lmax + t) * a.d + h * DH + lane * 8;
      ld8(reinterpret_cast<const T*>(a.knew) + (long long)r * a.ldkv + h * DH + lane * 8, tmp);
      st8(reinterpret_cast<T*>(a.kc) + dst, tmp);
      ld8(reinterpret_cast<const T*>(a.vnew) + (long long)r * a.ldkv + h * DH + lane * 8, tmp);
      st8(reinterpret_cast<T*>(a.vc) + dst, tmp);
    }
    nkeys = t + 1;
    Kb = reinterpret_cast<const T*>(a.kc) + h * DH + sub * 8;
    Vb = reinterpret_cast<const T*>(a.vc) + h * DH + sub * 8;
    pitch = a.d;  // cache [R][Lmax][d]: a row's positions are contiguous (see decode_self_attn2_kernel)
    // the beam step ping-pongs the ancestor table on cur_len parity: [2][R][Lmax]
    anc = a.anc ? a.anc + ((long long)((t + 1) & 1) * a.R + r) * a.Lmax : nullptr;
  } else {
    const int b = r / a.beams;
    nkeys = a.S;
    Kb = reinterpret_cast<const T*>(a.kc) + (long long)b * a.S * a.ldm + h * DH + sub * 8;
    Vb = reinterpret_cast<const T*>(a.vc) + (long long)b * a.S * a.ldm + h * DH + sub * 8;
    pitch = a.ldm;
    km = a.kmask ? a.kmask + (long long)b * a.S : nullptr;
  }
  __syncwarp();
  float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) o[c] = 0.f;
  for (int j0 = 0; j0 < nkeys; j0 += KPP) {
    const int j = j0 + kslot;
    const bool valid = j < nkeys && (!km || km[j]);
    long long off = 0;
    if (j < nkeys) {
      if (!a.cross) {
        const int src = (j == t || !anc) ? r : anc[j];
        off = ((long long)src * a.Lmax + j) * pitch;
      } else {
        off = (long long)j * pitch;
      }
    }
    float s = 0.f;
    if (valid) {
      float kx[8];
      ld8(Kb + off, kx);
#pragma unroll
      for (int c = 0; c < 8; ++c) s = fmaf(qreg[c], kx[c], s);
    }
#pragma unroll
    for (int w = 1; w < G; w <<= 1) s += __shfl_xor_sync(0xffffffffu, s, w);
    if (!valid) s = -INFINITY;
    const float mt = warp_max(s);
    const float mn = fmaxf(m, mt);
    float p = 0.f, corr = 1.f;
    if (mn != -INFINITY) {
      p = valid ? __expf(s - mn) : 0.f;
      corr = m == -INFINITY ? 0.f : __expf(m - mn);
    }
    l = l * corr + warp_sum(p) * (1.0f / G);
    m = mn;
    float vx[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) vx[c] = 0.f;
    if (p != 0.f) ld8(Vb + off, vx);
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] = fmaf(p, vx[c], o[c] * corr);
  }
  // combine the KPP key slots (lanes with equal `sub`)
#pragma unroll
  for (int w = G; w < 32; w <<= 1) {
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] += __shfl_xor_sync(0xffffffffu, o[c], w);
  }
  if (kslot == 0) {
    const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] *= inv;
    st8(reinterpret_cast<T*>(a.o) + (long long)r * a.ldo + h * DH + sub * 8, o);
  }
}

// ---------------------------------------------------------------------------------------------
// beam step
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

struct BeamArgs {
  const float* logits; long long ldl;   // [B*K, ldl] fp32, last-position logits
  const float* extra_bias;              // optional [B*K, V] additive processor output (log-prob domain)
  int B, K, V, L;
  int pad_id, eos_id;
  const int* cur_len;
  int* run_seq;      // [2][B][K][L]
  int* fin_seq;      // [2][B][K][L]
  float* run_score;  // [B][K]
  float* fin_score;  // [B][K]
  unsigned char* fin_flag;   // [B][K]
  int* fin_len;      // [B][K]
  unsigned char* improvable; // [B]
  unsigned char* all_hit;    // [B]
  int* anc;          // [2][B*K][L]
  int* next_tok;     // [B*K]
  int* parent_row;   // [B*K]
  // processed-score mode: `logits` already holds log_softmax -> forced EOS -> processors (skip stages 1 + forced EOS)
  int prenorm;
  // formula-guided decoding (GuidedFormulaProcessor, generation/logit_processors.py:89-152), nullptr = off
  const int* g_cur;        // [B*K][g_na] atom counts of each running hypothesis (host chemistry, refreshed per step)
  const int* g_tgt;        // [B][g_na] atom counts of the target formula
  const unsigned* g_tok;   // [V] bit a set: the token adds one atom of element a
  int g_na, g_ncheck;      // elements compared for <eos> (14) / for the look-ahead (9: H and rarer elements skipped)
};

// Row verdict of the formula guide: bit 31 = formula matches (force <eos> to 0), bit 30 = some element still short
// (ban <eos>), bit 29 = an element already over target (ban everything), bits 0..g_ncheck-1 = elements at their
// target count (ban tokens that add one).  Order of the writes follows logit_processors.py:122-150.
__device__ __forceinline__ unsigned guide_row(const int* cur, const int* tgt, int na, int ncheck) {
  bool match = true, small = false, over = false;
  unsigned full = 0u;
  for (int e = 0; e < na; ++e) {
    const int c = cur[e], t = tgt[e];
    match = match && (c == t);
    small = small || (c < t);
    if (e < ncheck) {
      over = over || (c > t);
      if (c >= t) full |= 1u << e;
    }
  }
  return full | (match ? 0x80000000u : 0u) | (small ? 0x40000000u : 0u) | (over ? 0x20000000u : 0u);
}
__device__ __forceinline__ float guide_apply(float lp, unsigned verdict, unsigned tok_bits, bool is_eos) {
  if (is_eos) {
    if (verdict & 0x80000000u) lp = 0.f;
    if (verdict & 0x40000000u) lp = -INFINITY;
  }
  if ((verdict & 0x20000000u) || (tok_bits & verdict & 0x1fffffffu)) lp = -INFINITY;
  return lp;
}

constexpr int MAXK = 64;  // beams (2K candidates <= 128)

__global__ void __launch_bounds__(256) beam_step_kernel(BeamArgs a) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ uint32_t cand[];  // [K*V] ordered keys of accumulated log-probs
  __shared__ float s_lse[MAXK];
  __shared__ unsigned long long s_wkeys[8 * 2 * MAXK];  // per-warp top-2K keys
  __shared__ float s_live[2 * MAXK];
  __shared__ float s_ms[3 * MAXK];
  __shared__ unsigned char s_hit[2 * MAXK];
  __shared__ float c_score[2 * MAXK];
  __shared__ int c_idx[2 * MAXK];
  __shared__ int n_run_src[MAXK];     // candidate slot feeding running beam k
  __shared__ int n_fin_src[MAXK];     // merged slot feeding finished slot k (<K: old finished, >=K: candidate)
  __shared__ float n_fin_score[MAXK];
  __shared__ unsigned char n_fin_flag[MAXK];
  __shared__ int n_fin_len[MAXK];
  __shared__ float n_run_score[MAXK];
  __shared__ unsigned s_guide[MAXK];

  const int b = blockIdx.x;
  const int K = a.K, V = a.V, L = a.L;
  const int cur = *a.cur_len;
  const int rd = cur & 1, wr = rd ^ 1;
  const float NEG = -1.0e9f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const bool force_eos = cur == L - 1;

  if (a.g_cur && threadIdx.x < K)
    s_guide[threadIdx.x] = guide_row(a.g_cur + (long long)(b * K + threadIdx.x) * a.g_na, a.g_tgt + (long long)b * a.g_na,
                                     a.g_na, a.g_ncheck);
  // 1. log-softmax statistics per beam row (one warp per row)
  for (int k = warp; k < K && !a.prenorm; k += nwarp) {
    const float* x = a.logits + (long long)(b * K + k) * a.ldl;
    float mx = -INFINITY;
    for (int c = lane; c < V; c += 32) mx = fmaxf(mx, x[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < V; c += 32) s += expf(x[c] - mx);
    s = warp_sum(s);
    if (lane == 0) s_lse[k] = mx + logf(s);
  }
  __syncthreads();
  // 2. accumulated scores of the K*V continuations
  for (int i = threadIdx.x; i < K * V; i += blockDim.x) {
    const int k = i / V, c = i % V;
    float lp = a.logits[(long long)(b * K + k) * a.ldl + c];
    if (!a.prenorm) {
      lp -= s_lse[k];
      if (force_eos) lp = c == a.eos_id ? 0.f : -INFINITY;
    }
    if (a.extra_bias) lp += a.extra_bias[(long long)(b * K + k) * V + c];
    if (a.g_cur) lp = guide_apply(lp, s_guide[k], a.g_tok[c], c == a.eos_id);
    cand[i] = f2ord(lp + a.run_score[b * K + k]);
  }
  __syncthreads();
  // 3. top-2K, descending, ties -> lowest flat index (keys (score, ~index) are unique).  Two levels without block-wide
  //    syncs per pick: every warp extracts the top-2K of its own slice of the K*V candidates, then warp 0 extracts the
  //    top-2K of the nwarp*2K survivors.  A lane keeps the best key of ITS strided elements in a register; a pick is a
  //    shuffle arg-max over the 32 lane heads, and only the winning lane re-scans its own elements for its next head
  //    (measured at 1 spectrum x 10 beams, V = 120: 12.6 + 8.1 us for the two levels when every pick re-scanned the whole
  //    slice from shared memory and wrote the "taken" marker through lane 0; 7.8 + 7.6 us with lane heads and a 64-bit
  //    shuffle ladder).
  const int keep = 2 * K;
  const int KV = K * V;
  // warp-wide maximum of a 64-bit key with two REDUX instructions (high words, then low words among the lanes that hold
  // the maximal high word) instead of a 5-step shuffle ladder of 64-bit values
  auto wmax64 = [](unsigned long long v) {
    const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
    const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return ((unsigned long long)mh << 32) | (unsigned long long)ml;
  };
  {
    const int per_warp = (KV + nwarp - 1) / nwarp;
    const int lo = warp * per_warp, hi = min(KV, lo + per_warp);
    auto head = [&]() {  // best remaining key among this lane's elements lo + lane, lo + lane + 32, ...
      unsigned long long best = 0ull;
      for (int i = lo + lane; i < hi; i += 32) {
        const uint32_t o = cand[i];
        if (o) {
          const unsigned long long key = ((unsigned long long)o << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
          best = key > best ? key : best;
        }
      }
      return best;
    };
    unsigned long long mine = head();
    for (int sel = 0; sel < keep; ++sel) {
      const unsigned long long best = wmax64(mine);
      if (lane == 0) s_wkeys[warp * keep + sel] = best;  // dense [nwarp][keep]: the merge below indexes it flat
      if (best && mine == best) {  // the owner retires the element (only this lane ever touches it) and finds its next head
        cand[(int)(0xffffffffu - (uint32_t)(best & 0xffffffffull))] = 0u;
        mine = head();
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
    const int n = nwarp * keep;
    auto head = [&](int& where) {
      unsigned long long best = 0ull;
      where = -1;
      for (int i = lane; i < n; i += 32) {
        const unsigned long long key = s_wkeys[i];
        if (key > best) { best = key; where = i; }
      }
      return best;
    };
    int where;
    unsigned long long mine = head(where);
    for (int sel = 0; sel < keep; ++sel) {
      const unsigned long long best = wmax64(mine);
      if (lane == 0) {
        if (best) {
          c_idx[sel] = (int)(0xffffffffu - (uint32_t)(best & 0xffffffffull));
          c_score[sel] = ord2f((uint32_t)(best >> 32));
        } else {  // fewer than 2K candidates (K*V < 2K): pad with -inf on slot 0
          c_idx[sel] = 0;
          c_score[sel] = -INFINITY;
        }
      }
      if (best && mine == best) {
        s_wkeys[where] = 0ull;
        mine = head(where);
      }
    }
  }
  __syncthreads();
  // 4. live beams / finished pool by parallel ranking (rank = number of strictly better entries, ties -> lower slot)
  const bool improvable = a.improvable[b] != 0;
  const float glen = (float)(cur + 1 - 1);  // generated length incl. this token (prompt length 1)
  if (threadIdx.x < keep) {
    const int c = threadIdx.x;
    const int tok = c_idx[c] % V;
    const bool hit = (tok == a.eos_id) || (cur + 1 >= L);
    s_hit[c] = hit ? 1 : 0;
    s_live[c] = c_score[c] + (hit ? 1.0f : 0.0f) * NEG;
    const bool just = hit && c < K;
    float fs = c_score[c] / glen;
    fs = fs + (improvable ? 0.0f : 1.0f) * NEG;
    fs = fs + (just ? 0.0f : 1.0f) * NEG;
    s_ms[K + c] = fs;
  }
  if (threadIdx.x < K) s_ms[threadIdx.x] = a.fin_score[b * K + threadIdx.x];
  __syncthreads();
  if (threadIdx.x < keep) {
    const int c = threadIdx.x;
    const float mine = s_live[c];
    int rank = 0;
    for (int u = 0; u < keep; ++u) {
      const float o = s_live[u];
      rank += (o > mine || (o == mine && u < c)) ? 1 : 0;
    }
    if (rank < K) {
      n_run_src[rank] = c;
      n_run_score[rank] = mine;
    }
  }
  if (threadIdx.x < K + keep) {
    const int t = threadIdx.x;
    const float mine = s_ms[t];
    int rank = 0;
    for (int u = 0; u < K + keep; ++u) {
      const float o = s_ms[u];
      rank += (o > mine || (o == mine && u < t)) ? 1 : 0;
    }
    if (rank < K) {
      n_fin_src[rank] = t;
      n_fin_score[rank] = mine;
      if (t < K) {
        n_fin_flag[rank] = a.fin_flag[b * K + t];
        n_fin_len[rank] = a.fin_len[b * K + t];
      } else {
        const int c = t - K;
        n_fin_flag[rank] = (s_hit[c] && c < K) ? 1 : 0;
        n_fin_len[rank] = cur + 1 - 1;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    bool allh = true;
    for (int c = 0; c < keep; ++c) allh = allh && s_hit[c];
    float worst = INFINITY;
    for (int k = 0; k < K; ++k) worst = fminf(worst, n_fin_score[k]);
    // early-stop heuristic with cur_len already advanced: best live sum / (cur_len+1 - prompt)
    const float best_possible = n_run_score[0] / (float)(cur + 1 - 1);
    bool can = false;
    for (int k = 0; k < K; ++k) {
      const float w = n_fin_flag[k] ? worst : NEG;
      can = can || (best_possible > w);
    }
    a.improvable[b] = (improvable && can) ? 1 : 0;
    a.all_hit[b] = allh ? 1 : 0;
  }
  __syncthreads();
  // 5. parallel state rewrite (ping-pong buffers)
  const long long seq_stride = (long long)a.B * K * L;
  // (ping-pong halves never alias: __restrict__ lets the loads of several iterations be in flight together)
  const int* __restrict__ rs_old = a.run_seq + rd * seq_stride + (long long)b * K * L;
  int* __restrict__ rs_new = a.run_seq + wr * seq_stride + (long long)b * K * L;
  const int* __restrict__ fs_old = a.fin_seq + rd * seq_stride + (long long)b * K * L;
  int* __restrict__ fs_new = a.fin_seq + wr * seq_stride + (long long)b * K * L;
  const long long anc_stride = (long long)a.B * K * L;
  const int* __restrict__ an_old = a.anc + rd * anc_stride + (long long)b * K * L;
  int* __restrict__ an_new = a.anc + wr * anc_stride + (long long)b * K * L;
  // only positions 0 .. cur carry information: beyond them both halves still hold the fill value they were initialised
  // with (sequences) or stale entries that are rewritten before they are read (ancestors)
  const int Lc = min(L, cur + 1);
#pragma unroll 4
  for (int j = threadIdx.x; j < K * Lc; j += blockDim.x) {
    const int k = j / Lc, p = j - k * Lc;
    const int i = k * L + p;
    {  // running beams
      const int c = n_run_src[k];
      const int beam = c_idx[c] / V, tok = c_idx[c] % V;
      rs_new[i] = p == cur ? tok : rs_old[beam * L + p];
      // ancestors: positions < cur-1 inherit, position cur-1 is the parent's own row
      an_new[i] = p == cur - 1 ? b * K + beam : an_old[beam * L + p];
    }
    {  // finished pool
      const int s = n_fin_src[k];
      int v;
      if (s < K) v = fs_old[s * L + p];
      else {
        const int c = s - K;
        const int beam = c_idx[c] / V, tok = c_idx[c] % V;
        v = p == cur ? tok : rs_old[beam * L + p];
      }
      fs_new[i] = v;
    }
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const int c = n_run_src[k];
    a.run_score[b * K + k] = n_run_score[k];
    a.next_tok[b * K + k] = c_idx[c] % V;
    a.parent_row[b * K + k] = b * K + c_idx[c] / V;
    a.fin_score[b * K + k] = n_fin_score[k];
    a.fin_flag[b * K + k] = n_fin_flag[k];
    a.fin_len[b * K + k] = n_fin_len[k];
  }
}

// greedy: one warp per row.  seq [R][L]; unfinished [R]
__global__ void __launch_bounds__(256) greedy_step_kernel(const float* __restrict__ logits, long long ldl,
                                                          const float* __restrict__ extra_bias, int R, int V, int L,
                                                          int pad_id, int eos_id, const int* __restrict__ cur_len,
                                                          int* __restrict__ seq, unsigned char* __restrict__ unfinished,
                                                          int* __restrict__ next_tok, int prenorm,
                                                          const int* __restrict__ g_cur, const int* __restrict__ g_tgt,
                                                          const unsigned* __restrict__ g_tok, int g_na, int g_ncheck) {
  pdl_trigger();
  pdl_wait();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= R) return;
  const int cur = *cur_len;
  const float* x = logits + (long long)r * ldl;
  unsigned long long best = 0ull;
  // processor order of transformers' merged list: ForcedEOS first, then the caller's processors
  const unsigned verdict = g_cur ? guide_row(g_cur + (long long)r * g_na, g_tgt + (long long)r * g_na, g_na, g_ncheck) : 0u;
  for (int c = lane; c < V; c += 32) {
    float v = x[c];
    if (!prenorm && cur == L - 1) v = c == eos_id ? 0.f : -INFINITY;
    if (extra_bias) v += extra_bias[(long long)r * V + c];
    if (g_cur) v = guide_apply(v, verdict, g_tok[c], c == eos_id);
    const unsigned long long key = ((unsigned long long)f2ord(v) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)c);
    best = key > best ? key : best;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, off);
    best = other > best ? other : best;
  }
  if (lane == 0) {
    int tok = (int)(0xffffffffu - (uint32_t)(best & 0xffffffffull));
    const bool unf = unfinished[r] != 0;
    if (!unf) tok = pad_id;
    seq[(long long)r * L + cur] = tok;
    next_tok[r] = tok;
    unfinished[r] = (unf && tok != eos_id) ? 1 : 0;
  }
}

// Dense processed scores for host-visible logits processors: beam search sees log_softmax(logits) then ForcedEOS,
// greedy sees the raw logits then ForcedEOS (transformers `_beam_search` / `_sample`).  One warp per row.
__global__ void __launch_bounds__(256) score_rows_kernel(const float* __restrict__ logits, long long ldl,
                                                         float* __restrict__ out, long long ldo, int R, int V, int L,
                                                         int eos_id, const int* __restrict__ cur_len, int log_softmax) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= R) return;
  const int cur = *cur_len;
  const float* x = logits + (long long)r * ldl;
  float* o = out + (long long)r * ldo;
  if (cur == L - 1) {
    for (int c = lane; c < V; c += 32) o[c] = c == eos_id ? 0.f : -INFINITY;
    return;
  }
  float lse = 0.f;
  if (log_softmax) {
    float mx = -INFINITY;
    for (int c = lane; c < V; c += 32) mx = fmaxf(mx, x[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < V; c += 32) s += expf(x[c] - mx);
    lse = mx + logf(warp_sum(s));
  }
  for (int c = lane; c < V; c += 32) o[c] = x[c] - lse;
}

// GuidedFormulaProcessor.__call__ on a dense score matrix (in place); tgt row = r / beams.
__global__ void __launch_bounds__(256) guided_mask_kernel(float* __restrict__ scores, long long lds, int R, int V,
                                                          int eos_id, int beams, const int* __restrict__ g_cur,
                                                          const int* __restrict__ g_tgt,
                                                          const unsigned* __restrict__ g_tok, int g_na, int g_ncheck) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= R) return;
  const unsigned verdict = guide_row(g_cur + (long long)r * g_na, g_tgt + (long long)(r / beams) * g_na, g_na, g_ncheck);
  float* x = scores + (long long)r * lds;
  for (int c = lane; c < V; c += 32) x[c] = guide_apply(x[c], verdict, g_tok[c], c == eos_id);
}

__global__ void advance_kernel(int* cur_len) {
  pdl_trigger();
  pdl_wait();
  *cur_len += 1;
}

// Self-attention of one decode step, bf16 / head dim 64: one CTA per (row, group of 4 heads).
// A cached key row is gathered ONCE for the four heads (a 512-byte coalesced read, lane = (head, 16-byte chunk)): the
// ancestor look-up and the address arithmetic - most of the instructions of the one-warp-per-(row, head) kernel above
// (ncu: 2 500 warp instructions per (row, head) at 64 positions of which 256 were FMAs) - are shared by the heads.
// The CTA's 4 warps walk interleaved chunks of 8 positions (split-K over the history) with a per-chunk online softmax:
// the 8 key gathers of a chunk are independent loads in flight together, then its 8 value gathers; the partial
// (max, sum, output) states of the warps are merged through shared memory.
constexpr int SA_KC = 8;  // positions per chunk
__global__ void __launch_bounds__(128) decode_self_attn2_kernel(AttnArgs a) {
  pdl_trigger();
  pdl_wait();
  constexpr int DH = 64;
  __shared__ float s_m[4][32], s_l[4][32], s_o[4][32][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x;
  const int col = blockIdx.y * (4 * DH) + lane * 8;  // this lane's 8 columns: head = col / 64
  const int t = *a.cur_len - 1;
  const bf16* kc = reinterpret_cast<const bf16*>(a.kc) + col;
  const bf16* vc = reinterpret_cast<const bf16*>(a.vc) + col;
  const bf16* knew = reinterpret_cast<const bf16*>(a.knew) + (long long)r * a.ldkv + col;
  const bf16* vnew = reinterpret_cast<const bf16*>(a.vnew) + (long long)r * a.ldkv + col;
  const long long rowpitch = (long long)a.Lmax * a.d;  // cache layout [R][Lmax][d]
  if (warp == 0) {  // append this step's K / V of (r, these heads) at position t; position t is read from knew / vnew
    const long long dst = (long long)r * rowpitch + (long long)t * a.d;
    *reinterpret_cast<uint4*>(const_cast<bf16*>(kc) + dst) = *reinterpret_cast<const uint4*>(knew);
    *reinterpret_cast<uint4*>(const_cast<bf16*>(vc) + dst) = *reinterpret_cast<const uint4*>(vnew);
  }
  float qreg[8];
  ld8(reinterpret_cast<const bf16*>(a.q) + (long long)r * a.ldq + col, qreg);
#pragma unroll
  for (int c = 0; c < 8; ++c) qreg[c] *= a.scale;
  const int* anc = a.anc ? a.anc + ((long long)((t + 1) & 1) * a.R + r) * a.Lmax : nullptr;

  float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) o[c] = 0.f;
  if (warp == 3) {
    // position t (this step's own key / value, still in registers of the projection output): the initial state of the
    // warp that gets the fewest chunks
    float kx[8];
    ld8(knew, kx);
    float sv = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) sv = fmaf(qreg[c], kx[c], sv);
    sv += __shfl_xor_sync(0xffffffffu, sv, 1);
    sv += __shfl_xor_sync(0xffffffffu, sv, 2);
    sv += __shfl_xor_sync(0xffffffffu, sv, 4);
    m = sv;
    l = 1.f;
    ld8(vnew, o);
  }
  // cached positions j < t; offsets are 32-bit element indices into one layer's cache (checked on the host)
  const unsigned rp = (unsigned)a.Lmax * (unsigned)a.d, dd = (unsigned)a.d;
  for (int j0 = warp * SA_KC; j0 < t; j0 += 4 * SA_KC) {
    int src[SA_KC];
    if (anc) {
      const int4 a0 = *reinterpret_cast<const int4*>(anc + j0), a1 = *reinterpret_cast<const int4*>(anc + j0 + 4);
      src[0] = a0.x; src[1] = a0.y; src[2] = a0.z; src[3] = a0.w;
      src[4] = a1.x; src[5] = a1.y; src[6] = a1.z; src[7] = a1.w;
    } else {
#pragma unroll
      for (int i = 0; i < SA_KC; ++i) src[i] = r;
    }
    unsigned off[SA_KC];
    uint4 raw[SA_KC];
#pragma unroll
    for (int i = 0; i < SA_KC; ++i) {
      // positions past the history read the last cached one (always a valid address) and get weight 0 below
      const int j = min(j0 + i, t - 1);
      const int sr = j0 + i < t ? src[i] : r;
      off[i] = (unsigned)sr * rp + (unsigned)j * dd;
      raw[i] = *reinterpret_cast<const uint4*>(kc + off[i]);
    }
    float sc[SA_KC];
    float cm = -INFINITY;
#pragma unroll
    for (int i = 0; i < SA_KC; ++i) {
      const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&raw[i]);
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float2 f = __bfloat1622float2(hp[w]);
        s0 = fmaf(qreg[2 * w], f.x, s0);
        s1 = fmaf(qreg[2 * w + 1], f.y, s1);
      }
      float sv = s0 + s1;
      sv += __shfl_xor_sync(0xffffffffu, sv, 1);
      sv += __shfl_xor_sync(0xffffffffu, sv, 2);
      sv += __shfl_xor_sync(0xffffffffu, sv, 4);
      sc[i] = j0 + i < t ? sv : -INFINITY;
      cm = fmaxf(cm, sc[i]);
    }
    // value gathers of the chunk go out before the softmax arithmetic needs them
#pragma unroll
    for (int i = 0; i < SA_KC; ++i) raw[i] = *reinterpret_cast<const uint4*>(vc + off[i]);
    const float mn = fmaxf(m, cm);  // finite: the chunk holds at least one valid position
    const float corr = __expf(m - mn);
    l *= corr;
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] *= corr;
#pragma unroll
    for (int i = 0; i < SA_KC; ++i) {
      const float pw = __expf(sc[i] - mn);  // exp(-inf) = 0 for positions past the history
      l += pw;
      const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&raw[i]);
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float2 f = __bfloat1622float2(hp[w]);
        o[2 * w] = fmaf(pw, f.x, o[2 * w]);
        o[2 * w + 1] = fmaf(pw, f.y, o[2 * w + 1]);
      }
    }
    m = mn;
  }
  s_m[warp][lane] = m;
  s_l[warp][lane] = l;
#pragma unroll
  for (int c = 0; c < 8; ++c) s_o[warp][lane][c] = o[c];
  __syncthreads();
  if (warp == 0) {
    float M = fmaxf(fmaxf(s_m[0][lane], s_m[1][lane]), fmaxf(s_m[2][lane], s_m[3][lane]));
    float L = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float f = __expf(s_m[w][lane] - M);  // a warp without chunks has m = -inf: weight 0
      L = fmaf(s_l[w][lane], f, L);
#pragma unroll
      for (int c = 0; c < 8; ++c) o[c] = fmaf(s_o[w][lane][c], f, o[c]);
    }
    const float inv = L > 0.f ? 1.f / L : 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] *= inv;
    st8(reinterpret_cast<bf16*>(a.o) + (long long)r * a.ldo + col, o);
  }
}

// Beam-grouped variant for large batches: one CTA per (spectrum, group of 4 heads), one WARP per beam row walking the whole
// history (no split over warps, no merge).  The K beams of a spectrum descend from few ancestors, so their gathers hit the
// same cache rows: grouped in one CTA they are served by that SM's L1 instead of K separate trips to L2 (ncu on the per-row
// kernel at 2560 rows x 61 positions: 300 MB of L2 -> SM traffic per launch for 53 MB of DRAM reads).
__global__ void __launch_bounds__(1024) decode_self_attn3_kernel(AttnArgs a) {
  pdl_trigger();
  pdl_wait();
  constexpr int DH = 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * a.beams + warp;  // blockDim.x = 32 * beams
  const int col = blockIdx.y * (4 * DH) + lane * 8;
  const int t = *a.cur_len - 1;
  const bf16* kc = reinterpret_cast<const bf16*>(a.kc) + col;
  const bf16* vc = reinterpret_cast<const bf16*>(a.vc) + col;
  const bf16* knew = reinterpret_cast<const bf16*>(a.knew) + (long long)r * a.ldkv + col;
  const bf16* vnew = reinterpret_cast<const bf16*>(a.vnew) + (long long)r * a.ldkv + col;
  const unsigned rp = (unsigned)a.Lmax * (unsigned)a.d, dd = (unsigned)a.d;
  {
    const long long dst = (long long)r * rp + (long long)t * a.d;
    *reinterpret_cast<uint4*>(const_cast<bf16*>(kc) + dst) = *reinterpret_cast<const uint4*>(knew);
    *reinterpret_cast<uint4*>(const_cast<bf16*>(vc) + dst) = *reinterpret_cast<const uint4*>(vnew);
  }
  float qreg[8];
  ld8(reinterpret_cast<const bf16*>(a.q) + (long long)r * a.ldq + col, qreg);
#pragma unroll
  for (int c = 0; c < 8; ++c) qreg[c] *= a.scale;
  const int* anc = a.anc ? a.anc + ((long long)((t + 1) & 1) * a.R + r) * a.Lmax : nullptr;
  float m, l, o[8];
  {  // position t: this step's own key / value
    float kx[8];
    ld8(knew, kx);
    float sv = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) sv = fmaf(qreg[c], kx[c], sv);
    sv += __shfl_xor_sync(0xffffffffu, sv, 1);
    sv += __shfl_xor_sync(0xffffffffu, sv, 2);
    sv += __shfl_xor_sync(0xffffffffu, sv, 4);
    m = sv;
    l = 1.f;
    ld8(vnew, o);
  }
  for (int j0 = 0; j0 < t; j0 += SA_KC) {
    int src[SA_KC];
    if (anc) {
      const int4 a0 = *reinterpret_cast<const int4*>(anc + j0), a1 = *reinterpret_cast<const int4*>(anc + j0 + 4);
      src[0] = a0.x; src[1] = a0.y; src[2] = a0.z; src[3] = a0.w;
      src[4] = a1.x; src[5] = a1.y; src[6] = a1.z; src[7] = a1.w;
    } else {
#pragma unroll
      for (int i = 0; i < SA_KC; ++i) src[i] = r;
    }
    unsigned off[SA_KC];
    uint4 raw[SA_KC];
#pragma unroll
    for (int i = 0; i < SA_KC; ++i) {
      const int j = min(j0 + i, t - 1);
      const int sr = j0 + i < t ? src[i] : r;
      off[i] = (unsigned)sr * rp + (unsigned)j * dd;
      raw[i] = *reinterpret_cast<const uint4*>(kc + off[i]);
    }
    float sc[SA_KC];
    float cm = -INFINITY;
#pragma unroll
    for (int i = 0; i < SA_KC; ++i) {
      const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&raw[i]);
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float2 f = __bfloat1622float2(hp[w]);
        s0 = fmaf(qreg[2 * w], f.x, s0);
        s1 = fmaf(qreg[2 * w + 1], f.y, s1);
      }
      float sv = s0 + s1;
      sv += __shfl_xor_sync(0xffffffffu, sv, 1);
      sv += __shfl_xor_sync(0xffffffffu, sv, 2);
      sv += __shfl_xor_sync(0xffffffffu, sv, 4);
      sc[i] = j0 + i < t ? sv : -INFINITY;
      cm = fmaxf(cm, sc[i]);
    }
#pragma unroll
    for (int i = 0; i < SA_KC; ++i) raw[i] = *reinterpret_cast<const uint4*>(vc + off[i]);
    const float mn = fmaxf(m, cm);
    const float corr = __expf(m - mn);
    l *= corr;
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] *= corr;
#pragma unroll
    for (int i = 0; i < SA_KC; ++i) {
      const float pw = __expf(sc[i] - mn);
      l += pw;
      const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&raw[i]);
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float2 f = __bfloat1622float2(hp[w]);
        o[2 * w] = fmaf(pw, f.x, o[2 * w]);
        o[2 * w + 1] = fmaf(pw, f.y, o[2 * w + 1]);
      }
    }
    m = mn;
  }
  const float inv = 1.f / l;
#pragma unroll
  for (int c = 0; c < 8; ++c) o[c] *= inv;
  st8(reinterpret_cast<bf16*>(a.o) + (long long)r * a.ldo + col, o);
}

template <typename T>
static int launch_attn(const AttnArgs& a, int dh, cudaStream_t s) {
  const int blocks = (a.R * a.H + 3) / 4;
  switch (dh) {
    case 16: decode_attn_kernel<T, 16><<<blocks, 128, 0, s>>>(a); break;
    case 32: decode_attn_kernel<T, 32><<<blocks, 128, 0, s>>>(a); break;
    case 64: decode_attn_kernel<T, 64><<<blocks, 128, 0, s>>>(a); break;
    case 128: decode_attn_kernel<T, 128><<<blocks, 128, 0, s>>>(a); break;
    default: return MMA_ERR_UNSUPPORTED;
  }
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

}  // namespace dec

using namespace dec;

extern "C" int mma_decode_embed(const int* tok, const float* table, const float* gamma, const float* beta, float eps,
                                const float* pos, const int* cur_len, float* out, int rows, int d,
                                cudaStream_t stream) {
  if (rows <= 0) return MMA_ERR_ARG;
  launch_rowop(decode_embed_kernel, dim3((rows + 7) / 8), dim3(256), 0, stream, tok, table, gamma, beta, eps, pos, cur_len, out,
               rows, d);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

extern "C" int mma_decode_self_attn(const void* q, long long ldq, const void* knew, const void* vnew, long long ldkv,
                                    void* kcache, void* vcache, const int* anc, const int* cur_len, void* o,
                                    long long ldo, int R, int H, int dh, int Lmax, float scale, int type, int beams,
                                    cudaStream_t stream) {
  AttnArgs a{};
  a.q = q; a.ldq = ldq; a.knew = knew; a.vnew = vnew; a.ldkv = ldkv; a.kc = kcache; a.vc = vcache; a.anc = anc;
  a.cur_len = cur_len; a.o = o; a.ldo = ldo; a.R = R; a.H = H; a.d = H * dh; a.Lmax = Lmax; a.scale = scale;
  a.cross = 0; a.beams = 1;
  const bool anc_ok = !anc || (reinterpret_cast<uintptr_t>(anc) & 15) == 0;
  static int v2 = -1;
  if (v2 < 0) {
    const char* e = getenv("MMA_DECODE_ATTN2");
    v2 = e ? atoi(e) : 1;
  }
  if (v2 && type == MMA_BF16 && dh == 64 && (H % 4) == 0 && (Lmax % 8) == 0 && R > 0 && anc_ok &&
      (unsigned long long)R * Lmax * H * dh < 4294967296ull && (ldq % 8) == 0 && (ldkv % 8) == 0 && (ldo % 8) == 0 &&
      (reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(knew) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(vnew) & 15) == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
    // beams > 1: `beams` rows of a spectrum are consecutive; group them in one CTA when the batch is large enough to
    // fill the machine that way (MMA_DECODE_ATTN2=3 forces, =1 forbids the grouped kernel).  Measured on B200 (ms / step,
    // grouped vs per-row): 1024 spectra x 10 beams 2.41 vs 2.49, 1024 x 30 5.88 vs 6.10, but 256 x 10 1.09 vs 0.96 and
    // 256 x 30 1.87 vs 1.82 - with few CTAs the serial walk of a warp over the whole history is exposed
    if (beams >= 4 && beams <= 32 && R % beams == 0 && (v2 == 3 || (v2 != 1 && R / beams * (H / 4) >= 2048))) {
      a.beams = beams;
      launch_rowop(decode_self_attn3_kernel, dim3(R / beams, H / 4), dim3(32 * beams), 0, stream, a);
    } else {
      launch_rowop(decode_self_attn2_kernel, dim3(R, H / 4), dim3(128), 0, stream, a);
    }
    MMA_CHECK_LAUNCH();
    return MMA_OK;
  }
  return type == MMA_F32 ? launch_attn<float>(a, dh, stream) : launch_attn<bf16>(a, dh, stream);
}

extern "C" int mma_decode_cross_attn(const void* q, long long ldq, const void* kmem, const void* vmem, long long ldm,
                                     const unsigned char* kmask, const int* cur_len, void* o, long long ldo, int R,
                                     int H, int dh, int S, int beams, float scale, int type, cudaStream_t stream) {
  AttnArgs a{};
  a.q = q; a.ldq = ldq; a.kc = const_cast<void*>(kmem); a.vc = const_cast<void*>(vmem); a.ldm = ldm; a.kmask = kmask;
  a.cur_len = cur_len; a.o = o; a.ldo = ldo; a.R = R; a.H = H; a.d = H * dh; a.S = S; a.beams = beams;
  a.scale = scale; a.cross = 1;
  return type == MMA_F32 ? launch_attn<float>(a, dh, stream) : launch_attn<bf16>(a, dh, stream);
}

extern "C" int mma_beam_step_ex(const float* logits, long long ldl, const float* extra_bias, int B, int K, int V, int L,
                                int pad_id, int eos_id, const int* cur_len, int* run_seq, int* fin_seq,
                                float* run_score, float* fin_score, unsigned char* fin_flag, int* fin_len,
                                unsigned char* improvable, unsigned char* all_hit, int* anc, int* next_tok,
                                int* parent_row, int prenorm, const int* g_cur, const int* g_tgt,
                                const unsigned* g_tok_atoms, int n_atoms, int n_check, cudaStream_t stream) {
  if (K > MAXK || K < 1 || B < 1) return MMA_ERR_ARG;
  if (g_cur && (!g_tgt || !g_tok_atoms || n_atoms < 1 || n_atoms > 29 || n_check < 0 || n_check > n_atoms))
    return MMA_ERR_ARG;
  BeamArgs a{logits, ldl, extra_bias, B, K, V, L, pad_id, eos_id, cur_len, run_seq, fin_seq, run_score, fin_score,
             fin_flag, fin_len, improvable, all_hit, anc, next_tok, parent_row, prenorm, g_cur, g_tgt, g_tok_atoms,
             n_atoms, n_check};
  const size_t smem = sizeof(uint32_t) * (size_t)K * V;
  if (smem > 200 * 1024) return MMA_ERR_UNSUPPORTED;
  if (smem > 40 * 1024) cudaFuncSetAttribute(beam_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  launch_rowop(beam_step_kernel, dim3(B), dim3(256), smem, stream, a);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

extern "C" int mma_beam_step(const float* logits, long long ldl, const float* extra_bias, int B, int K, int V, int L,
                             int pad_id, int eos_id, const int* cur_len, int* run_seq, int* fin_seq, float* run_score,
                             float* fin_score, unsigned char* fin_flag, int* fin_len, unsigned char* improvable,
                             unsigned char* all_hit, int* anc, int* next_tok, int* parent_row, cudaStream_t stream) {
  return mma_beam_step_ex(logits, ldl, extra_bias, B, K, V, L, pad_id, eos_id, cur_len, run_seq, fin_seq, run_score,
                          fin_score, fin_flag, fin_len, improvable, all_hit, anc, next_tok, parent_row, 0, nullptr,
                          nullptr, nullptr, 0, 0, stream);
}

extern "C" int mma_greedy_step_ex(const float* logits, long long ldl, const float* extra_bias, int R, int V, int L,
                                  int pad_id, int eos_id, const int* cur_len, int* seq, unsigned char* unfinished,
                                  int* next_tok, int prenorm, const int* g_cur, const int* g_tgt,
                                  const unsigned* g_tok_atoms, int n_atoms, int n_check, cudaStream_t stream) {
  if (R < 1) return MMA_ERR_ARG;
  if (g_cur && (!g_tgt || !g_tok_atoms || n_atoms < 1 || n_atoms > 29 || n_check < 0 || n_check > n_atoms))
    return MMA_ERR_ARG;
  launch_rowop(greedy_step_kernel, dim3((R + 7) / 8), dim3(256), 0, stream, logits, ldl, extra_bias, R, V, L, pad_id, eos_id,
               cur_len, seq, unfinished, next_tok, prenorm, g_cur, g_tgt, g_tok_atoms, n_atoms, n_check);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

extern "C" int mma_greedy_step(const float* logits, long long ldl, const float* extra_bias, int R, int V, int L,
                               int pad_id, int eos_id, const int* cur_len, int* seq, unsigned char* unfinished,
                               int* next_tok, cudaStream_t stream) {
  return mma_greedy_step_ex(logits, ldl, extra_bias, R, V, L, pad_id, eos_id, cur_len, seq, unfinished, next_tok, 0,
                            nullptr, nullptr, nullptr, 0, 0, stream);
}

extern "C" int mma_score_rows(const float* logits, long long ldl, float* out, long long ldo, int R, int V, int L,
                              int eos_id, const int* cur_len, int log_softmax, cudaStream_t stream) {
  if (R < 1 || V < 1) return MMA_ERR_ARG;
  score_rows_kernel<<<(R + 7) / 8, 256, 0, stream>>>(logits, ldl, out, ldo, R, V, L, eos_id, cur_len, log_softmax);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

extern "C" int mma_guided_mask(float* scores, long long lds, int R, int V, int eos_id, int beams, const int* g_cur,
                               const int* g_tgt, const unsigned* g_tok_atoms, int n_atoms, int n_check,
                               cudaStream_t stream) {
  if (R < 1 || V < 1 || beams < 1 || !g_cur || !g_tgt || !g_tok_atoms || n_atoms < 1 || n_atoms > 29 || n_check < 0 ||
      n_check > n_atoms)
    return MMA_ERR_ARG;
  guided_mask_kernel<<<(R + 7) / 8, 256, 0, stream>>>(scores, lds, R, V, eos_id, beams, g_cur, g_tgt, g_tok_atoms,
                                                      n_atoms, n_check);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}

extern "C" int mma_advance(int* cur_len, cudaStream_t stream) {
  launch_rowop(advance_kernel, dim3(1), dim3(1), 0, stream, cur_len);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}
