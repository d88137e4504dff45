// Multi-head attention forward / backward for the training and teacher-forced paths.
//
// Three mask modes as the reference builds them (custom_modeling.py:233-234, 299-318 on top of torch's
// F.multi_head_attention_forward): encoder self (key padding), decoder self (causal + key padding),
// cross (memory key padding).  The key-padding mask is a byte vector [B, Lk] (1 = real token); the causal
// mask is generated in-kernel.  Attention-probability dropout is regenerated from (seed, site, element).
//
// v1 kernels: fp32 arithmetic from shared-memory tiles with online softmax (flash-style, no [Lq,Lk]
// matrix in HBM), templated on the storage type (fp32 parity path / bf16).  Saved for backward: the
// output and the per-row log-sum-exp.
#include "common.cuh"

namespace attn {

constexpr int KT = 64;    // keys per tile
constexpr int QPW = 4;    // queries per warp
constexpr int NWARP = 4;  // warps per CTA
constexpr int QT = QPW * NWARP;

struct Args {
  const void* q; const void* k; const void* v;
  long long ldq, ldk, ldv;          // row pitch (elements); head h lives at columns [h*DH, (h+1)*DH)
  const unsigned char* kmask;       // [B, Lk] or nullptr
  void* o; long long ldo;
  float* lse;                       // [B, H, Lq]
  int B, H, Lq, Lk;
  int causal;
  float scale;
  float p_drop; unsigned long long seed; unsigned int site;
  // backward only
  const void* dout; long long lddo;
  void* dq; void* dk; void* dv;
  long long lddq, lddk, lddv;
};

template <typename T, int DH>
__global__ void __launch_bounds__(NWARP * 32) attn_fwd_kernel(Args a) {
  pdl_trigger();
  extern __shared__ float sm[];
  float* Ks = sm;                          // [KT][DH+1]
  float* Vs = Ks + KT * (DH + 1);          // [KT][DH+1]
  float* qs = Vs + KT * (DH + 1);          // [NWARP][QPW][DH]
  float* ps = qs + NWARP * QPW * DH;       // [NWARP][KT]
  constexpr int CPL = (DH + 31) / 32;      // output columns per lane
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* Q = reinterpret_cast<const T*>(a.q) + (long long)b * a.Lq * a.ldq + h * DH;
  const T* K = reinterpret_cast<const T*>(a.k) + (long long)b * a.Lk * a.ldk + h * DH;
  const T* V = reinterpret_cast<const T*>(a.v) + (long long)b * a.Lk * a.ldv + h * DH;
  const unsigned char* km = a.kmask ? a.kmask + (long long)b * a.Lk : nullptr;

  for (int idx = threadIdx.x; idx < QT * DH; idx += NWARP * 32) {
    const int qi = idx / DH, c = idx % DH;
    const int i = q0 + qi;
    qs[idx] = i < a.Lq ? to_f(Q[(long long)i * a.ldq + c]) * a.scale : 0.f;
  }
  float m[QPW], l[QPW], o[QPW][CPL];
#pragma unroll
  for (int qi = 0; qi < QPW; ++qi) {
    m[qi] = -INFINITY;
    l[qi] = 0.f;
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) o[qi][cc] = 0.f;
  }
  const bool drop = a.p_drop > 0.f;
  const uint32_t thr = drop ? drop_threshold(a.p_drop) : 0u;
  const float inv_keep = drop ? 1.f / (1.f - a.p_drop) : 1.f;
  const uint32_t dkey = drop_key(a.seed, a.site);
  // causal: keys beyond the last query of this CTA are never needed
  const int k_end = a.causal ? min(a.Lk, q0 + QT) : a.Lk;

  for (int j0 = 0; j0 < k_end; j0 += KT) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < KT * DH; idx += NWARP * 32) {
      const int j = idx / DH, c = idx % DH;
      const bool ok = j0 + j < a.Lk;
      Ks[j * (DH + 1) + c] = ok ? to_f(K[(long long)(j0 + j) * a.ldk + c]) : 0.f;
      Vs[j * (DH + 1) + c] = ok ? to_f(V[(long long)(j0 + j) * a.ldv + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int qi = 0; qi < QPW; ++qi) {
      const int i = q0 + warp * QPW + qi;
      if (i >= a.Lq) continue;
      const float* qrow = qs + (warp * QPW + qi) * DH;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
      for (int c = 0; c < DH; ++c) {
        const float qc = qrow[c];
        s0 = fmaf(qc, Ks[lane * (DH + 1) + c], s0);
        s1 = fmaf(qc, Ks[(lane + 32) * (DH + 1) + c], s1);
      }
      const int ja = j0 + lane, jb = j0 + lane + 32;
      const bool va = ja < a.Lk && (!km || km[ja]) && (!a.causal || ja <= i);
      const bool vb = jb < a.Lk && (!km || km[jb]) && (!a.causal || jb <= i);
      if (!va) s0 = -INFINITY;
      if (!vb) s1 = -INFINITY;
      const float mt = warp_max(fmaxf(s0, s1));
      const float mn = fmaxf(m[qi], mt);
      float p0 = 0.f, p1 = 0.f, corr = 1.f;
      if (mn != -INFINITY) {
        p0 = va ? __expf(s0 - mn) : 0.f;
        p1 = vb ? __expf(s1 - mn) : 0.f;
        corr = m[qi] == -INFINITY ? 0.f : __expf(m[qi] - mn);
      }
      l[qi] = l[qi] * corr + warp_sum(p0 + p1);
      m[qi] = mn;
      if (drop) {
        const unsigned long long e = (((unsigned long long)b * a.H + h) * a.Lq + i) * (unsigned long long)((a.Lk + 1) & ~1);
        p0 *= drop_scale1(dkey, e + ja, thr, inv_keep);
        p1 *= drop_scale1(dkey, e + jb, thr, inv_keep);
      }
      float* pw = ps + warp * KT;
      __syncwarp();
      pw[lane] = p0;
      pw[lane + 32] = p1;
      __syncwarp();
#pragma unroll
      for (int cc = 0; cc < CPL; ++cc) {
        const int c = lane + cc * 32;
        float acc = o[qi][cc] * corr;
        if (c < DH) {
#pragma unroll 8
          for (int j = 0; j < KT; ++j) acc = fmaf(pw[j], Vs[j * (DH + 1) + c], acc);
        }
        o[qi][cc] = acc;
      }
    }
  }
  T* O = reinterpret_cast<T*>(a.o) + (long long)b * a.Lq * a.ldo + h * DH;
#pragma unroll
  for (int qi = 0; qi < QPW; ++qi) {
    const int i = q0 + warp * QPW + qi;
    if (i >= a.Lq) continue;
    const float inv = l[qi] > 0.f ? 1.f / l[qi] : 0.f;
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) {
      const int c = lane + cc * 32;
      if (c < DH) O[(long long)i * a.ldo + c] = from_f<T>(o[qi][cc] * inv);
    }
    if (lane == 0 && a.lse) a.lse[((long long)b * a.H + h) * a.Lq + i] = l[qi] > 0.f ? m[qi] + __logf(l[qi]) : -INFINITY;
  }
}

// ---- backward, pass 1: dQ.  Same tiling as forward; per query i (one warp):
//   p_ij = exp(s_ij - lse_i);  dp_ij = dO_i . V_j;  ds_ij = p_ij * (dp_ij * keep_ij - D_i);  dQ_i = scale * sum_j ds_ij K_j
template <typename T, int DH>
__global__ void __launch_bounds__(NWARP * 32) attn_bwd_dq_kernel(Args a) {
  pdl_trigger();
  extern __shared__ float sm[];
  float* Ks = sm;
  float* Vs = Ks + KT * (DH + 1);
  float* qs = Vs + KT * (DH + 1);           // [QT][DH]  (pre-scaled q)
  float* dos = qs + QT * DH;                // [QT][DH]
  float* ps = dos + QT * DH;                // [NWARP][KT]
  constexpr int CPL = (DH + 31) / 32;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* Q = reinterpret_cast<const T*>(a.q) + (long long)b * a.Lq * a.ldq + h * DH;
  const T* K = reinterpret_cast<const T*>(a.k) + (long long)b * a.Lk * a.ldk + h * DH;
  const T* V = reinterpret_cast<const T*>(a.v) + (long long)b * a.Lk * a.ldv + h * DH;
  const T* O = reinterpret_cast<const T*>(a.o) + (long long)b * a.Lq * a.ldo + h * DH;
  const T* DO = reinterpret_cast<const T*>(a.dout) + (long long)b * a.Lq * a.lddo + h * DH;
  const unsigned char* km = a.kmask ? a.kmask + (long long)b * a.Lk : nullptr;

  for (int idx = threadIdx.x; idx < QT * DH; idx += NWARP * 32) {
    const int qi = idx / DH, c = idx % DH;
    const int i = q0 + qi;
    qs[idx] = i < a.Lq ? to_f(Q[(long long)i * a.ldq + c]) * a.scale : 0.f;
    dos[idx] = i < a.Lq ? to_f(DO[(long long)i * a.lddo + c]) : 0.f;
  }
  __syncthreads();
  float Di[QPW], lse[QPW], dq[QPW][CPL];
#pragma unroll
  for (int qi = 0; qi < QPW; ++qi) {
    const int i = q0 + warp * QPW + qi;
    float s = 0.f;
    if (i < a.Lq)
      for (int c = lane; c < DH; c += 32) s += dos[(warp * QPW + qi) * DH + c] * to_f(O[(long long)i * a.ldo + c]);
    Di[qi] = warp_sum(s);
    lse[qi] = i < a.Lq ? a.lse[((long long)b * a.H + h) * a.Lq + i] : 0.f;
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) dq[qi][cc] = 0.f;
  }
  const bool drop = a.p_drop > 0.f;
  const uint32_t thr = drop ? drop_threshold(a.p_drop) : 0u;
  const float inv_keep = drop ? 1.f / (1.f - a.p_drop) : 1.f;
  const uint32_t dkey = drop_key(a.seed, a.site);
  const int k_end = a.causal ? min(a.Lk, q0 + QT) : a.Lk;

  for (int j0 = 0; j0 < k_end; j0 += KT) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < KT * DH; idx += NWARP * 32) {
      const int j = idx / DH, c = idx % DH;
      const bool ok = j0 + j < a.Lk;
      Ks[j * (DH + 1) + c] = ok ? to_f(K[(long long)(j0 + j) * a.ldk + c]) : 0.f;
      Vs[j * (DH + 1) + c] = ok ? to_f(V[(long long)(j0 + j) * a.ldv + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int qi = 0; qi < QPW; ++qi) {
      const int i = q0 + warp * QPW + qi;
      if (i >= a.Lq) continue;
      const float* qrow = qs + (warp * QPW + qi) * DH;
      const float* drow = dos + (warp * QPW + qi) * DH;
      float s0 = 0.f, s1 = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll 8
      for (int c = 0; c < DH; ++c) {
        const float qc = qrow[c], dc = drow[c];
        s0 = fmaf(qc, Ks[lane * (DH + 1) + c], s0);
        s1 = fmaf(qc, Ks[(lane + 32) * (DH + 1) + c], s1);
        d0 = fmaf(dc, Vs[lane * (DH + 1) + c], d0);
        d1 = fmaf(dc, Vs[(lane + 32) * (DH + 1) + c], d1);
      }
      const int ja = j0 + lane, jb = j0 + lane + 32;
      const bool va = ja < a.Lk && (!km || km[ja]) && (!a.causal || ja <= i);
      const bool vb = jb < a.Lk && (!km || km[jb]) && (!a.causal || jb <= i);
      float p0 = va ? __expf(s0 - lse[qi]) : 0.f;
      float p1 = vb ? __expf(s1 - lse[qi]) : 0.f;
      if (drop) {
        const unsigned long long e = (((unsigned long long)b * a.H + h) * a.Lq + i) * (unsigned long long)((a.Lk + 1) & ~1);
        d0 *= drop_scale1(dkey, e + ja, thr, inv_keep);
        d1 *= drop_scale1(dkey, e + jb, thr, inv_keep);
      }
      float* pw = ps + warp * KT;
      __syncwarp();
      pw[lane] = p0 * (d0 - Di[qi]);
      pw[lane + 32] = p1 * (d1 - Di[qi]);
      __syncwarp();
#pragma unroll
      for (int cc = 0; cc < CPL; ++cc) {
        const int c = lane + cc * 32;
        if (c < DH) {
          float acc = dq[qi][cc];
#pragma unroll 8
          for (int j = 0; j < KT; ++j) acc = fmaf(pw[j], Ks[j * (DH + 1) + c], acc);
          dq[qi][cc] = acc;
        }
      }
    }
  }
  T* DQ = reinterpret_cast<T*>(a.dq) + (long long)b * a.Lq * a.lddq + h * DH;
#pragma unroll
  for (int qi = 0; qi < QPW; ++qi) {
    const int i = q0 + warp * QPW + qi;
    if (i >= a.Lq) continue;
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) {
      const int c = lane + cc * 32;
      if (c < DH) DQ[(long long)i * a.lddq + c] = from_f<T>(dq[qi][cc] * a.scale);
    }
  }
}

// ---- backward, pass 2: dK, dV.  CTA owns QT (=16) keys, 4 per warp, and streams query tiles of 64:
//   dV_j = sum_i (p_ij * keep_ij) dO_i;   dK_j = scale * sum_i ds_ij Q_i
template <typename T, int DH>
__global__ void __launch_bounds__(NWARP * 32) attn_bwd_dkv_kernel(Args a) {
  pdl_trigger();
  extern __shared__ float sm[];
  float* Qs = sm;                            // [KT][DH+1]   query tile (pre-scaled)
  float* DOs = Qs + KT * (DH + 1);           // [KT][DH+1]
  float* ks = DOs + KT * (DH + 1);           // [QT][DH]  this CTA's keys
  float* vs = ks + QT * DH;                  // [QT][DH]
  float* ps = vs + QT * DH;                  // [NWARP][2][KT]
  float* lses = ps + NWARP * 2 * KT;         // [KT]
  float* Dis = lses + KT;                    // [KT]
  constexpr int CPL = (DH + 31) / 32;
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * QT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* Q = reinterpret_cast<const T*>(a.q) + (long long)b * a.Lq * a.ldq + h * DH;
  const T* K = reinterpret_cast<const T*>(a.k) + (long long)b * a.Lk * a.ldk + h * DH;
  const T* V = reinterpret_cast<const T*>(a.v) + (long long)b * a.Lk * a.ldv + h * DH;
  const T* O = reinterpret_cast<const T*>(a.o) + (long long)b * a.Lq * a.ldo + h * DH;
  const T* DO = reinterpret_cast<const T*>(a.dout) + (long long)b * a.Lq * a.lddo + h * DH;
  const unsigned char* km = a.kmask ? a.kmask + (long long)b * a.Lk : nullptr;

  for (int idx = threadIdx.x; idx < QT * DH; idx += NWARP * 32) {
    const int kj = idx / DH, c = idx % DH;
    const int j = k0 + kj;
    ks[idx] = j < a.Lk ? to_f(K[(long long)j * a.ldk + c]) : 0.f;
    vs[idx] = j < a.Lk ? to_f(V[(long long)j * a.ldv + c]) : 0.f;
  }
  float dk[QPW][CPL], dv[QPW][CPL];
#pragma unroll
  for (int kj = 0; kj < QPW; ++kj)
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) { dk[kj][cc] = 0.f; dv[kj][cc] = 0.f; }
  const bool drop = a.p_drop > 0.f;
  const uint32_t thr = drop ? drop_threshold(a.p_drop) : 0u;
  const float inv_keep = drop ? 1.f / (1.f - a.p_drop) : 1.f;
  const uint32_t dkey = drop_key(a.seed, a.site);
  // causal: queries before this CTA's first key never see it
  const int i_begin = a.causal ? (k0 / KT) * KT : 0;

  for (int i0 = i_begin; i0 < a.Lq; i0 += KT) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < KT * DH; idx += NWARP * 32) {
      const int ii = idx / DH, c = idx % DH;
      const bool ok = i0 + ii < a.Lq;
      Qs[ii * (DH + 1) + c] = ok ? to_f(Q[(long long)(i0 + ii) * a.ldq + c]) * a.scale : 0.f;
      DOs[ii * (DH + 1) + c] = ok ? to_f(DO[(long long)(i0 + ii) * a.lddo + c]) : 0.f;
    }
    __syncthreads();
    // per-query D_i and lse_i for this tile (each warp does 16 of the 64 rows)
    for (int ii = warp * (KT / NWARP); ii < (warp + 1) * (KT / NWARP); ++ii) {
      const int i = i0 + ii;
      float s = 0.f;
      if (i < a.Lq)
        for (int c = lane; c < DH; c += 32) s += DOs[ii * (DH + 1) + c] * to_f(O[(long long)i * a.ldo + c]);
      s = warp_sum(s);
      if (lane == 0) {
        Dis[ii] = s;
        lses[ii] = i < a.Lq ? a.lse[((long long)b * a.H + h) * a.Lq + i] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kj = 0; kj < QPW; ++kj) {
      const int j = k0 + warp * QPW + kj;
      if (j >= a.Lk) continue;
      const bool jvalid = !km || km[j];
      const float* krow = ks + (warp * QPW + kj) * DH;
      const float* vrow = vs + (warp * QPW + kj) * DH;
      float s0 = 0.f, s1 = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll 8
      for (int c = 0; c < DH; ++c) {
        const float kc = krow[c], vc = vrow[c];
        s0 = fmaf(kc, Qs[lane * (DH + 1) + c], s0);
        s1 = fmaf(kc, Qs[(lane + 32) * (DH + 1) + c], s1);
        d0 = fmaf(vc, DOs[lane * (DH + 1) + c], d0);
        d1 = fmaf(vc, DOs[(lane + 32) * (DH + 1) + c], d1);
      }
      const int ia = i0 + lane, ib = i0 + lane + 32;
      const bool va = jvalid && ia < a.Lq && (!a.causal || j <= ia);
      const bool vb = jvalid && ib < a.Lq && (!a.causal || j <= ib);
      float p0 = va ? __expf(s0 - lses[lane]) : 0.f;
      float p1 = vb ? __expf(s1 - lses[lane + 32]) : 0.f;
      float k0s = 1.f, k1s = 1.f;
      if (drop) {
        const unsigned long long eb = ((unsigned long long)b * a.H + h) * a.Lq;
        k0s = drop_scale1(dkey, (eb + ia) * (unsigned long long)((a.Lk + 1) & ~1) + j, thr, inv_keep);
        k1s = drop_scale1(dkey, (eb + ib) * (unsigned long long)((a.Lk + 1) & ~1) + j, thr, inv_keep);
      }
      float* pw = ps + warp * 2 * KT;
      __syncwarp();
      pw[lane] = p0 * k0s;                               // dropped probabilities -> dV
      pw[lane + 32] = p1 * k1s;
      pw[KT + lane] = p0 * (d0 * k0s - Dis[lane]);       // dS -> dK
      pw[KT + lane + 32] = p1 * (d1 * k1s - Dis[lane + 32]);
      __syncwarp();
#pragma unroll
      for (int cc = 0; cc < CPL; ++cc) {
        const int c = lane + cc * 32;
        if (c < DH) {
          float av = dv[kj][cc], ak = dk[kj][cc];
#pragma unroll 8
          for (int ii = 0; ii < KT; ++ii) {
            av = fmaf(pw[ii], DOs[ii * (DH + 1) + c], av);
            ak = fmaf(pw[KT + ii], Qs[ii * (DH + 1) + c], ak);
          }
          dv[kj][cc] = av;
          dk[kj][cc] = ak;
        }
      }
    }
  }
  T* DK = reinterpret_cast<T*>(a.dk) + (long long)b * a.Lk * a.lddk + h * DH;
  T* DV = reinterpret_cast<T*>(a.dv) + (long long)b * a.Lk * a.lddv + h * DH;
#pragma unroll
  for (int kj = 0; kj < QPW; ++kj) {
    const int j = k0 + warp * QPW + kj;
    if (j >= a.Lk) continue;
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) {
      const int c = lane + cc * 32;
      // Qs was pre-scaled by `scale`, so dk already carries it
      if (c < DH) {
        DK[(long long)j * a.lddk + c] = from_f<T>(dk[kj][cc]);
        DV[(long long)j * a.lddv + c] = from_f<T>(dv[kj][cc]);
      }
    }
  }
}

template <typename T, int DH>
static int launch_fwd(const Args& a, cudaStream_t s) {
  const size_t smem = sizeof(float) * (2 * KT * (DH + 1) + NWARP * QPW * DH + NWARP * KT);
  auto k = attn_fwd_kernel<T, DH>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((a.Lq + QT - 1) / QT, a.H, a.B);
  k<<<grid, NWARP * 32, smem, s>>>(a);
  MMA_CHECK_LAUNCH();
  return MMA_OK;
}
template <typename T, int DH>
static int launch_bwd(const Args& a, cudaStream_t s) {
  {
    const size_t smem = sizeof(float) * (2 * KT * (DH + 1) + 2 * QT * DH + NWARP * KT);
    auto k = attn_bwd_dq_kernel<T, DH>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((a.Lq + QT - 1) / QT, a.H, a.B);
    k<<<grid, NWARP * 32, smem, s>>>(a);
    MMA_CHECK_LAUNCH();
  }
  {
    const size_t smem = sizeof(float) * (2 * KT * (DH + 1) + 2 * QT * DH + NWARP * 2 * KT + 2 * KT);
    auto k = attn_bwd_dkv_kernel<T, DH>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((a.Lk + QT - 1) / QT, a.H, a.B);
    k<<<grid, NWARP * 32, smem, s>>>(a);
    MMA_CHECK_LAUNCH();
  }
  return MMA_OK;
}

template <typename T>
static int dispatch(const Args& a, int dh, bool bwd, cudaStream_t s) {
  switch (dh) {
    case 16: return bwd ? launch_bwd<T, 16>(a, s) : launch_fwd<T, 16>(a, s);
    case 32: return bwd ? launch_bwd<T, 32>(a, s) : launch_fwd<T, 32>(a, s);
    case 64: return bwd ? launch_bwd<T, 64>(a, s) : launch_fwd<T, 64>(a, s);
    case 128: return bwd ? launch_bwd<T, 128>(a, s) : launch_fwd<T, 128>(a, s);
    default: return MMA_ERR_UNSUPPORTED;
  }
}

}  // namespace attn

// q/k/v/o: [B, L, ld] views with head h at columns [h*dh, (h+1)*dh); type: MMA_BF16 / MMA_F32 for all of them.
extern "C" int mma_attn_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                            const unsigned char* kmask, void* o, long long ldo, float* lse, int B, int H, int Lq,
                            int Lk, int dh, int causal, float scale, float p_drop, unsigned long long seed,
                            unsigned int site, int type, cudaStream_t stream) {
  if (B <= 0 || Lq <= 0 || Lk <= 0) return MMA_OK;
  attn::Args a{};
  a.q = q; a.k = k; a.v = v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.kmask = kmask; a.o = o; a.ldo = ldo;
  a.lse = lse; a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk; a.causal = causal; a.scale = scale;
  a.p_drop = p_drop; a.seed = seed; a.site = site;
  return type == MMA_F32 ? attn::dispatch<float>(a, dh, false, stream) : attn::dispatch<bf16>(a, dh, false, stream);
}

extern "C" int mma_attn_bwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                            const unsigned char* kmask, const void* o, long long ldo, const float* lse,
                            const void* dout, long long lddo, void* dq, long long lddq, void* dk, long long lddk,
                            void* dv, long long lddv, int B, int H, int Lq, int Lk, int dh, int causal, float scale,
                            float p_drop, unsigned long long seed, unsigned int site, int type, cudaStream_t stream) {
  if (B <= 0 || Lq <= 0 || Lk <= 0) return MMA_OK;
  attn::Args a{};
  a.q = q; a.k = k; a.v = v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.kmask = kmask;
  a.o = const_cast<void*>(o); a.ldo = ldo; a.lse = const_cast<float*>(lse);
  a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk; a.causal = causal; a.scale = scale;
  a.p_drop = p_drop; a.seed = seed; a.site = site;
  a.dout = dout; a.lddo = lddo; a.dq = dq; a.dk = dk; a.dv = dv; a.lddq = lddq; a.lddk = lddk; a.lddv = lddv;
  return type == MMA_F32 ? attn::dispatch<float>(a, dh, true, stream) : attn::dispatch<bf16>(a, dh, true, stream);
}
