"""Tensor-level wrappers over the C-ABI kernels.  torch is used for device memory and streams only:
every function launches exactly the CUDA kernels of `lib/libmma_b200.so` on torch's current stream and
raises if the library is unavailable.  No CPU or ATen fallback exists on purpose.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import (EPI_ACCUM, EPI_DGELU, EPI_DGLU, EPI_DRELU, EPI_GELU, EPI_GLU_MUL, EPI_RELU, EPI_RESID, EPI_STORE,
                   MMA_BF16, MMA_F32, Epi, check)

LAUNCHES = 0  # number of kernel-launching C-ABI calls made (bench.py reports it as `gpu_launches`)
# bf16 / head-dim-64 attention: "tcgen05" = single-tile TMEM kernels when Lq, Lk <= 128 (else the streaming mma.sync
# kernels); "mma" = always the mma.sync kernels (kept selectable for A/B measurements and tests)
ATTN_IMPL = os.environ.get("MMA_ATTN_IMPL", "tcgen05")


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _ty(t):
    if t.dtype == torch.float32:
        return MMA_F32
    if t.dtype == torch.bfloat16:
        return MMA_BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("multimodalanalytical_b200 kernels need CUDA tensors (no CPU fallback)")


def make_epi(kind, out, out2=None, bias=None, resid=None, aux=None, aux2=None, p_drop=0.0, seed=0, site=0,
             alpha=1.0, accumulate=0, drop_ld=0):
    e = Epi()
    e.kind = kind
    e.out_f32 = _ty(out)
    e.out, e.ldo = out.data_ptr(), out.stride(0)
    if out2 is not None:
        assert out2.dtype == out.dtype
        e.out2, e.ldo2 = out2.data_ptr(), out2.stride(0)
    if bias is not None:
        assert bias.dtype == torch.float32
        e.bias = bias.data_ptr()
    if resid is not None:
        e.resid, e.ldr, e.resid_f32 = resid.data_ptr(), resid.stride(0), _ty(resid)
    if aux is not None:
        e.aux, e.lda, e.aux_f32 = aux.data_ptr(), aux.stride(0), _ty(aux)
    if aux2 is not None:
        assert aux2.dtype == aux.dtype
        e.aux2, e.lda2 = aux2.data_ptr(), aux2.stride(0)
    e.p_drop, e.alpha, e.seed, e.site = float(p_drop), float(alpha), int(seed), int(site)
    e.accumulate = int(accumulate)
    e.drop_ld = int(drop_ld)
    return e


def _tc_ok(t, mn_major, rows, k):
    """TMA needs a 16-byte aligned base and a 16-byte multiple row pitch."""
    return t.dtype == torch.bfloat16 and t.data_ptr() % 16 == 0 and (t.stride(0) * 2) % 16 == 0 and t.stride(1) == 1


def gemm(A, B, M, N, K, epi, a_mn=False, b_mn=False, splits=1, force_simt=False, max_ctas=0):
    """C[M,N] = epi(A_op[M,K] @ B_op[N,K]^T).  `a_mn`/`b_mn`: operand memory is [K, rows] instead of
    [rows, K].  bf16 operands run on the tcgen05 kernel, fp32 (or TMA-incompatible layouts) on the SIMT one."""
    _need_cuda(A, B)
    lib = _lib.load()
    if epi.p_drop > 0 and epi.drop_ld == 0:
        epi.drop_ld = N
    if not force_simt and _tc_ok(A, a_mn, M, K) and _tc_ok(B, b_mn, N, K):
        rc = lib.mma_gemm_bf16(A.data_ptr(), A.stride(0), int(a_mn), B.data_ptr(), B.stride(0), int(b_mn), M, N, K,
                               C.byref(epi), splits, max_ctas, _stream())
        check(rc, "mma_gemm_bf16")
    else:
        ssplit = 1
        if epi.kind == EPI_ACCUM and epi.accumulate in (1, 2):
            # few output tiles + long reduction (weight gradients): split K over CTAs with atomic accumulation
            tiles = ((M + 63) // 64) * ((N + 63) // 64)
            if tiles < 148 and K >= 1024:
                ssplit = max(1, min(64, K // 256, 296 // tiles))
            epi.accumulate = 2 if ssplit > 1 else 1
        sam, sak = (A.stride(1), A.stride(0)) if a_mn else (A.stride(0), A.stride(1))
        sbn, sbk = (B.stride(1), B.stride(0)) if b_mn else (B.stride(0), B.stride(1))
        rc = lib.mma_gemm_simt(A.data_ptr(), _ty(A), sam, sak, B.data_ptr(), _ty(B), sbn, sbk, M, N, K, C.byref(epi),
                               ssplit, _stream())
        check(rc, "mma_gemm_simt")
    _count()


def gemm_dual(A1, B1, A2, B2, M, N, K1, K2, epi, b_mn=False):
    """C[M,N] = epi(A1 B1_op^T + A2 B2_op^T) in one CTA-pair launch (one accumulator over both reductions).
    Returns False (nothing launched) outside the pair kernel's envelope."""
    _need_cuda(A1, B1, A2, B2)
    if not all(_tc_ok(t, False, 0, 0) for t in (A1, B1, A2, B2)):
        return False
    rc = _lib.load().mma_gemm2_dual(A1.data_ptr(), A1.stride(0), B1.data_ptr(), B1.stride(0), A2.data_ptr(),
                                    A2.stride(0), B2.data_ptr(), B2.stride(0), int(b_mn), M, N, K1, K2, C.byref(epi),
                                    _stream())
    if rc == -3:
        return False
    check(rc, "mma_gemm2_dual")
    _count()
    return True


def ffn_glu_fwd(h, W1, Wg, b1, bg, M, N, K, a, z1=None, z2=None, p_drop=0.0, seed=0, site=0):
    """a = drop(gelu(h W1^T + b1) * (h Wg^T + bg)) (+ saved bf16 pre-activations z1, z2) in ONE tcgen05 launch.
    Returns False (nothing launched) outside the fused kernel's envelope."""
    _need_cuda(h, W1, Wg, a)
    ts = (h, W1, Wg, a) + ((z1, z2) if z1 is not None else ())
    if not all(_tc_ok(t, False, 0, 0) for t in ts):
        return False
    rc = _lib.load().mma_ffn_glu_fwd(h.data_ptr(), h.stride(0), W1.data_ptr(), W1.stride(0), Wg.data_ptr(),
                                     Wg.stride(0), b1.data_ptr(), bg.data_ptr(), M, N, K, a.data_ptr(), a.stride(0),
                                     _p(z1), z1.stride(0) if z1 is not None else 0, _p(z2),
                                     z2.stride(0) if z2 is not None else 0, float(p_drop), int(seed), int(site),
                                     _stream())
    if rc == -3:
        return False
    check(rc, "mma_ffn_glu_fwd")
    _count()
    return True


def ffn_dglu(dy, W2, M, N, K, z1, z2, dz1, dz2, p_drop=0.0, seed=0, site=0, drop_ld=0):
    """dz1, dz2 of the gate from dy [M, K] and linear2.weight W2 [K, N] in one CTA-pair launch; False outside the
    kernel's envelope."""
    _need_cuda(dy, W2, z1, z2, dz1, dz2)
    if not all(_tc_ok(t, False, 0, 0) for t in (dy, W2, z1, z2, dz1, dz2)):
        return False
    rc = _lib.load().mma_ffn_dglu(dy.data_ptr(), dy.stride(0), W2.data_ptr(), W2.stride(0), M, N, K, z1.data_ptr(),
                                  z1.stride(0), z2.data_ptr(), z2.stride(0), dz1.data_ptr(), dz1.stride(0),
                                  dz2.data_ptr(), dz2.stride(0), float(p_drop), int(seed), int(site), int(drop_ld or N),
                                  _stream())
    if rc == -3:
        return False
    check(rc, "mma_ffn_dglu")
    _count()
    return True


# Engine-level switch for the fused residual-product + LayerNorm kernel. Measured on B200 at the C2 shapes (CUDA-graph
# timing, scripts/ln_fuse_bench.py): 26.9 vs 33.1 us for the decoder out-projection, but 43.2 vs 28.5 us for the
# encoder FFN-2 - a CTA pair owns 256 rows x all 512 columns, so M = 9216 fills only 36 of the 74 pairs - and the
# training step as a whole is 2.5 % slower with it.  Off by default; it pays from M ~ 32k rows upwards.
# 0 off (default), 1 always, 2 only where the 256-row tiles fill the CTA pairs (K <= d), 3 the same for any K.  Round 2, C2
# step on one box: 6.80 / 6.85 / 6.81 ms for modes 0 / 2 / 3 - the selective modes do not pay either.
FUSE_LN = int(os.environ.get("MMA_FUSE_LN", "0"))


def fuse_ln_wanted(M, K):
    """Should the residual product [M, 512] x K be fused with the LayerNorm that follows it?  The fused kernel gives a CTA
    pair 256 rows x all 512 columns: it pays when ceil(M / 256) tiles fill the 74 pairs of a wave to >= 85 %."""
    if FUSE_LN in (0, 1):
        return bool(FUSE_LN)
    tiles = (M + 255) // 256
    fill = tiles / (-(-tiles // 74) * 74)
    return fill >= 0.85 and (FUSE_LN == 3 or K <= 512)


def gemm_resid_ln(A, W, M, N, K, epi, gamma, beta, h, eps=1e-5):
    """epi.out = epi.resid + drop(A W^T + bias) and h = LayerNorm(epi.out) * gamma + beta in ONE launch (N == 512: up to
    4736 rows without dropout 128 x 128 tiles in 4-CTA clusters that exchange the row statistics through distributed
    shared memory, else the CTA-pair kernel).  Returns False (nothing launched) when the shape / layout is outside the
    fused kernels' envelope - the caller then runs `gemm` and `ln_fwd`."""
    _need_cuda(A, W, h)
    if not (N == 512 and M >= 65 and epi.kind == EPI_RESID and epi.out_f32 and epi.resid_f32
            and h.dtype == torch.bfloat16 and _tc_ok(A, False, M, K) and _tc_ok(W, False, N, K)):
        return False
    if epi.p_drop > 0 and M < 512:
        return False
    if epi.p_drop > 0 and epi.drop_ld == 0:
        epi.drop_ld = N
    rc = _lib.load().mma_gemm2_resid_ln(A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0), M, N, K, C.byref(epi),
                                        gamma.data_ptr(), beta.data_ptr(), float(eps), h.data_ptr(), h.stride(0),
                                        _stream())
    if rc == -3:
        return False
    check(rc, "mma_gemm2_resid_ln")
    _count()
    return True


def wgrad_group_ok(dy, x, n_out, k_in):
    return (_tc_ok(dy, True, 0, 0) and _tc_ok(x, True, 0, 0) and n_out % 8 == 0 and k_in % 8 == 0)


def wgrad_group(items):
    """items: list of (dy [R, Nout] bf16, x [R, Kin] bf16, dW [Nout, Kin] fp32, db [Nout] fp32 | None, Nout, Kin, R).
    One persistent tcgen05 launch per <= 8 items: dW += dy^T x, db += colsum(dy)."""
    lib = _lib.load()
    for i0 in range(0, len(items), 8):
        chunk = items[i0:i0 + 8]
        n = len(chunk)
        VP, LL, IN = C.c_void_p * n, C.c_longlong * n, C.c_int * n
        dy = VP(*[it[0].data_ptr() for it in chunk])
        lddy = LL(*[it[0].stride(0) for it in chunk])
        x = VP(*[it[1].data_ptr() for it in chunk])
        ldx = LL(*[it[1].stride(0) for it in chunk])
        out = VP(*[it[2].data_ptr() for it in chunk])
        ldo = LL(*[it[2].stride(0) for it in chunk])
        db = VP(*[(it[3].data_ptr() if it[3] is not None else 0) for it in chunk])
        nout = IN(*[it[4] for it in chunk])
        kin = IN(*[it[5] for it in chunk])
        rr = IN(*[it[6] for it in chunk])
        check(lib.mma_wgrad_group(n, dy, lddy, x, ldx, out, ldo, db, nout, kin, rr, _stream()), "mma_wgrad_group")
        _count()


def gather_rows(ids, table, out, scale=None):
    _need_cuda(ids, table, out)
    assert ids.dtype == torch.int64 and ids.is_contiguous() and table.dtype == torch.float32 and out.dtype == torch.float32
    check(_lib.load().mma_gather_rows(ids.data_ptr(), _p(scale), table.data_ptr(), out.data_ptr(), ids.numel(),
                                      table.shape[1], _stream()), "mma_gather_rows")
    _count()


def scatter_add_rows(ids, g, dtable, pad_idx, scale=None):
    _need_cuda(ids, g, dtable)
    check(_lib.load().mma_scatter_add_rows(ids.data_ptr(), _p(scale), g.data_ptr(), dtable.data_ptr(), ids.numel(),
                                           dtable.shape[1], int(pad_idx), _stream()), "mma_scatter_add_rows")
    _count()


def ln_fwd(x, gamma, beta, y, y2=None, add=None, group=0, out_group_stride=0, out_offset=0, eps=1e-5, rows=None,
           d=None):
    _need_cuda(x, y)
    rows = x.shape[0] if rows is None else rows
    d = x.shape[1] if d is None else d
    check(_lib.load().mma_ln_fwd(
        x.data_ptr(), _ty(x), x.stride(0), _p(gamma), _p(beta), eps, y.data_ptr(), _ty(y), y.stride(0),
        _p(y2), _ty(y2) if y2 is not None else 0, y2.stride(0) if y2 is not None else 0,
        _p(add), add.stride(0) if add is not None else 0, rows, d, group, out_group_stride, out_offset, _stream()),
        "mma_ln_fwd")
    _count()


def ln_bwd(dy, x, gamma, dx=None, dres=None, dxb=None, dgamma=None, dbeta=None, p_drop=0.0, seed=0, site=0, group=0,
           in_group_stride=0, in_offset=0, eps=1e-5, rows=None, d=None):
    _need_cuda(dy, x)
    rows = x.shape[0] if rows is None else rows
    d = x.shape[1] if d is None else d
    check(_lib.load().mma_ln_bwd(
        dy.data_ptr(), _ty(dy), dy.stride(0), group, in_group_stride, in_offset, x.data_ptr(), _ty(x), x.stride(0),
        _p(gamma), eps, _p(dres), dres.stride(0) if dres is not None else 0, _p(dx),
        dx.stride(0) if dx is not None else 0, _p(dxb), _ty(dxb) if dxb is not None else 0,
        dxb.stride(0) if dxb is not None else 0, float(p_drop), int(seed), int(site), _p(dgamma), _p(dbeta), rows, d,
        _stream()), "mma_ln_bwd")
    _count()


def colsum(x, out, rows=None, cols=None):
    _need_cuda(x, out)
    rows = x.shape[0] if rows is None else rows
    cols = x.shape[1] if cols is None else cols
    check(_lib.load().mma_colsum(x.data_ptr(), _ty(x), x.stride(0), out.data_ptr(), rows, cols, _stream()),
          "mma_colsum")
    _count()


def patchify(raw, out, mean, std, offset=0, hop=None, pad=None, missing=None, masking=False, rows=None):
    """raw fp32 [B, n_points] -> out fp32 [B, P, ps] standardised patches (PatchPreprocessor.__call__ on the device).
    rows (int32 [B]): `raw` / `missing` are a dataset table and batch element b reads row rows[b]."""
    _need_cuda(raw, out)
    assert raw.dtype == torch.float32 and raw.stride(1) == 1 and out.dtype == torch.float32 and out.is_contiguous()
    B, P, ps = out.shape
    hop = ps if hop is None else hop
    assert offset + (P - 1) * hop + ps <= raw.shape[1], "patches run past the spectrum"
    if rows is None:
        check(_lib.load().mma_patchify(raw.data_ptr(), raw.stride(0), int(offset), float(mean), float(std),
                                       out.data_ptr(), _p(pad), _p(missing), int(masking), B, P, ps, hop, _stream()),
              "mma_patchify")
    else:
        _need_cuda(rows)
        assert rows.dtype == torch.int32 and rows.numel() == B
        check(_lib.load().mma_patchify_rows(raw.data_ptr(), raw.stride(0), rows.data_ptr(), int(offset), float(mean),
                                            float(std), out.data_ptr(), _p(pad), _p(missing), int(masking), B, P, ps,
                                            hop, _stream()), "mma_patchify_rows")
    _count()


def patchify_deriv(raw, out, mean, std, n_patches, offset=0, n_use=None, hop=None, pad=None, missing=None, masking=False,
                   rows=None):
    """PatchPreprocessor(derivative=True): out fp32 [B, P + Pd, ps] = the P standardised patches, then the Pd = `n_patches`
    patches of torch.gradient over the `n_use` raw points from `offset` on (patches.py:91-95)."""
    _need_cuda(raw, out)
    assert raw.dtype == torch.float32 and raw.stride(1) == 1 and out.dtype == torch.float32 and out.is_contiguous()
    B, Ptot, ps = out.shape
    Pd = int(n_patches)
    P = Ptot - Pd
    hop = ps if hop is None else hop
    n_use = raw.shape[1] - offset if n_use is None else int(n_use)
    assert P > 0 and offset + n_use <= raw.shape[1] and (P - 1) * hop + ps <= n_use and Pd * ps <= n_use
    if rows is not None:
        _need_cuda(rows)
        assert rows.dtype == torch.int32 and rows.numel() == B
    check(_lib.load().mma_patchify_deriv(raw.data_ptr(), raw.stride(0), _p(rows), int(offset), n_use, float(mean),
                                         float(std), out.data_ptr(), _p(pad), _p(missing), int(masking), B, P, Pd, ps,
                                         hop, _stream()), "mma_patchify_deriv")
    _count(2)


def _ragged_ok(flat, offsets, rows):
    _need_cuda(flat, offsets, rows)
    assert offsets.dtype == torch.int64 and rows.dtype == torch.int32 and flat.is_contiguous()


def collate_tokens(flat, offsets, row_valid, rows, pad_id, max_len, ids, mask):
    """Ragged int32 token rows -> ids int64 [B, L] + validity mask u8 [B, L] for the samples `rows` (int32 [B])."""
    _ragged_ok(flat, offsets, rows)
    assert flat.dtype == torch.int32 and ids.dtype == torch.int64 and ids.is_contiguous()
    assert mask is None or (mask.dtype == torch.uint8 and mask.shape == ids.shape and mask.is_contiguous())
    B, L = ids.shape
    check(_lib.load().mma_collate_tokens(flat.data_ptr(), offsets.data_ptr(), _p(row_valid), rows.data_ptr(), B, L,
                                         int(pad_id), int(max_len), ids.data_ptr(), _p(mask), _stream()),
          "mma_collate_tokens")
    _count()


def collate_target(flat, offsets, rows, pad_id, max_len, dec_in, dec_mask, labels):
    """Ragged target rows -> teacher-forcing tensors [B, T]: dec_in = tokens[:-1], labels = tokens[1:] (pad -> -100)."""
    _ragged_ok(flat, offsets, rows)
    assert flat.dtype == torch.int32 and dec_in.dtype == torch.int64 and labels.dtype == torch.int64
    assert dec_mask.dtype == torch.uint8 and dec_in.shape == labels.shape == dec_mask.shape
    B, T = dec_in.shape
    check(_lib.load().mma_collate_target(flat.data_ptr(), offsets.data_ptr(), rows.data_ptr(), B, T, int(pad_id),
                                         int(max_len), dec_in.data_ptr(), dec_mask.data_ptr(), labels.data_ptr(),
                                         _stream()), "mma_collate_target")
    _count()


def collate_values(flat, offsets, rows, pad_value, max_len, out, mask):
    """Ragged fp32 rows [n, width] -> out [B, L, width] (+ validity mask u8 [B, L])."""
    _ragged_ok(flat, offsets, rows)
    assert flat.dtype == torch.float32 and flat.dim() == 2 and out.dtype == torch.float32 and out.is_contiguous()
    B, L, width = out.shape
    assert width == flat.shape[1]
    check(_lib.load().mma_collate_values(flat.data_ptr(), offsets.data_ptr(), rows.data_ptr(), B, L, width,
                                         float(pad_value), int(max_len), out.data_ptr(), _p(mask), _stream()),
          "mma_collate_values")
    _count()


ALIGN_LOSS_KINDS = {"mae": 0, "mse": 1, "sid": 2}


def masked_mean_fwd(mem, mask, pooled, B, S):
    _need_cuda(mem, mask, pooled)
    assert mask.dtype == torch.uint8 and mask.is_contiguous() and pooled.dtype == torch.float32 and pooled.is_contiguous()
    check(_lib.load().mma_masked_mean_fwd(mem.data_ptr(), _ty(mem), mem.stride(0), mask.data_ptr(), pooled.data_ptr(),
                                          B, S, mem.shape[1], _stream()), "mma_masked_mean_fwd")
    _count()


def masked_mean_bwd(dpooled, mask, dmem, B, S):
    _need_cuda(dpooled, mask, dmem)
    assert dmem.dtype == torch.float32 and dpooled.dtype == torch.float32 and dpooled.is_contiguous()
    check(_lib.load().mma_masked_mean_bwd(dpooled.data_ptr(), mask.data_ptr(), dmem.data_ptr(), dmem.stride(0), B, S,
                                          dmem.shape[1], _stream()), "mma_masked_mean_bwd")
    _count()


def align_loss(z, target, kind, lam, lm_loss, out, dz=None, dscale=1.0):
    """out[0] = loss(sigmoid(z), target), out[1] = lm_loss + lam * out[0]; dz = dscale * lam * dloss/dz."""
    _need_cuda(z, target, out)
    assert z.dtype == torch.float32 and target.dtype == torch.float32 and out.dtype == torch.float32
    check(_lib.load().mma_align_loss(z.data_ptr(), z.stride(0), target.data_ptr(), target.stride(0), z.shape[0],
                                     z.shape[1], ALIGN_LOSS_KINDS[kind], float(lam), _p(lm_loss), out.data_ptr(),
                                     _p(dz), 0 if dz is None else dz.stride(0), float(dscale), _stream()),
          "mma_align_loss")
    _count()


def add_strided(dst, src):
    """dst (a 1-D strided fp32 view) += src (contiguous, same element count)."""
    _need_cuda(dst, src)
    assert dst.dim() == 1 and dst.dtype == torch.float32 and src.dtype == torch.float32 and src.is_contiguous()
    assert dst.numel() == src.numel()
    check(_lib.load().mma_add_strided(dst.data_ptr(), dst.stride(0), src.data_ptr(), src.numel(), _stream()),
          "mma_add_strided")
    _count()


def cast_f32_bf16(src, dst):
    _need_cuda(src, dst)
    assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.numel() == dst.numel()
    check(_lib.load().mma_cast_f32_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "mma_cast_f32_bf16")
    _count()


def cast_bf16_f32(src, dst):
    _need_cuda(src, dst)
    check(_lib.load().mma_cast_bf16_f32(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "mma_cast_bf16_f32")
    _count()


_ATT_WS = {}


def _att_ws(device, name, shape, dtype=torch.bfloat16):
    """workspace of the blocked attention kernels (bf16 partials, fp32 log-sum-exp), one per (device, role, shape)."""
    key = (str(device), name, tuple(shape))
    t = _ATT_WS.get(key)
    if t is None:
        t = _ATT_WS[key] = torch.empty(shape, dtype=dtype, device=device)
    return t


def _attn_tc_ok(dh, *ts):
    return dh == 64 and all(t.dtype == torch.bfloat16 and t.stride(0) % 8 == 0 and t.data_ptr() % 16 == 0 for t in ts)


def attn_fwd(q, k, v, o, lse, B, H, Lq, Lk, dh, kmask=None, causal=False, p_drop=0.0, seed=0, site=0):
    """q/k/v/o: 2-D views [B*L, ld] (row pitch = stride(0)); heads are column blocks of width dh.
    bf16 with head dim 64 runs on the tensor-core kernels, everything else on the SIMT fp32-arithmetic ones."""
    _need_cuda(q, k, v, o)
    if _attn_tc_ok(dh, q, k, v, o) and Lq <= 128 and Lk <= 128 and ATTN_IMPL == "tcgen05":
        check(_lib.load().mma_attn_fwd_t5(
            q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), _p(kmask), o.data_ptr(),
            o.stride(0), _p(lse), B, H, Lq, Lk, int(causal), dh ** -0.5, float(p_drop), int(seed), int(site),
            _stream()), "mma_attn_fwd_t5")
        _count()
        return
    if _attn_tc_ok(dh, q, k, v, o) and Lq <= 512 and Lk <= 512 and ATTN_IMPL == "tcgen05" and lse is not None:
        # 128 < L <= 512: blocked tcgen05 kernels (128 x 128 tile problems + merge); workspaces are cached per shape so
        # that captured CUDA graphs keep stable addresses
        nkb = (Lk + 127) // 128
        ws_o = _att_ws(q.device, "fwd_o", (nkb, B * Lq, H * dh))
        ws_l = _att_ws(q.device, "fwd_lse", (nkb, B * H * Lq), torch.float32)
        check(_lib.load().mma_attn_fwd_t5b(
            q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), _p(kmask), o.data_ptr(),
            o.stride(0), lse.data_ptr(), ws_o.data_ptr(), ws_l.data_ptr(), B, H, Lq, Lk, int(causal), dh ** -0.5,
            float(p_drop), int(seed), int(site), _stream()), "mma_attn_fwd_t5b")
        _count(2)
        return
    if _attn_tc_ok(dh, q, k, v, o):
        check(_lib.load().mma_attn_fwd_tc(
            q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), _p(kmask), o.data_ptr(),
            o.stride(0), _p(lse), B, H, Lq, Lk, int(causal), dh ** -0.5, float(p_drop), int(seed), int(site),
            _stream()), "mma_attn_fwd_tc")
        _count()
        return
    check(_lib.load().mma_attn_fwd(
        q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), _p(kmask), o.data_ptr(),
        o.stride(0), _p(lse), B, H, Lq, Lk, dh, int(causal), dh ** -0.5, float(p_drop), int(seed), int(site), _ty(q),
        _stream()), "mma_attn_fwd")
    _count()


def attn_bwd(q, k, v, o, lse, dout, dq, dk, dv, B, H, Lq, Lk, dh, kmask=None, causal=False, p_drop=0.0, seed=0,
             site=0, dsum=None):
    _need_cuda(q, k, v, o, dout, dq, dk, dv)
    if _attn_tc_ok(dh, q, k, v, o, dout, dq, dk, dv) and Lq <= 128 and Lk <= 128 and ATTN_IMPL == "tcgen05":
        check(_lib.load().mma_attn_bwd_t5(
            q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), _p(kmask), o.data_ptr(),
            o.stride(0), lse.data_ptr(), dout.data_ptr(), dout.stride(0), dq.data_ptr(), dq.stride(0), dk.data_ptr(),
            dk.stride(0), dv.data_ptr(), dv.stride(0), B, H, Lq, Lk, int(causal), dh ** -0.5, float(p_drop), int(seed),
            int(site), _stream()), "mma_attn_bwd_t5")
        _count()
        return
    if _attn_tc_ok(dh, q, k, v, o, dout, dq, dk, dv) and Lq <= 512 and Lk <= 512 and ATTN_IMPL == "tcgen05":
        nqb, nkb = (Lq + 127) // 128, (Lk + 127) // 128
        ws_dq = _att_ws(q.device, "bwd_dq", (nkb, B * Lq, H * dh))
        ws_dk = _att_ws(q.device, "bwd_dk", (nqb, B * Lk, H * dh))
        ws_dv = _att_ws(q.device, "bwd_dv", (nqb, B * Lk, H * dh))
        check(_lib.load().mma_attn_bwd_t5b(
            q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), _p(kmask), o.data_ptr(),
            o.stride(0), lse.data_ptr(), dout.data_ptr(), dout.stride(0), dq.data_ptr(), dq.stride(0), dk.data_ptr(),
            dk.stride(0), dv.data_ptr(), dv.stride(0), ws_dq.data_ptr(), ws_dk.data_ptr(), ws_dv.data_ptr(), B, H, Lq, Lk,
            int(causal), dh ** -0.5, float(p_drop), int(seed), int(site), _stream()), "mma_attn_bwd_t5b")
        _count(2)
        return
    if _attn_tc_ok(dh, q, k, v, o, dout, dq, dk, dv):
        if dsum is None:
            dsum = torch.empty(B * H * Lq, dtype=torch.float32, device=q.device)
        check(_lib.load().mma_attn_bwd_tc(
            q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), _p(kmask), o.data_ptr(),
            o.stride(0), lse.data_ptr(), dsum.data_ptr(), dout.data_ptr(), dout.stride(0), dq.data_ptr(), dq.stride(0),
            dk.data_ptr(), dk.stride(0), dv.data_ptr(), dv.stride(0), B, H, Lq, Lk, int(causal), dh ** -0.5,
            float(p_drop), int(seed), int(site), _stream()), "mma_attn_bwd_tc")
        _count(2)
        return
    check(_lib.load().mma_attn_bwd(
        q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), _p(kmask), o.data_ptr(),
        o.stride(0), lse.data_ptr(), dout.data_ptr(), dout.stride(0), dq.data_ptr(), dq.stride(0), dk.data_ptr(),
        dk.stride(0), dv.data_ptr(), dv.stride(0), B, H, Lq, Lk, dh, int(causal), dh ** -0.5, float(p_drop), int(seed),
        int(site), _ty(q), _stream()), "mma_attn_bwd")
    _count(2)


def ce_fwd(logits, labels, V, row_loss, row_lse, stats, smoothing=0.0, ignore_index=-100):
    _need_cuda(logits, labels)
    assert logits.dtype == torch.float32 and labels.dtype == torch.int64
    check(_lib.load().mma_ce_fwd(logits.data_ptr(), logits.stride(0), labels.data_ptr(), labels.numel(), V,
                                 float(smoothing), ignore_index, row_loss.data_ptr(), row_lse.data_ptr(),
                                 stats.data_ptr(), _stream()), "mma_ce_fwd")
    _count(2)


def ce_bwd(logits, labels, V, row_lse, stats, dlogits, gscale=1.0, smoothing=0.0, ignore_index=-100):
    _need_cuda(logits, labels, dlogits)
    check(_lib.load().mma_ce_bwd(logits.data_ptr(), logits.stride(0), labels.data_ptr(), row_lse.data_ptr(),
                                 stats.data_ptr(), float(gscale), labels.numel(), V, float(smoothing), ignore_index,
                                 dlogits.data_ptr(), _ty(dlogits), dlogits.stride(0), _stream()), "mma_ce_bwd")
    _count()


def grad_norm(g, workspace, norm):
    _need_cuda(g, workspace, norm)
    check(_lib.load().mma_grad_norm(g.data_ptr(), g.numel(), workspace.data_ptr(), norm.data_ptr(), _stream()),
          "mma_grad_norm")
    _count(2)


def adam_step(p, g, m, v, p_bf16, hyper, norm=None, decoupled=True, zero_grad=True):
    _need_cuda(p, g, m, v, hyper)
    check(_lib.load().mma_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _p(p_bf16), p.numel(),
                                    hyper.data_ptr(), _p(norm), int(decoupled), int(zero_grad), _stream()),
          "mma_adam_step")
    _count()


def _ptr_array(ptrs):
    return (C.c_void_p * len(ptrs))(*[int(x) for x in ptrs])


def p2p_barrier(peer_flags, epoch_ctr, world, rank):
    """Cross-GPU barrier over flag words in symmetric memory (peer_flags: the ranks' device pointers)."""
    check(_lib.load().mma_p2p_barrier(_ptr_array(peer_flags), epoch_ctr.data_ptr(), world, rank, _stream()),
          "mma_p2p_barrier")
    _count()


def p2p_reduce_shard(peer_g, world, rank, lo, hi, workspace, sumsq_out, mc_g=0):
    check(_lib.load().mma_p2p_reduce_shard(_ptr_array(peer_g), int(mc_g) or None, world, rank, lo, hi, workspace.data_ptr(),
                                           sumsq_out.data_ptr(), _stream()), "mma_p2p_reduce_shard")
    _count(2)


def p2p_adam_shard(p, g, m, v, peer_pb, peer_sumsq, world, rank, lo, hi, hyper, decoupled=True, mc_pb=0):
    check(_lib.load().mma_p2p_adam_shard(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _ptr_array(peer_pb),
                                         int(mc_pb) or None, _ptr_array(peer_sumsq), world, rank, p.numel(), lo, hi,
                                         hyper.data_ptr(), int(decoupled), _stream()), "mma_p2p_adam_shard")
    _count()


def add_u64(t, inc=1):
    """t: int64 device scalar holding the dropout seed; advanced on the stream (graph-capturable)."""
    check(_lib.load().mma_add_u64(t.data_ptr(), int(inc), _stream()), "mma_add_u64")
    _count()


SMALL_KINDS = {"store": 0, "gelu": 1, "resid": 2, "glu": 3}


def small_linear(x, w, out, R, N, K, kind="store", bias=None, gamma=None, beta=None, w2=None, bias2=None, resid=None,
                 eps=1e-5):
    """out[R, N] = epilogue(LN?(x)[R, K] w^T + bias) for R <= 512 rows (blocks of <= 64) in ONE launch (decode of a few spectra).
    Returns False (nothing launched) outside the kernel's envelope."""
    _need_cuda(x, w, out)
    rc = _lib.load().mma_small_linear(
        x.data_ptr(), _ty(x), x.stride(0), _p(gamma), _p(beta), float(eps), w.data_ptr(), _p(w2), w.stride(0), _p(bias),
        _p(bias2), _p(resid), resid.stride(0) if resid is not None else 0, out.data_ptr(), _ty(out), out.stride(0), R, N, K,
        SMALL_KINDS[kind], _stream())
    if rc == -3:
        return False
    check(rc, "mma_small_linear")
    _count()
    return True


def decode_step(args, cluster_size):
    """The whole decoder step (logits of every row) in one launch; `args` is a filled `_lib.DecodeStep`.
    Returns False (nothing launched) outside the kernel's envelope."""
    import ctypes
    rc = _lib.load().mma_decode_step(ctypes.addressof(args), int(cluster_size), _stream())
    if rc == -3:
        return False
    check(rc, "mma_decode_step")
    _count()
    return True


def decode_step_max_clusters(cluster_size):
    """Resident clusters of the one-launch decode step on this device (0: the cluster size cannot be scheduled)."""
    return int(_lib.load().mma_decode_step_max_clusters(int(cluster_size)))


def decode_embed(tok, table, gamma, beta, pos, cur_len, out, eps=1e-5):
    _need_cuda(tok, table, out)
    check(_lib.load().mma_decode_embed(tok.data_ptr(), table.data_ptr(), _p(gamma), _p(beta), eps, pos.data_ptr(),
                                       cur_len.data_ptr(), out.data_ptr(), tok.numel(), table.shape[1], _stream()),
          "mma_decode_embed")
    _count()


def decode_self_attn(q, knew, vnew, kcache, vcache, anc, cur_len, o, R, H, dh, Lmax, beams=1):
    _need_cuda(q, knew, vnew, kcache, vcache, o)
    check(_lib.load().mma_decode_self_attn(
        q.data_ptr(), q.stride(0), knew.data_ptr(), vnew.data_ptr(), knew.stride(0), kcache.data_ptr(),
        vcache.data_ptr(), _p(anc), cur_len.data_ptr(), o.data_ptr(), o.stride(0), R, H, dh, Lmax, dh ** -0.5, _ty(q),
        int(beams), _stream()), "mma_decode_self_attn")
    _count()


def decode_cross_attn(q, kmem, vmem, kmask, cur_len, o, R, H, dh, S, beams):
    _need_cuda(q, kmem, vmem, o)
    check(_lib.load().mma_decode_cross_attn(
        q.data_ptr(), q.stride(0), kmem.data_ptr(), vmem.data_ptr(), kmem.stride(0), _p(kmask), cur_len.data_ptr(),
        o.data_ptr(), o.stride(0), R, H, dh, S, beams, dh ** -0.5, _ty(q), _stream()), "mma_decode_cross_attn")
    _count()


def _guide_args(guide):
    """guide: None or (cur_counts int32 [rows, NA], target_counts int32 [spectra, NA], tok_atoms int32 [V], n_check)."""
    if guide is None:
        return 0, 0, 0, 0, 0
    cur, tgt, tok, n_check = guide
    _need_cuda(cur, tgt, tok)
    assert cur.dtype == torch.int32 and tgt.dtype == torch.int32 and tok.dtype == torch.int32
    assert cur.is_contiguous() and tgt.is_contiguous() and cur.shape[1] == tgt.shape[1]
    return cur.data_ptr(), tgt.data_ptr(), tok.data_ptr(), cur.shape[1], int(n_check)


def beam_step(logits, V, st, extra_bias=None, prenorm=False, guide=None):
    """st: decode.BeamState (device buffers).  prenorm: `logits` already holds processed log-probabilities."""
    check(_lib.load().mma_beam_step_ex(
        logits.data_ptr(), logits.stride(0), _p(extra_bias), st.B, st.K, V, st.L, st.pad_id, st.eos_id,
        st.cur_len.data_ptr(), st.run_seq.data_ptr(), st.fin_seq.data_ptr(), st.run_score.data_ptr(),
        st.fin_score.data_ptr(), st.fin_flag.data_ptr(), st.fin_len.data_ptr(), st.improvable.data_ptr(),
        st.all_hit.data_ptr(), st.anc.data_ptr(), st.next_tok.data_ptr(), st.parent_row.data_ptr(),
        1 if prenorm else 0, *_guide_args(guide), _stream()), "mma_beam_step_ex")
    _count()


def greedy_step(logits, V, st, extra_bias=None, prenorm=False, guide=None):
    check(_lib.load().mma_greedy_step_ex(
        logits.data_ptr(), logits.stride(0), _p(extra_bias), st.B, V, st.L, st.pad_id, st.eos_id,
        st.cur_len.data_ptr(), st.run_seq.data_ptr(), st.unfinished.data_ptr(), st.next_tok.data_ptr(),
        1 if prenorm else 0, *_guide_args(guide), _stream()), "mma_greedy_step_ex")
    _count()


def score_rows(logits, out, V, L, eos_id, cur_len, log_softmax):
    """Dense scores for host-visible logits processors (log_softmax or raw logits, then ForcedEOS)."""
    _need_cuda(logits, out)
    check(_lib.load().mma_score_rows(logits.data_ptr(), logits.stride(0), out.data_ptr(), out.stride(0),
                                     logits.shape[0], V, L, eos_id, cur_len.data_ptr(), 1 if log_softmax else 0,
                                     _stream()), "mma_score_rows")
    _count()


def guided_mask(scores, eos_id, beams, guide):
    """GuidedFormulaProcessor.__call__ on a dense fp32 [R, V] matrix, in place."""
    _need_cuda(scores)
    assert scores.dtype == torch.float32 and scores.stride(1) == 1
    check(_lib.load().mma_guided_mask(scores.data_ptr(), scores.stride(0), scores.shape[0], scores.shape[1], eos_id,
                                      beams, *_guide_args(guide), _stream()), "mma_guided_mask")
    _count()


def advance(cur_len):
    check(_lib.load().mma_advance(cur_len.data_ptr(), _stream()), "mma_advance")
    _count()
