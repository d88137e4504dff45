"""GPU unit tests of the individual CUDA kernels (called through the C ABI via ops.py) against plain
PyTorch fp32 references of the same op."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from multimodalanalytical_b200 import ops
    from multimodalanalytical_b200._lib import (EPI_ACCUM, EPI_DGELU, EPI_DGLU, EPI_GELU, EPI_GLU_MUL, EPI_RESID,
                                                EPI_STORE)

DEV = "cuda"


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _rand(*shape, dtype=torch.float32, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


# ------------------------------------------------------------------------------------------- GEMM
GEMM_SHAPES = [(128, 128, 64), (256, 256, 512), (300, 200, 136), (1000, 512, 2048), (64, 48, 48), (130, 37, 48),
               (4096, 1536, 512)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_gemm_fwd_kmajor(M, N, K, dtype):
    dt = torch.bfloat16 if dtype == "bf16" else torch.float32
    A, W = _rand(M, K, dtype=dt), _rand(N, K, dtype=dt, scale=K ** -0.5)
    bias = _rand(N)
    out = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.gemm(A, W, M, N, K, ops.make_epi(EPI_STORE, out, bias=bias))
    ref = A.float() @ W.float().T + bias
    assert rel(out, ref) < (2e-5 if dtype == "f32" else 1e-5), "fp32-out"
    out2 = torch.empty(M, N, device=DEV, dtype=dt)
    ops.gemm(A, W, M, N, K, ops.make_epi(EPI_STORE, out2, bias=bias))
    assert rel(out2.float(), ref) < (2e-5 if dtype == "f32" else 1e-2)


@pytest.mark.parametrize("M,N,K", [(256, 512, 200), (1000, 48, 37 + 3), (512, 512, 2048), (130, 48, 40)])
@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_gemm_dgrad_b_mnmajor(M, N, K, dtype):
    """dx[M,N] = dy[M,K] @ W[K,N]  (W stored [K(out of fwd), N(in of fwd)] row-major = MN-major B)."""
    dt = torch.bfloat16 if dtype == "bf16" else torch.float32
    dy, W = _rand(M, K, dtype=dt), _rand(K, N, dtype=dt, scale=K ** -0.5)
    out = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.gemm(dy, W, M, N, K, ops.make_epi(EPI_STORE, out), b_mn=True)
    assert rel(out, dy.float() @ W.float()) < 2e-5


@pytest.mark.parametrize("R,N,K", [(512, 128, 64), (1000, 200, 48), (4096, 512, 512), (777, 37, 48), (9216, 2048, 512)])
@pytest.mark.parametrize("dtype", ["bf16", "f32"])
@pytest.mark.parametrize("splits", [1, 5])
def test_gemm_wgrad_mnmajor_splitk(R, N, K, dtype, splits):
    """dW[N,K] = dy[R,N]^T @ x[R,K]; both operands MN-major, reduction over rows, optional split-K."""
    dt = torch.bfloat16 if dtype == "bf16" else torch.float32
    dy, x = _rand(R, N, dtype=dt, scale=R ** -0.5), _rand(R, K, dtype=dt)
    out = torch.zeros(N, K, device=DEV, dtype=torch.float32)
    ops.gemm(dy, x, N, K, R, ops.make_epi(EPI_ACCUM, out, accumulate=2 if splits > 1 else 0), a_mn=True, b_mn=True,
             splits=splits)
    assert rel(out, dy.float().T @ x.float()) < 3e-5


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_gemm_epilogues(dtype):
    dt = torch.bfloat16 if dtype == "bf16" else torch.float32
    tol = 1e-2 if dtype == "bf16" else 2e-5
    M, N, K = 300, 256, 128
    A, W, bias = _rand(M, K, dtype=dt), _rand(N, K, dtype=dt, scale=K ** -0.5), _rand(N)
    acc = A.float() @ W.float().T
    # GELU (+ pre-activation copy)
    z, a = torch.empty(M, N, device=DEV, dtype=dt), torch.empty(M, N, device=DEV, dtype=dt)
    ops.gemm(A, W, M, N, K, ops.make_epi(EPI_GELU, a, out2=z, bias=bias))
    assert rel(z.float(), acc + bias) < tol
    assert rel(a.float(), torch.nn.functional.gelu(acc + bias)) < tol
    # residual
    resid = _rand(M, N)
    o = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.gemm(A, W, M, N, K, ops.make_epi(EPI_RESID, o, bias=bias, resid=resid))
    assert rel(o, resid + acc + bias) < 2e-5
    # dGELU
    zz = _rand(M, N, dtype=dt)
    o2 = torch.empty(M, N, device=DEV, dtype=dt)
    ops.gemm(A, W, M, N, K, ops.make_epi(EPI_DGELU, o2, aux=zz))
    zf = zz.float().requires_grad_(True)
    torch.nn.functional.gelu(zf).backward(acc)
    assert rel(o2.float(), zf.grad) < tol
    # GLU forward: out = gelu(z1) * (acc + bias), out2 = acc + bias
    o3, z2 = torch.empty(M, N, device=DEV, dtype=dt), torch.empty(M, N, device=DEV, dtype=dt)
    ops.gemm(A, W, M, N, K, ops.make_epi(EPI_GLU_MUL, o3, out2=z2, bias=bias, aux=zz))
    assert rel(o3.float(), torch.nn.functional.gelu(zz.float()) * (acc + bias)) < tol
    # GLU backward
    z1, z2b = _rand(M, N, dtype=dt, seed=3), _rand(M, N, dtype=dt, seed=4)
    d1, d2 = torch.empty(M, N, device=DEV, dtype=dt), torch.empty(M, N, device=DEV, dtype=dt)
    ops.gemm(A, W, M, N, K, ops.make_epi(EPI_DGLU, d1, out2=d2, aux=z1, aux2=z2b))
    z1f, z2f = z1.float().requires_grad_(True), z2b.float().requires_grad_(True)
    (torch.nn.functional.gelu(z1f) * z2f).backward(acc)
    assert rel(d1.float(), z1f.grad) < tol and rel(d2.float(), z2f.grad) < tol
    # accumulate (beta = 1)
    base = _rand(M, N)
    o4 = base.clone()
    ops.gemm(A, W, M, N, K, ops.make_epi(EPI_ACCUM, o4, accumulate=1))
    assert rel(o4, base + acc) < 2e-5


@pytest.mark.parametrize("M,N,K", [(16384, 512, 512), (12300, 1536, 512), (12416, 600, 200), (9216, 2048, 512),
                                   (16384, 512, 2048)])
def test_gemm_pair_kernel(M, N, K):
    """The CTA-pair kernel (gemm_tc2.cu: cta_group::2, 256x256 tiles, TMA epilogue) takes the large products; every
    instantiated epilogue against torch, and with dropout against the single-CTA kernel (max_ctas > 0 forces it)."""
    dt, tol = torch.bfloat16, 1e-2
    A, W, bias = _rand(M, K, dtype=dt), _rand(N, K, dtype=dt, scale=K ** -0.5), _rand(N)
    acc = A.float() @ W.float().T
    # forward STORE bf16 / fp32
    o = torch.empty(M, N, device=DEV, dtype=dt)
    ops.gemm(A, W, M, N, K, ops.make_epi(EPI_STORE, o, bias=bias))
    assert rel(o.float(), acc + bias) < tol
    of = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.gemm(A, W, M, N, K, ops.make_epi(EPI_STORE, of, bias=bias))
    assert rel(of, acc + bias) < 1e-5
    # GELU + pre-activation copy
    z, a = torch.empty(M, N, device=DEV, dtype=dt), torch.empty(M, N, device=DEV, dtype=dt)
    ops.gemm(A, W, M, N, K, ops.make_epi(EPI_GELU, a, out2=z, bias=bias))
    assert rel(z.float(), acc + bias) < tol
    assert rel(a.float(), torch.nn.functional.gelu(acc + bias)) < tol
    a1 = torch.empty(M, N, device=DEV, dtype=dt)
    ops.gemm(A, W, M, N, K, ops.make_epi(EPI_GELU, a1, bias=bias))
    assert torch.equal(a1, a)
    # residual, fp32 stream
    resid = _rand(M, N)
    r = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.gemm(A, W, M, N, K, ops.make_epi(EPI_RESID, r, bias=bias, resid=resid))
    assert rel(r, resid + acc + bias) < 2e-5
    # dgrad products: B is [K, N] row-major
    Wt = W.t().contiguous()
    d = torch.empty(M, N, device=DEV, dtype=dt)
    ops.gemm(A, Wt, M, N, K, ops.make_epi(EPI_STORE, d), b_mn=True)
    assert rel(d.float(), acc) < tol
    zz = _rand(M, N, dtype=dt, seed=5)
    dg = torch.empty(M, N, device=DEV, dtype=dt)
    ops.gemm(A, Wt, M, N, K, ops.make_epi(EPI_DGELU, dg, aux=zz), b_mn=True)
    zf = zz.float().requires_grad_(True)
    torch.nn.functional.gelu(zf).backward(acc)
    assert rel(dg.float(), zf.grad) < tol
    base = _rand(M, N, seed=9)
    ac = base.clone()
    ops.gemm(A, Wt, M, N, K, ops.make_epi(EPI_ACCUM, ac, accumulate=1), b_mn=True)
    assert rel(ac, base + acc) < 2e-5
    ops.gemm(A, Wt, M, N, K, ops.make_epi(EPI_ACCUM, ac, accumulate=0), b_mn=True)
    assert rel(ac, acc) < 2e-5
    # dropout: same masks and values as the single-CTA kernel
    for kind, kw, bmn in ((EPI_GELU, dict(bias=bias), False), (EPI_RESID, dict(bias=bias, resid=resid), False),
                          (EPI_DGELU, dict(aux=zz), True)):
        odt = torch.float32 if kind == EPI_RESID else dt
        x1, x2 = torch.empty(M, N, device=DEV, dtype=odt), torch.empty(M, N, device=DEV, dtype=odt)
        Bm = Wt if bmn else W
        ops.gemm(A, Bm, M, N, K, ops.make_epi(kind, x1, p_drop=0.1, seed=123, site=7, **kw), b_mn=bmn)
        ops.gemm(A, Bm, M, N, K, ops.make_epi(kind, x2, p_drop=0.1, seed=123, site=7, **kw), b_mn=bmn, max_ctas=148)
        assert rel(x1.float(), x2.float()) < 1e-2
        if kind != EPI_RESID:
            assert torch.equal(x1 == 0, x2 == 0)


@pytest.mark.parametrize("M,N,K", [(16384, 2048, 512), (9216, 2048, 512), (12300, 1040, 200), (2560, 2048, 512),
                                   (25472, 3072, 768)])
def test_ffn_glu_pair_kernels(M, N, K):
    """Gated FFN on the CTA-pair path (gemm_glu2.cu): the fused W1 | Wg forward, the DGLU dgrad and the two-reduction
    dh product against torch autograd of gelu(h W1^T + b1) * (h Wg^T + bg); with dropout against the single-CTA
    kernels' masks and values."""
    dt, tol = torch.bfloat16, 1e-2
    F = torch.nn.functional
    h = _rand(M, K, dtype=dt)
    W1, Wg = _rand(N, K, dtype=dt, scale=K ** -0.5, seed=1), _rand(N, K, dtype=dt, scale=K ** -0.5, seed=2)
    b1, bg = _rand(N, seed=3), _rand(N, seed=4)
    a, z1, z2 = (torch.empty(M, N, device=DEV, dtype=dt) for _ in range(3))
    assert ops.ffn_glu_fwd(h, W1, Wg, b1, bg, M, N, K, a, z1=z1, z2=z2)
    r1 = h.float() @ W1.float().T + b1
    r2 = h.float() @ Wg.float().T + bg
    assert rel(z1.float(), r1) < tol and rel(z2.float(), r2) < tol
    assert rel(a.float(), F.gelu(r1) * r2) < tol
    a_inf = torch.empty_like(a)
    assert ops.ffn_glu_fwd(h, W1, Wg, b1, bg, M, N, K, a_inf)  # inference form: no saved pre-activations
    assert torch.equal(a_inf, a)
    # dropout: mask identical to the EPI_GLU_MUL epilogue of the single-CTA kernel (same seed / site / index)
    ad, zs, ad2, z2b = (torch.empty(M, N, device=DEV, dtype=dt) for _ in range(4))
    assert ops.ffn_glu_fwd(h, W1, Wg, b1, bg, M, N, K, ad, p_drop=0.1, seed=321, site=9)
    ops.gemm(h, W1, M, N, K, ops.make_epi(EPI_STORE, zs, bias=b1), max_ctas=148)
    ops.gemm(h, Wg, M, N, K, ops.make_epi(EPI_GLU_MUL, ad2, out2=z2b, bias=bg, aux=zs, p_drop=0.1, seed=321, site=9),
             max_ctas=148)
    # (the one-MUFU GELU is exactly 0 once tanh saturates, z1 < -5.6: there a zero is not a dropped element, and the two
    # kernels see z1 in fp32 / rounded to bf16)
    nz0 = z1.float() > -5.0
    assert torch.equal((ad == 0) & nz0, (ad2 == 0) & nz0)
    assert rel(ad.float(), ad2.float()) < 2e-2
    keep = (ad != 0).float().mean().item()
    assert abs(keep - 0.9) < 0.01

    # backward through the gate: dy [M, Kd] @ W2 [Kd, N]
    Kd = K
    dy = _rand(M, Kd, dtype=dt, seed=5)
    W2 = _rand(Kd, N, dtype=dt, scale=N ** -0.5, seed=6)
    dz1, dz2 = torch.empty(M, N, device=DEV, dtype=dt), torch.empty(M, N, device=DEV, dtype=dt)
    assert ops.ffn_dglu(dy, W2, M, N, Kd, z1, z2, dz1, dz2)
    da = dy.float() @ W2.float()
    zf1, zf2 = z1.float().requires_grad_(True), z2.float().requires_grad_(True)
    (F.gelu(zf1) * zf2).backward(da)
    assert rel(dz1.float(), zf1.grad) < tol and rel(dz2.float(), zf2.grad) < tol
    e1, e2, f1, f2 = (torch.empty(M, N, device=DEV, dtype=dt) for _ in range(4))
    assert ops.ffn_dglu(dy, W2, M, N, Kd, z1, z2, e1, e2, p_drop=0.1, seed=321, site=9, drop_ld=N)
    ops.gemm(dy, W2, M, N, Kd, ops.make_epi(EPI_DGLU, f1, out2=f2, aux=z1, aux2=z2, p_drop=0.1, seed=321, site=9,
                                            drop_ld=N), b_mn=True, max_ctas=148)
    assert torch.equal(e2 == 0, f2 == 0)
    # backward mask == forward mask, wherever the un-dropped values are not exactly zero themselves (one fp32 sum in
    # ~1e8 cancels its bias exactly)
    nz = (a != 0) & (dz2 != 0)
    assert torch.equal((e2 == 0) & nz, (ad == 0) & nz)
    assert rel(e1.float(), f1.float()) < 2e-2 and rel(e2.float(), f2.float()) < 2e-2

    # dh = dz1 W1 + dz2 Wg in one accumulation (W1 / Wg are [N, K] = [reduction, out]: MN-major B operands)
    dh = torch.empty(M, K, device=DEV, dtype=dt)
    ok = ops.gemm_dual(dz1, W1, dz2, Wg, M, K, N, N, ops.make_epi(EPI_STORE, dh), b_mn=True)
    tiles = ((M + 255) // 256) * ((K + 255) // 256)
    assert ok == (K >= 256 and tiles >= 48)
    if ok:
        assert rel(dh.float(), dz1.float() @ W1.float() + dz2.float() @ Wg.float()) < tol


def test_ffn_glu_pair_kernels_decline_small_or_misaligned():
    dt = torch.bfloat16
    h, W = _rand(64, 512, dtype=dt), _rand(2048, 512, dtype=dt)
    b = _rand(2048)
    a = torch.empty(64, 2048, device=DEV, dtype=dt)
    assert not ops.ffn_glu_fwd(h, W, W, b, b, 64, 2048, 512, a)  # too few rows for the pair kernel
    h2, W3 = _rand(4096, 512, dtype=dt), _rand(2040, 512, dtype=dt)
    a2 = torch.empty(4096, 2040, device=DEV, dtype=dt)
    assert not ops.ffn_glu_fwd(h2, W3, W3, b[:2040], b[:2040], 4096, 2040, 512, a2)  # N % 16 != 0


@pytest.mark.parametrize("M,K", [(16384, 512), (9216, 2048), (12300, 512), (640, 512), (2560, 512), (2560, 2048), (200, 512),
                                 (4700, 2048)])
def test_gemm_resid_layernorm_fused(M, K):
    """x_new = resid + drop(A W^T + b) and h = LN(x_new) in one launch (gemm2_ln_kernel) against torch; with dropout
    against the unfused pair of launches (same mask)."""
    N = 512
    A, W, bias = _rand(M, K, dtype=torch.bfloat16), _rand(N, K, dtype=torch.bfloat16, scale=K ** -0.5), _rand(N)
    resid = _rand(M, N, seed=3)
    gamma, beta = _rand(N, seed=4) * 0.5 + 1.0, _rand(N, seed=5) * 0.1
    x = torch.empty(M, N, device=DEV)
    h = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ok = ops.gemm_resid_ln(A, W, M, N, K, ops.make_epi(EPI_RESID, x, bias=bias, resid=resid), gamma, beta, h)
    assert ok  # <= 4736 rows: 4-CTA-cluster kernel (gemm_tc.cu, LNF); above: CTA-pair kernel (gemm2_ln_kernel)
    ref_x = resid + A.float() @ W.float().T + bias
    assert rel(x, ref_x) < 2e-5
    ref_h = torch.nn.functional.layer_norm(ref_x, (N,), gamma, beta, 1e-5)
    assert rel(h.float(), ref_h) < 1e-2
    # in place on the residual stream (out == resid) must work too
    x3 = resid.clone()
    h1 = torch.empty_like(h)
    assert ops.gemm_resid_ln(A, W, M, N, K, ops.make_epi(EPI_RESID, x3, bias=bias, resid=x3), gamma, beta, h1)
    assert rel(x3, ref_x) < 2e-5 and rel(h1.float(), ref_h) < 1e-2
    if M < 512:  # with dropout only the CTA-pair kernel applies
        assert not ops.gemm_resid_ln(A, W, M, N, K, ops.make_epi(EPI_RESID, x3, bias=bias, resid=resid, p_drop=0.1, seed=11,
                                                                 site=5), gamma, beta, h1)
        return
    # dropout: same mask / values as gemm (RESID epilogue) + ln_fwd
    x1, x2 = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    h1, h2 = torch.empty_like(h), torch.empty_like(h)
    assert ops.gemm_resid_ln(A, W, M, N, K, ops.make_epi(EPI_RESID, x1, bias=bias, resid=resid, p_drop=0.1, seed=11, site=5),
                             gamma, beta, h1)
    ops.gemm(A, W, M, N, K, ops.make_epi(EPI_RESID, x2, bias=bias, resid=resid, p_drop=0.1, seed=11, site=5))
    ops.ln_fwd(x2, gamma, beta, h2)
    assert rel(x1, x2) < 1e-5
    assert rel(h1.float(), h2.float()) < 1e-2


def test_gemm_dropout_mask_consistency():
    """RESID-epilogue dropout (forward) and ln_bwd's masked copy (backward) must use the same mask."""
    M, N, K, p = 256, 128, 64, 0.25
    A, W = _rand(M, K, dtype=torch.bfloat16), _rand(N, K, dtype=torch.bfloat16)
    resid = torch.zeros(M, N, device=DEV)
    o = torch.empty(M, N, device=DEV)
    ops.gemm(A, W, M, N, K, ops.make_epi(EPI_RESID, o, resid=resid, p_drop=p, seed=77, site=5))
    acc = A.float() @ W.float().T
    keep = (o != 0)
    frac = keep.float().mean().item()
    assert abs(frac - (1 - p)) < 0.02
    assert rel(o[keep], (acc / (1 - p))[keep]) < 1e-5
    x = _rand(M, N)
    dy = torch.ones(M, N, device=DEV)
    dxb = torch.empty(M, N, device=DEV)
    ops.ln_bwd(dy, x, None, dxb=dxb, p_drop=p, seed=77, site=5)
    assert torch.equal(dxb != 0, keep)


# ------------------------------------------------------------------------------------------- row ops
@pytest.mark.parametrize("d", [512, 48, 1024])
@pytest.mark.parametrize("xdt", [torch.float32, torch.bfloat16])
def test_layernorm_fwd_bwd(d, xdt):
    rows = 777
    x = _rand(rows, d, dtype=xdt) * 2 + 0.5
    gamma, beta = _rand(d) + 1.0, _rand(d)
    y = torch.empty(rows, d, device=DEV)
    yb = torch.empty(rows, d, device=DEV, dtype=torch.bfloat16)
    ops.ln_fwd(x, gamma, beta, y, y2=yb)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (d,), gr, br)
    assert rel(y, ref) < 1e-5
    assert rel(yb.float(), ref) < 1e-2
    dy = _rand(rows, d, seed=9)
    dres = _rand(rows, d, seed=10)
    ref.backward(dy)
    dx = torch.empty(rows, d, device=DEV)
    dg, db = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    ops.ln_bwd(dy, x, gamma, dx=dx, dres=dres, dgamma=dg, dbeta=db)
    assert rel(dx, xr.grad + dres) < 2e-5
    assert rel(dg, gr.grad) < 1e-4 and rel(db, br.grad) < 1e-4


@pytest.mark.parametrize("xdt,dydt", [(torch.float32, torch.bfloat16), (torch.float32, torch.float32),
                                      (torch.bfloat16, torch.bfloat16)])
def test_layernorm_bwd_prefetching_variant(xdt, dydt):
    """rows >= 1024 at d = 512 take the bulk-copy prefetching kernel: against torch's LayerNorm backward, and bit for
    bit against the plain kernel (which rows < 1024 still use) on a prefix of the same data, in-place dres == dx,
    dropout-masked low-precision copy and remapped dy rows included."""
    rows, d, small = 5000, 512, 777
    x = _rand(rows, d, dtype=xdt) * 2 + 0.5
    gamma = _rand(d) + 1.0
    dy = _rand(rows, d, seed=9, dtype=dydt)
    dres = _rand(rows, d, seed=10)
    xr = x.detach().clone().float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), torch.zeros(d, device=DEV, requires_grad=True)
    torch.nn.functional.layer_norm(xr, (d,), gr, br).backward(dy.float())

    def run(n):
        dx = dres[:n].clone()  # in place: dres aliases dx, as the engine calls it
        dxb = torch.empty(n, d, device=DEV, dtype=torch.bfloat16)
        dg, db = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
        ops.ln_bwd(dy[:n], x[:n], gamma, dx=dx, dres=dx, dxb=dxb, dgamma=dg, dbeta=db, p_drop=0.1, seed=5, site=3)
        return dx, dxb, dg, db

    dx, dxb, dg, db = run(rows)
    assert rel(dx, xr.grad + dres) < 2e-5
    assert rel(dg, gr.grad) < 2e-4 and rel(db, br.grad) < 2e-4
    keep = dxb != 0
    assert abs(keep.float().mean().item() - 0.9) < 0.01
    assert rel(dxb.float()[keep], (dx / 0.9)[keep]) < 1e-2
    sdx, sdxb, _, _ = run(small)
    assert torch.equal(sdx, dx[:small]) and torch.equal(sdxb, dxb[:small])
    # dy read through the concat-by-offset row mapping (embedding backward): groups of 20 rows inside 36-row samples
    B, S, S_tot, off = 250, 20, 36, 16
    dyc = _rand(B * S_tot, d, seed=11, dtype=dydt)
    xs = x[: B * S]
    dx2 = torch.empty(B * S, d, device=DEV)
    ops.ln_bwd(dyc, xs, gamma, dx=dx2, group=S, in_group_stride=S_tot, in_offset=off)
    xr2 = xs.detach().clone().float().requires_grad_(True)
    sel = dyc.view(B, S_tot, d)[:, off: off + S].reshape(B * S, d).float()
    torch.nn.functional.layer_norm(xr2, (d,), gamma, None).backward(sel)
    assert rel(dx2, xr2.grad) < 2e-5


def test_embed_gather_ln_pos_concat_and_scatter():
    B, S1, S2, d, vocab = 5, 7, 4, 64, 30
    ids = torch.randint(0, vocab, (B, S1), device=DEV)
    table = _rand(vocab, d)
    scale = _rand(B, S1).abs() + 0.5
    pre2 = _rand(B * S2, d)
    g1, b1, g2, b2 = _rand(d) + 1, _rand(d), _rand(d, seed=2) + 1, _rand(d, seed=3)
    pos = _rand(64, d)
    out = torch.zeros(B * (S1 + S2), d, device=DEV)
    pre1 = torch.empty(B * S1, d, device=DEV)
    ops.gather_rows(ids.reshape(-1), table, pre1, scale=scale.reshape(-1))
    ops.ln_fwd(pre1, g1, b1, out, add=pos, group=S1, out_group_stride=S1 + S2, out_offset=0)
    ops.ln_fwd(pre2, g2, b2, out, add=pos, group=S2, out_group_stride=S1 + S2, out_offset=S1)
    e1 = torch.nn.functional.layer_norm(table[ids] * scale[..., None], (d,), g1, b1)
    e2 = torch.nn.functional.layer_norm(pre2.view(B, S2, d), (d,), g2, b2)
    ref = torch.cat([e1, e2], dim=1) + pos[: S1 + S2]
    assert rel(out.view(B, S1 + S2, d), ref) < 1e-5
    # scatter-add with padding row skipped
    g = _rand(B * S1, d, seed=5)
    dt = torch.zeros(vocab, d, device=DEV)
    ops.scatter_add_rows(ids.reshape(-1), g, dt, pad_idx=0, scale=scale.reshape(-1))
    ref_dt = torch.zeros(vocab, d, device=DEV).index_add_(0, ids.reshape(-1), g * scale.reshape(-1, 1))
    ref_dt[0] = 0
    assert rel(dt, ref_dt) < 1e-5


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_colsum(dt):
    x = _rand(3001, 300, dtype=dt)
    out = torch.zeros(300, device=DEV)
    ops.colsum(x, out)
    assert rel(out, x.float().sum(0)) < 1e-4


# ------------------------------------------------------------------------------------------- attention
def _attn_ref(q, k, v, kmask, causal, H):
    B, Lq, d = q.shape
    Lk = k.shape[1]
    dh = d // H
    qh = q.view(B, Lq, H, dh).transpose(1, 2)
    kh = k.view(B, Lk, H, dh).transpose(1, 2)
    vh = v.view(B, Lk, H, dh).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(dh)
    if kmask is not None:
        s = s.masked_fill(~kmask.bool()[:, None, None, :], float("-inf"))
    if causal:
        s = s + torch.triu(torch.full((Lq, Lk), float("-inf"), device=q.device), diagonal=1)
    return (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B, Lq, d)


@pytest.mark.parametrize("B,H,Lq,Lk,dh,causal,masked", [
    (3, 4, 57, 57, 16, True, True), (2, 8, 36, 36, 64, False, True), (2, 8, 64, 36, 64, False, True),
    (2, 2, 130, 199, 32, False, False), (1, 8, 128, 128, 64, True, False)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_attention_fwd_bwd(B, H, Lq, Lk, dh, causal, masked, dt):
    d = H * dh
    tol = 2e-5 if dt == torch.float32 else 2e-2
    q, k, v = _rand(B, Lq, d, dtype=dt, seed=1), _rand(B, Lk, d, dtype=dt, seed=2), _rand(B, Lk, d, dtype=dt, seed=3)
    kmask = None
    if masked:
        kmask = torch.ones(B, Lk, dtype=torch.uint8, device=DEV)
        for b in range(B):
            kmask[b, Lk - 1 - 3 * b - (b % 2) * 2: Lk - (b % 2) * 2] = 0
        kmask[:, 0] = 1
    o = torch.empty(B * Lq, d, device=DEV, dtype=dt)
    lse = torch.empty(B * H * Lq, device=DEV)
    ops.attn_fwd(q.view(-1, d), k.view(-1, d), v.view(-1, d), o, lse, B, H, Lq, Lk, dh, kmask=kmask, causal=causal)
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    ref = _attn_ref(qr, kr, vr, kmask, causal, H)
    assert rel(o.view(B, Lq, d).float(), ref) < tol
    do = _rand(B, Lq, d, dtype=dt, seed=4)
    ref.backward(do.float())
    dq, dk, dv = (torch.empty_like(t).view(-1, d) for t in (q, k, v))
    ops.attn_bwd(q.view(-1, d), k.view(-1, d), v.view(-1, d), o, lse, do.view(-1, d), dq, dk, dv, B, H, Lq, Lk, dh,
                 kmask=kmask, causal=causal)
    assert rel(dq.view_as(q).float(), qr.grad) < tol
    assert rel(dk.view_as(k).float(), kr.grad) < tol
    assert rel(dv.view_as(v).float(), vr.grad) < tol


def test_attention_packed_qkv_views():
    """q/k/v as column slices of one packed [M, 3d] buffer (how the model calls it)."""
    B, H, L, dh = 2, 4, 33, 16
    d = H * dh
    qkv = _rand(B * L, 3 * d)
    o = torch.empty(B * L, d, device=DEV)
    lse = torch.empty(B * H * L, device=DEV)
    ops.attn_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], o, lse, B, H, L, L, dh, causal=True)
    ref = _attn_ref(qkv[:, :d].reshape(B, L, d), qkv[:, d:2 * d].reshape(B, L, d), qkv[:, 2 * d:].reshape(B, L, d),
                    None, True, H)
    assert rel(o.view(B, L, d), ref) < 2e-5


def test_attention_dropout_statistics_and_grad_consistency():
    B, H, L, dh, p = 2, 2, 64, 16, 0.3
    d = H * dh
    q, k = _rand(B * L, d, seed=1), _rand(B * L, d, seed=2)
    v = torch.ones(B * L, d, device=DEV)
    o = torch.empty(B * L, d, device=DEV)
    lse = torch.empty(B * H * L, device=DEV)
    ops.attn_fwd(q, k, v, o, lse, B, H, L, L, dh, p_drop=p, seed=5, site=3)
    # with V == 1 the output is sum_j keep_ij p_ij / (1-p): mean 1
    assert abs(o.mean().item() - 1.0) < 0.02 and o.std().item() > 0.01
    o2 = torch.empty_like(o)
    ops.attn_fwd(q, k, v, o2, lse, B, H, L, L, dh, p_drop=p, seed=5, site=3)
    assert torch.equal(o, o2)


# ------------------------------------------------------------------------------------------- loss / optimiser
@pytest.mark.parametrize("V,smooth", [(200, 0.0), (37, 0.0), (26, 0.1)])
def test_cross_entropy(V, smooth):
    rows = 501
    ld = (V + 7) // 8 * 8
    logits = torch.zeros(rows, ld, device=DEV)
    logits[:, :V] = _rand(rows, V) * 3
    labels = torch.randint(0, V, (rows,), device=DEV)
    labels[::7] = -100
    rl, rlse, stats = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV), torch.empty(2, device=DEV)
    ops.ce_fwd(logits, labels, V, rl, rlse, stats, smoothing=smooth)
    xr = logits[:, :V].clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(xr, labels, ignore_index=-100, label_smoothing=smooth)
    assert abs(stats[0].item() - ref.item()) < 1e-5 * abs(ref.item())
    assert stats[1].item() == (labels != -100).sum().item()
    ref.backward()
    for dt in (torch.float32, torch.bfloat16):
        dl = torch.full((rows, ld), 7.0, device=DEV, dtype=dt)
        ops.ce_bwd(logits, labels, V, rlse, stats, dl, smoothing=smooth)
        assert rel(dl[:, :V].float(), xr.grad) < (1e-5 if dt == torch.float32 else 1e-2)
        assert (dl[:, V:] == 0).all()


@pytest.mark.parametrize("decoupled", [True, False])
def test_adam_step_matches_torch(decoupled):
    n = 100_003
    p0, g0 = _rand(n), _rand(n, seed=1) * 3
    ref_p = torch.nn.Parameter(p0.clone())
    opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)([ref_p], lr=1e-2, betas=(0.9, 0.999), eps=1e-8,
                                                                weight_decay=0.01)
    p, g = p0.clone(), g0.clone()
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    pb = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    ws, norm = torch.empty(1024, device=DEV), torch.empty(1, device=DEV)
    for t in range(1, 4):
        ref_p.grad = g0.clone()
        total = torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
        opt.step()
        g.copy_(g0)
        ops.grad_norm(g, ws, norm)
        assert abs(norm.item() - total.item()) < 1e-4 * total.item()
        hyper = torch.tensor([1e-2, 0.9, 0.999, 1e-8, 0.01, 1 - 0.9 ** t, 1 - 0.999 ** t, 1.0, 1.0], device=DEV)
        ops.adam_step(p, g, m, v, pb, hyper, norm=norm, decoupled=decoupled)
        assert rel(p, ref_p.data) < 1e-5
        assert (g == 0).all()
    assert rel(pb.float(), p) < 1e-2


def test_tc_attention_dropout_forward_backward_use_one_mask():
    """bf16 / head-dim-64 tensor-core path: recover the dropout mask from a forward with V = I, then check the
    forward output and all three gradients against autograd with that mask."""
    B, H, L, dh, p = 2, 2, 64, 64, 0.2
    d = H * dh
    dt = torch.bfloat16
    q, k = _rand(B * L, d, dtype=dt, seed=1), _rand(B * L, d, dtype=dt, seed=2)
    eye = torch.eye(L, dh, device=DEV).repeat(B, H).to(dt)  # V[b, j, h*dh + c] = [j == c]
    o = torch.empty(B * L, d, device=DEV, dtype=dt)
    lse = torch.empty(B * H * L, device=DEV)
    ops.attn_fwd(q, k, eye, o, lse, B, H, L, L, dh, causal=True, p_drop=p, seed=11, site=7)
    pd = o.view(B, L, H, dh).permute(0, 2, 1, 3).float()  # P_drop [B,H,Lq,Lk]
    qh = q.view(B, L, H, dh).permute(0, 2, 1, 3).float()
    kh = k.view(B, L, H, dh).permute(0, 2, 1, 3).float()
    s = qh @ kh.transpose(-1, -2) / math.sqrt(dh) + torch.triu(torch.full((L, L), float("-inf"), device=DEV), 1)
    pr = torch.softmax(s, -1)
    causal = torch.tril(torch.ones(L, L, device=DEV, dtype=torch.bool))
    keep = (pd != 0)
    frac = keep[..., causal].float().mean().item()
    assert abs(frac - (1 - p)) < 0.03, frac
    assert rel(pd[keep], (pr / (1 - p))[keep]) < 2e-2
    # gradients with the recovered mask
    v = _rand(B * L, d, dtype=dt, seed=3)
    do = _rand(B * L, d, dtype=dt, seed=4)
    o2 = torch.empty_like(o)
    ops.attn_fwd(q, k, v, o2, lse, B, H, L, L, dh, causal=True, p_drop=p, seed=11, site=7)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ops.attn_bwd(q, k, v, o2, lse, do, dq, dk, dv, B, H, L, L, dh, causal=True, p_drop=p, seed=11, site=7)
    qr, kr, vr = (t.float().view(B, L, H, dh).permute(0, 2, 1, 3).detach().requires_grad_(True) for t in (q, k, v))
    sr = qr @ kr.transpose(-1, -2) / math.sqrt(dh) + torch.triu(torch.full((L, L), float("-inf"), device=DEV), 1)
    ref = (torch.softmax(sr, -1) * keep / (1 - p)) @ vr
    assert rel(o2.view(B, L, H, dh).permute(0, 2, 1, 3).float(), ref) < 2e-2
    ref.backward(do.view(B, L, H, dh).permute(0, 2, 1, 3).float())
    for got, want in ((dq, qr.grad), (dk, kr.grad), (dv, vr.grad)):
        assert rel(got.view(B, L, H, dh).permute(0, 2, 1, 3).float(), want) < 3e-2


@pytest.mark.parametrize("impl", ["tcgen05", "mma"])
@pytest.mark.parametrize("B,H,Lq,Lk,causal", [(2, 8, 199, 199, False), (2, 8, 96, 199, False), (3, 8, 100, 100, True),
                                              (2, 8, 295, 295, False), (1, 8, 200, 200, True), (1, 4, 129, 512, False)])
def test_tc_attention_multi_tile(B, H, Lq, Lk, causal, impl, monkeypatch):
    """sequences longer than one tile (C4 multimodal shapes): blocked tcgen05 kernels (128 x 128 tile problems + merge)
    and the streaming mma.sync kernels, forward and backward against torch."""
    monkeypatch.setattr(ops, "ATTN_IMPL", impl)
    dh, dt = 64, torch.bfloat16
    d = H * dh
    q, k, v = _rand(B, Lq, d, dtype=dt, seed=1), _rand(B, Lk, d, dtype=dt, seed=2), _rand(B, Lk, d, dtype=dt, seed=3)
    kmask = torch.ones(B, Lk, dtype=torch.uint8, device=DEV)
    kmask[0, 70:90] = 0
    kmask[B - 1, Lk - 5:] = 0
    o = torch.empty(B * Lq, d, device=DEV, dtype=dt)
    lse = torch.empty(B * H * Lq, device=DEV)
    ops.attn_fwd(q.view(-1, d), k.view(-1, d), v.view(-1, d), o, lse, B, H, Lq, Lk, dh, kmask=kmask, causal=causal)
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    ref = _attn_ref(qr, kr, vr, kmask, causal, H)
    assert rel(o.view(B, Lq, d).float(), ref) < 2e-2
    do = _rand(B, Lq, d, dtype=dt, seed=4)
    ref.backward(do.float())
    dq, dk, dv = (torch.empty_like(t).view(-1, d) for t in (q, k, v))
    ops.attn_bwd(q.view(-1, d), k.view(-1, d), v.view(-1, d), o, lse, do.view(-1, d), dq, dk, dv, B, H, Lq, Lk, dh,
                 kmask=kmask, causal=causal)
    assert rel(dq.view_as(q).float(), qr.grad) < 3e-2
    assert rel(dk.view_as(k).float(), kr.grad) < 3e-2
    assert rel(dv.view_as(v).float(), vr.grad) < 3e-2


def test_blocked_tcgen05_attention_dropout_equals_streaming_kernel(monkeypatch):
    """Same dropout stream (global (i, j) indexing) in the blocked tcgen05 kernels and the streaming mma.sync kernels:
    outputs and gradients of the two implementations agree with dropout on, forward and backward use one mask."""
    B, H, Lq, Lk, dh, dt, p = 2, 8, 199, 199, 64, torch.bfloat16, 0.2
    d = H * dh
    q, k, v = (_rand(B * L, d, dtype=dt, seed=sd) for L, sd in ((Lq, 1), (Lk, 2), (Lk, 3)))
    do = _rand(B * Lq, d, dtype=dt, seed=4)
    kmask = torch.ones(B, Lk, dtype=torch.uint8, device=DEV)
    kmask[1, 150:] = 0
    res = {}
    for impl in ("tcgen05", "mma"):
        monkeypatch.setattr(ops, "ATTN_IMPL", impl)
        o = torch.empty(B * Lq, d, device=DEV, dtype=dt)
        lse = torch.empty(B * H * Lq, device=DEV)
        ops.attn_fwd(q, k, v, o, lse, B, H, Lq, Lk, dh, kmask=kmask, p_drop=p, seed=5, site=3)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        ops.attn_bwd(q, k, v, o, lse, do, dq, dk, dv, B, H, Lq, Lk, dh, kmask=kmask, p_drop=p, seed=5, site=3)
        res[impl] = (o.float(), lse.clone(), dq.float(), dk.float(), dv.float())
    for a, b in zip(res["tcgen05"], res["mma"]):
        assert rel(a, b) < 2e-2


@pytest.mark.parametrize("impl", ["tcgen05", "mma"])
@pytest.mark.parametrize("B,H,Lq,Lk,causal", [(5, 8, 36, 36, False), (3, 8, 64, 64, True), (3, 8, 64, 36, False),
                                              (2, 8, 57, 27, False), (2, 8, 100, 128, False), (3, 4, 128, 128, True),
                                              (1, 1, 20, 20, True)])
def test_bf16_dh64_attention_both_tensor_core_paths(impl, B, H, Lq, Lk, causal, monkeypatch):
    """tcgen05 single-tile kernels (packed pairs when L <= 64, incl. an odd number of problems) and the streaming
    mma.sync kernels on the shapes of this path."""
    monkeypatch.setattr(ops, "ATTN_IMPL", impl)
    dh, dt = 64, torch.bfloat16
    d = H * dh
    if causal and Lq != Lk:
        pytest.skip("causal is self-attention only")
    q, k, v = _rand(B, Lq, d, dtype=dt, seed=1), _rand(B, Lk, d, dtype=dt, seed=2), _rand(B, Lk, d, dtype=dt, seed=3)
    kmask = torch.ones(B, Lk, dtype=torch.uint8, device=DEV)
    kmask[0, 5:9] = 0
    kmask[B - 1, Lk - 3:] = 0
    o = torch.empty(B * Lq, d, device=DEV, dtype=dt)
    lse = torch.empty(B * H * Lq, device=DEV)
    ops.attn_fwd(q.view(-1, d), k.view(-1, d), v.view(-1, d), o, lse, B, H, Lq, Lk, dh, kmask=kmask, causal=causal)
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    ref = _attn_ref(qr, kr, vr, kmask, causal, H)
    assert rel(o.view(B, Lq, d).float(), ref) < 2e-2
    do = _rand(B, Lq, d, dtype=dt, seed=4)
    ref.backward(do.float())
    dq, dk, dv = (torch.empty_like(t).view(-1, d) for t in (q, k, v))
    ops.attn_bwd(q.view(-1, d), k.view(-1, d), v.view(-1, d), o, lse, do.view(-1, d), dq, dk, dv, B, H, Lq, Lk, dh,
                 kmask=kmask, causal=causal)
    assert rel(dq.view_as(q).float(), qr.grad) < 3e-2
    assert rel(dk.view_as(k).float(), kr.grad) < 3e-2
    assert rel(dv.view_as(v).float(), vr.grad) < 3e-2


def test_grouped_wgrad_with_fused_bias_grad():
    """One persistent launch for several dW += dy^T x products (+ db += colsum(dy)), incl. ragged sizes, pitched
    operand views and pre-existing gradient content (accumulation)."""
    dt = torch.bfloat16
    specs = [(4096, 1536, 512), (4096, 512, 512), (2304, 1024, 512), (4096, 2048, 512), (4096, 512, 2048),
             (1000, 200, 48), (777, 144, 48)]
    items, refs = [], []
    for i, (R, n_out, k_in) in enumerate(specs):
        wide = _rand(R, n_out + 64, dtype=dt, scale=R ** -0.5, seed=i)
        dy = wide[:, 32:32 + n_out] if n_out % 64 == 0 else wide[:, :n_out]  # a pitched view
        x = _rand(R, k_in, dtype=dt, seed=100 + i)
        dw0, db0 = _rand(n_out, k_in, seed=200 + i), _rand(n_out, seed=300 + i)
        dw, db = dw0.clone(), db0.clone()
        items.append((dy, x, dw, db, n_out, k_in, R))
        refs.append((dw0 + dy.float().T @ x.float(), db0 + dy.float().sum(0)))
    ops.wgrad_group(items)
    for it, (rw, rb) in zip(items, refs):
        assert rel(it[2], rw) < 3e-5, it[4:]
        assert rel(it[3], rb) < 3e-5, it[4:]


@pytest.mark.parametrize("specs", [
    [(9216, 1536, 512), (9216, 512, 512), (9216, 2048, 512), (9216, 512, 2048)],   # encoder layer: 48 tiles -> 3 k-segments
    [(16384, 200, 512)],                                                           # LM head: 2 tiles -> 8 k-segments
])
def test_grouped_wgrad_split_reduction(specs):
    """Few output tiles + a long reduction: the grouped kernel cuts every tile's row range into work items that add into
    the gradient with TMA reduce-add (bias gradients with atomics)."""
    dt = torch.bfloat16
    items, refs = [], []
    for i, (R, n_out, k_in) in enumerate(specs):
        dy = _rand(R, n_out, dtype=dt, scale=R ** -0.5, seed=i)
        x = _rand(R, k_in, dtype=dt, seed=100 + i)
        dw0, db0 = _rand(n_out, k_in, seed=200 + i), _rand(n_out, seed=300 + i)
        dw, db = dw0.clone(), db0.clone()
        items.append((dy, x, dw, db, n_out, k_in, R))
        refs.append((dw0 + dy.float().T @ x.float(), db0 + dy.float().sum(0)))
    ops.wgrad_group(items)
    for it, (rw, rb) in zip(items, refs):
        assert rel(it[2], rw) < 3e-5, it[4:]
        assert rel(it[3], rb) < 3e-5, it[4:]


@pytest.mark.parametrize("R", [1, 10, 30, 64, 65, 80, 320, 449, 512])
def test_small_linear_decode_products(R):
    """mma_small_linear (decode_small.cu): LayerNorm prologue + product + bias / GELU / gate / residual epilogue for a
    handful of rows, every kind against torch (bf16 operands, fp32 accumulate)."""
    F = torch.nn.functional
    d, f, V = 512, 2048, 200
    x = _rand(R, d, seed=1) * 2.0 + 0.3
    gamma, beta = _rand(d, seed=2) * 0.3 + 1.0, _rand(d, seed=3) * 0.1
    hn = F.layer_norm(x, (d,), gamma, beta, 1e-5).to(torch.bfloat16).float()
    # LayerNorm + QKV projection (bf16 out)
    W, b = _rand(3 * d, d, dtype=torch.bfloat16, scale=d ** -0.5, seed=4), _rand(3 * d, seed=5)
    out = torch.empty(R, 3 * d, device=DEV, dtype=torch.bfloat16)
    assert ops.small_linear(x, W, out, R, 3 * d, d, bias=b, gamma=gamma, beta=beta)
    assert rel(out.float(), hn @ W.float().T + b) < 1e-2
    # LayerNorm + FFN-1 with GELU, and the gated form
    W1, Wg = _rand(f, d, dtype=torch.bfloat16, scale=d ** -0.5, seed=6), _rand(f, d, dtype=torch.bfloat16, scale=d ** -0.5, seed=7)
    b1, bg = _rand(f, seed=8), _rand(f, seed=9)
    a = torch.empty(R, f, device=DEV, dtype=torch.bfloat16)
    assert ops.small_linear(x, W1, a, R, f, d, kind="gelu", bias=b1, gamma=gamma, beta=beta)
    assert rel(a.float(), F.gelu(hn @ W1.float().T + b1)) < 1e-2
    ag = torch.empty_like(a)
    assert ops.small_linear(x, W1, ag, R, f, d, kind="glu", bias=b1, gamma=gamma, beta=beta, w2=Wg, bias2=bg)
    assert rel(ag.float(), F.gelu(hn @ W1.float().T + b1) * (hn @ Wg.float().T + bg)) < 1e-2
    # bf16 input (no norm), K = 2048, residual epilogue, fp32 out
    W2, b2 = _rand(d, f, dtype=torch.bfloat16, scale=f ** -0.5, seed=10), _rand(d, seed=11)
    resid = _rand(R, d, seed=12)
    y = torch.empty(R, d, device=DEV)
    assert ops.small_linear(a, W2, y, R, d, f, kind="resid", bias=b2, resid=resid)
    assert rel(y, resid + a.float() @ W2.float().T + b2) < 2e-3
    # LM head: N = 200 (not a multiple of 16), fp32 out with a padded pitch
    Wv, bv = _rand(V, d, dtype=torch.bfloat16, scale=d ** -0.5, seed=13), _rand(V, seed=14)
    lg = torch.zeros(R, 208, device=DEV)
    assert ops.small_linear(x, Wv, lg, R, V, d, bias=bv, gamma=gamma, beta=beta)
    assert rel(lg[:, :V], hn @ Wv.float().T + bv) < 2e-3 and float(lg[:, V:].abs().max()) == 0.0
    # outside the envelope: more than 512 rows (8 blocks of 64), K not a multiple of 256
    big = _rand(513, d)
    assert not ops.small_linear(big, W, torch.empty(513, 3 * d, device=DEV, dtype=torch.bfloat16), 513, 3 * d, d, bias=b)
    assert not ops.small_linear(_rand(4, 200), _rand(16, 200, dtype=torch.bfloat16), torch.empty(4, 16, device=DEV), 4, 16, 200)
