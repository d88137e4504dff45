"""`torch.ops.mma_b200.*` (the C-ABI kernels as torch custom ops).  CPU: the namespace registers the
minimum export set K1-K8 with mutable-output schemas and refuses CPU tensors.  GPU: ops called through the dispatcher
against plain PyTorch fp32 references of the same operation."""
import math

import pytest
import torch

import multimodalanalytical_b200.torch_ops as T

gpu = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_namespace_registers_the_export_set():
    need = {"embed_fwd", "embed_bwd", "layernorm_fwd", "layernorm_bwd", "gemm_bias", "gemm_bias_gelu", "gemm_bias_glu",
            "gemm_bias_residual", "gemm_dgrad", "gemm_wgrad", "attn_fwd", "attn_bwd", "lmhead_ce_fwd", "lmhead_ce_bwd",
            "beam_step", "greedy_step", "decode_embed", "decode_self_attn", "adamw_clip_step"}
    assert need <= set(T.OPS)
    for name in T.OPS:
        op = getattr(torch.ops.mma_b200, name)
        schema = str(op.default._schema)
        assert schema.endswith("-> ()"), schema          # outputs are pre-allocated by the caller
        assert "!" in schema, schema                       # ... and declared as mutated arguments


def test_ops_have_no_cpu_kernel():
    x, y = torch.randn(4, 128), torch.empty(4, 128)
    with pytest.raises(NotImplementedError):
        torch.ops.mma_b200.layernorm_fwd(x, None, None, y, None)
    with pytest.raises(NotImplementedError):
        torch.ops.mma_b200.gemm_bias(torch.randn(8, 16), torch.randn(4, 16), None, torch.empty(8, 4))


@gpu
def test_gpu_layernorm_and_embed_ops():
    g = torch.Generator(device="cpu").manual_seed(0)
    rows, d, vocab = 300, 512, 40
    table = torch.randn(vocab, d, generator=g).to(DEV)
    ids = torch.randint(0, vocab, (rows,), generator=g).to(DEV)
    pre = torch.empty(rows, d, device=DEV)
    torch.ops.mma_b200.embed_fwd(ids, table, pre)
    assert torch.equal(pre, table[ids])
    gamma, beta = (torch.randn(d, generator=g) + 1).to(DEV), torch.randn(d, generator=g).to(DEV)
    y = torch.empty(rows, d, device=DEV)
    yb = torch.empty(rows, d, device=DEV, dtype=torch.bfloat16)
    torch.ops.mma_b200.layernorm_fwd(pre, gamma, beta, y, yb)
    xr = pre.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (d,), gamma, beta)
    assert rel(y, ref) < 1e-5 and rel(yb.float(), ref) < 1e-2
    dy = torch.randn(rows, d, generator=g).to(DEV)
    ref.backward(dy)
    dx = torch.empty(rows, d, device=DEV)
    dg, db = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    torch.ops.mma_b200.layernorm_bwd(dy, pre, gamma, dx, None, None, dg, db)
    assert rel(dx, xr.grad) < 2e-5 and rel(db, dy.sum(0)) < 1e-4
    dtable = torch.zeros(vocab, d, device=DEV)
    torch.ops.mma_b200.embed_bwd(ids, dx, dtable, -1)
    assert rel(dtable, torch.zeros(vocab, d, device=DEV).index_add_(0, ids, dx)) < 1e-5


@gpu
def test_gpu_linear_ops_forward_and_backward():
    g = torch.Generator(device="cpu").manual_seed(1)
    M, K, N = 1024, 512, 768
    a = (torch.randn(M, K, generator=g) * 0.5).to(DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.05).to(DEV).to(torch.bfloat16)
    bias = torch.randn(N, generator=g).to(DEV)
    af, wf = a.float(), w.float()
    z_ref = af @ wf.T + bias
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    torch.ops.mma_b200.gemm_bias(a, w, bias, out)
    assert rel(out.float(), z_ref) < 1e-2
    act, pre = torch.empty_like(out), torch.empty_like(out)
    torch.ops.mma_b200.gemm_bias_gelu(a, w, bias, act, pre)
    assert rel(pre.float(), z_ref) < 1e-2 and rel(act.float(), torch.nn.functional.gelu(z_ref)) < 1e-2
    resid = torch.randn(M, N, generator=g).to(DEV)
    x_new = torch.empty(M, N, device=DEV)
    torch.ops.mma_b200.gemm_bias_residual(a, w, bias, resid, x_new)
    assert rel(x_new, resid + z_ref) < 1e-2
    lin_pre = (torch.randn(M, N, generator=g)).to(DEV).to(torch.bfloat16)
    glu = torch.empty_like(out)
    torch.ops.mma_b200.gemm_bias_glu(a, w, bias, lin_pre, glu, None)
    assert rel(glu.float(), torch.nn.functional.gelu(lin_pre.float()) * z_ref) < 1.5e-2
    dy = (torch.randn(M, N, generator=g) * 0.1).to(DEV).to(torch.bfloat16)
    dx = torch.empty(M, K, device=DEV, dtype=torch.bfloat16)
    torch.ops.mma_b200.gemm_dgrad(dy, w, dx)
    assert rel(dx.float(), dy.float() @ wf) < 1e-2
    dw = torch.full((N, K), 0.25, device=DEV)
    dbias = torch.zeros(N, device=DEV)
    torch.ops.mma_b200.gemm_wgrad(dy, a, dw, dbias)
    assert rel(dw, 0.25 + dy.float().T @ af) < 1e-2 and rel(dbias, dy.float().sum(0)) < 1e-3


@gpu
def test_gpu_attention_and_ce_ops():
    g = torch.Generator(device="cpu").manual_seed(2)
    B, H, L, dh = 6, 8, 40, 64
    d = H * dh
    qkv = (torch.randn(B * L, 3 * d, generator=g) * 0.5).to(DEV).to(torch.bfloat16)
    q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    kmask = torch.ones(B, L, dtype=torch.uint8)
    kmask[1, 30:] = 0
    kmask = kmask.to(DEV)
    o = torch.empty(B * L, d, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B * H * L, device=DEV)
    torch.ops.mma_b200.attn_fwd(q, k, v, o, lse, B, H, L, L, dh, kmask, True)

    def heads(t):
        return t.float().view(B, L, H, dh).transpose(1, 2)

    s = heads(q) @ heads(k).transpose(-1, -2) / math.sqrt(dh)
    s = s.masked_fill(~kmask.bool()[:, None, None, :], float("-inf"))
    s = s + torch.triu(torch.full((L, L), float("-inf"), device=DEV), diagonal=1)
    ref = (torch.softmax(s, -1) @ heads(v)).transpose(1, 2).reshape(B * L, d)
    assert rel(o.float(), ref) < 2e-2
    rows, V = 500, 200
    logits = torch.randn(rows, 208, generator=g).to(DEV)[:, :V]
    labels = torch.randint(0, V, (rows,), generator=g)
    labels[::7] = -100
    labels = labels.to(DEV)
    row_loss, row_lse, stats = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV), torch.zeros(4, device=DEV)
    torch.ops.mma_b200.lmhead_ce_fwd(logits, labels, V, row_loss, row_lse, stats)
    lr = logits.clone().requires_grad_(True)
    want = torch.nn.functional.cross_entropy(lr, labels, ignore_index=-100)
    assert abs(float(stats[0]) - float(want)) < 1e-5 * abs(float(want))
    want.backward()
    dlogits = torch.zeros(rows, 208, device=DEV)
    torch.ops.mma_b200.lmhead_ce_bwd(logits, labels, V, row_lse, stats, dlogits[:, :V])
    assert rel(dlogits[:, :V], lr.grad) < 1e-5


@gpu
def test_gpu_beam_step_op_equals_the_direct_call():
    from multimodalanalytical_b200 import ops
    from multimodalanalytical_b200.decode import BeamState

    B, K, L, V = 5, 4, 16, 37
    g = torch.Generator(device="cpu").manual_seed(3)
    a, b = BeamState(B, K, L, 0, 2, 3, torch.device(DEV)), BeamState(B, K, L, 0, 2, 3, torch.device(DEV))
    for step in range(6):
        logits = torch.randn(B * K, 40, generator=g).to(DEV)
        ops.beam_step(logits, V, a)
        ops.advance(a.cur_len)
        torch.ops.mma_b200.beam_step(logits, V, B, K, L, 0, 3, b.cur_len, b.run_seq, b.fin_seq, b.run_score, b.fin_score,
                                     b.fin_flag, b.fin_len, b.improvable, b.all_hit, b.anc, b.next_tok, b.parent_row)
        ops.advance(b.cur_len)
    for name in ("run_seq", "fin_seq", "run_score", "fin_score", "fin_flag", "fin_len", "anc", "next_tok", "parent_row"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
