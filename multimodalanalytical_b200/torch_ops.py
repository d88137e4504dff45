"""`torch.ops.mma_b200.*`: the C-ABI kernels registered as torch custom ops (SURVEY.md §8b "C-ABI op layer",
minimum export set K1-K8).

Each op is a thin `torch.library.custom_op` over the matching `ops.*` wrapper, which calls the `extern "C"` entry point
of `lib/libmma_b200.so` with raw device pointers on the current stream.  Conventions of the boundary are kept: every
output is pre-allocated by the caller and listed in `mutates_args` (optional outputs are passed explicitly, `None` to
skip them), nothing allocates, ops return None, and only a CUDA implementation is registered - calling an op with CPU
tensors fails in the dispatcher instead of falling back.
`import multimodalanalytical_b200.torch_ops` registers the namespace; the engine itself calls `ops.*` directly (same
kernels, one Python frame less per launch).

  K1  embed_fwd / embed_bwd                         nn.Embedding (+ XVal scaling)           modeling/utils.py:93-106,154-160
  K2  layernorm_fwd / layernorm_bwd                 LayerNorm (+ pos-enc add, concat)       modeling/utils.py:165-180
  K3  gemm_bias / _gelu / _residual / _glu,         Linear + epilogues, dgrad, wgrad        custom_modeling.py:108-199
      gemm_dgrad, gemm_wgrad
  K4  attn_fwd / attn_bwd                           encoder / causal / cross attention      custom_modeling.py:131-199
  K5  lmhead_ce_fwd / lmhead_ce_bwd                 CrossEntropyLoss on the LM head         custom_modeling.py:486-491
  K6  beam_step / greedy_step                       transformers _beam_search / _sample     wrapper.py:443-451
  K7  decode_embed / decode_self_attn               KV-cached decoder step
  K8  adamw_clip_step                               clip_grad_norm_ + Adam/AdamW            wrapper.py:329-344
  +   collate_tokens / collate_target / patchify    device collator / patch preprocessor    data/datamodules.py:140-351
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch
from torch import Tensor
from torch.library import custom_op

from . import ops
from ._lib import EPI_ACCUM, EPI_GELU, EPI_GLU_MUL, EPI_RESID, EPI_STORE

NS = "mma_b200"
_CUDA = "cuda"


# ------------------------------------------------------------------------------------------------------ K1
@custom_op(f"{NS}::embed_fwd", mutates_args=("out",), device_types=_CUDA)
def embed_fwd(ids: Tensor, table: Tensor, out: Tensor, scale: Optional[Tensor] = None) -> None:
    ops.gather_rows(ids, table, out, scale=scale)


@custom_op(f"{NS}::embed_bwd", mutates_args=("dtable",), device_types=_CUDA)
def embed_bwd(ids: Tensor, grad: Tensor, dtable: Tensor, pad_idx: int, scale: Optional[Tensor] = None) -> None:
    ops.scatter_add_rows(ids, grad, dtable, pad_idx, scale=scale)


# ------------------------------------------------------------------------------------------------------ K2
@custom_op(f"{NS}::layernorm_fwd", mutates_args=("y", "y2"), device_types=_CUDA)
def layernorm_fwd(x: Tensor, gamma: Optional[Tensor], beta: Optional[Tensor], y: Tensor, y2: Optional[Tensor],
                  add: Optional[Tensor] = None, group: int = 0, out_group_stride: int = 0, out_offset: int = 0,
                  eps: float = 1e-5) -> None:
    ops.ln_fwd(x, gamma, beta, y, y2=y2, add=add, group=group, out_group_stride=out_group_stride,
               out_offset=out_offset, eps=eps)


@custom_op(f"{NS}::layernorm_bwd", mutates_args=("dx", "dxb", "dgamma", "dbeta"), device_types=_CUDA)
def layernorm_bwd(dy: Tensor, x: Tensor, gamma: Optional[Tensor], dx: Optional[Tensor], dres: Optional[Tensor],
                  dxb: Optional[Tensor], dgamma: Optional[Tensor], dbeta: Optional[Tensor],
                  p_drop: float = 0.0, seed: int = 0, site: int = 0, group: int = 0, in_group_stride: int = 0,
                  in_offset: int = 0, eps: float = 1e-5) -> None:
    ops.ln_bwd(dy, x, gamma, dx=dx, dres=dres, dxb=dxb, dgamma=dgamma, dbeta=dbeta, p_drop=p_drop, seed=seed, site=site,
               group=group, in_group_stride=in_group_stride, in_offset=in_offset, eps=eps)


# ------------------------------------------------------------------------------------------------------ K3
def _mnk(a: Tensor, w: Tensor):
    return a.shape[0], w.shape[0], a.shape[1]


@custom_op(f"{NS}::gemm_bias", mutates_args=("out",), device_types=_CUDA)
def gemm_bias(a: Tensor, w: Tensor, bias: Optional[Tensor], out: Tensor) -> None:
    """out[M, N] = a[M, K] @ w[N, K]^T + bias."""
    M, N, K = _mnk(a, w)
    ops.gemm(a, w, M, N, K, ops.make_epi(EPI_STORE, out, bias=bias))


@custom_op(f"{NS}::gemm_bias_gelu", mutates_args=("out", "pre_act"), device_types=_CUDA)
def gemm_bias_gelu(a: Tensor, w: Tensor, bias: Optional[Tensor], out: Tensor, pre_act: Optional[Tensor],
                   p_drop: float = 0.0, seed: int = 0, site: int = 0) -> None:
    """z = a w^T + bias; pre_act = z (saved for backward); out = dropout(gelu(z))."""
    M, N, K = _mnk(a, w)
    ops.gemm(a, w, M, N, K, ops.make_epi(EPI_GELU, out, out2=pre_act, bias=bias, p_drop=p_drop, seed=seed, site=site))


@custom_op(f"{NS}::gemm_bias_residual", mutates_args=("out",), device_types=_CUDA)
def gemm_bias_residual(a: Tensor, w: Tensor, bias: Optional[Tensor], resid: Tensor, out: Tensor, p_drop: float = 0.0,
                       seed: int = 0, site: int = 0) -> None:
    """out = resid + dropout(a w^T + bias)."""
    M, N, K = _mnk(a, w)
    ops.gemm(a, w, M, N, K, ops.make_epi(EPI_RESID, out, bias=bias, resid=resid, p_drop=p_drop, seed=seed, site=site))


@custom_op(f"{NS}::gemm_bias_glu", mutates_args=("out", "gate_pre"), device_types=_CUDA)
def gemm_bias_glu(a: Tensor, w_gate: Tensor, bias: Optional[Tensor], lin_pre: Tensor, out: Tensor,
                  gate_pre: Optional[Tensor], p_drop: float = 0.0, seed: int = 0, site: int = 0) -> None:
    """GLU feed-forward (custom_modeling.py:143-150): z2 = a w_gate^T + bias; out = dropout(gelu(lin_pre) * z2)."""
    M, N, K = _mnk(a, w_gate)
    ops.gemm(a, w_gate, M, N, K, ops.make_epi(EPI_GLU_MUL, out, out2=gate_pre, bias=bias, aux=lin_pre, p_drop=p_drop,
                                              seed=seed, site=site))


@custom_op(f"{NS}::gemm_dgrad", mutates_args=("dx",), device_types=_CUDA)
def gemm_dgrad(dy: Tensor, w: Tensor, dx: Tensor) -> None:
    """dx[M, K_in] = dy[M, N_out] @ w[N_out, K_in]."""
    M, n_out, k_in = dy.shape[0], w.shape[0], w.shape[1]
    ops.gemm(dy, w, M, k_in, n_out, ops.make_epi(EPI_STORE, dx), b_mn=True)


@custom_op(f"{NS}::gemm_wgrad", mutates_args=("dw", "db"), device_types=_CUDA)
def gemm_wgrad(dy: Tensor, x: Tensor, dw: Tensor, db: Optional[Tensor]) -> None:
    """dw[N_out, K_in] += dy[R, N_out]^T @ x[R, K_in]; db[N_out] += column sums of dy."""
    R, n_out, k_in = dy.shape[0], dy.shape[1], x.shape[1]
    ops.gemm(dy, x, n_out, k_in, R, ops.make_epi(EPI_ACCUM, dw, accumulate=1), a_mn=True, b_mn=True)
    if db is not None:
        ops.colsum(dy, db, rows=R, cols=n_out)


# ------------------------------------------------------------------------------------------------------ K4
@custom_op(f"{NS}::attn_fwd", mutates_args=("o", "lse"), device_types=_CUDA)
def attn_fwd(q: Tensor, k: Tensor, v: Tensor, o: Tensor, lse: Optional[Tensor], B: int, H: int, Lq: int, Lk: int,
             dh: int, kmask: Optional[Tensor] = None, causal: bool = False, p_drop: float = 0.0, seed: int = 0,
             site: int = 0) -> None:
    ops.attn_fwd(q, k, v, o, lse, B, H, Lq, Lk, dh, kmask=kmask, causal=causal, p_drop=p_drop, seed=seed, site=site)


@custom_op(f"{NS}::attn_bwd", mutates_args=("dq", "dk", "dv"), device_types=_CUDA)
def attn_bwd(q: Tensor, k: Tensor, v: Tensor, o: Tensor, lse: Tensor, dout: Tensor, dq: Tensor, dk: Tensor, dv: Tensor,
             B: int, H: int, Lq: int, Lk: int, dh: int, kmask: Optional[Tensor] = None, causal: bool = False,
             p_drop: float = 0.0, seed: int = 0, site: int = 0) -> None:
    ops.attn_bwd(q, k, v, o, lse, dout, dq, dk, dv, B, H, Lq, Lk, dh, kmask=kmask, causal=causal, p_drop=p_drop,
                 seed=seed, site=site)


# ------------------------------------------------------------------------------------------------------ K5
@custom_op(f"{NS}::lmhead_ce_fwd", mutates_args=("row_loss", "row_lse", "stats"), device_types=_CUDA)
def lmhead_ce_fwd(logits: Tensor, labels: Tensor, vocab: int, row_loss: Tensor, row_lse: Tensor, stats: Tensor,
                  smoothing: float = 0.0, ignore_index: int = -100) -> None:
    """stats[0] = mean CE over the non-ignored rows (+ label smoothing), stats[1] = their count."""
    ops.ce_fwd(logits, labels, vocab, row_loss, row_lse, stats, smoothing=smoothing, ignore_index=ignore_index)


@custom_op(f"{NS}::lmhead_ce_bwd", mutates_args=("dlogits",), device_types=_CUDA)
def lmhead_ce_bwd(logits: Tensor, labels: Tensor, vocab: int, row_lse: Tensor, stats: Tensor, dlogits: Tensor,
                  gscale: float = 1.0, smoothing: float = 0.0, ignore_index: int = -100) -> None:
    ops.ce_bwd(logits, labels, vocab, row_lse, stats, dlogits, gscale=gscale, smoothing=smoothing,
               ignore_index=ignore_index)


# ------------------------------------------------------------------------------------------------------ K6
@custom_op(f"{NS}::beam_step", mutates_args=("run_seq", "fin_seq", "run_score", "fin_score", "fin_flag", "fin_len",
                                            "improvable", "all_hit", "anc", "next_tok", "parent_row"),
           device_types=_CUDA)
def beam_step(logits: Tensor, vocab: int, n_spectra: int, n_beams: int, max_length: int, pad_id: int, eos_id: int,
              cur_len: Tensor, run_seq: Tensor, fin_seq: Tensor, run_score: Tensor, fin_score: Tensor, fin_flag: Tensor,
              fin_len: Tensor, improvable: Tensor, all_hit: Tensor, anc: Tensor, next_tok: Tensor, parent_row: Tensor,
              extra_bias: Optional[Tensor] = None, prenorm: bool = False) -> None:
    st = SimpleNamespace(B=n_spectra, K=n_beams, L=max_length, pad_id=pad_id, eos_id=eos_id, cur_len=cur_len,
                         run_seq=run_seq, fin_seq=fin_seq, run_score=run_score, fin_score=fin_score, fin_flag=fin_flag,
                         fin_len=fin_len, improvable=improvable, all_hit=all_hit, anc=anc, next_tok=next_tok,
                         parent_row=parent_row)
    ops.beam_step(logits, vocab, st, extra_bias, prenorm=prenorm)


@custom_op(f"{NS}::greedy_step", mutates_args=("seq", "unfinished", "next_tok"), device_types=_CUDA)
def greedy_step(logits: Tensor, vocab: int, rows: int, max_length: int, pad_id: int, eos_id: int, cur_len: Tensor,
                seq: Tensor, unfinished: Tensor, next_tok: Tensor, extra_bias: Optional[Tensor] = None,
                prenorm: bool = False) -> None:
    st = SimpleNamespace(B=rows, K=1, L=max_length, pad_id=pad_id, eos_id=eos_id, cur_len=cur_len, run_seq=seq,
                         unfinished=unfinished, next_tok=next_tok)
    ops.greedy_step(logits, vocab, st, extra_bias, prenorm=prenorm)


# ------------------------------------------------------------------------------------------------------ K7
@custom_op(f"{NS}::decode_embed", mutates_args=("out",), device_types=_CUDA)
def decode_embed(tok: Tensor, table: Tensor, gamma: Optional[Tensor], beta: Optional[Tensor], pos: Tensor,
                 cur_len: Tensor, out: Tensor, eps: float = 1e-5) -> None:
    ops.decode_embed(tok, table, gamma, beta, pos, cur_len, out, eps=eps)


@custom_op(f"{NS}::decode_self_attn", mutates_args=("kcache", "vcache", "o"), device_types=_CUDA)
def decode_self_attn(q: Tensor, knew: Tensor, vnew: Tensor, kcache: Tensor, vcache: Tensor, anc: Optional[Tensor],
                     cur_len: Tensor, o: Tensor, rows: int, heads: int, dh: int, max_length: int) -> None:
    ops.decode_self_attn(q, knew, vnew, kcache, vcache, anc, cur_len, o, rows, heads, dh, max_length)


# ------------------------------------------------------------------------------------------------------ K8
@custom_op(f"{NS}::adamw_clip_step", mutates_args=("p", "g", "m", "v", "p_bf16", "norm", "workspace"),
           device_types=_CUDA)
def adamw_clip_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, p_bf16: Optional[Tensor], hyper: Tensor, norm: Tensor,
                    workspace: Tensor, decoupled: bool = True, zero_grad: bool = True) -> None:
    """norm = ||g||; then clip + Adam / AdamW + bf16 mirror refresh + gradient zeroing in one pass
    (hyper: lr, beta1, beta2, eps, weight_decay, bias corrections, clip, 1/world - trainer.FusedTrainer._write_hyper)."""
    ops.grad_norm(g, workspace, norm)
    ops.adam_step(p, g, m, v, p_bf16, hyper, norm=norm, decoupled=decoupled, zero_grad=zero_grad)


# -------------------------------------------------------------------------------------------- data pipeline
@custom_op(f"{NS}::collate_tokens", mutates_args=("ids", "mask"), device_types=_CUDA)
def collate_tokens(flat: Tensor, offsets: Tensor, row_valid: Optional[Tensor], rows: Tensor, pad_id: int, max_len: int,
                   ids: Tensor, mask: Optional[Tensor]) -> None:
    ops.collate_tokens(flat, offsets, row_valid, rows, pad_id, max_len, ids, mask)


@custom_op(f"{NS}::collate_target", mutates_args=("dec_in", "dec_mask", "labels"), device_types=_CUDA)
def collate_target(flat: Tensor, offsets: Tensor, rows: Tensor, pad_id: int, max_len: int, dec_in: Tensor,
                   dec_mask: Tensor, labels: Tensor) -> None:
    ops.collate_target(flat, offsets, rows, pad_id, max_len, dec_in, dec_mask, labels)


@custom_op(f"{NS}::patchify", mutates_args=("out", "pad"), device_types=_CUDA)
def patchify(raw: Tensor, out: Tensor, pad: Optional[Tensor], mean: float, std: float, offset: int = 0, hop: int = 0,
             missing: Optional[Tensor] = None, masking: bool = False, rows: Optional[Tensor] = None) -> None:
    ops.patchify(raw, out, mean, std, offset=offset, hop=hop if hop > 0 else None, pad=pad, missing=missing,
                 masking=masking, rows=rows)


OPS = ("embed_fwd", "embed_bwd", "layernorm_fwd", "layernorm_bwd", "gemm_bias", "gemm_bias_gelu", "gemm_bias_residual",
       "gemm_bias_glu", "gemm_dgrad", "gemm_wgrad", "attn_fwd", "attn_bwd", "lmhead_ce_fwd", "lmhead_ce_bwd", "beam_step",
       "greedy_step", "decode_embed", "decode_self_attn", "adamw_clip_step", "collate_tokens", "collate_target",
       "patchify")
