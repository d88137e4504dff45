"""KV-cached greedy / beam-search generation.

Reference path: `HFWrapper.generate` (wrapper.py:409-453) -> transformers `generate(num_beams=K,
num_return_sequences=K, use_cache=False)`; the semantics of `_sample` / `_beam_search` (length penalty 1.0,
early_stopping False, ForcedEOS at max_length 128, `pad_token_id or eos_token_id` fill) are restated in
`oracle/spectra_oracle.py` and implemented on the device in csrc/decode.cu.

Differences in HOW (not WHAT): the encoder runs once per spectrum; cross-attention K/V are projected once per
spectrum and shared by its K beams; self-attention K/V are cached and beams are re-ordered through an
ancestor table instead of copying the cache; the per-step schedule reads `cur_len` from device memory so it
is captured once in a CUDA graph and replayed for all 127 steps.
"""
from __future__ import annotations

import os
from typing import Any, Dict, Optional

import torch

from . import ops
from ._lib import EPI_GELU, EPI_GLU_MUL, EPI_RESID, EPI_STORE
from .model import Engine


# Residual products of the decode step fused with the LayerNorm that follows them (mma_gemm2_resid_ln): opt-in
# (MMA_DECODE_FUSE_LN=1 every residual product, 2 only the K = d_model ones, 3 only up to 4736 rows, where the 4-CTA-cluster
# kernel applies).  Measured on B200, beam-10, ms / step fused vs separate launches: CTA-pair kernel at 2560 rows 1.75 vs
# 1.58 (a pair owns 256 rows x all 512 columns, so only 10 pairs work); 4-CTA-cluster kernel (128 x 128 tiles, row
# statistics exchanged through distributed shared memory) 1.003 vs 0.953 at 2560 rows, 0.654 vs 0.590 at 320 rows - the
# cluster barrier in the middle of the epilogue and the row-per-thread stores of h cost more than the ~3 us LayerNorm launch.
FUSE_DECODE_LN = int(os.environ.get("MMA_DECODE_FUSE_LN", "0"))


# finished-spectrum compaction inside a decode batch (MMA_DECODE_COMPACT=0 disables it); batches below COMPACT_MIN_B
# spectra are latency-bound and not worth a re-capture
COMPACT = os.environ.get("MMA_DECODE_COMPACT", "1") != "0"
COMPACT_MIN_B = int(os.environ.get("MMA_DECODE_COMPACT_MIN_B", "16"))
COMPACT_LIVE_FRAC = 0.5  # compact when at most this fraction of the batch is still searching


SMALL_DECODE = os.environ.get("MMA_DECODE_SMALL", "1") != "0"  # few rows: fused LayerNorm + product launches (decode_small.cu)
SMALL_MAX_ROWS = int(os.environ.get("MMA_DECODE_SMALL_ROWS", "64"))  # rows (spectra x beams) up to which that path is taken


# few spectra (every cluster resident at once, <= 16 rows per cluster): the whole step as ONE cluster-synchronised launch
# (decode_step.cu); MMA_DECODE_PERSIST=0 keeps the per-op launches
PERSIST_DECODE = os.environ.get("MMA_DECODE_PERSIST", "1") != "0"
PERSIST_CLUSTER = int(os.environ.get("MMA_DECODE_PERSIST_CLUSTER", "16"))
PERSIST_MAX_CLUSTERS = int(os.environ.get("MMA_DECODE_PERSIST_MAX_CLUSTERS", "0"))  # 0: what the device can hold at once
PERSIST_MIN_ROWS = int(os.environ.get("MMA_DECODE_PERSIST_MIN_ROWS", "0"))  # 0: no lower bound (tests force the path)


class BeamState:
    """Device-resident search state for B spectra x K beams (K == 1: greedy)."""

    def __init__(self, B, K, L, pad_id, bos_id, eos_id, device):
        self.B, self.K, self.L = B, K, L
        self.pad_id, self.bos_id, self.eos_id = pad_id, bos_id, eos_id
        dev = device
        R = B * K
        i32, f32, u8 = torch.int32, torch.float32, torch.uint8
        self.cur_len = torch.empty(1, dtype=i32, device=dev)
        self.next_tok = torch.empty(R, dtype=i32, device=dev)
        self.parent_row = torch.empty(R, dtype=i32, device=dev)
        if K == 1:
            self.run_seq = torch.empty(B, L, dtype=i32, device=dev)
            self.unfinished = torch.empty(B, dtype=u8, device=dev)
            self.anc = None
        else:
            self.run_seq = torch.empty(2, B, K, L, dtype=i32, device=dev)
            self.fin_seq = torch.empty(2, B, K, L, dtype=i32, device=dev)
            self.run_score = torch.empty(B, K, dtype=f32, device=dev)
            self.fin_score = torch.empty(B, K, dtype=f32, device=dev)
            self.fin_flag = torch.empty(B, K, dtype=u8, device=dev)
            self.fin_len = torch.empty(B, K, dtype=i32, device=dev)
            self.improvable = torch.empty(B, dtype=u8, device=dev)
            self.all_hit = torch.empty(B, dtype=u8, device=dev)
            self.anc = torch.empty(2, R, L, dtype=i32, device=dev)
        self.reset()

    def reset(self):
        """(Re)initialise in place, so a captured CUDA graph over these buffers stays valid."""
        self.cur_len.fill_(1)
        self.next_tok.fill_(self.bos_id)
        self.parent_row.copy_(torch.arange(self.B * self.K, dtype=torch.int32, device=self.cur_len.device))
        if self.K == 1:
            self.run_seq.fill_(self.pad_id)
            self.run_seq[:, 0] = self.bos_id
            self.unfinished.fill_(1)
            return
        fill = self.pad_id if self.pad_id else self.eos_id  # transformers: `pad_token_id or eos_token_id`
        self.run_seq.fill_(fill)
        self.run_seq[:, :, :, 0] = self.bos_id
        self.fin_seq.copy_(self.run_seq)
        self.run_score.zero_()
        self.run_score[:, 1:] = -1.0e9
        self.fin_score.fill_(-1.0e9)
        self.fin_flag.zero_()
        self.fin_len.zero_()
        self.improvable.fill_(1)
        self.all_hit.zero_()
        self.anc.zero_()


class Generator:
    def __init__(self, engine: Engine):
        self.eng = engine
        self._graphs: Dict[Any, Any] = {}
        self._states: Dict[Any, BeamState] = {}
        engine.on_release.append(self._graphs.clear)

    # one decoder step for R rows; every shape is static and cur_len lives on the device
    def _step(self, st: BeamState, ctx: Dict[str, Any], extra_bias=None):
        logits = self._forward_logits(st, ctx)
        self._select(st, logits, extra_bias)

    def _select(self, st: BeamState, scores, extra_bias=None, prenorm=False, guide=None):
        V = self.eng.cfg.vocab_size
        if st.K == 1:
            ops.greedy_step(scores, V, st, extra_bias, prenorm=prenorm, guide=guide)
        else:
            ops.beam_step(scores, V, st, extra_bias, prenorm=prenorm, guide=guide)
        ops.advance(st.cur_len)

    def _forward_logits_small(self, st: BeamState, ctx: Dict[str, Any]):
        """The decoder step for <= 64 rows (a few spectra x beams), bf16: every LayerNorm + product (+ GELU / gate /
        residual) is ONE `mma_small_linear` launch (52 launches per step instead of 74, none of them a tcgen05 tile that
        is > 90 % padding).  Same math as `_forward_logits`."""
        eng, cfg = self.eng, self.eng.cfg
        d, H = cfg.d_model, cfg.decoder_attention_heads
        dh = d // H
        R, K, L = st.B * st.K, st.K, st.L
        T, f, e, tm = eng.adt, cfg.decoder_ffn_dim, eng.ps.EMB, cfg.target_modality
        gam = bet = None
        if cfg.multimodal_norm:
            gam, bet = eng.P(f"{e}embedding_norm_dict.{tm}.weight"), eng.P(f"{e}embedding_norm_dict.{tm}.bias")
        x = eng.buf("g.x", (R, d), torch.float32)
        ops.decode_embed(st.next_tok, eng.P(f"{e}embedding_layer_dict.{tm}.weight"), gam, bet, ctx["pos"], st.cur_len, x)
        qkv = eng.buf("g.qkv", (R, 3 * d), T)
        att = eng.buf("g.att", (R, d), T)
        q = eng.buf("g.q", (R, d), T)
        a = eng.buf("g.a", (R, f), T)
        xa = eng.buf("g.xa", (R, d), torch.float32)
        xb = eng.buf("g.xb", (R, d), torch.float32)
        P, W = eng.P, eng.W

        def lin(xin, wname, bname, out, n_out, k_in, kind="store", norm=None, resid=None, rows=None, w2=None, b2=None):
            w = W(wname) if rows is None else W(wname)[rows]
            b = P(bname) if rows is None else P(bname)[rows]
            ok = ops.small_linear(xin, w, out, R, n_out, k_in, kind=kind, bias=b,
                                  gamma=P(norm + "weight") if norm else None, beta=P(norm + "bias") if norm else None,
                                  w2=W(w2) if w2 else None, bias2=P(b2) if b2 else None, resid=resid)
            assert ok, "small_linear declined a shape the caller checked"

        for i in range(cfg.decoder_layers):
            p = f"hf_model.decoder.layers.{i}."
            lin(x, p + "self_attn.in_proj_weight", p + "self_attn.in_proj_bias", qkv, 3 * d, d, norm=p + "norm1.")
            ops.decode_self_attn(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], ctx["kc"][i], ctx["vc"][i], st.anc,
                                 st.cur_len, att, R, H, dh, L, beams=K)
            lin(att, p + "self_attn.out_proj.weight", p + "self_attn.out_proj.bias", xa, d, d, kind="resid", resid=x)
            lin(xa, p + "multihead_attn.in_proj_weight", p + "multihead_attn.in_proj_bias", q, d, d, norm=p + "norm2.",
                rows=slice(0, d))
            kv = ctx["kvmem"][i]
            ops.attn_fwd(q, kv[:, :d], kv[:, d:], att, None, st.B, H, K, ctx["S"], dh, kmask=ctx["enc_mask"])
            lin(att, p + "multihead_attn.out_proj.weight", p + "multihead_attn.out_proj.bias", xb, d, d, kind="resid",
                resid=xa)
            if cfg.gated_linear:
                lin(xb, p + "linear1.weight", p + "linear1.bias", a, f, d, kind="glu", norm=p + "norm3.",
                    w2=p + "gate.weight", b2=p + "gate.bias")
            else:
                lin(xb, p + "linear1.weight", p + "linear1.bias", a, f, d, kind="gelu", norm=p + "norm3.")
            lin(a, p + "linear2.weight", p + "linear2.bias", x, d, f, kind="resid", resid=xb)
        logits = eng.buf("g.logits", (R, eng.ldv), torch.float32)
        lin(x, "hf_model.token_ff.weight", "hf_model.token_ff.bias", logits, cfg.vocab_size, d,
            norm="hf_model.decoder.norm.")
        return logits

    def _persist_plan(self, st: BeamState):
        """(rows per cluster, clusters) of the one-launch step, or None outside its envelope: bf16, d 512 / ffn 2048 /
        8 heads of 64, <= 16 beams, and few enough spectra that every cluster is resident at once."""
        cfg, eng = self.eng.cfg, self.eng
        if not (PERSIST_DECODE and eng.precision == "bf16" and cfg.d_model == 512 and cfg.decoder_ffn_dim == 2048
                and cfg.decoder_attention_heads == 8 and st.K <= 16 and cfg.decoder_layers <= 12):
            return None
        global PERSIST_MAX_CLUSTERS
        if PERSIST_MAX_CLUSTERS == 0:
            # -1: the device cannot hold a 16-CTA cluster of this kernel (a partitioned GPU): per-op launches only
            PERSIST_MAX_CLUSTERS = ops.decode_step_max_clusters(PERSIST_CLUSTER) or -1
        if PERSIST_MAX_CLUSTERS < 0:
            return None
        per = max(1, min(16 // st.K, -(-st.B // PERSIST_MAX_CLUSTERS)))  # spectra per cluster
        clusters = -(-st.B // per)
        if clusters > PERSIST_MAX_CLUSTERS:
            return None
        # measured on B200 (C5 model, ms per step, one-launch vs per-op launches): beam-10 x 1 / 2 / 4 / 7 spectra
        # 0.327 / 0.327 / 0.333 / 0.332 vs 0.344 / 0.388 / 0.497 / 0.532; greedy x 1 / 7 / 16 spectra 0.212 / 0.219 / 0.236 vs
        # 0.250 / 0.240 / 0.290
        if st.B * st.K < PERSIST_MIN_ROWS:
            return None
        return per * st.K, clusters

    def _forward_logits_persist(self, st: BeamState, ctx: Dict[str, Any], plan):
        """One `mma_decode_step` launch; same buffers, same math as `_forward_logits_small`."""
        from ._lib import DecodeStep
        eng, cfg = self.eng, self.eng.cfg
        d, H, f = cfg.d_model, cfg.decoder_attention_heads, cfg.decoder_ffn_dim
        R, K, L = st.B * st.K, st.K, st.L
        T, e, tm = eng.adt, eng.ps.EMB, cfg.target_modality
        logits = eng.buf("g.logits", (R, eng.ldv), torch.float32)
        key = ("persist", R, K, L, ctx["S"], ctx["kc"].data_ptr(), ctx["kvmem"].data_ptr(), st.next_tok.data_ptr())
        args = self._persist_args.get(key) if hasattr(self, "_persist_args") else None
        if args is None:
            if not hasattr(self, "_persist_args"):
                self._persist_args = {}
                eng.on_release.append(self._persist_args.clear)
            P, W = eng.P, eng.W
            a = DecodeStep()

            def w(name, rows=None):
                t = W(name) if rows is None else W(name)[rows]
                assert t.dtype == torch.bfloat16 and t.stride(1) == 1 and t.stride(0) == t.shape[1], name
                return t.data_ptr()

            def p_(name, rows=None):
                t = P(name) if rows is None else P(name)[rows]
                assert t.dtype == torch.float32 and t.is_contiguous(), name
                return t.data_ptr()

            for i in range(cfg.decoder_layers):
                p = f"hf_model.decoder.layers.{i}."
                ly = a.layer[i]
                ly.w_qkv, ly.b_qkv = w(p + "self_attn.in_proj_weight"), p_(p + "self_attn.in_proj_bias")
                ly.w_so, ly.b_so = w(p + "self_attn.out_proj.weight"), p_(p + "self_attn.out_proj.bias")
                ly.w_cq = w(p + "multihead_attn.in_proj_weight", slice(0, d))
                ly.b_cq = p_(p + "multihead_attn.in_proj_bias", slice(0, d))
                ly.w_co, ly.b_co = w(p + "multihead_attn.out_proj.weight"), p_(p + "multihead_attn.out_proj.bias")
                ly.w_f1, ly.b_f1 = w(p + "linear1.weight"), p_(p + "linear1.bias")
                if cfg.gated_linear:
                    ly.w_fg, ly.b_fg = w(p + "gate.weight"), p_(p + "gate.bias")
                ly.w_f2, ly.b_f2 = w(p + "linear2.weight"), p_(p + "linear2.bias")
                for j in (1, 2, 3):
                    setattr(ly, f"n{j}g", p_(p + f"norm{j}.weight"))
                    setattr(ly, f"n{j}b", p_(p + f"norm{j}.bias"))
                ly.kc, ly.vc, ly.kvmem = ctx["kc"][i].data_ptr(), ctx["vc"][i].data_ptr(), ctx["kvmem"][i].data_ptr()
            a.tok, a.emb = st.next_tok.data_ptr(), p_(f"{e}embedding_layer_dict.{tm}.weight")
            if cfg.multimodal_norm:
                a.emb_g, a.emb_b = p_(f"{e}embedding_norm_dict.{tm}.weight"), p_(f"{e}embedding_norm_dict.{tm}.bias")
            a.pos, a.cur_len = ctx["pos"].data_ptr(), st.cur_len.data_ptr()
            a.fin_g, a.fin_b = p_("hf_model.decoder.norm.weight"), p_("hf_model.decoder.norm.bias")
            a.w_lm, a.b_lm = w("hf_model.token_ff.weight"), p_("hf_model.token_ff.bias")
            for name, shape, dt in (("x", (R, d), torch.float32), ("xa", (R, d), torch.float32),
                                    ("xb", (R, d), torch.float32), ("qkv", (R, 3 * d), T), ("att", (R, d), T),
                                    ("q", (R, d), T), ("a", (R, f), T)):
                setattr(a, name, eng.buf("g." + name, shape, dt).data_ptr())
            a.logits = logits.data_ptr()
            a.anc = st.anc.data_ptr() if st.anc is not None else None
            a.enc_mask = ctx["enc_mask"].data_ptr()
            if os.environ.get("MMA_DECODE_PERSIST_DBG"):  # per-phase %globaltimer stamps of cluster 0 (scripts/decode_step_phases.py)
                self.dbg_times = torch.zeros(192, dtype=torch.int64, device=eng.dev)
                a.dbg_times = self.dbg_times.data_ptr()
            a.ldv = eng.ldv
            a.layers, a.R, a.rows_per_cluster, a.beams = cfg.decoder_layers, R, plan[0], K
            a.d, a.f, a.H, a.Lmax, a.S, a.V, a.gated = d, f, H, L, ctx["S"], cfg.vocab_size, int(bool(cfg.gated_linear))
            a.eps, a.scale = 1e-5, (d // H) ** -0.5
            assert ctx["pos"].dtype == torch.float32 and ctx["pos"].stride(0) == d
            args = self._persist_args[key] = a
        ok = ops.decode_step(args, PERSIST_CLUSTER)
        assert ok, "mma_decode_step declined a shape the caller checked"
        return logits

    def _small_ok(self, R):
        cfg = self.eng.cfg
        return (SMALL_DECODE and self.eng.precision == "bf16" and R <= min(SMALL_MAX_ROWS, 512) and cfg.d_model % 256 == 0
                and cfg.decoder_ffn_dim % 256 == 0 and (cfg.d_model // cfg.decoder_attention_heads) == 64)

    def _forward_logits_post(self, st: BeamState, ctx: Dict[str, Any]):
        """The decoder step of a `post_layer_normalisation=False` model, LN(x + f(x)) layers (custom_modeling.py:166-176
        with torch's norm_first=False): per-op launches, each LayerNorm writing the fp32 stream and the next product's
        operand at once.  No shipped config uses it, so the fused small-row / one-launch steps do not cover it."""
        eng, cfg = self.eng, self.eng.cfg
        d, H, f = cfg.d_model, cfg.decoder_attention_heads, cfg.decoder_ffn_dim
        dh = d // H
        R, K, L = st.B * st.K, st.K, st.L
        T, e, tm = eng.adt, eng.ps.EMB, cfg.target_modality
        gam = bet = None
        if cfg.multimodal_norm:
            gam, bet = eng.P(f"{e}embedding_norm_dict.{tm}.weight"), eng.P(f"{e}embedding_norm_dict.{tm}.bias")
        x = eng.buf("g.x", (R, d), torch.float32)
        ops.decode_embed(st.next_tok, eng.P(f"{e}embedding_layer_dict.{tm}.weight"), gam, bet, ctx["pos"], st.cur_len, x)
        bf = eng.precision == "bf16"
        h = eng.buf("g.h", (R, d), T) if bf else x  # the stream in the activation dtype
        if bf:
            ops.cast_f32_bf16(x, h)
        qkv = eng.buf("g.qkv", (R, 3 * d), T)
        att = eng.buf("g.att", (R, d), T)
        q = eng.buf("g.q", (R, d), T)
        a = eng.buf("g.a", (R, f), T)
        z = eng.buf("g.z", (R, f), T)
        xs = eng.buf("g.xa", (R, d), torch.float32)
        P, W = eng.P, eng.W

        def resid_norm(a_in, wname, bname, k_in, norm):
            """x, h = LayerNorm(x + a_in W^T + b)"""
            ops.gemm(a_in, W(wname), R, d, k_in, ops.make_epi(EPI_RESID, xs, bias=P(bname), resid=x))
            ops.ln_fwd(xs, P(norm + "weight"), P(norm + "bias"), x, y2=h if bf else None)

        for i in range(cfg.decoder_layers):
            p = f"hf_model.decoder.layers.{i}."
            ops.gemm(h, W(p + "self_attn.in_proj_weight"), R, 3 * d, d,
                     ops.make_epi(EPI_STORE, qkv, bias=P(p + "self_attn.in_proj_bias")))
            ops.decode_self_attn(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], ctx["kc"][i], ctx["vc"][i], st.anc,
                                 st.cur_len, att, R, H, dh, L, beams=K)
            resid_norm(att, p + "self_attn.out_proj.weight", p + "self_attn.out_proj.bias", d, p + "norm1.")
            ops.gemm(h, W(p + "multihead_attn.in_proj_weight")[:d], R, d, d,
                     ops.make_epi(EPI_STORE, q, bias=P(p + "multihead_attn.in_proj_bias")[:d]))
            kv = ctx["kvmem"][i]
            ops.attn_fwd(q, kv[:, :d], kv[:, d:], att, None, st.B, H, K, ctx["S"], dh, kmask=ctx["enc_mask"])
            resid_norm(att, p + "multihead_attn.out_proj.weight", p + "multihead_attn.out_proj.bias", d, p + "norm2.")
            if not cfg.gated_linear:
                ops.gemm(h, W(p + "linear1.weight"), R, f, d, ops.make_epi(EPI_GELU, a, bias=P(p + "linear1.bias")))
            else:
                ops.gemm(h, W(p + "linear1.weight"), R, f, d, ops.make_epi(EPI_STORE, z, bias=P(p + "linear1.bias")))
                ops.gemm(h, W(p + "gate.weight"), R, f, d, ops.make_epi(EPI_GLU_MUL, a, bias=P(p + "gate.bias"), aux=z))
            resid_norm(a, p + "linear2.weight", p + "linear2.bias", f, p + "norm3.")
        hT = eng.buf("g.hT", (R, d), T)
        ops.ln_fwd(x, P("hf_model.decoder.norm.weight"), P("hf_model.decoder.norm.bias"), hT)
        logits = eng.buf("g.logits", (R, eng.ldv), torch.float32)
        ops.gemm(hT, W("hf_model.token_ff.weight"), R, cfg.vocab_size, d,
                 ops.make_epi(EPI_STORE, logits, bias=P("hf_model.token_ff.bias")))
        return logits

    def _forward_logits(self, st: BeamState, ctx: Dict[str, Any]):
        if not self.eng.norm_first:
            return self._forward_logits_post(st, ctx)
        plan = self._persist_plan(st)
        if plan is not None:
            return self._forward_logits_persist(st, ctx, plan)
        if self._small_ok(st.B * st.K):
            return self._forward_logits_small(st, ctx)
        eng, cfg = self.eng, self.eng.cfg
        d, H = cfg.d_model, cfg.decoder_attention_heads
        dh = d // H
        R, K, L = st.B * st.K, st.K, st.L
        T = eng.adt
        f = cfg.decoder_ffn_dim
        e = eng.ps.EMB
        tm = cfg.target_modality
        gam = bet = None
        if cfg.multimodal_norm:
            gam, bet = eng.P(f"{e}embedding_norm_dict.{tm}.weight"), eng.P(f"{e}embedding_norm_dict.{tm}.bias")
        x = eng.buf("g.x", (R, d), torch.float32)
        ops.decode_embed(st.next_tok, eng.P(f"{e}embedding_layer_dict.{tm}.weight"), gam, bet, ctx["pos"], st.cur_len, x)
        h = eng.buf("g.h", (R, d), T)
        qkv = eng.buf("g.qkv", (R, 3 * d), T)
        att = eng.buf("g.att", (R, d), T)
        q = eng.buf("g.q", (R, d), T)
        a = eng.buf("g.a", (R, f), T)
        z = eng.buf("g.z", (R, f), T)
        xa = eng.buf("g.xa", (R, d), torch.float32)
        xb = eng.buf("g.xb", (R, d), torch.float32)
        bf = eng.precision == "bf16"

        def resid_ln(a_in, wname, bname, k_in, resid, out, gname, bename):
            """out = resid + a_in W^T + b and h = LayerNorm(out) with the NEXT sub-layer's norm: one CTA-pair launch when
            the fused kernel applies (bf16, d_model 512, >= 256 rows), else the product followed by the LayerNorm."""
            epi = ops.make_epi(EPI_RESID, out, bias=eng.P(bname), resid=resid)
            if not (bf and (FUSE_DECODE_LN == 1 or (FUSE_DECODE_LN == 2 and k_in <= d) or
                            (FUSE_DECODE_LN == 3 and R <= 4736)) and
                    ops.gemm_resid_ln(a_in, eng.W(wname), R, d, k_in, epi, eng.P(gname), eng.P(bename), h)):
                ops.gemm(a_in, eng.W(wname), R, d, k_in, epi)
                ops.ln_fwd(out, eng.P(gname), eng.P(bename), h)

        L_ = cfg.decoder_layers
        ops.ln_fwd(x, eng.P("hf_model.decoder.layers.0.norm1.weight"), eng.P("hf_model.decoder.layers.0.norm1.bias"), h)
        for i in range(L_):
            p = f"hf_model.decoder.layers.{i}."
            ops.gemm(h, eng.W(p + "self_attn.in_proj_weight"), R, 3 * d, d,
                     ops.make_epi(EPI_STORE, qkv, bias=eng.P(p + "self_attn.in_proj_bias")))
            ops.decode_self_attn(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], ctx["kc"][i], ctx["vc"][i], st.anc,
                                 st.cur_len, att, R, H, dh, L, beams=K)
            resid_ln(att, p + "self_attn.out_proj.weight", p + "self_attn.out_proj.bias", d, x, xa,
                     p + "norm2.weight", p + "norm2.bias")
            ops.gemm(h, eng.W(p + "multihead_attn.in_proj_weight")[:d], R, d, d,
                     ops.make_epi(EPI_STORE, q, bias=eng.P(p + "multihead_attn.in_proj_bias")[:d]))
            kv = ctx["kvmem"][i]
            # the K beams of a spectrum attend over the same memory: one (spectrum, head) problem with K queries on
            # the tensor-core forward-attention kernel (K/V are read once per spectrum, not once per beam)
            # (a per-row SIMT kernel in the shape of decode_self_attn2 was measured here: 1.11 vs 0.96 ms / step at 2560
            # rows - every beam's CTA re-reads its spectrum's K / V from L2, 10x the tile kernel's traffic)
            ops.attn_fwd(q, kv[:, :d], kv[:, d:], att, None, st.B, H, K, ctx["S"], dh, kmask=ctx["enc_mask"])
            resid_ln(att, p + "multihead_attn.out_proj.weight", p + "multihead_attn.out_proj.bias", d, xa, xb,
                     p + "norm3.weight", p + "norm3.bias")
            if not cfg.gated_linear:
                ops.gemm(h, eng.W(p + "linear1.weight"), R, f, d, ops.make_epi(EPI_GELU, a, bias=eng.P(p + "linear1.bias")))
            else:
                if not (bf and ops.ffn_glu_fwd(h, eng.W(p + "linear1.weight"), eng.W(p + "gate.weight"),
                                               eng.P(p + "linear1.bias"), eng.P(p + "gate.bias"), R, f, d, a)):
                    ops.gemm(h, eng.W(p + "linear1.weight"), R, f, d,
                             ops.make_epi(EPI_STORE, z, bias=eng.P(p + "linear1.bias")))
                    ops.gemm(h, eng.W(p + "gate.weight"), R, f, d,
                             ops.make_epi(EPI_GLU_MUL, a, bias=eng.P(p + "gate.bias"), aux=z))
            nxt = f"hf_model.decoder.layers.{i + 1}.norm1." if i + 1 < L_ else "hf_model.decoder.norm."
            resid_ln(a, p + "linear2.weight", p + "linear2.bias", f, xb, x, nxt + "weight", nxt + "bias")
        V = cfg.vocab_size
        logits = eng.buf("g.logits", (R, eng.ldv), torch.float32)
        ops.gemm(h, eng.W("hf_model.token_ff.weight"), R, V, d,
                 ops.make_epi(EPI_STORE, logits, bias=eng.P("hf_model.token_ff.bias")))
        return logits

    @torch.no_grad()
    def generate(self, enc_inputs, enc_mask, n_beams: int = 1, max_length: Optional[int] = None, extra_bias=None,
                 use_graph: bool = True, check_every: int = 8, return_scores: bool = False, processors=None):
        """enc_inputs: {modality: batch-first tensor}; enc_mask: uint8 [B, S] (1 = real token).
        processors: list of `(input_ids, scores) -> scores` callables on CUDA tensors (transformers' LogitsProcessor
        protocol, applied after ForcedEOS as transformers orders them); a single `GuidedFormulaProcessor` is fused
        into the step kernel.
        Returns int64 [B * n_beams, L <= max_length]; row b*K + r is the r-th best hypothesis of spectrum b."""
        eng, cfg = self.eng, self.eng.cfg
        eng.sync_weights()
        B, S = enc_mask.shape
        K = int(n_beams)
        L = int(max_length or cfg.max_length)
        d = cfg.d_model
        R = B * K
        T = eng.adt
        mask_buf = eng.buf("g.enc_mask", (B, S), torch.uint8)
        mask_buf.copy_(enc_mask)
        mem = eng.encode(enc_inputs, mask_buf, train=False)
        skey = (B, K, L)
        st = self._states.get(skey)
        if st is None:
            st = self._states[skey] = BeamState(B, K, L, cfg.pad_token_id, cfg.bos_token_id, cfg.eos_token_id, eng.dev)
        else:
            st.reset()
        # cross-attention K/V once per spectrum and layer
        kvmem = eng.buf("g.kvmem", (cfg.decoder_layers, B * S, 2 * d), T)
        for i in range(cfg.decoder_layers):
            p = f"hf_model.decoder.layers.{i}.multihead_attn."
            ops.gemm(mem, eng.W(p + "in_proj_weight")[d:], B * S, 2 * d, d,
                     ops.make_epi(EPI_STORE, kvmem[i], bias=eng.P(p + "in_proj_bias")[d:]))
        kc = eng.buf("g.kc", (cfg.decoder_layers, R, L, d), T)
        vc = eng.buf("g.vc", (cfg.decoder_layers, R, L, d), T)
        pos = eng._pos_rows(L, "g")
        ctx = dict(kvmem=kvmem, kc=kc, vc=vc, pos=pos, enc_mask=mask_buf, S=S)

        if processors:
            self._run_processed(st, ctx, list(processors), extra_bias, use_graph, (B, K, L, S))
            return self._collect(st, return_scores)

        graph = None
        if use_graph:
            # every buffer touched by a step (workspace, caches, search state) is cached by shape, so the captured
            # step is reusable across calls of the same (B, K, L, S)
            gkey = (B, K, L, S, 0 if extra_bias is None else extra_bias.data_ptr())
            graph = self._capture(gkey, st, lambda: self._step(st, ctx, extra_bias))

        steps = 0
        max_steps = L - 1
        # finished-spectrum compaction: a spectrum whose search is over (beam: its finished pool can no longer improve,
        # greedy: it emitted <eos>) is frozen by the step kernels, but its K rows still ride through every product of the
        # step.  When at most half of the batch is still searching, the live spectra (state, K/V caches, cross K/V) are
        # gathered into a batch of half the size (shapes stay powers-of-two fractions of B, so the captured step graphs
        # are reused across calls) and the results of the retired ones are parked.  Output is unchanged: every spectrum's
        # state is independent of its neighbours.
        compact = COMPACT and extra_bias is None
        orig = torch.arange(B, device=eng.dev)  # original index of each spectrum of the current (compacted) batch
        parked = None                            # (seq [B, K, L] int32, score [B, K], len [B, K]) of retired spectra
        while steps < max_steps:
            n = min(check_every, max_steps - steps)
            for _ in range(n):
                if graph is not None:
                    graph.replay()
                else:
                    self._step(st, ctx, extra_bias)
            steps += n
            live = st.unfinished if K == 1 else st.improvable
            live_host = live.cpu()
            if K == 1:
                if not bool(live_host.any()):
                    break
            else:
                if not (bool(live_host.any()) and not bool(st.all_hit.all().item())):
                    break
            n_live = int(live_host.sum())
            if compact and st.B >= COMPACT_MIN_B and n_live < st.B and n_live <= COMPACT_LIVE_FRAC * st.B and steps < max_steps:
                if parked is None:
                    parked = self._park_init(B, K, L)
                st, ctx, graph, orig = self._compact(st, ctx, live_host, orig, parked, use_graph, S)
        if parked is None:
            return self._collect(st, return_scores)
        self._park(st, torch.arange(st.B, device=eng.dev), orig, parked)
        return self._collect_parked(parked, K, return_scores)

    # ------------------------------------------------------------------------------------ compaction
    def _park_init(self, B, K, L):
        cfg, dev = self.eng.cfg, self.eng.dev
        fill = cfg.pad_token_id if (K == 1 or cfg.pad_token_id) else cfg.eos_token_id
        return dict(seq=torch.full((B, K, L), fill, dtype=torch.int32, device=dev),
                    score=torch.full((B, K), -1.0e9, dtype=torch.float32, device=dev),
                    len=torch.zeros(B, K, dtype=torch.int32, device=dev))

    def _park(self, st: BeamState, idx, orig, parked):
        """Copy the final hypotheses of the spectra `idx` (indices into the current batch) to their original slots."""
        if idx.numel() == 0:
            return
        cur = int(st.cur_len.item())
        dst = orig[idx]
        if st.K == 1:
            parked["seq"][dst, 0] = st.run_seq[idx]
            parked["len"][dst, 0] = cur
        else:
            parked["seq"][dst] = st.fin_seq[cur & 1][idx]
            parked["score"][dst] = st.fin_score[idx]
            parked["len"][dst] = st.fin_len[idx]

    def _compact(self, st: BeamState, ctx, live_host, orig, parked, use_graph, S):
        eng, cfg = self.eng, self.eng.cfg
        K, L, d, T = st.K, st.L, cfg.d_model, self.eng.adt
        dev = eng.dev
        B_old = st.B
        n_live = max(int(live_host.sum()), 1)
        B_new = max(B_old // 2, 1)
        while B_new // 2 >= n_live:
            B_new //= 2
        B_new = max(B_new, n_live)
        live_idx = live_host.nonzero().flatten()
        dead_idx = (live_host == 0).nonzero().flatten()
        self._park(st, dead_idx.to(dev), orig, parked)
        # the new batch: every live spectrum + retired ones as filler (frozen, their results are already parked)
        keep = torch.cat([live_idx, dead_idx[: B_new - live_idx.numel()]]).to(dev)
        new = self._states.get((B_new, K, L))
        if new is None:
            new = self._states[(B_new, K, L)] = BeamState(B_new, K, L, cfg.pad_token_id, cfg.bos_token_id,
                                                          cfg.eos_token_id, dev)
        R_new = B_new * K
        nctx = dict(kvmem=eng.buf("g.kvmem", (cfg.decoder_layers, B_new * S, 2 * d), T),
                    kc=eng.buf("g.kc", (cfg.decoder_layers, R_new, L, d), T),
                    vc=eng.buf("g.vc", (cfg.decoder_layers, R_new, L, d), T), pos=ctx["pos"],
                    enc_mask=eng.buf("g.enc_mask", (B_new, S), torch.uint8), S=S)
        graph = None
        if use_graph:  # capture (warm-up + reset) BEFORE the state moves in
            graph = self._capture((B_new, K, L, S, 0), new, lambda: self._step(new, nctx, None))
        rows = (keep[:, None] * K + torch.arange(K, device=dev)[None, :]).reshape(-1)  # old row of every new row
        cur = int(st.cur_len.item())
        new.cur_len.copy_(st.cur_len)
        new.next_tok.copy_(st.next_tok[rows])
        if K == 1:
            new.run_seq.copy_(st.run_seq[keep])
            new.unfinished.copy_(st.unfinished[keep])
            new.parent_row.copy_(torch.arange(R_new, dtype=torch.int32, device=dev))
        else:
            shift = ((torch.arange(B_new, device=dev) - keep) * K).to(torch.int32)  # new row = old row + shift
            shift_rows = shift.repeat_interleave(K)
            new.run_seq.copy_(st.run_seq[:, keep])
            new.fin_seq.copy_(st.fin_seq[:, keep])
            for name in ("run_score", "fin_score", "fin_flag", "fin_len", "improvable", "all_hit"):
                getattr(new, name).copy_(getattr(st, name)[keep])
            new.parent_row.copy_(st.parent_row[rows] + shift_rows)
            new.anc.copy_(st.anc[:, rows] + shift_rows[None, :, None])
        nctx["enc_mask"].copy_(ctx["enc_mask"][keep])
        nctx["kvmem"].view(cfg.decoder_layers, B_new, S, 2 * d).copy_(
            ctx["kvmem"].view(cfg.decoder_layers, B_old, S, 2 * d)[:, keep])
        # only the positions written so far carry information
        nctx["kc"][:, :, :cur].copy_(ctx["kc"][:, rows, :cur])
        nctx["vc"][:, :, :cur].copy_(ctx["vc"][:, rows, :cur])
        self.compactions = getattr(self, "compactions", 0) + 1
        return new, nctx, graph, orig[keep]

    def _collect_parked(self, parked, K, return_scores):
        seq = parked["seq"].to(torch.int64)
        B, _, L = seq.shape
        cfg = self.eng.cfg
        if K == 1:
            s = seq[:, 0]
            is_eos = s == cfg.eos_token_id
            has = is_eos.any(dim=1)
            cur = parked["len"][:, 0].to(torch.int64)
            first = torch.where(has, is_eos.float().argmax(dim=1), cur - 1)
            out_len = int(first.max().item()) + 1
            return s[:, :out_len].contiguous()
        out_len = 1 + int(parked["len"].max().item())
        out = seq.reshape(B * K, L)[:, :out_len].contiguous()
        if return_scores:
            return out, parked["score"].reshape(B * K).clone()
        return out

    def _capture(self, gkey, st: BeamState, fn):
        graph = self._graphs.get(gkey)
        if graph is None:
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()  # warm-up: builds tensor maps, sets kernel attributes
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            st.reset()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                fn()
            st.reset()
            self._graphs[gkey] = graph
        return graph

    def _run_processed(self, st: BeamState, ctx, processors, extra_bias, use_graph, shape_key):
        """Decode loop with logits processors (wrapper.py:443-451 `logits_processor=`).  The decoder forward of a
        step is one graph replay.  A lone `GuidedFormulaProcessor` is fused: while the forward runs, the host reads
        the previous step's `next_tok` / `parent_row` (2 ints per row, copied to pinned memory before the replay was
        queued), extends its per-row strings, looks the element counts up in its memo and sends [rows, 14] counts
        back - the chemistry overlaps the GPU instead of serialising with it - and the step kernel applies the guide.
        Any other processor list gets dense scores (`mma_score_rows`) and CUDA `input_ids`, transformers' protocol.
        Selection follows un-captured."""
        from .guided import GuidedFormulaProcessor

        eng, cfg = self.eng, self.eng.cfg
        B, K, L = st.B, st.K, st.L
        R, V = B * K, cfg.vocab_size
        fused = len(processors) == 1 and isinstance(processors[0], GuidedFormulaProcessor)
        graph = self._capture(shape_key + ("fwd",), st, lambda: self._forward_logits(st, ctx)) if use_graph else None
        logits = eng.buf("g.logits", (R, eng.ldv), torch.float32)
        step_host = torch.empty(2, R, dtype=torch.int32).pin_memory()  # fused: next_tok | parent_row of the last step
        nflag = B if K == 1 else 2 * B
        flag_dev = eng.buf("g.flags", (nflag,), torch.uint8)
        flag_host = torch.ones(nflag, dtype=torch.uint8).pin_memory()
        cnt_host = torch.zeros(R, 14, dtype=torch.int32).pin_memory()
        cnt_dev = eng.buf("g.guide_counts", (R, 14), torch.int32)
        scores = None if fused else eng.buf("g.scores", (R, V), torch.float32)
        ev = torch.cuda.Event()
        if fused:
            processors[0].begin(R)
        steps = 0
        for cur in range(1, L):
            running = st.run_seq if K == 1 else st.run_seq[cur & 1].view(R, L)
            if fused:
                # the fused guide follows the search incrementally: 2 ints per row per step instead of the sequences
                if cur > 1:
                    step_host[0].copy_(st.next_tok, non_blocking=True)
                    step_host[1].copy_(st.parent_row, non_blocking=True)
            if K == 1:
                flag_dev.copy_(st.unfinished)
            else:
                flag_dev[:B].copy_(st.improvable)
                flag_dev[B:].copy_(st.all_hit)
            flag_host.copy_(flag_dev, non_blocking=True)
            ev.record()
            if graph is not None:
                graph.replay()
            else:
                self._forward_logits(st, ctx)
            ev.synchronize()
            if cur > 1:  # stop tests of transformers' `_sample` / `_beam_search` on the state after the last step
                if K == 1:
                    if not bool(flag_host.any()):
                        break
                elif not (bool(flag_host[:B].any()) and not bool(flag_host[B:].all())):
                    break
            if fused:
                if cur > 1:
                    processors[0].advance(step_host[1].tolist(), step_host[0].tolist())
                processors[0].counts_current(cnt_host)
                cnt_dev.copy_(cnt_host, non_blocking=True)
                self._select(st, logits, extra_bias, guide=processors[0].guide(cnt_dev))
            else:
                ops.score_rows(logits, scores, V, L, cfg.eos_token_id, st.cur_len, log_softmax=K > 1)
                ids_dev = running[:, :cur].to(torch.int64)
                sc = scores
                for proc in processors:
                    sc = proc(ids_dev, sc)
                if sc.data_ptr() != scores.data_ptr():
                    scores.copy_(sc)
                self._select(st, scores, extra_bias, prenorm=True)
            steps += 1
        torch.cuda.current_stream().synchronize()
        self.last_steps = steps  # search steps taken (diagnostics / benchmarks)

    def _collect(self, st: BeamState, return_scores: bool = False):
        cfg = self.eng.cfg
        K, L, R = st.K, st.L, st.B * st.K
        cur = int(st.cur_len.item())
        if K == 1:
            seq = st.run_seq.to(torch.int64)
            is_eos = seq == cfg.eos_token_id
            has = is_eos.any(dim=1)
            first = torch.where(has, is_eos.float().argmax(dim=1), torch.full_like(has, cur - 1, dtype=torch.int64))
            out_len = min(cur, int(first.max().item()) + 1)
            return seq[:, :out_len].contiguous()
        fin = st.fin_seq[cur & 1].reshape(R, L).to(torch.int64)
        out_len = 1 + int(st.fin_len.max().item())
        out = fin[:, :out_len].contiguous()
        if return_scores:
            return out, st.fin_score.reshape(R).clone()
        return out

