"""Forward / backward of the multimodal encoder-decoder, written as an explicit schedule of kernel launches.

Reference semantics (custom_modeling.py:108-199,220-243,271-320,420-508; modeling/utils.py:44-182,198-272):
pre-LN encoder/decoder layers over packed-QKV attention, exact-erf GELU FFN (optionally GLU-gated), final
LayerNorms, untied LM head, mean cross-entropy over non-pad targets.  There is no autograd graph: the
backward pass is the hand-scheduled mirror of the forward pass, so every buffer, stream and gradient bucket
is under our control (what a CUDA-graph capture and the overlapped gradient all-reduce need).

Precision modes
  "bf16": bf16 GEMM operands on tcgen05 (fp32 TMEM accumulate), fp32 residual stream / LN / softmax / loss,
          fp32 master weights with a bf16 mirror.
  "fp32": everything fp32 (SIMT GEMM) - the parity mode (logits/loss 1e-5, identical beam sequences).
"""
from __future__ import annotations

import os
from typing import Any, Callable, Dict, List, Optional

import torch

from . import ops
from ._lib import (EPI_ACCUM, EPI_DGELU, EPI_DGLU, EPI_DRELU, EPI_GELU, EPI_GLU_MUL, EPI_RELU, EPI_RESID, EPI_STORE)
from .params import TOKEN_TYPES, ModelConfig, ParamStore

NUM_SMS = 148


class Engine:
    def __init__(self, cfg: ModelConfig, params: ParamStore, precision: str = "bf16"):
        if precision not in ("bf16", "fp32"):
            raise ValueError(precision)
        self.cfg, self.ps = cfg, params
        self.precision = precision
        self.adt = torch.bfloat16 if precision == "bf16" else torch.float32
        self.dev = params.device
        self._bufs: Dict[Any, torch.Tensor] = {}
        self.on_release: List[Callable[[], None]] = []
        self.seed = 0x5EED0001
        # dropout seed lives on the device: a captured CUDA graph advances it between replays (ops.add_u64) while its
        # launch parameters stay constant (kernels dereference it: MMA_SITE_SEED_INDIRECT)
        self.seed_dev = torch.full((1,), self.seed, dtype=torch.int64, device=self.dev) if self.dev.type == "cuda" else None
        self.saved: Dict[str, Any] = {}
        self.grad_ready_hook: Optional[Callable[[int], None]] = None  # called with the flat offset from which
        #                                                               all gradients are final
        self.ldv = (cfg.vocab_size + 7) // 8 * 8
        self._wq: List[Any] = []  # weight / bias gradient products deferred to one grouped launch per layer
        self.group_wgrad = True
        # custom_modeling.py:129,176 hand `post_layer_normalisation` to torch as norm_first: True (every shipped config) is
        # x + f(LN(x)); False is LN(x + f(x)) - same kernels, other order (the `xb=` / `_post` paths below)
        self.norm_first = bool(cfg.post_layer_normalisation)
        # The grouped weight-gradient launch of a layer is off the critical path of backward (nothing but the
        # optimiser reads its output), so it CAN run on a side stream next to the memory-bound LayerNorm / attention
        # kernels of the next layer down (MMA_WGRAD_STREAM=1).  Measured on B200: no gain - the persistent 200 KB CTAs
        # leave no room for co-resident blocks - so the default keeps everything on one stream.  Its `dy` operands live in one of three rotating buffer sets (`wbuf`), so
        # the layer after next may overwrite them only once that launch is done (`_begin_group`).
        self.wgrad_stream = None
        if self.dev.type == "cuda" and os.environ.get("MMA_WGRAD_STREAM", "0") != "0":
            self.wgrad_stream = torch.cuda.Stream(device=self.dev)
        self._wev: Dict[int, Any] = {}
        self._bset = 0
        self.last_wgrad_event = None

    # ------------------------------------------------------------------------------------- utils
    def buf(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=self.dev)
            self._bufs[key] = t
        return t

    def release_buffers(self):
        """Drop every shape-keyed workspace.  Captured CUDA graphs hold raw pointers into them, so every graph owner
        (trainer, generator) registered in `on_release` drops its graphs first."""
        for fn in self.on_release:
            fn()
        self._bufs.clear()
        self.saved = {}
        self._wq = []

    def sync_weights(self):
        """Refresh the bf16 mirror of the master weights (after load_state_dict / an external optimiser)."""
        if self.precision == "bf16" and self.ps.bf16_dirty:
            ops.cast_f32_bf16(self.ps.p, self.ps.pb)
        self.ps.bf16_dirty = False

    def W(self, name):
        return self.ps.PB(name) if self.precision == "bf16" else self.ps.P(name)

    def P(self, name):
        return self.ps.P(name)

    def G(self, name):
        return self.ps.G(name)

    def _site(self, dec: bool, layer: int, k: int) -> int:
        return ((1024 if dec else 0) + 16 * layer + k) | 0x80000000

    @property
    def seed_arg(self) -> int:
        return self.seed_dev.data_ptr()

    def set_seed(self, seed: Optional[int], rank: int = 0):
        """Dropout stream of this process: a function of the user's seed and the rank, so DDP ranks draw different
        masks and two runs with different seeds differ (the reference gets both from torch's per-process generator)."""
        base = 0x5EED0001 if seed is None else (int(seed) * 0x9E3779B97F4A7C15 + 0x5EED0001)
        self.seed = (base + (int(rank) << 40)) & 0x7FFFFFFFFFFFFFFF
        if self.seed_dev is not None:
            self.seed_dev.fill_(self.seed)

    def next_seed(self):
        """Fresh dropout masks for the next training step (stream-ordered, graph-capturable)."""
        ops.add_u64(self.seed_dev, 1)

    def _splits(self, n_out, k_in, rows):
        tiles = ((n_out + 127) // 128) * ((k_in + 127) // 128)
        kb = (rows + 63) // 64
        s = max(1, NUM_SMS // tiles)
        return max(1, min(s, kb // 2 if kb >= 2 else 1, 64))

    # ----------------------------------------------------------------------------- linear helpers
    def _lin_bwd(self, dy, x, wname, bname, rows, n_out, k_in, dx_epi=None, row_slice=None):
        """dy: [rows, n_out]; x: [rows, k_in]; W: [n_out, k_in] (optionally a row slice of the named tensor).
        dgrad is launched now; wgrad + bias grad are queued for the layer's grouped launch (`_flush_wgrads`)."""
        Wt, Gw, Gb = self.W(wname), self.G(wname), self.G(bname)
        if row_slice is not None:
            Wt, Gw, Gb = Wt[row_slice], Gw[row_slice], Gb[row_slice]
        if dx_epi is not None:
            ops.gemm(dy, Wt, rows, k_in, n_out, dx_epi, b_mn=True)
        if self.group_wgrad and self.precision == "bf16" and ops.wgrad_group_ok(dy, x, n_out, k_in):
            self._wq.append((dy, x, Gw, Gb, n_out, k_in, rows))
            return
        ops.gemm(dy, x, n_out, k_in, rows, ops.make_epi(EPI_ACCUM, Gw, accumulate=2), a_mn=True, b_mn=True,
                 splits=self._splits(n_out, k_in, rows))
        ops.colsum(dy, Gb, rows=rows, cols=n_out)

    def wbuf(self, name, shape, dtype, ahead=0):
        """Buffer read by the grouped wgrad launch of backward group `_bset + ahead` (three rotating sets)."""
        return self.buf(f"{name}.s{(self._bset + ahead) % 3}", shape, dtype)

    def _begin_group(self, b: int):
        """Backward group b (LM head = 0, then one per layer) starts: its buffer set (and, at its end, the next one)
        is rewritten, so the wgrad launch of group b - 2 must have finished reading."""
        self._bset = b
        ev = self._wev.pop(b - 2, None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def _flush_wgrads(self):
        if not self._wq:
            return
        if self.wgrad_stream is None:
            ops.wgrad_group(self._wq)
        else:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream())
            self.wgrad_stream.wait_event(ready)
            with torch.cuda.stream(self.wgrad_stream):
                ops.wgrad_group(self._wq)
                done = torch.cuda.Event()
                done.record(self.wgrad_stream)
            self._wev[self._bset] = done
            self.last_wgrad_event = done
        self._wq = []

    def _join_wgrads(self):
        if self.wgrad_stream is not None:
            torch.cuda.current_stream().wait_stream(self.wgrad_stream)
            self._wev.clear()

    # --------------------------------------------------------------------------------- embedding
    def _pos_rows(self, L, tag):
        ps, cfg = self.ps, self.cfg
        e = ps.EMB + "positional_encodings."
        if cfg.positional_encoding_type == "sin_cos":
            return ps.buffers[e + "pos_enc"]
        pl = self.buf(f"{tag}.posln", (L, cfg.d_model), torch.float32)
        ops.ln_fwd(self.P(e + "pos_encodings.weight")[:L], self.P(e + "norm.weight"), self.P(e + "norm.bias"), pl)
        return pl

    def _embed(self, inputs: Dict[str, Any], L_total: int, tag: str, B: int):
        cfg, d = self.cfg, self.cfg.d_model
        x0 = self.buf(f"{tag}.x0", (B * L_total, d), torch.float32)
        pos = self._pos_rows(L_total, tag)
        recs: List[Dict[str, Any]] = []
        off = 0
        e = self.ps.EMB
        for m, val in inputs.items():
            mc = cfg.data_config[m]
            base = f"{e}embedding_layer_dict.{m}."
            if mc["type"] in TOKEN_TYPES:
                ids = (val["tokenized_input"] if isinstance(val, dict) else val).contiguous()
                scale = val["numerical_values"].contiguous().float() if isinstance(val, dict) else None
                S_m = ids.shape[1]
                pre = self.buf(f"{tag}.{m}.pre", (B * S_m, d), torch.float32)
                ops.gather_rows(ids.view(-1), self.P(base + "weight"), pre, scale=None if scale is None else scale.view(-1))
                rec = dict(kind="tok", ids=ids, scale=scale, pre=pre)
            else:
                xin = val.contiguous().float()
                S_m = xin.shape[1]
                h = xin.view(B * S_m, -1)
                layers = cfg.embed_layers(m)
                hs = [h]
                for li, (o, i) in enumerate(layers):
                    pre_n = base if len(layers) == 1 else f"{base}{2 * li}."
                    out = self.buf(f"{tag}.{m}.h{li}", (B * S_m, o), torch.float32)
                    kind = EPI_STORE if li == len(layers) - 1 else EPI_RELU
                    ops.gemm(h, self.P(pre_n + "weight"), B * S_m, o, i,
                             ops.make_epi(kind, out, bias=self.P(pre_n + "bias")), force_simt=True)
                    h = out
                    hs.append(out)
                pre = h
                rec = dict(kind="patch", hs=hs)
            gam = bet = None
            if cfg.multimodal_norm:
                gam, bet = self.P(f"{e}embedding_norm_dict.{m}.weight"), self.P(f"{e}embedding_norm_dict.{m}.bias")
            ops.ln_fwd(pre, gam, bet, x0, add=pos, group=S_m, out_group_stride=L_total, out_offset=off)
            rec.update(m=m, off=off, S=S_m, pre=pre)
            recs.append(rec)
            off += S_m
        if off != L_total:
            raise ValueError(f"modalities cover {off} positions, mask has {L_total}")
        return x0, recs

    def _embed_bwd(self, dx0, recs, L_total: int, tag: str, B: int):
        cfg, d = self.cfg, self.cfg.d_model
        e = self.ps.EMB
        for rec in recs:
            m, off, S_m = rec["m"], rec["off"], rec["S"]
            base = f"{e}embedding_layer_dict.{m}."
            dpre = self.buf(f"{tag}.{m}.dpre", (B * S_m, d), torch.float32)
            gam = dg = db = None
            if cfg.multimodal_norm:
                gam = self.P(f"{e}embedding_norm_dict.{m}.weight")
                dg, db = self.G(f"{e}embedding_norm_dict.{m}.weight"), self.G(f"{e}embedding_norm_dict.{m}.bias")
            ops.ln_bwd(dx0, rec["pre"], gam, dx=dpre, dgamma=dg, dbeta=db, group=S_m, in_group_stride=L_total,
                       in_offset=off)
            if rec["kind"] == "tok":
                ops.scatter_add_rows(rec["ids"].view(-1), dpre, self.G(base + "weight"),
                                     pad_idx=cfg.data_config[m]["pad_token_id"],
                                     scale=None if rec["scale"] is None else rec["scale"].view(-1))
            else:
                layers = cfg.embed_layers(m)
                hs = rec["hs"]
                dy = dpre
                for li in reversed(range(len(layers))):
                    o, i = layers[li]
                    pre_n = base if len(layers) == 1 else f"{base}{2 * li}."
                    rows = B * S_m
                    ops.gemm(dy, hs[li], o, i, rows, ops.make_epi(EPI_ACCUM, self.G(pre_n + "weight"), accumulate=1),
                             a_mn=True, b_mn=True, force_simt=True)
                    ops.colsum(dy, self.G(pre_n + "bias"), rows=rows, cols=o)
                    if li > 0:
                        dh = self.buf(f"{tag}.{m}.dh{li}", (rows, i), torch.float32)
                        ops.gemm(dy, self.P(pre_n + "weight"), rows, i, o, ops.make_epi(EPI_DRELU, dh, aux=hs[li]),
                                 b_mn=True, force_simt=True)
                        dy = dh
        if cfg.positional_encoding_type == "learned":
            pe = e + "positional_encodings."
            dpos = self.buf(f"{tag}.dposln", (L_total * d,), torch.float32)
            dpos.zero_()
            ops.colsum(dx0.view(B, L_total * d), dpos, rows=B, cols=L_total * d)
            gtab = self.G(pe + "pos_encodings.weight")[:L_total]
            ops.ln_bwd(dpos.view(L_total, d), self.P(pe + "pos_encodings.weight")[:L_total], self.P(pe + "norm.weight"),
                       dx=gtab, dres=gtab, dgamma=self.G(pe + "norm.weight"), dbeta=self.G(pe + "norm.bias"))

    def _resid_gemm(self, a, wname, bname, M, n_out, k_in, xo, x, p, site_r, nxt):
        """xo = x + drop(a W^T + b).  `nxt` = (gamma, beta, h buffer) of the LayerNorm that consumes xo next: when the
        fused CTA-pair kernel applies (d_model 512, bf16), h = LN(xo) comes out of the same launch.  Returns h or None."""
        epi = ops.make_epi(EPI_RESID, xo, bias=self.P(bname), resid=x, p_drop=p, seed=self.seed_arg, site=site_r)
        if nxt is not None and self.precision == "bf16" and ops.fuse_ln_wanted(M, k_in):
            gam, bet, hbuf = nxt
            if ops.gemm_resid_ln(a, self.W(wname), M, n_out, k_in, epi, gam, bet, hbuf):
                return hbuf
        ops.gemm(a, self.W(wname), M, n_out, k_in, epi)
        return None

    def _next_norm(self, wname, bname, tag, M):
        """(gamma, beta, output buffer) for the LayerNorm that opens the block `tag`."""
        return self.P(wname), self.P(bname), self.buf(tag + ".h", (M, self.cfg.d_model), self.adt)

    def _operand_copy(self, x, name):
        """The fp32 residual stream as a GEMM operand (LN(x + f(x)) layers feed the stream itself to the next block)."""
        if self.adt == torch.float32:
            return x
        xb = self.buf(name, tuple(x.shape), self.adt)
        ops.cast_f32_bf16(x, xb)
        return xb

    def _post_norm_fwd(self, tag, xo, wp, M):
        """LN(x + f(x)) layers: the block's output is LayerNorm(xo) - in fp32 for the residual stream and in the
        activation dtype as the next block's operand, from one launch."""
        xn = self.buf(tag + ".xn", (M, self.cfg.d_model), torch.float32)
        if self.adt == torch.float32:
            ops.ln_fwd(xo, self.P(wp["n_w"]), self.P(wp["n_b"]), xn)
            return xn, xn
        xnb = self.buf(tag + ".xnb", (M, self.cfg.d_model), self.adt)
        ops.ln_fwd(xo, self.P(wp["n_w"]), self.P(wp["n_b"]), xn, y2=xnb)
        return xn, xnb

    def _post_norm_bwd(self, tag, kind, dxn, M, wp, p, site_r):
        """Backward of `_post_norm_fwd`: dxn (fp32, total gradient of the block output) -> ds (fp32 gradient of
        s = x + drop(f(x)); the sub-layer's input gradient is accumulated into it afterwards, which makes it the total
        gradient of the block input) and its dropout-masked low-precision copy, the gradient of f's output."""
        d = self.cfg.d_model
        ds = self._other_dx(dxn, M)
        dsb = self.wbuf("bw.dsb." + kind, (M, d), self.adt)
        ops.ln_bwd(dxn, self.saved[tag]["x"], self.P(wp["n_w"]), dx=ds, dxb=dsb, dgamma=self.G(wp["n_w"]),
                   dbeta=self.G(wp["n_b"]), p_drop=p, seed=self.seed_arg, site=site_r)
        return ds, dsb

    # ----------------------------------------------------------------------------- layer forward
    def _attn_block_fwd(self, tag, x, B, L, heads, wp, kmask, causal, p, site_a, site_r, train, h_pre=None, nxt=None,
                        xb=None):
        """x -> x + drop(out_proj(attn(LN(x))))  (self-attention).  `h_pre`: LN(x) if the previous block's fused
        product already made it; `nxt`: the LayerNorm that follows.  Returns (new residual stream, LN of it or None).
        `xb` (x in the activation dtype) selects the LN(x + f(x)) form: returns (LN(x + drop(out_proj(attn(x)))), its
        activation-dtype copy)."""
        d = self.cfg.d_model
        M = B * L
        dh = d // heads
        T = self.adt
        if xb is not None:
            h = xb
        else:
            h = self.buf(tag + ".h", (M, d), T)
            if h_pre is None:
                ops.ln_fwd(x, self.P(wp["n_w"]), self.P(wp["n_b"]), h)
            else:
                assert h_pre.data_ptr() == h.data_ptr()
        qkv = self.buf(tag + ".qkv", (M, 3 * d), T)
        ops.gemm(h, self.W(wp["in_w"]), M, 3 * d, d, ops.make_epi(EPI_STORE, qkv, bias=self.P(wp["in_b"])))
        ctx = self.buf(tag + ".ctx", (M, d), T)
        lse = self.buf(tag + ".lse", (B * heads * L,), torch.float32)
        ops.attn_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], ctx, lse, B, heads, L, L, dh, kmask=kmask,
                     causal=causal, p_drop=p, seed=self.seed_arg, site=site_a)
        xo = self.buf(tag + ".xo", (M, d), torch.float32)
        h_next = self._resid_gemm(ctx, wp["out_w"], wp["out_b"], M, d, d, xo, x, p, site_r, None if xb is not None else nxt)
        if train:
            self.saved[tag] = dict(x=xo if xb is not None else x, h=h, qkv=qkv, ctx=ctx, lse=lse)
        if xb is not None:
            return self._post_norm_fwd(tag, xo, wp, M)
        return xo, h_next

    def _attn_block_bwd(self, tag, dx, dyb, B, L, heads, wp, kmask, causal, p, site_a, prev_site, first=False,
                        out_tag="x"):
        """dx: fp32 grad of the block output; dyb: low-precision (dropout-masked) copy.  Returns (dx_in, dyb_in)."""
        d = self.cfg.d_model
        M = B * L
        dh = d // heads
        T = self.adt
        s = self.saved[tag]
        dctx = self.wbuf("bw.dctx", (M, d), T)
        self._lin_bwd(dyb, s["ctx"], wp["out_w"], wp["out_b"], M, d, d, dx_epi=ops.make_epi(EPI_STORE, dctx))
        dqkv = self.wbuf("bw.dqkv", (M, 3 * d), T)
        qkv = s["qkv"]
        ops.attn_bwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], s["ctx"], s["lse"], dctx, dqkv[:, :d],
                     dqkv[:, d:2 * d], dqkv[:, 2 * d:], B, heads, L, L, dh, kmask=kmask, causal=causal, p_drop=p,
                     seed=self.seed_arg, site=site_a, dsum=self.buf("bw.dsum", (B * heads * L,), torch.float32))
        dh_ = self.buf("bw.dh", (M, d), T)
        self._lin_bwd(dqkv, s["h"], wp["in_w"], wp["in_b"], M, 3 * d, d, dx_epi=ops.make_epi(EPI_STORE, dh_))
        dx_in = self._other_dx(dx, M)
        dyb_in = None if first else self.wbuf("bw.dyb.x", (M, d), T, ahead=1)  # read by the next group down
        ops.ln_bwd(dh_, s["x"], self.P(wp["n_w"]), dx=dx_in, dres=dx, dxb=dyb_in, dgamma=self.G(wp["n_w"]),
                   dbeta=self.G(wp["n_b"]), p_drop=p, seed=self.seed_arg, site=prev_site)
        return dx_in, dyb_in

    def _ffn_block_fwd(self, tag, x, M, f, wp, p, site_i, site_r, train, h_pre=None, nxt=None, xb=None):
        d = self.cfg.d_model
        T = self.adt
        if xb is not None:  # LN(x + f(x)) form, see _attn_block_fwd
            h = xb
        else:
            h = self.buf(tag + ".h", (M, d), T)
            if h_pre is None:
                ops.ln_fwd(x, self.P(wp["n_w"]), self.P(wp["n_b"]), h)
            else:
                assert h_pre.data_ptr() == h.data_ptr()
        a = self.buf(tag + ".a", (M, f), T)
        z = self.buf(tag + ".z", (M, f), T)
        z2 = None
        if not self.cfg.gated_linear:
            ops.gemm(h, self.W(wp["w1"]), M, f, d,
                     ops.make_epi(EPI_GELU, a, out2=z if train else None, bias=self.P(wp["b1"]), p_drop=p,
                                  seed=self.seed_arg, site=site_i))
        else:
            z2 = self.buf(tag + ".z2", (M, f), T)
            # both products of the gate in one CTA-pair launch (h read once, z1 never re-read); small / fp32 shapes
            # take the two single-CTA launches
            if not (self.precision == "bf16" and
                    ops.ffn_glu_fwd(h, self.W(wp["w1"]), self.W(wp["wg"]), self.P(wp["b1"]), self.P(wp["bg"]), M, f, d, a,
                                    z1=z if train else None, z2=z2 if train else None, p_drop=p, seed=self.seed_arg,
                                    site=site_i)):
                ops.gemm(h, self.W(wp["w1"]), M, f, d, ops.make_epi(EPI_STORE, z, bias=self.P(wp["b1"])))
                ops.gemm(h, self.W(wp["wg"]), M, f, d,
                         ops.make_epi(EPI_GLU_MUL, a, out2=z2, bias=self.P(wp["bg"]), aux=z, p_drop=p,
                                      seed=self.seed_arg, site=site_i))
        xo = self.buf(tag + ".xo", (M, d), torch.float32)
        h_next = self._resid_gemm(a, wp["w2"], wp["b2"], M, d, f, xo, x, p, site_r, None if xb is not None else nxt)
        if train:
            self.saved[tag] = dict(x=xo if xb is not None else x, h=h, a=a, z=z, z2=z2)
        if xb is not None:
            return self._post_norm_fwd(tag, xo, wp, M)
        return xo, h_next

    def _ffn_block_bwd(self, tag, dx, dyb, M, f, wp, p, site_i, prev_site):
        d = self.cfg.d_model
        T = self.adt
        s = self.saved[tag]
        dz = self.wbuf("bw.dz", (M, f), T)
        dh_ = self.buf("bw.dh", (M, d), T)
        if not self.cfg.gated_linear:
            self._lin_bwd(dyb, s["a"], wp["w2"], wp["b2"], M, d, f,
                          dx_epi=ops.make_epi(EPI_DGELU, dz, aux=s["z"], p_drop=p, seed=self.seed_arg, site=site_i, drop_ld=f))
            self._lin_bwd(dz, s["h"], wp["w1"], wp["b1"], M, f, d, dx_epi=ops.make_epi(EPI_STORE, dh_))
        else:
            dz2 = self.wbuf("bw.dz2", (M, f), T)
            bf = self.precision == "bf16"
            if bf and ops.ffn_dglu(dyb, self.W(wp["w2"]), M, f, d, s["z"], s["z2"], dz, dz2, p_drop=p,
                                   seed=self.seed_arg, site=site_i, drop_ld=f):
                self._lin_bwd(dyb, s["a"], wp["w2"], wp["b2"], M, d, f)  # weight / bias gradient only
            else:
                self._lin_bwd(dyb, s["a"], wp["w2"], wp["b2"], M, d, f,
                              dx_epi=ops.make_epi(EPI_DGLU, dz, out2=dz2, aux=s["z"], aux2=s["z2"], p_drop=p,
                                                  seed=self.seed_arg, site=site_i, drop_ld=f))
            # dh = dz W1 + dz2 Wg: one accumulation over both reductions on the pair kernel, else the second product
            # lands through the accumulate epilogue (fp32)
            if bf and ops.gemm_dual(dz, self.W(wp["w1"]), dz2, self.W(wp["wg"]), M, d, f, f,
                                    ops.make_epi(EPI_STORE, dh_), b_mn=True):
                self._lin_bwd(dz, s["h"], wp["w1"], wp["b1"], M, f, d)
                self._lin_bwd(dz2, s["h"], wp["wg"], wp["bg"], M, f, d)
            else:
                dh_ = self.buf("bw.dhs", (M, d), torch.float32)
                self._lin_bwd(dz, s["h"], wp["w1"], wp["b1"], M, f, d, dx_epi=ops.make_epi(EPI_ACCUM, dh_, accumulate=0))
                self._lin_bwd(dz2, s["h"], wp["wg"], wp["bg"], M, f, d, dx_epi=ops.make_epi(EPI_ACCUM, dh_, accumulate=1))
        dx_in = self._other_dx(dx, M)
        dyb_in = self.wbuf("bw.dyb.ffn_in", (M, d), T)
        ops.ln_bwd(dh_, s["x"], self.P(wp["n_w"]), dx=dx_in, dres=dx, dxb=dyb_in, dgamma=self.G(wp["n_w"]),
                   dbeta=self.G(wp["n_b"]), p_drop=p, seed=self.seed_arg, site=prev_site)
        return dx_in, dyb_in

    # ---- LN(x + f(x)) layers: backward.  dxn = total fp32 gradient of the block output; returns that of the block input
    def _attn_block_bwd_post(self, tag, dxn, B, L, heads, wp, kmask, causal, p, site_a, site_r):
        d = self.cfg.d_model
        M = B * L
        dh = d // heads
        T = self.adt
        s = self.saved[tag]
        ds, dsb = self._post_norm_bwd(tag, "sa", dxn, M, wp, p, site_r)
        dctx = self.wbuf("bw.dctx", (M, d), T)
        self._lin_bwd(dsb, s["ctx"], wp["out_w"], wp["out_b"], M, d, d, dx_epi=ops.make_epi(EPI_STORE, dctx))
        dqkv = self.wbuf("bw.dqkv", (M, 3 * d), T)
        qkv = s["qkv"]
        ops.attn_bwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], s["ctx"], s["lse"], dctx, dqkv[:, :d],
                     dqkv[:, d:2 * d], dqkv[:, 2 * d:], B, heads, L, L, dh, kmask=kmask, causal=causal, p_drop=p,
                     seed=self.seed_arg, site=site_a, dsum=self.buf("bw.dsum", (B * heads * L,), torch.float32))
        self._lin_bwd(dqkv, s["h"], wp["in_w"], wp["in_b"], M, 3 * d, d,
                      dx_epi=ops.make_epi(EPI_ACCUM, ds, accumulate=1))
        return ds

    def _cross_block_bwd_post(self, tag, dxn, mem, dmem, first_mem, B, T_, S, heads, wp, kmask, p, site_a, site_r):
        d = self.cfg.d_model
        M, Me = B * T_, B * S
        dh = d // heads
        T = self.adt
        s = self.saved[tag]
        ds, dsb = self._post_norm_bwd(tag, "ca", dxn, M, wp, p, site_r)
        dctx = self.wbuf("bw.dctx", (M, d), T)
        self._lin_bwd(dsb, s["ctx"], wp["out_w"], wp["out_b"], M, d, d, dx_epi=ops.make_epi(EPI_STORE, dctx))
        dq = self.wbuf("bw.dq", (M, d), T)
        dkv = self.wbuf("bw.dkv", (Me, 2 * d), T)
        kv = s["kv"]
        ops.attn_bwd(s["q"], kv[:, :d], kv[:, d:], s["ctx"], s["lse"], dctx, dq, dkv[:, :d], dkv[:, d:], B, heads, T_,
                     S, dh, kmask=kmask, causal=False, p_drop=p, seed=self.seed_arg, site=site_a,
                     dsum=self.buf("bw.dsum", (B * heads * T_,), torch.float32))
        self._lin_bwd(dq, s["h"], wp["in_w"], wp["in_b"], M, d, d, dx_epi=ops.make_epi(EPI_ACCUM, ds, accumulate=1),
                      row_slice=slice(0, d))
        self._lin_bwd(dkv, mem, wp["in_w"], wp["in_b"], Me, 2 * d, d,
                      dx_epi=ops.make_epi(EPI_ACCUM, dmem, accumulate=0 if first_mem else 1),
                      row_slice=slice(d, 3 * d))
        return ds

    def _ffn_block_bwd_post(self, tag, dxn, M, f, wp, p, site_i, site_r):
        d = self.cfg.d_model
        T = self.adt
        s = self.saved[tag]
        ds, dsb = self._post_norm_bwd(tag, "ff", dxn, M, wp, p, site_r)
        dz = self.wbuf("bw.dz", (M, f), T)
        acc = ops.make_epi(EPI_ACCUM, ds, accumulate=1)
        if not self.cfg.gated_linear:
            self._lin_bwd(dsb, s["a"], wp["w2"], wp["b2"], M, d, f,
                          dx_epi=ops.make_epi(EPI_DGELU, dz, aux=s["z"], p_drop=p, seed=self.seed_arg, site=site_i, drop_ld=f))
            self._lin_bwd(dz, s["h"], wp["w1"], wp["b1"], M, f, d, dx_epi=acc)
            return ds
        dz2 = self.wbuf("bw.dz2", (M, f), T)
        if self.precision == "bf16" and ops.ffn_dglu(dsb, self.W(wp["w2"]), M, f, d, s["z"], s["z2"], dz, dz2, p_drop=p,
                                                     seed=self.seed_arg, site=site_i, drop_ld=f):
            self._lin_bwd(dsb, s["a"], wp["w2"], wp["b2"], M, d, f)  # weight / bias gradient only
        else:
            self._lin_bwd(dsb, s["a"], wp["w2"], wp["b2"], M, d, f,
                          dx_epi=ops.make_epi(EPI_DGLU, dz, out2=dz2, aux=s["z"], aux2=s["z2"], p_drop=p,
                                              seed=self.seed_arg, site=site_i, drop_ld=f))
        self._lin_bwd(dz, s["h"], wp["w1"], wp["b1"], M, f, d, dx_epi=acc)
        self._lin_bwd(dz2, s["h"], wp["wg"], wp["bg"], M, f, d, dx_epi=ops.make_epi(EPI_ACCUM, ds, accumulate=1))
        return ds

    def _other_dx(self, dx, M):
        a = self.buf(f"bw.dx.{M}.0", (M, self.cfg.d_model), torch.float32)
        if dx is None or dx.data_ptr() != a.data_ptr():
            return a
        return self.buf(f"bw.dx.{M}.1", (M, self.cfg.d_model), torch.float32)

    @staticmethod
    def _wp_attn(prefix, attn, norm):
        return dict(in_w=f"{prefix}{attn}.in_proj_weight", in_b=f"{prefix}{attn}.in_proj_bias",
                    out_w=f"{prefix}{attn}.out_proj.weight", out_b=f"{prefix}{attn}.out_proj.bias",
                    n_w=f"{prefix}{norm}.weight", n_b=f"{prefix}{norm}.bias")

    @staticmethod
    def _wp_ffn(prefix, norm):
        return dict(w1=prefix + "linear1.weight", b1=prefix + "linear1.bias", wg=prefix + "gate.weight",
                    bg=prefix + "gate.bias", w2=prefix + "linear2.weight", b2=prefix + "linear2.bias",
                    n_w=f"{prefix}{norm}.weight", n_b=f"{prefix}{norm}.bias")

    # -------------------------------------------------------------------------------- cross-attn
    def _cross_block_fwd(self, tag, x, mem, B, T_, S, heads, wp, kmask, p, site_a, site_r, train, h_pre=None, nxt=None,
                         xb=None):
        d = self.cfg.d_model
        M, Me = B * T_, B * S
        dh = d // heads
        T = self.adt
        if xb is not None:  # LN(x + f(x)) form, see _attn_block_fwd
            h = xb
        else:
            h = self.buf(tag + ".h", (M, d), T)
            if h_pre is None:
                ops.ln_fwd(x, self.P(wp["n_w"]), self.P(wp["n_b"]), h)
            else:
                assert h_pre.data_ptr() == h.data_ptr()
        q = self.buf(tag + ".q", (M, d), T)
        Win, bin_ = self.W(wp["in_w"]), self.P(wp["in_b"])
        ops.gemm(h, Win[:d], M, d, d, ops.make_epi(EPI_STORE, q, bias=bin_[:d]))
        kv = self.buf(tag + ".kv", (Me, 2 * d), T)
        ops.gemm(mem, Win[d:], Me, 2 * d, d, ops.make_epi(EPI_STORE, kv, bias=bin_[d:]))
        ctx = self.buf(tag + ".ctx", (M, d), T)
        lse = self.buf(tag + ".lse", (B * heads * T_,), torch.float32)
        ops.attn_fwd(q, kv[:, :d], kv[:, d:], ctx, lse, B, heads, T_, S, dh, kmask=kmask, causal=False, p_drop=p,
                     seed=self.seed_arg, site=site_a)
        xo = self.buf(tag + ".xo", (M, d), torch.float32)
        h_next = self._resid_gemm(ctx, wp["out_w"], wp["out_b"], M, d, d, xo, x, p, site_r, None if xb is not None else nxt)
        if train:
            self.saved[tag] = dict(x=xo if xb is not None else x, h=h, q=q, kv=kv, ctx=ctx, lse=lse)
        if xb is not None:
            return self._post_norm_fwd(tag, xo, wp, M)
        return xo, h_next

    def _cross_block_bwd(self, tag, dx, dyb, mem, dmem, first_mem, B, T_, S, heads, wp, kmask, p, site_a, prev_site):
        d = self.cfg.d_model
        M, Me = B * T_, B * S
        dh = d // heads
        T = self.adt
        s = self.saved[tag]
        dctx = self.wbuf("bw.dctx", (M, d), T)
        self._lin_bwd(dyb, s["ctx"], wp["out_w"], wp["out_b"], M, d, d, dx_epi=ops.make_epi(EPI_STORE, dctx))
        dq = self.wbuf("bw.dq", (M, d), T)
        dkv = self.wbuf("bw.dkv", (Me, 2 * d), T)
        kv = s["kv"]
        ops.attn_bwd(s["q"], kv[:, :d], kv[:, d:], s["ctx"], s["lse"], dctx, dq, dkv[:, :d], dkv[:, d:], B, heads, T_,
                     S, dh, kmask=kmask, causal=False, p_drop=p, seed=self.seed_arg, site=site_a,
                     dsum=self.buf("bw.dsum", (B * heads * T_,), torch.float32))
        dh_ = self.buf("bw.dh", (M, d), T)
        self._lin_bwd(dq, s["h"], wp["in_w"], wp["in_b"], M, d, d, dx_epi=ops.make_epi(EPI_STORE, dh_),
                      row_slice=slice(0, d))
        self._lin_bwd(dkv, mem, wp["in_w"], wp["in_b"], Me, 2 * d, d,
                      dx_epi=ops.make_epi(EPI_ACCUM, dmem, accumulate=0 if first_mem else 1),
                      row_slice=slice(d, 3 * d))
        dx_in = self._other_dx(dx, M)
        dyb_in = self.wbuf("bw.dyb.ca_in", (M, d), T)
        ops.ln_bwd(dh_, s["x"], self.P(wp["n_w"]), dx=dx_in, dres=dx, dxb=dyb_in, dgamma=self.G(wp["n_w"]),
                   dbeta=self.G(wp["n_b"]), p_drop=p, seed=self.seed_arg, site=prev_site)
        return dx_in, dyb_in

    # ------------------------------------------------------------------------------------ encoder
    def encode(self, enc_inputs, enc_mask, train=False):
        """-> memory [B*S, d] in the activation dtype (final encoder LayerNorm applied)."""
        cfg = self.cfg
        B, S = enc_mask.shape
        p = cfg.dropout if train else 0.0
        x, recs = self._embed(enc_inputs, S, "enc", B)
        Me = B * S
        mem = self.buf("enc.mem", (Me, cfg.d_model), self.adt)
        h = None  # LN of x for the next block, when the previous block's fused product already produced it
        if not self.norm_first:
            xb = self._operand_copy(x, "enc.x0b")
            for i in range(cfg.encoder_layers):
                pre = f"hf_model.encoder.layers.{i}."
                tg = f"enc{i if train else ''}"
                x, xb = self._attn_block_fwd(tg + ".sa", x, B, S, cfg.encoder_attention_heads,
                                             self._wp_attn(pre, "self_attn", "norm1"), enc_mask, False, p,
                                             self._site(False, i, 0), self._site(False, i, 1), train, xb=xb)
                x, xb = self._ffn_block_fwd(tg + ".ff", x, Me, cfg.encoder_ffn_dim, self._wp_ffn(pre, "norm2"), p,
                                            self._site(False, i, 2), self._site(False, i, 3), train, xb=xb)
            ops.ln_fwd(x, self.P("hf_model.encoder.norm.weight"), self.P("hf_model.encoder.norm.bias"), mem)
            if train:
                self.saved["enc"] = dict(recs=recs, xL=x, mem=mem, B=B, S=S, mask=enc_mask)
            return mem
        for i in range(cfg.encoder_layers):
            pre = f"hf_model.encoder.layers.{i}."
            tg = f"enc{i if train else ''}"
            x, h = self._attn_block_fwd(tg + ".sa", x, B, S, cfg.encoder_attention_heads,
                                        self._wp_attn(pre, "self_attn", "norm1"), enc_mask, False, p,
                                        self._site(False, i, 0), self._site(False, i, 1), train, h_pre=h,
                                        nxt=self._next_norm(pre + "norm2.weight", pre + "norm2.bias", tg + ".ff", Me))
            if i + 1 < cfg.encoder_layers:
                npre = f"hf_model.encoder.layers.{i + 1}."
                nxt = self._next_norm(npre + "norm1.weight", npre + "norm1.bias", f"enc{i + 1 if train else ''}.sa", Me)
            else:
                nxt = (self.P("hf_model.encoder.norm.weight"), self.P("hf_model.encoder.norm.bias"), mem)
            x, h = self._ffn_block_fwd(tg + ".ff", x, Me, cfg.encoder_ffn_dim, self._wp_ffn(pre, "norm2"), p,
                                       self._site(False, i, 2), self._site(False, i, 3), train, h_pre=h, nxt=nxt)
        if h is None:
            ops.ln_fwd(x, self.P("hf_model.encoder.norm.weight"), self.P("hf_model.encoder.norm.bias"), mem)
        if train:
            self.saved["enc"] = dict(recs=recs, xL=x, mem=mem, B=B, S=S, mask=enc_mask)
        return mem

    def decode_teacher_forced(self, dec_ids, dec_mask, mem, enc_mask, train=False):
        """-> final decoder hidden states [B*T, d] (activation dtype)."""
        cfg = self.cfg
        B, T_ = dec_ids.shape
        S = enc_mask.shape[1]
        p = cfg.dropout if train else 0.0
        x, recs = self._embed({cfg.target_modality: dec_ids}, T_, "dec", B)
        H = cfg.decoder_attention_heads
        M = B * T_
        hT = self.buf("dec.hT", (M, cfg.d_model), self.adt)
        h = None
        if not self.norm_first:
            xb = self._operand_copy(x, "dec.x0b")
            for i in range(cfg.decoder_layers):
                pre = f"hf_model.decoder.layers.{i}."
                tg = f"dec{i if train else ''}"
                x, xb = self._attn_block_fwd(tg + ".sa", x, B, T_, H, self._wp_attn(pre, "self_attn", "norm1"), dec_mask,
                                             True, p, self._site(True, i, 0), self._site(True, i, 1), train, xb=xb)
                x, xb = self._cross_block_fwd(tg + ".ca", x, mem, B, T_, S, H,
                                              self._wp_attn(pre, "multihead_attn", "norm2"), enc_mask, p,
                                              self._site(True, i, 4), self._site(True, i, 5), train, xb=xb)
                x, xb = self._ffn_block_fwd(tg + ".ff", x, M, cfg.decoder_ffn_dim, self._wp_ffn(pre, "norm3"), p,
                                            self._site(True, i, 2), self._site(True, i, 3), train, xb=xb)
            ops.ln_fwd(x, self.P("hf_model.decoder.norm.weight"), self.P("hf_model.decoder.norm.bias"), hT)
            if train:
                self.saved["dec"] = dict(recs=recs, xL=x, hT=hT, B=B, T=T_, mask=dec_mask)
            return hT
        for i in range(cfg.decoder_layers):
            pre = f"hf_model.decoder.layers.{i}."
            tg = f"dec{i if train else ''}"
            x, h = self._attn_block_fwd(tg + ".sa", x, B, T_, H, self._wp_attn(pre, "self_attn", "norm1"), dec_mask,
                                        True, p, self._site(True, i, 0), self._site(True, i, 1), train, h_pre=h,
                                        nxt=self._next_norm(pre + "norm2.weight", pre + "norm2.bias", tg + ".ca", M))
            x, h = self._cross_block_fwd(tg + ".ca", x, mem, B, T_, S, H, self._wp_attn(pre, "multihead_attn", "norm2"),
                                         enc_mask, p, self._site(True, i, 4), self._site(True, i, 5), train, h_pre=h,
                                         nxt=self._next_norm(pre + "norm3.weight", pre + "norm3.bias", tg + ".ff", M))
            if i + 1 < cfg.decoder_layers:
                npre = f"hf_model.decoder.layers.{i + 1}."
                nxt = self._next_norm(npre + "norm1.weight", npre + "norm1.bias", f"dec{i + 1 if train else ''}.sa", M)
            else:
                nxt = (self.P("hf_model.decoder.norm.weight"), self.P("hf_model.decoder.norm.bias"), hT)
            x, h = self._ffn_block_fwd(tg + ".ff", x, M, cfg.decoder_ffn_dim, self._wp_ffn(pre, "norm3"), p,
                                       self._site(True, i, 2), self._site(True, i, 3), train, h_pre=h, nxt=nxt)
        if h is None:
            ops.ln_fwd(x, self.P("hf_model.decoder.norm.weight"), self.P("hf_model.decoder.norm.bias"), hT)
        if train:
            self.saved["dec"] = dict(recs=recs, xL=x, hT=hT, B=B, T=T_, mask=dec_mask)
        return hT

    # ------------------------------------------------------------------------------------ forward
    def forward(self, enc_inputs, enc_mask, dec_ids, dec_mask, labels=None, train=False, align_target=None):
        """enc_mask/dec_mask: uint8 [B, L], 1 = real token.  labels: int64 [B, T], -100 = ignore.
        align_target: fp32 [B, output_dimension] (models with an align head, custom_modeling.py:453-475).
        Returns dict(logits=[B, T, V] fp32 view, loss=device scalar or None, lm_loss, align_loss)."""
        cfg = self.cfg
        self.sync_weights()
        if train:
            self.saved = {}
            self._wq = []
        B, T_ = dec_ids.shape
        mem = self.encode(enc_inputs, enc_mask, train)
        hT = self.decode_teacher_forced(dec_ids, dec_mask, mem, enc_mask, train)
        M = B * T_
        V = cfg.vocab_size
        logits = self.buf("logits", (M, self.ldv), torch.float32)
        ops.gemm(hT, self.W("hf_model.token_ff.weight"), M, V, cfg.d_model,
                 ops.make_epi(EPI_STORE, logits, bias=self.P("hf_model.token_ff.bias")))
        out = {"logits": logits[:, :V].view(B, T_, V), "loss": None, "memory": mem}
        if labels is not None:
            labels = labels.contiguous().view(-1)
            row_loss = self.buf("ce.row_loss", (M,), torch.float32)
            row_lse = self.buf("ce.row_lse", (M,), torch.float32)
            stats = self.buf("ce.stats", (2,), torch.float32)
            ops.ce_fwd(logits, labels, V, row_loss, row_lse, stats, smoothing=cfg.label_smoothing)
            out["loss"] = out["lm_loss"] = stats[0]
            out["align_loss"] = None
            if train:
                self.saved["ce"] = dict(logits=logits, labels=labels, row_lse=row_lse, stats=stats, M=M)
            if cfg.align_config and align_target is not None:
                al = self._align_fwd(mem, enc_mask, align_target, stats[0:1], train)
                out["align_loss"], out["loss"] = al[0], al[1]
        return out

    # ------------------------------------------------------------------------------- align head
    def _align_layers(self):
        """[(weight view [out, in], bias, grad-weight view, grad bias, relu?)] of the align network; the Conv1d layers
        act on a length-1 sequence, so only the centre tap (k // 2) of `4.weight` contributes."""
        ac, p = self.cfg.align_config, "hf_model.align_network."
        names = [("0", True)]
        names += [("2", False), ("4", True), ("6", False)] if ac["align_network"] == "convolutional" else [("2", False)]
        out = []
        for n, relu in names:
            W, G = self.P(f"{p}{n}.weight"), self.G(f"{p}{n}.weight")
            if W.dim() == 3:
                c = W.shape[2] // 2
                W, G = W[:, :, c], G[:, :, c]
            out.append((W, self.P(f"{p}{n}.bias"), G, self.G(f"{p}{n}.bias"), relu))
        return out

    def _align_fwd(self, mem, enc_mask, target, lm_loss, train):
        ac = self.cfg.align_config
        if ac["align_network"] not in ("convolutional", "mlp"):
            raise ValueError(f"unknown align network {ac['align_network']}")
        if ac["loss_function"] not in ops.ALIGN_LOSS_KINDS:
            raise ValueError(f"Loss function {ac['loss_function']} not supported for alignment!")
        B, S = enc_mask.shape
        d = self.cfg.d_model
        pooled = self.buf("al.pooled", (B, d), torch.float32)
        ops.masked_mean_fwd(mem, enc_mask, pooled, B, S)
        acts = [pooled]
        h = pooled
        layers = self._align_layers()
        for li, (W, b, _, _, relu) in enumerate(layers):
            o, i = W.shape
            y = self.buf(f"al.h{li}", (B, o), torch.float32)
            ops.gemm(h, W, B, o, i, ops.make_epi(EPI_RELU if relu else EPI_STORE, y, bias=b), force_simt=True)
            acts.append(y)
            h = y
        target = target.contiguous().float()
        res = self.buf("al.loss", (2,), torch.float32)
        ops.align_loss(h, target, ac["loss_function"], ac["loss_lambda"], lm_loss, res)
        if train:
            self.saved["align"] = dict(acts=acts, target=target, mask=enc_mask, B=B, S=S)
        return res

    def _align_bwd(self, dmem, gscale):
        """Weight / bias gradients of the align network and its contribution to d(memory)."""
        s = self.saved["align"]
        acts, B = s["acts"], s["B"]
        ac = self.cfg.align_config
        z = acts[-1]
        dy = self.buf("al.dz", tuple(z.shape), torch.float32)
        ops.align_loss(z, s["target"], ac["loss_function"], ac["loss_lambda"], None, self.buf("al.loss_bw", (2,), torch.float32),
                       dz=dy, dscale=gscale)
        layers = self._align_layers()
        for li in reversed(range(len(layers))):
            W, _, G, Gb, relu = layers[li]
            o, i = W.shape
            x = acts[li]
            if G.stride(1) == 1:
                ops.gemm(dy, x, o, i, B, ops.make_epi(EPI_ACCUM, G, accumulate=1), a_mn=True, b_mn=True, force_simt=True)
            else:  # centre tap of a Conv1d weight: accumulate through a contiguous scratch
                tmp = self.buf(f"al.dw{li}", (o, i), torch.float32)
                ops.gemm(dy, x, o, i, B, ops.make_epi(EPI_ACCUM, tmp, accumulate=0), a_mn=True, b_mn=True, force_simt=True)
                ops.add_strided(G.as_strided((o * i,), (G.stride(1),), G.storage_offset()), tmp.view(-1))
            ops.colsum(dy, Gb, rows=B, cols=o)
            dx = self.buf(f"al.dx{li}", (B, i), torch.float32)
            prev_relu = li > 0 and layers[li - 1][4]
            epi = ops.make_epi(EPI_DRELU, dx, aux=x) if prev_relu else ops.make_epi(EPI_STORE, dx)
            ops.gemm(dy, W, B, i, o, epi, b_mn=True, force_simt=True)
            dy = dx
        ops.masked_mean_bwd(dy, s["mask"], dmem, B, s["S"])

    # ----------------------------------------------------------------------------------- backward
    def backward(self, gscale: float = 1.0):
        """Accumulates d(loss * gscale)/d(param) into the flat gradient buffer (ParamStore.g)."""
        cfg, ps = self.cfg, self.ps
        d, V = cfg.d_model, cfg.vocab_size
        T = self.adt
        p = cfg.dropout
        ce, dec, enc = self.saved["ce"], self.saved["dec"], self.saved["enc"]
        B, T_, S = dec["B"], dec["T"], enc["S"]
        M, Me = B * T_, B * S
        H = cfg.decoder_attention_heads
        notify = self.grad_ready_hook or (lambda off: None)

        self._wev.clear()
        self._begin_group(0)
        dlogits = self.wbuf("bw.dlogits", (M, self.ldv), T)
        ops.ce_bwd(ce["logits"], ce["labels"], V, ce["row_lse"], ce["stats"], dlogits, gscale=gscale,
                   smoothing=cfg.label_smoothing)
        dl = dlogits[:, :V]
        dhT = self.buf("bw.dh", (M, d), T)
        self._lin_bwd(dl, dec["hT"], "hf_model.token_ff.weight", "hf_model.token_ff.bias", M, V, d,
                      dx_epi=ops.make_epi(EPI_STORE, dhT))
        self._flush_wgrads()
        last_site = self._site(True, cfg.decoder_layers - 1, 3)
        dx = self._other_dx(None, M)
        # the masked low-precision gradient handed to the layer below lives in the buffer set of the group that reads
        # it (it must stay alive until that group's weight-gradient launch is done)
        dyb = self.wbuf("bw.dyb.x", (M, d), T, ahead=1)
        ops.ln_bwd(dhT, dec["xL"], self.P("hf_model.decoder.norm.weight"), dx=dx, dxb=dyb,
                   dgamma=self.G("hf_model.decoder.norm.weight"), dbeta=self.G("hf_model.decoder.norm.bias"),
                   p_drop=p, seed=self.seed_arg, site=last_site)
        notify(ps.offsets["hf_model.decoder.norm.weight"][0])

        dmem = self.buf("bw.dmem", (Me, d), torch.float32)
        if not self.norm_first:
            return self._backward_post_layers(dx, dmem, gscale, notify)
        for i in reversed(range(cfg.decoder_layers)):
            self._begin_group(cfg.decoder_layers - i)
            pre = f"hf_model.decoder.layers.{i}."
            tg = f"dec{i}"
            dx, dyb = self._ffn_block_bwd(tg + ".ff", dx, dyb, M, cfg.decoder_ffn_dim, self._wp_ffn(pre, "norm3"), p,
                                          self._site(True, i, 2), prev_site=self._site(True, i, 5))
            dx, dyb = self._cross_block_bwd(tg + ".ca", dx, dyb, enc["mem"], dmem, i == cfg.decoder_layers - 1, B, T_,
                                            S, H, self._wp_attn(pre, "multihead_attn", "norm2"), enc["mask"], p,
                                            self._site(True, i, 4), prev_site=self._site(True, i, 1))
            prev = self._site(True, i - 1, 3) if i > 0 else self._site(True, 0, 15)
            dx, dyb = self._attn_block_bwd(tg + ".sa", dx, dyb, B, T_, H, self._wp_attn(pre, "self_attn", "norm1"),
                                           dec["mask"], True, p, self._site(True, i, 0), prev_site=prev, first=(i == 0),
                                           out_tag=f"x{i % 2}")
            self._flush_wgrads()
            notify(ps.offsets[pre + "self_attn.in_proj_weight"][0])
        self._embed_bwd(dx, dec["recs"], T_, "dec", B)
        if "align" in self.saved:
            self._align_bwd(dmem, gscale)

        # encoder: d(mem) arrives in fp32 from the cross-attention K/V projections
        He = cfg.encoder_attention_heads
        dxe = self._other_dx(None, Me)
        dybe = self.wbuf("bw.dyb.x", (Me, d), T, ahead=1)
        ops.ln_bwd(dmem, enc["xL"], self.P("hf_model.encoder.norm.weight"), dx=dxe, dxb=dybe,
                   dgamma=self.G("hf_model.encoder.norm.weight"), dbeta=self.G("hf_model.encoder.norm.bias"),
                   p_drop=p, seed=self.seed_arg, site=self._site(False, cfg.encoder_layers - 1, 3))
        notify(ps.offsets["hf_model.encoder.norm.weight"][0])
        for i in reversed(range(cfg.encoder_layers)):
            self._begin_group(cfg.decoder_layers + cfg.encoder_layers - i)
            pre = f"hf_model.encoder.layers.{i}."
            tg = f"enc{i}"
            dxe, dybe = self._ffn_block_bwd(tg + ".ff", dxe, dybe, Me, cfg.encoder_ffn_dim, self._wp_ffn(pre, "norm2"),
                                            p, self._site(False, i, 2), prev_site=self._site(False, i, 1))
            prev = self._site(False, i - 1, 3) if i > 0 else self._site(False, 0, 15)
            dxe, dybe = self._attn_block_bwd(tg + ".sa", dxe, dybe, B, S, He, self._wp_attn(pre, "self_attn", "norm1"),
                                             enc["mask"], False, p, self._site(False, i, 0), prev_site=prev,
                                             first=(i == 0), out_tag=f"x{i % 2}")
            self._flush_wgrads()
            notify(ps.offsets[pre + "self_attn.in_proj_weight"][0])
        self._embed_bwd(dxe, enc["recs"], S, "enc", B)
        notify(0)
        self._join_wgrads()

    def _backward_post_layers(self, dx, dmem, gscale, notify):
        """Layer loops of `backward` for LN(x + f(x)) layers (post_layer_normalisation=False).  dx: fp32 gradient of the
        last decoder block's output (the input of decoder.norm)."""
        cfg, ps = self.cfg, self.ps
        d, p = cfg.d_model, cfg.dropout
        dec, enc = self.saved["dec"], self.saved["enc"]
        B, T_, S = dec["B"], dec["T"], enc["S"]
        M, Me = B * T_, B * S
        H, He = cfg.decoder_attention_heads, cfg.encoder_attention_heads
        for i in reversed(range(cfg.decoder_layers)):
            self._begin_group(cfg.decoder_layers - i)
            pre = f"hf_model.decoder.layers.{i}."
            tg = f"dec{i}"
            dx = self._ffn_block_bwd_post(tg + ".ff", dx, M, cfg.decoder_ffn_dim, self._wp_ffn(pre, "norm3"), p,
                                          self._site(True, i, 2), self._site(True, i, 3))
            dx = self._cross_block_bwd_post(tg + ".ca", dx, enc["mem"], dmem, i == cfg.decoder_layers - 1, B, T_, S, H,
                                            self._wp_attn(pre, "multihead_attn", "norm2"), enc["mask"], p,
                                            self._site(True, i, 4), self._site(True, i, 5))
            dx = self._attn_block_bwd_post(tg + ".sa", dx, B, T_, H, self._wp_attn(pre, "self_attn", "norm1"),
                                           dec["mask"], True, p, self._site(True, i, 0), self._site(True, i, 1))
            self._flush_wgrads()
            notify(ps.offsets[pre + "self_attn.in_proj_weight"][0])
        self._embed_bwd(dx, dec["recs"], T_, "dec", B)
        if "align" in self.saved:
            self._align_bwd(dmem, gscale)
        dxe = self._other_dx(None, Me)
        ops.ln_bwd(dmem, enc["xL"], self.P("hf_model.encoder.norm.weight"), dx=dxe,
                   dgamma=self.G("hf_model.encoder.norm.weight"), dbeta=self.G("hf_model.encoder.norm.bias"))
        notify(ps.offsets["hf_model.encoder.norm.weight"][0])
        for i in reversed(range(cfg.encoder_layers)):
            self._begin_group(cfg.decoder_layers + cfg.encoder_layers - i)
            pre = f"hf_model.encoder.layers.{i}."
            tg = f"enc{i}"
            dxe = self._ffn_block_bwd_post(tg + ".ff", dxe, Me, cfg.encoder_ffn_dim, self._wp_ffn(pre, "norm2"), p,
                                           self._site(False, i, 2), self._site(False, i, 3))
            dxe = self._attn_block_bwd_post(tg + ".sa", dxe, B, S, He, self._wp_attn(pre, "self_attn", "norm1"),
                                            enc["mask"], False, p, self._site(False, i, 0), self._site(False, i, 1))
            self._flush_wgrads()
            notify(ps.offsets[pre + "self_attn.in_proj_weight"][0])
        self._embed_bwd(dxe, enc["recs"], S, "enc", B)
        notify(0)
        self._join_wgrads()
