"""Model configuration and the flat parameter store.

All trainable tensors live in ONE contiguous fp32 master buffer (plus a same-shaped gradient buffer, Adam
moments, and a bf16 mirror the tcgen05 GEMMs read), laid out in forward order so that gradient buckets
complete back-to-front during backward and can be all-reduced while earlier layers are still running.
Names and shapes follow the reference checkpoint layout (`HFWrapper.state_dict()` keys
`hf_model.*`, with the shared embedding also visible as `hf_model.decoder.embedding.*` and
`multimodal_embedding.*`, custom_modeling.py:347,409-415, wrapper.py:298).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import torch

TOKEN_TYPES = ("text", "text_spectrum", "peak_positional_encoding", "run_length_encoding", "multiplets", "carbon",
               "msms_text")
PATCH_TYPES = ("1D_patches", "msms_number")
ALIGN = 128  # elements; keeps every tensor 256-byte aligned in the bf16 mirror (TMA needs 16)


@dataclass
class ModelConfig:
    """The subset of `CustomConfig` (custom_modeling.py:40-105) + wrapper arguments the hot path needs."""

    data_config: Dict[str, Any]
    vocab_size: int
    d_model: int = 512
    encoder_layers: int = 6
    decoder_layers: int = 6
    encoder_attention_heads: int = 8
    decoder_attention_heads: int = 8
    encoder_ffn_dim: int = 2048
    decoder_ffn_dim: int = 2048
    dropout: float = 0.1
    gated_linear: bool = False
    # custom_modeling.py:119-129,166-176: handed to torch as `norm_first` - True (every shipped yaml): x + f(LN(x));
    # False: LN(x + f(x))
    post_layer_normalisation: bool = True
    positional_encoding_type: str = "sin_cos"
    multimodal_norm: bool = True
    max_position_embeddings: int = 1024
    pad_token_id: int = 0
    bos_token_id: int = 2
    eos_token_id: int = 3
    max_length: int = 128
    align_config: Optional[Dict[str, Any]] = None
    label_smoothing: float = 0.0
    target_modality: str = field(default="")

    def __post_init__(self):
        targets = [m for m, c in self.data_config.items() if c.get("target") and not c.get("alignment")]
        if len(targets) != 1:
            raise ValueError("Only 1 target modality can be specified.")  # data/datamodules.py:57-60
        self.target_modality = targets[0]
        for m, c in self.data_config.items():
            if c["type"] not in TOKEN_TYPES + PATCH_TYPES and not c.get("alignment"):
                raise NotImplementedError(f"Unknown modality type: {c['type']}")  # modeling/utils.py:137-138
        if self.d_model % self.encoder_attention_heads or self.d_model % self.decoder_attention_heads:
            raise ValueError("d_model must be divisible by the number of heads")
        if self.positional_encoding_type not in ("sin_cos", "learned"):
            raise KeyError(self.positional_encoding_type)

    def embed_layers(self, modality) -> List[Tuple[int, int]]:
        """[(out_features, in_features), ...] of a patch modality's Linear stack (utils.py:107-134)."""
        mc = self.data_config[modality]
        d = self.d_model
        ps = 2 if mc["type"] == "msms_number" else mc["preprocessor_arguments"]["patch_size"]
        et = (mc.get("preprocessor_arguments") or {}).get("encoding_type", "linear")
        if et == "linear":
            return [(d, ps)]
        if et == "linear_2_layer":
            return [(d // 2, ps), (d, d // 2)]
        if et == "linear_3_layer":
            return [(d // 3, ps), (2 * (d // 3), d // 3), (d, 2 * (d // 3))]
        raise NotImplementedError(et)


def sincos_table(d_model: int, max_len: int) -> torch.Tensor:
    """Interleaved sin/cos rows, w_i = 10000^(2i/d) (modeling/utils.py:226-239); evaluated position by position
    with a python-int numerator exactly as the reference builds its buffer so checkpoints compare bit-equal."""
    w = 10000 ** torch.tensor([dim / d_model for dim in range(0, d_model, 2)])
    out = torch.empty(max_len, 2 * w.numel())
    for p in range(max_len):
        a = p / w
        out[p, 0::2] = torch.sin(a)
        out[p, 1::2] = torch.cos(a)
    return out[:, :d_model].contiguous()


def param_specs(cfg: ModelConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(canonical name, shape, init kind) in forward order."""
    d = cfg.d_model
    out: List[Tuple[str, Tuple[int, ...], str]] = []
    e = "hf_model.embedding."
    for m, mc in cfg.data_config.items():
        # an alignment-target modality (`alignment: true`) also gets its layer + norm: the reference builds one for every
        # data_config entry (modeling/utils.py:73-77) and its checkpoints carry them, although no forward pass uses them
        base = f"{e}embedding_layer_dict.{m}."
        if mc["type"] in TOKEN_TYPES:
            out.append((base + "weight", (mc["vocab_size"], d), "xavier"))
        else:
            layers = cfg.embed_layers(m)
            for li, (o, i) in enumerate(layers):
                pre = base if len(layers) == 1 else f"{base}{2 * li}."
                out.append((pre + "weight", (o, i), "xavier"))
                out.append((pre + "bias", (o,), f"lin_bias:{i}"))
        if cfg.multimodal_norm:
            out.append((f"{e}embedding_norm_dict.{m}.weight", (d,), "ones"))
            out.append((f"{e}embedding_norm_dict.{m}.bias", (d,), "zeros"))
    if cfg.positional_encoding_type == "learned":
        out.append((e + "positional_encodings.pos_encodings.weight", (cfg.max_position_embeddings, d), "xavier"))
        out.append((e + "positional_encodings.norm.weight", (d,), "ones"))
        out.append((e + "positional_encodings.norm.bias", (d,), "zeros"))

    def attn(p):
        out.append((p + "in_proj_weight", (3 * d, d), "xavier"))
        out.append((p + "in_proj_bias", (3 * d,), "zeros"))
        out.append((p + "out_proj.weight", (d, d), "xavier"))
        out.append((p + "out_proj.bias", (d,), "zeros"))

    def ffn(p, f):
        out.append((p + "linear1.weight", (f, d), "xavier"))
        out.append((p + "linear1.bias", (f,), f"lin_bias:{d}"))
        if cfg.gated_linear:
            out.append((p + "gate.weight", (f, d), "xavier"))
            out.append((p + "gate.bias", (f,), f"lin_bias:{d}"))
        out.append((p + "linear2.weight", (d, f), "xavier"))
        out.append((p + "linear2.bias", (d,), f"lin_bias:{f}"))

    def norm(p):
        out.append((p + "weight", (d,), "ones"))
        out.append((p + "bias", (d,), "zeros"))

    for i in range(cfg.encoder_layers):
        p = f"hf_model.encoder.layers.{i}."
        attn(p + "self_attn.")
        ffn(p, cfg.encoder_ffn_dim)
        norm(p + "norm1.")
        norm(p + "norm2.")
    norm("hf_model.encoder.norm.")
    if cfg.align_config:
        ac = cfg.align_config
        hd = ac["hidden_dimension"]
        p = "hf_model.align_network."
        out.append((p + "0.weight", (hd, d), "xavier"))
        out.append((p + "0.bias", (hd,), f"lin_bias:{d}"))
        if ac["align_network"] == "convolutional":
            cc, ks, od = ac["conv_channels"], ac["kernel_size"], ac["output_dimension"]
            out.append((p + "2.weight", (hd, hd), "xavier"))
            out.append((p + "2.bias", (hd,), f"lin_bias:{hd}"))
            out.append((p + "4.weight", (cc, hd, ks), "xavier"))
            out.append((p + "4.bias", (cc,), f"lin_bias:{hd * ks}"))
            out.append((p + "6.weight", (od, cc, 1), "xavier"))
            out.append((p + "6.bias", (od,), f"lin_bias:{cc}"))
        elif ac["align_network"] == "mlp":
            out.append((p + "2.weight", (ac["output_dimension"], hd), "xavier"))
            out.append((p + "2.bias", (ac["output_dimension"],), f"lin_bias:{hd}"))
    for i in range(cfg.decoder_layers):
        p = f"hf_model.decoder.layers.{i}."
        attn(p + "self_attn.")
        attn(p + "multihead_attn.")
        ffn(p, cfg.decoder_ffn_dim)
        norm(p + "norm1.")
        norm(p + "norm2.")
        norm(p + "norm3.")
    norm("hf_model.decoder.norm.")
    out.append(("hf_model.token_ff.weight", (cfg.vocab_size, d), "xavier"))
    out.append(("hf_model.token_ff.bias", (cfg.vocab_size,), f"lin_bias:{d}"))
    return out


class ParamStore:
    """Flat fp32 master / gradient / Adam-moment buffers + bf16 mirror, with named views."""

    EMB = "hf_model.embedding."
    ALIASES = ("hf_model.decoder.embedding.", "multimodal_embedding.")

    def __init__(self, cfg: ModelConfig, device="cuda", seed: Optional[int] = None):
        self.cfg = cfg
        self.device = torch.device(device)
        self.specs = param_specs(cfg)
        self.offsets: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        off = 0
        for name, shape, _ in self.specs:
            self.offsets[name] = (off, shape)
            n = int(math.prod(shape))
            off += (n + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        self.p = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.g = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.pb = torch.zeros(off, dtype=torch.bfloat16, device=self.device)
        self.m: Optional[torch.Tensor] = None
        self.v: Optional[torch.Tensor] = None
        self.buffers: Dict[str, torch.Tensor] = {}
        if cfg.positional_encoding_type == "sin_cos":
            self.buffers[self.EMB + "positional_encodings.pos_enc"] = sincos_table(
                cfg.d_model, cfg.max_position_embeddings).to(self.device)
        self._views: Dict[Tuple[str, str], torch.Tensor] = {}
        self.bf16_dirty = True
        self.g_dirty = False  # the gradient buffer holds leftovers of a backward run outside FusedTrainer
        self.shard: Optional[Tuple[int, int]] = None  # [lo, hi) of the master weights this rank's optimiser owns
        self.init_parameters(seed)

    # -- views ---------------------------------------------------------------------------------
    def _view(self, buf, kind, name):
        key = (kind, name)
        v = self._views.get(key)
        if v is None:
            off, shape = self.offsets[name]
            v = buf[off: off + int(math.prod(shape))].view(shape)
            self._views[key] = v
        return v

    def P(self, name):
        return self._view(self.p, "p", name)

    def G(self, name):
        return self._view(self.g, "g", name)

    def PB(self, name):
        return self._view(self.pb, "pb", name)

    def has(self, name):
        return name in self.offsets

    def rebind(self, g: Optional[torch.Tensor] = None, pb: Optional[torch.Tensor] = None):
        """Move the gradient buffer / bf16 mirror into caller-provided storage of the same size (symmetric memory that
        peer GPUs can address: trainer.PeerShardedStep).  Contents are carried over; cached views are rebuilt."""
        for kind, new in (("g", g), ("pb", pb)):
            if new is None:
                continue
            old = getattr(self, kind)
            if new.numel() != old.numel() or new.dtype != old.dtype:
                raise ValueError(f"rebind({kind}): expected {old.numel()} x {old.dtype}")
            new.copy_(old)
            setattr(self, kind, new)
            for key in [k for k in self._views if k[0] == kind]:
                del self._views[key]

    def gather_master(self, process_group=None):
        """Reassemble the fp32 master weights (and Adam moments) on every rank after rank-sharded optimiser steps
        (`self.shard` = this rank's [lo, hi)): checkpoints / state_dict() need the full tensors."""
        if self.shard is None:
            return
        lo, hi = self.shard
        for buf in (self.p, self.m, self.v):
            if buf is None:
                continue
            keep = buf[lo:hi].clone()
            buf.zero_()
            buf[lo:hi] = keep
            torch.distributed.all_reduce(buf, group=process_group)

    def ensure_optimizer_state(self):
        if self.m is None:
            self.m = torch.zeros_like(self.p)
            self.v = torch.zeros_like(self.p)

    # -- init (wrapper.py:320-327: xavier_uniform_ on every parameter with dim > 1) --------------
    def init_parameters(self, seed: Optional[int] = None):
        g = torch.Generator().manual_seed(3247 if seed is None else seed)
        host = torch.zeros(self.numel, dtype=torch.float32)
        for name, shape, kind in self.specs:
            off, _ = self.offsets[name]
            n = int(math.prod(shape))
            if kind == "xavier":
                rf = int(math.prod(shape[2:])) if len(shape) > 2 else 1
                fan_in, fan_out = shape[1] * rf, shape[0] * rf
                a = math.sqrt(6.0 / (fan_in + fan_out))
                t = (torch.rand(n, generator=g) * 2 - 1) * a
            elif kind == "ones":
                t = torch.ones(n)
            elif kind == "zeros":
                t = torch.zeros(n)
            else:  # lin_bias:<fan_in>  (torch Linear / Conv default)
                b = 1.0 / math.sqrt(int(kind.split(":")[1]))
                t = (torch.rand(n, generator=g) * 2 - 1) * b
            host[off: off + n] = t
        self.p.copy_(host)
        self.bf16_dirty = True

    # -- checkpoint layout ------------------------------------------------------------------------
    def state_dict(self, with_aliases=True) -> Dict[str, torch.Tensor]:
        sd = {name: self.P(name) for name, _, _ in self.specs}
        sd.update(self.buffers)
        if with_aliases:
            for k in [k for k in sd if k.startswith(self.EMB)]:
                for a in self.ALIASES:
                    sd[a + k[len(self.EMB):]] = sd[k]
        return sd

    def canonical(self, key: str) -> str:
        for a in self.ALIASES:
            if key.startswith(a):
                return self.EMB + key[len(a):]
        return key

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict=True):
        seen = set()
        unexpected = []
        for k, v in sd.items():
            c = self.canonical(k)
            if c in self.offsets:
                tgt = self.P(c)
                if tuple(v.shape) != tuple(tgt.shape):
                    raise RuntimeError(f"size mismatch for {k}: {tuple(v.shape)} vs {tuple(tgt.shape)}")
                tgt.copy_(v.to(torch.float32))
                seen.add(c)
            elif c in self.buffers:
                self.buffers[c].copy_(v.to(torch.float32))
                seen.add(c)
            else:
                unexpected.append(k)
        missing = [n for n in self.offsets if n not in seen]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing={missing[:8]} unexpected={unexpected[:8]}")
        self.bf16_dirty = True
        return missing, unexpected
