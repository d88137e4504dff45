"""CPU oracle for the spectra -> SMILES hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch CPU restatement (plain torch-CPU tensor algebra, fp32) of the algorithm
that `rxn4chemistry/MultimodalAnalytical` runs on its hot path.  It exists so that the CUDA path can
be checked against something that does not need `/root/reference` at run time.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.
The product package (`multimodalanalytical_b200/`) never does.

Parity status: PINNED.  The reference's own tests hold no numerical vectors, so the
oracle is pinned against outputs of the imported, unmodified reference modules run in the build
container (`tests/golden/make_golden.py` -> `tests/golden/*.pt`); `tests/test_oracle_golden.py`
replays them.

Reference anchors (paths relative to /root/reference/src/analytical_fm):
  * modeling/utils.py:44-182      MultimodalEmbedding      -> embed_modalities
  * modeling/utils.py:198-239     SincCosPositionalEncoding -> sincos_table
  * modeling/utils.py:242-272     LearnedPositionalEncoding -> embed_modalities (learned branch)
  * modeling/custom_modeling.py:108-199  encoder/decoder layers (+GLU)  -> encoder_stack / decoder_stack
  * modeling/custom_modeling.py:220-243,271-320  mask conventions   -> encoder_stack / decoder_stack
  * modeling/custom_modeling.py:420-508  CustomModel.forward (align head, LM head, CE) -> model_forward
  * modeling/wrapper.py:346-407   HFWrapper.forward (batch re-layout)   -> wrapper_forward
  * modeling/wrapper.py:409-453   HFWrapper.generate                    -> generate
  * third-party: torch `nn.TransformerEncoderLayer/DecoderLayer` norm_first branch and
    `F.multi_head_attention_forward` (pinned torch==2.5.1, requirements.txt:69); transformers
    `GenerationMixin._sample/_beam_search` (pinned 4.47.0, requirements.txt:72; the container runs
    5.5.0 whose vectorised `_beam_search` is what the golden vectors come from).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import torch
import torch.nn.functional as F

TOKEN_TYPES = (
    "text",
    "text_spectrum",
    "peak_positional_encoding",
    "run_length_encoding",
    "multiplets",
    "carbon",
    "msms_text",
)
PATCH_TYPES = ("1D_patches", "msms_number")


@dataclass
class OracleConfig:
    d_model: int = 512
    encoder_layers: int = 6
    decoder_layers: int = 6
    encoder_attention_heads: int = 8
    decoder_attention_heads: int = 8
    gated_linear: bool = False
    post_layer_normalisation: bool = True  # the reference hands this to torch as `norm_first` (custom_modeling.py:129,176)
    positional_encoding_type: str = "sin_cos"
    multimodal_norm: bool = True
    pad_token_id: int = 0
    bos_token_id: int = 2
    eos_token_id: int = 3
    max_length: int = 128
    align_config: Optional[Dict[str, Any]] = None
    target_modality: str = "Smiles"
    data_config: Dict[str, Any] = field(default_factory=dict)
    dropout: float = 0.0       # reference default 0.1 (bart-base config); only applied when `training`
    training: bool = False


# --------------------------------------------------------------------------------------------
# embedding (modeling/utils.py:44-182, 198-272)
# --------------------------------------------------------------------------------------------
def sincos_table(d_model: int, max_len: int = 1024) -> torch.Tensor:
    """Interleaved [sin(p/w0), cos(p/w0), sin(p/w1), ...], w_i = 10000^(2i/d)  (utils.py:226-239)."""
    w = 10000 ** torch.tensor([dim / d_model for dim in range(0, d_model, 2)])
    rows = []
    for p in range(max_len):
        ang = p / w  # python-int / tensor, as the reference does it: keeps the buffer bit-identical
        rows.append(torch.stack((torch.sin(ang), torch.cos(ang)), dim=1).reshape(-1))
    tab = torch.stack(rows)
    return tab[:, :d_model].contiguous()


def _ln(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def _embed_one(sd, emb_prefix, modality, mcfg, value):
    """One modality -> [B, S_m, d] before LayerNorm (utils.py:84-140, 152-162)."""
    base = f"{emb_prefix}embedding_layer_dict.{modality}."
    if mcfg["type"] in TOKEN_TYPES:
        table = sd[base + "weight"]
        if isinstance(value, dict):  # XVal: embedding scaled by the numeric value (utils.py:154-160)
            return table[value["tokenized_input"]] * value["numerical_values"].unsqueeze(-1)
        return table[value]
    if mcfg["type"] in PATCH_TYPES:
        if base + "weight" in sd:  # single Linear
            return F.linear(value, sd[base + "weight"], sd[base + "bias"])
        h = value
        idx = 0
        while base + f"{idx}.weight" in sd:  # Sequential(Linear, ReLU, Linear[, ReLU, Linear])
            if idx > 0:
                h = torch.relu(h)
            h = F.linear(h, sd[base + f"{idx}.weight"], sd[base + f"{idx}.bias"])
            idx += 2
        return h
    raise NotImplementedError(f"Unknown modality type: {mcfg['type']}")


def embed_modalities(sd, cfg: OracleConfig, inputs: Dict[str, Any], emb_prefix="hf_model.embedding."):
    """MultimodalEmbedding.forward: per-modality embed -> LN(float) -> cat(dim=1) -> + pos-enc."""
    parts = []
    for modality, value in inputs.items():
        e = _embed_one(sd, emb_prefix, modality, cfg.data_config[modality], value)
        if cfg.multimodal_norm:
            nb = f"{emb_prefix}embedding_norm_dict.{modality}."
            e = _ln(e.float(), sd[nb + "weight"], sd[nb + "bias"])
        parts.append(e)
    x = torch.cat(parts, dim=1)
    S = x.shape[1]
    pb = f"{emb_prefix}positional_encodings."
    if cfg.positional_encoding_type == "sin_cos":
        pos = sd[pb + "pos_enc"][:S]
    else:  # learned: Embedding(arange(S)) then LayerNorm (utils.py:267-272)
        pos = _ln(sd[pb + "pos_encodings.weight"][:S], sd[pb + "norm.weight"], sd[pb + "norm.bias"])
    return x + pos.unsqueeze(0)


# --------------------------------------------------------------------------------------------
# transformer blocks (custom_modeling.py:108-199 over torch's norm_first layer equations)
# --------------------------------------------------------------------------------------------
def _drop(x, cfg):
    if cfg is not None and cfg.training and cfg.dropout > 0:
        return F.dropout(x, cfg.dropout, True)
    return x


def _mha(xq, xkv, sd, p, n_heads, add_mask, cfg=None):
    """Packed-QKV multi-head attention; `add_mask` is an additive float mask broadcastable to
    [B, H, Lq, Lk] (0 / -inf), as torch's F.multi_head_attention_forward builds it."""
    B, Lq, d = xq.shape
    Lk = xkv.shape[1]
    dh = d // n_heads
    w, b = sd[p + "in_proj_weight"], sd[p + "in_proj_bias"]
    q = F.linear(xq, w[:d], b[:d])
    k = F.linear(xkv, w[d : 2 * d], b[d : 2 * d])
    v = F.linear(xkv, w[2 * d :], b[2 * d :])
    q = q.view(B, Lq, n_heads, dh).transpose(1, 2)
    k = k.view(B, Lk, n_heads, dh).transpose(1, 2)
    v = v.view(B, Lk, n_heads, dh).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dh)
    if add_mask is not None:
        s = s + add_mask
    a = _drop(torch.softmax(s, dim=-1), cfg)
    o = torch.matmul(a, v).transpose(1, 2).reshape(B, Lq, d)
    return F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def _ffn(h, sd, p, gated, cfg=None):
    """W2 * (gelu(W1 h) [ * (Wg h) ])  (custom_modeling.py:137-152,184-199); exact erf GELU."""
    u = F.gelu(F.linear(h, sd[p + "linear1.weight"], sd[p + "linear1.bias"]))
    if gated:
        u = u * F.linear(h, sd[p + "gate.weight"], sd[p + "gate.bias"])
    return F.linear(_drop(u, cfg), sd[p + "linear2.weight"], sd[p + "linear2.bias"])


def _key_pad_mask(valid):  # valid: [B, L] (1 = real token) -> additive [B,1,1,L]
    m = torch.zeros(valid.shape, dtype=torch.float32)
    m = m.masked_fill(~valid.bool(), float("-inf"))
    return m[:, None, None, :]


def encoder_stack(sd, cfg: OracleConfig, x, attention_mask, prefix="hf_model.encoder."):
    """Encoder layers + final LayerNorm (custom_modeling.py:220-243,350-360).  `post_layer_normalisation=True` (every
    shipped config) is torch's norm_first branch, x + f(LN(x)); False is LN(x + f(x))."""
    km = _key_pad_mask(attention_mask)
    for i in range(cfg.encoder_layers):
        p = f"{prefix}layers.{i}."
        if not cfg.post_layer_normalisation:
            x = _ln(x + _drop(_mha(x, x, sd, p + "self_attn.", cfg.encoder_attention_heads, km, cfg), cfg),
                    sd[p + "norm1.weight"], sd[p + "norm1.bias"])
            x = _ln(x + _drop(_ffn(x, sd, p, cfg.gated_linear, cfg), cfg), sd[p + "norm2.weight"], sd[p + "norm2.bias"])
            continue
        h = _ln(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        x = x + _drop(_mha(h, h, sd, p + "self_attn.", cfg.encoder_attention_heads, km, cfg), cfg)
        h = _ln(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
        x = x + _drop(_ffn(h, sd, p, cfg.gated_linear, cfg), cfg)
    return _ln(x, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])


def decoder_stack(sd, cfg: OracleConfig, ids, memory, memory_mask, dec_mask=None, prefix="hf_model.decoder."):
    """Target embed + causal/pad self-attn + cross-attn + FFN, pre-LN, final LN
    (custom_modeling.py:271-320).  `dec_mask=None` == the generating branch (no target pad mask)."""
    x = embed_modalities(sd, cfg, {cfg.target_modality: ids}, emb_prefix=prefix + "embedding.")
    T = ids.shape[1]
    causal = torch.triu(torch.full((T, T), float("-inf")), diagonal=1)[None, None]
    self_mask = causal if dec_mask is None else causal + _key_pad_mask(dec_mask)
    mem_mask = _key_pad_mask(memory_mask)
    for i in range(cfg.decoder_layers):
        p = f"{prefix}layers.{i}."
        if not cfg.post_layer_normalisation:
            H = cfg.decoder_attention_heads
            x = _ln(x + _drop(_mha(x, x, sd, p + "self_attn.", H, self_mask, cfg), cfg),
                    sd[p + "norm1.weight"], sd[p + "norm1.bias"])
            x = _ln(x + _drop(_mha(x, memory, sd, p + "multihead_attn.", H, mem_mask, cfg), cfg),
                    sd[p + "norm2.weight"], sd[p + "norm2.bias"])
            x = _ln(x + _drop(_ffn(x, sd, p, cfg.gated_linear, cfg), cfg), sd[p + "norm3.weight"], sd[p + "norm3.bias"])
            continue
        h = _ln(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        x = x + _drop(_mha(h, h, sd, p + "self_attn.", cfg.decoder_attention_heads, self_mask, cfg), cfg)
        h = _ln(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
        x = x + _drop(_mha(h, memory, sd, p + "multihead_attn.", cfg.decoder_attention_heads, mem_mask, cfg), cfg)
        h = _ln(x, sd[p + "norm3.weight"], sd[p + "norm3.bias"])
        x = x + _drop(_ffn(h, sd, p, cfg.gated_linear, cfg), cfg)
    return _ln(x, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])


# --------------------------------------------------------------------------------------------
# align head (custom_modeling.py:363-396, 453-475)
# --------------------------------------------------------------------------------------------
def _kl(p, q, eps=1e-16):
    q = q.clamp(min=eps)
    p = p.clamp(min=eps)
    return (p * (p / q).log()).sum() / p.shape[0]


def _align_loss(sd, cfg: OracleConfig, enc_out, attention_mask, target):
    ac = cfg.align_config
    m = attention_mask.unsqueeze(-1)
    pooled = (enc_out * m).sum(dim=1) / m.sum(dim=1)
    p = "hf_model.align_network."
    h = torch.relu(F.linear(pooled, sd[p + "0.weight"], sd[p + "0.bias"]))
    if ac["align_network"] == "convolutional":
        h = F.linear(h, sd[p + "2.weight"], sd[p + "2.bias"]).unsqueeze(-1)
        h = torch.relu(F.conv1d(h, sd[p + "4.weight"], sd[p + "4.bias"], padding=ac["kernel_size"] // 2))
        pred = torch.sigmoid(F.conv1d(h, sd[p + "6.weight"], sd[p + "6.bias"])).squeeze(-1)
    elif ac["align_network"] == "mlp":
        pred = torch.sigmoid(F.linear(h, sd[p + "2.weight"], sd[p + "2.bias"]))
    else:
        raise ValueError(f"unknown align network {ac['align_network']}")
    fn = ac["loss_function"]
    if fn == "mae":
        return (pred - target).abs().mean()
    if fn == "mse":
        return ((pred - target) ** 2).mean()
    if fn == "sid":
        return _kl(pred, target) + _kl(target, pred)
    raise ValueError(f"Loss function {fn} not supported for alignment!")


# --------------------------------------------------------------------------------------------
# whole model (custom_modeling.py:420-508, wrapper.py:346-407)
# --------------------------------------------------------------------------------------------
def model_forward(sd, cfg: OracleConfig, enc_inputs, attention_mask, dec_ids, dec_mask, labels=None,
                  align_target=None):
    """enc_inputs: {modality: batch-first tensor}; masks: [B,L] 1=valid; labels: -100 = ignore."""
    x = embed_modalities(sd, cfg, enc_inputs)
    enc = encoder_stack(sd, cfg, x, attention_mask)
    dec = decoder_stack(sd, cfg, dec_ids, enc, attention_mask, dec_mask)
    logits = F.linear(dec, sd["hf_model.token_ff.weight"], sd["hf_model.token_ff.bias"])
    out = {"logits": logits, "encoder_hidden_states": enc, "decoder_hidden_states": dec}
    if labels is not None:
        V = logits.shape[-1]
        lm = F.cross_entropy(logits.reshape(-1, V), labels.reshape(-1), ignore_index=-100)
        out["model_only_loss"] = lm
        total = lm
        if cfg.align_config and align_target is not None:
            al = _align_loss(sd, cfg, enc, attention_mask, align_target)
            out["alignment_loss"] = al
            total = lm + cfg.align_config["loss_lambda"] * al
        out["loss"] = total
    return out


def relayout_batch(cfg: OracleConfig, batch):
    """Collator dict (seq-first, True = pad) -> batch-first tensors (wrapper.py:356-365,389)."""
    enc = {}
    for m, v in batch["encoder_input"].items():
        if isinstance(v, dict):
            enc[m] = {k: t.transpose(1, 0) for k, t in v.items()}
        else:
            enc[m] = v.transpose(1, 0)
    attention_mask = (~batch["encoder_pad_mask"]).int().T
    dec_ids = batch["decoder_input"][cfg.target_modality].transpose(1, 0)
    dec_mask = (~batch["decoder_pad_mask"]).int().T
    labels = batch["target"].T.contiguous().clone()
    labels[labels == cfg.pad_token_id] = -100
    return enc, attention_mask, dec_ids, dec_mask, labels


def wrapper_forward(sd, cfg: OracleConfig, batch):
    enc, am, dec_ids, dm, labels = relayout_batch(cfg, batch)
    return model_forward(sd, cfg, enc, am, dec_ids, dm, labels, batch.get("encoder_alignment_input"))


# --------------------------------------------------------------------------------------------
# generation (wrapper.py:409-453 + transformers GenerationMixin semantics)
# --------------------------------------------------------------------------------------------
def _next_logits(sd, cfg, seqs, memory, memory_mask):
    """Reference behaviour: use_cache=False, the whole prefix is re-decoded every step
    (wrapper.py:450; custom_modeling.py:447-450,478-486) and no target pad mask is passed."""
    dec = decoder_stack(sd, cfg, seqs, memory, memory_mask, None)
    return F.linear(dec[:, -1], sd["hf_model.token_ff.weight"], sd["hf_model.token_ff.bias"]).float()


def _force_eos(scores, cur_len, cfg):
    """ForcedEOSTokenLogitsProcessor: at cur_len == max_length-1 everything but <eos> -> -inf, <eos> -> 0."""
    if cur_len == cfg.max_length - 1:
        scores = torch.full_like(scores, float("-inf"))
        scores[:, cfg.eos_token_id] = 0.0
    return scores


def _apply_hooks(hooks, seqs, scores):
    """`logits_processor=` of wrapper.py:443-451: one callable or a list, `(input_ids, scores) -> scores`, in order."""
    if hooks is None:
        return scores
    for h in (hooks if isinstance(hooks, (list, tuple)) else [hooks]):
        scores = h(seqs, scores)
    return scores


def generate(sd, cfg: OracleConfig, batch, n_beams: int = 1, logits_hook=None, return_scores=False):
    """Greedy (n_beams=1) or beam search with num_return_sequences = n_beams.
    Output: int64 [B*n_beams, L<=max_length]; row b*K+r is the r-th best hypothesis of sample b."""
    enc, am, _, _, _ = relayout_batch(cfg, batch)
    with torch.no_grad():
        memory = encoder_stack(sd, cfg, embed_modalities(sd, cfg, enc), am)
        if n_beams == 1:
            return _greedy(sd, cfg, memory, am, logits_hook)
        return _beam(sd, cfg, memory, am, n_beams, logits_hook, return_scores)


def _greedy(sd, cfg, memory, am, logits_hook):
    B = memory.shape[0]
    seqs = torch.full((B, 1), cfg.bos_token_id, dtype=torch.long)
    unfinished = torch.ones(B, dtype=torch.bool)
    while seqs.shape[1] < cfg.max_length and unfinished.any():
        cur_len = seqs.shape[1]
        scores = _next_logits(sd, cfg, seqs, memory, am)
        # transformers merges the caller's processors AFTER its own (ForcedEOS first): generation/utils.py
        # `_get_logits_processor` -> `_merge_criteria_processor_list`
        scores = _force_eos(scores, cur_len, cfg)
        scores = _apply_hooks(logits_hook, seqs, scores)
        nxt = scores.argmax(dim=-1)
        nxt = torch.where(unfinished, nxt, torch.full_like(nxt, cfg.pad_token_id))
        seqs = torch.cat([seqs, nxt[:, None]], dim=1)
        unfinished = unfinished & (nxt != cfg.eos_token_id)
    return seqs


def _beam(sd, cfg, memory, am, K, logits_hook, return_scores):
    """Vectorised beam search, length_penalty 1.0, early_stopping False (heuristic stop)."""
    NEG = -1.0e9
    B = memory.shape[0]
    L = cfg.max_length
    V = sd["hf_model.token_ff.weight"].shape[0]
    mem = memory.repeat_interleave(K, dim=0)
    mmask = am.repeat_interleave(K, dim=0)
    prompt = 1
    # transformers fills with `pad_token_id or eos_token_id`: a pad id of 0 is falsy, so the
    # reference's beam outputs are <eos>-padded (golden-vector verified)
    fill = cfg.pad_token_id if cfg.pad_token_id else cfg.eos_token_id
    run_seq = torch.full((B, K, L), fill, dtype=torch.long)
    run_seq[:, :, 0] = cfg.bos_token_id
    fin_seq = run_seq.clone()
    run_score = torch.zeros(B, K)
    run_score[:, 1:] = NEG
    fin_score = torch.full((B, K), NEG)
    fin_flag = torch.zeros(B, K, dtype=torch.bool)
    fin_len = torch.zeros(B, K, dtype=torch.long)  # generated length of each kept hypothesis
    improvable = torch.ones(B, 1, dtype=torch.bool)
    keep = 2 * K
    top_mask = torch.arange(keep) < K
    cur_len = prompt
    while True:
        flat = run_seq[:, :, :cur_len].reshape(B * K, cur_len)
        logits = _next_logits(sd, cfg, flat, mem, mmask)
        logp = torch.log_softmax(logits, dim=-1)
        logp = _force_eos(logp, cur_len, cfg)
        logp = _apply_hooks(logits_hook, flat, logp)
        acc = (logp.view(B, K, V) + run_score[:, :, None]).reshape(B, K * V)
        cand_score, cand_idx = torch.topk(acc, k=keep)
        cand_beam = cand_idx // V
        cand_tok = cand_idx % V
        cand_seq = torch.take_along_dim(run_seq, cand_beam[:, :, None], dim=1)
        cand_seq[:, :, cur_len] = cand_tok
        hits = (cand_tok == cfg.eos_token_id) | (cur_len + 1 >= L)
        # live beams for the next step: best K candidates that did not stop
        live = cand_score + hits.float() * NEG
        nxt = torch.topk(live, k=K)[1]
        run_seq = torch.take_along_dim(cand_seq, nxt[:, :, None], dim=1)
        run_score = torch.take_along_dim(live, nxt, dim=1)
        # finished pool: only candidates ranked < K may finish; score = sum_logp / generated_len
        just = hits & top_mask[None, :]
        fs = cand_score / float(cur_len + 1 - prompt)
        fs = fs + (~improvable).float() * NEG
        fs = fs + (~just).float() * NEG
        m_seq = torch.cat([fin_seq, cand_seq], dim=1)
        m_score = torch.cat([fin_score, fs], dim=1)
        m_flag = torch.cat([fin_flag, just], dim=1)
        m_len = torch.cat([fin_len, torch.full((B, keep), cur_len + 1 - prompt, dtype=torch.long)], dim=1)
        sel = torch.topk(m_score, k=K)[1]
        fin_seq = torch.take_along_dim(m_seq, sel[:, :, None], dim=1)
        fin_score = torch.take_along_dim(m_score, sel, dim=1)
        fin_flag = torch.take_along_dim(m_flag, sel, dim=1)
        fin_len = torch.take_along_dim(m_len, sel, dim=1)
        cur_len += 1
        best_possible = run_score[:, :1] / float(cur_len - prompt)
        worst = torch.where(fin_flag, fin_score.min(dim=1, keepdim=True)[0], torch.full_like(fin_score, NEG))
        improvable = improvable & (best_possible > worst).any(dim=-1, keepdim=True)
        if not (improvable.any() and not hits.all()):
            break
    out_len = prompt + int(fin_len.max())
    seqs = fin_seq.reshape(B * K, L)[:, :out_len]
    if return_scores:
        return seqs, fin_score.reshape(B * K)
    return seqs


# --------------------------------------------------------------------------------------------
# parameter construction for synthetic benchmarks (wrapper.py:320-327: xavier_uniform on dim>1)
# --------------------------------------------------------------------------------------------
def init_state_dict(cfg: OracleConfig, vocab: int, enc_ffn: int, dec_ffn: int, seed: int = 3247,
                    max_pos: int = 1024) -> Dict[str, torch.Tensor]:
    """Random-init a state dict with the reference's key layout and init rule (torch defaults for
    1-D params, xavier_uniform for everything with dim > 1)."""
    g = torch.Generator().manual_seed(seed)
    d = cfg.d_model
    sd: Dict[str, torch.Tensor] = {}

    def xav(*shape):
        fan_out, fan_in = shape[0], shape[1]
        a = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(*shape, generator=g) * 2 - 1) * a

    def bias(n, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(n, generator=g) * 2 - 1) * b

    def lin(p, out_f, in_f):
        sd[p + "weight"] = xav(out_f, in_f)
        sd[p + "bias"] = bias(out_f, in_f)

    def ln(p):
        sd[p + "weight"] = torch.ones(d)
        sd[p + "bias"] = torch.zeros(d)

    e = "hf_model.embedding."
    for m, mc in cfg.data_config.items():
        base = f"{e}embedding_layer_dict.{m}."
        if mc["type"] in TOKEN_TYPES:
            sd[base + "weight"] = xav(mc["vocab_size"], d)
        else:
            ps = 2 if mc["type"] == "msms_number" else mc["preprocessor_arguments"]["patch_size"]
            et = mc["preprocessor_arguments"].get("encoding_type", "linear")
            if et == "linear":
                lin(base, d, ps)
            elif et == "linear_2_layer":
                lin(base + "0.", d // 2, ps)
                lin(base + "2.", d, d // 2)
            elif et == "linear_3_layer":
                lin(base + "0.", d // 3, ps)
                lin(base + "2.", 2 * (d // 3), d // 3)
                lin(base + "4.", d, 2 * (d // 3))
            else:
                raise NotImplementedError
        if cfg.multimodal_norm:
            ln(f"{e}embedding_norm_dict.{m}.")
    if cfg.positional_encoding_type == "sin_cos":
        sd[e + "positional_encodings.pos_enc"] = sincos_table(d, max_pos)
    else:
        sd[e + "positional_encodings.pos_encodings.weight"] = xav(max_pos, d)
        ln(e + "positional_encodings.norm.")

    def attn(p):
        sd[p + "in_proj_weight"] = xav(3 * d, d)
        sd[p + "in_proj_bias"] = torch.zeros(3 * d)
        sd[p + "out_proj.weight"] = xav(d, d)
        sd[p + "out_proj.bias"] = torch.zeros(d)

    for i in range(cfg.encoder_layers):
        p = f"hf_model.encoder.layers.{i}."
        attn(p + "self_attn.")
        lin(p + "linear1.", enc_ffn, d)
        lin(p + "linear2.", d, enc_ffn)
        if cfg.gated_linear:
            lin(p + "gate.", enc_ffn, d)
        ln(p + "norm1.")
        ln(p + "norm2.")
    ln("hf_model.encoder.norm.")
    if cfg.align_config:  # custom_modeling.py:363-396
        ac, p = cfg.align_config, "hf_model.align_network."
        hd = ac["hidden_dimension"]
        lin(p + "0.", hd, d)
        if ac["align_network"] == "convolutional":
            lin(p + "2.", hd, hd)
            ks, cc = ac["kernel_size"], ac["conv_channels"]
            a = math.sqrt(6.0 / (hd * ks + cc * ks))
            sd[p + "4.weight"] = (torch.rand(cc, hd, ks, generator=g) * 2 - 1) * a
            sd[p + "4.bias"] = bias(cc, hd * ks)
            a = math.sqrt(6.0 / (cc + ac["output_dimension"]))
            sd[p + "6.weight"] = (torch.rand(ac["output_dimension"], cc, 1, generator=g) * 2 - 1) * a
            sd[p + "6.bias"] = bias(ac["output_dimension"], cc)
        else:
            lin(p + "2.", ac["output_dimension"], hd)
    for i in range(cfg.decoder_layers):
        p = f"hf_model.decoder.layers.{i}."
        attn(p + "self_attn.")
        attn(p + "multihead_attn.")
        lin(p + "linear1.", dec_ffn, d)
        lin(p + "linear2.", d, dec_ffn)
        if cfg.gated_linear:
            lin(p + "gate.", dec_ffn, d)
        ln(p + "norm1.")
        ln(p + "norm2.")
        ln(p + "norm3.")
    ln("hf_model.decoder.norm.")
    lin("hf_model.token_ff.", vocab, d)
    # the decoder shares the embedding module (custom_modeling.py:409-415)
    for k in [k for k in sd if k.startswith(e)]:
        sd["hf_model.decoder.embedding." + k[len(e):]] = sd[k]
    return sd
