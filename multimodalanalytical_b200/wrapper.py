"""Drop-in boundary: the reference's Lightning-module surface (`analytical_fm/modeling/wrapper.py`) over the
B200 engine.

Mirrored names and behaviour (reference file:line):
  * MODEL_REGISTRY / load_custom_model           wrapper.py:144-180, 222-227   (no network: the three facts the
    reference pulls from facebook/bart-base - is_encoder_decoder, dropout=0.1, gelu - are constants here)
  * HFWrapper.__init__ signature and attributes   wrapper.py:233-318
  * forward(batch) batch re-layout, modality dropout, pad -> -100          wrapper.py:346-407
  * generate(batch, n_beams, logits_processor)    wrapper.py:409-453
  * training_step / validation_step / predict_step / configure_optimizers  wrapper.py:329-344, 455-578
  * _calc_token_acc, score_val_sequences           wrapper.py:606-655
  * state_dict key layout                           SURVEY.md §8(b) checkpoint contract
The module subclasses `pytorch_lightning.LightningModule` when Lightning is importable and a plain
`torch.nn.Module` (with a no-op `log`) otherwise, so it also runs under the built-in loop in trainer.py.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import nn

from . import ops
from .decode import Generator
from .model import Engine
from .params import ModelConfig, ParamStore

try:  # pragma: no cover - Lightning is not installed in the build image
    import pytorch_lightning as pl

    _Base = pl.LightningModule
except Exception:  # noqa: BLE001
    pl = None

    class _Base(nn.Module):  # type: ignore[no-redef]
        def log(self, *a, **k):
            pass

try:  # ListConfig is what the reference checks for modality dropout (wrapper.py:368)
    from omegaconf.listconfig import ListConfig  # type: ignore
except Exception:  # noqa: BLE001
    class ListConfig(list):  # type: ignore[no-redef]
        """Stand-in used when omegaconf is absent; pass `ListConfig([...])` to enable modality dropout."""


@dataclasses.dataclass
class CustomLMOutput:
    """Same fields the callers read from the reference's CustomLMOutput (modeling/utils.py:25-30).  A dataclass, so
    that a DistributedDataParallel wrapper finds the loss tensor in the module's output (it walks lists, dicts and
    dataclasses when `find_unused_parameters=True`, the reference's strategy) as it does in HF's ModelOutput."""

    loss: Any = None
    logits: Any = None
    loss_dict: Any = None
    encoder_hidden_states: Any = None
    decoder_hidden_states: Any = None

    def __getitem__(self, k):
        return getattr(self, k)


class _EngineLoss(torch.autograd.Function):
    """Connects the engine's hand-scheduled backward to `loss.backward()` (what Lightning / a plain torch loop calls).

    Every exposed `nn.Parameter` is an INPUT of this node, so autograd delivers one gradient per parameter through the
    parameter's own AccumulateGrad node: `.grad` is populated / accumulated by autograd itself (it survives
    `optimizer.zero_grad(set_to_none=True)`, which Lightning issues before every backward), gradient accumulation over
    micro-batches adds up, and a stock `DistributedDataParallel` wrapper sees its per-parameter hooks fire and
    all-reduces the gradients as it does for any module (reference: Lightning DDP, trainer/trainer.py:58-71).  The
    engine writes into its (freshly zeroed) flat gradient buffer; the node returns views of ONE flat copy of it (a
    177 MB device copy at d_model 512, ~0.05 ms)."""

    @staticmethod
    def forward(ctx, loss_value, module, *params):
        ctx.module = module
        return loss_value.detach().clone().reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        m = ctx.module
        store = m.store
        store.g.zero_()
        m.engine.backward(gscale=float(grad_out))
        flat = store.g.clone()
        store.g_dirty = True  # holds this backward's gradients (read by tests / diagnostics); FusedTrainer re-zeroes
        store.bf16_dirty = True  # an optimiser outside the engine is about to change the master weights
        grads = []
        for name in m._names:
            off, shape = store.offsets[name]
            n = 1
            for x in shape:
                n *= x
            grads.append(flat[off: off + n].view(shape))
        return (None, None, *grads)


class B200CustomModel:
    """What `MODEL_REGISTRY['CustomModel']` returns in place of the reference's `CustomModel`: a handle on the
    engine that owns encoder, decoder, LM head and loss."""

    def __init__(self, cfg: ModelConfig, store: ParamStore, engine: Engine):
        self.config, self.store, self.engine = cfg, store, engine
        self.target_modality = cfg.target_modality
        self.decoder_vocab_size = cfg.vocab_size


def load_custom_model(model_name: str, target_tokenizer, target_modality: str, data_config: Dict[str, Any],
                      multimodal_norm: bool, precision: str = "bf16", device="cuda", seed: Optional[int] = None,
                      **kwargs) -> Tuple[B200CustomModel, ParamStore]:
    """wrapper.py:144-180.  `model_name` ('facebook/bart-base') is accepted and ignored: nothing is fetched."""
    del model_name, target_modality
    heads = kwargs.get("num_heads", 8)
    cfg = ModelConfig(
        data_config=data_config,
        vocab_size=target_tokenizer.vocab_size,
        d_model=kwargs.get("d_model", 512),
        encoder_layers=kwargs.get("encoder_layers", 6),
        decoder_layers=kwargs.get("decoder_layers", 6),
        encoder_attention_heads=kwargs.get("encoder_attention_heads", heads),
        decoder_attention_heads=kwargs.get("decoder_attention_heads", heads),
        encoder_ffn_dim=kwargs.get("encoder_ffn_dim", 2048),
        decoder_ffn_dim=kwargs.get("decoder_ffn_dim", 2048),
        dropout=kwargs.get("dropout", 0.1),
        gated_linear=bool(kwargs.get("gated_linear", False)),
        post_layer_normalisation=bool(kwargs.get("post_layer_normalisation", True)),
        positional_encoding_type=kwargs.get("positional_encoding_type", "sin_cos"),
        multimodal_norm=bool(multimodal_norm),
        max_position_embeddings=kwargs.get("max_position_embeddings", 1024),
        pad_token_id=target_tokenizer.pad_token_id,
        bos_token_id=target_tokenizer.bos_token_id,
        eos_token_id=target_tokenizer.eos_token_id,
        align_config=kwargs.get("align_config"),
        label_smoothing=float(kwargs.get("label_smoothing", 0.0)),
    )
    if cfg.align_config:
        ac = cfg.align_config = dict(cfg.align_config)
        if ac.get("align_network") not in ("convolutional", "mlp"):
            raise ValueError(f"unknown align network {ac.get('align_network')}")
    store = ParamStore(cfg, device=device, seed=seed)
    engine = Engine(cfg, store, precision=precision)
    return B200CustomModel(cfg, store, engine), store


def _dist_rank() -> int:
    import os

    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank()
    return int(os.environ.get("RANK", "0"))


def _unsupported(name):
    def loader(*a, **k):
        raise NotImplementedError(f"model_type {name} is outside the accelerated hot path (SURVEY.md §2)")

    return loader


MODEL_REGISTRY: Dict[str, Callable[..., Tuple[Any, Any]]] = {
    "T5ForConditionalGeneration": _unsupported("T5ForConditionalGeneration"),
    "BartForConditionalGeneration": _unsupported("BartForConditionalGeneration"),
    "CustomBartForConditionalGeneration": _unsupported("CustomBartForConditionalGeneration"),
    "CustomModel": load_custom_model,
}

OPTIMISER_REGISTRY = {"adam": torch.optim.Adam, "adamw": torch.optim.AdamW}


class HFWrapper(_Base):
    """Same constructor, attributes and hooks as the reference's HFWrapper."""

    def __init__(self, data_config: Dict[str, Any], model_type: str, model_name: str, target_tokenizer,
                 optimiser: str = "adam", num_steps: int = 1000, lr: float = 0.001, weight_decay: float = 0,
                 adam_beta1: float = 0.9, adam_beta2: float = 0.999, multimodal_norm: bool = True,
                 modality_dropout: Optional[List[str]] = None, **kwargs) -> None:
        super().__init__()
        if isinstance(target_tokenizer, str):
            raise NotImplementedError("pass a tokenizer object: nothing is fetched from the network")
        self.target_tokenizer = target_tokenizer
        self.model_type, self.model_name = model_type, model_name
        self.data_config = data_config
        self.multimodal_norm = multimodal_norm
        self.modality_dropout = modality_dropout
        self.guided_generation = kwargs.get("guided_generation", False)
        # chemistry of guided decoding (rdkit in the reference); None -> rdkit, imported when first needed
        self.chem_backend = kwargs.pop("chem_backend", None)
        self.target_modality = ""
        for modality, modality_config in self.data_config.items():
            if modality_config["target"]:
                self.target_modality = modality
        self.optimiser, self.lr, self.weight_decay = optimiser, float(lr), float(weight_decay)
        self.adam_beta1, self.adam_beta2, self.num_steps = float(adam_beta1), float(adam_beta2), num_steps
        self.validation_step_outputs: List[Dict[str, Any]] = []
        self.test_step_outputs: List[Dict[str, Any]] = []

        self.hf_model, self.multimodal_embedding = MODEL_REGISTRY[self.model_type](
            self.model_name, self.target_tokenizer, self.target_modality, self.data_config, self.multimodal_norm,
            **kwargs)
        self.store: ParamStore = self.hf_model.store
        self.engine: Engine = self.hf_model.engine
        self.generator = Generator(self.engine)
        self.generation_config = dict(
            bos_token_id=target_tokenizer.bos_token_id, decoder_start_token_id=target_tokenizer.bos_token_id,
            eos_token_id=target_tokenizer.eos_token_id, forced_eos_token_id=target_tokenizer.eos_token_id,
            max_length=128, pad_token_id=target_tokenizer.pad_token_id)
        self.n_beams = kwargs.get("n_beams", 10)
        # parameters exposed to torch / Lightning / DDP are views of the flat master buffer; their gradients arrive
        # through autograd (`_EngineLoss`)
        self._flat = nn.ParameterDict()
        self._names: List[str] = []
        for name, _, _ in self.store.specs:
            self._flat[name.replace(".", "|")] = nn.Parameter(self.store.P(name), requires_grad=True)
            self._names.append(name)
        # dropout stream: a function of the user's seed and of the rank (DDP ranks must draw different masks)
        self.engine.set_seed(kwargs.get("seed"), _dist_rank())
        self.engine_seed_probe = self.engine.seed  # diagnostics / tests: differs across ranks

    # --------------------------------------------------------------------------- checkpoint layout
    def state_dict(self, *args, destination=None, prefix="", keep_vars=False, **kw):  # noqa: D401
        self.store.gather_master()  # no-op unless the optimiser state is sharded by rank (trainer.PeerShardedStep)
        sd = self.store.state_dict(with_aliases=True)
        out = destination if destination is not None else {}
        for k, v in sd.items():
            out[prefix + k] = v if keep_vars else v.detach()
        return out

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        missing, unexpected = self.store.load_state_dict(state_dict, strict=strict)
        return missing, unexpected

    # ------------------------------------------------------------------------------------ optimiser
    def configure_optimizers(self):
        """wrapper.py:329-344: Adam/AdamW over all parameters + OneCycleLR stepped per optimiser step."""
        from torch.optim.lr_scheduler import OneCycleLR

        optim = OPTIMISER_REGISTRY[self.optimiser](self.parameters(), lr=self.lr, weight_decay=self.weight_decay,
                                                   betas=(self.adam_beta1, self.adam_beta2))
        sch = {"scheduler": OneCycleLR(optim, self.lr, total_steps=self.num_steps), "interval": "step"}
        self.store.bf16_dirty = True
        return [optim], [sch]

    def optimizer_step(self, *a, **k):  # Lightning hook: master weights changed -> bf16 mirror is stale
        self.store.bf16_dirty = True
        return super().optimizer_step(*a, **k) if hasattr(super(), "optimizer_step") else None

    def named_gradients(self) -> Dict[str, Optional[torch.Tensor]]:
        """{checkpoint name: .grad} of the exposed parameters (what a torch optimiser / DDP reducer sees)."""
        return {n: self._flat[n.replace(".", "|")].grad for n in self._names}

    # -------------------------------------------------------------------------------------- forward
    def _to_dev(self, t):
        return t.to(self.store.device, non_blocking=True) if isinstance(t, torch.Tensor) else t

    def _relayout(self, batch: Dict[str, Any], training: bool):
        """wrapper.py:356-389: seq-first -> batch-first, pad masks -> validity masks, modality dropout."""
        input_ids: Dict[str, Any] = {}
        for modality, v in batch["encoder_input"].items():
            if isinstance(v, dict):
                input_ids[modality] = {k: self._to_dev(t).transpose(1, 0).contiguous() for k, t in v.items()}
            else:
                input_ids[modality] = self._to_dev(v).transpose(1, 0).contiguous()
        attention_mask = (~self._to_dev(batch["encoder_pad_mask"])).T
        if isinstance(self.modality_dropout, ListConfig) and training and len(self.modality_dropout) > 0:
            drop = np.random.choice(self.modality_dropout, np.random.randint(0, len(self.modality_dropout)),
                                    replace=False)
            keep_cols, idx = [], 0
            for modality, v in input_ids.items():
                n = (v["tokenized_input"] if isinstance(v, dict) else v).shape[1]
                if modality not in drop:
                    keep_cols.append(attention_mask[:, idx: idx + n])
                idx += n
            for modality in drop:
                input_ids.pop(modality)
            attention_mask = torch.cat(keep_cols, dim=-1)
        return input_ids, attention_mask.to(torch.uint8).contiguous()

    def forward(self, batch: Dict[str, Any]) -> CustomLMOutput:
        input_ids, attention_mask = self._relayout(batch, self.training)
        dec_in = self._to_dev(batch["decoder_input"][self.target_modality]).transpose(1, 0).contiguous()
        dec_mask = (~self._to_dev(batch["decoder_pad_mask"])).T.to(torch.uint8).contiguous()
        labels = self._to_dev(batch["target"]).T.contiguous().clone()
        labels[labels == self.target_tokenizer.pad_token_id] = -100
        train = self.training and torch.is_grad_enabled()
        # wrapper.py:394-396: the align target rides along when the collator provides it
        align_target = None
        if self.engine.cfg.align_config and "encoder_alignment_input" in batch:
            align_target = self._to_dev(batch["encoder_alignment_input"]).float().contiguous()
        out = self.engine.forward(input_ids, attention_mask, dec_in, dec_mask, labels=labels, train=train,
                                  align_target=align_target)
        loss = out["loss"]
        if train:
            loss = _EngineLoss.apply(loss, self, *[self._flat[n.replace(".", "|")] for n in self._names])
        loss_dict = {"model_only_loss": out["lm_loss"], "alignment_loss": out["align_loss"]}
        return CustomLMOutput(loss=loss, logits=out["logits"], loss_dict=loss_dict,
                              encoder_hidden_states=out["memory"])

    def generate(self, batch: Dict[str, Any], n_beams: int = 1, logits_processor=None, **kw) -> torch.Tensor:
        """wrapper.py:409-453.  `logits_processor`: None; a processor or a list of processors with transformers'
        protocol `(input_ids [rows, cur_len] int64, scores [rows, V] fp32) -> scores` on CUDA tensors, applied after
        ForcedEOS exactly where transformers applies them (a lone `guided.GuidedFormulaProcessor` is fused into the
        step kernel); or an additive fp32 [B*n_beams, V] device tensor."""
        input_ids, attention_mask = self._relayout(batch, training=False)
        extra_bias, processors = None, None
        if isinstance(logits_processor, torch.Tensor):
            extra_bias = logits_processor
        elif logits_processor is not None:
            processors = list(logits_processor) if isinstance(logits_processor, (list, tuple)) else [logits_processor]
            for proc in processors:
                if not callable(proc):
                    raise TypeError(f"logits processor {type(proc).__name__} is not callable")
        return self.generator.generate(input_ids, attention_mask, n_beams=n_beams,
                                       max_length=self.generation_config["max_length"], extra_bias=extra_bias,
                                       processors=processors, **kw)

    # ---------------------------------------------------------------------------------------- hooks
    def training_step(self, batch: Dict[str, Any], batch_idx: int) -> torch.Tensor:
        self.train()
        self.engine.next_seed()  # fresh dropout masks every step
        model_output = self.forward(batch)
        loss = model_output.loss
        if (batch_idx % 10) == 0:
            self.log("train_loss", loss, prog_bar=True, on_step=True, logger=True, sync_dist=True)
            for key, val in (model_output.loss_dict or {}).items():
                if val is not None:
                    self.log(f"train_{key}", val, prog_bar=True, on_step=True, logger=True, sync_dist=True)
        return loss

    @torch.no_grad()
    def validation_step(self, batch: Dict[str, Any], batch_idx: int) -> Dict[str, Any]:  # noqa: ARG002
        self.eval()
        model_output = self.forward(batch)
        loss = model_output.loss
        token_acc = self._calc_token_acc(batch, model_output)
        generated = self.generate(batch, n_beams=1)
        scores = self.score_val_sequences(generated, self._to_dev(batch["target"]).T.clone(), n_beams=1)
        val_outputs = {
            "val_loss": loss, "val_token_acc": token_acc,
            "val_molecular_accuracy_tensorboard": torch.tensor([scores["Top-1"]], device=loss.device),
            "val_molecular_accuracy": torch.tensor([scores["Top-1"]], device=loss.device),
        }
        for key, val in (model_output.loss_dict or {}).items():
            val_outputs[f"val_{key}"] = val
        self.validation_step_outputs.append(val_outputs)
        return val_outputs

    def on_validation_epoch_end(self):
        colls = self.validation_step_outputs
        if colls:
            keys = list(colls[0].keys())
            for key in keys:
                vals = [c[key] for c in colls]
                if any(v is None for v in vals):
                    continue
                avg = sum(vals) / len(vals)
                if key == "val_molecular_accuracy":
                    self.log(key, avg, prog_bar=True, logger=False, sync_dist=True)
                else:
                    self.log(key, avg, sync_dist=True)
        self.validation_step_outputs = []

    @torch.no_grad()
    def predict_step(self, batch, batch_idx):  # noqa: ARG002
        self.eval()
        model_output = self.forward(batch)
        loss = model_output.loss
        if self.guided_generation:  # wrapper.py:546-556
            from .guided import GuidedFormulaProcessor, RDKitChem

            chem = self.chem_backend if self.chem_backend is not None else RDKitChem()
            target_formula = [chem.formula(smiles) for smiles in batch["target_smiles"]]
            processor = GuidedFormulaProcessor(self.n_beams, target_formula, self.target_tokenizer, chem=chem)
            generated = self.generate(batch, n_beams=self.n_beams, logits_processor=[processor])
        else:
            generated = self.generate(batch, n_beams=self.n_beams)
        decoded = self.target_tokenizer.batch_decode(generated, skip_special_tokens=True)
        extra = {k: v for k, v in batch.items()
                 if not k.startswith("encoder_") and not k.startswith("decoder_") and not k.startswith("target_")}
        return {"loss": loss, "predictions": decoded, "targets": batch.get("target_smiles"), **extra}

    # -------------------------------------------------------------------------------------- metrics
    def score_val_sequences(self, generated_sequences, targets, n_beams: int) -> Dict[str, float]:
        targets[targets == -100] = self.target_tokenizer.pad_token_id
        tgt = self.target_tokenizer.batch_decode(targets, skip_special_tokens=True)
        dec = self.target_tokenizer.batch_decode(generated_sequences, skip_special_tokens=True)
        dec = [dec[i * n_beams: (i + 1) * n_beams] for i in range(len(dec) // n_beams)]
        return top_n_string_accuracy(dec, tgt)

    def _calc_token_acc(self, batch_input, model_output):
        token_ids = self._to_dev(batch_input["target"]).T
        pred = torch.argmax(model_output.logits, dim=-1)
        mask = token_ids != -100
        correct = torch.eq(token_ids, pred) * mask
        return correct.sum().float() / mask.sum().float()


def _clean(sample: str) -> str:
    return sample.replace("<bos>", "").replace("<pad>", "").replace("<eos>", "").replace(" ", "")


def top_n_string_accuracy(samples: List[List[str]], targets: List[str]) -> Dict[str, float]:
    """Top-N bookkeeping of `calc_sampling_metrics(..., molecules=False)` (analytical_fm/utils.py:86-153):
    rank of the first exact string match among the n_beams candidates."""
    n_beams = len(samples[0])
    ranks = []
    for preds, tgt in zip(samples, targets):
        preds = [_clean(p) for p in preds]
        t = _clean(tgt)
        ranks.append(preds.index(t) if t in preds else n_beams)
    return {f"Top-{i + 1}": float(sum(r <= i for r in ranks) / len(ranks)) for i in range(n_beams)}
