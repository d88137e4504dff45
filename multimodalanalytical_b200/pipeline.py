"""Collator -> device pipeline.

Reference: `MultiModalDataCollator` (analytical_fm/data/datamodules.py:18-399) runs inside DataLoader workers for every
batch: tokenizer / preprocessor calls on Python lists, list -> tensor conversions, seq-first transposes, bool pad
masks; `HFWrapper.forward` then transposes everything back (modeling/wrapper.py:356-389).  At 10^4-10^5 spectra/s per
GPU that host work is the bottleneck.

Here the dataset is pushed through the reference's own preprocessor objects ONCE (`pretokenise`, duck-typed: the same
calls the collator makes), stored ragged (flat values + row offsets, raw spectra as one fp32 table) and kept resident
in HBM (a million 1791-point spectra with their token rows are < 8 GB of the 180 GB).  A batch is then nothing but B
sample indices: `DeviceDataset.collate` copies them to the GPU and the `mma_collate_*` / `mma_patchify_rows` kernels
gather + pad + mask + shift straight into the batch-first layout the engine consumes (ids int64 [B, L], validity masks
u8, labels with -100) - no per-batch tokenisation, no transposes, no bool masks.  `wire_batch` re-expresses the same
batch in the reference collator's wire format (seq-first, True = pad) for drop-in use and for the parity tests, which
compare it tensor for tensor with batches produced by the reference collator (tests/golden/collate.pt).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .preprocess import INTERP_OFFSET, INTERP_POINTS

TOKENISED_TYPES = ("multiplets", "carbon", "msms_text")  # datamodules.py:254-259 (padded to the batch's longest row)
# spectra written out as text: their preprocessors pad every batch to `max_sequence_length` (text_spectrum.py:113-119,
# 465-471), so these columns have a fixed padded length like the plain text inputs
SPECTRUM_TEXT_TYPES = ("text_spectrum", "run_length_encoding")  # datamodules.py:277-304, 321-332


# ---------------------------------------------------------------------------------------------- host: ragged columns
@dataclass
class Ragged:
    """Variable-length rows: `flat` int32 [n] or fp32 [n, width], `offsets` int64 [N + 1], `valid` u8 [N] or None
    (0 = the sample has no data for this modality -> fully masked, carbon.py:52-56 / multiplets.py:82-86)."""

    flat: np.ndarray
    offsets: np.ndarray
    valid: Optional[np.ndarray] = None

    @classmethod
    def from_rows(cls, rows: Sequence[np.ndarray], dtype, valid=None, width: int = 0):
        lens = np.fromiter((len(r) for r in rows), dtype=np.int64, count=len(rows))
        offsets = np.zeros(len(rows) + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        if len(rows) and offsets[-1]:
            flat = np.concatenate([np.asarray(r, dtype=dtype).reshape(-1, width) if width else np.asarray(r, dtype=dtype)
                                   for r in rows if len(r)])
        else:
            flat = np.zeros((0, width) if width else (0,), dtype=dtype)
        return cls(np.ascontiguousarray(flat), offsets, None if valid is None else np.asarray(valid, dtype=np.uint8))

    @property
    def lengths(self) -> np.ndarray:
        return np.diff(self.offsets)

    def row(self, i: int) -> np.ndarray:
        return self.flat[self.offsets[i]: self.offsets[i + 1]]


@dataclass
class Column:
    """One modality of the pre-tokenised dataset."""

    kind: str                      # "tokens" | "values" | "patches"
    pad_len: Optional[int] = None  # tokens: fixed padded length (text inputs, datamodules.py:238-245); None = longest
    max_len: int = 1 << 30         # truncation bound (tokenizer truncation=True)
    pad_id: int = 0
    tokens: Optional[Ragged] = None
    values: Optional[Ragged] = None      # XVal numerical values (pad 1.0) or msms_number peaks (width 2, pad 0.0)
    pad_value: float = 0.0
    raw: Optional[np.ndarray] = None     # patches: fp32 [N, n_points]
    missing: Optional[np.ndarray] = None  # patches: u8 [N], 1 = sample has no spectrum
    patch: Dict[str, Any] = field(default_factory=dict)  # patch_size, mean, std, interpolation, overlap, masking


@dataclass
class HostDataset:
    columns: Dict[str, Column]           # input modalities in config order
    target_modality: str
    target: Column
    n: int
    passthrough: Dict[str, List[Any]] = field(default_factory=dict)  # e.g. target strings, extra columns
    alignment: Optional[np.ndarray] = None  # fp32 [N, 1800]


def _column(rows, key):
    if hasattr(rows, "column_names") or isinstance(rows, dict):
        return list(rows[key])
    return [r[key] for r in rows]


def _ragged_from_padded(ids: np.ndarray, pad_id: int, mask: Optional[np.ndarray]):
    lens = (ids != pad_id).sum(axis=1)
    rows = [ids[i, : lens[i]] for i in range(ids.shape[0])]
    valid = None if mask is None else (mask.reshape(mask.shape[0], -1).sum(axis=1) > 0)
    return rows, lens, valid


def pretokenise(rows, preprocessors: Dict[str, Any], data_config: Dict[str, Any],
                max_source_length: Dict[str, int], max_target_length: int, chunk: int = 2048,
                extra_columns: Sequence[str] = ()) -> HostDataset:
    """Run every sample through the reference's preprocessors once.  `rows`: a `datasets.Dataset`, a dict of columns
    or a list of row dicts; `preprocessors` / `data_config` / `max_*_length`: what the reference hands its collator
    (datamodules.py:402-414).  The calls made on the preprocessor objects are the collator's own
    (datamodules.py:237-341,354-361), so any object that satisfies the collator satisfies this function."""
    inputs = [m for m, c in data_config.items() if not c["target"]]
    targets = [m for m, c in data_config.items() if c["target"] and not c.get("alignment")]
    aligns = [m for m, c in data_config.items() if c["target"] and c.get("alignment")]
    if len(aligns) > 1:
        raise ValueError("At most 1 target alignment modality can be specified.")
    if len(targets) != 1:
        raise ValueError("Only 1 target modality can be specified.")  # datamodules.py:57-60
    tgt = targets[0]
    n = len(_column(rows, tgt))
    cols: Dict[str, Column] = {}
    for m in inputs:
        mtype = data_config[m]["type"]
        pre = preprocessors[m]
        # a text_spectrum modality reads other columns (its own name need not be one)
        data = None if mtype == "text_spectrum" else _column(rows, m)
        if mtype == "text":
            toks: List[np.ndarray] = []
            for lo in range(0, n, chunk):
                enc = pre(data[lo: lo + chunk], padding=False, truncation=True, max_length=max_source_length[m])
                toks += [np.asarray(t, dtype=np.int32) for t in enc["input_ids"]]
            cols[m] = Column("tokens", pad_len=int(max_source_length[m]), max_len=int(max_source_length[m]),
                             pad_id=int(pre.pad_token_id), tokens=Ragged.from_rows(toks, np.int32))
        elif mtype in TOKENISED_TYPES:
            pad_id = int(pre.tokenizer.pad_token_id)
            toks, vals, valid = [], [], []
            for lo in range(0, n, chunk):
                enc = pre(data[lo: lo + chunk])
                ids = enc["input_ids"].numpy()
                r, lens, v = _ragged_from_padded(ids, pad_id, enc["attention_mask"].numpy())
                toks += [x.astype(np.int32) for x in r]
                valid += list(v)
                if "numerical_values" in enc:
                    nv = enc["numerical_values"].numpy()
                    vals += [nv[i, : lens[i]].astype(np.float32) for i in range(len(r))]
            col = Column("tokens", pad_len=None, max_len=int(pre.max_sequence_length), pad_id=pad_id,
                         tokens=Ragged.from_rows(toks, np.int32, valid=valid))
            if vals:
                col.values, col.pad_value = Ragged.from_rows(vals, np.float32, width=1), 1.0
            cols[m] = col
        elif mtype in SPECTRUM_TEXT_TYPES:
            mc = data_config[m]
            formulae = None
            if mtype == "text_spectrum":
                data = _column(rows, mc["spectra_column"])
                formulae = None if mc["spectra_only"] else _column(rows, mc["formula_column"])
            pad_id = int(pre.tokenizer.pad_token_id)
            toks, vals = [], []
            for lo in range(0, n, chunk):
                spectra = np.asarray(data[lo: lo + chunk])  # add_padding_numerical_values reads spectra.shape
                if mtype == "text_spectrum":
                    enc = pre(formulae=None if formulae is None else formulae[lo: lo + chunk], spectra=spectra)
                else:
                    enc = pre(spectra=spectra)
                ids = enc["input_ids"].numpy()
                r, lens, _ = _ragged_from_padded(ids, pad_id, None)
                toks += [x.astype(np.int32) for x in r]
                if "numerical_values" in enc:
                    nv = enc["numerical_values"].numpy()
                    vals += [nv[i, : lens[i]].astype(np.float32) for i in range(len(r))]
            width = int(pre.max_sequence_length)
            col = Column("tokens", pad_len=width, max_len=width, pad_id=pad_id, tokens=Ragged.from_rows(toks, np.int32))
            if vals:
                col.values, col.pad_value = Ragged.from_rows(vals, np.float32, width=1), 1.0
            cols[m] = col
        elif mtype == "msms_number":
            peaks = []
            for lo in range(0, n, chunk):
                enc = pre(data[lo: lo + chunk])
                x, msk = enc["input_ids"].numpy(), enc["attention_mask"].numpy()
                lens = msk.sum(axis=1).astype(np.int64)
                peaks += [x[i, : lens[i]].astype(np.float32) for i in range(x.shape[0])]
            cols[m] = Column("values", values=Ragged.from_rows(peaks, np.float32, width=2), pad_value=0.0)
        elif mtype == "1D_patches":
            sizes = [len(s) if s is not None else -1 for s in data]
            width = max(sizes) if max(sizes) != -1 else 500  # patches.py:63-67
            raw = np.zeros((n, width), dtype=np.float32)
            for i, s in enumerate(data):
                if s is not None:
                    if len(s) != width:
                        raise ValueError(f"{m}: spectra of different lengths ({len(s)} vs {width}) cannot share a table")
                    raw[i] = np.asarray(s, dtype=np.float32)
            cols[m] = Column("patches", raw=raw, missing=np.asarray([s == -1 for s in sizes], dtype=np.uint8),
                             patch=dict(patch_size=int(pre.patch_size), mean=float(pre.mean), std=float(pre.std),
                                        interpolation=bool(pre.interpolation), overlap=int(pre.overlap),
                                        masking=bool(pre.masking), derivative=bool(getattr(pre, "derivative", False))))
        else:
            raise NotImplementedError(f"modality type {mtype} is not on the accelerated path")
    if data_config[tgt]["type"] != "text":
        raise NotImplementedError("only text targets are on the accelerated path")
    pre, data, toks = preprocessors[tgt], _column(rows, tgt), []
    for lo in range(0, n, chunk):
        enc = pre(text=data[lo: lo + chunk], padding=False, truncation=True, max_length=max_target_length)
        toks += [np.asarray(t, dtype=np.int32) for t in enc["input_ids"]]
    target = Column("tokens", pad_len=None, max_len=int(max_target_length), pad_id=int(pre.pad_token_id),
                    tokens=Ragged.from_rows(toks, np.int32))
    ds = HostDataset(cols, tgt, target, n, passthrough={"target_smiles": list(data)})
    for c in extra_columns:
        ds.passthrough[c] = _column(rows, c)
    if aligns:  # datamodules.py:151-170 (without the optional interpolation of the alignment target)
        a = np.zeros((n, 1800), dtype=np.float32)
        for i, s in enumerate(_column(rows, aligns[0])):
            a[i, : min(len(s), 1800)] = np.asarray(s, dtype=np.float32)[:1800]
        ds.alignment = a
    return ds


# ------------------------------------------------------------------------------------------------ device-resident set
class DeviceDataset:
    """`HostDataset` uploaded to HBM + batch assembly by index.  `collate` returns the tuple
    `FusedTrainer.train_step` / `Engine.forward` take: (enc_inputs {modality: batch-first tensor}, enc_mask u8 [B, S],
    dec_in int64 [B, T], dec_mask u8 [B, T], labels int64 [B, T] with -100[, align_target fp32 [B, 1800]])."""

    def __init__(self, host: HostDataset, device="cuda"):
        self.host = host
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceDataset lives in HBM: a CUDA device is required (no CPU fallback)")
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)  # noqa: E731
        self.dev: Dict[str, Dict[str, torch.Tensor]] = {}
        for name, col in list(host.columns.items()) + [("__target__", host.target)]:
            d: Dict[str, torch.Tensor] = {}
            if col.tokens is not None:
                d["tok"], d["tok_off"] = up(col.tokens.flat), up(col.tokens.offsets)
                if col.tokens.valid is not None:
                    d["valid"] = up(col.tokens.valid)
            if col.values is not None:
                d["val"], d["val_off"] = up(col.values.flat), up(col.values.offsets)
            if col.raw is not None:
                d["raw"], d["missing"] = up(col.raw), up(col.missing)
            self.dev[name] = d
        self.align = None if host.alignment is None else up(host.alignment)
        # pinned staging slots for the index vectors; a slot is rewritten only after the H2D copy that last read it has
        # completed (the host may run many steps ahead of the GPU when the step is a graph replay)
        self._idx_ring = [torch.empty(0, dtype=torch.int32).pin_memory() for _ in range(8)]
        self._idx_done: List[Optional[torch.cuda.Event]] = [None] * len(self._idx_ring)
        self._ring_pos = 0

    def __len__(self):
        return self.host.n

    def bytes_resident(self) -> int:
        return sum(t.numel() * t.element_size() for d in self.dev.values() for t in d.values()) + \
            (0 if self.align is None else self.align.numel() * 4)

    def _indices(self, indices) -> Tuple[np.ndarray, torch.Tensor]:
        idx = np.ascontiguousarray(np.asarray(indices, dtype=np.int32))
        if idx.ndim != 1 or idx.size == 0:
            raise ValueError("indices must be a non-empty 1-D sequence")
        if idx.min() < 0 or idx.max() >= self.host.n:
            raise IndexError("sample index out of range")
        slot = self._ring_pos = (self._ring_pos + 1) % len(self._idx_ring)
        if self._idx_done[slot] is not None:
            self._idx_done[slot].synchronize()
        if self._idx_ring[slot].numel() < idx.size:
            self._idx_ring[slot] = torch.empty(max(idx.size, 1024), dtype=torch.int32).pin_memory()
        stage = self._idx_ring[slot][: idx.size]
        stage.copy_(torch.from_numpy(idx))
        rows = stage.to(self.device, non_blocking=True)
        ev = self._idx_done[slot] or torch.cuda.Event()
        ev.record()
        self._idx_done[slot] = ev
        return idx, rows

    def collate(self, indices):
        idx, rows = self._indices(indices)
        B = idx.size
        dev = self.device
        enc: Dict[str, Any] = {}
        masks: List[torch.Tensor] = []
        for name, col in self.host.columns.items():
            d = self.dev[name]
            if col.kind == "tokens":
                lens = np.minimum(col.tokens.lengths[idx], col.max_len)
                L = col.pad_len if col.pad_len is not None else int(lens.max())
                ids = torch.empty(B, L, dtype=torch.int64, device=dev)
                mask = torch.empty(B, L, dtype=torch.uint8, device=dev)
                ops.collate_tokens(d["tok"], d["tok_off"], d.get("valid"), rows, col.pad_id, col.max_len, ids, mask)
                if col.values is not None:
                    vals = torch.empty(B, L, 1, dtype=torch.float32, device=dev)
                    ops.collate_values(d["val"], d["val_off"], rows, col.pad_value, col.max_len, vals, None)
                    enc[name] = {"tokenized_input": ids, "numerical_values": vals.view(B, L)}
                else:
                    enc[name] = ids
            elif col.kind == "values":
                L = int(col.values.lengths[idx].max())
                width = col.values.flat.shape[1]
                out = torch.empty(B, L, width, dtype=torch.float32, device=dev)
                mask = torch.empty(B, L, dtype=torch.uint8, device=dev)
                ops.collate_values(d["val"], d["val_off"], rows, col.pad_value, col.max_len, out, mask)
                enc[name] = out
            else:  # patches: gather + standardise + patch in one pass over the selected spectra
                p = col.patch
                n_pts = col.raw.shape[1]
                offset, n_use = (INTERP_OFFSET, INTERP_POINTS) if p["interpolation"] else (0, n_pts)
                if p["interpolation"] and n_pts not in (1791, 1800):
                    raise ValueError(f"interpolation expects 1791 or 1800 points, got {n_pts}")
                ps = p["patch_size"]
                hop = ps // p["overlap"]
                n_patches = n_use // ps
                P = n_patches if p["overlap"] == 1 else (n_patches * ps - ps) // hop + 1
                Pd = n_patches if p.get("derivative") else 0  # patches.py:91-95: gradient patches appended
                out = torch.empty(B, P + Pd, ps, dtype=torch.float32, device=dev)
                pad = torch.empty(B, P + Pd, dtype=torch.uint8, device=dev)
                if Pd:
                    ops.patchify_deriv(d["raw"], out, p["mean"], p["std"], Pd, offset=offset, n_use=n_use, hop=hop, pad=pad,
                                       missing=d["missing"], masking=p["masking"], rows=rows)
                else:
                    ops.patchify(d["raw"], out, p["mean"], p["std"], offset=offset, hop=hop, pad=pad,
                                 missing=d["missing"], masking=p["masking"], rows=rows)
                mask = pad ^ 1  # the kernel writes the reference's pad flag; the engine wants validity
                enc[name] = out
            masks.append(mask)
        enc_mask = masks[0] if len(masks) == 1 else torch.cat(masks, dim=1)
        tcol, td = self.host.target, self.dev["__target__"]
        T = int(np.minimum(tcol.tokens.lengths[idx], tcol.max_len).max()) - 1
        dec_in = torch.empty(B, T, dtype=torch.int64, device=dev)
        dec_mask = torch.empty(B, T, dtype=torch.uint8, device=dev)
        labels = torch.empty(B, T, dtype=torch.int64, device=dev)
        ops.collate_target(td["tok"], td["tok_off"], rows, tcol.pad_id, tcol.max_len, dec_in, dec_mask, labels)
        if self.align is not None:
            return enc, enc_mask, dec_in, dec_mask, labels, self.align.index_select(0, rows.long())
        return enc, enc_mask, dec_in, dec_mask, labels

    def wire_batch(self, indices) -> Dict[str, Any]:
        """The same batch in the reference collator's wire format (datamodules.py:201-218): seq-first tensors, bool
        masks with True = pad, `target` holding pad ids (not -100), plus the passthrough columns."""
        out = self.collate(indices)
        enc, enc_mask, dec_in, dec_mask, labels = out[:5]
        pad = self.host.target.pad_id
        wire_in = {}
        for m, v in enc.items():
            if isinstance(v, dict):
                wire_in[m] = {k: t.transpose(0, 1) for k, t in v.items()}
            else:
                wire_in[m] = v.transpose(0, 1)
        batch = {
            "encoder_input": wire_in,
            "encoder_pad_mask": ~enc_mask.bool().T,
            "decoder_input": {self.host.target_modality: dec_in.T},
            "decoder_pad_mask": ~dec_mask.bool().T,
            "target": torch.where(labels == -100, torch.full_like(labels, pad), labels).T,
            "target_mask": (labels == -100).T,
        }
        for k, col in self.host.passthrough.items():
            batch[k] = [col[int(i)] for i in np.asarray(indices)]
        if len(out) > 5:
            batch["encoder_alignment_input"] = out[5]
        return batch


# ------------------------------------------------------------------------------------------------------- index stream
class IndexSampler:
    """Per-epoch sample order.  `shuffle=True` mirrors DataLoader(shuffle=True) (datamodules.py:425-431); with
    world > 1 rank r takes every world-th index of the (padded) epoch permutation, torch DistributedSampler's rule,
    which is what Lightning DDP installs (trainer/trainer.py:58-71)."""

    def __init__(self, n: int, batch_size: int, shuffle: bool = True, seed: int = 3247, rank: int = 0, world: int = 1,
                 drop_last: bool = False):
        if not 0 <= rank < world:
            raise ValueError("rank must be in [0, world)")
        self.n, self.bs, self.shuffle, self.seed = n, batch_size, shuffle, seed
        self.rank, self.world, self.drop_last = rank, world, drop_last
        self.epoch = 0

    def set_epoch(self, epoch: int):
        self.epoch = epoch

    def indices(self) -> np.ndarray:
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            order = torch.randperm(self.n, generator=g).numpy()
        else:
            order = np.arange(self.n)
        if self.world > 1:
            total = -(-self.n // self.world) * self.world
            if total > self.n:  # pad by wrapping around, as DistributedSampler does
                order = np.concatenate([order, order[: total - self.n]])
            order = order[self.rank:: self.world]
        return order

    def __len__(self):
        m = len(self.indices())
        return m // self.bs if self.drop_last else -(-m // self.bs)

    def __iter__(self) -> Iterator[np.ndarray]:
        order = self.indices()
        for lo in range(0, len(order), self.bs):
            chunk = order[lo: lo + self.bs]
            if len(chunk) < self.bs and self.drop_last:
                return
            yield chunk


class DeviceLoader:
    """Iterates engine-format batches (or wire-format dicts with `wire=True`) assembled on the device."""

    def __init__(self, dataset: DeviceDataset, batch_size: int, shuffle: bool = True, seed: int = 3247, rank: int = 0,
                 world: int = 1, drop_last: bool = False, wire: bool = False):
        self.ds, self.wire = dataset, wire
        self.sampler = IndexSampler(len(dataset), batch_size, shuffle, seed, rank, world, drop_last)

    def set_epoch(self, epoch: int):
        self.sampler.set_epoch(epoch)

    def __len__(self):
        return len(self.sampler)

    def __iter__(self):
        for idx in self.sampler:
            yield self.ds.wire_batch(idx) if self.wire else self.ds.collate(idx)
