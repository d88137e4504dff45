"""CPU restatement (numpy) of the reference collator's batch assembly, starting from the pre-tokenised ragged columns
(TEST INFRASTRUCTURE ONLY - never imported by product code).

Follows `MultiModalDataCollator.__call__` / `prepare_encoder_input` / `prepare_target`
(analytical_fm/data/datamodules.py:140-361): text inputs padded to `max_source_length` (:238-252), tokenised peak lists
padded to the batch's longest row with fully-masked empty samples (:254-275; carbon.py:52-56), XVal values padded with
1.0 (multiplets.py:199-230), MS/MS peak lists padded with zeros (msms_number.py:50-80), patches (:334-341 ->
patch_oracle), masks concatenated in modality order (:343-349), target padded to the batch's longest row, decoder
input = tokens[:-1], target = tokens[1:] (:178-206).  Output is the collator's wire format (seq-first, True = pad).
Parity pinned by tests/golden/collate.pt (batches produced by the reference collator itself).
"""
import numpy as np
import torch

from oracle import patch_oracle


def _pad_rows(rows, L, fill, dtype, width=0):
    shape = (len(rows), L, width) if width else (len(rows), L)
    out = np.full(shape, fill, dtype=dtype)
    for b, r in enumerate(rows):
        n = min(len(r), L)
        out[b, :n] = r[:n]
    return out


def collate(host, indices):
    """host: pipeline.HostDataset (plain numpy containers); indices: sample ids of the batch."""
    idx = [int(i) for i in indices]
    enc, masks = {}, []
    for name, col in host.columns.items():
        if col.kind == "tokens":
            rows = [col.tokens.row(i)[: col.max_len] for i in idx]
            L = col.pad_len if col.pad_len is not None else max(len(r) for r in rows)
            ids = _pad_rows(rows, L, col.pad_id, np.int64)
            valid = np.ones(len(idx), bool) if col.tokens.valid is None else col.tokens.valid[idx].astype(bool)
            pad = ~((np.arange(L)[None, :] < np.array([min(len(r), L) for r in rows])[:, None]) & valid[:, None])
            if col.values is not None:
                vals = _pad_rows([col.values.row(i)[: col.max_len, 0] for i in idx], L, col.pad_value, np.float32)
                enc[name] = {"tokenized_input": torch.from_numpy(ids.T.copy()),
                             "numerical_values": torch.from_numpy(vals.T.copy())}
            else:
                enc[name] = torch.from_numpy(ids.T.copy())
        elif col.kind == "values":
            rows = [col.values.row(i) for i in idx]
            L = max(len(r) for r in rows)
            width = col.values.flat.shape[1]
            x = _pad_rows(rows, L, col.pad_value, np.float32, width=width)
            pad = ~(np.arange(L)[None, :] < np.array([len(r) for r in rows])[:, None])
            enc[name] = torch.from_numpy(np.transpose(x, (1, 0, 2)).copy())
        else:
            p = col.patch
            spectra = [None if col.missing[i] else col.raw[i].tolist() for i in idx]
            if all(s is None for s in spectra):  # the table width stands in for the reference's 500-point default
                spectra = [[0.0] * col.raw.shape[1] for _ in idx]
                patches, _ = patch_oracle.patch_preprocess(spectra, p["mean"], p["std"], p["patch_size"], p["masking"],
                                                           p["interpolation"], p["overlap"], p.get("derivative", False))
                pad = np.ones(patches.shape[:2], bool) if not p["masking"] else patches.sum(-1) == 0
            else:
                patches, pad = patch_oracle.patch_preprocess(spectra, p["mean"], p["std"], p["patch_size"], p["masking"],
                                                             p["interpolation"], p["overlap"], p.get("derivative", False))
            enc[name] = torch.from_numpy(np.transpose(patches, (1, 0, 2)).copy())
        masks.append(pad.T)
    tcol = host.target
    rows = [tcol.tokens.row(i)[: tcol.max_len] for i in idx]
    L = max(len(r) for r in rows)
    ids = _pad_rows(rows, L, tcol.pad_id, np.int64).T
    pad = ~(np.arange(L)[None, :] < np.array([len(r) for r in rows])[:, None]).T
    batch = {
        "encoder_input": enc,
        "encoder_pad_mask": torch.from_numpy(np.concatenate(masks, axis=0)),
        "decoder_input": {host.target_modality: torch.from_numpy(ids[:-1].copy())},
        "decoder_pad_mask": torch.from_numpy(pad[:-1].copy()),
        "target": torch.from_numpy(ids[1:].copy()),
        "target_mask": torch.from_numpy(pad[1:].copy()),
    }
    for k, colv in host.passthrough.items():
        batch[k] = [colv[i] for i in idx]
    return batch
