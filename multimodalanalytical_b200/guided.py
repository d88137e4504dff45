"""Formula-guided decoding and rejection sampling.

Reference: `GuidedFormulaProcessor` (analytical_fm/generation/logit_processors.py:12-152), built in
`HFWrapper.predict_step` when `guided_generation` is set (modeling/wrapper.py:546-556) and handed to transformers'
`generate(logits_processor=[...])`; `reject_sample` / `clean_sample` (analytical_fm/utils.py:22-83).

Split of the work here:
  * host, once per batch: formula strings -> target count vectors; vocabulary -> token/element bit table
    (the reference's substring rule, logit_processors.py:46-62);
  * host, once per step: running hypotheses -> canonical SMILES -> element counts.  That is rdkit's job in the
    reference (a third-party dependency, absent in this image); it sits behind `ChemBackend` and is memoised per
    decoded string, because the K beams of a spectrum and consecutive steps keep asking about the same prefixes;
  * device, every step: the three writes of `__call__` (<eos> := 0 on a formula match, <eos> := -inf while an
    element is short, token := -inf when it would overshoot one of the first nine elements) are fused into the
    beam / greedy step kernels (csrc/decode.cu `guide_row` / `guide_apply`), so the [rows, V] score matrix is never
    materialised and nothing but [rows, 14] counts crosses PCIe.
The object is also a plain `(input_ids, scores) -> scores` processor (`__call__`, CUDA tensors) so it composes with
other processors in a list, in which case `mma_guided_mask` edits the dense scores in place.
"""
from __future__ import annotations

import re
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops

ATOM_LIST = ["C", "N", "O", "S", "P", "F", "Cl", "Br", "I", "B", "Si", "H", "Se", "As"]  # logit_processors.py:26-41
SPECIAL_TOKENS = ("<bos>", "<unk>", "<eos>", "<pad>")                                    # logit_processors.py:52
N_CHECK = 9  # the look-ahead stops before B / Si / H / Se / As (logit_processors.py:148-149)
_FORMULA_RE = re.compile(r"([A-Z][a-z]?)(\d*)")


class ChemBackend:
    """The two chemistry questions guided decoding asks.  `canonical` returns None for an unparsable string."""

    def canonical(self, smiles: str) -> Optional[str]:
        raise NotImplementedError

    def formula(self, smiles: str) -> str:
        raise NotImplementedError


class RDKitChem(ChemBackend):
    """rdkit, exactly as the reference calls it (logit_processors.py:107-118).  Import is deferred and loud."""

    def __init__(self):
        try:
            from rdkit import Chem, RDLogger
            from rdkit.Chem import rdMolDescriptors
        except ImportError as e:  # no silent degradation: guided decoding without chemistry is meaningless
            raise ImportError("guided generation / rejection sampling need rdkit (or pass chem_backend=...)") from e
        RDLogger.DisableLog("rdApp.*")
        self._chem, self._desc = Chem, rdMolDescriptors

    def canonical(self, smiles):
        mol = self._chem.MolFromSmiles(smiles)
        return self._chem.MolToSmiles(mol) if mol else None

    def formula(self, smiles):
        return self._desc.CalcMolFormula(self._chem.MolFromSmiles(smiles))


def formula_counts(formula: str) -> List[int]:
    """`make_formula_encoding` (logit_processors.py:72-87): unknown elements raise ValueError, as there."""
    out = [0] * len(ATOM_LIST)
    for atom, count in _FORMULA_RE.findall(formula):
        out[ATOM_LIST.index(atom)] = int(count) if count else 1
    return out


def token_atom_bits(vocab: Dict[str, int], vocab_size: int) -> List[int]:
    """Bit e of entry t is set when token t adds one atom of ATOM_LIST[e] (logit_processors.py:46-62): case-blind
    substring match, "Cl" does not count as carbon, hydrogen is never counted, special tokens are skipped."""
    bits = [0] * vocab_size
    for token, tid in vocab.items():
        if token in SPECIAL_TOKENS:
            continue
        low = token.lower()
        for e, atom in enumerate(ATOM_LIST):
            if atom == "H":
                continue
            if atom.lower() in low and not (atom == "C" and low == "cl"):
                bits[tid] |= 1 << e
    return bits


def _vocab_of(tokenizer) -> Dict[str, int]:
    v = getattr(tokenizer, "vocab", None)
    if v is None:
        v = tokenizer.get_vocab()
    return dict(v)


class GuidedFormulaProcessor:
    """Same constructor as the reference's processor (n_beams, chemical_formula, target_tokenizer) plus the
    chemistry backend.  `chemical_formula` holds one formula string per spectrum of the batch."""

    def __init__(self, n_beams: int, chemical_formula: Sequence[str], target_tokenizer, chem: Optional[ChemBackend] = None):
        self.n_beams = int(n_beams)
        self.target_tokenizer = target_tokenizer
        self.eos_token_id = target_tokenizer.eos_token_id
        self.vocab_size = target_tokenizer.vocab_size
        self.chem = chem if chem is not None else RDKitChem()
        vocab = _vocab_of(target_tokenizer)
        self._piece = [""] * self.vocab_size  # id -> text; specials decode to nothing (skip_special_tokens=True)
        for token, tid in vocab.items():
            if token not in SPECIAL_TOKENS and tid < self.vocab_size:
                self._piece[tid] = token
        self.tok_atoms = torch.tensor(token_atom_bits(vocab, self.vocab_size), dtype=torch.int32)
        self.target_counts = torch.tensor([formula_counts(f) for f in chemical_formula], dtype=torch.int32)
        self._memo: Dict[str, List[int]] = {}
        self._dev: Dict[Any, Any] = {}
        # incremental state of the fused decode loop: text of every running hypothesis, and the memo as an id -> counts
        # table so a step's [rows, 14] matrix is one vectorised gather
        self._texts: List[str] = []
        self._ids: Dict[str, int] = {}
        self._table = np.zeros((256, len(ATOM_LIST)), dtype=np.int32)

    # ---------------------------------------------------------------------------------------- host chemistry
    def _counts_of(self, text: str) -> List[int]:
        c = self._memo.get(text)
        if c is None:
            canon = self.chem.canonical(text) or ""       # invalid -> "" (logit_processors.py:107-110)
            try:
                c = formula_counts(self.chem.formula(canon))  # :113-121
            except Exception:  # noqa: BLE001  (the reference swallows everything here, :116-117)
                c = [0] * len(ATOM_LIST)
            self._memo[text] = c
        return c

    def counts(self, input_ids: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """input_ids: host integer tensor [rows, cur_len] -> int32 [rows, 14] element counts of each hypothesis."""
        piece = self._piece
        rows = [self._counts_of("".join(piece[t] for t in row)) for row in input_ids.tolist()]
        res = torch.tensor(rows, dtype=torch.int32)
        if out is not None:
            out.copy_(res)
            return out
        return res

    # ------------------------------------------------------------------ incremental form (fused decode loop)
    def begin(self, rows: int) -> None:
        """All hypotheses are `<bos>` (decodes to nothing)."""
        self._texts = [""] * rows

    def advance(self, parent_rows: Sequence[int], next_tokens: Sequence[int]) -> None:
        """One search step: row r now continues row parent_rows[r] with token next_tokens[r] (what the step kernel
        wrote to `parent_row` / `next_tok`).  O(rows) per step instead of re-decoding [rows, cur_len] ids."""
        old, piece = self._texts, self._piece
        self._texts = [old[p] + piece[t] for p, t in zip(parent_rows, next_tokens)]

    def _row_id(self, text: str) -> int:
        i = self._ids.get(text)
        if i is None:
            i = self._ids[text] = len(self._ids)
            if i >= self._table.shape[0]:
                self._table = np.concatenate([self._table, np.zeros_like(self._table)], axis=0)
            self._table[i] = self._counts_of(text)
        return i

    def counts_current(self, out: torch.Tensor) -> torch.Tensor:
        """Element counts of the current hypotheses -> `out` (host int32 [rows, 14], e.g. pinned)."""
        idx = [self._row_id(t) for t in self._texts]
        out.copy_(torch.from_numpy(self._table[idx]))
        return out

    # --------------------------------------------------------------------------------------------- device side
    def _device_tables(self, device):
        d = self._dev.get(device)
        if d is None:
            d = self._dev[device] = (self.target_counts.to(device).contiguous(), self.tok_atoms.to(device))
        return d

    def guide(self, cur_counts_dev: torch.Tensor):
        """The tuple the step kernels take (ops.beam_step / ops.greedy_step `guide=`)."""
        tgt, tok = self._device_tables(cur_counts_dev.device)
        return (cur_counts_dev, tgt, tok, N_CHECK)

    def __call__(self, input_ids: torch.Tensor, scores: torch.Tensor) -> torch.Tensor:
        """Processor protocol on CUDA tensors: rows of `scores` follow `input_ids` (row = spectrum * n_beams + beam)."""
        if not scores.is_cuda:
            raise RuntimeError("GuidedFormulaProcessor runs on the CUDA path only (no CPU fallback)")
        cur = self.counts(input_ids.cpu()).to(scores.device)
        ops.guided_mask(scores, self.eos_token_id, self.n_beams, self.guide(cur))
        return scores


# ------------------------------------------------------------------------------------------ rejection sampling
def clean_sample(sample: str, canonicalise: bool, chem: Optional[ChemBackend] = None) -> Optional[str]:
    """utils.py:22-41."""
    sample = sample.replace("<bos>", "").replace("<pad>", "").replace("<eos>", "").replace(" ", "")
    if canonicalise:
        sample = (chem or RDKitChem()).canonical(sample)
    return sample


def reject_sample(predictions: Dict[str, Any], molecules: bool = True, chem: Optional[ChemBackend] = None):
    """utils.py:44-83: keep the hypotheses whose formula equals the target's, left-packed, ""-padded to n_beams."""
    chem = chem or RDKitChem()
    n_beams = len(predictions["predictions"][0])
    for i in range(len(predictions["predictions"])):
        kept = []
        for p in predictions["predictions"][i]:
            sample = clean_sample(p, molecules, chem)
            try:
                if sample is None:
                    raise TypeError("unparsable prediction")
                pred_formula = chem.formula(sample)
                target_formula = chem.formula(predictions["targets"][i])
            except TypeError:
                continue
            if pred_formula == target_formula:
                kept.append(sample)
        predictions["predictions"][i] = kept + [""] * (n_beams - len(kept))
    assert len(predictions["predictions"]) == len(predictions["targets"])
    return predictions


def calc_sampling_metrics(samples: List[Any], targets: List[str], classes: Optional[List[Any]] = None,
                          molecules: bool = True, chem: Optional[ChemBackend] = None) -> Dict[Any, Any]:
    """Top-N accuracies of ranked hypotheses (analytical_fm/utils.py:86-153): rank = position of the cleaned target
    among the cleaned predictions (n_beams when absent), `Top-(i+1)` = share of samples with rank <= i; with `classes`
    one such dict per class value.  `molecules=True` canonicalises both sides through the chemistry backend."""
    n_beams = len(samples[0])
    if molecules and chem is None:
        chem = RDKitChem()
    ranks = []
    for preds, tgt in zip(samples, targets):
        clean = [clean_sample(p, molecules, chem) for p in preds]
        t = clean_sample(tgt, molecules, chem)
        ranks.append(clean.index(t) if t in clean else n_beams)
    metrics: Dict[Any, Any] = {}
    for i in range(n_beams):
        if classes:
            for cl in dict.fromkeys(classes):
                sel = [r for r, c in zip(ranks, classes) if c == cl]
                metrics.setdefault(float(cl), {})[f"Top-{i + 1}"] = float(sum(r <= i for r in sel) / len(sel))
        else:
            metrics[f"Top-{i + 1}"] = float(sum(r <= i for r in ranks) / len(ranks))
    return metrics
