"""CPU restatement (numpy) of the reference's formula-guided logits processor.  TEST INFRASTRUCTURE ONLY: imported by
tests/ (and nothing in the product path).

Follows `analytical_fm/generation/logit_processors.py`:
  * atom list and the token -> element table            :26-62   (substring rule, "Cl" is not carbon, "H" skipped)
  * formula string -> count vector                      :72-87
  * __call__: decode, canonicalise, count, three writes :89-152
Parity pinned: tests/golden/guided_c1.pt holds outputs of the UNMODIFIED reference processor (direct calls and whole
greedy / beam generations) with rdkit's three functions replaced by tests/toy_chem.py; tests/test_guided.py replays them.
The chemistry itself (rdkit) is a third-party dependency that is absent here and is injected, never restated.
"""
import re
from typing import Dict, List

import numpy as np
import torch

ATOMS = ["C", "N", "O", "S", "P", "F", "Cl", "Br", "I", "B", "Si", "H", "Se", "As"]
SPECIAL = ("<bos>", "<unk>", "<eos>", "<pad>")
N_CHECK = 9  # look-ahead compares C..I only (logit_processors.py:148-149)


def token_atoms(vocab: Dict[str, int]) -> Dict[int, List[int]]:
    """logit_processors.py:42-62."""
    table: Dict[int, List[int]] = {i: [] for i in range(len(ATOMS))}
    for token, tid in vocab.items():
        if token in SPECIAL:
            continue
        for i, atom in enumerate(ATOMS):
            if atom == "H":
                continue
            if atom.lower() in token.lower():
                if atom.lower() == "c" and token.lower() == "cl":
                    continue
                table[i].append(tid)
    return table


def formula_counts(formula: str) -> np.ndarray:
    """logit_processors.py:72-87."""
    out = np.zeros(len(ATOMS))
    for atom, count in re.findall(r"([A-Z][a-z]?)(\d*)", formula):
        out[ATOMS.index(atom)] = int(count) if count else 1
    return out


class GuidedOracle:
    def __init__(self, n_beams, formulas, vocab, eos_id, chem):
        self.vocab = dict(vocab)
        self.id2tok = {i: t for t, i in self.vocab.items()}
        self.eos_id, self.chem = eos_id, chem
        self.table = token_atoms(self.vocab)
        self.target = np.repeat(np.stack([formula_counts(f) for f in formulas]), n_beams, axis=0)

    def decode(self, row) -> str:
        return "".join(self.id2tok[int(t)] for t in row if self.id2tok[int(t)] not in SPECIAL)

    def counts(self, input_ids) -> np.ndarray:
        rows = []
        for row in input_ids.tolist():
            canon = self.chem.canonical(self.decode(row))
            canon = canon if canon else ""       # :107-110 (invalid -> "")
            try:
                f = self.chem.formula(canon)     # :113-118
            except Exception:  # noqa: BLE001
                f = ""
            rows.append(formula_counts(f))
        return np.stack(rows)

    def __call__(self, input_ids, scores):
        cur = self.counts(input_ids)
        V = scores.shape[1]
        scores[torch.from_numpy(np.all(self.target == cur, axis=1)), self.eos_id] = 0            # :123-124
        scores[torch.from_numpy(np.any(cur < self.target, axis=1)), self.eos_id] = -float("inf")  # :127-128
        nxt = np.repeat(cur[:, None, :], V, axis=1)                                               # :131-146
        for a, ids in self.table.items():
            nxt[:, ids, a] += 1
        too_large = np.any(nxt[:, :, :N_CHECK] > self.target[:, None, :N_CHECK], axis=2)          # :149
        scores[torch.from_numpy(too_large)] = -float("inf")                                       # :150
        return scores
