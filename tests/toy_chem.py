"""A deterministic stand-in for the three rdkit calls the reference's guided decoding makes
(`Chem.MolFromSmiles`, `Chem.MolToSmiles`, `rdMolDescriptors.CalcMolFormula`;
generation/logit_processors.py:107-121, modeling/wrapper.py:546-550).

rdkit is not installable here (no network), so the golden vectors of the guided path are produced by the UNMODIFIED
reference processor with these functions patched in for rdkit (tests/golden/make_guided_golden.py), and the tests
inject the same functions into the product as its chemistry backend.  What the vectors pin is therefore everything
around the chemistry: token -> element table, <eos> forcing / banning, the look-ahead mask, processor ordering
against ForcedEOS, and the interplay with beam search.  The rules below are NOT chemistry: a syntactic validity
check (no dangling bond, no empty or unopened branch) and a plain element count.
"""
import re

_ELEM = re.compile(r"Cl|Br|Si|Se|As|H\d*|[CNOSPFIB]|[cnosp]")
_ORDER = ["C", "H", "N", "O", "S", "P", "F", "Cl", "Br", "I", "B", "Si", "Se", "As"]


class ToyMol:
    def __init__(self, smiles):
        self.smiles = smiles


def is_valid(s: str) -> bool:
    if s == "":
        return True  # rdkit parses "" into an empty molecule
    if s[-1] in "(=#" or "()" in s:
        return False  # dangling bond / open branch, empty branch
    depth = 0
    for ch in s:
        depth += (ch == "(") - (ch == ")")
        if depth < 0:
            return False
    return True


def mol_from_smiles(s):
    return ToyMol(s) if isinstance(s, str) and is_valid(s) else None


def mol_to_smiles(m, **kw):
    return m.smiles


def calc_mol_formula(m) -> str:
    if m is None:
        raise TypeError("no molecule")
    counts = {}
    for tok in _ELEM.findall(m.smiles):
        if tok[0] == "H":
            counts["H"] = counts.get("H", 0) + (int(tok[1:]) if len(tok) > 1 else 1)
        else:
            el = tok if tok[0].isupper() else tok.upper()
            counts[el] = counts.get(el, 0) + 1
    return "".join(f"{el}{counts[el] if counts[el] > 1 else ''}" for el in _ORDER if counts.get(el))


class ToyChem:
    """Chemistry backend in the shape `multimodalanalytical_b200.guided` expects."""

    def canonical(self, smiles: str):
        m = mol_from_smiles(smiles)
        return mol_to_smiles(m) if m is not None else None

    def formula(self, smiles: str) -> str:
        return calc_mol_formula(mol_from_smiles(smiles))


def patch_rdkit_stub():
    """Route the stub `rdkit` modules installed by tests/golden/_ref_stubs.py to the toy rules."""
    import sys

    chem = sys.modules["rdkit.Chem"]
    chem.MolFromSmiles = mol_from_smiles
    chem.MolToSmiles = mol_to_smiles
    sys.modules["rdkit.Chem.rdMolDescriptors"].CalcMolFormula = calc_mol_formula
