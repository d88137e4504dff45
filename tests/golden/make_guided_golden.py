"""Golden vectors for logits processors inside `generate`, produced by the UNMODIFIED reference:
`HFWrapper.generate(batch, n_beams, logits_processor=[...])` (modeling/wrapper.py:409-453) with the reference's own
`GuidedFormulaProcessor` (generation/logit_processors.py:12-152) on the c1 fixture's model, batch and tokenizer.

rdkit is absent in this image; the three rdkit calls the processor makes are routed to `tests/toy_chem.py`
(syntactic validity + element count), the same functions the tests inject into the product as chemistry backend.

    python tests/golden/make_guided_golden.py        # rewrites tests/golden/guided_c1.pt
"""
import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _ref_stubs  # noqa: E402

_ref_stubs.install()
from tests import toy_chem  # noqa: E402

toy_chem.patch_rdkit_stub()

import pytorch_lightning as pl  # noqa: E402  (stub)
from analytical_fm.data import data_utils, datamodules, datasets  # noqa: E402
from analytical_fm.generation.logit_processors import GuidedFormulaProcessor  # noqa: E402
from analytical_fm.modeling import wrapper  # noqa: E402
from transformers.generation.logits_process import LogitsProcessor  # noqa: E402

REF = "/root/reference"


class BanTokens(LogitsProcessor):
    """A second, chemistry-free processor: exercises the generic (callable) processor path."""

    def __init__(self, banned, boost, amount):
        self.banned, self.boost, self.amount = list(banned), int(boost), float(amount)

    def __call__(self, input_ids, scores):
        scores[:, self.banned] = -float("inf")
        # depends on the prefix: boost one token after an odd number of generated tokens
        if input_ids.shape[1] % 2 == 0:
            scores[:, self.boost] += self.amount
        return scores


def main():
    fx = torch.load(os.path.join(HERE, "c1_ir_tiny.pt"), weights_only=False)
    # the tokenizer is rebuilt exactly as in make_golden.case_c1 (same seeds -> same vocabulary)
    pl.seed_everything(3247)
    dc = yaml.safe_load(open(f"{REF}/configs/data/ir/patches.yaml"))
    data_config, ds = datasets.build_dataset_multimodal(
        dc, data_path=f"{REF}/tests/test_data/ir_dataset", splitting="random", cv_split=0,
        augment_config=None, num_cpu=1, mixture_config=None)
    np.random.seed(3247)
    data_config, pre = data_utils.load_preprocessors(ds["train"], data_config)
    tok = pre["Smiles"]
    assert dict(tok.get_vocab()) == fx["smiles_vocab"], "tokenizer differs from the c1 fixture"

    model = wrapper.HFWrapper(data_config=fx["data_config"], target_tokenizer=tok, num_steps=100, **fx["model_kwargs"])
    model.load_state_dict(fx["state_dict"])
    model.eval()
    batch = fx["batch"]
    formulas = [toy_chem.calc_mol_formula(toy_chem.mol_from_smiles(s)) for s in batch["target_smiles"]]

    out = {"formulas": formulas, "target_smiles": list(batch["target_smiles"])}

    # Instrumentation only: HFWrapper.generate does not return the hypothesis scores.  The beam pool of transformers
    # starts at -1e9 and un-finished candidates are pushed to `score - 1e9`, which fp32 rounds to exactly -1e9, so
    # when fewer than K hypotheses finish the pool is padded by whatever `torch.topk` picks among ties (device- and
    # version-dependent).  The scores tell the tests which returned rows are real finished hypotheses.
    stash = {}
    hf_generate = model.hf_model.generate

    def generate_with_scores(**kw):
        if kw.get("num_beams", 1) == 1:
            return hf_generate(**kw)
        res = hf_generate(return_dict_in_generate=True, output_scores=True, **kw)
        stash["scores"] = res.sequences_scores.detach().clone()
        return res.sequences

    model.hf_model.generate = generate_with_scores

    def run(key, k, procs):
        out[key] = model.generate(batch, n_beams=k, logits_processor=procs).clone()
        if k > 1:
            out[key + "_scores"] = stash.pop("scores")

    with torch.no_grad():
        for k in (1, 3, 10):
            proc = GuidedFormulaProcessor(k, formulas, tok)
            run(f"guided_beam{k}", k, [proc])
            if k == 3:
                out["atom_id_token_id_dict"] = {a: sorted(v) for a, v in proc.atom_id_token_id_dict.items()}
                out["chemical_formula_beams"] = torch.from_numpy(proc.chemical_formula_beams.copy())
        for k in (1, 4):
            run(f"ban_beam{k}", k, [BanTokens(banned=[17, 9], boost=10, amount=1.5)])
        # both together, in list order
        run("guided_ban_beam3", 3, [GuidedFormulaProcessor(3, formulas, tok), BanTokens([17, 14], 10, 1.5)])
    # a direct call of the processor on a fixed score matrix (unit-level vector)
    g = torch.Generator().manual_seed(1)
    proc = GuidedFormulaProcessor(2, formulas[:4], tok)
    ids = torch.tensor([[2, 4, 4, 10], [2, 4, 6, 4], [2, 4, 4, 11], [2, 9, 5, 9], [2, 4, 16, 4], [2, 4, 7, 4],
                        [2, 17, 17, 17], [2, 4, 4, 4]])
    # make row 0's formula match exactly: overwrite target 0 with the toy formula of "CCO"
    proc.chemical_formula_beams[0] = proc.make_formula_encoding("C2O")
    proc.chemical_formula_beams[1] = proc.make_formula_encoding("C2O")
    scores = torch.randn(8, tok.vocab_size, generator=g)
    out["call_ids"] = ids
    out["call_formulas"] = torch.from_numpy(proc.chemical_formula_beams.copy())
    out["call_scores_in"] = scores.clone()
    out["call_scores_out"] = proc(ids, scores.clone()).clone()

    # the token -> element table on a realistic SMILES vocabulary (bracket atoms, two-letter elements, charges, ...):
    # only the constructor of the reference processor runs here (logit_processors.py:42-62)
    import types

    big = ["<pad>", "<unk>", "<bos>", "<eos>", "C", "c", "N", "n", "O", "o", "S", "s", "P", "p", "F", "Cl", "Br", "I", "B",
           "b", "(", ")", "[", "]", "=", "#", "-", "+", "/", "\\", ".", ":", "1", "2", "3", "%10", "%11", "[nH]", "[C@@H]",
           "[C@H]", "[C@]", "[C@@]", "[N+]", "[O-]", "[n+]", "[S+]", "[Si]", "[Se]", "[se]", "[As]", "[B-]", "[2H]", "[13C]",
           "[NH3+]", "[Cl-]", "[Br-]", "[I-]", "[Na+]", "[SiH]", "[PH]", "[te]", "[Te]", "[cH-]", "[Sn]", "[Cu]", "@", "H"]
    big_vocab = {t: i for i, t in enumerate(big)}
    fake = types.SimpleNamespace(vocab=big_vocab, eos_token_id=3, vocab_size=len(big))
    bp = GuidedFormulaProcessor(1, ["C2H6O", "CCl4", "C6H5Br"], fake)
    out["big_vocab"] = big_vocab
    out["big_vocab_table"] = {a: sorted(v) for a, v in bp.atom_id_token_id_dict.items()}
    out["big_vocab_formulas"] = torch.from_numpy(bp.chemical_formula_beams.copy())

    path = os.path.join(HERE, "guided_c1.pt")
    torch.save(out, path)
    for k, v in out.items():
        if isinstance(v, torch.Tensor) and v.dtype == torch.long and v.dim() == 2:
            dec = tok.batch_decode(v[:3], skip_special_tokens=True)
            real = int((out[k + "_scores"] > -1e8).sum()) if k + "_scores" in out else -1
            print(k, tuple(v.shape), "real finished rows:", real, [d.replace(" ", "") for d in dec])
    print("formulas", formulas[:5], f"{os.path.getsize(path) / 1e3:.1f} kB")


if __name__ == "__main__":
    main()
