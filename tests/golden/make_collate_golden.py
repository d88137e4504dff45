"""Golden batches for the collator -> device pipeline, produced by the UNMODIFIED reference
collator `MultiModalDataCollator` (analytical_fm/data/datamodules.py:18-399) and its own preprocessors
(`load_preprocessors`, data/data_utils.py:40-126).

  c1     the bundled IR parquet (Formula text + IR patches -> Smiles), same pipeline as make_golden.case_c1
  multi  a synthetic multimodal set: Formula text, 13C peak lists (carbon), 1H multiplets as text and as XVal numerical
         encoding, MS/MS peak lists (msms_number), IR with interpolation; some samples lack a modality (None)
  spectext  spectra written out as text (the legacy ablation inputs): text_spectrum with a formula prefix and integer
         intensities, text_spectrum with XVal numerical encoding (spectra only), run_length_encoding; plus IR patches
         with derivative=True (gradient patches appended, patches.py:91-95)
For each case the fixture holds the reference batches for several index lists and the `HostDataset` that
`multimodalanalytical_b200.pipeline.pretokenise` extracted from the same preprocessor objects (the reference cannot
travel to the GPU box); for c1 also the raw rows and the tokenizers' JSON so a CPU test can re-run `pretokenise`.

    python tests/golden/make_collate_golden.py        # rewrites tests/golden/collate.pt
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

import pytorch_lightning as pl  # noqa: E402  (stub)
from analytical_fm.data import data_utils, datamodules, datasets  # noqa: E402
from datasets import Dataset  # noqa: E402

from multimodalanalytical_b200.pipeline import pretokenise  # noqa: E402

REF = "/root/reference"


def _clone(x):
    if isinstance(x, torch.Tensor):
        return x.clone()
    if isinstance(x, dict):
        return {k: _clone(v) for k, v in x.items()}
    return x


def reference_batches(collator, ds, index_lists):
    out = []
    for idx in index_lists:
        batch = collator([ds[int(i)] for i in idx])
        out.append({"indices": list(idx), "batch": {k: _clone(v) for k, v in batch.items() if v is not None}})
    return out


def case_c1():
    pl.seed_everything(3247)
    dc = yaml.safe_load(open(f"{REF}/configs/data/ir/patches.yaml"))
    data_config, dsd = datasets.build_dataset_multimodal(
        dc, data_path=f"{REF}/tests/test_data/ir_dataset", splitting="random", cv_split=0,
        augment_config=None, num_cpu=1, mixture_config=None)
    np.random.seed(3247)
    data_config, pre = data_utils.load_preprocessors(dsd["train"], data_config)
    dm = datamodules.MultiModalDataModule(dataset=dsd, preprocessors=pre, data_config=data_config,
                                          model_type="CustomModel", batch_size=16, num_workers=0, extra_columns=[None])
    ds = dsd["train"]
    n = len(ds)
    lists = [list(range(n)), [3, 1, 7, 2], [n - 1], [5, 5, 0, 9, 14, 2, 11]]
    col = dm.collator
    host = pretokenise(ds, pre, data_config, col.max_source_length, col.max_target_length)
    rows = {k: list(ds[k]) for k in ("Formula", "IR", "Smiles")}
    return {
        "data_config": data_config, "batches": reference_batches(col, ds, lists), "host": host, "rows": rows,
        "max_source_length": dict(col.max_source_length), "max_target_length": int(col.max_target_length),
        "tokenizers": {m: pre[m].backend_tokenizer.to_str() for m in ("Formula", "Smiles")},
        "patch": {k: getattr(pre["IR"], k) for k in ("patch_size", "masking", "interpolation", "overlap", "derivative")}
                 | {"mean": float(pre["IR"].mean), "std": float(pre["IR"].std)},
    }


def case_multi():
    rng = np.random.default_rng(11)
    n = 24
    atoms = ["C", "c", "N", "O", "(", ")", "=", "1", "Cl", "Br", "S", "n", "F"]
    cats = ["s", "d", "t", "q", "m", "dd"]
    rows = {k: [] for k in ("Formula", "Carbon", "Multiplets", "MultipletsNum", "MSMS", "IR", "Smiles")}
    for i in range(n):
        rows["Formula"].append(f"C{rng.integers(2, 30)}H{rng.integers(2, 40)}" + ("N2" if i % 3 == 0 else "") +
                               ("O" if i % 2 else "ClBr"))
        rows["Carbon"].append(None if i in (3, 10) else [
            {"delta (ppm)": float(rng.uniform(5, 200)), "intensity": float(rng.uniform(1, 50))}
            for _ in range(int(rng.integers(1, 14)))])
        mult = None if i == 5 else [
            {"rangeMax": float(rng.uniform(0.5, 12)), "rangeMin": float(rng.uniform(0.1, 0.5)),
             "category": str(rng.choice(cats)), "nH": int(rng.integers(1, 4)),
             "j_values": "_".join(f"{rng.uniform(1, 16):.2f}" for _ in range(int(rng.integers(0, 3)))) or "None"}
            for _ in range(int(rng.integers(1, 7)))]
        rows["Multiplets"].append(mult)
        rows["MultipletsNum"].append(mult if mult is not None else [
            {"rangeMax": 1.0, "rangeMin": 0.5, "category": "s", "nH": 1, "j_values": "None"}])
        rows["MSMS"].append([[float(rng.uniform(50, 500)), float(rng.uniform(0, 100) if rng.random() > 0.2 else 0.3)]
                             for _ in range(int(rng.integers(2, 12)))] + [[123.4, 55.0]])
        rows["IR"].append(None if i == 7 else (rng.random(1791) * (rng.random(1791) > 0.05)).astype(np.float32).tolist())
        rows["Smiles"].append("".join(rng.choice(atoms) for _ in range(int(rng.integers(3, 40)))))
    ds = Dataset.from_dict(rows)
    # PatchPreprocessor.initialise (patches.py:37-39) cannot digest a missing spectrum: fit on a complete copy
    fit_rows = dict(rows)
    fit_rows["IR"] = [r if r is not None else rows["IR"][0] for r in rows["IR"]]
    ds_fit = Dataset.from_dict(fit_rows)
    dc = {
        "Formula": {"type": "text", "target": False,
                    "preprocessor_arguments": {"tokenizer_regex": "([A-Z]{1}[a-z]?[0-9]*)"}},
        "Carbon": {"type": "carbon", "target": False, "preprocessor_arguments": {"intensities": True}},
        "Multiplets": {"type": "multiplets", "target": False,
                       "preprocessor_arguments": {"encoding": "text", "j_values": True}},
        "MultipletsNum": {"type": "multiplets", "target": False,
                          "preprocessor_arguments": {"encoding": "numerical_encoding", "j_values": True,
                                                     "normalise": True}},
        "MSMS": {"type": "msms_number", "target": False, "preprocessor_arguments": None},
        "IR": {"type": "1D_patches", "target": False,
               "preprocessor_arguments": {"patch_size": 75, "interpolation": True, "masking": False}},
        "Smiles": {"type": "text", "target": True,
                   "preprocessor_arguments": {"tokenizer_regex": yaml.safe_load(
                       open(f"{REF}/configs/data/ir/patches.yaml"))["Smiles"]["preprocessor_arguments"]["tokenizer_regex"]}},
    }
    np.random.seed(7)
    data_config, pre = data_utils.load_preprocessors(ds_fit, dc)
    np.random.seed(8)
    col = datamodules.MultiModalDataCollator(preprocessors=pre, data_config=data_config, model_type="CustomModel",
                                             dataset={"train": ds_fit}, extra_columns=[None])
    lists = [list(range(n)), [3, 10, 5, 7], [1, 2], [23, 0, 11, 7, 3], [10, 3]]
    host = pretokenise(ds, pre, data_config, col.max_source_length, col.max_target_length)
    return {"data_config": data_config, "batches": reference_batches(col, ds, lists), "host": host,
            "max_source_length": dict(col.max_source_length), "max_target_length": int(col.max_target_length)}


def case_spectext():
    rng = np.random.default_rng(5)
    n = 14
    rows = {"Formula": [f"C{rng.integers(2, 30)}H{rng.integers(2, 40)}" + ("N2" if i % 3 == 0 else "O") for i in range(n)],
            "IR": [np.round(rng.random(1791) * (rng.random(1791) > 0.3), 3).astype(np.float32).tolist() for _ in range(n)],
            "Smiles": ["".join(rng.choice(list("CcNO()=1")) for _ in range(int(rng.integers(3, 30)))) for _ in range(n)]}
    rows["RLE"] = [np.repeat(rng.random(200), 9)[:1791].astype(np.float32).tolist() for _ in range(n)]  # runs to encode
    rows["IRd"] = rows["IR"]
    ds = Dataset.from_dict(rows)
    cols = {"spectra_column": "IR", "formula_column": "Formula"}
    dc = {
        "SpecText": {"type": "text_spectrum", "target": False, "spectra_only": False, **cols,
                     "preprocessor_arguments": {"spectrum_tokens_x": 40, "spectrum_tokens_y": 20, **cols}},
        "SpecNum": {"type": "text_spectrum", "target": False, "spectra_only": True, **cols,
                    "preprocessor_arguments": {"spectrum_tokens_x": 30, "spectrum_to_text_y": "numerical_encoding",
                                               "spectra_only": True, **cols}},
        "RLE": {"type": "run_length_encoding", "target": False,
                "preprocessor_arguments": {"spectrum_tokens_x": 25, "spectrum_tokens_y": 6}},
        "IRd": {"type": "1D_patches", "target": False,
                "preprocessor_arguments": {"patch_size": 125, "interpolation": False, "masking": False, "derivative": True}},
        "Smiles": {"type": "text", "target": True, "preprocessor_arguments": {"tokenizer_regex": "(.)"}},
    }
    np.random.seed(9)
    data_config, pre = data_utils.load_preprocessors(ds, dc)
    col = datamodules.MultiModalDataCollator(preprocessors=pre, data_config=data_config, model_type="CustomModel",
                                             dataset={"train": ds}, extra_columns=[None])
    lists = [list(range(n)), [3, 1, 7], [n - 1], [5, 5, 0, 9, 13, 2]]
    host = pretokenise(ds, pre, data_config, col.max_source_length, col.max_target_length)
    return {"data_config": data_config, "batches": reference_batches(col, ds, lists), "host": host,
            "max_source_length": dict(col.max_source_length), "max_target_length": int(col.max_target_length)}


CASES = {"c1": case_c1, "multi": case_multi, "spectext": case_spectext}


def main():
    # `python make_collate_golden.py spectext` regenerates only the named cases and keeps the others as committed
    path = os.path.join(HERE, "collate.pt")
    names = [a for a in sys.argv[1:] if a in CASES]
    out = torch.load(path, weights_only=False) if names and os.path.exists(path) else {}
    for name in names or list(CASES):
        out[name] = CASES[name]()
    torch.save(out, path)
    for name, fx in out.items():
        b = fx["batches"][0]["batch"]
        shapes = {m: (tuple(v.shape) if isinstance(v, torch.Tensor) else {k: tuple(t.shape) for k, t in v.items()})
                  for m, v in b["encoder_input"].items()}
        print(name, shapes, "pad mask", tuple(b["encoder_pad_mask"].shape), "target", tuple(b["target"].shape),
              "max_src", fx["max_source_length"], "max_tgt", fx["max_target_length"])
    print(f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
