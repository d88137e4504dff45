"""Golden vectors for the patch preprocessor: runs the UNMODIFIED reference class
(/root/reference/src/analytical_fm/data/preprocessing/patches.py) on the bundled IR parquet and on synthetic
spectra, for the configurations the reference's data yamls use.  Run in the build container:
    python tests/golden/make_patches_golden.py      -> tests/golden/patches.pt
"""
import os
import sys

import numpy as np
import pandas as pd
import torch

sys.path.insert(0, "/root/reference/src")
from analytical_fm.data.preprocessing.patches import PatchPreprocessor  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
df = pd.read_parquet("/root/reference/tests/test_data/ir_dataset/ir_data.parquet")
col = [c for c in df.columns if "ir" in c.lower()][0]
real = [list(map(float, s)) for s in df[col].tolist()]
rng = np.random.default_rng(3247)
synth1800 = (rng.random((6, 1800)) * (rng.random((6, 1800)) > 0.2)).tolist()
cases = []
for name, spectra, kw in (
    ("ir_patch125", real, dict(patch_size=125, masking=False, interpolation=False)),
    ("ir_interp_patch75", real, dict(patch_size=75, masking=False, interpolation=True)),
    ("ir_patch75_overlap3", real[:8], dict(patch_size=75, masking=False, interpolation=False, overlap=3)),
    ("mix1800_interp_patch75_masking", synth1800, dict(patch_size=75, masking=True, interpolation=True)),
    ("with_missing", [real[0], None, real[2]], dict(patch_size=150, masking=False, interpolation=False)),
    # derivative=True (patches.py:91-95): gradient patches appended; incl. a patch size that divides the spectrum (the
    # one-sided difference at the last point lands in a patch), interpolation, masking, a missing spectrum
    ("ir_patch125_derivative", real, dict(patch_size=125, masking=False, interpolation=False, derivative=True)),
    ("mix1800_patch75_derivative_exact", synth1800, dict(patch_size=75, masking=True, interpolation=False, derivative=True)),
    ("ir_interp_patch65_derivative", real[:6] + [None], dict(patch_size=65, masking=False, interpolation=True, derivative=True)),
):
    pp = PatchPreprocessor(**kw)
    pp.initialise({"m": [s for s in spectra if s is not None]}, "m")
    inp = [None if s is None else list(s) for s in spectra]
    patches, mask = pp([None if s is None else list(s) for s in spectra])
    cases.append(dict(name=name, kwargs=kw, spectra=inp, mean=float(pp.mean), std=float(pp.std),
                      patches=patches.clone(), mask=mask.clone()))
    print(name, tuple(patches.shape), tuple(mask.shape), float(pp.mean), float(pp.std))
torch.save(cases, os.path.join(HERE, "patches.pt"))
