"""Stub modules that let the UNMODIFIED reference (`/root/reference/src/analytical_fm`) import in
this container, where pytorch_lightning / omegaconf / rdkit / hydra are absent.

Only used by `make_golden.py` (golden-vector generation, run in the build container, never on the GPU
box).  Nothing in the product path imports this file.
"""
import sys
import types

import torch
from torch import nn

REFERENCE_SRC = "/root/reference/src"


def install():
    if "analytical_fm" in sys.modules:
        return
    # --- pytorch_lightning -------------------------------------------------------------------
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(nn.Module):
        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    class LightningDataModule:
        def __init__(self, *a, **k):
            pass

    def seed_everything(seed=None, workers=False):
        import random

        import numpy as np

        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        return seed

    pl.LightningModule = LightningModule
    pl.LightningDataModule = LightningDataModule
    pl.seed_everything = seed_everything
    pl.Trainer = object
    sys.modules["pytorch_lightning"] = pl
    for sub in ("callbacks", "loggers", "utilities", "strategies"):
        m = types.ModuleType(f"pytorch_lightning.{sub}")
        sys.modules[f"pytorch_lightning.{sub}"] = m
        setattr(pl, sub, m)

    # --- omegaconf ----------------------------------------------------------------------------
    oc = types.ModuleType("omegaconf")

    class ListConfig(list):
        pass

    class DictConfig(dict):
        pass

    class OmegaConf:
        @staticmethod
        def to_container(x, resolve=True):
            return x

    oc.ListConfig, oc.DictConfig, oc.OmegaConf = ListConfig, DictConfig, OmegaConf
    lc = types.ModuleType("omegaconf.listconfig")
    lc.ListConfig = ListConfig
    dc = types.ModuleType("omegaconf.dictconfig")
    dc.DictConfig = DictConfig
    sys.modules["omegaconf"] = oc
    sys.modules["omegaconf.listconfig"] = lc
    sys.modules["omegaconf.dictconfig"] = dc

    # --- rdkit ----------------------------------------------------------------------------------
    rd = types.ModuleType("rdkit")
    chem = types.ModuleType("rdkit.Chem")
    chem.Mol = object
    chem.MolFromSmiles = lambda s: object()
    chem.MolToSmiles = lambda m, **k: ""
    chem.MolFromSmarts = lambda s: object()
    rdm = types.ModuleType("rdkit.Chem.rdMolDescriptors")
    rdm.CalcMolFormula = lambda m: ""
    rdl = types.ModuleType("rdkit.RDLogger")
    rdl.DisableLog = lambda *a, **k: None
    chem.rdMolDescriptors = rdm
    rd.Chem = chem
    rd.RDLogger = rdl
    sys.modules["rdkit"] = rd
    sys.modules["rdkit.Chem"] = chem
    sys.modules["rdkit.Chem.rdMolDescriptors"] = rdm
    sys.modules["rdkit.RDLogger"] = rdl

    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)

    # torch's eval-mode MHA/encoder fast path drops the GLU gate (SURVEY.md §3.3 hazard)
    torch.backends.mha.set_fastpath_enabled(False)

    # offline CustomConfig.from_pretrained: facebook/bart-base is only used for
    # is_encoder_decoder=True, dropout=0.1, activation_function="gelu" 
    from analytical_fm.modeling import custom_modeling

    def _from_pretrained(cls, name, **kw):
        kw.setdefault("is_encoder_decoder", True)
        return cls(**kw)

    custom_modeling.CustomConfig.from_pretrained = classmethod(_from_pretrained)
