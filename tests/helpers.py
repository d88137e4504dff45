"""Shared test helpers: golden-fixture loading and oracle-config construction."""
import os

import torch

from oracle import spectra_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("c1_ir_tiny", "mm_gated_learned", "align_conv", "align_modality", "post_ln")


def load_case(name):
    return torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), weights_only=False)


def oracle_cfg(fx):
    mk = fx["model_kwargs"]
    target = [m for m, c in fx["data_config"].items() if c["target"] and not c.get("alignment")][0]
    return orc.OracleConfig(
        d_model=mk["d_model"],
        encoder_layers=mk["encoder_layers"],
        decoder_layers=mk["decoder_layers"],
        encoder_attention_heads=mk["encoder_attention_heads"],
        decoder_attention_heads=mk["decoder_attention_heads"],
        gated_linear=mk["gated_linear"],
        post_layer_normalisation=mk.get("post_layer_normalisation", True),
        positional_encoding_type=mk["positional_encoding_type"],
        multimodal_norm=mk["multimodal_norm"],
        align_config=mk.get("align_config"),
        target_modality=target,
        data_config=fx["data_config"],
    )


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
