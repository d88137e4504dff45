"""Golden vectors for a data_config that CONTAINS the alignment modality (configs/data/ir/patches_mixture_text_align.yaml:
`IR_target: {target: True, alignment: True}`), produced by the UNMODIFIED reference:

    python tests/golden/make_align_modality_golden.py     # rewrites tests/golden/align_modality.pt

The reference's MultimodalEmbedding builds an (unused) embedding layer + LayerNorm for EVERY data_config entry
(modeling/utils.py:73-77), so its checkpoints hold `...embedding_layer_dict.IR_target.*` / `...embedding_norm_dict.IR_target.*`
under the three embedding prefixes.  The fixture pins that key layout (strict load both ways) next to logits / losses.
"""
import os

import torch

import make_golden as mg  # installs the reference stubs on import

HERE = os.path.dirname(os.path.abspath(__file__))


def case_align_modality():
    g = torch.Generator().manual_seed(13)
    B = 4
    patch = {"type": "1D_patches", "preprocessor_arguments": {"patch_size": 20, "interpolation": False, "masking": False}}
    data_config = {
        "Formula": {"type": "text", "target": False, "vocab_size": 30, "pad_token_id": 0, "preprocessor_arguments": {}},
        "IR": dict(patch, target=False),
        "IR_target": dict(patch, target=True, alignment=True),
        "Smiles": {"type": "text", "target": True, "vocab_size": 28, "pad_token_id": 0, "preprocessor_arguments": {}},
    }
    f_ids, f_pad = mg._tok(B, 8, 30, g, 3)
    ir = torch.randn(6, B, 20, generator=g)
    t_ids, t_pad = mg._tok(B, 13, 28, g, 5)
    batch = {
        "encoder_input": {"Formula": f_ids, "IR": ir},
        "encoder_pad_mask": torch.cat([f_pad, torch.zeros(6, B, dtype=torch.bool)], dim=0),
        "decoder_input": {"Smiles": t_ids[:-1]},
        "decoder_pad_mask": t_pad[:-1],
        "target": t_ids[1:],
        "target_mask": t_pad[1:],
        "encoder_alignment_input": torch.rand(B, 90, generator=g),
    }
    ac = {"align_network": "mlp", "hidden_dimension": 24, "conv_channels": 40, "kernel_size": 5,
          "output_dimension": 90, "loss_lambda": 50, "loss_function": "mse"}
    mk = mg.model_kwargs(d_model=32, num_heads=2, encoder_attention_heads=2, decoder_attention_heads=2,
                         encoder_layers=1, decoder_layers=1, encoder_ffn_dim=64, decoder_ffn_dim=64, align_config=ac)
    sd, res = mg.run_reference(data_config, mg.FakeTokenizer(28), mk, batch, beams=(1, 3), seed=37)
    return {"model_kwargs": mk, "data_config": data_config, "state_dict": sd, "batch": batch, "ref": res}


if __name__ == "__main__":
    fx = case_align_modality()
    path = os.path.join(HERE, "align_modality.pt")
    torch.save(fx, path)
    print("align_modality loss", float(fx["ref"]["loss"]), "keys", len(fx["state_dict"]),
          [k for k in fx["state_dict"] if "IR_target" in k], f"{os.path.getsize(path) / 1e6:.2f} MB")
