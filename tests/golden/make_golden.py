"""Generate golden vectors by running the UNMODIFIED reference (imported from /root/reference/src
through `_ref_stubs`) in the build container.  The reference cannot travel to the GPU box, so the
vectors are committed as small fixtures next to this script:

    python tests/golden/make_golden.py        # rewrites tests/golden/*.pt

Cases
  c1_ir_tiny        real bundled parquet through the reference's own dataset/collator pipeline
                    (configs/data/ir/patches.yaml), custom_model.yaml shrunk to d=64/2+2 layers
  mm_gated_learned  synthetic multimodal batch (tokens, XVal tokens, patches via 2-layer MLP,
                    msms_number peaks), learned pos-enc, GLU FFN, padding inside the sequence
  align_conv        custom_model_align.yaml head (convolutional, MAE) shrunk
Each fixture holds: model kwargs, data_config, state_dict, batch, and the reference's outputs
(logits, loss, selected grads, greedy ids, beam ids).
"""
import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_stubs  # noqa: E402

_ref_stubs.install()

import pytorch_lightning as pl  # noqa: E402  (stub)
from analytical_fm.data import data_utils, datamodules, datasets  # noqa: E402
from analytical_fm.modeling import wrapper  # noqa: E402

REF = "/root/reference"
GRAD_KEYS_SUFFIX = (
    "token_ff.weight",
    "encoder.layers.0.self_attn.in_proj_weight",
    "encoder.layers.0.linear1.weight",
    "decoder.layers.1.multihead_attn.out_proj.weight",
    "decoder.layers.0.norm3.weight",
    "decoder.norm.bias",
)


class FakeTokenizer:
    """Just the attributes HFWrapper / CustomModel read from the target tokenizer."""

    def __init__(self, vocab_size):
        self.vocab_size = vocab_size
        self.pad_token_id, self.bos_token_id, self.eos_token_id = 0, 2, 3


def model_kwargs(**over):
    cfg = yaml.safe_load(open(f"{REF}/configs/model/custom_model.yaml"))
    for k in ("lr", "weight_decay", "adam_beta1", "adam_beta2"):
        cfg[k] = float(cfg[k])
    cfg.update(over)
    return cfg


def run_reference(data_config, tokenizer, mk, batch, beams, seed, eos_bias=2.0):
    torch.manual_seed(seed)
    model = wrapper.HFWrapper(data_config=data_config, target_tokenizer=tokenizer, num_steps=100, **mk)
    # make the generation less degenerate than pure xavier noise: sharpen the LM head a little
    with torch.no_grad():
        model.hf_model.token_ff.weight.mul_(6.0)
        model.hf_model.token_ff.bias[tokenizer.eos_token_id] += eos_bias  # let hypotheses finish early
        for n, p in model.named_parameters():
            if p.dim() == 1 and "norm" in n:
                p.add_(0.1 * torch.randn_like(p))
            elif p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    model.eval()  # dropout off; fast path is disabled in _ref_stubs
    out = model.forward(batch)
    out.loss.backward()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    grads = {}
    for n, p in model.named_parameters():
        if n.endswith(GRAD_KEYS_SUFFIX) or "embedding_layer_dict" in n or "embedding_norm_dict" in n \
                or "positional_encodings" in n or "align_network" in n:
            if p.grad is not None and n.startswith("hf_model.") and ".decoder.embedding." not in n:
                grads[n] = p.grad.detach().clone()
    res = {
        "logits": out.logits.detach().clone(),
        "loss": out.loss.detach().clone(),
        "model_only_loss": out.loss_dict["model_only_loss"].detach().clone(),
        "grads": grads,
    }
    if out.loss_dict.get("alignment_loss") is not None:
        res["alignment_loss"] = out.loss_dict["alignment_loss"].detach().clone()
    with torch.no_grad():
        for k in beams:
            res[f"gen_beam{k}"] = model.generate(batch, n_beams=k).clone()
    return sd, res


def case_c1():
    pl.seed_everything(3247)
    dc = yaml.safe_load(open(f"{REF}/configs/data/ir/patches.yaml"))
    data_config, ds = datasets.build_dataset_multimodal(
        dc, data_path=f"{REF}/tests/test_data/ir_dataset", splitting="random", cv_split=0,
        augment_config=None, num_cpu=1, mixture_config=None)
    np.random.seed(3247)
    data_config, pre = data_utils.load_preprocessors(ds["train"], data_config)
    dm = datamodules.MultiModalDataModule(dataset=ds, preprocessors=pre, data_config=data_config,
                                          model_type="CustomModel", batch_size=16, num_workers=0,
                                          extra_columns=[None])
    rows = [ds["train"][i] for i in range(len(ds["train"]))]
    batch = dm.collator(rows)
    batch = {k: v for k, v in batch.items() if v is not None}
    mk = model_kwargs(d_model=64, num_heads=4, encoder_attention_heads=4, decoder_attention_heads=4,
                      encoder_layers=2, decoder_layers=2, encoder_ffn_dim=128, decoder_ffn_dim=128)
    tok = pre["Smiles"]
    sd, res = run_reference(data_config, tok, mk, batch, beams=(1, 3, 10), seed=11)
    vocab = {t: i for t, i in tok.get_vocab().items()}
    return {"model_kwargs": mk, "data_config": data_config, "state_dict": sd, "batch": batch,
            "ref": res, "smiles_vocab": vocab,
            "patch_mean_std": (float(pre["IR"].mean), float(pre["IR"].std)) if hasattr(pre["IR"], "mean") else None}


def _tok(B, S, vocab, g, min_len):
    ids = torch.zeros(S, B, dtype=torch.long)
    pad = torch.ones(S, B, dtype=torch.bool)
    for b in range(B):
        n = int(torch.randint(min_len, S + 1, (1,), generator=g))
        ids[:n, b] = torch.randint(4, vocab, (n,), generator=g)
        ids[0, b], ids[n - 1, b] = 2, 3
        pad[:n, b] = False
    return ids, pad


def case_mm():
    g = torch.Generator().manual_seed(5)
    B = 6
    data_config = {
        "Formula": {"type": "text", "target": False, "vocab_size": 40, "pad_token_id": 0,
                    "preprocessor_arguments": {}},
        "Multiplets": {"type": "multiplets", "target": False, "vocab_size": 50, "pad_token_id": 0,
                       "preprocessor_arguments": {}},
        "Carbon": {"type": "carbon", "target": False, "vocab_size": 60, "pad_token_id": 0,
                   "preprocessor_arguments": {}},
        "IR": {"type": "1D_patches", "target": False,
               "preprocessor_arguments": {"patch_size": 30, "encoding_type": "linear_2_layer"}},
        "MSMS": {"type": "msms_number", "target": False, "preprocessor_arguments": {}},
        "Percentage": {"type": "1D_patches", "target": False,
                       "preprocessor_arguments": {"patch_size": 1, "encoding_type": "linear_3_layer"}},
        "Smiles": {"type": "text", "target": True, "vocab_size": 37, "pad_token_id": 0,
                   "preprocessor_arguments": {}},
    }
    f_ids, f_pad = _tok(B, 9, 40, g, 3)
    m_ids, m_pad = _tok(B, 17, 50, g, 4)
    m_val = torch.where(m_pad, torch.ones(17, B), 1.0 + 0.5 * torch.randn(17, B, generator=g))
    c_ids, c_pad = _tok(B, 11, 60, g, 2)
    ir = torch.randn(7, B, 30, generator=g)
    ir_pad = torch.zeros(7, B, dtype=torch.bool)
    ir_pad[:, 4] = True  # a sample without spectrum: whole modality masked (patches.py:98-105)
    ms = torch.randn(5, B, 2, generator=g)
    ms_pad = torch.zeros(5, B, dtype=torch.bool)
    ms_pad[3:, 1] = True
    ms_pad[2:, 3] = True
    pc = torch.rand(1, B, 1, generator=g)
    pc_pad = torch.zeros(1, B, dtype=torch.bool)
    t_ids, t_pad = _tok(B, 21, 37, g, 6)
    batch = {
        "encoder_input": {"Formula": f_ids, "Multiplets": m_ids,
                          "Carbon": c_ids, "IR": ir, "MSMS": ms, "Percentage": pc},
        "encoder_pad_mask": torch.cat([f_pad, m_pad, c_pad, ir_pad, ms_pad, pc_pad], dim=0),
        "decoder_input": {"Smiles": t_ids[:-1]},
        "decoder_pad_mask": t_pad[:-1],
        "target": t_ids[1:],
        "target_mask": t_pad[1:],
    }
    mk = model_kwargs(d_model=48, num_heads=3, encoder_attention_heads=3, decoder_attention_heads=3,
                      encoder_layers=2, decoder_layers=3, encoder_ffn_dim=80, decoder_ffn_dim=112,
                      positional_encoding_type="learned", gated_linear=True, n_beams=4)
    sd, res = run_reference(data_config, FakeTokenizer(37), mk, batch, beams=(1, 4), seed=23)
    # XVal (utils.py:154-160) is reachable through MultimodalEmbedding only: HFWrapper.forward calls
    # .transpose on every encoder input (wrapper.py:356-359) and fails on the XVal dict.
    from analytical_fm.modeling.utils import MultimodalEmbedding
    torch.manual_seed(4)
    emb = MultimodalEmbedding({k: data_config[k] for k in ("Formula", "Multiplets")}, 48, True,
                              do_positional_encodings=True, positional_encodings_type="sin_cos")
    for p in emb.parameters():
        if p.dim() > 1:
            torch.nn.init.xavier_uniform_(p)
    xin = {"Formula": f_ids.T, "Multiplets": {"tokenized_input": m_ids.T, "numerical_values": m_val.T}}
    xval = {"state_dict": {"hf_model.embedding." + k: v.detach().clone() for k, v in emb.state_dict().items()},
            "inputs": xin, "out": emb(xin).detach().clone()}
    return {"model_kwargs": mk, "data_config": data_config, "state_dict": sd, "batch": batch, "ref": res,
            "xval": xval}


def case_align():
    g = torch.Generator().manual_seed(9)
    B = 5
    data_config = {
        "Formula": {"type": "text", "target": False, "vocab_size": 30, "pad_token_id": 0,
                    "preprocessor_arguments": {}},
        "IR": {"type": "1D_patches", "target": False, "preprocessor_arguments": {"patch_size": 20}},
        "Smiles": {"type": "text", "target": True, "vocab_size": 28, "pad_token_id": 0,
                   "preprocessor_arguments": {}},
    }
    f_ids, f_pad = _tok(B, 8, 30, g, 3)
    ir = torch.randn(6, B, 20, generator=g)
    t_ids, t_pad = _tok(B, 15, 28, g, 5)
    batch = {
        "encoder_input": {"Formula": f_ids, "IR": ir},
        "encoder_pad_mask": torch.cat([f_pad, torch.zeros(6, B, dtype=torch.bool)], dim=0),
        "decoder_input": {"Smiles": t_ids[:-1]},
        "decoder_pad_mask": t_pad[:-1],
        "target": t_ids[1:],
        "target_mask": t_pad[1:],
        "encoder_alignment_input": torch.rand(B, 90, generator=g),
    }
    ac = {"align_network": "convolutional", "hidden_dimension": 24, "conv_channels": 40, "kernel_size": 5,
          "output_dimension": 90, "loss_lambda": 50, "loss_function": "mae"}
    mk = model_kwargs(d_model=32, num_heads=2, encoder_attention_heads=2, decoder_attention_heads=2,
                      encoder_layers=1, decoder_layers=1, encoder_ffn_dim=64, decoder_ffn_dim=64,
                      align_config=ac)
    sd, res = run_reference(data_config, FakeTokenizer(28), mk, batch, beams=(1, 3), seed=31)
    return {"model_kwargs": mk, "data_config": data_config, "state_dict": sd, "batch": batch, "ref": res}


def case_postln():
    """`post_layer_normalisation: False` -> torch's norm_first=False layers, LN(x + f(x)) (custom_modeling.py:119-129,
    166-176), with the gated FFN; no shipped yaml sets it, the model surface has it."""
    g = torch.Generator().manual_seed(13)
    B = 5
    data_config = {
        "Formula": {"type": "text", "target": False, "vocab_size": 30, "pad_token_id": 0,
                    "preprocessor_arguments": {}},
        "IR": {"type": "1D_patches", "target": False, "preprocessor_arguments": {"patch_size": 20}},
        "Smiles": {"type": "text", "target": True, "vocab_size": 31, "pad_token_id": 0,
                   "preprocessor_arguments": {}},
    }
    f_ids, f_pad = _tok(B, 8, 30, g, 3)
    ir = torch.randn(6, B, 20, generator=g)
    t_ids, t_pad = _tok(B, 17, 31, g, 5)
    batch = {
        "encoder_input": {"Formula": f_ids, "IR": ir},
        "encoder_pad_mask": torch.cat([f_pad, torch.zeros(6, B, dtype=torch.bool)], dim=0),
        "decoder_input": {"Smiles": t_ids[:-1]},
        "decoder_pad_mask": t_pad[:-1],
        "target": t_ids[1:],
        "target_mask": t_pad[1:],
    }
    mk = model_kwargs(d_model=64, num_heads=2, encoder_attention_heads=2, decoder_attention_heads=2,
                      encoder_layers=2, decoder_layers=2, encoder_ffn_dim=96, decoder_ffn_dim=128,
                      gated_linear=True, post_layer_normalisation=False, n_beams=3)
    sd, res = run_reference(data_config, FakeTokenizer(31), mk, batch, beams=(1, 3), seed=37)
    return {"model_kwargs": mk, "data_config": data_config, "state_dict": sd, "batch": batch, "ref": res}


CASES = (("c1_ir_tiny", case_c1), ("mm_gated_learned", case_mm), ("align_conv", case_align), ("post_ln", case_postln))


def main():
    # `python make_golden.py post_ln` rewrites only the named fixtures
    only = set(sys.argv[1:])
    for name, fn in CASES:
        if only and name not in only:
            continue
        fx = fn()
        path = os.path.join(HERE, f"{name}.pt")
        torch.save(fx, path)
        r = fx["ref"]
        for k, v in r.items():
            if k.startswith("gen_"):
                print("  ", k, "lengths", sorted(set((v != 0).sum(1).tolist())))
        print(name, "loss", float(r["loss"]), "logits", tuple(r["logits"].shape),
              {k: tuple(v.shape) for k, v in r.items() if k.startswith("gen_")},
              f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
