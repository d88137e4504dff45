"""CPU tests of the host-side logic: parameter store / checkpoint layout, schedules, config errors, metrics."""
import math

import pytest
import torch

from multimodalanalytical_b200.params import ModelConfig, ParamStore, sincos_table
from multimodalanalytical_b200.trainer import one_cycle
from multimodalanalytical_b200.wrapper import top_n_string_accuracy
from tests.helpers import load_case


def cfg_from(fx):
    mk = fx["model_kwargs"]
    tgt = [m for m, c in fx["data_config"].items() if c["target"] and not c.get("alignment")][0]
    return ModelConfig(data_config=fx["data_config"], vocab_size=fx["data_config"][tgt]["vocab_size"],
                       d_model=mk["d_model"], encoder_layers=mk["encoder_layers"], decoder_layers=mk["decoder_layers"],
                       encoder_attention_heads=mk["encoder_attention_heads"],
                       decoder_attention_heads=mk["decoder_attention_heads"], encoder_ffn_dim=mk["encoder_ffn_dim"],
                       decoder_ffn_dim=mk["decoder_ffn_dim"], gated_linear=mk["gated_linear"],
                       positional_encoding_type=mk["positional_encoding_type"], multimodal_norm=mk["multimodal_norm"],
                       max_position_embeddings=mk["max_position_embeddings"], align_config=mk.get("align_config"))


@pytest.mark.parametrize("name", ["c1_ir_tiny", "mm_gated_learned", "align_conv", "align_modality", "post_ln"])
def test_param_store_has_reference_checkpoint_layout(name):
    fx = load_case(name)
    ps = ParamStore(cfg_from(fx), device="cpu")
    sd = ps.state_dict()
    assert set(sd) == set(fx["state_dict"]), set(sd) ^ set(fx["state_dict"])
    for k, v in fx["state_dict"].items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    ps.load_state_dict(fx["state_dict"])
    for k, v in fx["state_dict"].items():
        assert torch.equal(ps.state_dict()[k], v), k
    # shared embedding: the three aliases are the same storage
    emb = [k for k in sd if k.startswith("hf_model.embedding.")][0]
    tail = emb[len("hf_model.embedding."):]
    assert sd["hf_model.decoder.embedding." + tail].data_ptr() == sd[emb].data_ptr()
    assert sd["multimodal_embedding." + tail].data_ptr() == sd[emb].data_ptr()


def test_post_layer_normalisation_reaches_the_engine():
    """The model yaml's `post_layer_normalisation` (custom_modeling.py:55,119-129) selects the layer order of the engine
    (True, the default of every shipped config: x + f(LN(x)); False: LN(x + f(x))) - it is no longer refused."""
    from multimodalanalytical_b200.wrapper import load_custom_model
    fx = load_case("post_ln")
    tok = type("T", (), dict(vocab_size=31, pad_token_id=0, bos_token_id=2, eos_token_id=3))()
    for flag in (False, True):
        mk = dict(fx["model_kwargs"], post_layer_normalisation=flag)
        for k in ("multimodal_norm", "model_name", "model_type"):
            mk.pop(k, None)
        m, _ = load_custom_model("facebook/bart-base", tok, "Smiles", fx["data_config"], True, device="cpu", **mk)
        assert m.config.post_layer_normalisation is flag and m.engine.norm_first is flag


def test_flat_layout_is_aligned_and_forward_ordered():
    fx = load_case("c1_ir_tiny")
    ps = ParamStore(cfg_from(fx), device="cpu")
    offs = [ps.offsets[n][0] for n, _, _ in ps.specs]
    assert offs == sorted(offs) and all(o % 128 == 0 for o in offs)
    assert ps.offsets["hf_model.token_ff.weight"][0] > ps.offsets["hf_model.decoder.layers.0.self_attn.in_proj_weight"][0]
    with pytest.raises(RuntimeError):
        ps.load_state_dict({"hf_model.token_ff.weight": torch.zeros(3, 3)}, strict=False)


def test_init_rule_xavier_on_matrices():
    fx = load_case("c1_ir_tiny")
    ps = ParamStore(cfg_from(fx), device="cpu", seed=1)
    w = ps.P("hf_model.encoder.layers.0.self_attn.in_proj_weight")  # packed [3d, d]: fan computed on packed shape
    a = math.sqrt(6.0 / (w.shape[0] + w.shape[1]))
    assert w.abs().max() <= a and w.abs().max() > 0.9 * a
    assert torch.all(ps.P("hf_model.encoder.layers.0.norm1.weight") == 1)
    assert torch.all(ps.P("hf_model.encoder.layers.0.self_attn.in_proj_bias") == 0)


def test_sincos_table_bit_identical_to_reference_buffer():
    fx = load_case("c1_ir_tiny")
    ref = fx["state_dict"]["hf_model.embedding.positional_encodings.pos_enc"]
    assert torch.equal(sincos_table(ref.shape[1], ref.shape[0]), ref)


def test_config_errors_match_reference_conventions():
    dc = {"A": {"type": "text", "target": True, "vocab_size": 5, "pad_token_id": 0},
          "B": {"type": "text", "target": True, "vocab_size": 5, "pad_token_id": 0}}
    with pytest.raises(ValueError):
        ModelConfig(data_config=dc, vocab_size=5)
    dc = {"A": {"type": "hologram", "target": False}, "B": {"type": "text", "target": True, "vocab_size": 5, "pad_token_id": 0}}
    with pytest.raises(NotImplementedError):
        ModelConfig(data_config=dc, vocab_size=5)


@pytest.mark.parametrize("total", [10, 100, 1234])
def test_one_cycle_matches_torch(total):
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=1e-3)
    sch = torch.optim.lr_scheduler.OneCycleLR(opt, 1e-3, total_steps=total)
    for step in range(total):
        lr, b1 = one_cycle(step, total, 1e-3)
        assert abs(lr - opt.param_groups[0]["lr"]) < 1e-12 + 1e-9 * lr, step
        assert abs(b1 - opt.param_groups[0]["betas"][0]) < 1e-9, step
        opt.step()
        if step < total - 1:
            sch.step()


def test_top_n_bookkeeping():
    samples = [["C C O", "CCN"], ["c1ccccc1", "CC"], ["N", "O"]]
    targets = ["CCO", "CC", "S"]
    m = top_n_string_accuracy(samples, targets)
    assert m["Top-1"] == pytest.approx(1 / 3) and m["Top-2"] == pytest.approx(2 / 3)


def test_product_code_never_imports_the_oracle_or_the_reference():
    """The oracle is test infrastructure: only tests/, __graft_entry__.smoke()/build() and bench.py's CPU-baseline legs
    may import it; the package must not, and nothing shipped may reach into /root/reference at run time."""
    import ast
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "multimodalanalytical_b200")
    for fn in sorted(os.listdir(pkg)):
        if not fn.endswith(".py"):
            continue
        src = open(os.path.join(pkg, fn)).read()
        for node in ast.walk(ast.parse(src)):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                mods = [node.module or ""]
            for m in mods:
                assert not (m == "oracle" or m.startswith("oracle.")), (fn, m)
                assert not m.startswith("analytical_fm"), (fn, m)
        assert "/root/reference" not in src, fn
    for fn in ("bench.py", "__graft_entry__.py"):
        assert "/root/reference" not in open(os.path.join(root, fn)).read(), fn


def test_calculate_training_steps_follows_the_reference_formula():
    """analytical_fm/utils.py:155-172: ceil(ceil(n / batch) / acc) * epochs with a hard-coded GPU count of 1."""
    from multimodalanalytical_b200.trainer import calculate_training_steps

    assert calculate_training_steps(15, 128, 4, 1) == 1           # the tiny test run (tests/test_run.py)
    assert calculate_training_steps(1000, 256, 4, 60) == math.ceil(math.ceil(1000 / 256) / 4) * 60
    assert calculate_training_steps(100_000, 256, 4, 10, world=8) == calculate_training_steps(100_000, 256, 4, 10)
    assert calculate_training_steps(100_000, 256, 4, 10, world=8, reference_compat=False) == math.ceil(49 / 4) * 10
