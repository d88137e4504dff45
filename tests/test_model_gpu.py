"""End-to-end parity of the CUDA path (through HFWrapper / the C ABI) against the CPU oracle and the golden
vectors produced by the reference.  Tolerances are the north_star's: logits/loss 1e-5 relative in fp32,
1e-2 in bf16; greedy and beam token sequences identical in fp32."""
import copy

import pytest
import torch

from oracle import spectra_oracle as orc
from tests.helpers import load_case, oracle_cfg, rel_err

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from multimodalanalytical_b200.wrapper import HFWrapper


class FakeTokenizer:
    def __init__(self, vocab_size):
        self.vocab_size = vocab_size
        self.pad_token_id, self.bos_token_id, self.eos_token_id = 0, 2, 3

    def batch_decode(self, seqs, skip_special_tokens=True):
        out = []
        for s in seqs.tolist():
            out.append(" ".join(str(t) for t in s if not (skip_special_tokens and t in (0, 2, 3))))
        return out


def build(fx, precision, **over):
    mk = dict(fx["model_kwargs"])
    mk.update(over)
    tgt = [m for m, c in fx["data_config"].items() if c["target"] and not c.get("alignment")][0]
    tok = FakeTokenizer(fx["data_config"][tgt]["vocab_size"])
    m = HFWrapper(data_config=fx["data_config"], target_tokenizer=tok, num_steps=100, precision=precision, **mk)
    m.load_state_dict(fx["state_dict"])
    return m


def oracle_grads(fx):
    cfg = oracle_cfg(fx)
    sd = {k: v.clone() for k, v in fx["state_dict"].items()}
    for k in list(sd):
        if k.startswith("hf_model.decoder.embedding."):
            sd[k] = sd["hf_model.embedding." + k[len("hf_model.decoder.embedding."):]]
    leaves = {}
    for k, v in sd.items():
        if v.is_floating_point() and k.startswith("hf_model.") and ".decoder.embedding." not in k and "pos_enc" != k.split(".")[-1]:
            v.requires_grad_(True)
            leaves[k] = v
    out = orc.wrapper_forward(sd, cfg, fx["batch"])
    out["loss"].backward()
    return out, {k: v.grad for k, v in leaves.items() if v.grad is not None}


CASES = ("c1_ir_tiny", "mm_gated_learned", "align_conv", "align_modality", "post_ln")


@pytest.mark.parametrize("name", CASES)
def test_fp32_logits_loss_match_reference_golden(name):
    fx = load_case(name)
    m = build(fx, "fp32")
    m.eval()
    with torch.no_grad():
        out = m.forward(fx["batch"])
    ref = fx["ref"]
    assert rel_err(out.logits.cpu(), ref["logits"]) < 1e-5
    assert abs(float(out.loss) - float(ref["loss"])) < 1e-5 * abs(float(ref["loss"]))


@pytest.mark.parametrize("name", CASES)
def test_bf16_logits_loss_within_1e2(name):
    fx = load_case(name)
    m = build(fx, "bf16")
    m.eval()
    with torch.no_grad():
        out = m.forward(fx["batch"])
    ref = fx["ref"]
    assert rel_err(out.logits.cpu(), ref["logits"]) < 1e-2
    assert abs(float(out.loss) - float(ref["loss"])) < 1e-2 * abs(float(ref["loss"]))


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_gradients_match_oracle(name, precision):
    fx = load_case(name)
    m = build(fx, precision, dropout=0.0)
    m.train()
    m.store.g.zero_()
    out = m.forward(fx["batch"])
    out.loss.backward()
    torch.cuda.synchronize()
    _, want = oracle_grads(fx)
    tol = 2e-4 if precision == "fp32" else 6e-2
    worst = []
    for k, g in want.items():
        got = m.store.G(k).cpu()
        e = rel_err(got, g)
        worst.append((e, k))
    worst.sort(reverse=True)
    assert worst[0][0] < tol, worst[:5]
    # reference's own gradients (golden) for the sampled tensors
    for k, g in fx["ref"]["grads"].items():
        assert rel_err(m.store.G(k).cpu(), g) < tol, k


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("use_graph", [False, True])
def test_fp32_generation_identical_to_reference(name, use_graph):
    fx = load_case(name)
    m = build(fx, "fp32")
    m.eval()
    for key, want in fx["ref"].items():
        if not key.startswith("gen_beam"):
            continue
        k = int(key[len("gen_beam"):])
        got = m.generate(fx["batch"], n_beams=k, use_graph=use_graph, check_every=5).cpu()
        assert got.shape == want.shape, (key, got.shape, want.shape)
        if not torch.equal(got, want):
            diff = (got != want).nonzero()
            raise AssertionError(f"{key}: first divergence at {diff[0].tolist()} ({len(diff)} tokens differ)")


def test_bf16_generation_runs_and_mostly_agrees():
    fx = load_case("mm_gated_learned")
    m = build(fx, "bf16")
    m.eval()
    got = m.generate(fx["batch"], n_beams=4).cpu()
    assert got.shape[0] == fx["ref"]["gen_beam4"].shape[0]
    assert (got[:, 0] == 2).all()


def test_state_dict_roundtrip_reference_layout():
    fx = load_case("c1_ir_tiny")
    m = build(fx, "fp32")
    sd = m.state_dict()
    assert set(sd) == set(fx["state_dict"])
    for k, v in fx["state_dict"].items():
        assert torch.equal(sd[k].cpu(), v), k


def test_midsize_random_model_against_oracle():
    """d=512 / 8 heads / ffn 2048 (custom_model.yaml dims, 2+2 layers), IR patches + formula, fp32 and bf16."""
    torch.manual_seed(0)
    data_config = {
        "Formula": {"type": "text", "target": False, "vocab_size": 64, "pad_token_id": 0, "preprocessor_arguments": {}},
        "IR": {"type": "1D_patches", "target": False, "preprocessor_arguments": {"patch_size": 75}},
        "Smiles": {"type": "text", "target": True, "vocab_size": 200, "pad_token_id": 0, "preprocessor_arguments": {}},
    }
    B, S1, P, T = 8, 15, 21, 64
    f_ids = torch.randint(4, 64, (S1, B))
    f_pad = torch.zeros(S1, B, dtype=torch.bool)
    f_pad[10:, ::2] = True
    f_ids[f_pad] = 0
    ir = torch.randn(P, B, 75)
    t_ids = torch.randint(4, 200, (T + 1, B))
    t_pad = torch.zeros(T + 1, B, dtype=torch.bool)
    t_pad[40:, 1::3] = True
    t_ids[t_pad] = 0
    batch = {"encoder_input": {"Formula": f_ids, "IR": ir},
             "encoder_pad_mask": torch.cat([f_pad, torch.zeros(P, B, dtype=torch.bool)], 0),
             "decoder_input": {"Smiles": t_ids[:-1]}, "decoder_pad_mask": t_pad[:-1], "target": t_ids[1:]}
    mk = dict(model_type="CustomModel", model_name="facebook/bart-base", d_model=512, num_heads=8,
              encoder_attention_heads=8, decoder_attention_heads=8, encoder_layers=2, decoder_layers=2,
              encoder_ffn_dim=2048, decoder_ffn_dim=2048, multimodal_norm=True, positional_encoding_type="sin_cos",
              gated_linear=False, max_position_embeddings=1024, dropout=0.0)
    fx = {"model_kwargs": mk, "data_config": data_config, "batch": batch}
    cfg = oracle_cfg(fx)
    sd = orc.init_state_dict(cfg, vocab=200, enc_ffn=2048, dec_ffn=2048, seed=1)
    fx["state_dict"] = sd
    want_out, want_g = oracle_grads(fx)
    for precision, tol_l, tol_g in (("fp32", 1e-5, 3e-4), ("bf16", 1e-2, 8e-2)):
        m = build(fx, precision)
        m.train()
        m.store.g.zero_()
        out = m.forward(batch)
        out.loss.backward()
        torch.cuda.synchronize()
        assert rel_err(out.logits.cpu(), want_out["logits"].detach()) < tol_l, precision
        assert abs(float(out.loss) - float(want_out["loss"])) < tol_l * float(want_out["loss"]), precision
        worst = sorted(((rel_err(m.store.G(k).cpu(), g), k) for k, g in want_g.items()), reverse=True)
        assert worst[0][0] < tol_g, (precision, worst[:5])


def test_align_head_losses_match_reference_and_oracle():
    """custom_model_align: total = lm + lambda * align (golden from the reference: convolutional head + MAE); the MLP
    head and the MSE / SID losses against the oracle (loss and every gradient)."""
    fx = load_case("align_conv")
    m = build(fx, "fp32")
    m.eval()
    with torch.no_grad():
        out = m.forward(fx["batch"])
    ref = fx["ref"]
    assert abs(float(out.loss_dict["alignment_loss"]) - float(ref["alignment_loss"])) < 1e-5 * abs(float(ref["alignment_loss"]))
    assert abs(float(out.loss_dict["model_only_loss"]) - float(ref["model_only_loss"])) < 1e-5 * abs(float(ref["model_only_loss"]))
    assert abs(float(out.loss) - float(ref["loss"])) < 1e-5 * abs(float(ref["loss"]))
    for network, loss_fn in (("convolutional", "mse"), ("convolutional", "sid"), ("mlp", "mae"), ("mlp", "sid")):
        fy = copy.deepcopy(fx)
        ac = dict(fy["model_kwargs"]["align_config"], align_network=network, loss_function=loss_fn, loss_lambda=3.0)
        fy["model_kwargs"]["align_config"] = ac
        if loss_fn == "sid":  # a spectrum-like positive target
            fy["batch"]["encoder_alignment_input"] = fy["batch"]["encoder_alignment_input"].abs() + 0.05
        if network == "mlp":
            sd = {k: v for k, v in fy["state_dict"].items() if ".align_network." not in k or ".0." in k}
            g = torch.Generator().manual_seed(5)
            sd["hf_model.align_network.2.weight"] = torch.randn(ac["output_dimension"], ac["hidden_dimension"], generator=g) * 0.2
            sd["hf_model.align_network.2.bias"] = torch.randn(ac["output_dimension"], generator=g) * 0.1
            fy["state_dict"] = sd
        want_out, want_g = oracle_grads(fy)
        mm = build(fy, "fp32", dropout=0.0)
        mm.train()
        mm.store.g.zero_()
        o = mm.forward(fy["batch"])
        o.loss.backward()
        torch.cuda.synchronize()
        assert abs(float(o.loss) - float(want_out["loss"])) < 2e-5 * abs(float(want_out["loss"])), (network, loss_fn)
        assert abs(float(o.loss_dict["alignment_loss"]) - float(want_out["alignment_loss"])) < 2e-5 * abs(float(want_out["alignment_loss"]))
        worst = sorted(((rel_err(mm.store.G(k).cpu(), g), k) for k, g in want_g.items()), reverse=True)
        assert worst[0][0] < 3e-4, (network, loss_fn, worst[:5])


def test_align_model_trains_with_fused_trainer_and_generates():
    """The C5 model (align weights) through FusedTrainer (graph-captured step incl. the align target) and generate()
    (the head is skipped when generating, custom_modeling.py:447)."""
    from multimodalanalytical_b200.trainer import FusedTrainer
    fx = load_case("align_conv")
    m = build(fx, "bf16", dropout=0.0)
    tr = FusedTrainer(m)
    l0 = float(tr.train_step(fx["batch"]))
    want = float(fx["ref"]["loss"])
    assert abs(l0 - want) < 2e-2 * abs(want)
    losses = [float(tr.train_step(fx["batch"])) for _ in range(4)]
    assert all(torch.isfinite(torch.tensor(losses)))
    m.eval()
    got = m.generate(fx["batch"], n_beams=3).cpu()
    assert got.shape[0] == fx["ref"]["gen_beam3"].shape[0]
