"""BASELINE.json's configurations at their real model size (custom_model.yaml: d=512, 6+6 layers, 8 heads, ffn 2048)
through the CUDA path (pair GEMM kernel, tcgen05 attention, fused trainer), against the CPU oracle on small batches
that the oracle finishes in seconds, plus size-independent properties at the full C2 batch:

  C3  31P-NMR (one value, 2-layer patch MLP) + formula -> SMILES           (phosphor/formula_num.yaml)
  C4  formula + 1H multiplet tokens + 13C tokens + IR patches + MS/MS and HSQC peak lists, learned pos-enc + GLU
  C5  IR patches of 1800 points + formula, custom_model_align weights, beam-10 decode
"""
import pytest
import torch

from oracle import spectra_oracle as orc
from tests.helpers import oracle_cfg, rel_err

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from multimodalanalytical_b200.trainer import FusedTrainer
    from multimodalanalytical_b200.wrapper import HFWrapper
    from tests.test_model_gpu import FakeTokenizer, build, oracle_grads


def _tok(vocab, target=False):
    return {"type": "text", "target": target, "vocab_size": vocab, "pad_token_id": 0, "preprocessor_arguments": {}}


def _mk(**over):
    mk = dict(model_type="CustomModel", model_name="facebook/bart-base", d_model=512, num_heads=8,
              encoder_attention_heads=8, decoder_attention_heads=8, encoder_layers=6, decoder_layers=6,
              encoder_ffn_dim=2048, decoder_ffn_dim=2048, multimodal_norm=True, positional_encoding_type="sin_cos",
              gated_linear=False, max_position_embeddings=1024, dropout=0.0, align_config=None)
    mk.update(over)
    return mk


def _ragged_tokens(g, L, B, vocab, min_len):
    ids = torch.randint(4, vocab, (L, B), generator=g)
    lens = torch.randint(min_len, L + 1, (B,), generator=g)
    pad = torch.arange(L)[:, None] >= lens[None, :]
    ids[pad] = 0
    return ids, pad


def make_case(name, B, seed=11):
    g = torch.Generator().manual_seed(seed)
    if name == "c3":
        dc = {"Formula": _tok(64), "Phosphor_NMR": {"type": "1D_patches", "target": False, "preprocessor_arguments":
                                                     {"patch_size": 1, "encoding_type": "linear_2_layer"}},
              "Smiles": _tok(64, True)}
        f, fpad = _ragged_tokens(g, 13, B, 64, 6)
        enc = {"Formula": f, "Phosphor_NMR": torch.randn(1, B, 1, generator=g)}
        epad = torch.cat([fpad, torch.zeros(1, B, dtype=torch.bool)], 0)
        T, V, mk = 24, 64, _mk()
    elif name == "c4":
        dc = {"Formula": _tok(64), "Multiplets": dict(_tok(2048), type="multiplets"),
              "Carbon": dict(_tok(2304), type="carbon"),
              "IR": {"type": "1D_patches", "target": False, "preprocessor_arguments": {"patch_size": 75}},
              "MSMS": {"type": "msms_number", "target": False, "preprocessor_arguments": {}},
              "HSQC": {"type": "msms_number", "target": False, "preprocessor_arguments": {}},
              "Smiles": _tok(300, True)}
        f, fpad = _ragged_tokens(g, 16, B, 64, 8)
        m, mpad = _ragged_tokens(g, 120, B, 2048, 30)
        cb, cpad = _ragged_tokens(g, 40, B, 2304, 5)
        cpad[:, 0] = True  # a sample without a 13C spectrum: the whole modality is masked (carbon.py:60-88)
        cb[:, 0] = 0
        enc = {"Formula": f, "Multiplets": m, "Carbon": cb, "IR": torch.randn(23, B, 75, generator=g),
               "MSMS": torch.randn(64, B, 2, generator=g), "HSQC": torch.randn(32, B, 2, generator=g)}
        zeros = lambda n: torch.zeros(n, B, dtype=torch.bool)  # noqa: E731
        epad = torch.cat([fpad, mpad, cpad, zeros(23), zeros(64), zeros(32)], 0)
        T, V, mk = 96, 300, _mk(positional_encoding_type="learned", gated_linear=True)
    elif name == "c5":
        dc = {"Formula": _tok(64), "IR": {"type": "1D_patches", "target": False, "preprocessor_arguments": {"patch_size": 75}},
              "Smiles": _tok(120, True)}
        f, fpad = _ragged_tokens(g, 16, B, 64, 8)
        enc = {"Formula": f, "IR": torch.randn(24, B, 75, generator=g)}
        epad = torch.cat([fpad, torch.zeros(24, B, dtype=torch.bool)], 0)
        T, V = 48, 120
        mk = _mk(align_config=dict(align_network="convolutional", hidden_dimension=256, conv_channels=512, kernel_size=5,
                                   output_dimension=1800, loss_lambda=50, loss_function="mae"))
    else:
        raise KeyError(name)
    t, tpad = _ragged_tokens(g, T + 1, B, V, T // 2)
    t[0] = 2
    batch = {"encoder_input": enc, "encoder_pad_mask": epad, "decoder_input": {"Smiles": t[:-1].contiguous()},
             "decoder_pad_mask": tpad[:-1].contiguous(), "target": t[1:].contiguous()}
    if name == "c5":
        batch["encoder_alignment_input"] = torch.rand(B, 1800, generator=g)
    fx = {"model_kwargs": mk, "data_config": dc, "batch": batch}
    fx["state_dict"] = orc.init_state_dict(oracle_cfg(fx), vocab=V, enc_ffn=2048, dec_ffn=2048, seed=seed)
    return fx


@pytest.mark.parametrize("name,B", [("c3", 32), ("c4", 8), ("c5", 12)])
def test_named_configs_train_step_matches_oracle(name, B):
    """Loss, logits and every parameter gradient of one training step (dropout 0) in bf16 on the tensor-core path."""
    fx = make_case(name, B)
    want_out, want_g = oracle_grads(fx)
    m = build(fx, "bf16")
    m.train()
    m.store.g.zero_()
    out = m.forward(fx["batch"])
    out.loss.backward()
    torch.cuda.synchronize()
    assert rel_err(out.logits.float().cpu(), want_out["logits"].detach()) < 1e-2
    assert abs(float(out.loss) - float(want_out["loss"])) < 1e-2 * abs(float(want_out["loss"]))
    # the MAE align loss has a sign() gradient: where |pred - target| is below the bf16 noise of the encoder output the
    # sign flips, so the head's own gradients (sums over only B rows) get a looser bound; fp32 parity of the head is
    # checked at 3e-4 in test_model_gpu.py
    worst = sorted(((rel_err(m.store.G(k).cpu(), g), k) for k, g in want_g.items() if ".align_network." not in k),
                   reverse=True)
    assert worst[0][0] < 8e-2, worst[:5]
    worst_al = sorted(((rel_err(m.store.G(k).cpu(), g), k) for k, g in want_g.items() if ".align_network." in k),
                      reverse=True)
    assert not worst_al or worst_al[0][0] < 0.3, worst_al[:5]


@pytest.mark.parametrize("gated", [False, True])
def test_post_layer_normalisation_false_at_model_size(gated):
    """`post_layer_normalisation: False` = LN(x + f(x)) layers (custom_modeling.py:119-129,166-176) at d 512 / 6 + 6 layers on
    the tensor-core path: loss, logits and every parameter gradient of a bf16 train step against the oracle (whose post-LN
    branch is pinned to the unmodified reference by tests/golden/post_ln.pt), and fp32 greedy / beam-4 token identity."""
    fx = make_case("c3", 256)  # 6144 decoder rows: the residual-stream accumulations run on the CTA-pair kernel
    fx["model_kwargs"] = dict(fx["model_kwargs"], post_layer_normalisation=False, gated_linear=gated)
    fx["state_dict"] = orc.init_state_dict(oracle_cfg(fx), vocab=64, enc_ffn=2048, dec_ffn=2048, seed=5)
    want_out, want_g = oracle_grads(fx)
    m = build(fx, "bf16")
    m.train()
    m.store.g.zero_()
    out = m.forward(fx["batch"])
    out.loss.backward()
    torch.cuda.synchronize()
    assert rel_err(out.logits.float().cpu(), want_out["logits"].detach()) < 1e-2
    assert abs(float(out.loss) - float(want_out["loss"])) < 1e-2 * abs(float(want_out["loss"]))
    worst = sorted(((rel_err(m.store.G(k).cpu(), g), k) for k, g in want_g.items()), reverse=True)
    assert worst[0][0] < 8e-2, worst[:5]
    # the fused trainer (graph-captured step) runs the same schedule
    tr = FusedTrainer(m, clip_grad=1.0)
    l0 = float(tr.train_step(fx["batch"], 0))
    for i in range(1, 4):
        l1 = float(tr.train_step(fx["batch"], i))
    assert l1 < l0
    cfg = oracle_cfg(fx)
    cfg.max_length = 24
    m32 = build(fx, "fp32")
    m32.eval()
    m32.generation_config["max_length"] = 24
    small = {k: (v[:, :4] if torch.is_tensor(v) else {kk: vv[:, :4] for kk, vv in v.items()}) for k, v in fx["batch"].items()}
    for k in (1, 4):
        got = m32.generate(small, n_beams=k).cpu()
        with torch.no_grad():
            want = orc.generate(fx["state_dict"], cfg, small, n_beams=k)
        assert got.shape == want.shape and torch.equal(got, want), k
    # bf16 decode of the same model runs the per-op post-LN step
    m.eval()
    m.generation_config["max_length"] = 24
    got16 = m.generate(small, n_beams=4).cpu()
    assert got16.shape[0] == 16 and bool((got16[:, 0] == 2).all())


def test_c5_beam10_decode_identical_to_oracle_fp32():
    """IR-of-mixtures model (align weights present, head unused while generating): greedy and beam-10 sequences."""
    fx = make_case("c5", 3)
    cfg = oracle_cfg(fx)
    cfg.max_length = 40  # keeps the cache-less O(T^2) oracle to a few seconds; ForcedEOS fires at max_length - 1
    m = build(fx, "fp32")
    m.eval()
    m.generation_config["max_length"] = 40
    for k in (1, 10):
        got = m.generate(fx["batch"], n_beams=k).cpu()
        with torch.no_grad():
            want = orc.generate(fx["state_dict"], cfg, fx["batch"], n_beams=k)
        assert got.shape == want.shape and torch.equal(got, want), k


def test_c4_modality_dropout_shifts_positions_like_the_reference():
    """wrapper.py:367-386: dropped modalities are removed from the input AND the mask, so later modalities move up in
    the positional encoding; the result must equal a batch that never had those modalities."""
    import numpy as np
    from multimodalanalytical_b200.wrapper import ListConfig
    fx = make_case("c4", 4)
    m = build(fx, "fp32", modality_dropout=ListConfig(["IR", "Multiplets", "Carbon"]))
    m.train()
    np.random.seed(5)
    st = np.random.get_state()
    drop = list(np.random.choice(["IR", "Multiplets", "Carbon"], np.random.randint(0, 3), replace=False))
    np.random.set_state(st)
    with torch.no_grad():
        got = m.forward(fx["batch"])
    # the same batch without the dropped modalities, evaluated by the oracle
    b2 = dict(fx["batch"])
    sizes = {k: (v.shape[0]) for k, v in fx["batch"]["encoder_input"].items()}
    keep_rows, off = [], 0
    for k, n in sizes.items():
        if k not in drop:
            keep_rows.append(torch.arange(off, off + n))
        off += n
    b2["encoder_input"] = {k: v for k, v in fx["batch"]["encoder_input"].items() if k not in drop}
    b2["encoder_pad_mask"] = fx["batch"]["encoder_pad_mask"][torch.cat(keep_rows)]
    with torch.no_grad():
        want = orc.wrapper_forward(fx["state_dict"], oracle_cfg(fx), b2)
    assert rel_err(got.logits.cpu(), want["logits"]) < 2e-5, drop
    assert abs(float(got.loss) - float(want["loss"])) < 2e-5 * float(want["loss"])


def _c2_batch(B, seed):
    import bench
    return bench.synth_batch(bench.C2, B, seed), bench


def test_c2_full_batch_properties():
    """Full-size C2 batch (256 spectra): (1) the forward pass is bit-reproducible, (2) gradients are linear in
    the batch: grad(256) == mean of grad over its two halves (dropout 0, fp32 accumulation), (3) the loss of the
    CUDA path at B=256 equals the mean of per-chunk oracle losses on a 32-sample slice within bf16 tolerance."""
    batch, bench = _c2_batch(256, 1)
    c = bench.C2

    def fresh():
        m = HFWrapper(data_config=bench.data_config(c), target_tokenizer=bench.Tok(c["V"]), num_steps=100,
                      precision="bf16", seed=3, **bench.model_kwargs(c, dropout=0.0))
        return m

    def grads_of(m, b):
        m.train()
        m.store.g.zero_()
        out = m.forward(b)
        out.loss.backward()
        torch.cuda.synchronize()
        return float(out.loss), m.store.g.clone()

    m = fresh()
    l_full, g_full = grads_of(m, batch)
    l_again, g_again = grads_of(m, batch)
    # forward has no atomics: the loss is bit-reproducible; bias / LayerNorm / embedding gradients use fp32 atomics
    assert l_full == l_again, "forward pass is not bit-reproducible"
    assert rel_err(g_again, g_full) < 1e-5

    def half(b, lo, hi):
        return bench.map_batch(b, lambda x: x[:, lo:hi].contiguous())

    la, ga = grads_of(m, half(batch, 0, 128))
    lb, gb = grads_of(m, half(batch, 128, 256))
    assert abs(0.5 * (la + lb) - l_full) < 2e-3 * l_full
    assert rel_err(0.5 * (ga + gb), g_full) < 2e-2

    # oracle on a 32-sample slice with the same weights
    sl = half(batch, 0, 32)
    l32, _ = grads_of(m, sl)
    sd = {k: v.detach().float().cpu() for k, v in m.state_dict().items()}
    fx = {"model_kwargs": dict(bench.model_kwargs(c, dropout=0.0), align_config=None), "data_config": bench.data_config(c)}
    with torch.no_grad():
        want = orc.wrapper_forward(sd, oracle_cfg(fx), sl)
    assert abs(l32 - float(want["loss"])) < 1e-2 * float(want["loss"])


def test_decode_is_batch_invariant_and_beam1_equals_greedy_at_scale():
    """64 spectra decoded together give exactly the sequences of the same spectra decoded in chunks of 16 (fp32),
    and beam search with one beam equals greedy decoding."""
    batch, bench = _c2_batch(64, 9)
    c = bench.C2
    m = HFWrapper(data_config=bench.data_config(c), target_tokenizer=bench.Tok(c["V"]), num_steps=100,
                  precision="fp32", seed=4, **bench.model_kwargs(c))
    m.eval()
    m.generation_config["max_length"] = 24
    whole = m.generate(batch, n_beams=4).cpu()
    parts = []
    for lo in range(0, 64, 16):
        sub = bench.map_batch(batch, lambda x: x[:, lo:lo + 16].contiguous())
        parts.append(m.generate(sub, n_beams=4).cpu())
    L = max(p.shape[1] for p in parts)
    assert whole.shape[1] == L
    for i, p in enumerate(parts):
        assert torch.equal(whole[i * 64: (i + 1) * 64, :p.shape[1]], p), i
    greedy = m.generate(batch, n_beams=1).cpu()
    assert greedy.shape[0] == 64 and (greedy[:, 0] == 2).all()


def test_fused_trainer_loss_decreases_on_c3():
    fx = make_case("c3", 64)
    m = build(fx, "bf16", dropout=0.1, lr=3e-4)
    tr = FusedTrainer(m)
    losses = [float(tr.train_step(fx["batch"], i)) for i in range(12)]
    assert losses[-1] < losses[0] - 0.2, losses


def test_validation_epoch_runs_greedy_decode_and_token_accuracy():
    """N4: trainer.validate = the reference's validation loop (loss, token accuracy, greedy Top-1) over batches."""
    from multimodalanalytical_b200.trainer import validate
    fx = make_case("c3", 16)
    m = build(fx, "bf16")
    m.generation_config["max_length"] = 32
    res = validate(m, [fx["batch"], fx["batch"]])
    assert set(res) >= {"val_loss", "val_token_acc", "val_molecular_accuracy"}
    assert res["val_loss"] > 0 and 0.0 <= res["val_token_acc"] <= 1.0 and 0.0 <= res["val_molecular_accuracy"] <= 1.0
    assert m.validation_step_outputs == []


def test_side_stream_weight_gradients_match_single_stream(monkeypatch):
    """MMA_WGRAD_STREAM=1: grouped wgrad launches on a side stream with rotating operand buffer sets must produce
    the gradients of the single-stream schedule (eager and inside the captured step)."""
    fx = make_case("c3", 32)
    ref = build(fx, "bf16")
    ref.train()
    ref.store.g.zero_()
    ref.forward(fx["batch"]).loss.backward()
    torch.cuda.synchronize()
    want = ref.store.g.clone()
    monkeypatch.setenv("MMA_WGRAD_STREAM", "1")
    m = build(fx, "bf16")
    assert m.engine.wgrad_stream is not None
    m.train()
    m.store.g.zero_()
    m.forward(fx["batch"]).loss.backward()
    torch.cuda.synchronize()
    assert rel_err(m.store.g, want) < 1e-5
    tr = FusedTrainer(m)
    losses = [float(tr.train_step(fx["batch"], i)) for i in range(4)]  # step 2+ replays the captured graph
    assert all(l == l for l in losses) and losses[-1] < losses[0]


def _width_case(d, heads, ffn, layers, B, seed=5):
    """IR + formula -> SMILES at the widths of configs/model/custom_model_base.yaml (768 / 12 heads / 3072) and
    custom_model_large.yaml (1024 / 16 / 4096); fewer layers than the yaml so the oracle stays in seconds."""
    g = torch.Generator().manual_seed(seed)
    dc = {"Formula": _tok(64), "IR": {"type": "1D_patches", "target": False, "preprocessor_arguments": {"patch_size": 75}},
          "Smiles": _tok(200, True)}
    f, fpad = _ragged_tokens(g, 15, B, 64, 6)
    enc = {"Formula": f, "IR": torch.randn(21, B, 75, generator=g)}
    epad = torch.cat([fpad, torch.zeros(21, B, dtype=torch.bool)], 0)
    T, V = 40, 200
    t, tpad = _ragged_tokens(g, T + 1, B, V, T // 2)
    t[0] = 2
    batch = {"encoder_input": enc, "encoder_pad_mask": epad, "decoder_input": {"Smiles": t[:-1].contiguous()},
             "decoder_pad_mask": tpad[:-1].contiguous(), "target": t[1:].contiguous()}
    mk = _mk(d_model=d, num_heads=heads, encoder_attention_heads=heads, decoder_attention_heads=heads,
             encoder_layers=layers, decoder_layers=layers, encoder_ffn_dim=ffn, decoder_ffn_dim=ffn)
    fx = {"model_kwargs": mk, "data_config": dc, "batch": batch}
    fx["state_dict"] = orc.init_state_dict(oracle_cfg(fx), vocab=V, enc_ffn=ffn, dec_ffn=ffn, seed=seed)
    return fx


@pytest.mark.parametrize("d,heads,ffn,layers", [(768, 12, 3072, 2), (1024, 16, 4096, 2)])
def test_base_and_large_widths_match_oracle(d, heads, ffn, layers):
    """custom_model_base / custom_model_large widths: fp32 logits + loss at 1e-5, bf16 train step (loss, logits, every
    gradient) at the bf16 bounds, greedy + beam-4 decode token-identical in fp32."""
    fx = _width_case(d, heads, ffn, layers, B=24)
    want_out, want_g = oracle_grads(fx)
    m32 = build(fx, "fp32")
    m32.eval()
    with torch.no_grad():
        out32 = m32.forward(fx["batch"])
    assert rel_err(out32.logits.cpu(), want_out["logits"].detach()) < 1e-5
    assert abs(float(out32.loss) - float(want_out["loss"])) < 1e-5 * abs(float(want_out["loss"]))
    m = build(fx, "bf16")
    m.train()
    m.store.g.zero_()
    out = m.forward(fx["batch"])
    out.loss.backward()
    torch.cuda.synchronize()
    assert rel_err(out.logits.float().cpu(), want_out["logits"].detach()) < 1e-2
    assert abs(float(out.loss) - float(want_out["loss"])) < 1e-2 * abs(float(want_out["loss"]))
    worst = sorted(((rel_err(m.store.G(k).cpu(), g), k) for k, g in want_g.items()), reverse=True)
    assert worst[0][0] < 8e-2, worst[:5]
    # decode on 3 spectra, short horizon (the oracle re-decodes the whole prefix every step)
    sub = {"encoder_input": {k: v[:, :3] for k, v in fx["batch"]["encoder_input"].items()},
           "encoder_pad_mask": fx["batch"]["encoder_pad_mask"][:, :3],
           "decoder_input": {"Smiles": fx["batch"]["decoder_input"]["Smiles"][:, :3]},
           "decoder_pad_mask": fx["batch"]["decoder_pad_mask"][:, :3], "target": fx["batch"]["target"][:, :3]}
    cfg = oracle_cfg(fx)
    cfg.max_length = 24
    m32.generation_config["max_length"] = 24
    for k in (1, 4):
        got = m32.generate(sub, n_beams=k).cpu()
        with torch.no_grad():
            want = orc.generate(fx["state_dict"], cfg, sub, n_beams=k)
        assert got.shape == want.shape and torch.equal(got, want), (d, k)
