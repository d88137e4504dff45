"""Collator -> device pipeline against batches produced by the reference's own collator
(tests/golden/collate.pt, written by tests/golden/make_collate_golden.py from the unmodified
`MultiModalDataCollator` + `load_preprocessors`).  Integer / bool tensors must be identical; float tensors 2e-6."""
import os
import types

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN_DIR, load_case

gpu = pytest.mark.gpu
_FX = {}


def fixture():
    if not _FX:
        _FX.update(torch.load(os.path.join(GOLDEN_DIR, "collate.pt"), weights_only=False))
    return _FX


def _pad_id(fx, m):
    return fx["data_config"][m].get("pad_token_id", 0)


def _same(got, want, name):
    if isinstance(want, dict):
        assert set(got) == set(want), name
        for k in want:
            _same(got[k], want[k], f"{name}.{k}")
        return
    if isinstance(want, torch.Tensor):
        got = got.cpu()
        assert got.shape == want.shape, (name, got.shape, want.shape)
        if want.is_floating_point():
            assert torch.allclose(got, want.float(), atol=2e-6, rtol=0), name
        else:
            assert got.dtype == want.dtype, (name, got.dtype, want.dtype)
            assert torch.equal(got, want), name
        return
    assert got == want, name



# ------------------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("case", ["c1", "multi", "spectext"])
def test_ragged_rows_are_the_valid_prefixes_of_the_reference_batch(case):
    """Host-side extraction: every ragged row equals the un-padded prefix of the reference's padded row, `valid`
    mirrors the fully-masked samples, lengths respect the truncation bounds."""
    fx = fixture()[case]
    host = fx["host"]
    full = fx["batches"][0]  # all samples in order
    assert full["indices"] == list(range(host.n))
    b = full["batch"]
    off = 0
    for m, col in host.columns.items():
        ref = b["encoder_input"][m]
        ids = ref["tokenized_input"] if isinstance(ref, dict) else ref
        S = ids.shape[0]
        pad = b["encoder_pad_mask"][off: off + S]
        off += S
        if col.kind == "tokens":
            assert int(col.tokens.lengths.max()) <= col.max_len
            for i in range(host.n):
                row = col.tokens.row(i)
                assert np.array_equal(ids[: len(row), i].numpy(), row)
                assert bool((ids[len(row):, i] == col.pad_id).all())
                valid = True if col.tokens.valid is None else bool(col.tokens.valid[i])
                assert np.array_equal((~pad[:, i]).numpy(), (np.arange(S) < len(row)) & valid)
            if col.values is not None:
                for i in range(host.n):
                    v = col.values.row(i)[:, 0]
                    assert np.array_equal(ref["numerical_values"][: len(v), i].numpy(), v)
                    assert bool((ref["numerical_values"][len(v):, i] == col.pad_value).all())
        elif col.kind == "values":
            for i in range(host.n):
                v = col.values.row(i)
                assert np.array_equal(ids[: len(v), i].numpy(), v)
                assert np.array_equal((~pad[:, i]).numpy(), np.arange(S) < len(v))
        else:
            assert col.raw.shape[0] == host.n and col.missing.dtype == np.uint8
            assert np.array_equal(pad.all(dim=0).numpy(), col.missing.astype(bool))
    assert off == b["encoder_pad_mask"].shape[0]
    tgt = host.target
    full_ids = torch.cat([b["decoder_input"][host.target_modality][:1], b["target"]], dim=0)
    for i in range(host.n):
        row = tgt.tokens.row(i)
        assert len(row) <= tgt.max_len
        assert np.array_equal(full_ids[: len(row), i].numpy(), row)
    assert host.passthrough["target_smiles"] == b["target_smiles"]


@pytest.mark.parametrize("case", ["c1", "multi", "spectext"])
def test_collate_oracle_matches_reference_batches(case):
    """oracle/collate_oracle.py (numpy restatement of the collator's padding / masking / shifting on the ragged columns)
    reproduces every batch the reference collator produced, for every index list of the fixture."""
    from oracle import collate_oracle

    fx = fixture()[case]
    for entry in fx["batches"]:
        got = collate_oracle.collate(fx["host"], entry["indices"])
        want = entry["batch"]
        assert set(got) == set(want), set(got) ^ set(want)
        assert list(got["encoder_input"]) == list(want["encoder_input"])
        for k in want:
            _same(got[k], want[k], f"{case}{entry['indices'][:4]}.{k}")


def test_pretokenise_from_raw_rows_with_rebuilt_tokenizers():
    """`pretokenise` re-run here on the c1 raw rows with tokenizers rebuilt from their JSON gives the fixture's
    HostDataset (which the golden script extracted next to the reference collator)."""
    from tokenizers import Tokenizer
    from transformers import PreTrainedTokenizerFast

    from multimodalanalytical_b200.pipeline import pretokenise

    fx = fixture()["c1"]
    pre = {m: PreTrainedTokenizerFast(tokenizer_object=Tokenizer.from_str(js), pad_token="<pad>", unk_token="<unk>",
                                      eos_token="<eos>", bos_token="<bos>", model_max_length=512)
           for m, js in fx["tokenizers"].items()}
    pre["IR"] = types.SimpleNamespace(**fx["patch"])
    host = pretokenise(fx["rows"], pre, fx["data_config"], fx["max_source_length"], fx["max_target_length"])
    want = fx["host"]
    assert host.n == want.n and list(host.columns) == list(want.columns)
    for m, col in host.columns.items():
        w = want.columns[m]
        assert (col.kind, col.pad_len, col.max_len, col.pad_id) == (w.kind, w.pad_len, w.max_len, w.pad_id)
        if col.tokens is not None:
            assert np.array_equal(col.tokens.flat, w.tokens.flat) and np.array_equal(col.tokens.offsets, w.tokens.offsets)
        if col.raw is not None:
            assert np.array_equal(col.raw, w.raw) and np.array_equal(col.missing, w.missing)
            assert col.patch == {"derivative": False, **w.patch}  # the committed c1 fixture predates the derivative key
    assert np.array_equal(host.target.tokens.flat, want.target.tokens.flat)
    assert host.passthrough["target_smiles"] == want.passthrough["target_smiles"]


def test_pretokenise_error_conventions():
    from multimodalanalytical_b200.pipeline import pretokenise

    dc = {"A": {"type": "text", "target": True}, "B": {"type": "text", "target": True}}
    with pytest.raises(ValueError, match="Only 1 target"):  # datamodules.py:57-60
        pretokenise({"A": ["x"], "B": ["y"]}, {}, dc, {}, 8)
    # the collator's `token_indices` dict (datamodules.py:306-319) is something the reference's own embedding cannot take
    # (modeling/utils.py:154-160 reads `numerical_values`): that modality type is not carried over
    dc = {"A": {"type": "peak_positional_encoding", "target": False}, "B": {"type": "text", "target": True}}
    with pytest.raises(NotImplementedError):
        pretokenise({"A": ["x"], "B": ["y"]}, {"A": None}, dc, {}, 8)


@pytest.mark.parametrize("n,bs,world", [(103, 16, 1), (103, 16, 2), (64, 8, 4), (7, 4, 2)])
def test_index_sampler_partitions_like_distributed_sampler(n, bs, world):
    from torch.utils.data.distributed import DistributedSampler

    from multimodalanalytical_b200.pipeline import IndexSampler

    seen = []
    for rank in range(world):
        s = IndexSampler(n, bs, shuffle=True, seed=5, rank=rank, world=world)
        s.set_epoch(3)
        idx = np.concatenate(list(s))
        assert len(s) == -(-len(idx) // bs)
        # same partition rule as torch's DistributedSampler: padded permutation, rank-strided
        ds = DistributedSampler(range(n), num_replicas=world, rank=rank, shuffle=False)
        assert len(idx) == len(list(ds))
        order = s.indices()
        assert np.array_equal(idx, order)
        seen.append(idx)
        # different epoch -> different order, same epoch -> same order
        s2 = IndexSampler(n, bs, shuffle=True, seed=5, rank=rank, world=world)
        s2.set_epoch(3)
        assert np.array_equal(s2.indices(), order)
        s2.set_epoch(4)
        assert not np.array_equal(s2.indices(), order) or n < 3
    allidx = np.concatenate(seen)
    assert set(allidx.tolist()) == set(range(n))          # every sample is visited
    assert len(allidx) == -(-n // world) * world          # ranks do equal work (wrap-around padding)
    plain = IndexSampler(n, bs, shuffle=False, drop_last=True)
    assert all(len(c) == bs for c in plain) and len(plain) == n // bs


def test_device_dataset_refuses_cpu():
    from multimodalanalytical_b200.pipeline import DeviceDataset

    with pytest.raises(RuntimeError, match="CUDA"):
        DeviceDataset(fixture()["c1"]["host"], device="cpu")


# ------------------------------------------------------------------------------------------------------- GPU
@gpu
@pytest.mark.parametrize("case", ["c1", "multi", "spectext"])
def test_gpu_wire_batches_identical_to_reference_collator(case):
    from multimodalanalytical_b200.pipeline import DeviceDataset

    fx = fixture()[case]
    ds = DeviceDataset(fx["host"])
    assert len(ds) == fx["host"].n and ds.bytes_resident() > 0
    for entry in fx["batches"]:
        got = ds.wire_batch(entry["indices"])
        want = entry["batch"]
        assert set(got) == set(want), (set(got) ^ set(want))
        assert list(got["encoder_input"]) == list(want["encoder_input"])  # modality order = positional order
        for k in want:
            _same(got[k], want[k], f"{case}{entry['indices'][:4]}.{k}")
    with pytest.raises(IndexError):
        ds.collate([0, len(ds)])


@gpu
def test_gpu_engine_batch_trains_like_the_reference_batch():
    """collate() output goes straight into the trainer; loss and every gradient equal those of the reference
    collator's batch through HFWrapper (same model, fp32, dropout off)."""
    from multimodalanalytical_b200.pipeline import DeviceDataset, DeviceLoader
    from multimodalanalytical_b200.wrapper import HFWrapper
    from tests.test_model_gpu import FakeTokenizer

    fx, cx = fixture()["c1"], load_case("c1_ir_tiny")
    mk = dict(cx["model_kwargs"])
    mk["dropout"] = 0.0
    tok = FakeTokenizer(cx["data_config"]["Smiles"]["vocab_size"])

    def model():
        m = HFWrapper(data_config=cx["data_config"], target_tokenizer=tok, num_steps=100, precision="fp32", **mk)
        m.load_state_dict(cx["state_dict"])
        return m

    ds = DeviceDataset(fx["host"])
    idx = fx["batches"][3]["indices"]
    a, b = model(), model()
    a.train()
    b.train()
    enc, enc_mask, dec_in, dec_mask, labels = ds.collate(idx)
    out_a = a.engine.forward(enc, enc_mask, dec_in, dec_mask, labels=labels, train=True)
    a.engine.backward(gscale=1.0)
    loss_b = b.forward(fx["batches"][3]["batch"]).loss
    loss_b.backward()
    assert abs(float(out_a["loss"]) - float(loss_b)) <= 1e-6 * abs(float(loss_b))
    assert torch.allclose(a.store.g, b.store.g, rtol=1e-5, atol=1e-7)
    # the loader walks the whole set once per epoch, in the engine's layout
    loader = DeviceLoader(ds, batch_size=4, shuffle=True, seed=1)
    seen = 0
    for batch in loader:
        assert isinstance(batch, tuple) and batch[1].dtype == torch.uint8 and batch[4].dtype == torch.int64
        seen += batch[1].shape[0]
    assert seen == len(ds) and len(loader) == 4


@gpu
def test_gpu_fused_trainer_takes_device_batches():
    from multimodalanalytical_b200.pipeline import DeviceDataset, DeviceLoader
    from multimodalanalytical_b200.trainer import FusedTrainer
    from multimodalanalytical_b200.wrapper import HFWrapper
    from tests.test_model_gpu import FakeTokenizer

    fx, cx = fixture()["c1"], load_case("c1_ir_tiny")
    tok = FakeTokenizer(cx["data_config"]["Smiles"]["vocab_size"])
    m = HFWrapper(data_config=cx["data_config"], target_tokenizer=tok, num_steps=100, precision="bf16", **cx["model_kwargs"])
    m.load_state_dict(cx["state_dict"])
    tr = FusedTrainer(m)
    ds = DeviceDataset(fx["host"])
    losses = []
    for epoch in range(6):
        loader = DeviceLoader(ds, batch_size=len(ds), shuffle=False)
        for batch in loader:
            losses.append(float(tr.train_step(batch)))
    assert losses[-1] < losses[0]  # same shape every step: eager, capture, then graph replays; the loss goes down


@gpu
def test_gpu_predict_sharded_equals_one_big_batch():
    """Sharded prediction (strided shard, device-assembled wire batches of 4) returns, in dataset order, the same
    hypotheses as one batch holding the whole set."""
    from multimodalanalytical_b200.pipeline import DeviceDataset
    from multimodalanalytical_b200.trainer import predict_sharded, shard_indices
    from multimodalanalytical_b200.wrapper import HFWrapper
    from tests.test_guided import VocabTokenizer

    fx, cx = fixture()["c1"], load_case("c1_ir_tiny")
    tok = VocabTokenizer(cx["smiles_vocab"])
    m = HFWrapper(data_config=cx["data_config"], target_tokenizer=tok, num_steps=100, precision="fp32",
                  **cx["model_kwargs"])
    m.load_state_dict(cx["state_dict"])
    m.eval()
    ds = DeviceDataset(fx["host"])
    n, K = len(ds), 3
    got = predict_sharded(m, ds, batch_size=4, n_beams=K)
    seqs = m.generate(ds.wire_batch(list(range(n))), n_beams=K)
    dec = tok.batch_decode(seqs, skip_special_tokens=True)
    assert got == [dec[i * K: (i + 1) * K] for i in range(n)]
    # and the reference's own golden beam-3 ids for this set (make_golden.case_c1 used the same rows in order)
    assert torch.equal(seqs.cpu(), cx["ref"]["gen_beam3"])
    assert shard_indices(n, 1, 2) == list(range(1, n, 2))


@gpu
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_gpu_collate_kernels_random_ragged(seed):
    """Collate kernels against a numpy restatement on random ragged columns: empty rows, rows longer than the
    truncation bound, rows flagged as missing, repeated and out-of-order indices, one-sample batches, width-2 values."""
    from multimodalanalytical_b200 import ops
    from multimodalanalytical_b200.pipeline import Ragged

    rng = np.random.default_rng(seed)
    N, pad_id, max_len = 57, 0, 19
    lens = rng.integers(0, 30, size=N)
    lens[3] = 0
    rows = [rng.integers(4, 90, size=n).astype(np.int32) for n in lens]
    valid = (rng.random(N) > 0.2).astype(np.uint8)
    col = Ragged.from_rows(rows, np.int32, valid=valid)
    vals = Ragged.from_rows([rng.standard_normal((n, 2)).astype(np.float32) for n in lens], np.float32, width=2)
    dev = torch.device("cuda")
    flat, off, vd = (torch.from_numpy(x).to(dev) for x in (col.flat, col.offsets, col.valid))
    vflat, voff = torch.from_numpy(vals.flat).to(dev), torch.from_numpy(vals.offsets).to(dev)
    for B in (1, 7, 64):
        idx = rng.integers(0, N, size=B).astype(np.int32)
        tl = np.minimum(lens[idx], max_len)
        for L in (int(max(tl.max(), 1)), 24):
            ids = torch.empty(B, L, dtype=torch.int64, device=dev)
            mask = torch.empty(B, L, dtype=torch.uint8, device=dev)
            r = torch.from_numpy(idx).to(dev)
            ops.collate_tokens(flat, off, vd, r, pad_id, max_len, ids, mask)
            want_ids = np.full((B, L), pad_id, dtype=np.int64)
            want_mask = np.zeros((B, L), dtype=np.uint8)
            for b, i in enumerate(idx):
                n = min(tl[b], L)
                want_ids[b, :n] = rows[i][:n]
                want_mask[b, :n] = valid[i]
            assert np.array_equal(ids.cpu().numpy(), want_ids) and np.array_equal(mask.cpu().numpy(), want_mask)
            out = torch.empty(B, L, 2, device=dev)
            vmask = torch.empty(B, L, dtype=torch.uint8, device=dev)
            ops.collate_values(vflat, voff, r, -1.5, max_len, out, vmask)
            want_v = np.full((B, L, 2), -1.5, dtype=np.float32)
            want_vm = np.zeros((B, L), dtype=np.uint8)
            for b, i in enumerate(idx):
                n = min(tl[b], L)
                want_v[b, :n] = vals.row(i)[:n]
                want_vm[b, :n] = 1
            assert np.array_equal(out.cpu().numpy(), want_v) and np.array_equal(vmask.cpu().numpy(), want_vm)
        # teacher forcing: T = longest (truncated) row - 1, at least 1
        T = int(max(tl.max() - 1, 1))
        dec_in = torch.empty(B, T, dtype=torch.int64, device=dev)
        dec_mask = torch.empty(B, T, dtype=torch.uint8, device=dev)
        labels = torch.empty(B, T, dtype=torch.int64, device=dev)
        ops.collate_target(flat, off, torch.from_numpy(idx).to(dev), pad_id, max_len, dec_in, dec_mask, labels)
        w_in = np.full((B, T), pad_id, dtype=np.int64)
        w_m = np.zeros((B, T), dtype=np.uint8)
        w_l = np.full((B, T), -100, dtype=np.int64)
        for b, i in enumerate(idx):
            t = rows[i][: tl[b]]
            n = min(len(t), T)
            w_in[b, :n] = t[:n]
            w_m[b, :n] = 1
            m = min(max(len(t) - 1, 0), T)
            w_l[b, :m] = t[1: 1 + m]
        assert np.array_equal(dec_in.cpu().numpy(), w_in) and np.array_equal(dec_mask.cpu().numpy(), w_m)
        assert np.array_equal(labels.cpu().numpy(), w_l)
