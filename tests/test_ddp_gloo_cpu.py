"""world_size-2 gloo test of the N>1 host logic: back-to-front bucketed gradient all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodalanalytical_b200.trainer import GradBucketer


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, bucket, offsets, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.arange(n, dtype=torch.float32) * (rank + 1)
    b = GradBucketer(g, bucket)
    b.reset()
    for off in offsets:
        b.on_ready(off)
        # nothing below the reported offset may have been reduced yet
        assert all(lo >= off for lo, _ in b.launched)
    b.finish()
    want = torch.arange(n, dtype=torch.float32) * sum(r + 1 for r in range(world))
    covered = sorted(b.launched)
    ok = torch.equal(g, want) and covered[0][0] == 0 and covered[-1][1] == n and \
        all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))
    q.put((rank, bool(ok), len(covered)))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    world, n, bucket = 2, 10_000, 3_000
    offsets = [9_500, 7_000, 6_900, 2_000, 100, 0]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, bucket, offsets, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(nb >= 3 for _, _, nb in res)


def _shard_worker(rank, world, port, n, q):
    from multimodalanalytical_b200.trainer import gather_outputs, shard_indices

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(n, rank, world)
    local = [[f"mol{i}-beam{k}" for k in range(3)] for i in mine]  # what a rank's decode would return
    full = gather_outputs(local, mine, n)
    ok = full == [[f"mol{i}-beam{k}" for k in range(3)] for i in range(n)]
    # a sample decoded twice, or never, is an error - not a silent overwrite
    try:
        gather_outputs(local, [0] * len(mine), n)
        dup = False
    except ValueError:
        dup = True
    q.put((rank, ok, dup, len(mine)))
    dist.destroy_process_group()


def test_sharded_inference_partition_and_gather_world2():
    """Inference shards independent spectra over the ranks with no data-path collective; the decoded strings are
    exchanged once at the end."""
    from multimodalanalytical_b200.trainer import shard_indices

    for n, world in ((11, 2), (8, 4), (3, 4), (0, 2)):
        for contiguous in (False, True):
            parts = [shard_indices(n, r, world, contiguous) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= (1 if not contiguous else -(-n // world))
    world, n = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok and dup for _, ok, dup, _ in res), res
    assert sorted(m for *_, m in res) == [5, 6]


# ---------------------------------------------------------------------------------------------------------------------
# The Lightning / torch-optimizer path under a STOCK DistributedDataParallel wrapper (reference: Lightning DDP,
# trainer/trainer.py:58-71).  The CUDA engine cannot run here, so a stand-in engine supplies loss and gradients; what is
# tested is the plumbing that ADVICE r1 found broken: parameters receive their gradients through autograd, so
# (1) `optimizer.zero_grad(set_to_none=True)` before every backward does not lose them, (2) gradient accumulation adds up,
# (3) DDP's reducer hooks fire and average the gradients over the ranks, (4) a torch optimiser actually moves the weights.
# ---------------------------------------------------------------------------------------------------------------------
class _FakeEngine:
    """loss = sum_p <w_p, c_p> * scale(batch); d loss / d w_p = c_p * scale.  Same surface as model.Engine for forward/backward."""

    def __init__(self, module):
        self.m = module
        self.cfg = module.engine.cfg
        self.precision = "fp32"
        self.scale = 1.0

    def next_seed(self):
        pass

    def forward(self, input_ids, attention_mask, dec_in, dec_mask, labels=None, train=False, align_target=None):
        self.scale = float(dec_in.float().mean())
        ps = self.m.store
        loss = sum((ps.P(n).double() * self._c(n)).sum() for n in self.m._names) * self.scale
        B, T = dec_in.shape
        return {"logits": torch.zeros(B, T, self.cfg.vocab_size), "loss": loss.float(), "lm_loss": loss.float(),
                "align_loss": None, "memory": None}

    def _c(self, name):
        off, shape = self.m.store.offsets[name]
        n = int(torch.tensor(shape).prod())
        return (torch.arange(n, dtype=torch.float64).view(shape) % 7 - 3) * 1e-3

    def backward(self, gscale=1.0):
        ps = self.m.store
        for n in self.m._names:
            ps.G(n).add_((self._c(n) * self.scale * gscale).float())


def _tiny_module():
    from multimodalanalytical_b200.wrapper import HFWrapper

    dc = {"Formula": {"type": "text", "target": False, "vocab_size": 12, "pad_token_id": 0, "preprocessor_arguments": {}},
          "Smiles": {"type": "text", "target": True, "vocab_size": 11, "pad_token_id": 0, "preprocessor_arguments": {}}}

    class Tok:
        vocab_size, pad_token_id, bos_token_id, eos_token_id = 11, 0, 2, 3

    m = HFWrapper(data_config=dc, model_type="CustomModel", model_name="x", target_tokenizer=Tok(), num_steps=20,
                  optimiser="adamw", lr=1e-2, device="cpu", precision="fp32", d_model=16, num_heads=2,
                  encoder_layers=1, decoder_layers=1, encoder_ffn_dim=32, decoder_ffn_dim=32, seed=5)
    m.engine = _FakeEngine(m)
    return m


def _batch(level):
    S, T, B = 3, 4, 2
    return {"encoder_input": {"Formula": torch.full((S, B), 4)}, "encoder_pad_mask": torch.zeros(S, B, dtype=torch.bool),
            "decoder_input": {"Smiles": torch.full((T, B), level)}, "decoder_pad_mask": torch.zeros(T, B, dtype=torch.bool),
            "target": torch.full((T, B), 5)}


def test_torch_optimizer_path_survives_zero_grad_and_accumulates():
    m = _tiny_module()
    (opt,), (sch,) = m.configure_optimizers()
    name = "hf_model.token_ff.weight"
    w0 = m.store.P(name).clone()
    for step in range(3):
        opt.zero_grad()  # set_to_none=True: what Lightning does before every backward
        loss = m.training_step(_batch(4 + step), step)
        loss.backward()
        g = m.named_gradients()
        assert all(v is not None for v in g.values())
        want = (m.engine._c(name) * (4 + step)).float()
        assert torch.allclose(g[name], want, rtol=1e-6), step
        opt.step()
        sch["scheduler"].step()
    assert not torch.equal(m.store.P(name), w0), "the optimiser did not move the weights"
    assert m.store.bf16_dirty  # the engine must refresh its bf16 mirror before the next forward
    # accumulation: two backwards without zero_grad add up
    opt.zero_grad()
    m.training_step(_batch(4), 0).backward()
    m.training_step(_batch(6), 1).backward()
    assert torch.allclose(m.named_gradients()[name], (m.engine._c(name) * 10).float(), rtol=1e-6)


def _ddp_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"] = str(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = _tiny_module()
    ddp = torch.nn.parallel.DistributedDataParallel(m, find_unused_parameters=True)  # the reference's strategy
    (opt,), _ = m.configure_optimizers()
    name = "hf_model.token_ff.weight"
    ok = True
    for step in range(2):
        opt.zero_grad()
        out = ddp(_batch(4 + 2 * rank + step))  # DDP.forward -> HFWrapper.forward; ranks see different batches
        out.loss.backward()
        mean_level = sum(4 + 2 * r + step for r in range(world)) / world
        want = (m.engine._c(name) * mean_level).float()
        ok = ok and torch.allclose(m.named_gradients()[name], want, rtol=1e-6)
        opt.step()
    # replicas stay identical
    ws = [torch.zeros_like(m.store.p) for _ in range(world)]
    dist.all_gather(ws, m.store.p)
    same = all(torch.equal(ws[0], w) for w in ws)
    seeds_differ = m.engine_seed_probe != 0
    q.put((rank, bool(ok), bool(same), m.engine_seed_probe))
    dist.destroy_process_group()


def test_stock_ddp_wrapper_averages_gradients_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok and same for _, ok, same, _ in res), res
    assert len({s for *_, s in res}) == world, "dropout seeds must differ across ranks"
