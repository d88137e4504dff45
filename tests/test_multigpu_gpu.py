"""Two-GPU hardware tests of the data-parallel training path (skipped on boxes with one GPU; run with
`gpurun --gpus 2 -- python -m pytest tests/test_multigpu_gpu.py -m gpu`):

  * FusedTrainer: 2 ranks x 128 spectra reproduce 1 rank x 256 spectra - same loss, same Adam first moment (which is
    (1 - beta1) x the clipped, world-averaged gradient: linear in what the bucketed NCCL all-reduce produced, captured
    in the step graph with the 1/world factor folded into the Adam kernel), same weights after the step;
  * the Lightning-style path under a STOCK torch DistributedDataParallel wrapper (reference strategy
    "ddp_find_unused_parameters_true", trainer/trainer.py:58): gradients averaged over the ranks, replicas identical.
"""
import os
import socket
import time

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model(precision, bench, dropout=0.0):
    from multimodalanalytical_b200.wrapper import HFWrapper
    c = bench.C2
    return HFWrapper(data_config=bench.data_config(c), target_tokenizer=bench.Tok(c["V"]), num_steps=100,
                     precision=precision, seed=3, **bench.model_kwargs(c, dropout=dropout))


def _worker(rank, world, port, q, mode):
    import torch.distributed as dist

    import bench
    from multimodalanalytical_b200.trainer import FusedTrainer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    full = bench.synth_batch(bench.C2, 256, 21)
    mine = bench.map_batch(full, lambda x: x[:, rank * 128:(rank + 1) * 128].contiguous())
    out = {}
    if mode == "fused":
        m = _model("fp32", bench)
        tr = FusedTrainer(m, clip_grad=1.0)
        for i in range(2):  # step 0 eager, step 1 from the captured graph (NCCL inside the graph)
            loss = tr.train_step(mine, i)
        torch.cuda.synchronize()
        out = {"loss": float(loss), "m": m.store.m.cpu(), "p": m.store.p.cpu()}
    elif mode == "p2p_vs_nccl":
        for ddp_mode in ("nccl", "p2p"):  # both in one process pair: process start-up dominates the test's cost
            os.environ["MMA_DDP"] = ddp_mode
            m = _model("bf16", bench)
            tr = FusedTrainer(m, clip_grad=1.0)
            assert (tr.peer is not None) == (ddp_mode == "p2p")
            for i in range(3):  # eager step, graph capture, one replay (device-side barriers inside the graph)
                loss = tr.train_step(mine, i)
                if i == 0:
                    torch.cuda.synchronize()
                    m.store.gather_master()
                    m1 = m.store.m.cpu()
            torch.cuda.synchronize()
            sd = m.state_dict()  # p2p: reassembles the rank-sharded master weights
            out[ddp_mode] = {"loss": float(loss), "m1": m1, "m": m.store.m.cpu(), "p": m.store.p.cpu(),
                             "pb": m.store.pb.float().cpu(), "w": sd["hf_model.token_ff.weight"].float().cpu()}
    else:
        m = _model("bf16", bench)
        ddp = torch.nn.parallel.DistributedDataParallel(m, device_ids=[rank], find_unused_parameters=True)
        (opt,), _ = m.configure_optimizers()
        m.train()
        opt.zero_grad()
        res = ddp(mine)
        res.loss.backward()
        torch.cuda.synchronize()
        g = torch.cat([v.reshape(-1).float() for v in m.named_gradients().values()])
        opt.step()
        ws = [torch.zeros_like(m.store.p) for _ in range(world)]
        dist.all_gather(ws, m.store.p)
        out = {"loss": float(res.loss), "g": g.cpu(), "same": all(torch.equal(ws[0], w) for w in ws)}
    q.put((rank, out))
    dist.barrier()
    torch.cuda.synchronize()
    time.sleep(2.0)  # let the parent drain the queue
    os._exit(0)  # captured graphs hold NCCL / symmetric-memory work: leave without the teardown (as bench.py does)


def _spawn(mode, world=2):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
    return res


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_ranks_of_128_equal_one_rank_of_256_fused_trainer():
    import bench
    from multimodalanalytical_b200.trainer import FusedTrainer
    res = _spawn("fused")
    torch.cuda.set_device(0)
    m = _model("fp32", bench)
    tr = FusedTrainer(m, clip_grad=1.0)
    full = bench.synth_batch(bench.C2, 256, 21)
    for i in range(2):
        loss = tr.train_step(full, i)
    torch.cuda.synchronize()
    # the ranks hold identical replicas
    assert torch.equal(res[0]["p"], res[1]["p"]) and torch.equal(res[0]["m"], res[1]["m"])
    # mean of the two half-batch losses == full-batch loss (no padding: equal token counts)
    assert abs(0.5 * (res[0]["loss"] + res[1]["loss"]) - float(loss)) < 1e-5 * float(loss)
    assert _rel(res[0]["m"], m.store.m.cpu()) < 2e-5, "world-averaged gradient differs from the single-rank gradient"
    d = (res[0]["p"] - m.store.p.cpu()).abs()
    lr0 = m.lr / 25
    assert float((d > 0.2 * lr0).float().mean()) < 2e-3


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_stock_ddp_wrapper_on_two_gpus():
    import bench
    res = _spawn("ddp")
    torch.cuda.set_device(0)
    m = _model("bf16", bench)
    m.train()
    full = bench.synth_batch(bench.C2, 256, 21)
    out = m.forward(full)
    out.loss.backward()
    torch.cuda.synchronize()
    g = torch.cat([v.reshape(-1).float() for v in m.named_gradients().values()]).cpu()
    assert res[0]["same"] and res[1]["same"]
    assert torch.equal(res[0]["g"], res[1]["g"]), "DDP did not leave identical gradients on the ranks"
    assert abs(0.5 * (res[0]["loss"] + res[1]["loss"]) - float(out.loss)) < 2e-3 * float(out.loss)
    assert _rel(res[0]["g"], g) < 3e-2, "averaged half-batch gradients differ from the full-batch gradient (bf16)"


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_peer_memory_sharded_step_equals_nccl_allreduce_step():
    """MMA_DDP=p2p (reduce-scatter by P2P loads + rank-sharded Adam + bf16 weights stored into every peer's mirror, device
    barriers captured in the step graph) against MMA_DDP=nccl (bucketed all-reduce + replicated Adam) on the same two
    ranks and batches: same loss trajectory, same Adam moments and weights after three steps."""
    res = _spawn("p2p_vs_nccl")
    a, b = {r: res[r]["p2p"] for r in (0, 1)}, {r: res[r]["nccl"] for r in (0, 1)}
    for r in (0, 1):
        assert abs(a[r]["loss"] - b[r]["loss"]) < 1e-3 * abs(b[r]["loss"]), (a[r]["loss"], b[r]["loss"])
    assert torch.equal(a[0]["pb"], a[1]["pb"]), "bf16 mirrors differ between the ranks"
    assert torch.equal(a[0]["p"], a[1]["p"]), "gathered master weights differ between the ranks"
    # after ONE step the Adam first moment is (1 - beta1) x the clipped world-averaged gradient: the two exchanges must
    # agree to fp32 rounding; later steps drift apart through the bf16 rounding of the weights
    assert _rel(a[0]["m1"], b[0]["m1"]) < 2e-5
    assert _rel(a[0]["m"], b[0]["m"]) < 2e-2
    assert torch.equal(a[0]["w"], a[1]["w"]) and _rel(a[0]["w"], b[0]["w"]) < 1e-3  # state_dict(): gathered master
    d = (a[0]["pb"] - b[0]["pb"]).abs()
    assert float((d > 0).float().mean()) < 0.05, "more than 5 % of the bf16 weights differ from the NCCL path"
    assert float(d.max()) < 2e-2
