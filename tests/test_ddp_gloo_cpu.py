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
