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
