"""World-size-2 gloo test of the N>1 host logic: batch sharding + the in-place all-gather of embeddings."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clearcam_b200.parallel import gather_rows, max_over_ranks, shard_range


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, D, per = 10, 8, 5
        lo, hi = shard_range(n, world, rank)
        full = torch.zeros(world * per, D)
        # "kernel output": this rank's rows, already in place inside the gather buffer
        full[rank * per:(rank + 1) * per] = torch.arange(lo, hi, dtype=torch.float32).reshape(-1, 1).expand(-1, D)
        gather_rows(full, per)
        ok = torch.equal(full[:, 0], torch.arange(n, dtype=torch.float32))
        t = max_over_ranks(float(rank + 1), "cpu")
        q.put((rank, bool(ok), t))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True, 2.0), (1, True, 2.0)]
