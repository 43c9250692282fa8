"""CPU, world_size 2 over gloo: the multi-GPU path shards independent images across ranks (no data-path collective)
and gathers the results once."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hedit_b200.dist import gather_results, max_over_ranks, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 64, 65):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_items, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32).reshape(-1, 1, 1).expand(-1, 2, 3).contiguous() * 10 + rank * 0
    full = gather_results(local, n_items)
    t = max_over_ranks(float(rank + 1))
    if rank == 0:
        q.put((full[:, 0, 0].tolist(), t))
    dist.destroy_process_group()


def test_gather_results_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 5
    procs = [ctx.Process(target=_worker, args=(r, 2, 29613, n, q)) for r in range(2)]
    [p.start() for p in procs]
    vals, t = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert vals == [float(i * 10) for i in range(n)] and t == 2.0
