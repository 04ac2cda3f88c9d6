"""N>1 host logic on CPU: two gloo ranks shard a batch, "compute" their shard, and gather it back in order."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pytorch_attention_b200.sharding import gather_outputs, shard_batch, shard_range


@pytest.mark.parametrize("batch,world", [(64, 8), (512, 8), (7, 2), (5, 4), (2, 4)])
def test_shard_ranges_partition_the_batch(batch, world):
    spans = [shard_range(batch, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == batch
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = torch.arange(batch * 3, dtype=torch.float32).reshape(batch, 3)
    local = shard_batch(x, rank, world) * 2.0          # stand-in for the per-rank forward
    full = gather_outputs(local, batch)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)            # the max-over-ranks timing reduction bench.py uses
    q.put((rank, torch.equal(full, x * 2.0), t.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7])
def test_two_rank_gloo_shard_and_gather(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + batch
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    assert all(mx == 2.0 for _, _, mx in res)
