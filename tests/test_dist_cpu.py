"""World-size-2 gloo test (CPU) of the N > 1 host logic: read sharding and max-over-ranks timing."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nanocall_b200 import dist as ncd


def test_shard_bounds_cover_and_balance():
    rng = np.random.default_rng(0)
    lengths = rng.integers(100, 20000, 1000)
    for world in (1, 2, 3, 4, 8):
        b = ncd.shard_bounds(lengths, world)
        assert b[0] == 0 and b[-1] == lengths.size and all(x <= y for x, y in zip(b, b[1:]))
        per = [int(lengths[b[r]:b[r + 1]].sum()) for r in range(world)]
        assert sum(per) == int(lengths.sum())
        assert max(per) - min(per) <= 2 * lengths.max()
    assert ncd.shard_bounds([5], 4) == [0, 1]
    assert ncd.shard_bounds([5, 5, 5], 8)[-1] == 3


def _worker(rank, world, port, lengths, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = ncd.shard_bounds(lengths, world)
    mine = lengths[b[rank]:b[rank + 1]]
    events = int(mine.sum())
    seconds = 1.0 + rank  # rank 1 is the slow one
    dist.barrier()
    rate, total, slowest = ncd.whole_job_rate(events, seconds)
    mx = ncd.max_over_ranks([seconds, float(events)])
    if rank == 0:
        torch.save(dict(rate=rate, total=total, slowest=slowest, mx=mx, events=events), out)
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    lengths = np.arange(1, 41) * 100
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, port, lengths, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["total"] == float(lengths.sum())
    assert r["slowest"] == 2.0 and r["rate"] == float(lengths.sum()) / 2.0
    assert r["mx"][0] == 2.0 and r["mx"][1] >= r["events"]
