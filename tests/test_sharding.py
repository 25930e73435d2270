"""Host-side logic of the N>1 path on CPU: world_size-2 (and 3) gloo process groups exercise the stream
partition and the result gather used by bench.py under torchrun."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from soundscope_b200.sharding import gather_results, shard_range, shard_sizes


def test_shard_range_partitions_exactly():
    for n in (1, 2, 7, 4096, 1000000, 125001):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(n, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (a, b), (c, d) in zip(ranges, ranges[1:]):
                assert b == c and a <= b
            sizes = shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_streams, stride, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_streams, world, rank)
        # row s holds s + column/100: any misordering or padding leak shows up in the gathered matrix
        rows = torch.arange(lo, hi, dtype=torch.float64)[:, None] + torch.arange(stride, dtype=torch.float64)[None, :] / 100.0
        got = gather_results(rows, n_streams)
        want = torch.arange(n_streams, dtype=torch.float64)[:, None] + torch.arange(stride, dtype=torch.float64)[None, :] / 100.0
        ok = got.shape == want.shape and bool(torch.equal(got, want))
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and t.item() == float(world)
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_streams", [(2, 4096), (2, 7), (3, 1000)])
def test_gather_results_gloo(world, n_streams):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, 8, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)
