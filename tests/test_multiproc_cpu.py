"""N > 1 host logic on CPU: world_size-2/3 gloo process groups exercise the instance partitioning, the per-instance
seed assignment, the final result gather and the max-over-ranks timing reduction that bench.py uses under torchrun."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hydrochrono_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard.shard_range(total, world, rank)
        seeds = shard.instance_seeds(lo, hi)
        # stand-in for the per-instance result of this rank's block: a function of the global instance index
        local = np.stack([seeds.astype(np.float64), 2.0 * np.arange(lo, hi)], axis=1)
        full = shard.gather_results(local, total, world, rank, dist=dist)
        tmax = shard.max_over_ranks(0.1 * (rank + 1), world, dist=dist)
        dist.barrier()
        q.put((rank, lo, hi, seeds.tolist(), None if full is None else full.tolist(), tmax))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 16), (3, 10)])
def test_partition_gather_and_timing_reduce(world, total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = []
    for rank, lo, hi, seeds, full, tmax in res:
        covered += list(range(lo, hi))
        assert seeds == [1 + i for i in range(lo, hi)]
        assert abs(tmax - 0.1 * world) < 1e-12            # every rank sees the max
        if rank == 0:
            full = np.array(full)
            assert full.shape == (total, 2)
            np.testing.assert_array_equal(full[:, 0], 1 + np.arange(total))   # global instance order
            np.testing.assert_array_equal(full[:, 1], 2.0 * np.arange(total))
        else:
            assert full is None
    assert covered == list(range(total))                  # disjoint, complete, contiguous


def test_shard_range_properties():
    for total in (1, 7, 16384, 16385):
        for world in (1, 2, 4, 8):
            blocks = [shard.shard_range(total, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert shard.shard_range(16384, 8, 3) == (6144, 8192)
