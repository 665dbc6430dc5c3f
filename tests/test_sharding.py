"""Multi-GPU host logic on CPU: barcode-group sharding and the max-over-ranks timing, world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lariat_b200 import sharding


def test_shard_groups_cover_and_balance():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 4, 8):
        sizes = rng.integers(1, 300, size=1000)
        sh = sharding.shard_groups(sizes, world)
        assert len(sh) == world and sh[0][0] == 0 and sh[-1][1] == len(sizes)
        assert all(sh[i][1] == sh[i + 1][0] for i in range(world - 1))
        loads = [int(sizes[a:b].sum()) for a, b in sh]
        assert max(loads) - min(loads) <= 2 * int(sizes.max())
    # degenerate: fewer groups than ranks, empty input, one huge group
    assert sharding.shard_groups([5], 4)[-1][1] == 1
    assert sharding.shard_groups([], 2) == [(0, 0), (0, 0)]
    sh = sharding.shard_groups([1, 1, 1000, 1, 1], 2)
    assert sh[0][1] == sh[1][0] and sh[1][1] == 5
    assert sharding.pair_range([3, 4, 5], 1, 3) == (3, 12)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes = np.random.default_rng(7).integers(50, 150, size=101)          # same on every rank, like the barcode-sorted input
    lo, hi = sharding.shard_groups(sizes, world)[rank]
    ranges = sharding.gather_ranges(lo, hi)
    slowest = sharding.max_over_ranks(1.0 + rank)
    total = sharding.sum_over_ranks(float(sizes[lo:hi].sum()))
    dist.barrier()
    q.put((rank, ranges, slowest, total, int(sizes.sum())))
    dist.destroy_process_group()


def test_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ranges, slowest, total, want in out:
        assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == 101      # disjoint cover, rank order
        assert slowest == 2.0                                                                 # max over ranks, not the local time
        assert total == want                                                                  # every pair owned exactly once
