"""CPU tests of the multi-rank host logic (gloo, world_size 2): region sharding (LPT), the hand-over of the
communicator id that precedes a sharded search, and the ownership arithmetic of the per-level score exchange.  The
sharded search itself is covered by the GPU tests (ranks emulated on one device, tests/test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nahrwhals_b200.distributed import broadcast_id, estimate_region_cost, own_chunk, own_range, shard_regions
from nahrwhals_b200.sdgrid import sdgrid, to_solver_inputs


def test_shard_regions_lpt():
    costs = [5, 9, 1, 7, 3, 3, 8]
    parts = shard_regions(costs, 3)
    assert sorted(k for p in parts for k in p) == list(range(len(costs)))
    loads = [sum(costs[k] for k in p) for p in parts]
    assert max(loads) - min(loads) <= max(costs)
    assert shard_regions(costs, 3) == parts  # deterministic
    assert shard_regions(costs, 1) == [list(range(len(costs)))]
    assert shard_regions([], 2) == [[], []]


def test_region_cost_monotone():
    small = to_solver_inputs(sdgrid(20, [2, 2], seed=1)[0])
    big = to_solver_inputs(sdgrid(60, [5, 4, 4], seed=1)[0])
    assert estimate_region_cost(big, 3) > estimate_region_cost(small, 3) > 0
    assert estimate_region_cost(big, 3, maxwidth=10) < estimate_region_cost(big, 3)


def test_own_ranges_partition_every_level():
    """Host mirror of own_chunk / owns (csrc/nw_kernels.cuh): the ranks' ranges are disjoint, cover every fresh
    candidate, and the exchanged segments are 16-byte aligned."""
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 3, 4, 5, 127, 128, 1000, 65537, 10722271):
            chunk = own_chunk(n, world)
            assert chunk % 4 == 0 and chunk >= 4 and chunk * world >= n
            seen = 0
            for r in range(world):
                lo, hi = own_range(n, world, r)
                assert lo == min(r * chunk, n) and hi == min((r + 1) * chunk, n) and lo <= hi
                seen += hi - lo
            assert seen == n


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ident = bytes((7 * k + 3) % 256 for k in range(128)) if rank == 0 else None
        got = broadcast_id(dist, ident, 0)
        ok = got == bytes((7 * k + 3) % 256 for k in range(128))
        # every rank derives the same shard of a regionfile without talking to the others
        costs = [float((k * 37) % 11 + 1) for k in range(40)]
        mine = shard_regions(costs, world)[rank]
        allp = [None] * world
        dist.all_gather_object(allp, mine)
        ok = ok and sorted(k for p in allp for k in p) == list(range(40)) and allp == shard_regions(costs, world)
        # the source rank must supply a full id
        if rank == 0:
            try:
                broadcast_id(dist, b"short", 0)
                ok = False
            except ValueError:
                pass
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_id_broadcast_and_sharding_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
