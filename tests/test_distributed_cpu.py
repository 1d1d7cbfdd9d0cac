"""CPU tests of the multi-rank host logic (gloo, world_size 2): region sharding and the padded variable-length
all-gather used by the sharded search.  The CUDA merge itself is covered by the GPU tests (emulated ranks)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nahrwhals_b200.distributed import all_gather_varlen, estimate_region_cost, pad_gather_layout, shard_regions
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


def test_pad_layout():
    assert pad_gather_layout([3, 0, 5]) == (5, [0, 5, 10])
    assert pad_gather_layout([0, 0]) == (0, [0, 0])


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        item = 40
        count = 3 + 2 * rank  # ragged: rank 0 sends 3 records, rank 1 sends 5
        local = torch.arange(count * item, dtype=torch.int64).to(torch.uint8) + rank * 100
        out, counts, stride = all_gather_varlen(dist, local, count, item)
        ok = counts == [3, 5] and stride == 5 and out.numel() == world * stride * item
        for r in range(world):
            exp = (torch.arange(counts[r] * item, dtype=torch.int64).to(torch.uint8) + r * 100)
            ok = ok and torch.equal(out[r * stride * item: r * stride * item + counts[r] * item], exp)
        # empty contribution from every rank
        out2, counts2, stride2 = all_gather_varlen(dist, torch.zeros(0, dtype=torch.uint8), 0, item)
        ok = ok and counts2 == [0, 0] and stride2 == 1
        # needed-flags reduction used between marking passes
        need = torch.zeros(8, dtype=torch.uint8)
        need[rank] = rank + 1
        dist.all_reduce(need, op=dist.ReduceOp.MAX)
        ok = ok and need.tolist() == [1, 2, 0, 0, 0, 0, 0, 0]
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_varlen_allgather_gloo_world2():
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
