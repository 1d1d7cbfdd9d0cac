#!/usr/bin/env python
"""Frontier-sharded deep search over the ranks of a torchrun launch (NCCL); checks rank results against the
single-GPU search and prints timings.  usage: torchrun --nproc-per-node N tools/dist_search.py [workload] [reps]"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nahrwhals_b200 as nb  # noqa: E402
from bench import make_workload  # noqa: E402
from nahrwhals_b200.distributed import solve_sharded  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "s150"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
aln, lx, ly, flags, desc = make_workload(name, 1)
s = nb.Solver(local)
ref = s.solve(aln, lx, ly, **flags)
for r in range(reps):
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    R = solve_sharded(s, aln, lx, ly, **flags)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ok = R.tsv == ref.tsv and R.scored == ref.scored and R.generated == ref.generated
    t1 = time.perf_counter()
    s.solve(aln, lx, ly, **flags)
    d1 = time.perf_counter() - t1
    print("[rank %d/%d] rep %d sharded %.2f ms (device %.2f ms) vs single-GPU %.2f ms; identical=%s; scored %s; "
          "cells on this rank %s" % (rank, world, r, dt * 1e3, R.stats.total_ms, d1 * 1e3, ok, R.scored, R.cells), flush=True)
    assert ok
dist.barrier()
dist.destroy_process_group()
s.close()
