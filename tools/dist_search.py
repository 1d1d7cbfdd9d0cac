#!/usr/bin/env python
"""Frontier-sharded deep search over the ranks of a torchrun launch: NCCL all-gather per level issued by libnwsolve
itself (nw_solve_sharded); checks every rank's result against the single-GPU search and prints timings.
usage: torchrun --nproc-per-node N tools/dist_search.py [workload] [reps]"""
import json
import os
import sys
import time

# stdout = the JSON line only: native writes to descriptor 1 (NCCL's version banner) are pointed at stderr
sys.stdout.flush()
_REAL_STDOUT_FD = os.dup(1)
os.dup2(2, 1)
sys.stdout = os.fdopen(_REAL_STDOUT_FD, "w", buffering=1)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nahrwhals_b200 as nb  # noqa: E402
from bench import make_workload  # noqa: E402
from nahrwhals_b200.distributed import init_sharded, solve_sharded  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "s150"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
aln, lx, ly, flags, desc = make_workload(name, 1)
s = nb.Solver(local)
rank, world = init_sharded(s)
ref = s.solve(aln, lx, ly, keep_ids=True, **flags)
R = solve_sharded(s, aln, lx, ly, keep_ids=True, **flags)
same = (R.tsv == ref.tsv and R.scored == ref.scored and R.generated == ref.generated and R.cells == ref.cells and
        all((R.ids[k] == ref.ids[k]).all() for k in ref.ids) and all((R.edges[k] == ref.edges[k]).all() for k in ref.edges))
assert same, "rank %d: sharded result differs from the single-GPU search" % rank
for _ in range(3):
    solve_sharded(s, aln, lx, ly, **flags)
    s.solve(aln, lx, ly, **flags)
t_sh, t_one, dp_sh, dp_one = [], [], [], []
for r in range(reps):
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    R = solve_sharded(s, aln, lx, ly, **flags)
    t_sh.append((time.perf_counter() - t0) * 1e3)
    dp_sh.append(R.stats.score_ms)
    dev_sh = R.stats.total_ms
    torch.cuda.synchronize()
    dist.barrier()
    t1 = time.perf_counter()
    Q = s.solve(aln, lx, ly, **flags)
    t_one.append((time.perf_counter() - t1) * 1e3)
    dp_one.append(Q.stats.score_ms)
    dev_one = Q.stats.total_ms
v = torch.tensor([min(t_sh), min(t_one), min(dp_sh), min(dp_one), dev_sh, dev_one], dtype=torch.float64, device="cuda")
dist.all_reduce(v, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps(dict(workload=name, world=world, identical_to_single_gpu=True, scored=sum(ref.scored),
                          sharded_ms_e2e=v[0].item(), one_gpu_ms_e2e=v[1].item(), sharded_dp_ms=v[2].item(),
                          one_gpu_dp_ms=v[3].item(), sharded_ms_device=v[4].item(), one_gpu_ms_device=v[5].item(),
                          note="best of %d, max over ranks; e2e = wall clock around the C-ABI call from host buffers" % reps)),
          flush=True)
dist.barrier()
dist.destroy_process_group()
s.close()
