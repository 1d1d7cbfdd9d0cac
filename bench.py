#!/usr/bin/env python
"""bench.py -- headline benchmark of the mutation-space solver (BASELINE.json metric: candidates scored/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload s64|s150]

One "step" = one complete exhaustive depth-3 search (BFS + dedup + banded DP + report) of a synthetic 64-block
signed gridmatrix (BASELINE.json configs[1]).  With N > 1 (launched under torchrun, one rank per GPU) every rank
solves its own region of the same shape (independent regions shard per GPU with no collective -- SURVEY 8e,
regionfile mode), so scaling is weak.

Prints ONE JSON line.  `value` is whole-job unique candidates scored per second, timed on the device with CUDA
events inside the library (inputs already resident in HBM); `e2e` is the same metric through the public API from
host numpy buffers (host->device copies of the inputs and device->host copies of the results inside the timed
region).  `--impl reference` times the CPU restatement of the reference (oracle/) on the host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE line (the JSON).  Native libraries write to file descriptor 1 behind Python's back (NCCL
# prints its version banner there, for torch's communicator and for the one libnwsolve opens itself, whatever
# NCCL_DEBUG_FILE says): descriptor 1 is pointed at stderr for the whole run and Python's own sys.stdout keeps the
# real one.
os.environ.setdefault("NCCL_DEBUG", "WARN")
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
if __name__ == "__main__":  # not when a tool imports make_workload from here
    sys.stdout.flush()
    _REAL_STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_REAL_STDOUT_FD, "w", buffering=1)

METRIC = "mutation_candidates_scored_per_s"
UNIT = "candidates/s"

WORKLOADS = {
    # name: (sdgrid args, solver flags, description)
    "s64": (dict(n_x=64, families=[10, 9, 8, 8, 7], chain_depth=3),
            dict(maxdepth=3, maxdup=2, maxwidth=None, minreport=0.98, maxreport=1000),
            "synthetic 64-block signed gridmatrix (sdgrid 64 blocks, families 10/9/8/8/7 = 158 repeat pairs), "
            "exhaustive depth-3 inv/del/dup search, -d 3 -u 2 -r 0.98 -R 1000"),
    "s150": (dict(n_x=150, families=[8, 7, 6, 6, 5, 5, 4], chain_depth=4, len_tol=0.12),
             dict(maxdepth=4, maxdup=2, maxwidth=1000, minreport=0.98, maxreport=1000),
             "synthetic 150-block gridmatrix (105 repeat pairs), depth-4 search, -d 4 -u 2 -w 1000 -r 0.98 -R 1000"),
    "s150d3": (dict(n_x=150, families=[8, 7, 6, 6, 5, 5, 4], chain_depth=3, len_tol=0.12),
               dict(maxdepth=3, maxdup=2, maxwidth=None, minreport=0.98, maxreport=1000),
               "synthetic 150-block gridmatrix (105 repeat pairs), exhaustive depth-3 search, -d 3 -u 2 -r 0.98 -R 1000"),
}


def make_workload(name, seed):
    from nahrwhals_b200.sdgrid import sdgrid, to_solver_inputs
    gen, flags, desc = WORKLOADS[name]
    mat, lx, ly, _ = sdgrid(seed=seed, **gen)
    return to_solver_inputs(mat), lx, ly, flags, desc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = str(gpu_index)
        self.lines = []
        self.p = None

    def start(self):
        if int(os.environ.get("RANK", "0")) != 0:  # one sampler per job: nvidia-smi polling competes with the ranks
            return
        if os.environ.get("NW_BENCH_NO_SAMPLER"):  # diagnostics only: a line without clocks is not a valid bench line
            return
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", self.idx, "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.lines.append(ln)

    def stop(self):
        if not self.p:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["not sampled on this rank / nvidia-smi unavailable"])
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8 or f[0] != self.idx:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


def cpu_sample(aln, lx, ly, flags, max_scored, threads):
    """Oracle (CPU restatement of solver.jl) on a bounded sample of the workload: the first `max_scored` unique
    candidates of the same search, one independent replica per host thread (what R's threads= does for regions).
    Returns (candidates/s over all threads, seconds, candidates scored, DP cells evaluated)."""
    from oracle import oracle_lib
    oracle_lib.lib()
    kw = dict(maxdepth=flags["maxdepth"], maxdup=flags["maxdup"], maxwidth=flags["maxwidth"],
              minreport=flags["minreport"], maxreport=flags["maxreport"], max_scored=max_scored)
    res = [None] * threads

    def work(k):
        res[k] = oracle_lib.solve(aln, lx, ly, want_trajs=False, full=False, **kw)

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    scored = sum(int(r.scored.sum()) for r in res)
    cells = sum(int(r.cells.sum()) for r in res)
    return scored / dt, dt, scored, cells


def full_search_cells_note(workload):
    """DP cells per scored candidate of the WHOLE search (oracle fixture of this workload), to put beside the sample's."""
    try:
        fx = json.load(open(os.path.join(ROOT, "tests", "golden", "fullsize_%s.json" % workload)))
        return " vs %.0f over the whole search (oracle, tests/golden/fullsize_%s.json)" % (fx["cells_per_scored"], workload)
    except Exception:
        return ""


def run_reference(args, rank, world):
    if rank != 0:
        return
    aln, lx, ly, flags, desc = make_workload(args.workload, 1)
    cores = os.cpu_count() or 1
    sample_n = args.cpu_sample
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_sample(aln, lx, ly, flags, max(1000, sample_n // 20), cores)
    tot_scored, tot_t, tot_cells = 0, 0.0, 0
    for _ in range(args.steps):
        v, dt, sc, ce = cpu_sample(aln, lx, ly, flags, sample_n, cores)
        tot_scored += sc
        tot_t += dt
        tot_cells += ce
    value = tot_scored / tot_t
    sample = ("first %d unique candidates of the same search per replica, %d independent replicas (one per host "
              "thread), C++ restatement of solver.jl (oracle/nw_oracle.cpp, -O2); %.0f DP cells per scored candidate in "
              "the sample%s" % (sample_n, cores, tot_cells / max(tot_scored, 1), full_search_cells_note(args.workload)))
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * tot_t / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=desc, l2="n/a (CPU)"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


def run_regions(args, rank, world, local):
    """BASELINE.json config 4: regionfile batch.  Every rank solves its own `--regions` synthetic SD-flanked regions
    (20-120 blocks) through nw_solve_many_ctxs: `--regions-per-pass` regions are searched TOGETHER by one pipeline
    instance (every launch of a BFS level covers all of them) and the passes are handed out to `--contexts` contexts
    so that the host side of one pass overlaps the device side of another.  No collective on the data path.
    Metric: regions/s (wall clock around the batch, host buffers in, reports out, max over ranks; several streams run
    concurrently so there is no single device timeline).  --impl reference runs the CPU restatement on all host
    threads over a 10 x smaller sample."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from region_batch import FLAGS, make_regions, run_cpu
    cores = os.cpu_count() or 1
    if args.impl == "reference":
        if rank != 0:
            return
        regs = make_regions(max(50, args.regions // 10), seed=0)
        r, sc, dt = run_cpu(regs, cores)
        line = dict(impl="reference", metric="regions_solved_per_s", value=r, unit="regions/s", n_gpus=args.gpus,
                    steps=1, warmup=0, ms_per_step=1e3 * dt, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f64", data="synthetic",
                    config=dict(workload="regionfile batch: %d synthetic SD-flanked regions, 20-120 blocks, "
                                         "-d 3 -u 2 -w 1000 -r 0.98 -R 1000" % len(regs)),
                    cpu_baseline=dict(value=r, unit="regions/s", cores=cores, kind="port",
                                      sample="%d regions, one per host thread at a time" % len(regs)),
                    e2e=dict(value=r, unit="regions/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        print(json.dumps(line), flush=True)
        return
    import torch
    import nahrwhals_b200 as nb
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    regs = make_regions(args.regions, seed=rank)
    # contexts per GPU: each one is a host thread issuing launches; do not oversubscribe the host cores
    nctx = max(1, min(args.contexts, (os.cpu_count() or 8) // max(world, 1)))
    per = max(1, args.regions_per_pass)
    solvers = [nb.Solver(local) for _ in range(nctx)]
    # warm-up: one untimed run over an independent set of regions of the same generator (memory pools of every
    # context, corridor tables of the shapes a long regionfile run keeps meeting); the timed run sees new regions
    warm = make_regions(args.regions, seed=1000 + rank)
    for _ in range(max(1, args.warmup_regions)):
        nb.solve_many(solvers, warm, regions_per_pass=per, **FLAGS)
    del warm
    sampler = ClockSampler(local)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sampler.start()
    t0 = time.perf_counter()
    steps = max(1, args.steps)  # a step = one pass over the whole region set
    dt = 0.0
    for _ in range(steps):
        tm = {}
        res = nb.solve_many(solvers, regs, regions_per_pass=per, timing=tm, **FLAGS)
        dt += tm["native_s"]           # the C-ABI call: host buffers in, reports in host memory out
    torch.cuda.synchronize()
    dt_py = (time.perf_counter() - t0) / steps  # incl. ctypes marshalling and numpy views of the reports
    dt /= steps
    clocks = sampler.stop()
    scored = sum(sum(R.scored) for R in res)
    firsts = res[::per]  # launch / byte counters are per pass (shared by the regions of the pass)
    launches = sum(R.stats.kernel_launches for R in firsts)
    h2d = sum(R.stats.h2d_bytes for R in firsts)
    d2h = sum(R.stats.d2h_bytes for R in firsts)
    n = len(regs)
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        s = torch.tensor([n, scored, launches, h2d, d2h], dtype=torch.float64, device="cuda")
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        n, scored, launches, h2d, d2h = s.tolist()
    line = dict(metric="regions_solved_per_s", value=n / dt, unit="regions/s", n_gpus=world, steps=steps, warmup=max(1, args.warmup_regions),
                ms_per_step=1e3 * dt, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="int32",
                data="synthetic",
                config=dict(workload="regionfile batch: %d synthetic SD-flanked regions per GPU (20-120 blocks, 2-6 "
                                     "repeat families), -d 3 -u 2 -w 1000 -r 0.98 -R 1000; %d regions per pass "
                                     "searched together, %d contexts per GPU; regions shard per GPU, no collective"
                                     % (args.regions, per, nctx),
                            candidates_scored_per_s=scored / dt, l2="not flushed (every region is new data)",
                            timed="wall clock around nw_solve_many_ctxs (host buffers in, H2D, search, D2H, report "
                                  "text in host memory); through the Python mirror incl. ctypes marshalling: "
                                  "%.0f regions/s" % (len(regs) * world / dt_py)),
                clocks=clocks,
                e2e=dict(value=n / dt, unit="regions/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                gpu_launches=int(launches))
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = make_regions(max(50, args.regions // 10), seed=0)
        r, sc, cdt = run_cpu(sample, cores)
        line["cpu_baseline"] = dict(value=r, unit="regions/s", cores=cores, kind="port",
                                    sample="%d regions of the same generator, %.1f s, all host threads" % (len(sample), cdt))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    for s_ in solvers:
        s_.close()


def run_stage(args, rank, world, local):
    """SURVEY 8f rows 1+3 around the solver: the whole solver STAGE of a region as R/wrapper_aln_and_analyse.R:176-191
    runs it (R depth-1 pre-pass, 99.8 gate, flip, solver, format_julia_output, merge, logfile summary) for a batch of
    regions through nw_region_analyse_many on one context per GPU.  Metric: regions/s, wall clock around the C-ABI
    call, host buffers in, result tables in host memory out.  --impl reference: the R restatement (Python) around the
    C++ oracle solver, one region per worker process."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from region_batch import genome_plan, make_genome_regions, make_stage_regions, run_stage_cpu
    cores = os.cpu_count() or 1
    genome = args.workload == "genome"
    if genome:
        # config 5: a fixed genome-wide window list, LPT-sharded over the GPUs (strong scaling), blacklist on the host
        from nahrwhals_b200.distributed import shard_regions
        plan = genome_plan(args.regions, seed=0)
        shards = shard_regions([c for _, _, c in plan], max(world, 1))
        desc = ("solver stage of whole-genome mode (synthetic T2T-sized pair: %%d candidate windows of 200 kb / 1 Mb / "
                "5 Mb -> 10-124 blocks per axis, %d left after the blacklist), R depth-1 pre-pass + solver -d 3 -u 2 "
                "-w 1000 -r 0.98 -R 1000 + format/merge/logfile summary" % len(plan))
    else:
        desc = ("solver stage of a regionfile batch (R depth-1 pre-pass + solver -d 3 -u 2 -w 1000 -r 0.98 -R 1000 + "
                "format/merge/logfile summary): %d synthetic SD-flanked regions per GPU, 20-120 blocks")
    if args.impl == "reference":
        if rank != 0:
            return
        if genome:
            step = max(1, len(plan) // max(32, args.regions // 40))
            regs = make_genome_regions(plan, list(range(0, len(plan), step)))
        else:
            regs = make_stage_regions(max(32, args.regions // 40), seed=0)
        r, _, dt = run_stage_cpu(regs, cores)
        print(json.dumps(dict(impl="reference", metric="regions_solved_per_s", value=r, unit="regions/s",
                              n_gpus=args.gpus, steps=1, warmup=0, ms_per_step=1e3 * dt, higher_is_better=True,
                              scaling="strong" if genome else "weak", vs_baseline=None, dtype="f64", data="synthetic",
                              config=dict(workload=desc % (args.regions if genome else len(regs)), sample_regions=len(regs)),
                              cpu_baseline=dict(value=r, unit="regions/s", cores=cores, kind="port",
                                                sample="%d regions, one per worker process" % len(regs)),
                              e2e=dict(value=r, unit="regions/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                              gpu_launches=0)), flush=True)
        return
    import torch
    import nahrwhals_b200 as nb
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if genome:
        regs = make_genome_regions(plan, shards[rank])
        warm = make_stage_regions(min(len(regs), 1500), seed=1000 + rank)
    else:
        regs = make_stage_regions(args.regions, seed=rank)
        warm = make_stage_regions(min(args.regions, 1500), seed=1000 + rank)
    # two contexts per GPU overlap the host phases of one batch with the device phases of another; more only
    # oversubscribe the host threads of the prepare / replay / format phases (measured: 8.4 k / 13.8 k / 8.2 k regions/s
    # with 1 / 2 / 4 contexts)
    nctx = max(1, min(args.contexts, 2, (os.cpu_count() or 8) // (2 * max(world, 1))))
    solvers = [nb.Solver(local) for _ in range(nctx)]
    solver = solvers if nctx > 1 else solvers[0]
    kw = dict(depth=3, maxdup=2, init_width=1000, minreport=0.98, regions_per_batch=512)
    for _ in range(max(1, args.warmup_regions)):
        nb.analyse_regions(solver, warm, **kw)
    del warm
    sampler = ClockSampler(local)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sampler.start()
    steps = max(1, args.steps)
    dt = 0.0
    for _ in range(steps):
        tm = {}
        res = nb.analyse_regions(solver, regs, timing=tm, **kw)
        dt += tm["native_s"]
    torch.cuda.synchronize()
    dt /= steps
    clocks = sampler.stop()
    n = len(regs)
    called = sum(1 for R in res if R.solver_called)
    launches = sum(R.gpu_launches for R in res)
    cands = sum(R.prepass_candidates for R in res)
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        s = torch.tensor([n, called, launches, cands], dtype=torch.float64, device="cuda")
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        n, called, launches, cands = s.tolist()
    line = dict(metric="regions_solved_per_s", value=n / dt, unit="regions/s", n_gpus=world, steps=steps,
                warmup=max(1, args.warmup_regions), ms_per_step=1e3 * dt, higher_is_better=True,
                scaling="strong" if genome else "weak", vs_baseline=None, dtype="int32", data="synthetic",
                config=dict(workload=(desc % args.regions) + ("; windows LPT-sharded over the GPUs by estimated work, "
                                                              "no collective on the data path" if genome else
                                                              "; %d context(s) per GPU, regions shard per GPU, no collective" % nctx),
                            solver_stage_seconds=dt,
                            solver_called=int(called), prepass_candidates=int(cands),
                            timed="wall clock around nw_region_analyse_many(_ctxs) (host buffers in, result tables out)"),
                clocks=clocks, e2e=dict(value=n / dt, unit="regions/s", h2d_bytes_per_step=None, d2h_bytes_per_step=None),
                gpu_launches=int(launches))
    if rank == 0 and world == 1 and not args.no_cpu:
        if genome:
            step = max(1, len(plan) // max(32, args.regions // 40))
            sample = make_genome_regions(plan, list(range(0, len(plan), step)))
        else:
            sample = make_stage_regions(max(32, args.regions // 40), seed=0)
        r, _, cdt = run_stage_cpu(sample, cores)
        line["cpu_baseline"] = dict(value=r, unit="regions/s", cores=cores, kind="port",
                                    sample="%d regions of the same generator, %.1f s, one region per worker process; R "
                                           "restatement in Python around the C++ oracle solver" % (len(sample), cdt))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    for s_ in solvers:
        s_.close()


def _grid_cpu_one(p):
    from oracle import nw_grid_oracle as go
    try:
        return go.paf_to_gridmatrix(*p)["mat"].shape
    except Exception:
        return None


def run_grid(args, rank, world, local):
    """the gridding step upstream of the solver (make_xy_grid_fast + bressiwrap +
    gridlist_to_gridmatrix) for a batch of regions through nw_grid_build.  Metric: regions/s, wall clock around the
    C-ABI call (host alignment arrays in, gridlines + gridmatrices in host memory out).  --impl reference: the R
    restatement (Python) on all host cores, one region per worker process."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from region_batch import make_stage_regions
    from nahrwhals_b200.sdgrid import matrix_to_alignments
    cores = os.cpu_count() or 1
    desc = "gridding of a regionfile batch: alignments of %d synthetic SD-flanked regions per GPU (20-120 blocks) -> gridlines -> gridmatrix"

    def cpu(pafs):
        import multiprocessing as mp
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_grid_cpu_one, pafs, chunksize=4)
        return len(pafs) / (time.perf_counter() - t0), time.perf_counter() - t0
    if args.impl == "reference":
        if rank != 0:
            return
        pafs = [matrix_to_alignments(m, lx, ly) for m, ly, lx, _ in make_stage_regions(max(64, args.regions // 20), seed=0)]
        r, dt = cpu(pafs)
        print(json.dumps(dict(impl="reference", metric="regions_gridded_per_s", value=r, unit="regions/s", n_gpus=args.gpus,
                              steps=1, warmup=0, ms_per_step=1e3 * dt, higher_is_better=True, scaling="weak",
                              vs_baseline=None, dtype="f64", data="synthetic", config=dict(workload=desc % len(pafs)),
                              cpu_baseline=dict(value=r, unit="regions/s", cores=cores, kind="port",
                                                sample="%d regions, one per worker process" % len(pafs)),
                              e2e=dict(value=r, unit="regions/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                              gpu_launches=0)), flush=True)
        return
    import torch
    import nahrwhals_b200 as nb
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pafs = [matrix_to_alignments(m, lx, ly) for m, ly, lx, _ in make_stage_regions(args.regions, seed=rank)]
    solver = nb.Solver(local)
    for _ in range(3):
        nb.build_grids(solver, pafs[:min(len(pafs), 1000)])
    sampler = ClockSampler(local)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sampler.start()
    steps = max(1, args.steps)
    dt = 0.0
    for _ in range(steps):
        tm = {}
        grids, rcs = nb.build_grids(solver, pafs, timing=tm)
        dt += tm["native_s"]
    torch.cuda.synchronize()
    dt /= steps
    clocks = sampler.stop()
    n, ok = len(pafs), sum(1 for r in rcs if r == 0)
    alns = sum(len(p[0]) for p in pafs)
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        s = torch.tensor([n, ok, alns], dtype=torch.float64, device="cuda")
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        n, ok, alns = s.tolist()
    line = dict(metric="regions_gridded_per_s", value=n / dt, unit="regions/s", n_gpus=world, steps=steps, warmup=3,
                ms_per_step=1e3 * dt, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=(desc % args.regions) + "; regions shard per GPU, no collective", grids_built=int(ok),
                            alignments=int(alns), timed="wall clock around nw_grid_build (host arrays in, grids out)"),
                clocks=clocks, e2e=dict(value=n / dt, unit="regions/s", h2d_bytes_per_step=int(alns) * 36,
                                        d2h_bytes_per_step=None), gpu_launches=3 * steps * max(world, 1))
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = pafs[:max(64, args.regions // 20)]
        r, cdt = cpu(sample)
        line["cpu_baseline"] = dict(value=r, unit="regions/s", cores=cores, kind="port",
                                    sample="%d regions of the same set, %.1f s, one region per worker process; R "
                                           "restatement in Python" % (len(sample), cdt))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    solver.close()


def _bitlocus_cpu_one(p):
    from oracle import nw_bitlocus_oracle as bo
    try:
        return bo.wrapper_paf_to_bitlocus(p, 1000.0, 1000.0)["status"]
    except Exception:
        return None


def run_bitlocus(args, rank, world, local):
    """SURVEY 8f rows 4 + 2 joined: wrapper_paf_to_bitlocus (PAF preparation, limits, gridding, SV-pair limit, coarser
    retries) for a regionfile batch through nw_bitlocus_build.  Metric: regions/s, wall clock around the C-ABI call
    (host PAF columns in, prepared tables + grids in host memory out).  --impl reference: the R restatement (Python) on
    all host cores, one region per worker process."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from region_batch import make_stage_regions
    from nahrwhals_b200.sdgrid import matrix_to_chunked_paf
    cores = os.cpu_count() or 1
    n_reg = max(16, args.regions // 4)
    desc = ("wrapper_paf_to_bitlocus of a regionfile batch: chunked (10 kb) alignment tables of %d synthetic SD-flanked "
            "regions per GPU (20-120 blocks) -> prepared PAF -> gridlines -> gridmatrix, minlen = compression = 1000, "
            "max_n_alns 150, max_size_col_plus_rows 250")

    def cpu(pafs):
        import multiprocessing as mp
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_bitlocus_cpu_one, pafs, chunksize=2)
        return len(pafs) / (time.perf_counter() - t0), time.perf_counter() - t0
    if args.impl == "reference":
        if rank != 0:
            return
        pafs = [matrix_to_chunked_paf(m, lx, ly) for m, ly, lx, _ in make_stage_regions(max(64, n_reg // 20), seed=0)]
        r, dt = cpu(pafs)
        print(json.dumps(dict(impl="reference", metric="regions_to_bitlocus_per_s", value=r, unit="regions/s",
                              n_gpus=args.gpus, steps=1, warmup=0, ms_per_step=1e3 * dt, higher_is_better=True,
                              scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                              config=dict(workload=desc % len(pafs)),
                              cpu_baseline=dict(value=r, unit="regions/s", cores=cores, kind="port",
                                                sample="%d regions, one per worker process" % len(pafs)),
                              e2e=dict(value=r, unit="regions/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                              gpu_launches=0)), flush=True)
        return
    import torch
    import nahrwhals_b200 as nb
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pafs = [matrix_to_chunked_paf(m, lx, ly) for m, ly, lx, _ in make_stage_regions(n_reg, seed=rank)]
    solver = nb.Solver(local)
    for k in range(3):  # steady state: the context's staging buffers have seen a batch of this size
        nb.paf_to_bitlocus(solver, pafs if k == 0 else pafs[:min(len(pafs), 200)], 1000.0, 1000.0)
    sampler = ClockSampler(local)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sampler.start()
    steps = max(1, args.steps)
    dt = 0.0
    for _ in range(steps):
        tm = {}
        res = nb.paf_to_bitlocus(solver, pafs, 1000.0, 1000.0, timing=tm)
        dt += tm["native_s"]
    torch.cuda.synchronize()
    dt /= steps
    clocks = sampler.stop()
    n, ok = len(pafs), sum(1 for r in res if r.status == "ok")
    alns = sum(len(p["tstart"]) for p in pafs)
    passes = sum(r.passes for r in res)
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        s = torch.tensor([n, ok, alns, passes], dtype=torch.float64, device="cuda")
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        n, ok, alns, passes = s.tolist()
    max_passes = max(r.passes for r in res)
    line = dict(metric="regions_to_bitlocus_per_s", value=n / dt, unit="regions/s", n_gpus=world, steps=steps, warmup=3,
                ms_per_step=1e3 * dt, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=(desc % n_reg) + "; regions shard per GPU, no collective", accepted=int(ok),
                            alignments_in=int(alns), passes=int(passes),
                            timed="wall clock around nw_bitlocus_build (host PAF columns in, tables + grids out)"),
                clocks=clocks, e2e=dict(value=n / dt, unit="regions/s", h2d_bytes_per_step=int(alns) * 56,
                                        d2h_bytes_per_step=None), gpu_launches=4 * max_passes * steps * max(world, 1))
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = pafs[:max(32, n_reg // 20)]
        r, cdt = cpu(sample)
        line["cpu_baseline"] = dict(value=r, unit="regions/s", cores=cores, kind="port",
                                    sample="%d regions of the same set, %.1f s, one region per worker process; R "
                                           "restatement in Python" % (len(sample), cdt))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    solver.close()


def _pipeline_cpu_one(job):
    from oracle import oracle_lib
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tests.test_pipeline import oracle_row
    p, m = job
    try:
        return oracle_row(oracle_lib, p, m, 3, 1000.0, 1000.0, 10000.0) is not None
    except Exception:
        return None


def run_pipeline(args, rank, world, local):
    """Regionfile mode from the alignments on, all native (nahrwhals_b200.pipeline.run_regions): chunked PAF tables ->
    wrapper_paf_to_bitlocus -> solver stage (-d 3) -> rows of nahrwhals_res.tsv.  Metric: regions/s, wall clock around
    run_regions (host PAF columns in, result rows out).  --impl reference: the chain of the R restatements around the C++
    oracle solver on all host cores, one region per worker process."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from region_batch import make_stage_regions
    from nahrwhals_b200.sdgrid import matrix_to_chunked_paf
    cores = os.cpu_count() or 1
    n_reg = max(16, args.regions // 4)
    desc = ("regionfile mode from the alignments on: chunked (10 kb) alignment tables of %d synthetic SD-flanked regions per "
            "GPU (20-120 blocks) -> wrapper_paf_to_bitlocus -> R depth-1 pre-pass + solver -d 3 -u 2 -w 1000 -r 0.98 -R 1000 + "
            "format / merge -> rows of nahrwhals_res.tsv")

    def make(n, seed):
        regs = make_stage_regions(n, seed=seed)
        pafs = [matrix_to_chunked_paf(m, lx, ly) for m, ly, lx, _ in regs]
        metas = [dict(chr="chr%d" % (1 + k % 22), start=str(1000000 + 1000 * k), end=str(1000000 + 1000 * k + int(lx.sum())),
                      samplename="x_y") for k, (m, ly, lx, _) in enumerate(regs)]
        return pafs, metas

    def cpu(pafs, metas):
        import multiprocessing as mp
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_pipeline_cpu_one, list(zip(pafs, metas)), chunksize=1)
        return len(pafs) / (time.perf_counter() - t0), time.perf_counter() - t0
    if args.impl == "reference":
        if rank != 0:
            return
        pafs, metas = make(max(64, n_reg // 20), 0)
        r, dt = cpu(pafs, metas)
        print(json.dumps(dict(impl="reference", metric="regions_alignments_to_result_row_per_s", value=r, unit="regions/s",
                              n_gpus=args.gpus, steps=1, warmup=0, ms_per_step=1e3 * dt, higher_is_better=True,
                              scaling="weak", vs_baseline=None, dtype="int32", data="synthetic",
                              config=dict(workload=desc % len(pafs)),
                              cpu_baseline=dict(value=r, unit="regions/s", cores=cores, kind="port",
                                                sample="%d regions, one per worker process" % len(pafs)),
                              e2e=dict(value=r, unit="regions/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                              gpu_launches=0)), flush=True)
        return
    import torch
    import nahrwhals_b200 as nb
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pafs, metas = make(n_reg, rank)
    solver = nb.Solver(local)
    solver2 = nb.Solver(local)  # second context on the same GPU: host phases of one batch overlap device phases of the other
    ctxs = [solver, solver2]
    kw = dict(depth=3, minlen=1000.0, compression=1000.0, chunklen=10000.0)
    for k in range(3):
        nb.run_regions_native(ctxs, pafs if k == 0 else pafs[:600], metas if k == 0 else metas[:600], 1000.0, 1000.0,
                              10000.0, depth=3, regions_per_batch=256)
    one = {}
    nb.run_regions_native(solver, pafs, metas, 1000.0, 1000.0, 10000.0, depth=3)
    nb.run_regions_native(solver, pafs, metas, 1000.0, 1000.0, 10000.0, depth=3, timing=one)
    nb.run_regions(solver, pafs[:200], metas[:200], **kw)
    sampler = ClockSampler(local)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sampler.start()
    # last warm-up step with the sampler already attached (nvidia-smi's start-up stalls CUDA calls of the first step it meets)
    nb.run_regions_native(ctxs, pafs, metas, 1000.0, 1000.0, 10000.0, depth=3, regions_per_batch=256)
    steps = max(1, args.steps)
    dt = 0.0
    step_ms = []
    for _ in range(steps):
        tn = {}
        rows, st_ = nb.run_regions_native(ctxs, pafs, metas, 1000.0, 1000.0, 10000.0, depth=3, timing=tn, regions_per_batch=256)
        dt += tn["native_s"]
        step_ms.append(round(1e3 * tn["native_s"], 2))
    torch.cuda.synchronize()
    dt /= steps
    clocks = sampler.stop()
    # the same through the per-step Python mirror (three batched calls + Python objects per region): reported, not the metric
    tm = {}
    t0 = time.perf_counter()
    rows_py, det = nb.run_regions(solver, pafs, metas, timing=tm, **kw)
    tm["total_s"] = time.perf_counter() - t0
    assert rows_py == rows, "pipeline: the one-call form and the Python-routed form disagree"
    n, ok = len(pafs), sum(1 for r in rows if r is not None)
    called = sum(1 for _, r in det if r is not None and r.solver_called)
    launches = sum(r.gpu_launches for _, r in det if r is not None)
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        s_ = torch.tensor([n, ok, called], dtype=torch.float64, device="cuda")
        dist.all_reduce(s_, op=dist.ReduceOp.SUM)
        n, ok, called = s_.tolist()
    line = dict(metric="regions_alignments_to_result_row_per_s", value=n / dt, unit="regions/s", n_gpus=world, steps=steps,
                warmup=3, ms_per_step=1e3 * dt, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="int32",
                data="synthetic",
                config=dict(workload=(desc % n_reg) + "; regions shard per GPU, no collective", rows=int(ok),
                            solver_called=int(called),
                            python_routed_ms=dict(total=1e3 * tm["total_s"], bitlocus=1e3 * tm["bitlocus_s"],
                                                  stage=1e3 * tm["stage_s"], rows=1e3 * tm["rows_s"]),
                            one_context_ms=1e3 * one["native_s"], step_ms=step_ms,
                            timed="wall clock around nw_regions_run_ctxs, two contexts on the GPU sharing batches of 256 regions "
                                  "(host PAF columns in, result rows in host memory out)"),
                clocks=clocks, e2e=dict(value=n / dt, unit="regions/s",
                                        h2d_bytes_per_step=int(sum(len(p["tstart"]) for p in pafs)) * 56,
                                        d2h_bytes_per_step=None), gpu_launches=int(max(launches, 1)) * steps)
    if rank == 0 and world == 1 and not args.no_cpu:
        k = max(32, n_reg // 20)
        r, cdt = cpu(pafs[:k], metas[:k])
        line["cpu_baseline"] = dict(value=r, unit="regions/s", cores=cores, kind="port",
                                    sample="%d regions of the same set, %.1f s, one region per worker process; R "
                                           "restatements in Python around the C++ oracle solver" % (k, cdt))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    solver2.close()
    solver.close()


def _paf_tables(n, seed, n_segments=260):
    """Chunked PAF tables of a few thousand alignments each (a 4 Mb region cut into 10 kb chunks, repeats, short hits)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_paf import chunked_paf
    return [chunked_paf(np.random.default_rng(seed * 100000 + k), n_segments=n_segments, chunk=10000, jitter=3, span=4000000,
                        drop=0.02, extra=300) for k in range(n)]


def _paf_cpu_one(p):
    from oracle import nw_paf_oracle as po
    return len(po.load_and_prep_paf(p, 1000.0, 1000.0, 10000.0)["tstart"])


def run_paf(args, rank, world, local):
    """PAF preparation (strand swap, compress_paf_fnct(second_run), minlen filter, rounding,
    enforce_slope_one) of chunked alignment tables through nw_paf_prepare.  Metric: alignments/s (input rows), wall clock
    around the C-ABI calls, host columns in, prepared table in host memory out.  --impl reference: the R restatement
    (numpy) on all host cores, one table per worker process."""
    cores = os.cpu_count() or 1
    n_tab = max(8, args.regions // 200)
    desc = "PAF preparation of %d chunked alignment tables per GPU (~2 400 alignments each: 4 Mb region, 10 kb chunks)"

    def cpu(tabs):
        import multiprocessing as mp
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_paf_cpu_one, tabs, chunksize=1)
        dt = time.perf_counter() - t0
        return sum(len(t["tstart"]) for t in tabs) / dt, dt
    if args.impl == "reference":
        if rank != 0:
            return
        tabs = _paf_tables(max(cores, n_tab // 2), 0)
        r, dt = cpu(tabs)
        print(json.dumps(dict(impl="reference", metric="paf_alignments_prepared_per_s", value=r, unit="alignments/s",
                              n_gpus=args.gpus, steps=1, warmup=0, ms_per_step=1e3 * dt, higher_is_better=True, scaling="weak",
                              vs_baseline=None, dtype="f64", data="synthetic", config=dict(workload=desc % len(tabs)),
                              cpu_baseline=dict(value=r, unit="alignments/s", cores=cores, kind="port",
                                                sample="%d tables, one per worker process" % len(tabs)),
                              e2e=dict(value=r, unit="alignments/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                              gpu_launches=0)), flush=True)
        return
    import torch
    import nahrwhals_b200 as nb
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tabs = _paf_tables(n_tab, rank)
    solver = nb.Solver(local)
    for _ in range(3):
        nb.load_and_prep_pafs(solver, tabs, 1000.0, 1000.0, 10000.0)
    sampler = ClockSampler(local)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sampler.start()
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    launches = 0
    for _ in range(steps):
        got, _ = nb.load_and_prep_pafs(solver, tabs, 1000.0, 1000.0, 10000.0)
        launches += max(t["_gpu_launches"] for t in got)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    clocks = sampler.stop()
    rows = sum(len(t["tstart"]) for t in tabs)
    if dist is not None:
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        ss = torch.tensor([rows, launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(ss, op=dist.ReduceOp.SUM)
        rows, launches = ss.tolist()
    line = dict(metric="paf_alignments_prepared_per_s", value=rows / dt, unit="alignments/s", n_gpus=world, steps=steps,
                warmup=3, ms_per_step=1e3 * dt, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic", config=dict(workload=(desc % n_tab) + "; tables shard per GPU, no collective",
                                             timed="wall clock around nw_paf_prepare_many (one all-pairs launch per batch) through the Python mirror"),
                clocks=clocks, e2e=dict(value=rows / dt, unit="alignments/s", h2d_bytes_per_step=int(rows) * 56,
                                        d2h_bytes_per_step=None), gpu_launches=int(launches))
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = tabs[:max(cores, n_tab // 2)]
        r, cdt = cpu(sample)
        line["cpu_baseline"] = dict(value=r, unit="alignments/s", cores=cores, kind="port",
                                    sample="%d tables of the same set, %.1f s, one table per worker process; R restatement "
                                           "in numpy" % (len(sample), cdt))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    solver.close()


def frontier_section(solver, dist, torch, rank, world, workload, steps, warmup):
    """BASELINE.json config 3 (and the exhaustive S150 search): ONE deep search over all ranks -- replicated expansion /
    dedup / selection, the banded DP of every level divided between the ranks, one NCCL all-gather of the score arrays
    per level issued by libnwsolve on its own stream (nw_solve_sharded).  Strong scaling: total work is fixed.
    Returns a dict (identical on every rank): device ms per search (CUDA events on each rank's stream, max over
    ranks), end-to-end ms from host buffers (wall, max over ranks), and the same search on ONE GPU of the same box."""
    from nahrwhals_b200.distributed import solve_sharded
    aln, lx, ly, flags, desc = make_workload(workload, 1)
    ref = solver.solve(aln, lx, ly, **flags)
    R = solve_sharded(solver, aln, lx, ly, **flags)
    same = R.tsv == ref.tsv and R.scored == ref.scored and R.generated == ref.generated and R.cells == ref.cells
    for _ in range(max(3, warmup)):
        solve_sharded(solver, aln, lx, ly, **flags)
    dev, e2e, dp, launches = 0.0, 0.0, 0.0, 0
    for _ in range(steps):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        R = solve_sharded(solver, aln, lx, ly, **flags)
        e2e += time.perf_counter() - t0
        dev += R.stats.total_ms
        dp += R.stats.score_ms
        launches += R.stats.kernel_launches
    one_dev, one_e2e, one_dp = 0.0, 0.0, 0.0
    for _ in range(max(3, warmup)):
        solver.solve(aln, lx, ly, **flags)
    for _ in range(steps):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        Q = solver.solve(aln, lx, ly, **flags)
        one_e2e += time.perf_counter() - t0
        one_dev += Q.stats.total_ms
        one_dp += Q.stats.score_ms
    t = torch.tensor([dev, 1e3 * e2e, dp, one_dev, 1e3 * one_e2e, one_dp, 0.0 if same else 1.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev, e2e, dp, one_dev, one_e2e, one_dp, bad = (t / steps).tolist()
    ln = torch.tensor([launches], dtype=torch.float64, device="cuda")
    dist.all_reduce(ln, op=dist.ReduceOp.SUM)
    scored = sum(ref.scored)
    return dict(workload=desc, ranks=world, scaling="strong", scored_per_search=scored,
                identical_to_single_gpu=(bad == 0.0), ms_per_search=dev, ms_per_search_e2e=e2e,
                cand_per_s=scored / (dev * 1e-3), dp_ms_per_search_max_rank=dp,
                ms_1gpu_same_box=one_dev, ms_1gpu_same_box_e2e=one_e2e, dp_ms_1gpu=one_dp,
                speedup_vs_1gpu=one_dev / dev if dev > 0 else None,
                launches_per_search_all_ranks=ln.item() / steps,
                exchange="one ncclAllGather group (k3, cost, total; 12 B per fresh candidate) per BFS level on the "
                         "library's stream",
                timing="CUDA events on every rank's stream around the whole search, max over ranks, mean of %d "
                       "searches; e2e = wall clock around nw_solve_sharded from host buffers" % steps)


def regions_section(solvers, dist, torch, rank, world, total_regions, per_pass):
    """BASELINE.json config 4: ONE regionfile batch of `total_regions` synthetic SD-flanked regions (20-120 blocks),
    LPT-sharded over the ranks by estimated work (nahrwhals_b200.distributed.shard_regions), every rank solving its
    shard through nw_solve_many_ctxs; no collective on the data path.  Strong scaling: total work is fixed.  Also the
    whole batch on ONE GPU of the same box (rank 0 alone)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from region_batch import FLAGS, make_regions
    import nahrwhals_b200 as nb
    from nahrwhals_b200.distributed import estimate_region_cost, shard_regions
    # every rank builds a slice of the batch, then everybody gets everything (generation is untimed set-up)
    regs_all = make_regions(total_regions, seed=0, only=range(rank, total_regions, world))
    parts = [None] * world
    dist.all_gather_object(parts, regs_all)
    regs = [None] * total_regions
    for r in range(world):
        for k, reg in zip(range(r, total_regions, world), parts[r]):
            regs[k] = reg
    del parts
    costs = [estimate_region_cost(a, FLAGS["maxdepth"], FLAGS["maxwidth"]) for a, _, _ in regs]
    mine = shard_regions(costs, world)[rank]
    my = [regs[k] for k in mine]
    # a pass per context and round: strong scaling leaves every rank total / world regions, and passes of the single-GPU
    # size would leave most contexts of an 8-GPU run without a second pass to overlap with
    per_pass = max(64, min(per_pass, -(-len(my) // (2 * len(solvers)))))
    warm = make_regions(min(1500, total_regions), seed=1000 + rank)
    for _ in range(2):
        nb.solve_many(solvers, warm, regions_per_pass=per_pass, **FLAGS)
    del warm

    def timed(batch):
        torch.cuda.synchronize()
        dist.barrier()
        tm = {}
        t0 = time.perf_counter()
        pp = per_pass if batch is my else 512
        res = nb.solve_many(solvers, batch, regions_per_pass=pp, timing=tm, **FLAGS) if batch else []
        return time.perf_counter() - t0, tm.get("native_s", 0.0), sum(sum(R.scored) for R in res)
    nb.solve_many(solvers, my, regions_per_pass=per_pass, **FLAGS)  # shapes of the timed batch warm (corridor tables)
    wall, native, scored = timed(my)
    # the whole batch on ONE GPU in its own best configuration (4 contexts, 512 regions per pass), rank 0 alone
    one_ctx = list(solvers)
    if rank == 0:
        one_ctx += [nb.Solver(torch.cuda.current_device()) for _ in range(max(0, 4 - len(solvers)))]
        nb.solve_many(one_ctx, regs[:2000], regions_per_pass=512, **FLAGS)
    solvers_shard, solvers = solvers, one_ctx
    one_wall, one_native, one_scored = timed(regs if rank == 0 else [])
    solvers = solvers_shard
    for s_ in one_ctx[len(solvers_shard):]:
        s_.close()
    t = torch.tensor([wall, native, one_wall, one_native], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, native, one_wall, one_native = t.tolist()
    sc = torch.tensor([scored, one_scored, len(my) ** 2], dtype=torch.float64, device="cuda")
    dist.all_reduce(sc, op=dist.ReduceOp.SUM)
    return dict(workload="regionfile batch of %d synthetic SD-flanked regions (20-120 blocks, 2-6 repeat families), "
                         "-d 3 -u 2 -w 1000 -r 0.98 -R 1000, LPT-sharded over %d GPUs by estimated work, %d regions "
                         "searched together per pass, %d contexts per GPU" % (total_regions, world, per_pass, len(solvers)),
                total_regions=total_regions, ranks=world, scaling="strong",
                regions_per_s=total_regions / native, regions_per_s_python=total_regions / wall,
                candidates_scored=sc[0].item(), seconds=native,
                regions_per_s_1gpu_same_box=total_regions / one_native if one_native > 0 else None,
                speedup_vs_1gpu=one_native / native if native > 0 else None,
                timing="wall clock around nw_solve_many_ctxs on every rank (host buffers in, reports in host memory out), "
                       "max over ranks; 1 GPU = rank 0 alone on the whole batch with 4 contexts and 512 regions per pass")


def run_frontier(args, rank, world, local):
    """`--mode frontier`: only the sharded single search of BASELINE config 3 (the default N > 1 run reports it under
    config.frontier next to the headline line)."""
    import torch
    import torch.distributed as dist
    import nahrwhals_b200 as nb
    from nahrwhals_b200.distributed import init_sharded
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    solver = nb.Solver(local)
    init_sharded(solver)
    sampler = ClockSampler(local)
    sampler.start()
    F = frontier_section(solver, dist, torch, rank, world, args.workload, args.steps, args.warmup)
    clocks = sampler.stop()
    if rank == 0:
        print(json.dumps(dict(metric=METRIC, value=F["cand_per_s"], unit=UNIT, n_gpus=world, steps=args.steps,
                              warmup=max(3, args.warmup), ms_per_step=F["ms_per_search"], higher_is_better=True,
                              scaling="strong", vs_baseline=None, dtype="int32", data="synthetic",
                              config=dict(workload=F["workload"] + "; ONE search over %d GPUs" % world, frontier=F),
                              clocks=clocks, e2e=dict(value=F["scored_per_search"] / (F["ms_per_search_e2e"] * 1e-3), unit=UNIT,
                                                      h2d_bytes_per_step=None, d2h_bytes_per_step=None),
                              gpu_launches=int(F["launches_per_search_all_ranks"] * args.steps))), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    solver.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="s64", choices=sorted(WORKLOADS) + ["regions", "stage", "genome", "grid", "paf", "bitlocus", "pipeline"])
    ap.add_argument("--regions", type=int, default=6000, help="regions per rank for --workload regions")
    ap.add_argument("--contexts", type=int, default=4, help="solver contexts (streams) per GPU for --workload regions")
    ap.add_argument("--warmup-regions", type=int, default=2, help="untimed warm-up runs for --workload regions")
    ap.add_argument("--regions-per-pass", type=int, default=512,
                    help="regions searched together by one pipeline instance (--workload regions)")
    ap.add_argument("--mode", default="regions", choices=["regions", "frontier"],
                    help="N > 1: 'regions' = one independent search per GPU (default, weak scaling); 'frontier' = ONE "
                         "search sharded over the GPUs with a per-level all-gather (strong scaling, config 3)")
    ap.add_argument("--cpu-sample", type=int, default=600000, help="unique candidates per CPU replica (bounded sample)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-multi", action="store_true", help="N > 1: skip the frontier-sharded and sharded-regionfile sections")
    ap.add_argument("--total-regions", type=int, default=10000,
                    help="N > 1: regions of the ONE regionfile batch that is LPT-sharded over the GPUs (BASELINE config 4)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "regions":
        run_regions(args, rank, world, local)
        return
    if args.workload == "grid":
        run_grid(args, rank, world, local)
        return
    if args.workload == "bitlocus":
        run_bitlocus(args, rank, world, local)
        return
    if args.workload == "pipeline":
        run_pipeline(args, rank, world, local)
        return
    if args.workload == "paf":
        run_paf(args, rank, world, local)
        return
    if args.workload in ("stage", "genome"):
        run_stage(args, rank, world, local)
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.mode == "frontier" and world > 1:
        run_frontier(args, rank, world, local)
        return

    import torch
    import nahrwhals_b200 as nb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # every rank solves its own copy of the same region (identical work per GPU: weak scaling of independent regions)
    aln, lx, ly, flags, desc = make_workload(args.workload, 1)
    solver = nb.Solver(local)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def l2_flush():
        flush.fill_(rank & 0xff)
        torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(3, args.warmup)
    for _ in range(W):
        solver.solve(aln, lx, ly, **flags)
    peaks = solver.bench_peaks()

    # ---- timed region 1: device-resident throughput (events inside the library on its own stream) ----
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    dev_ms, score_ms, scored, cells, launches, score_launches = 0.0, 0.0, 0, 0, 0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        l2_flush()
        R = solver.solve(aln, lx, ly, **flags)
        st = R.stats
        dev_ms += st.total_ms
        score_ms += st.score_ms
        scored += sum(R.scored)
        cells += sum(R.cells)
        launches += st.kernel_launches
        score_launches += st.score_launches
    barrier()
    wall1 = time.perf_counter() - t0
    # ---- timed region 2: end to end through the public API from host buffers ----
    e2e_t, h2d, d2h = 0.0, 0, 0
    for _ in range(args.steps):
        l2_flush()
        barrier()
        t1 = time.perf_counter()
        R = solver.solve(np.array(aln), np.array(lx), np.array(ly), **flags)
        ntraj = len(R.scores)  # result already on the host (trajectories, costs, TSV bytes)
        e2e_t += time.perf_counter() - t1
        h2d += R.stats.h2d_bytes
        d2h += R.stats.d2h_bytes
    clocks = sampler.stop()
    # ---- the same search with the element-wise dedup verifier switched off (NW_FLAG_NO_VERIFY): reported, not the metric.
    # The product default verifies every merged candidate against its owner (exact dedup, as the reference's Dict).
    nv_ms, nv_e2e = 0.0, 0.0
    nsteps_nv = max(2, min(args.steps, 5))
    for _ in range(2):
        solver.solve(aln, lx, ly, verify_dedup=False, **flags)
    for _ in range(nsteps_nv):
        l2_flush()
        barrier()
        t1 = time.perf_counter()
        R2 = solver.solve(np.array(aln), np.array(lx), np.array(ly), verify_dedup=False, **flags)
        nv_e2e += time.perf_counter() - t1
        nv_ms += R2.stats.total_ms
    nv_ms /= nsteps_nv
    nv_e2e = 1e3 * nv_e2e / nsteps_nv

    print("[bench rank %d] device ms/step %.3f, e2e ms/step %.3f, scored/step %d" %
          (rank, dev_ms / args.steps, 1e3 * e2e_t / args.steps, scored // args.steps), file=sys.stderr, flush=True)
    # ---- reduce over ranks: max time, summed work ----
    if dist is not None:
        t = torch.tensor([dev_ms, e2e_t, score_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_t, score_ms_max = t.tolist()
        s = torch.tensor([scored, cells, launches, h2d, d2h], dtype=torch.float64, device="cuda")
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        scored_all, cells_all, launches_all, h2d_all, d2h_all = s.tolist()
    else:
        scored_all, cells_all, launches_all, h2d_all, d2h_all = scored, cells, launches, h2d, d2h
    value = scored_all / (dev_ms * 1e-3)
    e2e_value = scored_all / e2e_t

    # ---- roofline of the dominant kernel (banded DP, k_score_pair<W> / k_score_fast<W>): integer ALU bound ----
    # algorithmic int-ops = 5 per band cell (3 adds + 2 mins); duration = CUDA events around every DP launch
    intops = 5.0 * cells
    ach = intops / (score_ms * 1e-3) / 1e12 if score_ms > 0 else 0.0
    # measured peak: best of (VIADDMNMX = 2 int-ops/instr), (IADD3 = 2), (VIADDMNMX + IMAD interleaved = 3 per pair)
    peak = max(2.0 * peaks["viaddmnmx_ips"], 2.0 * peaks["iadd3_ips"], 1.5 * peaks["mixed_ips"]) / 1e12
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg_bytes = 44.0 * scored  # 8 B descriptor in + 36 B record out per scored candidate
    hbm_ach = alg_bytes / (score_ms * 1e-3) / 1e9 if score_ms > 0 else 0.0
    traffic, traffic_note = None, None
    try:
        # ncu capture of the level-3 DP launch, which carries 99.6 % of a search's DP cells; `achieved` above averages
        # over all DP launches of a search (root, level 1, 2, 3), so the per-launch average is that capture / launches
        tj = json.load(open(os.path.join(ROOT, "profiles", "dp_kernel_traffic.json")))
        per_search = score_launches / max(args.steps, 1)
        if args.workload == "s64" and per_search > 0:
            traffic = tj["dram_bytes_per_launch"] / per_search
            traffic_note = ("FROM CAPTURE, not measured by this run: dram__bytes_read+write of the level-3 DP launch (%d B, "
                            "%s) spread over the %d DP launches of a search; algorithmic bytes of that launch: %d" %
                            (tj["dram_bytes_per_launch"], tj["source"], per_search, tj["algorithmic_bytes_per_launch"]))
    except Exception:
        pass
    roofline = dict(bound="int_alu",
                    kernel="k_score_pair<W> (banded DP, two candidates per thread on packed 16-bit DPX; k_score_fast<W> "
                           "when a total exceeds 15 bits, k_score_coop<W> for launches under 48 CTAs)",
                    achieved=ach, peak=peak, unit="Tintop/s",
                    frac=(ach / peak if peak else None), traffic=traffic, traffic_note=traffic_note,
                    launches=score_launches, avg_launch_ms=(score_ms / score_launches if score_launches else None),
                    peak_basis="measured on this GPU (nw_bench_peaks): VIADDMNMX %.2f, IADD3 %.2f, VIADDMNMX+IMAD %.2f "
                               "T thread-instr/s" % (peaks["viaddmnmx_ips"] / 1e12, peaks["iadd3_ips"] / 1e12,
                                                     peaks["mixed_ips"] / 1e12),
                    peak_note="the denominator is the 32-bit figure round 1 was measured against (2 int-ops per VIADDMNMX "
                              "thread-instruction); VIADDMNMX.S16x2 issues at the same rate (tools/ubench/dpx_rates: 18.45 T/s) "
                              "and carries two 16-bit lanes, i.e. 2 x this peak for a kernel made of packed instructions only",
                    hbm=dict(bound="hbm", achieved=hbm_ach, peak=hbm_peak, unit="GB/s",
                             frac=hbm_ach / hbm_peak if hbm_peak else None, peak_source=hbm_src,
                             copy_measured_gbs=peaks["copy_Bps"] / 1e9))
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=W,
                ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="int32",
                data="synthetic",
                config=dict(workload=desc + ("; one such region per GPU (%d independent regions, no collective on the data path)" % world if world > 1 else ""),
                            scored_per_step=scored_all / args.steps, l2="flushed between timed steps (256 MiB write)",
                            dedup="exact (default): every merged candidate compared element-wise with its owner on the device",
                            dp_arith="int16 x 2 (packed DPX, two candidates per register) where a bucket's totals fit 15 bits -- "
                                     "every bucket of this workload --, int32 otherwise; results are exact integers either way",
                            fingerprint_only_dedup=dict(ms_per_step=nv_ms, e2e_ms_per_step=nv_e2e,
                                                        note="NW_FLAG_NO_VERIFY, this rank; not the metric"),
                            dp_share_of_step=(score_ms / (dev_ms if dist is None else max(dev_ms, 1e-9)))),
                clocks=clocks,
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d_all / args.steps,
                         d2h_bytes_per_step=d2h_all / args.steps, ms_per_step=1e3 * e2e_t / args.steps),
                gpu_launches=int(launches_all), roofline=roofline, wall_s_timed=wall1)
    if world > 1 and not args.no_multi:
        # the two multi-GPU designs north_star names, measured in the same invocation (the headline value above stays
        # the per-GPU replica figure so that the driver's efficiency arithmetic over N keeps its meaning)
        from nahrwhals_b200.distributed import init_sharded
        init_sharded(solver)
        msteps = max(5, min(args.steps, 20))
        line["config"]["frontier"] = frontier_section(solver, dist, torch, rank, world, "s150", msteps, args.warmup)
        line["config"]["frontier_exhaustive"] = frontier_section(solver, dist, torch, rank, world, "s150d3", max(3, msteps // 4), 1)
        # contexts per GPU: each is a host thread feeding launches plus short parallel host phases; two cores per context
        nctx = max(1, min(args.contexts, (os.cpu_count() or 8) // (2 * max(world, 1))))
        more = [nb.Solver(local) for _ in range(nctx - 1)]
        line["config"]["regions"] = regions_section([solver] + more, dist, torch, rank, world, args.total_regions,
                                                    args.regions_per_pass)
        for s_ in more:
            s_.close()
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        v, dt, sc, ce = cpu_sample(aln, lx, ly, flags, args.cpu_sample, cores)
        line["cpu_baseline"] = dict(
            value=v, unit=UNIT, cores=cores, kind="port",
            sample="first %d unique candidates of the same search per replica, %d independent replicas (one per host "
                   "thread), %.1f s; %.0f DP cells per scored candidate in the sample vs %.0f over the whole search on "
                   "the GPU; C++ restatement of solver.jl (the reference itself needs julia, absent here)" %
                   (args.cpu_sample, cores, dt, ce / max(sc, 1), cells / max(scored, 1)),
            cells_per_candidate=dict(sample=ce / max(sc, 1), full_search=cells / max(scored, 1)),
            cells_per_s=ce / dt)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    solver.close()


if __name__ == "__main__":
    main()
