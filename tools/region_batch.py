#!/usr/bin/env python
"""Regionfile-mode throughput (BASELINE.json config 4): N synthetic SD-flanked regions (20-120 blocks, 2-6 repeat
families of size 2-4), -d 3 -u 2 -w 1000 -r 0.98 -R 1000, solved by T solver contexts on one GPU (one host thread
and one CUDA stream each).  Prints regions/s next to the CPU oracle on all host cores.
usage: python tools/region_batch.py [n_regions] [threads,...] [--cpu]"""
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nahrwhals_b200 as nb  # noqa: E402
from nahrwhals_b200.sdgrid import sdgrid, to_solver_inputs  # noqa: E402

FLAGS = dict(maxdepth=3, maxdup=2, maxwidth=1000, minreport=0.98, maxreport=1000)


def make_regions(n, seed=0, only=None):
    """only: build just these region indices of the n-region set (the random stream is consumed for all of them, so a
    region is the same whoever builds it)."""
    rng = np.random.default_rng(seed)
    want = None if only is None else set(only)
    out = []
    for k in range(n):
        n_x = int(rng.integers(20, 121))
        fams = [int(rng.integers(2, 5)) for _ in range(int(rng.integers(2, 7)))]
        while sum(fams) > n_x:
            fams.pop()
        depth = int(rng.integers(1, 4))
        if want is not None and k not in want:
            continue
        mat, lx, ly, _ = sdgrid(n_x, fams, seed=1000 + k, chain_depth=depth, len_tol=0.3)
        out.append((to_solver_inputs(mat), lx, ly))
    return out


def make_stage_regions(n, seed=0):
    """The same generator, but as the wrapper sees a region: (mat[n_y, n_x] signed bp, len_y, len_x, gridlines_x)."""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        n_x = int(rng.integers(20, 121))
        fams = [int(rng.integers(2, 5)) for _ in range(int(rng.integers(2, 7)))]
        while sum(fams) > n_x:
            fams.pop()
        mat, lx, ly, _ = sdgrid(n_x, fams, seed=1000 + k, chain_depth=int(rng.integers(1, 4)), len_tol=0.3)
        out.append((mat, ly, lx, np.concatenate([[0], np.cumsum(lx)])))
    return out


def genome_plan(n, seed=0, blacklist_frac=0.08):
    """BASELINE.json config 5: the solver stage of whole-genome mode on a synthetic T2T-sized pair.  Window sizes follow
    scan_for_windows (R/wholegenome.R:24: 200 kb / 1 Mb / 5 Mb) -> gridmatrices of 10-40 / 40-100 / 100-124 blocks per
    axis (the reference caps a grid at 250 gridlines, R/nahrwhals.R:55); a blacklist (ref_masked_regions) removes a
    fraction of the windows on the host before the solver stage.  Returns [(index, generator kwargs, cost estimate)] of
    the windows that survive -- cheap to compute for the whole genome on every rank; the matrices themselves are only
    built for a rank's own shard (make_genome_regions)."""
    rng = np.random.default_rng(seed)
    plan = []
    for k in range(n):
        cls = rng.choice(3, p=[0.70, 0.25, 0.05])
        n_x = int(rng.integers(*((10, 41), (40, 101), (100, 125))[cls]))
        fams = [int(rng.integers(2, 5)) for _ in range(int(rng.integers(1, (4, 6, 8)[cls])))]
        while sum(fams) > n_x:
            fams.pop()
        depth = int(rng.integers(1, 4))
        masked = rng.random() < blacklist_frac
        if masked:
            continue
        pairs = sum(m * (m - 1) // 2 for m in fams)
        cost = float(n_x * n_x) * (1.0 + min(float(pairs) ** 3, 1000.0 * pairs * 2))
        plan.append((k, dict(n_x=n_x, families=fams, seed=7000 + k, chain_depth=depth, len_tol=0.2), cost))
    return plan


def make_genome_regions(plan, idx):
    out = []
    for q in idx:
        _, kw, _ = plan[q]
        mat, lx, ly, _ = sdgrid(**kw)
        out.append((mat, ly, lx, np.concatenate([[0], np.cumsum(lx)])))
    return out


def _stage_cpu_one(reg):
    """One region through the R restatement (pre-pass, gate, flip, C++ oracle solver, format, merge, summary)."""
    from oracle import nw_region_oracle as ro, oracle_lib
    mat, ly, lx, gl = reg

    def solver(aln, lx_, ly_, d, u, w, r, R):
        return oracle_lib.solve(aln, lx_, ly_, maxdepth=d, maxdup=u, maxwidth=int(w), minreport=r, maxreport=R,
                                full=False).tsv
    A = ro.analyse_region(mat, ly, lx, gl, solver, depth=3, maxdup=2, init_width=1000, minreport=0.98)
    return A["mut_max"], A["res_max"], A["julia_called"]


def run_stage_cpu(regions, procs):
    """CPU arm of the region stage: one region per worker process at a time (what R's threads= does)."""
    import multiprocessing as mp
    from oracle import oracle_lib
    oracle_lib.lib()
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(procs) as pool:
        res = pool.map(_stage_cpu_one, regions, chunksize=1)
    dt = time.perf_counter() - t0
    return len(regions) / dt, res, dt


def run_gpu(regions, threads, device=0):
    solvers = [nb.Solver(device) for _ in range(threads)]
    for s in solvers:  # warm-up (pool, module load)
        s.solve(*regions[0], **FLAGS)
    t0 = time.perf_counter()
    res = nb.solve_batch(solvers, regions, **FLAGS)  # nw_solve_batch: native worker thread per context
    dt = time.perf_counter() - t0
    for s in solvers:
        s.close()
    return len(regions) / dt, sum(sum(R.scored) for R in res), dt


def run_cpu(regions, threads):
    from oracle import oracle_lib
    oracle_lib.lib()
    nxt = [0]
    lock = threading.Lock()
    scored = [0] * threads

    def work(t):
        while True:
            with lock:
                k = nxt[0]
                nxt[0] += 1
            if k >= len(regions):
                return
            aln, lx, ly = regions[k]
            R = oracle_lib.solve(aln, lx, ly, full=False, **FLAGS)
            scored[t] += int(R.scored.sum())

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    dt = time.perf_counter() - t0
    return len(regions) / dt, sum(scored), dt


def run_many(regions, per_pass, device=0, nctx=1):
    ss = [nb.Solver(device) for _ in range(nctx)]
    s = ss if nctx > 1 else ss[0]
    nb.solve_many(s, regions[:min(len(regions), per_pass * nctx)], regions_per_pass=per_pass, **FLAGS)  # warm-up
    t0 = time.perf_counter()
    res = nb.solve_many(s, regions, regions_per_pass=per_pass, **FLAGS)
    dt = time.perf_counter() - t0
    for q in ss:
        q.close()
    hs = [0.0, 0.0, 0.0]
    dev = 0.0
    for k in range(0, len(res), per_pass):
        for q in range(3):
            hs[q] += res[k].stats.host_ms[q]
        dev += res[k].stats.total_ms
    st = res[0].stats
    print("   first pass: levels %d, level_ms %s, generated %s, scored %s, launches %d" % (
        st.n_levels, ["%.2f" % v for v in list(st.level_ms)[:st.n_levels]], list(st.generated)[:st.n_levels],
        list(st.scored)[:st.n_levels], st.kernel_launches))
    print("   native phases over all passes: setup %.1f ms, search %.1f ms (device timeline %.1f ms), report %.1f ms; "
          "python wall %.1f ms" % (hs[0], hs[1], dev, hs[2], dt * 1e3))
    return len(regions) / dt, sum(sum(R.scored) for R in res), dt


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    tl = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 4, 16]
    t0 = time.time()
    regs = make_regions(n)
    print("generated %d regions in %.1f s" % (n, time.time() - t0), flush=True)
    if "--cpu" in sys.argv:
        c = os.cpu_count()
        r, sc, dt = run_cpu(regs, c)
        print("CPU oracle, %d threads: %.1f regions/s, %.3g candidates/s (%.2f s)" % (c, r, sc / dt, dt), flush=True)
    for t in tl:
        r, sc, dt = run_gpu(regs, t)
        print("GPU, %d contexts: %.1f regions/s, %.3g candidates/s (%.2f s)" % (t, r, sc / dt, dt), flush=True)
    for per, nctx in ((16, 1), (64, 1), (256, 1), (1024, 1), (256, 2), (256, 3), (512, 2), (512, 3), (1024, 2)):
        r, sc, dt = run_many(regs, per, nctx=nctx)
        print("GPU, %d pipeline(s), %d regions per pass: %.1f regions/s, %.3g candidates/s (%.2f s)" % (
            nctx, per, r, sc / dt, dt), flush=True)
