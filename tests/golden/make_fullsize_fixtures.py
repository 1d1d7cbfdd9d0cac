#!/usr/bin/env python
"""Full-size oracle fixtures: the CPU restatement of solver.jl (oracle/nw_oracle.cpp) run ONCE over the exact
workloads bench.py times (s64 = BASELINE config 2, s150 = config 3, s150d3 = the north_star target) and over a
200-window sample of the whole-genome plan (config 5).  What is kept is small: per-level counters
(solver.jl:589-642 emissions, slow_score calls :536, DP cells), the number of ids, the best score, and SHA-256 digests
of the per-id arrays (discovery level, DP terminal cost, total, Float32 score bits), of every provenance edge
(child-sorted: parent id, move kind/start/stop) and of the report TSV (write_trajectories :747-753).

    python tests/golden/make_fullsize_fixtures.py [s64 s150 s150d3 genome]      (minutes of CPU; s64 needs ~6 GB)

tests/test_gpu_fullsize.py compares the CUDA path with these files on the GPU box (where the oracle would need
minutes per search).  TEST INFRASTRUCTURE: this script is the only writer of tests/golden/fullsize_*.json."""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def id_digests(level, cost_bp, total_bp, score_f32):
    """Digests over the per-id arrays in id order (id 1 = root first).  cost_bp: -1 early exit, -2 unreachable."""
    return dict(level=sha(np.asarray(level, np.int32)), cost=sha(np.asarray(cost_bp, np.int64)),
                total=sha(np.asarray(total_bp, np.int64)), score=sha(np.asarray(score_f32, np.float32).view(np.uint32)))


def edge_digest(child, parent, moves3):
    """Edges sorted by child id (stable: discovery order within a child) -> one digest of (child, parent, kind, start, stop)."""
    child = np.asarray(child, np.int64)
    order = np.argsort(child, kind="stable")
    return sha(child[order].astype(np.int32), np.asarray(parent, np.int64)[order].astype(np.int32),
               np.asarray(moves3, np.int32).reshape(-1, 3)[order])


def canon_rows(rows):
    """One canonical text for a result table: keys in order, numbers as repr(float), missing / NaN as NA."""
    out = []
    for r in rows:
        cells = []
        for k, v in r.items():
            if hasattr(v, "item"):
                v = v.item()
            if v is None or (isinstance(v, float) and v != v):
                t = "NA"
            elif isinstance(v, bool):
                t = "T" if v else "F"
            elif isinstance(v, (int, float)):
                t = repr(float(v))
            else:
                t = str(v)
            cells.append("%s=%s" % (k, t))
        out.append("\t".join(cells))
    return "\n".join(out)


def oracle_search_fixture(aln, lx, ly, flags):
    from oracle import oracle_lib
    t0 = time.time()
    O = oracle_lib.solve(aln, lx.astype(np.float64), ly.astype(np.float64), **flags)
    dt = time.time() - t0
    ocost = np.where(np.isnan(O.cost), -1.0, np.where(np.isinf(O.cost), -2.0, O.cost)).astype(np.int64)
    child = np.repeat(np.arange(1, O.n_ids + 1), np.diff(O.edge_off))
    scored = [int(v) for v in O.scored]
    cells = [int(v) for v in O.cells]
    return dict(flags=flags, n_x=int(aln.shape[0]), n_y=int(aln.shape[1]),
                generated=[int(v) for v in O.generated], scored=scored, cells=cells,
                frontier=[int(v) for v in O.frontier], n_ids=int(O.n_ids), best=float(O.best),
                cells_per_scored=sum(cells) / max(1, sum(scored)),
                ids=id_digests(O.level, ocost, O.total.astype(np.int64), O.score),
                n_edges=int(len(O.edge_parent)), edges=edge_digest(child, O.edge_parent, O.edge_move),
                n_trajs=int(O.n_trajs), tsv_sha256=hashlib.sha256(O.tsv.encode()).hexdigest(),
                tsv_head=O.tsv.split("\n")[:3], oracle_seconds=round(dt, 1), oracle_search_seconds=round(O.seconds, 1))


def bench_workload(name):
    import bench
    aln, lx, ly, flags, desc = bench.make_workload(name, 1)
    fx = oracle_search_fixture(aln, lx, ly, flags)
    fx["workload"] = desc
    fx["bench_name"] = name
    return fx


GENOME_N, GENOME_SAMPLE = 16000, 200


def _genome_one(job):
    from oracle import nw_region_oracle as ro, oracle_lib
    q, (mat, ly, lx, gl) = job

    def solver(aln, lx_, ly_, d, u, w, r, R):
        return oracle_lib.solve(aln, lx_, ly_, maxdepth=d, maxdup=u, maxwidth=int(w), minreport=r, maxreport=R,
                                full=False).tsv
    A = ro.analyse_region(mat, ly, lx, gl, solver, depth=3, maxdup=2, init_width=1000, minreport=0.98)
    rows = A["res"]
    txt = canon_rows(rows)
    return dict(window=int(q), n_x=int(mat.shape[1]), n_y=int(mat.shape[0]), mut_max=A["mut_max"],
                res_ref=float(A["res_ref"]), res_max=float(A["res_max"]), n_res_max=int(A["n_res_max"]),
                solver_called=bool(A["julia_called"]), n_rows=len(rows), rows_sha256=hashlib.sha256(txt.encode()).hexdigest())


def genome_sample():
    """Every k-th window of the plan bench.py --workload genome builds (genome_plan(16000, seed 0)) through the R
    restatement of the solver stage around the oracle solver."""
    import multiprocessing as mp
    from region_batch import genome_plan, make_genome_regions
    plan = genome_plan(GENOME_N, seed=0)
    step = max(1, len(plan) // GENOME_SAMPLE)
    idx = list(range(0, len(plan), step))[:GENOME_SAMPLE]
    regs = make_genome_regions(plan, idx)
    with mp.get_context("fork").Pool(os.cpu_count() or 4) as pool:
        out = pool.map(_genome_one, list(zip(idx, regs)), chunksize=1)
    return dict(plan=dict(n=GENOME_N, seed=0, windows_after_blacklist=len(plan), step=step), windows=out)


if __name__ == "__main__":
    names = sys.argv[1:] or ["s150", "s150d3", "genome", "s64"]
    for nm in names:
        t0 = time.time()
        fx = genome_sample() if nm == "genome" else bench_workload(nm)
        path = os.path.join(HERE, "fullsize_%s.json" % nm)
        with open(path, "w") as f:
            json.dump(fx, f, indent=1)
            f.write("\n")
        print("%s: %.1f s -> %s" % (nm, time.time() - t0, path), flush=True)
