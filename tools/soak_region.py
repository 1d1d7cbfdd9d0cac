#!/usr/bin/env python
"""Soak test of the region stage and the gridding against the R restatements on many random inputs (slow; not part of
the pytest suites).  usage: python tools/soak_region.py [n_regions] [seed0]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nahrwhals_b200 as nb  # noqa: E402
from nahrwhals_b200.sdgrid import random_grid, sdgrid, matrix_to_alignments  # noqa: E402
from oracle import nw_grid_oracle as go, nw_region_oracle as ro, oracle_lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed0)


def osolver(a, x, y, d, u, w, r, R):
    return oracle_lib.solve(a, x, y, maxdepth=d, maxdup=u, maxwidth=int(w), minreport=r, maxreport=R).tsv


regs = []
for k in range(n):
    if rng.random() < 0.6:
        n_x = int(rng.integers(8, 60))
        fams = [int(rng.integers(2, 5)) for _ in range(int(rng.integers(1, 5)))]
        while sum(fams) > n_x:
            fams.pop()
        mat, lx, ly, _ = sdgrid(n_x, fams, chain_depth=int(rng.integers(1, 4)), seed=seed0 * 100000 + k)
        if rng.random() < 0.25:
            mat, ly = -mat[::-1], ly[::-1]
    else:
        n_x, n_y = int(rng.integers(4, 30)), int(rng.integers(4, 30))
        mat, lx, ly = random_grid(n_x, n_y, float(rng.choice([0.08, 0.15, 0.3])), seed0 * 100000 + k)
        if rng.random() < 0.7:  # clean borders so that most grids get past is_cluttered
            mat[0, 1:] = 0
            mat[-1, :-1] = 0
            mat[1:, 0] = 0
            mat[:-1, -1] = 0
            mat[0, 0], mat[-1, -1] = lx[0], lx[-1]
    regs.append((mat, ly, lx, np.concatenate([[0], np.cumsum(lx)])))
kw = dict(depth=int(rng.integers(1, 4)), init_width=int(rng.choice([50, 200, 1000])), minreport=float(rng.choice([0.9, 0.98])))
s = nb.Solver(0)
t0 = time.time()
B = nb.analyse_regions(s, regs, **kw)
print("product: %d regions in %.2f s (%s)" % (n, time.time() - t0, kw), flush=True)
bad = 0
t0 = time.time()
for k, (mat, ly, lx, gl) in enumerate(regs):
    O = ro.analyse_region(mat, ly, lx, gl, osolver, **kw)
    P = B[k]
    ok = len(P.rows) == len(O["res"]) and all(list(a.keys()) == list(b.keys()) and all(a[c] == b[c] for c in a)
                                               for a, b in zip(P.rows, O["res"]))
    ok = ok and P.mut_max == O["mut_max"] and P.res_ref == O["res_ref"] and P.res_max == O["res_max"] and \
        P.n_res_max == O["n_res_max"] and P.solver_called == O["julia_called"]
    if not ok:
        bad += 1
        print("MISMATCH region", k, mat.shape, P.mut_max, O["mut_max"], P.res_max, O["res_max"], flush=True)
print("region stage: %d regions checked against the restatement in %.1f s, %d mismatches; solver called for %d" %
      (n, time.time() - t0, bad, sum(1 for r in B if r.solver_called)), flush=True)
# gridding
pafs = [matrix_to_alignments(m, lx, ly) for (m, ly, lx, _) in regs if np.count_nonzero(m) > 0]
G, rcs = nb.build_grids(s, pafs)
gbad = 0
for k, p in enumerate(pafs):
    try:
        O = go.paf_to_gridmatrix(*p)
    except (IndexError, RuntimeError, KeyError):
        gbad += G[k] is not None
        continue
    ok = G[k] is not None and G[k].gridlines_x.tolist() == O["gridlines_x"] and G[k].gridlines_y.tolist() == O["gridlines_y"] \
        and np.array_equal(G[k].mat, O["mat"]) and G[k].converged == O["converged"]
    if not ok:
        gbad += 1
        print("GRID MISMATCH", k, rcs[k], flush=True)
print("gridding: %d alignment sets checked, %d mismatches" % (len(pafs), gbad), flush=True)
s.close()
sys.exit(1 if (bad or gbad) else 0)
