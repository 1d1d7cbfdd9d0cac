#!/usr/bin/env python
"""Writes tests/golden/testrun_region_d{2,3}.tsv: the `res` table of the whole solver stage of the bundled test run
(R depth-1 pre-pass -> solver -> format_julia_output -> merge) as computed by the CPU restatements
(oracle/nw_region_oracle.py around oracle/nw_oracle.cpp), in the text form of nw_table_tsv().  Depth 2 reproduces the
reference's published row (README.md:135-139): res_ref 71.378, res_max 99.595, mut_max 4_12_inv+11_18_del.
usage: python tests/golden/make_region_fixture.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import nw_region_oracle as ro, oracle_lib  # noqa: E402


def fmt(v):
    if v is None:
        return "NA"
    if isinstance(v, str):
        return v
    v = float(v)
    if v == int(v) and abs(v) < 1e15:
        return "%d" % int(v)
    return repr(v)


def render(rows):
    cols = list(rows[0].keys())
    out = ["\t".join(cols)]
    for r in rows:
        out.append("\t".join(fmt(r[c]) for c in cols))
    return "\n".join(out) + "\n"


def solver(aln, lx, ly, d, u, w, r, R):
    return oracle_lib.solve(aln, lx, ly, maxdepth=d, maxdup=u, maxwidth=int(w), minreport=r, maxreport=R).tsv


if __name__ == "__main__":
    mat = np.loadtxt(os.path.join(HERE, "testrun_grid_mut.tsv"), dtype=np.float64, ndmin=2)
    lens = open(os.path.join(HERE, "testrun_grid_mut_lens.tsv")).read().strip().split("\n")
    ly = np.array(lens[0].split("\t"), dtype=np.float64)
    lx = np.array(lens[1].split("\t"), dtype=np.float64)
    gl = np.concatenate([[0], np.cumsum(lx)])
    for d in (2, 3):
        A = ro.analyse_region(mat, ly, lx, gl, solver, depth=d)
        with open(os.path.join(HERE, "testrun_region_d%d.tsv" % d), "w") as f:
            f.write("# res_ref=%s res_max=%s n_res_max=%d mut_max=%s\n" % (fmt(A["res_ref"]), fmt(A["res_max"]), A["n_res_max"],
                                                                       A["mut_max"]))
            f.write(render(A["res"]))
        print(d, A["res_ref"], A["res_max"], A["n_res_max"], A["mut_max"], len(A["res"]))
