"""CPU tests of the region stage's HOST logic (nahrwhals_b200/csrc/nw_region_host.h through tests/harness): flip,
find_sv_opportunities, carry_out_compressed_sv incl. the colname repairs, the visited set, the sequential filters of
dfsutil, sorting, annotation, format_julia_output, merge and the logfile summary -- against the R restatement
(oracle/nw_region_oracle.py).  The per-child numbers the GPU phase delivers in production (estimate sums, DP cost /
total / score) are taken from the restatement here, so no GPU is needed; tests/test_gpu_region.py checks the kernels."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from nahrwhals_b200.region import NwRegionSummary
from nahrwhals_b200.sdgrid import random_grid, sdgrid
from oracle import nw_region_oracle as ro
from tests.test_region_oracle import load_testrun, oracle_solver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def HR():
    d = os.path.join(ROOT, "tests", "harness")
    subprocess.check_call(["make", "-s", "-C", d])
    L = C.CDLL(os.path.join(d, "libnw_region_harness.so"))
    L.hr_candidates.restype = C.c_int64
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def child_numbers(bf):
    """What the GPU phase computes, taken from the restatement: per entry (0 = reference arrangement, k = child k) the
    estimate sums sum(rowMaxs) / sum(colMaxs) of the diagonalised matrix and the DP cost / total / score (forced
    corridor for the reference arrangement, the R twin's own corridor for the children, no pre-filters)."""
    moves = ro.find_sv_opportunities(bf)
    ents = [bf] + [ro.carry_out_compressed_sv(bf, *m) for m in moves]
    n = len(ents)
    srow, scol = np.zeros(n, np.int64), np.zeros(n, np.int64)
    cost, total, score = np.full(n, -1, np.int64), np.zeros(n, np.int64), np.zeros(n, np.float64)
    for k, b in enumerate(ents):
        d = ro.diagonalize(b.m, 1 if b.nrow < 10 else 0.25)
        srow[k], scol[k] = int(d.max(axis=1).sum()), int(d.max(axis=0).sum())
        if b.ncol < 2:
            continue
        det = {}
        s = ro.calc_coarse_grained_aln_score(b, ro.ScorerState(), forcecalc=(k == 0), orig_symm=0.0, detail=det)
        if det:
            cost[k], total[k], score[k] = int(det["cost"]), int(det["total"]), s
        else:
            cost[k], score[k] = -2, 1.0
    return moves, ents, srow, scol, cost, total, score


def run_host(HR, mat, ly, lx, gl, kw, numbers, solver_tsv):
    g = np.ascontiguousarray(mat, dtype=np.int64)
    ly64, lx64 = np.ascontiguousarray(ly, np.int64), np.ascontiguousarray(lx, np.int64)
    glf = np.ascontiguousarray(gl, np.float64)
    srow, scol, cost, total, score = numbers
    out = C.create_string_buffer(1 << 22)
    mm = C.create_string_buffer(4096)
    S = NwRegionSummary()
    tsv = solver_tsv.encode() if solver_tsv is not None else None
    rc = HR.hr_region(_p(g), C.c_int(g.shape[1]), C.c_int(g.shape[0]), _p(lx64), _p(ly64), _p(glf), C.c_int(len(glf)),
                      C.c_int(kw.get("depth", 3)), C.c_int(kw.get("maxdup", 2)), C.c_double(kw.get("init_width", 1000)),
                      C.c_double(kw.get("minreport", 0.98)), C.c_double(kw.get("compression", 1000)), _p(srow), _p(scol),
                      _p(cost), _p(total), _p(score), C.c_int64(len(srow)), tsv, C.c_int64(len(tsv) if tsv else 0), out,
                      C.c_int64(len(out)), mm, C.c_int64(len(mm)), C.byref(S))
    return rc, out.value.decode(), mm.value.decode(), S


def parse_rows(tsv):
    lines = tsv.strip("\n").split("\n")
    cols = lines[0].split("\t")
    rows = []
    for ln in lines[1:]:
        row = {}
        for k, v in zip(cols, ln.split("\t")):
            row[k] = None if v == "NA" else (v if (k.startswith("mut") and "_" not in k) else float(v))
        rows.append(row)
    return rows


def check_region_host(HR, oracle, mat, ly, lx, **kw):
    gl = np.concatenate([[0], np.cumsum(lx)])
    captured = {}

    def solver(*a):
        captured["tsv"] = oracle_solver(oracle)(*a)
        return captured["tsv"]
    O = ro.analyse_region(mat, ly, lx, gl, solver, **kw)
    b = ro.Bitl(mat, ly, lx)
    if ro.is_cluttered(b):
        numbers = tuple(np.zeros(1, t) for t in (np.int64, np.int64, np.int64, np.int64, np.float64))
    else:
        bf = ro.flip_bitl_y_if_needed(b)[0]
        numbers = child_numbers(bf)[2:]
    rc, tsv, mut_max, S = run_host(HR, mat, ly, lx, gl, kw, numbers, captured.get("tsv"))
    assert rc == 0, rc
    rows = parse_rows(tsv)
    assert len(rows) == len(O["res"])
    for a, r in zip(rows, O["res"]):
        assert list(a.keys()) == list(r.keys()) and all(a[k] == r[k] for k in a), (a, r)
    assert mut_max == O["mut_max"] and bool(S.solver_called) == O["julia_called"]
    assert (S.res_max, S.n_res_max) == (O["res_max"], O["n_res_max"])
    assert S.res_ref == O["res_ref"]
    for k, v in O["log"].items():
        got = dict(depth=S.search_depth, mut_simulated=S.mut_simulated, mut_tested=S.mut_tested,
                   flip_unsure=bool(S.flip_unsure), cluttered_boundaries=bool(S.cluttered))[k]
        assert got == v, (k, got, v)
    return O, S


def test_host_children_match_carry_out(HR):
    """Position lists of the product vs the matrices carry_out_compressed_sv builds: same columns, same colnames (incl.
    the repairs that move lengths), same move order, on grids whose repeats touch the first and last block."""
    n_checked = 0
    for seed, (n_x, n_y) in enumerate([(6, 5), (9, 8), (14, 9), (12, 20), (25, 25), (7, 7), (10, 6), (5, 12)]):
        mat, lx, ly = random_grid(n_x, n_y, 0.3, 40 + seed)
        b = ro.Bitl(mat, ly, lx)
        if ro.is_cluttered(b):
            continue
        bf, flipped, unsure = ro.flip_bitl_y_if_needed(b)
        moves = ro.find_sv_opportunities(bf)
        g = np.ascontiguousarray(mat, np.int64)
        meta = np.zeros((max(len(moves), 1), 4), np.int32)
        blk = np.zeros(1 << 20, np.int16)
        ln = np.zeros(1 << 20, np.int64)
        flags = np.zeros(4, np.int32)
        n = HR.hr_candidates(_p(g), C.c_int(n_x), C.c_int(n_y), _p(np.ascontiguousarray(lx, np.int64)),
                             _p(np.ascontiguousarray(ly, np.int64)), _p(meta), C.c_int64(len(meta)), _p(blk), _p(ln),
                             C.c_int64(len(blk)), _p(flags))
        assert n == len(moves) and bool(flags[1]) == flipped and bool(flags[2]) == unsure
        off = 0
        for k, (p1, p2, sv) in enumerate(moves):
            assert (int(meta[k, 0]), int(meta[k, 1]), ("del", "dup", "inv")[meta[k, 2]]) == (p1, p2, sv)
            child = ro.carry_out_compressed_sv(bf, p1, p2, sv)
            L = int(meta[k, 3])
            assert L == child.ncol
            bl, le = blk[off:off + L].astype(int), ln[off:off + L]
            off += L
            assert le.tolist() == child.cn, (p1, p2, sv)
            cols = np.sign(bl)[None, :] * bf.m[:, np.abs(bl) - 1]
            assert np.array_equal(cols + 0.0, child.m), (p1, p2, sv)
            n_checked += 1
    assert n_checked > 100


def test_host_flow_readme_row(HR, oracle):
    mat, ly, lx = load_testrun()
    O, S = check_region_host(HR, oracle, mat, ly, lx, depth=2)
    assert (S.res_ref, S.res_max, O["mut_max"]) == (71.378, 99.595, "4_12_inv+11_18_del")
    check_region_host(HR, oracle, mat, ly, lx, depth=3)


@pytest.mark.parametrize("seed", range(1, 11))
def test_host_flow_random_regions(HR, oracle, seed):
    fams = [[3, 2], [4, 3], [4, 3, 2], [2]][seed % 4]
    mat, lx, ly, _ = sdgrid(10 + 3 * seed, fams, chain_depth=1 + seed % 3, seed=300 + seed)
    check_region_host(HR, oracle, mat, ly, lx, depth=3, init_width=150)
    check_region_host(HR, oracle, -mat[::-1], ly[::-1], lx, depth=2, minreport=0.9, compression=5000)
    n_x, n_y = [(6, 5), (9, 8), (14, 9), (12, 14), (16, 12)][seed % 5]
    mat, lx, ly = random_grid(n_x, n_y, 0.25, seed)
    check_region_host(HR, oracle, mat, ly, lx, depth=2, init_width=100)


def test_host_flow_single_block(HR, oracle):
    """A region with one clean alignment grids to a 1 x 1 matrix: the reference reports 100 % for 'ref' and never calls the
    solver (R/evaltools.R:100-101); a single row or column with several blocks stays rejected (see nw_region_host.h)."""
    for v in (250000, -250000):
        mat = np.array([[v]], dtype=np.int64)
        O, S = check_region_host(HR, oracle, mat, np.array([250000]), np.array([250000]), depth=3)
        assert (S.res_ref, S.res_max, S.n_res_max, O["mut_max"]) == (100.0, 100.0, 1, "ref") and not S.solver_called
        assert (S.search_depth, S.mut_simulated) == (1, 1)
    numbers = tuple(np.zeros(1, t) for t in (np.int64, np.int64, np.int64, np.int64, np.float64))
    rc = run_host(HR, np.array([[5000, 0, 3000]]), [5000], [5000, 2000, 3000], [0, 5000, 7000, 10000], {}, numbers, None)[0]
    assert rc == -1


def test_round3_near_ties(HR):
    """R's round(x, 3) (R/evaltools.R:300-302) as the region stage computes it: nearbyint(x * 1000) / 1000 in the host code
    == the restatement, on scores of the form (1 - cost / total) * 100 incl. every exact tie a total of 2^a 5^b units
    produces.  R >= 4.0 rounds by comparing the two 3-decimal candidates floor(x * 1000) / 1000 and ceil(x * 1000) / 1000
    with x itself (its exact rule is not restated here: there is no R in the image to pin it); this test states how far
    ANY rule of that family can be from what is computed here: the results agree except within one part in 10^12 of a
    half-way point, where they may differ by 0.001 ("Contract details")."""
    import math
    HR.hr_round3.restype = C.c_double
    HR.hr_round3.argtypes = [C.c_double]

    def r4_round3(x):  # a candidate-distance rounding (nearer of the two 3-decimal neighbours, ties to the even one)
        x10 = 1000.0 * x
        i10 = math.floor(x10)
        xd, xu = i10 / 1000.0, math.ceil(x10) / 1000.0
        dd, du = x - xd, xu - x
        return xd if (dd < du or (dd == du and i10 % 2 == 0)) else xu

    rng = np.random.default_rng(4)
    xs = []
    for total in (3200000, 2830000, 160000, 1 << 20, 5 ** 8, 64 * 5 ** 6, 2470000):
        for cost in rng.integers(0, total, 3000):
            xs.append((1 - int(cost) / total) * 100)
        step = total // 200000 if total % 200000 == 0 else 0  # cost = odd multiples of total / 200000: x * 1000 = k + 0.5
        if step:
            for k in range(1, 4000, 2):
                xs.append((1 - (k * step) / total) * 100)
    xs += [0.0, 100.0, 71.378, 99.595, 0.0005, 0.3125, 12.3455, 12.3465]
    ndiff = 0
    for x in xs:
        got = HR.hr_round3(x)
        assert got == ro.r_round3(x)
        alt = r4_round3(x)
        if got != alt:
            ndiff += 1
            frac = x * 1000.0 - math.floor(x * 1000.0)
            assert abs(frac - 0.5) < 1e-9 * max(1.0, x * 1000.0) and abs(got - alt) < 0.0011
    assert ndiff < len(xs) // 20
