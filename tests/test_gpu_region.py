"""GPU parity tests of the region steps (include/nwregion.h): the depth-1 pre-pass and the
whole solver stage of a region through the C ABI vs the R restatement (oracle/nw_region_oracle.py) on identical
gridmatrices -- every row of the result table, the log counters and the logfile summary."""
import os

import numpy as np
import pytest

import nahrwhals_b200 as nb
from nahrwhals_b200.sdgrid import random_grid, sdgrid
from oracle import nw_region_oracle as ro
from tests.test_region_oracle import _rows_equal, load_testrun, oracle_solver

pytestmark = pytest.mark.gpu


def check_prepass(solver, mat, ly, lx, compression=1000, generic=False):
    O = ro.solve_mutation_depth1(ro.Bitl(mat, ly, lx), compression)
    P = nb.solve_mutation_depth1(solver, mat, ly, lx, compression=compression, generic_dp=generic)
    _rows_equal(P.rows, O["res"])
    if O["log"].get("cluttered_boundaries"):
        assert P.log["cluttered_boundaries"]
        return O, P
    # every depth-1 child, whether or not the reference's sequential filters would have looked at it: the GPU's
    # estimate sums and DP result vs the R restatement applied to the child matrix carry_out_compressed_sv builds
    bf = ro.flip_bitl_y_if_needed(ro.Bitl(mat, ly, lx))[0]
    moves = ro.find_sv_opportunities(bf)
    assert [c[:3] for c in P.candidates] == moves
    for c, (p1, p2, sv) in zip(P.candidates, moves):
        child = ro.carry_out_compressed_sv(bf, p1, p2, sv)
        assert c[3] == child.ncol
        if child.ncol < 2:
            continue
        st = ro.ScorerState()
        det = {}
        s = ro.calc_coarse_grained_aln_score(child, st, forcecalc=False, orig_symm=0.0, detail=det)
        est = st.est_highest          # with est_highest = 0 before the call this is the child's own estimate
        assert ((c[4] + c[5]) / (sum(child.rn) + sum(child.cn))) * 100 == est, (p1, p2, sv)
        if det:
            assert (c[6], c[7], c[8]) == (det["cost"], det["total"], s), (p1, p2, sv)
        else:
            assert c[6] == -2 and s == 1.0, (p1, p2, sv)
    assert P.log["depth"] == O["log"]["depth"] and P.log["mut_simulated"] == O["log"]["mut_simulated"]
    assert P.prepass_dp == O["state"].n_eval_calc - 1
    assert P.gpu_launches >= 2
    return O, P


def check_region(solver, oracle, mat, ly, lx, **kw):
    gl = np.concatenate([[0], np.cumsum(lx)])
    O = ro.analyse_region(mat, ly, lx, gl, oracle_solver(oracle), **kw)
    P = nb.analyse_region(solver, mat, ly, lx, gl, **kw)
    _rows_equal(P.rows, O["res"])
    assert P.solver_called == O["julia_called"]
    assert P.mut_max == O["mut_max"]
    assert (P.res_max, P.n_res_max) == (O["res_max"], O["n_res_max"])
    assert P.res_ref == O["res_ref"]
    for k, v in O["log"].items():
        assert P.log[k] == v, (k, P.log, O["log"])
    for k, v in P.maxres.items():
        assert v == O["maxres"].get(k), (k, P.maxres, O["maxres"])
    return O, P


def test_testrun_readme_row(solver, oracle):
    """README.md:135-139 through the CUDA path: res_ref 71.378, res_max 99.595, mut_max 4_12_inv+11_18_del."""
    mat, ly, lx = load_testrun()
    O, P = check_prepass(solver, mat, ly, lx)
    assert P.res_ref == 71.378 and P.res_max == 82.114
    O, P = check_region(solver, oracle, mat, ly, lx, depth=2)
    assert (P.res_ref, P.res_max, P.mut_max) == (71.378, 99.595, "4_12_inv+11_18_del")
    O, P = check_region(solver, oracle, mat, ly, lx, depth=3)
    assert P.res_max == 100.0 and P.solver_called


@pytest.mark.parametrize("seed", range(1, 13))
def test_prepass_sdgrid(solver, seed):
    fams = [[3, 2], [4, 3, 2], [5, 4, 3], [6, 4, 4, 3]][seed % 4]
    n_x = 10 + 4 * seed
    mat, lx, ly, _ = sdgrid(n_x, fams, chain_depth=1 + seed % 3, seed=seed)
    check_prepass(solver, mat, ly, lx, generic=(seed % 5 == 0))
    # upside-down y axis: flip_bitl_y_if_needed has to undo it (and the row names flip with the rows)
    check_prepass(solver, -mat[::-1], ly[::-1], lx, compression=10000)


@pytest.mark.parametrize("seed", range(8))
def test_prepass_random_grids(solver, seed):
    """Unstructured grids: repeats touching the first / last block (border cases of carry_out_compressed_sv and its
    colname repairs -> virtual blocks), short (n_y < 10: full estimate band, threshold 0.9) and lopsided shapes."""
    shapes = [(6, 5), (9, 8), (14, 9), (12, 20), (25, 25), (33, 30), (40, 12), (8, 30)]
    n_x, n_y = shapes[seed]
    mat, lx, ly = random_grid(n_x, n_y, 0.12 if n_x > 20 else 0.25, seed)
    O, P = check_prepass(solver, mat, ly, lx)
    if not P.log["cluttered_boundaries"]:
        assert P.prepass_candidates == len(ro.find_sv_opportunities(ro.flip_bitl_y_if_needed(ro.Bitl(mat, ly, lx))[0]))


def test_prepass_cluttered_and_tiny(solver):
    mat, lx, ly = random_grid(12, 12, 0.05, 3)
    mat[0, 1:8] = lx[1:8]                       # five or more dots on a border: is_cluttered
    O, P = check_prepass(solver, mat, ly, lx)
    assert P.log["cluttered_boundaries"] and P.rows[0]["mut1"] == "ref" and P.rows[0]["eval"] == 0.0
    m = np.array([[5, 5], [0, 5]], dtype=np.int64) * 1000   # 2 x 2: deletion leaves a single column
    check_prepass(solver, m, np.array([5000, 5000]), np.array([5000, 5000]))
    with pytest.raises(Exception):
        nb.solve_mutation_depth1(solver, m[:1], np.array([5000]), np.array([5000, 5000]))


@pytest.mark.parametrize("seed", range(1, 9))
def test_region_sdgrid(solver, oracle, seed):
    """The whole solver stage: pre-pass, 99.8 gate, flip, solver, format_julia_output, ref row, summary."""
    fams = [[3, 2], [4, 3], [4, 3, 2]][seed % 3]
    mat, lx, ly, _ = sdgrid(12 + 5 * seed, fams, chain_depth=1 + seed % 3, seed=100 + seed)
    check_region(solver, oracle, mat, ly, lx, depth=3, init_width=200)
    check_region(solver, oracle, -mat[::-1], ly[::-1], lx, depth=2, minreport=0.9)


def test_region_solved_by_prepass(solver, oracle):
    """A region the depth-1 pre-pass solves (>= 99.8) never reaches the solver."""
    for seed in range(1, 30):
        mat, lx, ly, hidden = sdgrid(20, [2], chain_depth=1, seed=seed)
        O, P = check_region(solver, oracle, mat, ly, lx)
        if not P.solver_called:
            assert P.res_max >= 99.8
            return
    pytest.fail("no single-move region found")


def test_region_batch_equals_single(solver, oracle):
    """nw_region_analyse_many: the pre-pass of all regions in one GPU batch and the forest search for the regions that
    need the solver give, region by region, the tables of nw_region_analyse -- and of the R restatement."""
    regs = []
    for seed in range(40):
        n_x = 10 + (seed * 7) % 45
        fams = [[2], [3, 2], [4, 3], [4, 3, 2], [3, 3, 2, 2]][seed % 5]
        mat, lx, ly, _ = sdgrid(n_x, fams, chain_depth=1 + seed % 3, seed=500 + seed)
        if seed % 4 == 3:
            mat, ly = -mat[::-1], ly[::-1]
        if seed % 9 == 8:
            mat, lx, ly = random_grid(9 + seed % 7, 8 + seed % 5, 0.2, seed)
        regs.append((mat, ly, lx, np.concatenate([[0], np.cumsum(lx)])))
    kw = dict(depth=3, init_width=150, minreport=0.95)
    B = nb.analyse_regions(solver, regs, **kw)
    assert len(B) == len(regs)
    called = 0
    for k, (mat, ly, lx, gl) in enumerate(regs):
        S = nb.analyse_region(solver, mat, ly, lx, gl, **kw)
        assert B[k].tsv == S.tsv and B[k].mut_max == S.mut_max
        assert (B[k].res_ref, B[k].res_max, B[k].n_res_max) == (S.res_ref, S.res_max, S.n_res_max)
        assert B[k].log == S.log and B[k].maxres == S.maxres and B[k].solver_called == S.solver_called
        assert B[k].candidates == S.candidates
        called += S.solver_called
        if k % 5 == 0:
            O = ro.analyse_region(mat, ly, lx, gl, oracle_solver(oracle), **kw)
            _rows_equal(B[k].rows, O["res"])
            assert B[k].mut_max == O["mut_max"]
    assert 5 < called < len(regs)          # both branches of the 99.8 gate are in the batch
    # several contexts sharing the batches (nw_region_analyse_many_ctxs): same tables
    s2 = nb.Solver(0)
    try:
        M = nb.analyse_regions([solver, s2], regs, regions_per_batch=7, **kw)
    finally:
        s2.close()
    for a, b in zip(M, B):
        assert a.tsv == b.tsv and a.mut_max == b.mut_max and a.log == b.log and a.solver_called == b.solver_called
    assert nb.analyse_regions(solver, []) == []


def test_region_golden_fixture_gpu(solver):
    """nw_table_tsv of the CUDA path, byte for byte against the committed tables of the test run."""
    from tests.test_region_oracle import GOLD
    mat, ly, lx = load_testrun()
    gl = np.concatenate([[0], np.cumsum(lx)])
    for d in (2, 3):
        P = nb.analyse_region(solver, mat, ly, lx, gl, depth=d)
        head, body = open(os.path.join(GOLD, "testrun_region_d%d.tsv" % d)).read().split("\n", 1)
        assert P.tsv == body
        assert "mut_max=%s" % P.mut_max in head and "n_res_max=%d " % P.n_res_max in head


def test_region_batch_isolates_bad_regions(solver):
    """A region the stage cannot take (non-positive block length) fails alone when per-region status is requested."""
    good = []
    for seed in range(3):
        mat, lx, ly, _ = sdgrid(16 + 4 * seed, [3, 2], chain_depth=2, seed=900 + seed)
        good.append((mat, ly, lx, np.concatenate([[0], np.cumsum(lx)])))
    mat, ly, lx, gl = good[1]
    bad = (mat, ly, np.where(np.arange(len(lx)) == 2, 0, lx), gl)
    regs = [good[0], bad, good[2]]
    res, rcs = nb.analyse_regions(solver, regs, isolate=True)
    assert rcs[0] == 0 and rcs[2] == 0 and rcs[1] == -5 and res[1] is None
    for k in (0, 2):
        S = nb.analyse_region(solver, *regs[k])
        assert res[k].tsv == S.tsv and res[k].mut_max == S.mut_max
    with pytest.raises(Exception):
        nb.analyse_regions(solver, regs)


@pytest.mark.parametrize("shape", [(40, 8), (8, 40), (35, 25), (28, 45), (60, 31), (31, 9)])
def test_prepass_lopsided_uncluttered(solver, shape):
    """Shapes where the R twin's corridor (chosen by n_y) differs from slow_score's (chosen by the child length), with
    the borders cleaned so that is_cluttered does not short-circuit the pre-pass."""
    n_x, n_y = shape
    mat, lx, ly = random_grid(n_x, n_y, 0.10, 1000 + n_x * 7 + n_y)
    mat[0, 1:] = 0
    mat[-1, :-1] = 0
    mat[1:, 0] = 0
    mat[:-1, -1] = 0
    mat[0, 0] = lx[0]
    mat[-1, -1] = lx[-1]
    assert not ro.is_cluttered(ro.Bitl(mat, ly, lx))
    O, P = check_prepass(solver, mat, ly, lx)
    assert P.prepass_candidates > 0


@pytest.mark.parametrize("shape", [(20, 9), (30, 30), (64, 64), (100, 70), (150, 150), (140, 31), (45, 160), (8, 60)])
def test_rtwin_corridor_scores_large_shapes(solver, oracle, shape):
    """nw_score_batch(force = 2): the R twin's own corridor (chosen by n_y, no early exit) against the C++ restatement
    at sizes the Python one is too slow for -- random arrangements of every length from n_y / 3 to 2 n_x, through the
    register-resident kernels and through the generic kernel.  force = 1 (full corridor) on the same arrangements."""
    n_x, n_y = shape
    rng = np.random.default_rng(n_x * 1000 + n_y)
    mat, lx, ly = random_grid(n_x, n_y, 0.08, n_x + n_y)
    aln = np.ascontiguousarray(np.sign(mat).T.astype(np.int8))
    seqs = [list(range(1, n_x + 1))]
    for _ in range(160):
        L = int(rng.integers(min(max(2, n_y // 3), 2 * n_x), 2 * n_x + 1))
        start = int(rng.integers(1, n_x + 1))
        seq = [((start + t - 1) % n_x) + 1 for t in range(L)]          # a run of consecutive blocks ...
        for _ in range(int(rng.integers(0, 4))):                        # ... with a few inverted stretches
            a = int(rng.integers(0, L))
            b = int(rng.integers(a, min(L, a + 12))) + 1
            seq[a:b] = [-v for v in reversed(seq[a:b])]
        seqs.append(seq)
    lxf, lyf = lx.astype(np.float64), ly.astype(np.float64)
    for force in (2, 1):
        osc, ocost, otot, _ = oracle.score_batch(aln, lxf, lyf, seqs, force=force)
        for generic in (False, True):
            if generic and force == 1 and n_x * n_y > 12000:
                continue                                                 # the generic kernel on full corridors is slow
            sc, cost, tot = solver.score_batch(aln, lx, ly, seqs, force=force, generic_dp=generic)
            want_cost = np.where(np.isinf(ocost), -2, ocost).astype(np.int64)
            assert np.array_equal(cost, want_cost), (force, generic)
            assert np.array_equal(tot, otot.astype(np.int64))
            reach = ~np.isinf(ocost)
            assert np.array_equal(sc[reach], osc[reach])


@pytest.mark.gpu
def test_single_block_region(solver):
    """1 x 1 gridmatrix (one clean alignment): 'ref' at 100 %, no solver; alone and inside a batch."""
    import nahrwhals_b200 as nb
    one = (np.array([[-300000]]), np.array([300000]), np.array([300000]), np.array([0.0, 300000.0]))
    R = nb.analyse_region(solver, *one, depth=3)
    assert (R.res_ref, R.res_max, R.n_res_max, R.mut_max, R.solver_called) == (100.0, 100.0, 1, "ref", False)
    assert R.rows == [dict(eval=100.0, mut1="ref", mut1_start=None, mut1_end=None, mut1_pos_pm=None, mut1_len=None,
                           mut1_len_pm=None, mut2_len=None, mut2_len_pm=None, mut3_len=None, mut3_len_pm=None)]
    from nahrwhals_b200.sdgrid import sdgrid
    mat, lx, ly, _ = sdgrid(16, [3, 2], chain_depth=2, seed=5)
    other = (mat, ly, lx, np.concatenate([[0], np.cumsum(lx)]))
    out = nb.analyse_regions(solver, [other, one, other], depth=2)
    assert out[1].rows == R.rows and out[0].rows == out[2].rows == nb.analyse_region(solver, *other, depth=2).rows
    with pytest.raises(Exception):
        nb.analyse_region(solver, np.array([[5000, 0, 3000]]), np.array([5000]), np.array([5000, 2000, 3000]),
                          np.array([0.0, 5000, 7000, 10000]))
