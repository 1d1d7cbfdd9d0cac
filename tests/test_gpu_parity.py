"""GPU parity tests (B200): the CUDA path through the C ABI vs the oracle on identical inputs.
Bit-exact integer DP costs, Float32 scores, discovery ids, provenance edges and the output TSV."""
import os
import subprocess

import numpy as np
import pytest

from nahrwhals_b200.sdgrid import random_grid, sdgrid, to_solver_inputs, write_tsvs
from tests.util import assert_search_equal

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_moveset(solver, oracle, testrun):
    from oracle import nw_oracle_py as op
    for aln in (testrun[0], to_solver_inputs(sdgrid(40, [5, 4, 3], seed=3)[0])):
        dd, iv = solver.moveset(aln)
        pdd, piv = op.to_moveset(aln.tolist())
        assert np.array_equal(dd, np.array(pdd)) and np.array_equal(iv, np.array(piv))


def test_readme_known_answers(solver, testrun):
    """README.md:135-139 of the reference through the CUDA scorer."""
    aln, lx, ly = testrun
    sc, cost, total = solver.score_batch(aln, lx, ly, [list(range(1, 20))], force=True)
    assert (sc[0], cost[0], total[0]) == (71.378, 810000, 2830000)
    sc, cost, total = solver.score_batch(aln, lx, ly, [[1, 2, 3, 4, -11, -10, -9, -8, -7, -6, 18, 19]])
    assert (sc[0], cost[0], total[0]) == (99.595, 10000, 2470000)
    sc, cost, total = solver.score_batch(aln, lx, ly, [list(range(1, 20))])
    assert sc[0] == 0.0 and cost[0] == -1


@pytest.mark.parametrize("generic", [False, True])
def test_testrun_search(solver, oracle, testrun, generic):
    """Config 1: the bundled test run with the flags R passes (R/juliasolver.R:42-47).  generic=True also forces the
    fallback kernels (generic DP, pair-testing expansion)."""
    aln, lx, ly = testrun
    kw = dict(maxdepth=3, maxdup=2, maxwidth=1000, minreport=0.98, maxreport=1000)
    G = solver.solve(aln, lx, ly, keep_ids=True, generic_dp=generic, pair_expand=generic, **kw)
    O = oracle.solve(aln, lx, ly, **kw)
    assert_search_equal(G, O)
    assert G.tsv == open(os.path.join(GOLD, "testrun_oracle_res.tsv")).read()
    assert G.generated == [17, 291, 4484]


def test_testrun_files_and_cli(solver, tmp_path):
    """File + CLI contract (solver.jl:755-812): same TSV bytes as the committed golden result."""
    out = tmp_path / "res.tsv"
    solver.solve_files(os.path.join(GOLD, "testrun_grid_mut.tsv"), os.path.join(GOLD, "testrun_grid_mut_lens.tsv"),
                       out, maxdepth=3, maxdup=2, maxwidth=1000, minreport=0.98, maxreport=1000)
    gold = open(os.path.join(GOLD, "testrun_oracle_res.tsv")).read()
    assert out.read_text() == gold
    out2 = tmp_path / "res2.tsv"
    cli = os.path.join(ROOT, "nahrwhals_b200", "csrc", "nwsolve")
    subprocess.check_call([cli, "-d", "3", "-u", "2", "-w", "1000", "-r", "0.98", "-R", "1000",
                           os.path.join(GOLD, "testrun_grid_mut.tsv"), os.path.join(GOLD, "testrun_grid_mut_lens.tsv"),
                           str(out2)])
    assert out2.read_text() == gold
    # failure: non-zero exit and no output file (R/juliasolver.R:56-59 relies on the file's absence)
    out3 = tmp_path / "res3.tsv"
    rc = subprocess.call([cli, os.path.join(GOLD, "nope.tsv"), os.path.join(GOLD, "testrun_grid_mut_lens.tsv"),
                          str(out3)], stderr=subprocess.DEVNULL)
    assert rc != 0 and not out3.exists()


def _random_candidates(rng, n_x, count, maxmult=3):
    seqs = []
    for _ in range(count):
        mode = rng.integers(0, 3)
        if mode == 0:
            L = int(rng.integers(1, maxmult * n_x + 1))
            s = rng.integers(1, n_x + 1, L) * rng.choice([-1, 1], L)
        else:
            s = np.arange(1, n_x + 1)
            for _ in range(int(rng.integers(0, 4))):
                L = len(s)
                if L < 3:
                    break
                a = int(rng.integers(0, L - 1))
                b = int(rng.integers(a + 1, L))
                k = rng.integers(0, 3)
                if k == 0 and L + (b - a) <= maxmult * n_x:
                    s = np.concatenate([s[:b + 1], s[a + 1:]])
                elif k == 1:
                    s = np.concatenate([s[:a], s[b:]])
                else:
                    s = np.concatenate([s[:a + 1], -s[a + 1:b][::-1], s[b:]])
        seqs.append([int(v) for v in s])
    return seqs


@pytest.mark.parametrize("shape", [(19, 11), (64, 64), (150, 150), (150, 110), (9, 8), (30, 10), (5, 1), (1, 7),
                                   (33, 12), (120, 250), (250, 120), (200, 330)])
def test_score_batch_random(solver, oracle, shape):
    """slow_score parity on raw candidates: every corridor regime (row<10, <30, >=30), early exit boundary,
    unreachable terminals, n_y == 1, both DP kernels, and the forcecalc twin."""
    n_x, n_y = shape
    rng = np.random.default_rng(n_x * 1000 + n_y)
    if min(n_x, n_y) >= 19:
        fams = [int(rng.integers(2, 7)) for _ in range(4)]
        mat, lx, ly, _ = sdgrid(n_x, fams, seed=n_y, chain_depth=2)
        if mat.shape[0] != n_y:  # fix the target length by cropping/tiling rows
            idx = np.arange(n_y) % mat.shape[0]
            mat, ly = mat[idx], ly[idx]
    else:
        mat, lx, ly = random_grid(n_x, n_y, 0.25, n_x + n_y)
    aln = to_solver_inputs(mat)
    seqs = _random_candidates(rng, n_x, 600 if n_x * n_y < 30000 else 200)
    for force in (False, True):
        osc, ocost, ototal, _ = oracle.score_batch(aln, lx, ly, seqs, force=force)
        ocost_i = np.where(np.isnan(ocost), -1.0, np.where(np.isinf(ocost), -2.0, ocost)).astype(np.int64)
        # the three DP kernels: packed 16-bit register form (default when the totals fit 15 bits), 32-bit register form, generic
        for generic, dp32 in ((False, False), (False, True), (True, False)):
            sc, cost, total = solver.score_batch(aln, lx, ly, seqs, force=force, generic_dp=generic, dp32=dp32)
            assert np.array_equal(cost, ocost_i), (shape, force, generic, dp32)
            assert np.array_equal(total, ototal.astype(np.int64))
            assert np.array_equal(sc, osc)  # Float64 scores incl. -inf, exact


def test_score_batch_packed_range_boundary(solver, oracle):
    """Packed 16-bit DP (k_score_pair) at the edge of its range: the same gridmatrix with block lengths scaled so that the
    totals of the candidates straddle 2^15 units -- the short children run packed, the long ones on the 32-bit kernel,
    in one call -- and with the largest scale that still fits.  Costs / totals / scores exact vs the oracle."""
    rng = np.random.default_rng(77)
    mat, lx, ly, _ = sdgrid(48, [6, 5, 4, 3], seed=9, chain_depth=2)
    aln = to_solver_inputs(mat)
    n_x = aln.shape[0]
    seqs = _random_candidates(rng, n_x, 500)
    base_lx, base_ly = lx // np.gcd.reduce(np.concatenate([lx, ly])), ly // np.gcd.reduce(np.concatenate([lx, ly]))
    for target in (20000, 32767, 40000, 70000):
        # scale so that sum(ly) + 48 * max(lx) ~ target units; +1 on one block keeps the gcd at 1
        f = max(1, int(target / float(base_ly.sum() + n_x * base_lx.max())))
        lx2, ly2 = base_lx * f, base_ly * f
        lx2 = lx2.copy()
        lx2[0] += 1
        for force in (False, True):
            osc, ocost, ototal, _ = oracle.score_batch(aln, lx2, ly2, seqs, force=force)
            ocost_i = np.where(np.isnan(ocost), -1.0, np.where(np.isinf(ocost), -2.0, ocost)).astype(np.int64)
            for dp32 in (False, True):
                sc, cost, total = solver.score_batch(aln, lx2, ly2, seqs, force=force, dp32=dp32)
                assert np.array_equal(cost, ocost_i), (target, force, dp32)
                assert np.array_equal(total, ototal.astype(np.int64))
                assert np.array_equal(sc, osc)


CASES = [
    # (name, generator args, solver flags)
    ("sd24_exh", dict(n_x=24, families=[4, 3, 3], seed=11, chain_depth=3), dict(maxdepth=3, maxdup=2)),
    ("sd40_w50", dict(n_x=40, families=[5, 4, 3, 2], seed=12, chain_depth=3), dict(maxdepth=4, maxdup=2, maxwidth=50)),
    ("sd64_exh", dict(n_x=64, families=[6, 5, 5, 4], seed=1, chain_depth=3), dict(maxdepth=3, maxdup=2)),
    ("sd64_d2", dict(n_x=64, families=[10, 9, 8, 8, 7], seed=1, chain_depth=3), dict(maxdepth=2, maxdup=2)),
    ("sd150_w100", dict(n_x=150, families=[6, 5, 4, 4, 3], seed=7, chain_depth=4),
     dict(maxdepth=4, maxdup=2, maxwidth=100)),
    # beam of 1 000 out of >100 000 candidates per level (the bench's S150 workload): histogram threshold selection
    ("sd150_w1000", dict(n_x=150, families=[6, 5, 4, 4, 3], seed=7, chain_depth=4),
     dict(maxdepth=4, maxdup=2, maxwidth=1000)),
    ("sd64_w300", dict(n_x=64, families=[6, 5, 5, 4], seed=3, chain_depth=3), dict(maxdepth=3, maxdup=2, maxwidth=300)),
    ("sd30_dup0", dict(n_x=30, families=[4, 4, 3], seed=5, chain_depth=2), dict(maxdepth=3, maxdup=0, maxwidth=1000)),
    ("sd30_dup3_w7", dict(n_x=30, families=[5, 4], seed=6, chain_depth=3), dict(maxdepth=5, maxdup=3, maxwidth=7)),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_search_parity(solver, oracle, case):
    """Full BFS parity: counts per level, every id's level/cost/total/score, every provenance edge, TSV."""
    _, gen, flags = case
    mat, lx, ly, _ = sdgrid(**gen)
    aln = to_solver_inputs(mat)
    kw = dict(minreport=0.98, maxreport=1000, **flags)
    O = oracle.solve(aln, lx, ly, **kw)
    G = solver.solve(aln, lx, ly, keep_ids=True, **kw)
    assert_search_equal(G, O)
    if case[0] in ("sd40_w50", "sd30_dup3_w7"):  # fallback expansion kernel must agree too
        G2 = solver.solve(aln, lx, ly, keep_ids=True, pair_expand=True, **kw)
        assert_search_equal(G2, O)
    if case[0] in ("sd64_exh", "sd150_w1000", "sd30_dup3_w7"):  # the default ran the packed 16-bit DP: 32-bit form must agree
        G3 = solver.solve(aln, lx, ly, keep_ids=True, dp32=True, **kw)
        assert_search_equal(G3, O)


def test_search_small_and_degenerate(solver, oracle):
    """n_y <= 10 (no early exit, unreachable terminals -> -Inf dropped from the frontier), no moves at all,
    minreport 0 (everything reported incl. the 0-move ref line with its trailing tab), maxreport truncation."""
    for seed, (n_x, n_y, dens) in enumerate([(6, 5, 0.5), (10, 9, 0.15), (8, 3, 0.6), (4, 4, 0.0), (14, 1, 0.7)]):
        mat, lx, ly = random_grid(n_x, n_y, dens, 100 + seed)
        aln = to_solver_inputs(mat)
        for kw in (dict(maxdepth=3, maxdup=2, minreport=0.0, maxreport=300),
                   dict(maxdepth=2, maxdup=1, maxwidth=5, minreport=0.5)):
            O = oracle.solve(aln, lx, ly, **kw)
            G = solver.solve(aln, lx, ly, keep_ids=True, **kw)
            assert_search_equal(G, O)


def test_roundtrip_property_full_size(solver):
    """Size-independent property at a bench-sized problem (no oracle): the hidden move chain that generated the
    target must be found, i.e. best == 100.0 within chain_depth moves, and rerunning is deterministic."""
    mat, lx, ly, hidden = sdgrid(64, [6, 5, 5, 4, 4], seed=21, chain_depth=2)
    aln = to_solver_inputs(mat)
    A = solver.solve(aln, lx, ly, maxdepth=2, maxdup=2, minreport=0.98, maxreport=100)
    B = solver.solve(aln, lx, ly, maxdepth=2, maxdup=2, minreport=0.98, maxreport=100)
    assert A.stats.best == 100.0 and A.tsv == B.tsv
    assert A.tsv.startswith("100.0\t")
    assert A.generated == B.generated and A.scored == B.scored


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_search_emulated(oracle, world):
    """Frontier-sharded single search (nw_solve_sharded_ctxs) with the ranks emulated on one GPU: every rank
    must reproduce the single-GPU / oracle result bit for bit (every id, score, edge, counter, TSV), for a beam search
    and an exhaustive one, and the DP work must really be divided (cells are summed over the ranks)."""
    import nahrwhals_b200 as nb
    from nahrwhals_b200.distributed import solve_sharded_local
    from tests.util import assert_search_equal
    solvers = [nb.Solver(0) for _ in range(world)]
    try:
        for gen, flags in [(dict(n_x=40, families=[5, 4, 3, 2], seed=12, chain_depth=3), dict(maxdepth=4, maxdup=2, maxwidth=50)),
                           (dict(n_x=24, families=[4, 3, 3], seed=11, chain_depth=3), dict(maxdepth=3, maxdup=2)),
                           (dict(n_x=150, families=[6, 5, 4, 4, 3], seed=7, chain_depth=4), dict(maxdepth=4, maxdup=2, maxwidth=100))]:
            mat, lx, ly, _ = sdgrid(**gen)
            aln = to_solver_inputs(mat)
            kw = dict(minreport=0.98, maxreport=1000, **flags)
            O = oracle.solve(aln, lx, ly, **kw)
            res = solve_sharded_local(solvers, aln, lx, ly, keep_ids=True, **kw)
            assert len(res) == world
            for R in res:
                assert_search_equal(R, O)
                assert R.stats.score_launches >= 1
    finally:
        for s in solvers:
            s.close()


def test_solve_batch_regions(oracle):
    """Regionfile mode (nw_solve_batch): independent regions on several contexts sharing one GPU give, region by
    region, exactly the TSV of the oracle."""
    import nahrwhals_b200 as nb
    rng = np.random.default_rng(3)
    regions = []
    for k in range(24):
        n_x = int(rng.integers(20, 70))
        fams = [int(rng.integers(2, 5)) for _ in range(int(rng.integers(2, 5)))]
        mat, lx, ly, _ = sdgrid(n_x, fams, seed=500 + k, chain_depth=int(rng.integers(1, 4)), len_tol=0.3)
        regions.append((to_solver_inputs(mat), lx, ly))
    kw = dict(maxdepth=3, maxdup=2, maxwidth=1000, minreport=0.98, maxreport=1000)
    solvers = [nb.Solver(0) for _ in range(4)]
    try:
        res = nb.solve_batch(solvers, regions, **kw)
    finally:
        for s in solvers:
            s.close()
    assert len(res) == len(regions)
    for (aln, lx, ly), R in zip(regions, res):
        O = oracle.solve(aln, lx, ly, full=False, **kw)
        assert R.tsv == O.tsv and R.scored == [int(v) for v in O.scored]


def test_dedup_verifier(solver, oracle):
    """By default every merged candidate is re-materialised and compared element-wise with its owner (include/nwsolve.h,
    "Dedup exactness"): a returned search is exact.  With fingerprints truncated to 8 bits (test-only flag) collisions
    are certain and the search must fail with NW_ERR_COLLISION instead of returning a wrong result."""
    from nahrwhals_b200._lib import NwError
    mat, lx, ly, _ = sdgrid(n_x=64, families=[6, 5, 5, 4], seed=1, chain_depth=3)
    aln = to_solver_inputs(mat)
    kw = dict(maxdepth=3, maxdup=2, minreport=0.98, maxreport=1000)
    O = oracle.solve(aln, lx, ly, **kw)
    G = solver.solve(aln, lx, ly, keep_ids=True, **kw)
    assert_search_equal(G, O)
    with pytest.raises(NwError) as e:
        solver.solve(aln, lx, ly, debug_weak_fp=True, **kw)
    assert e.value.code == -7 and "collision" in str(e.value)
    # without the verifier the truncated fingerprints silently merge different indices (fewer unique ids)
    W = solver.solve(aln, lx, ly, debug_weak_fp=True, verify_dedup=False, **kw)
    assert sum(W.scored) < int(O.scored.sum())
    # a collision is cured by repeating the search with other hash bases: weak fingerprints on the first attempt only
    # -> the verifier rejects it, the second attempt (other bases, full fingerprints) is the oracle's search
    T = solver.solve(aln, lx, ly, keep_ids=True, debug_weak_fp_once=True, **kw)
    assert_search_equal(T, O)
    # the fingerprint-only path (NW_FLAG_NO_VERIFY) gives the same search when nothing collides
    U = solver.solve(aln, lx, ly, keep_ids=True, verify_dedup=False, **kw)
    assert_search_equal(U, O)
    # and the context is still usable afterwards
    G2 = solver.solve(aln, lx, ly, **kw)
    assert G2.tsv == O.tsv


def test_edge_shapes_and_options(solver, oracle):
    """Shapes at the reference's size envelope (x + y grid lines = 250, R/nahrwhals.R:55), bp lengths with gcd 1 near
    the 5 Mb window size, single-block inputs, depth 1, report limits 0 / none, minreport above 1."""
    rng = np.random.default_rng(77)
    # 130 x 120 blocks, realistic sparse structure, lengths without a common factor
    mat, lx, ly, _ = sdgrid(130, [5, 4, 4, 3, 3, 2], seed=9, chain_depth=2, len_tol=0.1)
    lx = lx // 10 + rng.integers(1, 97, lx.shape)          # ~1e3..2e4 bp, gcd 1
    fam_fix = {}
    n_y = mat.shape[0]
    ly = np.array([lx[np.nonzero(mat[k])[0][0]] if np.any(mat[k]) else 1000 + k for k in range(n_y)], dtype=np.int64)
    mat = np.sign(mat) * lx[None, :]
    aln = to_solver_inputs(mat)
    for kw in (dict(maxdepth=2, maxdup=2, maxwidth=200, minreport=0.98, maxreport=1000),
               dict(maxdepth=1, maxdup=3, minreport=0.0, maxreport=None),
               dict(maxdepth=2, maxdup=2, maxwidth=50, minreport=0.5, maxreport=0),
               dict(maxdepth=2, maxdup=2, maxwidth=50, minreport=1.5, maxreport=10)):
        O = oracle.solve(aln, lx, ly, **kw)
        G = solver.solve(aln, lx, ly, keep_ids=True, **kw)
        assert_search_equal(G, O)
    # big bp lengths (5 Mb window, 1 bp resolution): int32 DP range after gcd reduction
    big_lx = np.array([1_234_567, 999_983, 1_500_001, 765_432], dtype=np.int64)
    big = np.array([[1, 0, 0, 0], [0, 1, 0, -1], [0, -1, 1, 0], [0, 0, 0, 1]], dtype=np.int64) * big_lx[None, :]
    big_ly = np.array([1_234_567, 999_983, 1_500_001, 765_432], dtype=np.int64)
    O = oracle.solve(to_solver_inputs(big), big_lx, big_ly, maxdepth=3, maxdup=2, minreport=0.0)
    G = solver.solve(to_solver_inputs(big), big_lx, big_ly, keep_ids=True, maxdepth=3, maxdup=2, minreport=0.0)
    assert_search_equal(G, O)
    # one block: no pair, only the reference line with its trailing tab
    one = np.array([[1]], dtype=np.int8)
    O = oracle.solve(one, [5000], [5000], maxdepth=3, maxdup=2, minreport=0.0)
    G = solver.solve(one, [5000], [5000], keep_ids=True, maxdepth=3, maxdup=2, minreport=0.0)
    assert_search_equal(G, O)
    assert G.tsv == "100.0\t0\t\n"


def test_error_behaviour(solver, tmp_path):
    """Failures are loud and leave no output file (R/juliasolver.R:56-59 only looks for the file)."""
    import nahrwhals_b200 as nb
    from nahrwhals_b200._lib import NwError
    m = tmp_path / "m.tsv"
    l = tmp_path / "l.tsv"
    out = tmp_path / "out.tsv"
    m.write_text("10\t0\n0\t10\n")
    l.write_text("10\t10\n10\t10\t10\n")  # x lengths do not match the matrix
    with pytest.raises(NwError):
        solver.solve_files(m, l, out)
    assert not out.exists() and not (tmp_path / "out.tsv.tmp").exists()
    l.write_text("10\t10\n10.5\t10\n")  # non-integral bp length
    with pytest.raises(NwError):
        solver.solve_files(m, l, out)
    assert not out.exists()
    with pytest.raises(NwError):
        solver.solve_files(tmp_path / "missing.tsv", l, out)
    with pytest.raises(NwError):
        solver.solve(np.ones((2, 2), np.int8), [1, 1], [1, 1], maxdepth=0)
    with pytest.raises(NwError):  # lengths beyond the int32 DP range
        solver.solve(np.ones((2, 2), np.int8), [2**40 + 1, 2**40 + 3], [2**40 + 1, 3], maxdepth=3)
    # the context survives errors
    l.write_text("10\t10\n10\t10\n")
    solver.solve_files(m, l, out, maxdepth=2, minreport=0.0)
    assert out.read_text().startswith("100.0\t0\t\n")
    assert nb.main(["-d", "2", str(tmp_path / "missing.tsv"), str(l), str(tmp_path / "o2.tsv")]) == 1


@pytest.mark.parametrize("per_pass", [1, 5, 64])
def test_solve_many_regions(solver, oracle, per_pass):
    """Regionfile mode through ONE pipeline instance (nw_solve_many): regions of different shapes searched together,
    region by region identical to the oracle -- counts, every id's level/cost/total/score, every edge, the TSV."""
    rng = np.random.default_rng(5)
    regions = []
    for k in range(23):
        n_x = int(rng.integers(8, 70))
        fams = [int(rng.integers(2, 5)) for _ in range(int(rng.integers(1, 5)))]
        while sum(fams) > n_x:
            fams.pop()
        mat, lx, ly, _ = sdgrid(n_x, fams, seed=700 + k, chain_depth=int(rng.integers(0, 4)), len_tol=0.3)
        if k % 7 == 3:  # a region with different length units
            lx, ly = lx // 10 + 7, ly // 10 + 7
            mat = np.sign(mat) * lx[None, :]
        regions.append((to_solver_inputs(mat), lx, ly))
    regions.append((np.array([[1]], dtype=np.int8), np.array([5000]), np.array([5000])))  # no pair at all
    import nahrwhals_b200 as nb
    for kw in (dict(maxdepth=3, maxdup=2, maxwidth=1000, minreport=0.98, maxreport=1000),
               dict(maxdepth=3, maxdup=2, maxwidth=3, minreport=0.5, maxreport=5),
               dict(maxdepth=2, maxdup=1, minreport=0.0)):
        res = nb.solve_many(solver, regions, regions_per_pass=per_pass, keep_ids=True, **kw)
        assert len(res) == len(regions)
        for (aln, lx, ly), G in zip(regions, res):
            O = oracle.solve(aln, lx, ly, **kw)
            assert_search_equal(G, O)


@pytest.mark.gpu
def test_solve_many_several_contexts(solver, oracle):
    """nw_solve_many_ctxs: the passes of a regionfile run handed out to three contexts; results as from one."""
    import nahrwhals_b200 as nb
    rng = np.random.default_rng(11)
    regions = []
    for k in range(29):
        n_x = int(rng.integers(8, 50))
        fams = [int(rng.integers(2, 4)) for _ in range(int(rng.integers(1, 4)))]
        while sum(fams) > n_x:
            fams.pop()
        mat, lx, ly, _ = sdgrid(n_x, fams, seed=900 + k, chain_depth=int(rng.integers(0, 3)), len_tol=0.3)
        regions.append((to_solver_inputs(mat), lx, ly))
    kw = dict(maxdepth=3, maxdup=2, maxwidth=200, minreport=0.9, maxreport=50)
    extra = [nb.Solver(0), nb.Solver(0)]
    try:
        res = nb.solve_many([solver] + extra, regions, regions_per_pass=4, keep_ids=True, **kw)
        bulk = nb.solve_many([solver] + extra, regions, regions_per_pass=8, **kw)  # nw_results_gather path
    finally:
        for s in extra:
            s.close()
    assert len(res) == len(regions)
    for (aln, lx, ly), G in zip(regions, res):
        assert_search_equal(G, oracle.solve(aln, lx, ly, **kw))
    for G, B in zip(res, bulk):
        assert B.tsv == G.tsv and B.moves == G.moves
        for f in ("scores", "n_moves", "move_off", "cost_bp", "total_bp", "index_id"):
            assert np.array_equal(getattr(B, f), getattr(G, f)), f
        assert B.generated == G.generated and B.scored == G.scored and B.stats.n_ids == G.stats.n_ids


@pytest.mark.gpu
def test_files_many_and_cli_batch(solver, oracle, tmp_path):
    """Regionfile run from files (nw_solve_files_many, `nwsolve --batch`): every report byte-identical to the oracle's
    writer; a region whose gridmatrix is missing / malformed fails alone and leaves no output file."""
    import nahrwhals_b200 as nb
    rng = np.random.default_rng(3)
    kw = dict(maxdepth=3, maxdup=2, maxwidth=300, minreport=0.95, maxreport=100)
    triples, want = [], []
    for k in range(9):
        n_x = int(rng.integers(8, 40))
        mat, lx, ly, _ = sdgrid(n_x, [3, 2], seed=40 + k, chain_depth=int(rng.integers(0, 3)), len_tol=0.3)
        a, b, c = (str(tmp_path / ("r%d_%s.tsv" % (k, s))) for s in ("mat", "lens", "out"))
        write_tsvs(mat, lx, ly, a, b)
        triples.append((a, b, c))
        ref = str(tmp_path / ("r%d_ref.tsv" % k))
        assert oracle.solve_files(a, b, ref, **kw) == 0
        want.append(open(ref, "rb").read())
    bad = len(triples)
    triples.append((str(tmp_path / "missing.tsv"), triples[0][1], str(tmp_path / "bad_out.tsv")))
    with open(tmp_path / "ragged.tsv", "w") as f:
        f.write("1000\t0\n0\n")
    triples.append((str(tmp_path / "ragged.tsv"), triples[0][1], str(tmp_path / "bad_out2.tsv")))
    rcs = nb.solve_files_many([solver], triples, regions_per_pass=4, **kw)
    assert all(rc == 0 for rc in rcs[:bad]) and rcs[bad] < 0 and rcs[bad + 1] < 0
    for (a, b, c), w in zip(triples[:bad], want):
        assert open(c, "rb").read() == w
        os.remove(c)
    assert not os.path.exists(triples[bad][2]) and not os.path.exists(triples[bad + 1][2])
    # the same through the command line
    lst = tmp_path / "list.tsv"
    with open(lst, "w") as f:
        for t in triples:
            f.write("\t".join(t) + "\n")
    cli = os.path.join(ROOT, "nahrwhals_b200", "csrc", "nwsolve")
    rc = subprocess.call([cli, "-d", "3", "-u", "2", "-w", "300", "-r", "0.95", "-R", "100", "--batch", str(lst),
                          "--contexts", "2", "--regions-per-pass", "3"], stderr=subprocess.DEVNULL)
    assert rc == 1  # two regions failed
    for (a, b, c), w in zip(triples[:bad], want):
        assert open(c, "rb").read() == w
    assert not os.path.exists(triples[bad][2]) and not os.path.exists(triples[bad + 1][2])
    with open(lst, "w") as f:
        for t in triples[:bad]:
            f.write("\t".join(t) + "\n")
    assert subprocess.call([cli, "-d", "3", "-u", "2", "-w", "300", "-r", "0.95", "-R", "100", "--batch", str(lst)]) == 0


@pytest.mark.gpu
def test_full_size_properties_bench_workload(solver, oracle):
    """BASELINE config 2 at full size (the bench workload: 17.5 M generated / 10.8 M scored candidates), where the
    oracle would need minutes: size-independent properties instead.
      * two independent DP kernels (register-resident savings form vs the literal cost-form transcription) agree on
        the cost / total / score of every one of the 10.8 M ids, and on every provenance edge;
      * the fingerprint dedup of this very search is exact (every merged candidate compared element-wise);
      * both expansion kernels (occurrence masks vs pair testing) generate the same candidates in the same order;
      * the search is deterministic and sharding it over two emulated ranks changes nothing;
      * round trip: every reported trajectory, replayed move by move on the host, yields an index whose score by the
        CPU oracle equals the reported score and whose DP cost equals the reported cost."""
    import sys
    sys.path.insert(0, ROOT)
    from bench import make_workload
    from oracle.nw_oracle_py import apply_move
    from nahrwhals_b200.distributed import solve_sharded_local
    import nahrwhals_b200 as nb
    aln, lx, ly, flags, _ = make_workload("s64", 1)
    A = solver.solve(aln, lx, ly, keep_ids=True, **flags)
    assert sum(A.generated) > 1.5e7 and sum(A.scored) > 1.0e7  # really the full-size workload
    assert A.stats.n_ids == 1 + sum(A.scored)
    B = solver.solve(aln, lx, ly, keep_ids=True, generic_dp=True, **flags)
    assert A.generated == B.generated and A.scored == B.scored and A.cells == B.cells
    for f in ("score", "cost_bp", "total_bp", "level"):
        assert np.array_equal(A.ids[f], B.ids[f]), f
    for f in ("child", "parent", "moves"):
        assert np.array_equal(A.edges[f], B.edges[f]), f
    assert A.tsv == B.tsv
    del B
    C = solver.solve(aln, lx, ly, keep_ids=True, pair_expand=True, **flags)
    assert C.generated == A.generated and C.scored == A.scored and C.tsv == A.tsv
    for f in ("child", "parent", "moves"):
        assert np.array_equal(A.edges[f], C.edges[f]), f
    assert np.array_equal(A.ids["cost_bp"], C.ids["cost_bp"])
    del C
    extra = nb.Solver(0)
    try:
        S = solve_sharded_local([solver, extra], aln, lx, ly, **flags)
    finally:
        extra.close()
    for R in S:
        assert R.tsv == A.tsv and R.generated == A.generated and R.scored == A.scored and R.cells == A.cells
    # round trip of the report through the CPU oracle
    n_x = aln.shape[0]
    root = tuple(range(1, n_x + 1))
    seqs = []
    for mv in A.moves:
        d = root
        for kind, p, q in mv:
            d = apply_move(kind, d, p, q)
        seqs.append(list(d))
    assert len(seqs) == len(A.scores) and len(seqs) >= 100
    osc, ocost, ototal, _ = oracle.score_batch(aln, lx, ly, seqs)
    assert np.array_equal(osc.astype(np.float32), A.scores)
    ocost_i = np.where(np.isnan(ocost), -1.0, ocost).astype(np.int64)
    assert np.array_equal(ocost_i, A.cost_bp) and np.array_equal(ototal.astype(np.int64), A.total_bp)
    assert A.tsv.count("\n") == len(seqs) and np.all(np.diff(A.scores) <= 0)  # report sorted by score desc


@pytest.mark.gpu
def test_forest_equals_single_searches_at_bench_shapes(solver):
    """BASELINE config 4 shapes (20-120 blocks, SD-flanked), too many for the oracle: the forest search
    (256 regions per pass, two contexts) must reproduce nw_solve region by region -- report bytes, trajectory arrays,
    per-level generated / scored / cell counts, id counts and best scores."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from region_batch import FLAGS, make_regions
    import nahrwhals_b200 as nb
    regs = make_regions(300, seed=5)
    extra = nb.Solver(0)
    try:
        many = nb.solve_many([solver, extra], regs, regions_per_pass=256, **FLAGS)
    finally:
        extra.close()
    assert len(many) == len(regs)
    for k, (aln, lx, ly) in enumerate(regs):
        one = solver.solve(aln, lx, ly, **FLAGS)
        M = many[k]
        assert M.tsv == one.tsv, k
        assert M.generated == one.generated and M.scored == one.scored and M.cells == one.cells, k
        assert M.stats.n_ids == one.stats.n_ids and M.stats.best == one.stats.best, k
        for f in ("scores", "n_moves", "cost_bp", "total_bp", "index_id"):
            assert np.array_equal(getattr(M, f), getattr(one, f)), (k, f)
        assert np.array_equal(M.moves3, one.moves3), k


@pytest.mark.gpu
def test_cli_devices_batch_and_sharded_search(oracle, tmp_path):
    """`nwsolve --devices` (the file + command-line boundary of R/juliasolver.R:34-53 over several GPUs, one process, no
    launcher): a regionfile batch spread over the listed devices and ONE search sharded over them must write the
    oracle's bytes.  With a single visible GPU the list names device 0 twice (two ranks / context sets on one device)."""
    import nahrwhals_b200 as nb
    ndev = nb._lib.load().nw_device_count()
    devs = "0,1" if ndev >= 2 else "0,0"
    cli = os.path.join(ROOT, "nahrwhals_b200", "csrc", "nwsolve")
    kw = dict(maxdepth=3, maxdup=2, maxwidth=200, minreport=0.95, maxreport=100)
    flags = ["-d", "3", "-u", "2", "-w", "200", "-r", "0.95", "-R", "100"]
    rng = np.random.default_rng(11)
    triples, want = [], []
    for k in range(12):
        n_x = int(rng.integers(10, 60))
        mat, lx, ly, _ = sdgrid(n_x, [4, 3, 2], seed=70 + k, chain_depth=int(rng.integers(1, 4)), len_tol=0.3)
        a, b, c = (str(tmp_path / ("d%d_%s.tsv" % (k, s))) for s in ("mat", "lens", "out"))
        write_tsvs(mat, lx, ly, a, b)
        triples.append((a, b, c))
        ref = str(tmp_path / ("d%d_ref.tsv" % k))
        assert oracle.solve_files(a, b, ref, **kw) == 0
        want.append(open(ref, "rb").read())
    lst = tmp_path / "dlist.tsv"
    with open(lst, "w") as f:
        for t in triples:
            f.write("\t".join(t) + "\n")
    assert subprocess.call([cli] + flags + ["--batch", str(lst), "--devices", devs, "--contexts", "2",
                                            "--regions-per-pass", "2"]) == 0
    for (a, b, c), w in zip(triples, want):
        assert open(c, "rb").read() == w
        os.remove(c)
    # one search over the listed devices
    for (a, b, c), w in list(zip(triples, want))[:4]:
        assert subprocess.call([cli] + flags + ["--devices", devs, a, b, c]) == 0
        assert open(c, "rb").read() == w
    # a malformed device list is refused before anything runs
    assert subprocess.call([cli] + flags + ["--devices", "0;1", triples[0][0], triples[0][1], str(tmp_path / "x.tsv")],
                           stderr=subprocess.DEVNULL) == 2
    assert not os.path.exists(tmp_path / "x.tsv")
