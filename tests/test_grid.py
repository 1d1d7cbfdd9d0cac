"""Gridding: the R restatement on the CPU, and the GPU path (include/nwgrid.h) against it."""
import numpy as np
import pytest

from nahrwhals_b200.sdgrid import random_grid, sdgrid
from oracle import nw_grid_oracle as go


def random_paf(rng, n, span=200, step=10, p_neg=0.4):
    """n slope +-1 alignments with integer coordinates on a coarse lattice (so that projected lines collide a lot)."""
    ts, te, qs, qe = [], [], [], []
    for _ in range(n):
        ln = step * int(rng.integers(1, 8))
        t0 = step * int(rng.integers(0, span // step))
        q0 = step * int(rng.integers(0, span // step))
        ts.append(t0)
        te.append(t0 + ln)
        if rng.random() < p_neg:
            qs.append(q0 + ln)
            qe.append(q0)
        else:
            qs.append(q0)
            qe.append(q0 + ln)
    return [np.array(v, dtype=np.float64) for v in (ts, te, qs, qe)]


def test_oracle_round_trip_and_rules():
    """A gridmatrix turned into its alignments and gridded again is the same matrix (when every block carries an
    alignment); two alignments through one cell with opposite strand cancel to 0; an empty generation adds line 0."""
    hits = 0
    for seed in range(1, 12):
        mat, lx, ly, _ = sdgrid(12 + 3 * seed, [3, 2, 2], chain_depth=2, seed=seed)
        ts, te, qs, qe = go.matrix_to_paf(mat, lx, ly)
        G = go.paf_to_gridmatrix(ts, te, qs, qe)
        assert G["converged"]
        if G["mat"].shape == mat.shape:
            hits += 1
            assert np.array_equal(G["mat"], mat) and np.array_equal(G["len_x"], lx) and np.array_equal(G["len_y"], ly)
    assert hits >= 4
    # crossing alignments: the shared cell gets +z and -z -> 0
    G = go.paf_to_gridmatrix([0, 0], [10, 10], [0, 10], [10, 0])
    assert G["mat"].tolist() == [[0.0]]
    # a single alignment starting at 5: no line is ever projected, so the reference's buffer quirk adds gridline 0
    G = go.paf_to_gridmatrix([5], [15], [5], [15])
    assert G["gridlines_x"] == [0.0, 5.0, 15.0] and G["gridlines_y"] == [0.0, 5.0, 15.0]
    assert G["mat"].tolist() == [[0.0, 0.0], [0.0, 10.0]]
    assert go.remove_duplicates_triple([(1, 1, 5.0), (1, 1, 5.0), (1, 1, -5.0), (2, 1, 3.0)]) == [(1, 1, 0.0), (2, 1, 3.0)]


def _check(G, O):
    assert G is not None
    assert G.gridlines_x.tolist() == O["gridlines_x"] and G.gridlines_y.tolist() == O["gridlines_y"]
    assert G.converged == O["converged"]
    assert np.array_equal(G.mat, O["mat"])
    assert np.array_equal(G.len_x, O["len_x"]) and np.array_equal(G.len_y, O["len_y"])


@pytest.mark.gpu
def test_gpu_grids_match_oracle(solver):
    import nahrwhals_b200 as nb
    rng = np.random.default_rng(7)
    pafs = []
    for seed in range(1, 16):                       # structured: alignments of synthetic gridmatrices
        mat, lx, ly, _ = sdgrid(10 + 4 * seed, [4, 3, 2], chain_depth=1 + seed % 3, seed=seed)
        pafs.append(go.matrix_to_paf(mat, lx, ly))
    for seed in range(6):
        mat, lx, ly = random_grid(8 + 3 * seed, 7 + 2 * seed, 0.2, seed)
        pafs.append(go.matrix_to_paf(mat, lx, ly))
    for k in range(40):                             # unstructured: many projected lines, crossings, quirk cases
        pafs.append(random_paf(rng, int(rng.integers(1, 14)), span=150 + 10 * k))
    pafs.append([np.array([5.0]), np.array([15.0]), np.array([5.0]), np.array([15.0])])
    pafs.append([np.array([0.0, 0.0]), np.array([10.0, 10.0]), np.array([0.0, 10.0]), np.array([10.0, 0.0])])
    # non-unit slopes (the reference enforces slope one upstream, but the grid code handles any slope)
    pafs.append([np.array([0.0, 20.0]), np.array([30.0, 40.0]), np.array([0.0, 45.0]), np.array([15.0, 5.0])])
    grids, rcs = nb.build_grids(solver, pafs)
    n_ok = 0
    for k, p in enumerate(pafs):
        try:
            O = go.paf_to_gridmatrix(*p)
        except (IndexError, RuntimeError, KeyError):
            assert grids[k] is None and rcs[k] == -1, k
            continue
        if max(len(O["gridlines_x"]), len(O["gridlines_y"])) > nb.grid.MAX_LINES:
            assert grids[k] is None and rcs[k] == -5
            continue
        assert rcs[k] == 0, (k, rcs[k])
        _check(grids[k], O)
        n_ok += 1
    assert n_ok >= 55
    # one region at a time gives the same as the batch
    one = nb.paf_to_gridmatrix(solver, *pafs[3])
    assert np.array_equal(one.mat, grids[3].mat) and one.gridlines_x.tolist() == grids[3].gridlines_x.tolist()


@pytest.mark.gpu
def test_gpu_grid_errors(solver):
    import nahrwhals_b200 as nb
    ok = [np.array([0.0]), np.array([10.0]), np.array([0.0]), np.array([10.0])]
    bad = [np.array([0.0]), np.array([0.0]), np.array([0.0]), np.array([10.0])]       # tstart == tend
    many = random_paf(np.random.default_rng(1), nb.grid.MAX_ALNS + 1)
    grids, rcs = nb.build_grids(solver, [ok, bad, many])
    assert rcs == [0, -1, -5] and grids[1] is None and grids[2] is None and grids[0].mat.tolist() == [[10.0]]
    assert nb.build_grids(solver, []) == ([], [])
    with pytest.raises(Exception):
        nb.paf_to_gridmatrix(solver, *bad)


@pytest.mark.gpu
def test_gpu_grid_to_region_stage(solver, oracle):
    """Gridding output feeds the region stage directly: alignments -> gridmatrix -> pre-pass + solver -> mut_max."""
    import nahrwhals_b200 as nb
    from oracle import nw_region_oracle as ro
    from tests.test_region_oracle import oracle_solver
    mat, lx, ly, _ = sdgrid(30, [3, 3, 2], chain_depth=2, seed=21)
    G = nb.paf_to_gridmatrix(solver, *go.matrix_to_paf(mat, lx, ly))
    P = nb.analyse_region(solver, G.mat, G.len_y, G.len_x, G.gridlines_x, depth=3)
    O = ro.analyse_region(G.mat, G.len_y, G.len_x, G.gridlines_x, oracle_solver(oracle), depth=3)
    assert P.mut_max == O["mut_max"] and P.res_max == O["res_max"] and P.res_ref == O["res_ref"]


@pytest.mark.gpu
def test_gpu_grid_many_regions(solver):
    """More regions than one internal pass holds (8192): results do not depend on the pass a region falls into."""
    import nahrwhals_b200 as nb
    rng = np.random.default_rng(3)
    base = [random_paf(rng, int(rng.integers(1, 6)), span=80) for _ in range(50)]
    pafs = [base[k % 50] for k in range(8300)]
    grids, rcs = nb.build_grids(solver, pafs)
    ref, rref = nb.build_grids(solver, base)
    for k in range(0, 8300, 83):
        a, b = grids[k], ref[k % 50]
        assert rcs[k] == rref[k % 50]
        assert (a is None) == (b is None)
        if a is not None:
            assert np.array_equal(a.mat, b.mat) and a.gridlines_x.tolist() == b.gridlines_x.tolist()
    assert all((grids[k] is None) == (ref[k % 50] is None) for k in range(8192, 8300))
