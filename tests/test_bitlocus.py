"""wrapper_paf_to_bitlocus (the recompression loop): the R restatement on the CPU, and the GPU path
(include/nwbitlocus.h) against it."""
import ctypes as C

import numpy as np
import pytest

from oracle import nw_bitlocus_oracle as bo
from oracle import nw_paf_oracle as po
from tests.test_paf import chunked_paf


def sd_paf(seed, n_blocks=24, chunk=10000):
    """Chunked alignments of a synthetic SD-flanked region: every non-zero cell of an sdgrid matrix is one slope +-1
    alignment, cut into chunk-long pieces as chunked minimap2 reports them."""
    from nahrwhals_b200.sdgrid import sdgrid
    from oracle.nw_grid_oracle import matrix_to_paf
    mat, len_x, len_y, _ = sdgrid(n_blocks, families=(3, 3, 2), chain_depth=1, seed=seed)
    ts, te, qs, qe = matrix_to_paf(mat, len_x, len_y)
    rows = []
    span = float(max(np.sum(len_x), np.sum(len_y)))
    for a in range(len(ts)):
        length = te[a] - ts[a]
        neg = qs[a] > qe[a]
        for k in np.arange(0, length, chunk):
            ln = min(chunk, length - k)
            if neg:
                q0, q1 = qs[a] - k - ln, qs[a] - k      # minimap2 order: qstart < qend
            else:
                q0, q1 = qs[a] + k, qs[a] + k + ln
            rows.append((chunk, q0, q1, -1.0 if neg else 1.0, span, ts[a] + k, ts[a] + k + ln, ln - 3, ln, 60))
    a = np.array(rows, dtype=np.float64).reshape(-1, 10)
    return {k: a[:, i].copy() for i, k in enumerate(po.COLS)}


def test_update_params_and_defaults():
    seq, v = [], 350.0
    for _ in range(8):
        v, c = bo.update_params(v, 123.0)
        assert c == v
        seq.append(v)
    assert seq == [700.0, 1000.0, 2000.0, 4000.0, 8000.0, 10000.0, 20000.0, 40000.0]
    assert bo.update_params(100.0, 100.0)[0] == 200.0 and bo.update_params(50000.0, 1.0)[0] == 100000.0
    # the bundled test run: chr1_partial:1700000-3300000
    assert bo.determine_chunklen(1700000, 3300000) == 10000 and bo.determine_compression(1700000, 3300000) == 10000
    assert bo.determine_compression(0, 50000) == 100 and bo.determine_compression(0, 50001) == 1000
    assert bo.determine_chunklen(0, 9999) == 100 and bo.determine_chunklen(0, 6e6) == 50000
    assert bo.determine_compression(0, 2e10) == 100          # which.min of an all-Inf vector is 1


def test_auto_params_c_abi_matches_restatement():
    """Pure host functions of the C ABI (no GPU needed)."""
    from nahrwhals_b200 import bitlocus
    for length in (1, 9999, 10000, 10001, 49999, 50000, 50001, 499999, 500000, 500001, 1e6, 1600000, 2e6, 2000001, 5e6,
                   5000001, 9e9, 1e10, 2e10):
        got = bitlocus.auto_params(1000.0, 1000.0 + length)
        assert got["chunklen"] == bo.determine_chunklen(1000.0, 1000.0 + length), length
        assert got["minlen"] == got["compression"] == bo.determine_compression(1000.0, 1000.0 + length), length
    p = bitlocus.default_params()
    assert (p.max_n_alns, p.max_size_col_plus_rows, p.max_pairs) == (150, 250, 2000)


def test_is_cluttered_paf():
    p = {k: np.arange(1, 13, dtype=np.float64) for k in po.COLS}
    assert not bo.is_cluttered_paf(p)
    p["qstart"] = np.array([0.0] * 7 + [5.0] * 5)             # >= 5 alignments start at the border AND few distinct starts
    assert bo.is_cluttered_paf(p)
    p["qstart"] = np.array([0.0] * 5 + list(range(1, 8)), dtype=np.float64)   # border hits but 8 distinct starts of 12
    assert not bo.is_cluttered_paf(p)


def test_oracle_loop_branches():
    p = sd_paf(0)
    ok = bo.wrapper_paf_to_bitlocus(p, 1000.0, 1000.0)
    assert ok["status"] == "ok" and ok["passes"] == 1 and ok["n_pairs"] > 0
    assert ok["grid"]["mat"].shape == (len(ok["grid"]["len_y"]), len(ok["grid"]["len_x"]))
    few = bo.wrapper_paf_to_bitlocus(p, 1000.0, 1000.0, max_n_alns=ok["n_alns"] - 1)
    assert few["passes"] > 1 and few["minlen"] > 1000.0 and few["status"] in ("ok", "empty")
    small = bo.wrapper_paf_to_bitlocus(p, 1000.0, 1000.0, max_size_col_plus_rows=len(ok["grid"]["len_x"]) + 2)
    assert small["passes"] > 1
    pairs = bo.wrapper_paf_to_bitlocus(p, 1000.0, 1000.0, max_pairs=ok["n_pairs"] - 1)
    assert pairs["passes"] > 1
    assert bo.wrapper_paf_to_bitlocus(p, 1e7, 1e7)["status"] == "empty"
    assert bo.wrapper_paf_to_bitlocus(p, 1000.0, 1000.0, max_n_alns=0, max_passes=3)["status"] in ("passes", "empty")


def _testrun():
    import os
    import nahrwhals_b200 as nb
    from nahrwhals_b200.sdgrid import matrix_to_chunked_paf
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    _, lx, ly = nb.read_problem(os.path.join(g, "testrun_grid_mut.tsv"), os.path.join(g, "testrun_grid_mut_lens.tsv"))
    mat = np.loadtxt(os.path.join(g, "testrun_grid_mut.tsv"), dtype=np.float64, ndmin=2)
    ap = nb.auto_params(1700000, 3300000)        # the bundled test run: chr1_partial:1700000-3300000
    return mat, lx, ly, ap, matrix_to_chunked_paf(mat, lx, ly, int(ap["chunklen"]))


def test_oracle_testrun_chunks_give_back_the_testrun_gridmatrix():
    """The test run's alignments, chunked as the pipeline chunks them, through the restated wrapper_paf_to_bitlocus with
    the reference's default parameters: 153 chunks -> 6 alignments -> the 11 x 19 gridmatrix of the reference's figure."""
    mat, lx, ly, ap, chunks = _testrun()
    assert ap == dict(chunklen=10000.0, minlen=10000.0, compression=10000.0)
    O = bo.wrapper_paf_to_bitlocus(chunks, ap["minlen"], ap["compression"], ap["chunklen"])
    assert (O["status"], O["passes"], O["n_alns"]) == ("ok", 1, 6)
    assert np.array_equal(O["grid"]["mat"], mat) and np.array_equal(O["grid"]["len_x"], lx)
    assert np.array_equal(O["grid"]["len_y"], ly)


@pytest.mark.gpu
def test_gpu_testrun_chunks_give_back_the_testrun_gridmatrix(solver):
    import nahrwhals_b200 as nb
    mat, lx, ly, ap, chunks = _testrun()
    B = nb.paf_to_bitlocus(solver, [chunks], ap["minlen"], ap["compression"], ap["chunklen"])[0]
    assert (B.status, B.passes, B.n_alns) == ("ok", 1, 6)
    assert np.array_equal(B.grid.mat, mat) and np.array_equal(B.grid.len_x, lx) and np.array_equal(B.grid.len_y, ly)
    R = nb.analyse_region(solver, B.grid.mat, B.grid.len_y, B.grid.len_x, B.grid.gridlines_x, depth=2)
    assert (R.res_ref, R.res_max, R.mut_max) == (71.378, 99.595, "4_12_inv+11_18_del")   # README.md:135-139


def _check(B, O):
    want = {"ok": "ok", "cluttered": "cluttered", "empty": "failed", "passes": "failed"}[O["status"]]
    assert B.status == want, (B.status, O["status"])
    assert B.passes == O["passes"] and B.minlen == O["minlen"] and B.compression == O["compression"]
    assert B.n_alns == O["n_alns"]
    if O["status"] in ("ok", "passes"):
        assert B.n_pairs == O["n_pairs"]
    if O["status"] == "empty":
        assert B.rc == -1
    if O["status"] == "passes":
        assert B.rc == -5
    if O["status"] != "ok":
        assert B.grid is None and B.paf is None
        return
    for k in po.COLS:
        assert np.array_equal(B.paf[k], O["paf"][k]), k
    g = O["grid"]
    assert np.array_equal(B.grid.gridlines_x, g["gridlines_x"]) and np.array_equal(B.grid.gridlines_y, g["gridlines_y"])
    assert np.array_equal(B.grid.mat, g["mat"])
    assert np.array_equal(B.grid.len_x, g["len_x"]) and np.array_equal(B.grid.len_y, g["len_y"])
    assert B.grid.converged == bool(g["converged"])


@pytest.mark.gpu
def test_gpu_bitlocus_matches_oracle(solver):
    """A batch of regions with limits chosen so that every branch of the loop is taken by some region."""
    import nahrwhals_b200 as nb
    pafs = [sd_paf(s, n_blocks=int(12 + 4 * (s % 4))) for s in range(6)]
    rng = np.random.default_rng(5)
    pafs += [chunked_paf(rng, n_segments=int(rng.integers(3, 12)), jitter=int(rng.choice([0, 2, 30])),
                         span=int(rng.choice([300000, 900000])), extra=int(rng.integers(0, 8))) for _ in range(6)]
    clutter = chunked_paf(np.random.default_rng(8), n_segments=8, span=400000)
    clutter["tstart"] = np.where(np.arange(len(clutter["tstart"])) % 2 == 0, 0.0, clutter["tstart"])
    clutter["tend"] = np.maximum(clutter["tend"], clutter["tstart"] + 2000)
    pafs.append(clutter)
    pafs.append({k: np.zeros(0) for k in po.COLS})
    base = [bo.wrapper_paf_to_bitlocus(p, 1000.0, 1000.0) for p in pafs]
    for B, O in zip(nb.paf_to_bitlocus(solver, pafs, 1000.0, 1000.0), base):
        _check(B, O)
    assert {o["status"] for o in base} >= {"ok", "empty"}
    for kw in (dict(max_n_alns=12), dict(max_size_col_plus_rows=20), dict(max_pairs=3), dict(max_n_alns=2, max_passes=3),
               dict(minlen=350.0, compression=350.0, max_size_col_plus_rows=30, max_pairs=10)):
        minlen, compression = kw.pop("minlen", 1000.0), kw.pop("compression", 1000.0)
        want = [bo.wrapper_paf_to_bitlocus(p, minlen, compression, **kw) for p in pafs]
        got = nb.paf_to_bitlocus(solver, pafs, minlen, compression, **kw)
        assert max(o["passes"] for o in want) > 1
        for B, O in zip(got, want):
            _check(B, O)


@pytest.mark.gpu
def test_gpu_bitlocus_chain_to_region_stage(solver):
    """PAF -> bitlocus (with the retry loop) -> region stage, all native."""
    import nahrwhals_b200 as nb
    B = nb.paf_to_bitlocus(solver, [sd_paf(2, n_blocks=20)], 1000.0, 1000.0)[0]
    assert B.status == "ok" and B.grid.mat.shape[0] >= 2
    R = nb.analyse_region(solver, B.grid.mat, B.grid.len_y, B.grid.len_x, B.grid.gridlines_x, depth=2)
    assert R.res_ref is not None and R.res_max >= R.res_ref
