"""CPU tests of the region steps: the R restatement (oracle/nw_region_oracle.py) against the
reference's published test-run row, and the product's host-side formatter (nw_region_format, no GPU needed) against
that restatement."""
import os

import numpy as np
import pytest

from oracle import nw_region_oracle as ro

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def load_testrun():
    mat = np.loadtxt(os.path.join(GOLD, "testrun_grid_mut.tsv"), dtype=np.float64, ndmin=2)
    lens = open(os.path.join(GOLD, "testrun_grid_mut_lens.tsv")).read().strip().split("\n")
    ly = np.array(lens[0].split("\t"), dtype=np.float64)
    lx = np.array(lens[1].split("\t"), dtype=np.float64)
    return mat, ly, lx


def oracle_solver(oracle):
    def run(aln, lx, ly, d, u, w, r, R):
        return oracle.solve(aln, lx, ly, maxdepth=d, maxdup=u, maxwidth=int(w), minreport=r, maxreport=R).tsv
    return run


def test_readme_row_of_the_test_run(oracle):
    """README.md:135-139 of the reference: res_ref 71.378, res_max 99.595, mut_max 4_12_inv+11_18_del.  The whole
    region path (R depth-1 pre-pass -> solver at depth 2 -> format_julia_output -> merge -> save_to_logfile)
    reproduces that row on the gridmatrix recovered from the reference's own figure."""
    mat, ly, lx = load_testrun()
    gl = np.concatenate([[0], np.cumsum(lx)])
    A = ro.analyse_region(mat, ly, lx, gl, oracle_solver(oracle), depth=2)
    assert A["julia_called"]
    assert (A["res_ref"], A["res_max"], A["mut_max"]) == (71.378, 99.595, "4_12_inv+11_18_del")
    # the depth-1 pre-pass alone: the ref row is scored with the full corridor, children above 90 would be listed
    pre = ro.solve_mutation_depth1(ro.Bitl(mat, ly, lx), 1000)
    assert [r for r in pre["res"] if r["mut1"] == "ref"][0]["eval"] == 71.378
    assert max(r["eval"] for r in pre["res"]) == 82.114
    # depth 3 (the wrapper's default): a perfect three-step chain exists
    A3 = ro.analyse_region(mat, ly, lx, gl, oracle_solver(oracle), depth=3)
    assert (A3["res_ref"], A3["res_max"]) == (71.378, 100.0)
    assert A3["mut_max"].startswith("4_12_inv+11_18_del+")


def test_find_sv_opportunities_order():
    """Row scan, combn order, unique(), inv/del labels in place, dup rows appended, length-1 inversions dropped."""
    m = np.array([[5, 0, 5, -5],
                  [0, 7, 0, 0],
                  [5, 0, -5, 5]], dtype=float)
    b = ro.Bitl(m, [1, 1, 1], [5, 7, 5, 5])
    got = ro.find_sv_opportunities(b)
    assert got == [(1, 3, "del"), (1, 4, "inv"), (1, 3, "inv"), (1, 4, "del"), (1, 3, "dup"), (1, 4, "dup")]


def test_carry_out_colname_repairs():
    """The colname repairs of R/seqmodder.R move lengths between positions; they are part of the semantics."""
    m = np.arange(1, 11, dtype=float).reshape(2, 5)
    b = ro.Bitl(m, [1, 1], [10, 20, 30, 40, 50])
    d = ro.carry_out_compressed_sv(b, 1, 2, "dup")           # :25-28: last colname := second colname
    assert d.cn == [10, 20, 20, 30, 40, 20]
    d = ro.carry_out_compressed_sv(b, 2, 4, "dup")
    assert d.cn == [10, 20, 30, 40, 30, 40, 50]
    d = ro.carry_out_compressed_sv(b, 2, 4, "del")
    assert d.cn == [10, 40, 50] and d.m[0].tolist() == [1, 4, 5]
    i = ro.carry_out_compressed_sv(b, 2, 5, "inv")           # border case 2: flanks inverted too, :85-87 repair
    assert i.m[0].tolist() == [1, -5, -4, -3, -2] and i.cn == [10, 50, 40, 30, 50]
    i = ro.carry_out_compressed_sv(b, 1, 4, "inv")           # border case 1
    assert i.m[0].tolist() == [-4, -3, -2, -1, 5] and i.cn == [40, 30, 20, 10, 50]
    i = ro.carry_out_compressed_sv(b, 2, 4, "inv")           # standard, one column: names reset
    assert i.m[0].tolist() == [1, 2, -3, 4, 5] and i.cn == [10, 20, 30, 40, 50]
    i = ro.carry_out_compressed_sv(b, 1, 3, "inv")           # border case 1 with p2 - p1 == 2: names reset (:83-84)
    assert i.m[0].tolist() == [-3, -2, -1, 4, 5] and i.cn == [10, 20, 30, 40, 50]


def test_turn_mut_max_into_svdf():
    f = ro.turn_mut_max_into_svdf
    assert f("4_12_inv") == [[4.0, 12.0, "inv"]]
    # inversion mirrors later breakpoints strictly inside it; deletion shifts later breakpoints behind its start
    assert f("4_12_inv+11_18_del") == [[4.0, 12.0, "inv"], [5.0, 18.0, "del"]]
    assert f("4_12_inv+11_18_del+7_8_del") == [[4.0, 12.0, "inv"], [5.0, 18.0, "del"], [21.0, 22.0, "del"]]
    assert f("2_5_dup+7_9_inv") == [[2.0, 5.0, "dup"], [4.0, 6.0, "inv"]]


def test_format_julia_output_shapes():
    gl = list(range(0, 2100, 100))
    out = ro.format_julia_output("100.0\t0\t\n", gl, 3)
    assert out["only_ref"] and out["rows"] == [dict(eval=100.0, nmut=0, mut1="ref")]
    txt = "99.5\t2\tmove_inv\t4\t12\tmove_del\t11\t18\n99.5\t1\tmove_del\t3\t5\n98.0\t0\t\n"
    out = ro.format_julia_output(txt, gl, 3)
    rows = out["rows"]
    assert [r["mut1"] for r in rows] == ["3_5_del", "4_12_inv", "ref"]      # fewer steps first inside an eval tie
    assert rows[1]["mut2"] == "11_18_del" and rows[1]["mut2_start"] == 400.0 and rows[1]["mut2_end"] == 1700.0
    assert rows[0]["mut2"] is None and rows[0]["mut1_len"] == 200.0
    s = ro.logfile_summary(rows)
    assert (s["res_max"], s["n_res_max"], s["mut_max"], s["res_ref"]) == (99.5, 2, "3_5_del", 98.0)


def _rows_equal(product_rows, oracle_rows):
    assert len(product_rows) == len(oracle_rows)
    for a, b in zip(product_rows, oracle_rows):
        assert list(a.keys()) == list(b.keys()), (list(a.keys()), list(b.keys()))
        for k in a:
            assert a[k] == b[k], (k, a, b)


@pytest.mark.parametrize("seed", range(6))
def test_product_formatter_matches_oracle(oracle, seed):
    """nw_region_format (C++ in libnwsolve.so, host side) vs the R restatement on real solver reports."""
    import nahrwhals_b200 as nb
    from nahrwhals_b200.sdgrid import sdgrid, to_solver_inputs
    if seed == 0:
        mat, ly, lx = load_testrun()
    else:
        mat, lx, ly, _ = sdgrid(18 + 3 * seed, [3, 3, 2], chain_depth=2 + seed % 2, seed=seed)
    aln = to_solver_inputs(mat)
    gl = np.concatenate([[0], np.cumsum(lx)])[: len(lx) + (seed % 2)]  # odd seeds: one gridline more than blocks
    for depth, minreport in ((1, 0.5), (3, 0.9)):
        tsv = oracle.solve(aln, np.asarray(lx, float), np.asarray(ly, float), maxdepth=depth, maxdup=2, maxwidth=200,
                           minreport=minreport, maxreport=300).tsv
        O = ro.format_julia_output(tsv, gl, depth)
        P = nb.format_julia_output(tsv, gl)
        _rows_equal(P.rows, O["rows"])
        s = ro.logfile_summary(O["rows"])
        assert P.mut_max == s["mut_max"]
        assert (P.res_max, P.n_res_max) == (s["res_max"], s["n_res_max"])
        if not isinstance(s["res_ref"], list):
            assert P.res_ref == s["res_ref"]
    P = nb.format_julia_output("100.0\t0\t\n", gl)
    assert P.rows == [dict(eval=100.0, nmut=0.0, mut1="ref")] and P.mut_max == "ref"
    with pytest.raises(Exception):
        nb.format_julia_output("", gl)


def test_region_golden_fixture(oracle):
    """The committed region tables of the test run (tests/golden/make_region_fixture.py) still come out of the oracle."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mrf", os.path.join(GOLD, "make_region_fixture.py"))
    mrf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mrf)
    mat, ly, lx = load_testrun()
    gl = np.concatenate([[0], np.cumsum(lx)])
    for d in (2, 3):
        A = ro.analyse_region(mat, ly, lx, gl, oracle_solver(oracle), depth=d)
        gold = open(os.path.join(GOLD, "testrun_region_d%d.tsv" % d)).read().split("\n", 1)
        assert gold[1] == mrf.render(A["res"])
        assert "mut_max=%s" % A["mut_max"] in gold[0]
