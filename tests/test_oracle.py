"""CPU tests: pin the oracle on the reference's own known answers and against the independent Python restatement."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from nahrwhals_b200.sdgrid import random_grid, sdgrid, to_solver_inputs
from oracle import nw_oracle_py as op

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_readme_known_answers(oracle, testrun):
    """README.md:135-139 of the reference: res_ref 71.378, res_max 99.595, mut_max 4_12_inv+11_18_del."""
    aln, lx, ly = testrun
    # res_ref comes from the R twin's forcecalc scorer on the unmodified index (R/wrapper_aln_and_analyse.R:176)
    sc, cost, total, cells = oracle.score_batch(aln, lx, ly, [list(range(1, 20))], force=True)
    assert (sc[0], cost[0], total[0]) == (71.378, 810000.0, 2830000.0)
    # res_max = score of the index after 4_12_inv then 11_18_del (matches ...-7.png of the reference)
    idx = [1, 2, 3, 4, -11, -10, -9, -8, -7, -6, 18, 19]
    sc, cost, total, cells = oracle.score_batch(aln, lx, ly, [idx])
    assert (sc[0], cost[0], total[0]) == (99.595, 10000.0, 2470000.0)
    # Julia's own scorer early-exits on the 19-long ref index (|19-11| > 5.5)
    sc, cost, _, _ = oracle.score_batch(aln, lx, ly, [list(range(1, 20))])
    assert sc[0] == 0.0 and np.isnan(cost[0])
    # and the move chain 4_12_inv + 11_18_del really produces that index
    seq = tuple(range(1, 20))
    seq = op.apply_move(2, seq, 4, 12)
    seq = op.apply_move(1, seq, 11, 18)
    assert list(seq) == idx


def test_testrun_search_counts(oracle, testrun):
    aln, lx, ly = testrun
    R = oracle.solve(aln, lx, ly, maxdepth=3, maxdup=2, maxwidth=1000, minreport=0.98, maxreport=1000)
    assert list(R.generated.cumsum()) == [17, 308, 4792]
    assert R.n_ids == 2420 and list(R.scored.cumsum()) == [17, 208, 2419]
    assert R.best == 100.0
    first = R.tsv.split("\n")[0].split("\t")
    assert first[0] == "100.0" and first[1] == "3"
    # the README answer is among the reported trajectories with its README score
    assert "99.595\t2\tmove_inv\t4\t12\tmove_del\t11\t18\n" in R.tsv


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_cpp_vs_python_restatement(oracle, seed):
    rng = np.random.default_rng(seed)
    if seed == 1:
        mat, lx, ly, _ = sdgrid(14, [3, 2, 2], seed=seed, chain_depth=2)
    elif seed == 2:
        mat, lx, ly = random_grid(9, 8, 0.2, seed)
    else:
        mat, lx, ly, _ = sdgrid(22, [4, 3], seed=seed, chain_depth=3)
    aln = to_solver_inputs(mat)
    for kw in (dict(maxdepth=3, maxdup=2, maxwidth=None), dict(maxdepth=3, maxdup=1, maxwidth=7, maxreport=50)):
        R = oracle.solve(aln, lx, ly, minreport=0.9, **kw)
        P = op.solve(aln.tolist(), lx, ly, minreport=0.9, **kw)
        assert R.n_ids == len(P["seqs"])
        assert np.array_equal(R.score, np.array(P["scores"], np.float32))
        pc = np.array(P["costs"], np.float64)
        assert np.array_equal(np.isnan(R.cost), np.isnan(pc))
        assert np.array_equal(np.nan_to_num(R.cost, nan=-1.0), np.nan_to_num(pc, nan=-1.0))
        assert [tuple(R.flat[R.off[k]:R.off[k + 1]]) for k in range(R.n_ids)] == P["seqs"]
        assert R.tsv == P["tsv"]


def test_cli_matches_library(oracle, testrun, tmp_path):
    g = os.path.join(ROOT, "tests", "golden")
    out = tmp_path / "res.tsv"
    subprocess.check_call([os.path.join(ROOT, "oracle", "nw_oracle"), "-d", "3", "-u", "2", "-w", "1000", "-r", "0.98",
                           "-R", "1000", os.path.join(g, "testrun_grid_mut.tsv"),
                           os.path.join(g, "testrun_grid_mut_lens.tsv"), str(out)])
    aln, lx, ly = testrun
    R = oracle.solve(aln, lx, ly, maxdepth=3, maxdup=2, maxwidth=1000, minreport=0.98, maxreport=1000)
    assert out.read_text() == R.tsv
    assert out.read_text() == open(os.path.join(g, "testrun_oracle_res.tsv")).read()


def test_float32_printing(oracle):
    assert oracle.f32_string(100.0) == "100.0"
    assert oracle.f32_string(99.595) == "99.595"
    assert oracle.f32_string(0.0) == "0.0"
    assert oracle.f32_string(99.5) == "99.5"
    assert oracle.f32_string(0.001) == "0.001"
    assert op.f32_str(np.float32(71.378)) == "71.378"
