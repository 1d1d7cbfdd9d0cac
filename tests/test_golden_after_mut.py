"""The reference's second golden picture (inst/extdata/output_png/Fasta_x_Fasta_y_all/Fasta_x_Fasta_y_all-7.png):
the test run's grid AFTER mut_max = 4_12_inv+11_18_del, drawn by make_modified_grid_plot (R/plot_matrix_ggplot.R:261)
via modify_gridmatrix / carry_out_compressed_sv (R/seqmodder.R:10-151).  tests/golden/testrun_grid_after_mut.tsv is its
transcription (sign and colour class per cell).  Everything that applies moves -- the R restatement, the solver
restatement's index map, the product's host position-list model and the CUDA scorer's index map -- must land on it."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import nw_oracle_py as op
from oracle import nw_region_oracle as ro
from tests.test_region_oracle import load_testrun

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CLASSES = (100000, 200000, 300000, 360000)  # legend of the picture: "Blocksize [kbp]"
INDEX_AFTER = [1, 2, 3, 4, -11, -10, -9, -8, -7, -6, 18, 19]


def golden_after():
    return np.loadtxt(os.path.join(GOLD, "testrun_grid_after_mut.tsv"), dtype=np.int64, ndmin=2)


def colour_class(m):
    """sign * smallest legend class that holds |value| (0 stays 0)."""
    a = np.abs(np.asarray(m, dtype=np.float64))
    cls = np.zeros(a.shape, np.int64)
    for c in reversed(CLASSES):
        cls = np.where((a > 0) & (a <= c), c, cls)
    return (np.sign(m) * cls).astype(np.int64)


def test_r_restatement_lands_on_the_picture():
    mat, ly, lx = load_testrun()
    b = ro.Bitl(mat, ly, lx)
    b = ro.carry_out_compressed_sv(b, 4, 12, "inv")
    b = ro.carry_out_compressed_sv(b, 11, 18, "del")
    G = golden_after()
    assert b.m.shape == G.shape == (11, 12)
    assert np.array_equal(colour_class(b.m), G)
    # colnames after the two moves = lengths of the permuted x blocks (no repair applies to this pair of moves)
    assert [int(v) for v in b.cn] == [int(lx[abs(t) - 1]) for t in INDEX_AFTER]


def test_solver_index_map_lands_on_the_picture(testrun):
    """solver.jl never builds the mutated matrix: a child is an index into the original alignment (apply_* :244-278).
    Rendering the index after inv(4,12) and del(11,18) through the original grid must give the picture."""
    aln, lx, ly = testrun
    seq = tuple(range(1, 20))
    seq = op.apply_move(2, seq, 4, 12)
    seq = op.apply_move(1, seq, 11, 18)
    assert list(seq) == INDEX_AFTER
    idx = np.array(seq)
    rendered = (np.sign(idx)[:, None] * aln[np.abs(idx) - 1, :] * lx[np.abs(idx) - 1][:, None]).T  # [n_y, L]
    assert np.array_equal(colour_class(rendered), golden_after())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _host_children(HR, mat, lx, ly):
    n_y, n_x = mat.shape
    g = np.ascontiguousarray(mat, np.int64)
    meta = np.zeros((4096, 4), np.int32)
    blk = np.zeros(1 << 18, np.int16)
    ln = np.zeros(1 << 18, np.int64)
    flags = np.zeros(4, np.int32)
    n = HR.hr_candidates(_p(g), C.c_int(n_x), C.c_int(n_y), _p(np.ascontiguousarray(lx, np.int64)),
                         _p(np.ascontiguousarray(ly, np.int64)), _p(meta), C.c_int64(len(meta)), _p(blk), _p(ln),
                         C.c_int64(len(blk)), _p(flags))
    assert n >= 0 and not flags[1], "the test run is not flipped"
    out, off = {}, 0
    for k in range(n):
        L = int(meta[k, 3])
        out[(int(meta[k, 0]), int(meta[k, 1]), ("del", "dup", "inv")[meta[k, 2]])] = (blk[off:off + L].astype(int), ln[off:off + L].copy())
        off += L
    return out


def test_product_host_model_lands_on_the_picture():
    """The product's position-list model of carry_out_compressed_sv (csrc/nw_region_host.h, through the CPU harness):
    child of 4_12_inv, rendered as the matrix R would hold, then its child 11_18_del."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    subprocess.check_call(["make", "-s", "-C", os.path.join(here, "harness")])
    HR = C.CDLL(os.path.join(here, "harness", "libnw_region_harness.so"))
    HR.hr_candidates.restype = C.c_int64
    mat, ly, lx = load_testrun()
    kids = _host_children(HR, mat.astype(np.int64), lx.astype(np.int64), ly.astype(np.int64))
    bl, le = kids[(4, 12, "inv")]
    m1 = (np.sign(bl)[None, :] * mat[:, np.abs(bl) - 1]).astype(np.int64)
    # R rewrites a cell as sign * (its column's length); the inverted columns keep their own magnitudes
    kids2 = _host_children(HR, m1, le, ly.astype(np.int64))
    bl2, le2 = kids2[(11, 18, "del")]
    m2 = np.sign(bl2)[None, :] * m1[:, np.abs(bl2) - 1]
    assert np.array_equal(colour_class(m2), golden_after())
    assert le2.tolist() == [int(lx[abs(t) - 1]) for t in INDEX_AFTER]


@pytest.mark.gpu
def test_cuda_scorer_agrees_on_the_mutated_grid(solver, testrun):
    """CUDA path: the index [1,2,3,4,-11..-6,18,19] scored on the ORIGINAL grid (index map inside the DP kernel) and the
    identity index scored on the grid of the picture (cells = sign * permuted x length) are the same alignment problem:
    cost 10 000 of 2 470 000 bp = 99.595 (README.md:137 res_max)."""
    aln, lx, ly = testrun
    score, cost, total = solver.score_batch(aln, lx, ly, [INDEX_AFTER])
    assert (int(cost[0]), int(total[0]), float(score[0])) == (10000, 2470000, 99.595)
    G = golden_after()
    aln2 = np.ascontiguousarray(np.sign(G).T.astype(np.int8))
    lx2 = np.array([int(lx[abs(t) - 1]) for t in INDEX_AFTER], np.int64)
    score2, cost2, total2 = solver.score_batch(aln2, lx2, ly, [list(range(1, 13))])
    assert (int(cost2[0]), int(total2[0]), float(score2[0])) == (10000, 2470000, 99.595)
