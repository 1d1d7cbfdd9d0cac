"""ctypes loader for the C++ oracle (oracle/nw_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class NwoOpts(C.Structure):
    _fields_ = [("maxdepth", C.c_int32), ("maxdup", C.c_int32), ("maxwidth", C.c_int64),
                ("minreport", C.c_double), ("maxreport", C.c_int64), ("max_scored", C.c_int64)]


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "libnw_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.nwo_solve.restype = C.c_void_p
        L.nwo_solve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(NwoOpts), C.c_int]
        L.nwo_solve_files.restype = C.c_int
        L.nwo_solve_files.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(NwoOpts)]
        L.nwo_free.argtypes = [C.c_void_p]
        for nm in ("nwo_n_ids", "nwo_index_total_len", "nwo_n_edges", "nwo_n_trajs", "nwo_tsv_len"):
            getattr(L, nm).restype = C.c_int64
            getattr(L, nm).argtypes = [C.c_void_p]
        for nm in ("nwo_best", "nwo_seconds"):
            getattr(L, nm).restype = C.c_double
            getattr(L, nm).argtypes = [C.c_void_p]
        L.nwo_n_levels.restype = C.c_int
        L.nwo_n_levels.argtypes = [C.c_void_p]
        L.nwo_level_stats.argtypes = [C.c_void_p] * 6
        L.nwo_id_arrays.argtypes = [C.c_void_p] * 6
        L.nwo_indices.argtypes = [C.c_void_p] * 3
        L.nwo_edges.argtypes = [C.c_void_p] * 4
        L.nwo_tsv.restype = C.c_void_p
        L.nwo_tsv.argtypes = [C.c_void_p]
        L.nwo_score_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nwo_f32_string.restype = C.c_int
        L.nwo_f32_string.argtypes = [C.c_float, C.c_char_p, C.c_int]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def make_opts(maxdepth=3, maxdup=3, maxwidth=None, minreport=0.95, maxreport=None, max_scored=None):
    return NwoOpts(maxdepth, maxdup, -1 if maxwidth is None else int(maxwidth), float(minreport),
                   -1 if maxreport is None else int(maxreport), -1 if max_scored is None else int(max_scored))


class OracleResult:
    pass


def solve(aln, len_x, len_y, want_trajs=True, full=True, **kw):
    """aln: int8 [n_x, n_y] in {-1,0,1}.  Returns OracleResult with per-id arrays, edges, tsv, stats."""
    L = lib()
    aln = np.ascontiguousarray(aln, dtype=np.int8)
    n_x, n_y = aln.shape
    lx = np.ascontiguousarray(len_x, dtype=np.float64)
    ly = np.ascontiguousarray(len_y, dtype=np.float64)
    o = make_opts(**kw)
    h = L.nwo_solve(_p(aln), n_x, n_y, _p(lx), _p(ly), C.byref(o), 1 if want_trajs else 0)
    try:
        R = OracleResult()
        n = L.nwo_n_ids(h)
        R.n_ids = n
        R.best = L.nwo_best(h)
        R.seconds = L.nwo_seconds(h)
        nl = L.nwo_n_levels(h)
        R.generated = np.zeros(nl, np.int64)
        R.scored = np.zeros(nl, np.int64)
        R.cells = np.zeros(nl, np.int64)
        R.frontier = np.zeros(nl, np.int64)
        R.level_seconds = np.zeros(nl, np.float64)
        L.nwo_level_stats(h, _p(R.generated), _p(R.scored), _p(R.cells), _p(R.frontier), _p(R.level_seconds))
        if full:
            R.score = np.zeros(n, np.float32)
            R.cost = np.zeros(n, np.float64)
            R.total = np.zeros(n, np.float64)
            R.level = np.zeros(n, np.int32)
            R.length = np.zeros(n, np.int32)
            L.nwo_id_arrays(h, _p(R.score), _p(R.cost), _p(R.total), _p(R.level), _p(R.length))
            tl = L.nwo_index_total_len(h)
            R.flat = np.zeros(tl, np.int16)
            R.off = np.zeros(n + 1, np.int64)
            L.nwo_indices(h, _p(R.flat), _p(R.off))
            ne = L.nwo_n_edges(h)
            R.edge_off = np.zeros(n + 1, np.int64)
            R.edge_parent = np.zeros(ne, np.int64)
            R.edge_move = np.zeros((ne, 3), np.int32)
            L.nwo_edges(h, _p(R.edge_off), _p(R.edge_parent), _p(R.edge_move))
        R.n_trajs = L.nwo_n_trajs(h)
        R.tsv = C.string_at(L.nwo_tsv(h), L.nwo_tsv_len(h)).decode()
        return R
    finally:
        L.nwo_free(h)


def solve_files(inmat, inlens, out, **kw):
    o = make_opts(**kw)
    return lib().nwo_solve_files(inmat.encode(), inlens.encode(), out.encode(), C.byref(o))


def score_batch(aln, len_x, len_y, seqs, force=False):
    """seqs: list of int sequences.  force: 0 slow_score, 1 / True the R twin's forcecalc branch, 2 the R twin's own
    corridor (chosen by n_y, no early exit).  Returns (score f64, cost f64 [NaN early exit / inf], total f64, cells)."""
    L = lib()
    aln = np.ascontiguousarray(aln, dtype=np.int8)
    n_x, n_y = aln.shape
    lx = np.ascontiguousarray(len_x, dtype=np.float64)
    ly = np.ascontiguousarray(len_y, dtype=np.float64)
    off = np.zeros(len(seqs) + 1, np.int64)
    for k, s in enumerate(seqs):
        off[k + 1] = off[k] + len(s)
    flat = np.zeros(max(int(off[-1]), 1), np.int16)
    for k, s in enumerate(seqs):
        flat[off[k]:off[k + 1]] = s
    n = len(seqs)
    score = np.zeros(n, np.float64)
    cost = np.zeros(n, np.float64)
    total = np.zeros(n, np.float64)
    cells = np.zeros(n, np.int64)
    L.nwo_score_batch(_p(aln), n_x, n_y, _p(lx), _p(ly), _p(flat), _p(off), n, int(force), _p(score), _p(cost),
                      _p(total), _p(cells))
    return score, cost, total, cells


def f32_string(v):
    buf = C.create_string_buffer(64)
    lib().nwo_f32_string(C.c_float(float(v)), buf, 64)
    return buf.value.decode()
