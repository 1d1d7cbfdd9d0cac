"""Host-side mirror of the reference's gridding step over the C ABI of include/nwgrid.h (GPU only, no CPU fallback).

    build_grids(solver, pafs)  <->  make_xy_grid_fast + bressiwrap + gridlist_to_gridmatrix for a batch of regions
                                    (R/tsv_to_bitlocus.R:186-215, :290-384; R/bresenham.R; R/gridlist_to_gridmatrix.R)
An alignment set is (tstart, tend, qstart, qend): the columns of the reference's paf data frame.
"""
import ctypes as C

import numpy as np

from . import _lib

SYMBOLS = ["nw_grid_build", "nw_grid_dims", "nw_grid_lines", "nw_grid_matrix", "nw_grid_free"]
MAX_LINES = 1024
MAX_ALNS = 512


class NwPaf(C.Structure):
    _fields_ = [("tstart", C.c_void_p), ("tend", C.c_void_p), ("qstart", C.c_void_p), ("qend", C.c_void_p),
                ("n", C.c_int64)]


_bound = False


def _lib_grid():
    global _bound
    L = _lib.load()
    if not _bound:
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        L.nw_grid_build.argtypes = [vp, vp, i64, vp, vp]
        L.nw_grid_dims.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
        L.nw_grid_lines.argtypes = [vp, vp, vp]
        L.nw_grid_matrix.argtypes = [vp, vp, vp, vp]
        L.nw_grid_free.argtypes = [vp]
        _bound = True
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class GridResult:
    """gridlines_x / gridlines_y (gxy[[1]], gxy[[2]]), mat[n_y, n_x] with len_y = row.names and len_x = colnames."""

    def __init__(self, L, h):
        try:
            nx, ny, cv, gen = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
            _lib.check(L.nw_grid_dims(h, C.byref(nx), C.byref(ny), C.byref(cv), C.byref(gen)))
            self.gridlines_x = np.zeros(nx.value + 1, np.float64)
            self.gridlines_y = np.zeros(ny.value + 1, np.float64)
            _lib.check(L.nw_grid_lines(h, _p(self.gridlines_x), _p(self.gridlines_y)))
            self.mat = np.zeros((ny.value, nx.value), np.float64)
            self.len_x = np.zeros(nx.value, np.float64)
            self.len_y = np.zeros(ny.value, np.float64)
            _lib.check(L.nw_grid_matrix(h, _p(self.mat), _p(self.len_x), _p(self.len_y)))
            self.converged = bool(cv.value)
            self.generations = int(gen.value)
        finally:
            L.nw_grid_free(h)


def build_grids(solver, pafs, timing=None):
    """pafs: list of (tstart, tend, qstart, qend) arrays.  Returns (grids, rcs): grids[k] is a GridResult or None,
    rcs[k] the status of region k (0, NW_ERR_RANGE = -5: too many gridlines / alignments, NW_ERR_ARG = -1)."""
    import time
    L = _lib_grid()
    n = len(pafs)
    keep = []
    arr = (NwPaf * max(n, 1))()
    for k, p in enumerate(pafs):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in p]
        if len({len(v) for v in a}) != 1:
            raise ValueError("tstart, tend, qstart, qend must have one length")
        keep.append(a)
        arr[k] = NwPaf(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), len(a[0]))
    outs = (C.c_void_p * max(n, 1))()
    rcs = np.zeros(max(n, 1), np.int32)
    t0 = time.perf_counter()
    _lib.check(L.nw_grid_build(solver.h, arr, n, outs, _p(rcs)))
    if timing is not None:
        timing["native_s"] = time.perf_counter() - t0
    return [GridResult(L, C.c_void_p(outs[k])) if outs[k] else None for k in range(n)], rcs[:n].tolist()


def paf_to_gridmatrix(solver, tstart, tend, qstart, qend):
    """One region; raises if the grid could not be built."""
    g, rc = build_grids(solver, [(tstart, tend, qstart, qend)])
    if g[0] is None:
        raise _lib.NwError(rc[0], "gridding failed (%s)" % ("too many gridlines or alignments" if rc[0] == -5
                                                             else "degenerate alignment or inconsistent grid"))
    return g[0]
