"""Host-side mirror of wrapper_paf_to_bitlocus over the C ABI of include/nwbitlocus.h (GPU only, no CPU fallback).

    paf_to_bitlocus(solver, pafs, minlen, compression, ...)  <->  wrapper_paf_to_bitlocus(inpaf, params) for a batch of
                                                                  regions, from the parsed PAF table on
                                                                  (R/tsv_to_bitlocus.R:151-241)
    auto_params(start_x, end_x)                              <->  chunklen / minlen / compression = "default"
                                                                  (R/nahrwhals_coremodule.R:120-128)
A PAF table is a dict of equally long arrays (see paf.py); a result carries the accepted grid (grid.GridResult), the
prepared table and the parameters of the accepted pass.
"""
import ctypes as C

import numpy as np

from . import _lib
from .grid import GridResult, _lib_grid
from .paf import NwPafCols, _cols, _lib_paf, _table

SYMBOLS = ["nw_bitlocus_default_params", "nw_bitlocus_auto_params", "nw_bitlocus_build"]


class NwBitlocusParams(C.Structure):
    _fields_ = [("minlen", C.c_double), ("compression", C.c_double), ("chunklen", C.c_double), ("max_n_alns", C.c_int32),
                ("max_size_col_plus_rows", C.c_int32), ("max_pairs", C.c_int32), ("max_passes", C.c_int32)]


class NwBitlocusInfo(C.Structure):
    _fields_ = [("minlen", C.c_double), ("compression", C.c_double), ("passes", C.c_int32), ("cluttered", C.c_int32),
                ("n_alns", C.c_int32), ("n_pairs", C.c_int32)]


_bound = False


def _lib_bitlocus():
    global _bound
    L = _lib_paf()
    _lib_grid()
    if not _bound:
        vp = C.c_void_p
        L.nw_bitlocus_default_params.argtypes = [C.POINTER(NwBitlocusParams)]
        L.nw_bitlocus_default_params.restype = None
        L.nw_bitlocus_auto_params.argtypes = [C.c_double, C.c_double, C.POINTER(NwBitlocusParams)]
        L.nw_bitlocus_auto_params.restype = None
        L.nw_bitlocus_build.argtypes = [vp, vp, C.c_int64, C.POINTER(NwBitlocusParams), vp, vp, vp, vp]
        _bound = True
    return L


def default_params():
    L = _lib_bitlocus()
    p = NwBitlocusParams()
    L.nw_bitlocus_default_params(C.byref(p))
    return p


def auto_params(start_x, end_x):
    """dict(chunklen, minlen, compression) the reference derives from the region's reference interval."""
    L = _lib_bitlocus()
    p = default_params()
    L.nw_bitlocus_auto_params(float(start_x), float(end_x), C.byref(p))
    return dict(chunklen=p.chunklen, minlen=p.minlen, compression=p.compression)


class Bitlocus:
    """status: 'ok' (grid / paf set), 'cluttered' (is_cluttered_paf: the reference returns NULL), or 'failed' (rc < 0)."""

    def __init__(self, rc, info, grid, paf):
        self.rc = int(rc)
        self.minlen, self.compression = info.minlen, info.compression
        self.passes, self.n_alns, self.n_pairs = info.passes, info.n_alns, info.n_pairs
        self.grid, self.paf = grid, paf
        self.status = "failed" if rc != 0 else ("cluttered" if info.cluttered else "ok")


def paf_to_bitlocus(solver, pafs, minlen, compression, chunklen=10000.0, max_n_alns=150, max_size_col_plus_rows=250,
                    max_pairs=2000, max_passes=64, timing=None):
    """pafs: list of PAF tables.  Returns a list of Bitlocus, one per region."""
    import time
    L = _lib_bitlocus()
    n = len(pafs)
    keep = []
    arr = (NwPafCols * max(n, 1))()
    for k, p in enumerate(pafs):
        c, a = _cols(p)
        keep.append(a)
        arr[k] = c
    prm = NwBitlocusParams(float(minlen), float(compression), float(chunklen), int(max_n_alns),
                           int(max_size_col_plus_rows), int(max_pairs), int(max_passes))
    grids = (C.c_void_p * max(n, 1))()
    tabs = (C.c_void_p * max(n, 1))()
    infos = (NwBitlocusInfo * max(n, 1))()
    rcs = np.zeros(max(n, 1), np.int32)
    t0 = time.perf_counter()
    _lib.check(L.nw_bitlocus_build(solver.h, arr, n, C.byref(prm), grids, tabs, infos, rcs.ctypes.data_as(C.c_void_p)))
    if timing is not None:
        timing["native_s"] = time.perf_counter() - t0
    out = []
    for k in range(n):
        g = GridResult(L, C.c_void_p(grids[k])) if grids[k] else None
        t = _table(L, C.c_void_p(tabs[k])) if tabs[k] else None
        out.append(Bitlocus(rcs[k], infos[k], g, t))
    return out
