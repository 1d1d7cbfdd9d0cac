"""Host-side mirror of the reference's PAF preparation over the C ABI of include/nwpaf.h (GPU only, no CPU fallback).

    compress_paf(solver, paf, second_run, chunklen, compression)  <-> compress_paf_fnct(inpaf_df = paf, save_outpaf = F, ...)
    load_and_prep_paf(solver, paf, minlen, compression, chunklen) <-> load_and_prep_paf_for_gridplot_transform from the
                                                                      parsed table on (R/tsv_to_bitlocus.R:38-82)
A PAF table is a dict of equally long arrays: qlen, qstart, qend, strand (+1 / -1), tlen, tstart, tend, nmatch, alen, mapq.
"""
import ctypes as C

import numpy as np

from . import _lib

SYMBOLS = ["nw_paf_compress", "nw_paf_prepare", "nw_paf_prepare_many", "nw_paf_compress_file", "nw_paf_compress_file_by_qname",
           "nw_paf_read_file",
           "nw_paf_rows", "nw_paf_columns", "nw_paf_stats", "nw_paf_free"]
COLS = ("qlen", "qstart", "qend", "strand", "tlen", "tstart", "tend", "nmatch", "alen", "mapq")


class NwPafCols(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in COLS] + [("n", C.c_int64)]


_bound = False


def _lib_paf():
    global _bound
    L = _lib.load()
    if not _bound:
        vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        L.nw_paf_compress.argtypes = [vp, C.POINTER(NwPafCols), i32, f64, f64, i32, C.POINTER(vp)]
        L.nw_paf_prepare.argtypes = [vp, C.POINTER(NwPafCols), f64, f64, f64, C.POINTER(vp)]
        L.nw_paf_prepare_many.argtypes = [vp, vp, i64, f64, f64, f64, vp, vp]
        L.nw_paf_compress_file.argtypes = [vp, C.c_char_p, C.c_char_p, f64, i32, C.POINTER(i64), C.POINTER(i64)]
        L.nw_paf_compress_file_by_qname.argtypes = [vp, C.c_char_p, C.c_char_p, f64, i32, C.POINTER(i64), C.POINTER(i64)]
        L.nw_paf_read_file.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.nw_paf_rows.restype = i64
        L.nw_paf_rows.argtypes = [vp]
        L.nw_paf_columns.argtypes = [vp] * 11
        L.nw_paf_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i32)]
        L.nw_paf_free.argtypes = [vp]
        _bound = True
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _cols(paf):
    arrs = {k: np.ascontiguousarray(paf[k], dtype=np.float64) for k in COLS}
    n = len(arrs["qstart"])
    if any(len(v) != n for v in arrs.values()):
        raise ValueError("PAF columns must have one length")
    c = NwPafCols(*[_p(arrs[k]) for k in COLS], n)
    return c, arrs


def _table(L, h):
    try:
        n = L.nw_paf_rows(h)
        out = {k: np.zeros(n, np.float64) for k in COLS}
        _lib.check(L.nw_paf_columns(h, *[_p(out[k]) for k in COLS]))
        pf, gl = C.c_int64(), C.c_int32()
        _lib.check(L.nw_paf_stats(h, C.byref(pf), C.byref(gl)))
        out["_pairs_found"], out["_gpu_launches"] = int(pf.value), int(gl.value)
        return out
    finally:
        L.nw_paf_free(h)


def compress_paf(solver, paf, second_run=False, chunklen=10000.0, compression=1000.0, n_quadrants_per_axis=None):
    L = _lib_paf()
    c, keep = _cols(paf)
    h = C.c_void_p()
    _lib.check(L.nw_paf_compress(solver.h, C.byref(c), 1 if second_run else 0, float(chunklen), float(compression),
                                 int(n_quadrants_per_axis or 0), C.byref(h)))
    return _table(L, h)


def load_and_prep_paf(solver, paf, minlen, compression, chunklen=10000.0):
    L = _lib_paf()
    c, keep = _cols(paf)
    h = C.c_void_p()
    _lib.check(L.nw_paf_prepare(solver.h, C.byref(c), float(minlen), float(compression), float(chunklen), C.byref(h)))
    return _table(L, h)


def load_and_prep_pafs(solver, pafs, minlen, compression, chunklen=10000.0, timing=None):
    """load_and_prep_paf for a batch of tables (one kernel launch for the all-pairs tests of all of them).
    Returns (tables, rcs); tables[k] is None for a malformed table."""
    import time
    L = _lib_paf()
    n = len(pafs)
    keep = []
    arr = (NwPafCols * max(n, 1))()
    for k, p in enumerate(pafs):
        c, a = _cols(p)
        keep.append(a)
        arr[k] = c
    outs = (C.c_void_p * max(n, 1))()
    rcs = np.zeros(max(n, 1), np.int32)
    t0 = time.perf_counter()
    _lib.check(L.nw_paf_prepare_many(solver.h, arr, n, float(minlen), float(compression), float(chunklen), outs, _p(rcs)))
    if timing is not None:
        timing["native_s"] = time.perf_counter() - t0
    return [_table(L, C.c_void_p(outs[k])) if outs[k] else None for k in range(n)], rcs[:n].tolist()


def compress_paf_file(solver, inpaf_link, outpaf_link, chunklen=10000.0, n_quadrants_per_axis=None):
    """compress_paf_fnct(inpaf_link, outpaf_link, inparam_chunklen = chunklen) (R/seqbuilder_functions.R:107): PAF file in,
    "molten" PAF file out.  Returns (rows read after unique(), rows written)."""
    L = _lib_paf()
    a, b = C.c_int64(), C.c_int64()
    _lib.check(L.nw_paf_compress_file(solver.h, str(inpaf_link).encode(), str(outpaf_link).encode(), float(chunklen),
                                      int(n_quadrants_per_axis or 0), C.byref(a), C.byref(b)))
    return int(a.value), int(b.value)


def read_paf(path):
    """read.table(inpaf) of load_and_prep_paf_for_gridplot_transform: the numeric columns of a PAF file as a table."""
    L = _lib_paf()
    h = C.c_void_p()
    _lib.check(L.nw_paf_read_file(str(path).encode(), C.byref(h)))
    t = _table(L, h)
    return {k: t[k] for k in COLS}


def compress_paf_file_by_qname(solver, inpaf_link, outpaf_link, chunklen=10000.0, n_quadrants_per_axis=10):
    """The all-vs-all melt of whole-genome mode (R/make_whole_genome_list.R:80-106): every query sequence of the chunked PAF
    melted on its own and appended to outpaf_link.  Returns (rows read, rows written)."""
    L = _lib_paf()
    a, b = C.c_int64(), C.c_int64()
    _lib.check(L.nw_paf_compress_file_by_qname(solver.h, str(inpaf_link).encode(), str(outpaf_link).encode(), float(chunklen),
                                               int(n_quadrants_per_axis or 0), C.byref(a), C.byref(b)))
    return int(a.value), int(b.value)
