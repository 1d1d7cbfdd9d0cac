"""Regionfile mode from the alignments on, all native: PAF tables -> wrapper_paf_to_bitlocus -> solver stage -> rows of
nahrwhals_res.tsv (the part of wrapper_aln_and_analyse, R/wrapper_aln_and_analyse.R:130-203, that follows minimap2;
plots are not produced).  Every step is a batched C-ABI call of libnwsolve.so (include/nwbitlocus.h, nwregion.h); this
module only routes the regions between them the way the reference's wrapper does:

    grid_xy <- wrapper_condense_paf(params, outlinks)              -> bitlocus.paf_to_bitlocus (per parameter group)
    is.null(grid_xy): res_empty, write_results  (:132-140)          -> the empty row, cluttered_boundaries TRUE
    gridmatrix, is_cluttered, solve_mutation, solver, merge (:160-191) -> region.analyse_regions
    write_results -> save_to_logfile (:201)                          -> RegionResult.logfile_row / append_logfile
"""
import ctypes as C

import numpy as np

from . import _lib
from .bitlocus import NwBitlocusParams, _lib_bitlocus, auto_params, paf_to_bitlocus
from .paf import NwPafCols, _cols
from .region import (NwLogMeta, NwRegionSummary, _lib_region, _opts, analyse_regions, append_logfile)

SYMBOLS = ["nw_regions_run", "nw_regions_run_ctxs", "nw_regions_run_files"]


def _empty_row(meta, depth, cluttered):
    """res_empty <- data.frame(eval = 0, mut1 = 'ref') annotated with NA positions, through save_to_logfile."""
    L = _lib_region()
    S = NwRegionSummary()
    S.res_ref = S.res_max = 0.0
    S.n_res_max = 1
    S.search_depth = int(depth)
    S.cluttered = 1 if cluttered else 0
    for k in range(15):
        S.maxres[k] = float("nan")
    enc = lambda v: None if v is None else str(v).encode()  # noqa: E731
    lg = lambda v: -1 if v is None else int(bool(v))  # noqa: E731
    m = NwLogMeta(enc(meta["chr"]), enc(meta["start"]), enc(meta["end"]), enc(meta["samplename"]), enc(meta.get("y_seqname")),
                  enc(meta.get("y_start")), enc(meta.get("y_end")), enc(meta.get("xpad", "1")), float("nan"),
                  lg(meta.get("exceeds_x", False)), lg(meta.get("exceeds_y", False)), lg(meta.get("grid_inconsistency", False)))
    buf = C.create_string_buffer(4096)
    n = C.c_int64()
    _lib.check(L.nw_region_logrow(C.byref(S), b"ref", C.byref(m), buf, 4096, C.byref(n)))
    return buf.value.decode()


def run_regions(solver, pafs, metas, depth=3, maxdup=2, init_width=1000, minreport=0.98, minlen=None, compression=None,
                chunklen=None, max_n_alns=150, max_size_col_plus_rows=250, logfile=None, timing=None):
    """pafs[k]: the parsed PAF table of region k (dict of columns, see paf.py); metas[k]: dict(chr, start, end, samplename
    [, y_seqname, y_start, y_end, xpad, exceeds_x, exceeds_y]) with chr / start / end as R would print them.
    minlen / compression / chunklen None = the reference's "default" (determine_compression / determine_chunklen of the
    region's interval).  Returns (rows, details): rows[k] is the region's line of nahrwhals_res.tsv, or None for a region
    the reference would have dropped with an error; details[k] = (Bitlocus, RegionResult or None).  logfile: rows are
    appended there in region order."""
    import time
    t0 = time.perf_counter()
    n = len(pafs)
    prm = []
    for m in metas:
        ap = auto_params(float(m["start"]), float(m["end"]))
        prm.append((float(minlen if minlen is not None else ap["minlen"]),
                    float(compression if compression is not None else ap["compression"]),
                    float(chunklen if chunklen is not None else ap["chunklen"])))
    bit = [None] * n
    for key in sorted(set(prm)):
        idx = [k for k in range(n) if prm[k] == key]
        for k, b in zip(idx, paf_to_bitlocus(solver, [pafs[k] for k in idx], key[0], key[1], key[2], max_n_alns=max_n_alns,
                                             max_size_col_plus_rows=max_size_col_plus_rows)):
            bit[k] = b
    t1 = time.perf_counter()
    res = [None] * n
    for comp in sorted({prm[k][1] for k in range(n) if bit[k].status == "ok"}):
        idx = [k for k in range(n) if bit[k].status == "ok" and prm[k][1] == comp]
        regs = [(bit[k].grid.mat, bit[k].grid.len_y, bit[k].grid.len_x, bit[k].grid.gridlines_x) for k in idx]
        out, rcs = analyse_regions(solver, regs, depth=depth, maxdup=maxdup, init_width=init_width, minreport=minreport,
                                   compression=comp, isolate=True)
        for k, r in zip(idx, out):
            res[k] = r
    t2 = time.perf_counter()
    rows = []
    for k in range(n):
        m = metas[k]
        if bit[k].status == "cluttered":
            rows.append(_empty_row(m, depth, True))
        elif bit[k].status == "ok" and res[k] is not None:
            rows.append(res[k].logfile_row(m["chr"], m["start"], m["end"], m["samplename"], m.get("y_seqname"),
                                           m.get("y_start"), m.get("y_end"), m.get("xpad", "1"), None,
                                           m.get("exceeds_x", False), m.get("exceeds_y", False), not bit[k].grid.converged))
        else:
            rows.append(None)
    if logfile is not None:
        for r in rows:
            if r is not None:
                append_logfile(logfile, r)
    if timing is not None:
        timing.update(bitlocus_s=t1 - t0, stage_s=t2 - t1, rows_s=time.perf_counter() - t2)
    return rows, list(zip(bit, res))


def _meta(m, keep):
    enc = lambda v: None if v is None else str(v).encode()  # noqa: E731
    lg = lambda v: -1 if v is None else int(bool(v))  # noqa: E731
    comp = m.get("compression")
    f = [enc(m["chr"]), enc(m["start"]), enc(m["end"]), enc(m["samplename"]), enc(m.get("y_seqname")), enc(m.get("y_start")),
         enc(m.get("y_end")), enc(m.get("xpad", "1"))]
    keep.append(f)
    return NwLogMeta(*f, float("nan") if comp is None else float(comp), lg(m.get("exceeds_x", False)),
                     lg(m.get("exceeds_y", False)), 0)


def run_regions_native(solver, pafs, metas, minlen, compression, chunklen=10000.0, depth=3, maxdup=2, init_width=1000,
                       minreport=0.98, max_n_alns=150, max_size_col_plus_rows=250, logfile=None, timing=None,
                       regions_per_batch=512):
    """run_regions as ONE call of the C ABI (nw_regions_run, include/nwpipeline.h): all regions share minlen / compression /
    chunklen.  `solver` may be a list of Solvers (contexts, on one GPU or several) that share the batches
    (nw_regions_run_ctxs).  Returns (rows, status): rows[k] the region's line of nahrwhals_res.tsv or None, status[k]
    0 / 1 (cluttered PAF: the empty row) / negative (dropped)."""
    import time
    L = _lib_bitlocus()
    _lib_region()
    L.nw_regions_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(NwBitlocusParams), C.c_void_p,
                                 C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_void_p, C.c_void_p]
    L.nw_regions_run_ctxs.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(NwBitlocusParams),
                                      C.c_void_p, C.c_int32, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_void_p,
                                      C.c_void_p]
    L.nw_free.argtypes = [C.c_void_p]
    n = len(pafs)
    keep = []
    cols = (NwPafCols * max(n, 1))()
    ms = (NwLogMeta * max(n, 1))()
    for k in range(n):
        c, a = _cols(pafs[k])
        keep.append(a)
        cols[k] = c
        ms[k] = _meta(metas[k], keep)
    bp = NwBitlocusParams(float(minlen), float(compression), float(chunklen), int(max_n_alns), int(max_size_col_plus_rows),
                          2000, 64)
    ro = _opts(L, depth, maxdup, init_width, minreport, compression)
    text, tlen = C.c_void_p(), C.c_int64()
    off = np.zeros(n + 1, np.int64)
    st = np.zeros(max(n, 1), np.int32)
    t0 = time.perf_counter()
    lf = None if logfile is None else str(logfile).encode()
    if isinstance(solver, (list, tuple)):
        hs = (C.c_void_p * len(solver))(*[s_.h for s_ in solver])
        _lib.check(L.nw_regions_run_ctxs(hs, len(solver), cols, ms, n, C.byref(bp), C.byref(ro), int(regions_per_batch), lf,
                                         C.byref(text), C.byref(tlen), off.ctypes.data_as(C.c_void_p),
                                         st.ctypes.data_as(C.c_void_p)))
    else:
        _lib.check(L.nw_regions_run(solver.h, cols, ms, n, C.byref(bp), C.byref(ro), lf, C.byref(text), C.byref(tlen),
                                    off.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p)))
    if timing is not None:
        timing["native_s"] = time.perf_counter() - t0
    try:
        blob = C.string_at(text, tlen.value).decode()
    finally:
        L.nw_free(text)
    rows = [blob[off[k]:off[k + 1]].rstrip("\n") if st[k] >= 0 else None for k in range(n)]
    return rows, st[:n].tolist()


def run_region_files(solver, paf_paths, metas, minlen, compression, chunklen=10000.0, depth=3, maxdup=2, init_width=1000,
                     minreport=0.98, max_n_alns=150, max_size_col_plus_rows=250, logfile=None, regions_per_batch=512):
    """run_regions_native with every region's alignments read from its PAF file (nw_regions_run_files): the regionfile mode
    of the reference from the files minimap2 / the melt step left behind to the rows of nahrwhals_res.tsv."""
    L = _lib_bitlocus()
    _lib_region()
    L.nw_regions_run_files.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(NwBitlocusParams),
                                       C.c_void_p, C.c_int32, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_void_p,
                                       C.c_void_p]
    L.nw_free.argtypes = [C.c_void_p]
    solvers = list(solver) if isinstance(solver, (list, tuple)) else [solver]
    n = len(paf_paths)
    keep = []
    paths = (C.c_char_p * max(n, 1))(*[str(p).encode() for p in paf_paths])
    ms = (NwLogMeta * max(n, 1))()
    for k in range(n):
        ms[k] = _meta(metas[k], keep)
    bp = NwBitlocusParams(float(minlen), float(compression), float(chunklen), int(max_n_alns), int(max_size_col_plus_rows),
                          2000, 64)
    ro = _opts(L, depth, maxdup, init_width, minreport, compression)
    hs = (C.c_void_p * len(solvers))(*[s_.h for s_ in solvers])
    text, tlen = C.c_void_p(), C.c_int64()
    off = np.zeros(n + 1, np.int64)
    st = np.zeros(max(n, 1), np.int32)
    _lib.check(L.nw_regions_run_files(hs, len(solvers), paths, ms, n, C.byref(bp), C.byref(ro), int(regions_per_batch),
                                      None if logfile is None else str(logfile).encode(), C.byref(text), C.byref(tlen),
                                      off.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p)))
    try:
        blob = C.string_at(text, tlen.value).decode()
    finally:
        L.nw_free(text)
    rows = [blob[off[k]:off[k + 1]].rstrip("\n") if st[k] >= 0 else None for k in range(n)]
    return rows, st[:n].tolist()
