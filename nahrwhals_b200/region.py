"""Host-side mirror of the reference's per-region steps either side of the solver call, over the C ABI of
include/nwregion.h (no CPU fallback; the scoring runs on the GPU).

Names follow the reference:
    solve_mutation_depth1(...)  <-> solve_mutation(bitlocus, maxdepth = 1, ...)   R/explore_mutation_space_v2.R:14-112
    format_julia_output(...)    <-> format_julia_output(juliares_path, gridlines_x, depth)     R/juliasolver.R:56-120
    analyse_region(...)         <-> R/wrapper_aln_and_analyse.R:176-191 + save_to_logfile's res_ref / res_max /
                                    n_res_max / mut_max (R/save_logfile.R:17-51)
A gridmatrix is passed as R holds it: mat[n_y, n_x] of signed bp, len_y = row.names, len_x = colnames.
"""
import ctypes as C

import numpy as np

from . import _lib


class NwRegionOpts(C.Structure):
    _fields_ = [("depth", C.c_int32), ("maxdup", C.c_int32), ("init_width", C.c_double), ("minreport", C.c_double),
                ("compression", C.c_double), ("flags", C.c_int32), ("reserved", C.c_int32)]


class NwRegionSummary(C.Structure):
    _fields_ = [("res_ref", C.c_double), ("res_max", C.c_double), ("n_res_max", C.c_int64),
                ("mut_simulated", C.c_int64), ("mut_tested", C.c_int64), ("search_depth", C.c_int32),
                ("solver_called", C.c_int32), ("flipped", C.c_int32), ("flip_unsure", C.c_int32),
                ("cluttered", C.c_int32), ("n_rows", C.c_int32), ("n_mut", C.c_int32), ("width_reduced", C.c_int32),
                ("maxres", C.c_double * 15), ("prepass_candidates", C.c_int64), ("prepass_dp", C.c_int64),
                ("gpu_launches", C.c_int32), ("reserved", C.c_int32)]


class NwLogMeta(C.Structure):
    _fields_ = [("chr", C.c_char_p), ("start", C.c_char_p), ("end", C.c_char_p), ("samplename", C.c_char_p),
                ("y_seqname", C.c_char_p), ("y_start", C.c_char_p), ("y_end", C.c_char_p), ("xpad", C.c_char_p),
                ("compression", C.c_double), ("exceeds_x", C.c_int32), ("exceeds_y", C.c_int32),
                ("grid_inconsistency", C.c_int32)]


SYMBOLS = ["nw_region_default_opts", "nw_region_prepass", "nw_region_format", "nw_region_analyse", "nw_table_summary",
           "nw_table_mut_max", "nw_table_tsv", "nw_table_rows", "nw_table_free", "nw_table_candidates", "nw_region_analyse_many",
           "nw_region_analyse_many_ctxs", "nw_region_logrow_header", "nw_region_logrow", "nw_region_log_append"]


class NwRegionProblem(C.Structure):
    _fields_ = [("grid", C.c_void_p), ("n_x", C.c_int32), ("n_y", C.c_int32), ("len_x", C.c_void_p),
                ("len_y", C.c_void_p), ("gridlines_x", C.c_void_p), ("n_gridlines", C.c_int32)]


class NwCandidate(C.Structure):
    _fields_ = [("p1", C.c_int32), ("p2", C.c_int32), ("sv", C.c_int32), ("length", C.c_int32),
                ("est_rows", C.c_int64), ("est_cols", C.c_int64), ("cost_bp", C.c_int64), ("total_bp", C.c_int64),
                ("score", C.c_double)]

MAXRES_NAMES = ["mut%d_%s" % (i, s) for i in (1, 2, 3) for s in ("start", "end", "pos_pm", "len", "len_pm")]

_bound = False


def _lib_region():
    global _bound
    L = _lib.load()
    if not _bound:
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        L.nw_region_default_opts.argtypes = [C.POINTER(NwRegionOpts)]
        L.nw_region_prepass.argtypes = [vp, vp, i32, i32, vp, vp, C.POINTER(NwRegionOpts), C.POINTER(vp),
                                        C.POINTER(NwRegionSummary)]
        L.nw_region_format.argtypes = [C.c_char_p, i64, vp, i32, C.POINTER(vp)]
        L.nw_region_analyse.argtypes = [vp, vp, i32, i32, vp, vp, vp, i32, C.POINTER(NwRegionOpts), C.POINTER(vp),
                                        C.POINTER(NwRegionSummary)]
        L.nw_region_analyse_many.argtypes = [vp, vp, i64, C.POINTER(NwRegionOpts), vp, vp, vp]
        L.nw_region_analyse_many_ctxs.argtypes = [vp, i32, vp, i64, C.POINTER(NwRegionOpts), i32, vp, vp, vp]
        L.nw_table_summary.argtypes = [vp, C.POINTER(NwRegionSummary)]
        L.nw_table_mut_max.restype = C.c_char_p
        L.nw_table_mut_max.argtypes = [vp]
        L.nw_table_tsv.argtypes = [vp, C.POINTER(vp), C.POINTER(i64)]
        L.nw_table_rows.restype = i32
        L.nw_table_rows.argtypes = [vp]
        L.nw_table_free.argtypes = [vp]
        L.nw_table_candidates.restype = i64
        L.nw_table_candidates.argtypes = [vp, C.POINTER(C.POINTER(NwCandidate))]
        L.nw_region_logrow_header.restype = C.c_char_p
        L.nw_region_logrow.argtypes = [C.POINTER(NwRegionSummary), C.c_char_p, C.POINTER(NwLogMeta), C.c_char_p, i64,
                                       C.POINTER(i64)]
        L.nw_region_log_append.argtypes = [C.c_char_p, C.c_char_p]
        _bound = True
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RegionResult:
    """The `res` table of the reference (rows of dicts, NA = None) with the logfile summary."""

    def __init__(self, L, handle, summary=None):
        try:
            data, n = C.c_void_p(), C.c_int64()
            _lib.check(L.nw_table_tsv(handle, C.byref(data), C.byref(n)))
            self.tsv = C.string_at(data, n.value).decode()
            S = NwRegionSummary()
            _lib.check(L.nw_table_summary(handle, C.byref(S)))
            self.mut_max = L.nw_table_mut_max(handle).decode()
            cp = C.POINTER(NwCandidate)()
            nc = L.nw_table_candidates(handle, C.byref(cp))
            # diagnostics: (p1, p2, sv, length, est_rows, est_cols, cost_bp, total_bp, score) per depth-1 child
            self.candidates = [(c.p1, c.p2, ("del", "dup", "inv")[c.sv], c.length, c.est_rows, c.est_cols, c.cost_bp,
                                c.total_bp, c.score) for c in cp[:nc]]
        finally:
            L.nw_table_free(handle)
        if summary is not None:
            S = summary
        lines = self.tsv.split("\n")
        self.columns = lines[0].split("\t")
        self.rows = []
        for ln in lines[1:]:
            if not ln:
                continue
            row = {}
            for k, v in zip(self.columns, ln.split("\t")):
                if v == "NA":
                    row[k] = None
                elif k.startswith("mut") and "_" not in k:
                    row[k] = v
                else:
                    row[k] = float(v)
            self.rows.append(row)
        nan = lambda v: None if v != v else v  # noqa: E731
        self.res_ref, self.res_max, self.n_res_max = nan(S.res_ref), nan(S.res_max), int(S.n_res_max)
        self.maxres = {k: nan(v) for k, v in zip(MAXRES_NAMES, S.maxres)}
        self.log = dict(depth=int(S.search_depth), mut_simulated=int(S.mut_simulated), mut_tested=int(S.mut_tested),
                        flip_unsure=bool(S.flip_unsure), cluttered_boundaries=bool(S.cluttered))
        self.solver_called = bool(S.solver_called)
        self.flipped = bool(S.flipped)
        self.width_reduced = bool(S.width_reduced)
        self.prepass_candidates = int(S.prepass_candidates)
        self.prepass_dp = int(S.prepass_dp)
        self.gpu_launches = int(S.gpu_launches)
        self._summary = S

    def logfile_row(self, chr, start, end, samplename, y_seqname=None, y_start=None, y_end=None, xpad="1", compression=None,
                    exceeds_x=False, exceeds_y=False, grid_inconsistency=False):
        """The row save_to_logfile appends to nahrwhals_res.tsv for this region (R/save_logfile.R:112-232).  chr / start /
        end / xpad are the CHARACTER values log_collection holds (pass them as R would print them)."""
        L = _lib_region()
        enc = lambda v: None if v is None else str(v).encode()  # noqa: E731
        lg = lambda v: -1 if v is None else int(bool(v))  # noqa: E731
        m = NwLogMeta(enc(chr), enc(start), enc(end), enc(samplename), enc(y_seqname), enc(y_start), enc(y_end), enc(xpad),
                      float("nan") if compression is None else float(compression), lg(exceeds_x), lg(exceeds_y),
                      lg(grid_inconsistency))
        buf = C.create_string_buffer(4096)
        n = C.c_int64()
        _lib.check(L.nw_region_logrow(C.byref(self._summary), self.mut_max.encode(), C.byref(m), buf, 4096, C.byref(n)))
        return buf.value.decode()


def _grid(mat, len_y, len_x):
    m = np.asarray(mat)
    if m.ndim != 2:
        raise ValueError("mat must be [n_y, n_x]")
    g = np.ascontiguousarray(np.rint(m), dtype=np.int64)
    if not np.array_equal(g, m):
        raise ValueError("gridmatrix values must be integers (bp)")
    ly = np.ascontiguousarray(len_y, dtype=np.int64)
    lx = np.ascontiguousarray(len_x, dtype=np.int64)
    if ly.shape != (g.shape[0],) or lx.shape != (g.shape[1],):
        raise ValueError("length vectors do not match the gridmatrix")
    return g, ly, lx


def _opts(L, depth=3, maxdup=2, init_width=1000, minreport=0.98, compression=1000, generic_dp=False):
    o = NwRegionOpts()
    L.nw_region_default_opts(C.byref(o))
    o.depth, o.maxdup = int(depth), int(maxdup)
    o.init_width, o.minreport, o.compression = float(init_width), float(minreport), float(compression)
    o.flags = _lib.FLAG_GENERIC_DP if generic_dp else 0
    return o


def solve_mutation_depth1(solver, mat, len_y, len_x, compression=1000, **kw):
    """solve_mutation(bitlocus, maxdepth = 1, compression = compression) on `solver`'s GPU."""
    L = _lib_region()
    g, ly, lx = _grid(mat, len_y, len_x)
    o = _opts(L, compression=compression, **kw)
    h, S = C.c_void_p(), NwRegionSummary()
    _lib.check(L.nw_region_prepass(solver.h, _p(g), g.shape[1], g.shape[0], _p(lx), _p(ly), C.byref(o), C.byref(h),
                                   C.byref(S)))
    return RegionResult(L, h, S)


def format_julia_output(tsv, gridlines_x):
    """format_julia_output on the text of a solver report (host only)."""
    L = _lib_region()
    b = tsv.encode() if isinstance(tsv, str) else bytes(tsv)
    gl = np.ascontiguousarray(gridlines_x, dtype=np.float64)
    h = C.c_void_p()
    _lib.check(L.nw_region_format(b, len(b), _p(gl), len(gl), C.byref(h)))
    return RegionResult(L, h)


def analyse_region(solver, mat, len_y, len_x, gridlines_x, depth=3, maxdup=2, init_width=1000, minreport=0.98,
                   compression=1000, **kw):
    """The solver stage of one region as R/wrapper_aln_and_analyse.R:176-191 runs it, with the logfile summary."""
    L = _lib_region()
    g, ly, lx = _grid(mat, len_y, len_x)
    gl = np.ascontiguousarray(gridlines_x, dtype=np.float64)
    o = _opts(L, depth, maxdup, init_width, minreport, compression, **kw)
    h, S = C.c_void_p(), NwRegionSummary()
    _lib.check(L.nw_region_analyse(solver.h, _p(g), g.shape[1], g.shape[0], _p(lx), _p(ly), _p(gl), len(gl),
                                   C.byref(o), C.byref(h), C.byref(S)))
    return RegionResult(L, h, S)


def analyse_regions(solver, regions, depth=3, maxdup=2, init_width=1000, minreport=0.98, compression=1000, timing=None,
                    regions_per_batch=1024, isolate=False, **kw):
    """Regionfile mode: the solver stage of many regions in one call (nw_region_analyse_many).  `regions` is a list of
    (mat[n_y, n_x], len_y, len_x, gridlines_x); `solver` may be a list of Solvers (contexts) that share the batches.
    Returns one RegionResult per region, identical to analyse_region.  isolate=True: a region that cannot be analysed
    fails alone (None in its place) and (results, rcs) is returned -- the reference drops such regions in a tryCatch."""
    import time
    L = _lib_region()
    n = len(regions)
    keep = []
    probs = (NwRegionProblem * max(n, 1))()
    for k, (mat, len_y, len_x, gridlines_x) in enumerate(regions):
        g, ly, lx = _grid(mat, len_y, len_x)
        gl = np.ascontiguousarray(gridlines_x, dtype=np.float64)
        keep.append((g, ly, lx, gl))
        probs[k] = NwRegionProblem(_p(g), g.shape[1], g.shape[0], _p(lx), _p(ly), _p(gl), len(gl))
    o = _opts(L, depth, maxdup, init_width, minreport, compression, **kw)
    outs = (C.c_void_p * max(n, 1))()
    summ = (NwRegionSummary * max(n, 1))()
    rcs = np.zeros(max(n, 1), np.int32)
    rcs_p = _p(rcs) if isolate else None
    t0 = time.perf_counter()
    if isinstance(solver, (list, tuple)):  # several contexts: batches handed out dynamically (nw_region_analyse_many_ctxs)
        hs = (C.c_void_p * len(solver))(*[s_.h for s_ in solver])
        _lib.check(L.nw_region_analyse_many_ctxs(hs, len(solver), probs, n, C.byref(o), int(regions_per_batch), outs, summ,
                                                 rcs_p))
    else:
        _lib.check(L.nw_region_analyse_many(solver.h, probs, n, C.byref(o), outs, summ, rcs_p))
    if timing is not None:
        timing["native_s"] = time.perf_counter() - t0
    res = [RegionResult(L, C.c_void_p(outs[k]), summ[k]) if outs[k] else None for k in range(n)]
    if isolate:
        return res, rcs[:n].tolist()
    return res


def logfile_header():
    return _lib_region().nw_region_logrow_header().decode()


def append_logfile(logfile, row):
    """Appends a logfile_row() to nahrwhals_res.tsv under an exclusive lock (header first if the file is empty)."""
    _lib.check(_lib_region().nw_region_log_append(str(logfile).encode(), row.encode()))
