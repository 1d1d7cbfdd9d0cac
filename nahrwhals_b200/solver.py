"""Host-side mirror of the reference solver interface over the C ABI (include/nwsolve.h).

The option names and meanings are solver.jl's (reference: inst/extdata/scripts/scripts_nw_main/solver.jl:756-791):
maxdepth (-d), maxdup (-u), maxwidth (-w, None = nothing), maxoffset (-O, parsed and ignored like the reference),
minreport (-r), maxreport (-R, None = nothing).  `main(argv)` is the drop-in for `julia solver.jl ...`
(R/juliasolver.R:34-53 builds exactly that command line).
"""
import argparse
import ctypes as C
import time
import sys

import numpy as np

from . import _lib

MOVE_NAMES = ("move_dup", "move_del", "move_inv")  # solver.jl:34


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class SolveResult:
    """Trajectories in report order (solver.jl:739-743) + counters."""

    _E = dict(f32=np.zeros(0, np.float32), i32=np.zeros(0, np.int32), i64=np.zeros(0, np.int64),
              m3=np.zeros((0, 3), np.int32))

    def __init__(self):
        E = SolveResult._E
        self._tsv = ""               # str, or bytes decoded on first access (bulk unpack)
        self.scores = E["f32"]
        self.n_moves = E["i32"]
        self.move_off = E["i64"]
        self.moves3 = E["m3"]        # (kind, start, stop) rows, trajectory k owns
        #                              moves3[move_off[k] : move_off[k] + n_moves[k]]
        self.cost_bp = E["i64"]
        self.total_bp = E["i64"]
        self.index_id = E["i64"]
        self.stats = None
        self.ids = None          # keep_ids: dict(score, cost_bp, total_bp, level)
        self.edges = None        # keep_ids: dict(child, parent, moves)

    @property
    def tsv(self):
        """the exact text write_trajectories (solver.jl:747-753) would write"""
        if not isinstance(self._tsv, str):
            self._tsv = self._tsv.tobytes().decode() if hasattr(self._tsv, "tobytes") else bytes(self._tsv).decode()
        return self._tsv

    @tsv.setter
    def tsv(self, v):
        self._tsv = v

    @property
    def moves(self):
        """list of [(kind, start, stop), ...] per trajectory (solver.jl:44-48)"""
        return [[tuple(int(v) for v in row) for row in self.moves3[o:o + n]]
                for o, n in zip(self.move_off, self.n_moves)]

    @property
    def generated(self):
        return [int(v) for v in self.stats.generated[:self.stats.n_levels]]

    @property
    def scored(self):
        return [int(v) for v in self.stats.scored[:self.stats.n_levels]]

    @property
    def cells(self):
        return [int(v) for v in self.stats.cells[:self.stats.n_levels]]


class Solver:
    """One solver context on one GPU (nw_ctx).  Not thread-safe; make one per thread / per rank."""

    def __init__(self, device=-1):
        self.L = _lib.load()
        h = C.c_void_p()
        _lib.check(self.L.nw_ctx_create(int(device), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.nw_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _opts(self, maxdepth=3, maxdup=3, maxwidth=None, minreport=0.95, maxreport=None, keep_ids=False,
              generic_dp=False, no_traj=False, pair_expand=False, verify_dedup=True, debug_weak_fp=False,
              debug_weak_fp_once=False, maxoffset=4, dp32=False):
        o = _lib.NwOpts()
        self.L.nw_default_opts(C.byref(o))
        o.maxdepth, o.maxdup = int(maxdepth), int(maxdup)
        o.maxwidth = -1 if maxwidth is None else int(maxwidth)
        o.minreport = float(minreport)
        o.maxreport = -1 if maxreport is None else int(maxreport)
        o.keep_ids = 1 if keep_ids else 0
        o.flags = ((_lib.FLAG_GENERIC_DP if generic_dp else 0) | (_lib.FLAG_NO_TRAJ if no_traj else 0) |
                   (_lib.FLAG_PAIR_EXPAND if pair_expand else 0) | (0 if verify_dedup else _lib.FLAG_NO_VERIFY) |
                   (_lib.FLAG_DEBUG_WEAK_FP if debug_weak_fp else 0) |
                   (_lib.FLAG_DEBUG_WEAK_FP_ONCE if debug_weak_fp_once else 0) | (_lib.FLAG_DP32 if dp32 else 0))
        return o

    @staticmethod
    def _problem(aln, len_x, len_y):
        aln = np.ascontiguousarray(aln, dtype=np.int8)
        if aln.ndim != 2:
            raise ValueError("aln must be [n_x, n_y]")
        lx = np.ascontiguousarray(len_x, dtype=np.int64)
        ly = np.ascontiguousarray(len_y, dtype=np.int64)
        if lx.shape != (aln.shape[0],) or ly.shape != (aln.shape[1],):
            raise ValueError("length vectors do not match the gridmatrix")
        return aln, lx, ly

    def solve(self, aln, len_x, len_y, **kw):
        """aln_bfs_scored + get_good_trajectories (solver.jl:644-686, :727-745).  aln: int8 [n_x, n_y]."""
        aln, lx, ly = self._problem(aln, len_x, len_y)
        o = self._opts(**kw)
        r = C.c_void_p()
        _lib.check(self.L.nw_solve(self.h, _p(aln), aln.shape[0], aln.shape[1], _p(lx), _p(ly), C.byref(o), C.byref(r)))
        try:
            return self._unpack(r, o)
        finally:
            self.L.nw_result_free(r)

    def _unpack(self, r, o):
        L = self.L
        R = SolveResult()
        st = _lib.NwStats()
        L.nw_result_stats(r, C.byref(st))
        R.stats = st
        data, ln = C.c_void_p(), C.c_int64()
        L.nw_result_tsv(r, C.byref(data), C.byref(ln))
        R.tsv = C.string_at(data, ln.value).decode() if ln.value else ""
        n = L.nw_result_count(r)
        nm = L.nw_result_move_count(r)
        R.scores = np.empty(n, np.float32)
        R.cost_bp = np.empty(n, np.int64)
        R.total_bp = np.empty(n, np.int64)
        R.index_id = np.empty(n, np.int64)
        R.n_moves = np.empty(n, np.int32)
        R.move_off = np.empty(n, np.int64)
        R.moves3 = np.empty((nm, 3), np.int32)
        L.nw_result_arrays(r, _p(R.scores), _p(R.n_moves), _p(R.move_off), _p(R.cost_bp), _p(R.total_bp),
                           _p(R.index_id), _p(R.moves3))
        if o.keep_ids:
            nid = int(st.n_ids)
            ids = dict(score=np.zeros(nid, np.float32), cost_bp=np.zeros(nid, np.int64),
                       total_bp=np.zeros(nid, np.int64), level=np.zeros(nid, np.int32))
            _lib.check(L.nw_result_ids(r, _p(ids["score"]), _p(ids["cost_bp"]), _p(ids["total_bp"]), _p(ids["level"])))
            R.ids = ids
            ne = L.nw_result_edge_count(r)
            ed = dict(child=np.zeros(ne, np.int32), parent=np.zeros(ne, np.int32), moves=np.zeros((ne, 3), np.int32))
            L.nw_result_edges(r, _p(ed["child"]), _p(ed["parent"]), _p(ed["moves"]))
            R.edges = ed
        return R

    def solve_files(self, inmat, inlens, out, **kw):
        """main() of solver.jl (:755-812): TSVs in, ranked trajectory TSV out.  Returns NwStats."""
        o = self._opts(**kw)
        st = _lib.NwStats()
        _lib.check(self.L.nw_solve_files(self.h, str(inmat).encode(), str(inlens).encode(), str(out).encode(),
                                         C.byref(o), C.byref(st)))
        return st

    def score_batch(self, aln, len_x, len_y, seqs, force=False, generic_dp=False, dp32=False):
        """slow_score (solver.jl:329-376) for raw candidate indices.  Returns (score f64, cost_bp, total_bp);
        cost_bp == -1: early exit (score 0.0, no DP), -2: unreachable (score -inf).
        force: 0/False = slow_score; 1/True = the R twin's forcecalc branch (corridor 1.0, no early exit,
        R/evaltools.R:203-205); 2 = the R twin without forcecalc (corridor chosen by n_y, R/evaltools.R:193-201)."""
        aln, lx, ly = self._problem(aln, len_x, len_y)
        n = len(seqs)
        off = np.zeros(n + 1, np.int64)
        for k, s in enumerate(seqs):
            off[k + 1] = off[k] + len(s)
        flat = np.zeros(max(int(off[-1]), 1), np.int16)
        for k, s in enumerate(seqs):
            flat[off[k]:off[k + 1]] = s
        cost = np.zeros(n, np.int64)
        total = np.zeros(n, np.int64)
        score = np.zeros(n, np.float64)
        _lib.check(self.L.nw_score_batch(self.h, _p(aln), aln.shape[0], aln.shape[1], _p(lx), _p(ly), _p(flat), _p(off),
                                         n, int(force), (_lib.FLAG_GENERIC_DP if generic_dp else 0) | (_lib.FLAG_DP32 if dp32 else 0),
                                         _p(cost),
                                         _p(total), _p(score)))
        return score, cost, total

    def bench_peaks(self):
        """Measured roofline denominators: dict of thread-instr/s for VIADDMNMX, IADD3, mixed alu+fma; copy B/s."""
        out = (C.c_double * 4)()
        _lib.check(self.L.nw_bench_peaks(self.h, C.byref(out)))
        return dict(viaddmnmx_ips=out[0], iadd3_ips=out[1], mixed_ips=out[2], copy_Bps=out[3])

    def moveset(self, aln):
        """to_moveset (solver.jl:217-231) -> (duplicate_delete, invert) bool [n_x, n_x]."""
        aln = np.ascontiguousarray(aln, dtype=np.int8)
        n_x, n_y = aln.shape
        dd = np.zeros((n_x, n_x), np.uint8)
        iv = np.zeros((n_x, n_x), np.uint8)
        _lib.check(self.L.nw_moveset(self.h, _p(aln), n_x, n_y, _p(dd), _p(iv)))
        return dd.astype(bool), iv.astype(bool)


def _problems(regions):
    n = len(regions)
    keep = [Solver._problem(*r) for r in regions]
    probs = (_lib.NwProblem * max(n, 1))()
    for k, (a, lx, ly) in enumerate(keep):
        q = probs[k]
        q.aln_sign = a.ctypes.data
        q.n_x, q.n_y = a.shape
        q.len_x = lx.ctypes.data
        q.len_y = ly.ctypes.data
    return keep, probs


def _unpack_many(solver, outs, n, o):
    """All results of a regionfile run through nw_results_sizes / nw_results_gather: two native calls, then views."""
    L = _lib.load()
    if n == 0:
        return []
    if o.keep_ids:  # diagnostics path: per-result accessors
        return [solver._unpack(C.c_void_p(outs[k]), o) for k in range(n)]
    nt = np.empty(n, np.int64)
    nm = np.empty(n, np.int64)
    tl = np.empty(n, np.int64)
    _lib.check(L.nw_results_sizes(outs, n, _p(nt), _p(nm), _p(tl)))
    T, M, B = int(nt.sum()), int(nm.sum()), int(tl.sum())
    scores = np.empty(T, np.float32)
    n_moves = np.empty(T, np.int32)
    move_off = np.empty(T, np.int64)
    cost = np.empty(T, np.int64)
    total = np.empty(T, np.int64)
    idx = np.empty(T, np.int64)
    moves3 = np.empty((M, 3), np.int32)
    tsv = np.empty(max(B, 1), np.uint8)
    stats = (_lib.NwStats * n)()
    _lib.check(L.nw_results_gather(outs, n, _p(scores), _p(n_moves), _p(move_off), _p(cost), _p(total), _p(idx),
                                   _p(moves3), _p(tsv), stats))
    t1 = np.cumsum(nt).tolist()
    m1 = np.cumsum(nm).tolist()
    b1 = np.cumsum(tl).tolist()
    res = []
    t0 = m0 = b0 = 0
    for k in range(n):
        R = SolveResult()
        te, me, be = t1[k], m1[k], b1[k]
        R.scores = scores[t0:te]
        R.n_moves = n_moves[t0:te]
        R.move_off = move_off[t0:te]
        R.cost_bp = cost[t0:te]
        R.total_bp = total[t0:te]
        R.index_id = idx[t0:te]
        R.moves3 = moves3[m0:me]
        R._tsv = tsv[b0:be]  # view into the shared buffer; decoded on first access
        R.stats = stats[k]
        res.append(R)
        t0, m0, b0 = te, me, be
    return res


def solve_batch(solvers, regions, **kw):
    """Regionfile mode: solve independent regions [(aln, len_x, len_y), ...] on several contexts concurrently
    (nw_solve_batch: one host thread + stream per context; contexts may share a GPU).  Returns SolveResults."""
    L = _lib.load()
    n = len(regions)
    keep, probs = _problems(regions)
    ctxs = (C.c_void_p * len(solvers))(*[s.h for s in solvers])
    outs = (C.c_void_p * max(n, 1))()
    rcs = (C.c_int32 * max(n, 1))()
    o = solvers[0]._opts(**kw)
    rc = L.nw_solve_batch(ctxs, len(solvers), probs, n, C.byref(o), outs, rcs)
    res = []
    try:
        _lib.check(rc)
        res = _unpack_many(solvers[0], outs, n, o)
    finally:
        for k in range(n):
            if outs[k]:
                L.nw_result_free(C.c_void_p(outs[k]))
    return res


def solve_many(solver, regions, regions_per_pass=256, timing=None, **kw):
    """Regionfile mode: `regions_per_pass` regions are searched together through one pipeline instance
    (nw_solve_many), so every kernel launch is shared by all of them.  `solver` may be a list of Solvers: the passes
    are then handed out to the contexts dynamically (nw_solve_many_ctxs).  Returns one SolveResult per region;
    `timing` (a dict) receives the wall time of the native call under "native_s"."""
    L = _lib.load()
    solvers = list(solver) if isinstance(solver, (list, tuple)) else [solver]
    solver = solvers[0]
    n = len(regions)
    keep, probs = _problems(regions)
    outs = (C.c_void_p * max(n, 1))()
    o = solver._opts(**kw)
    t0 = time.perf_counter()
    if len(solvers) == 1:
        rc = L.nw_solve_many(solver.h, probs, n, C.byref(o), int(regions_per_pass), outs)
    else:  # passes handed out to several contexts: host work of one pass overlaps device work of another
        hs = (C.c_void_p * len(solvers))(*[s.h for s in solvers])
        rc = L.nw_solve_many_ctxs(hs, len(solvers), probs, n, C.byref(o), int(regions_per_pass), outs)
    if timing is not None:  # wall time of the C-ABI call alone (host buffers in, reports in host memory out)
        timing["native_s"] = time.perf_counter() - t0
    res = []
    try:
        _lib.check(rc)
        res = _unpack_many(solver, outs, n, o)
    finally:
        for k in range(n):
            if outs[k]:
                L.nw_result_free(C.c_void_p(outs[k]))
    return res


def solve_files_many(solvers, triples, regions_per_pass=512, **kw):
    """Regionfile run from files: triples = [(inmat, inlens, out), ...] -> list of per-region return codes
    (nw_solve_files_many: regions that fail leave no output file; the others are written)."""
    L = _lib.load()
    solvers = list(solvers) if isinstance(solvers, (list, tuple)) else [solvers]
    n = len(triples)
    cols = [(C.c_char_p * max(n, 1))(*[str(t[q]).encode() for t in triples]) for q in range(3)]
    hs = (C.c_void_p * len(solvers))(*[s.h for s in solvers])
    rcs = (C.c_int32 * max(n, 1))()
    o = solvers[0]._opts(**kw)
    L.nw_solve_files_many(hs, len(solvers), cols[0], cols[1], cols[2], n, C.byref(o), int(regions_per_pass), rcs)
    return [int(rcs[k]) for k in range(n)]


def read_problem(inmat, inlens):
    """read_tsv + read_length_tsv (solver.jl:101-107, :121-138) -> (aln int8 [n_x,n_y], len_x, len_y)."""
    L = _lib.load()
    a, lx, ly = C.c_void_p(), C.c_void_p(), C.c_void_p()
    nx, ny = C.c_int32(), C.c_int32()
    _lib.check(L.nw_read_problem(str(inmat).encode(), str(inlens).encode(), C.byref(a), C.byref(nx), C.byref(ny),
                                 C.byref(lx), C.byref(ly)))
    try:
        aln = np.ctypeslib.as_array(C.cast(a, C.POINTER(C.c_int8)), (nx.value, ny.value)).copy()
        len_x = np.ctypeslib.as_array(C.cast(lx, C.POINTER(C.c_int64)), (nx.value,)).copy()
        len_y = np.ctypeslib.as_array(C.cast(ly, C.POINTER(C.c_int64)), (ny.value,)).copy()
    finally:
        L.nw_free(a)
        L.nw_free(lx)
        L.nw_free(ly)
    return aln, len_x, len_y


def main(argv=None):
    """Drop-in for `julia solver.jl` (argument table solver.jl:756-791)."""
    ap = argparse.ArgumentParser(prog="nwsolve")
    ap.add_argument("--maxdepth", "-d", type=int, default=3)
    ap.add_argument("--maxdup", "-u", type=int, default=3)
    ap.add_argument("--maxwidth", "-w", type=float, default=None)
    ap.add_argument("--maxoffset", "-O", type=int, default=4)
    ap.add_argument("--minreport", "-r", type=float, default=0.95)
    ap.add_argument("--maxreport", "-R", type=float, default=None)
    ap.add_argument("compressed_alignment")
    ap.add_argument("compressed_lengths")
    ap.add_argument("out_path")
    a = ap.parse_args(argv)
    try:
        s = Solver()
        s.solve_files(a.compressed_alignment, a.compressed_lengths, a.out_path, maxdepth=a.maxdepth, maxdup=a.maxdup,
                      maxwidth=None if a.maxwidth is None else int(a.maxwidth), minreport=a.minreport,
                      maxreport=None if a.maxreport is None else int(a.maxreport))
    except Exception as e:  # non-zero exit, message on stderr, no output file (R/juliasolver.R:56-59)
        print("nwsolve: %s" % e, file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
