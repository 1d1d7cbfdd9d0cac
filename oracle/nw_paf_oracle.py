"""CPU restatement of the reference's PAF preparation upstream of the gridding.

TEST INFRASTRUCTURE ONLY (imported by tests/ and bench.py's cpu_baseline leg, never by nahrwhals_b200/).

    load_and_prep_paf_for_gridplot_transform   R/tsv_to_bitlocus.R:15-110   (strand swap, compress, minlen filter,
                                               rounding to the compression grid, enforce_slope_one :258-275)
    compress_paf_fnct                          R/compress_paf.R:173-406    (inpaf_df path, save_outpaf = F)
    merge_paf_entries_intraloop                R/merge_paf_entries_intraloop.R:14-79
    merge_rows                                 R/compress_paf.R:13-138

A PAF table is a dict of equally long numpy arrays: qlen, qstart, qend, strand (+1 / -1), tlen, tstart, tend, nmatch,
alen, mapq (the qname / tname strings do not influence any number downstream and are not carried).
PARITY UNPINNED: R is not installed here.  R semantics relied upon: order() is stable; which(, arr.ind = T) walks
column-major; unique() / duplicated() keep first occurrences; dplyr::top_n(1, wt) keeps every row tied for the
maximum, in the original row order; round(x, 0) is half-even; seq(1, r, length.out = n) is 1 + k (r - 1) / (n - 1).
"""
import numpy as np

COLS = ("qlen", "qstart", "qend", "strand", "tlen", "tstart", "tend", "nmatch", "alen", "mapq")


def take(p, idx):
    return {k: np.asarray(v)[idx] for k, v in p.items()}


def swap_minus(p):
    """transform(paf, qend = ifelse(strand == '-', qstart, qend), qstart = ifelse(strand == '-', qend, qstart))"""
    q = dict(p)
    neg = p["strand"] < 0
    q["qend"] = np.where(neg, p["qstart"], p["qend"])
    q["qstart"] = np.where(neg, p["qend"], p["qstart"])
    return q


def merge_paf_entries_intraloop(p, names, second_run, chunklen, compression):
    """names: the R row names of the subset (1-based row numbers of the full table).  -> list of (row, col)."""
    if second_run:
        tol = min(0.5 * compression, float(np.min(p["qlen"])) * 0.9)
    else:
        tol = 0.05 * chunklen
    cond = (np.abs(p["tend"][:, None] - p["tstart"][None, :]) < tol) & \
           (np.abs(p["qend"][:, None] - p["qstart"][None, :]) < tol) & (p["strand"][:, None] == p["strand"][None, :])
    cols, rows = np.nonzero(cond.T)                       # column-major: col outer, row inner
    pairs = list(zip(rows.tolist(), cols.tolist()))
    if tol > 100:
        pairs = [(r, c) for r, c in pairs if p["alen"][r] > tol]
        pairs = [(r, c) for r, c in pairs if p["alen"][c] > tol]
    return [(int(names[r]), int(names[c])) for r, c in pairs]


def merge_rows(p, pairs):
    """R/compress_paf.R:13-138 (reduced_qnames irrelevant: names are not carried).  pairs: list of (row, col), 1-based."""
    qs, qe = p["qstart"].astype(np.float64).copy(), p["qend"].astype(np.float64).copy()
    ts, te = p["tstart"].astype(np.float64).copy(), p["tend"].astype(np.float64).copy()
    nm, al = p["nmatch"].astype(np.float64).copy(), p["alen"].astype(np.float64).copy()
    st = p["strand"]
    for (r, c) in reversed(pairs):
        nl1, nl2, bu = r - 1, c - 1, r - 1
        if st[nl1] > 0:
            q1, q2, e1, e2 = qs[nl1], qs[nl2], qe[nl1], qe[nl2]
            t1, t2, f1, f2 = ts[nl1], ts[nl2], te[nl1], te[nl2]
            qe[bu], qs[bu] = max(e1, e2), min(q1, q2)
            te[bu], ts[bu] = max(f1, f2), min(t1, t2)
        else:
            if ts[nl1] > ts[nl2]:
                nl1, nl2 = nl2, bu
            q1, q2, e1, e2 = qs[nl1], qs[nl2], qe[nl1], qe[nl2]
            t1, t2, f1, f2 = ts[nl1], ts[nl2], te[nl1], te[nl2]
            qe[bu], qs[bu] = min(e1, e2), max(q1, q2)
            te[bu], ts[bu] = max(f1, f2), min(t1, t2)
        nm[bu] = nm[nl1] + nm[nl2]
        al[bu] = al[nl1] + al[nl2]
    out = dict(p)
    out.update(qstart=qs, qend=qe, tstart=ts, tend=te, nmatch=nm, alen=al)
    cols = {c for _, c in pairs}
    keep = np.array([k + 1 not in cols for k in range(len(qs))], dtype=bool)
    out = take(out, keep)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = np.abs(out["tend"] - out["tstart"]) / np.abs(out["qend"] - out["qstart"])
    ok = (ratio >= 0.5) & (ratio <= 2)                    # dplyr::between; NaN (0 / 0) rows are treated as dropped
    return take(out, ok)


def compress_paf_fnct(p, second_run=False, chunklen=10000.0, compression=None, n_quadrants_per_axis=None):
    """R/compress_paf.R:173-406 with inpaf_df, save_outpaf = F, qname = NULL."""
    p = {k: np.asarray(v, dtype=np.float64) for k, v in p.items()}
    p = take(p, np.argsort(p["qstart"], kind="stable"))
    n = len(p["qstart"])
    if n <= 1:
        return p
    p = take(p, np.argsort(p["tstart"], kind="stable"))
    range_t = float(np.max(p["tend"]))
    minlen_for_merge = 0 if n < 15000 else chunklen * 0.9
    if n_quadrants_per_axis is None:
        n_quadrants_per_axis = 2 if range_t < 50000 else 10
    short = take(p, p["alen"] <= minlen_for_merge)
    p = take(p, p["alen"] > minlen_for_merge)
    nq = n_quadrants_per_axis
    tsteps = np.rint(1 + np.arange(nq) * ((range_t - 1) / (nq - 1))) if nq > 1 else np.array([1.0])  # from + k * by
    d = np.diff(tsteps)
    tstepsize = float(d[0]) if len(d) else float("nan")
    if nq == 1:
        tstepsize = 1e10
    rowpairs = []
    ts, te = p["tstart"], p["tend"]
    for tstep in tsteps[:-1]:
        sel = ((ts >= tstep) & (ts <= tstep + tstepsize)) | ((te >= tstep) & (te <= tstep + tstepsize))
        idx = np.nonzero(sel)[0]
        if len(idx) > 1:
            rowpairs += merge_paf_entries_intraloop(take(p, idx), idx + 1, second_run, chunklen, compression)
    if not rowpairs:
        return p
    order = sorted(range(len(rowpairs)), key=lambda k: rowpairs[k][1])      # order(rowpairs$col), stable
    rowpairs = [rowpairs[k] for k in order]
    rowpairs = list(dict.fromkeys(rowpairs))
    rowpairs = [(r, c) for r, c in rowpairs if r != c]
    ml = [p["nmatch"][r - 1] + p["nmatch"][c - 1] for r, c in rowpairs]
    keep = [k for k in range(len(rowpairs)) if ml[k] > chunklen / 5]
    rowpairs, ml = [rowpairs[k] for k in keep], [ml[k] for k in keep]
    # group_by(row) %>% top_n(1, combined_matchlen) %>% group_by(col) %>% top_n(1, combined_matchlen)
    best = {}
    for (r, c), m in zip(rowpairs, ml):
        best[r] = max(best.get(r, -np.inf), m)
    k1 = [k for k in range(len(rowpairs)) if ml[k] == best[rowpairs[k][0]]]
    rowpairs, ml = [rowpairs[k] for k in k1], [ml[k] for k in k1]
    best = {}
    for (r, c), m in zip(rowpairs, ml):
        best[c] = max(best.get(c, -np.inf), m)
    k2 = [k for k in range(len(rowpairs)) if ml[k] == best[rowpairs[k][1]]]
    rowpairs = [rowpairs[k] for k in k2]
    if rowpairs:
        p = merge_rows(p, rowpairs)
    return {k: np.concatenate([p[k], short[k]]) for k in p}


def enforce_slope_one(p):
    """R/tsv_to_bitlocus.R:258-275."""
    ql = np.abs(p["qend"] - p["qstart"])
    tl = np.abs(p["tend"] - p["tstart"])
    sq = -np.maximum(-ql, -tl)
    q = dict(p)
    q["qend"] = p["qstart"] + sq
    q["tend"] = p["tstart"] + sq
    return take(q, (q["qstart"] != q["qend"]) & (q["tstart"] != q["tend"]))


def load_and_prep_paf(p, minlen, compression, chunklen=10000.0):
    """R/tsv_to_bitlocus.R:15-110 from the parsed table on; returns the table make_xy_grid_fast reads."""
    p = {k: np.asarray(v, dtype=np.float64) for k, v in p.items()}
    p = swap_minus(p)
    p = compress_paf_fnct(p, second_run=True, chunklen=chunklen, compression=compression)
    p = take(p, p["alen"] > minlen)
    if len(p["alen"]) == 0:
        return p
    for k in ("qstart", "qend", "tstart", "tend"):
        p[k] = np.rint(p[k] / compression) * compression
    p = swap_minus(p)
    p = enforce_slope_one(p)
    p = swap_minus(p)
    return p
