"""CPU restatement of the reference's PAF preparation upstream of the gridding.

TEST INFRASTRUCTURE ONLY (imported by tests/ and bench.py's cpu_baseline leg, never by nahrwhals_b200/).

    load_and_prep_paf_for_gridplot_transform   R/tsv_to_bitlocus.R:15-110   (strand swap, compress, minlen filter,
                                               rounding to the compression grid, enforce_slope_one :258-275)
    compress_paf_fnct                          R/compress_paf.R:173-406    (inpaf_df path, save_outpaf = F; and the file
                                               path inpaf_link -> outpaf_link of R/seqbuilder_functions.R:107:
                                               compress_paf_file, with R's number formatting restated)
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


def merge_rows(p, pairs, names=None, reduced_qnames=False, q_integer=True):
    """R/compress_paf.R:13-138.  pairs: list of (row, col), 1-based.  names (file mode): list of qname strings aligned
    with the rows, rewritten in place as :49-59 / :86-97 do and filtered like the rows."""
    qs, qe = p["qstart"].astype(np.float64).copy(), p["qend"].astype(np.float64).copy()
    ts, te = p["tstart"].astype(np.float64).copy(), p["tend"].astype(np.float64).copy()
    nm, al = p["nmatch"].astype(np.float64).copy(), p["alen"].astype(np.float64).copy()
    st = p["strand"]
    for (r, c) in reversed(pairs):
        nl1, nl2, bu = r - 1, c - 1, r - 1
        if st[nl1] > 0:
            q1, q2, e1, e2 = qs[nl1], qs[nl2], qe[nl1], qe[nl2]
            t1, t2, f1, f2 = ts[nl1], ts[nl2], te[nl1], te[nl2]
            if names is not None and not reduced_qnames:
                import re
                names[bu] = re.sub(r"_[^_]*$", "", names[nl1]) + "_" + r_format(q1, q_integer) + "-" + r_format_double(e2 - 1)
            qe[bu], qs[bu] = max(e1, e2), min(q1, q2)
            te[bu], ts[bu] = max(f1, f2), min(t1, t2)
        else:
            if ts[nl1] > ts[nl2]:
                nl1, nl2 = nl2, bu
            q1, q2, e1, e2 = qs[nl1], qs[nl2], qe[nl1], qe[nl2]
            t1, t2, f1, f2 = ts[nl1], ts[nl2], te[nl1], te[nl2]
            if names is not None:
                names[bu] = names[nl1]                     # :97 overrides the rebuilt name with qname1
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
    if names is not None:
        kept = [names[i] for i in np.nonzero(keep)[0]]
        names[:] = [nm for nm, good in zip(kept, ok) if good]
    return take(out, ok)


def compress_paf_fnct(p, second_run=False, chunklen=10000.0, compression=None, n_quadrants_per_axis=None, file_mode=None):
    """R/compress_paf.R:173-406 with inpaf_df, save_outpaf = F, qname = NULL.  file_mode (compress_paf_file): a dict with
    'qname' (names by the '_row' column of p) and 'q_integer'; receives 'merged' and 'out_qname'."""
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
        if file_mode is not None:
            names = [file_mode["qname"][int(r)] for r in p["_row"]]
            p = merge_rows(p, rowpairs, names=names, reduced_qnames=file_mode.get("reduced", False),
                           q_integer=file_mode["q_integer"])
            file_mode["merged"] = True
            file_mode["out_qname"] = names + [file_mode["qname"][int(r)] for r in short["_row"]]
        else:
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


# ---------------------------------------------------------------------------------------------------------------
# file mode: compress_paf_fnct(inpaf_link, outpaf_link) with save_outpaf = T (R/compress_paf.R:184-216, :386-401),
# the call of R/seqbuilder_functions.R:107
# ---------------------------------------------------------------------------------------------------------------
def r_format_double(v):
    """as.character() / write.table() of an R double: fewest significant digits (<= 15) reproducing the value at 15
    digits; fixed notation unless scientific is narrower (formatReal, scipen = 0)."""
    v = float(v)
    if v == 0:
        return "0"
    target = float("%.14e" % v)
    nsig = 15
    for d in range(1, 16):
        if float("%.*e" % (d - 1, v)) == target:
            nsig = d
            break
    mant, ex = ("%.*e" % (nsig - 1, v)).split("e")
    kp = int(ex)
    neg = 1 if v < 0 else 0
    rgt = max(0, nsig - kp - 1)
    left = kp + 1 if kp >= 0 else 1
    w_fixed = neg + left + (rgt + 1 if rgt > 0 else 0)
    w_sci = neg + (nsig + 1 if nsig > 1 else 1) + (5 if abs(kp) >= 100 else 4)
    if w_fixed <= w_sci:
        return "%.*f" % (rgt, v)
    return "%se%s%02d" % (mant, "-" if kp < 0 else "+", abs(kp))


def r_format(v, integer_typed):
    return str(int(v)) if integer_typed else r_format_double(v)


NUMCOLS = ("qlen", "qstart", "qend", "tlen", "tstart", "tend", "nmatch", "alen", "mapq")


def read_paf_file(path, dedupe=False):
    """read.table(path, sep = '\\t') [+ unique()] -> (table with '_row', qname, tname, extra, integer_typed)."""
    seen, rows, qname, tname, extra = set(), [], [], [], []
    ityp = {k: True for k in NUMCOLS}
    for line in open(path).read().split("\n"):
        line = line.rstrip("\r")
        if not line:
            continue
        if dedupe:
            if line in seen:
                continue
            seen.add(line)
        f = line.split("\t")
        assert len(f) >= 12, "fewer than 12 columns"
        qname.append(f[0])
        tname.append(f[5])
        extra.append("\t".join(f[12:]))
        num = dict(zip(("qlen", "qstart", "qend"), f[1:4]))
        num.update(zip(("tlen", "tstart", "tend", "nmatch", "alen", "mapq"), f[6:12]))
        for k, txt in num.items():
            v = float(txt)
            if v != int(v) or abs(v) > 2147483647 or any(c in txt for c in ".eE"):
                ityp[k] = False
        rows.append([float(num["qlen"]), float(num["qstart"]), float(num["qend"]), 1.0 if f[4] == "+" else -1.0,
                     float(num["tlen"]), float(num["tstart"]), float(num["tend"]), float(num["nmatch"]), float(num["alen"]),
                     float(num["mapq"]), float(len(rows))])
    a = np.array(rows, dtype=np.float64).reshape(-1, 11)
    p = {k: a[:, i].copy() for i, k in enumerate(COLS + ("_row",))}
    return p, qname, tname, extra, ityp


def _write_rows(f, out, names, tname, extra, ityp, merged, qi):
    """write.table(inpaf, quote = F, col.names = F, row.names = F, sep = '\\t') after the final strand swap."""
    if merged:
        ityp = dict(ityp, nmatch=False, alen=False)
    out = swap_minus(out)
    for k in range(len(out["qstart"])):
        r = int(out["_row"][k])
        cells = [names[k], r_format(out["qlen"][k], ityp["qlen"]), r_format(out["qstart"][k], qi), r_format(out["qend"][k], qi),
                 "+" if out["strand"][k] > 0 else "-", tname[r], r_format(out["tlen"][k], ityp["tlen"]),
                 r_format(out["tstart"][k], ityp["tstart"]), r_format(out["tend"][k], ityp["tend"]),
                 r_format(out["nmatch"][k], ityp["nmatch"]), r_format(out["alen"][k], ityp["alen"]),
                 r_format(out["mapq"][k], ityp["mapq"])]
        if not merged and extra[r]:
            cells.append(extra[r])
        f.write("\t".join(cells) + "\n")


def compress_paf_file(in_path, out_path, chunklen, n_quadrants_per_axis=None):
    p, qname, tname, extra, ityp = read_paf_file(in_path, dedupe=True)
    p = swap_minus(p)
    fm = dict(qname=qname, q_integer=ityp["qstart"] and ityp["qend"], merged=False, out_qname=None)
    out = compress_paf_fnct(p, second_run=False, chunklen=chunklen, n_quadrants_per_axis=n_quadrants_per_axis, file_mode=fm)
    names = fm["out_qname"] if fm["merged"] else [qname[int(r)] for r in out["_row"]]
    with open(out_path, "w") as f:
        _write_rows(f, out, names, tname, extra, ityp, fm["merged"], fm["q_integer"])
    return len(qname), len(names)


def compress_paf_file_by_qname(in_path, out_path, chunklen, n_quadrants_per_axis=10):
    """The loop of R/make_whole_genome_list.R:80-106: chunk suffix stripped from qname, strand swap, then per query sequence
    compress_paf_fnct(inpaf_df = paf, outpaf_link, qname = q) -- reduced_qnames, appended to the output file."""
    import re
    p, qname, tname, extra, ityp = read_paf_file(in_path, dedupe=False)
    qname = [re.sub(r"_[0-9]+-[0-9]+$", "", q) for q in qname]
    p = swap_minus(p)
    order = list(dict.fromkeys(qname))
    total = 0
    with open(out_path, "a") as f:
        for q in order:
            sel = np.array([k for k in range(len(qname)) if qname[k] == q], dtype=np.int64)
            fm = dict(qname=qname, q_integer=ityp["qstart"] and ityp["qend"], merged=False, out_qname=None, reduced=True)
            out = compress_paf_fnct(take(p, sel), second_run=False, chunklen=chunklen, n_quadrants_per_axis=n_quadrants_per_axis,
                                    file_mode=fm)
            names = fm["out_qname"] if fm["merged"] else [qname[int(r)] for r in out["_row"]]
            _write_rows(f, out, names, tname, extra, ityp, fm["merged"], fm["q_integer"])
            total += len(names)
    return len(qname), total
