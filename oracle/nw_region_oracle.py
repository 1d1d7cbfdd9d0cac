"""CPU restatement of the R code either side of the Julia solver call.

TEST INFRASTRUCTURE ONLY (imported by tests/ and bench.py's cpu_baseline leg, never by nahrwhals_b200/).

Follows, function by function and in the matrix representation R uses (n_y rows x n_x columns, row.names = y block
lengths, colnames = x block lengths; the product works on index lists instead):

    flip_bitl_y_if_needed           R/flip_bitl_y_if_needed.R:12-76
    find_sv_opportunities           R/find_sv_opportunities.R:12-98
    carry_out_compressed_sv         R/seqmodder.R:10-109   (including its colname repairs, which change lengths)
    calc_coarse_grained_aln_score   R/evaltools.R:87-303   (the R twin of slow_score, with the symmetry / estimate
    calculate_estimated_aln_score   R/evaltools.R:51-71     pre-filters and the forcecalc branch)
    diagonalize, return_diag_values_new, calc_symm, is_cluttered, reduce_depth_if_needed, sort_new_by_penalised,
    transform_res_new_to_old, annotate_res_out_with_positions_lens      R/explore_mutation_space_v2_helpers.R
    dfs / dfsutil                   R/dfs_machinery.R:16-246   (as called with maxdepth <= 1)
    solve_mutation                  R/explore_mutation_space_v2.R:14-112 (as called from the wrapper: maxdepth = 1)
    solve_mutation_julia_wrapper    R/juliasolver.R:2-31   (flip, berzerk guard, solver call)
    format_julia_output, concat_triplet, process_row, turn_mut_max_into_svdf     R/juliasolver.R:56-263
    the merge of both result tables R/wrapper_aln_and_analyse.R:176-191
    save_to_logfile (summary part)  R/save_logfile.R:14-51

PARITY UNPINNED: R is not installed here, so nothing below could be executed against the reference.  The only
known answers are README.md:135-139 (res_ref 71.378 for the bundled test run), checked in tests/test_region_oracle.py.
R semantics relied upon: order() is stable; round(x, 0) is half-even; round(x, 3) is the nearest 3-decimal double
(for non-ties identical to Julia's); unnamed vector arguments of cbind() get the column name "" (as.numeric -> NA);
`x[1:0]` is x[1]; rlang::hash() of two numeric vectors is equal iff the vectors are identical.
"""
import math

import numpy as np

NA = None


def r_round0(x):
    """R round(x) (digits = 0): IEC 60559 half-even."""
    return float(np.rint(x))


def r_round3(x):
    y = float(np.rint(x * 1000.0)) / 1000.0
    return y if math.isfinite(y) else x


class Bitl:
    """An R numeric matrix with dimnames: m[y, x], rn = as.numeric(row.names), cn = as.numeric(colnames) (nan = NA)."""

    def __init__(self, m, rn, cn):
        self.m = np.array(m, dtype=np.float64).reshape(len(rn), len(cn))
        self.rn = [float(v) for v in rn]
        self.cn = [float(v) for v in cn]

    @property
    def nrow(self):
        return self.m.shape[0]

    @property
    def ncol(self):
        return self.m.shape[1]

    def cols(self, idx, neg=False):
        """bitl[, idx] (1-based list).  Returns (matrix part, names); a single column is R's dropped vector whose
        name is lost inside cbind() (nan)."""
        z = [i - 1 for i in idx]
        m = self.m[:, z]
        if neg:
            m = -m
        names = [self.cn[i] for i in z]
        return m, names


def flip_bitl_y_if_needed(b):
    """R/flip_bitl_y_if_needed.R:12-76.  Returns (bitl, flipped, flip_unsure)."""
    f = b.m.copy()
    nr, nc = f.shape
    if (nr >= 4 or nc >= 4) and (nr >= 2 and nc >= 2):
        y_min = max(r_round0(nr * (1 / 4)), 2)
        y_max = min(r_round0(nr * (3 / 4)), nr - 2)
        x_min = max(r_round0(nc * (1 / 4)), 2)
        x_max = min(r_round0(nc * (3 / 4)), nc - 2)

        def rng(a, z):  # R's a:z counts down when a > z; indices are positive here
            a, z = int(a), int(z)
            return list(range(a, z + 1)) if a <= z else list(range(a, z - 1, -1))
        for y in rng(y_min, y_max):
            if 1 <= y <= nr:
                f[y - 1, :] = 0
            elif y > nr:
                raise IndexError("subscript out of bounds")  # R would stop here
        for x in rng(x_min, x_max):
            if 1 <= x <= nc:
                f[:, x - 1] = 0
            elif x > nc:
                raise IndexError("subscript out of bounds")
    pos = float(f[f > 0].sum())
    neg = abs(float(f[f < 0].sum()))
    if pos + neg == 0:
        pos = 1.0
    lo = min(pos, neg)
    ratio = (max(pos, neg) / lo) if lo != 0 else math.inf
    unsure = ratio < 2
    if neg > pos:
        if nr == 1:
            m = -b.m
            rn = b.rn
        else:
            m = -b.m[::-1, :]            # -apply(bitl, 2, rev): names(rev(x)) == rev(names(x)), so the row names
            rn = b.rn[::-1]              # (y lengths) flip together with the rows
        m = m + 0.0                      # -0 -> 0
        return Bitl(m, rn, b.cn), True, unsure
    return b, False, unsure


def find_sv_opportunities(b):
    """R/find_sv_opportunities.R:12-98 -> list of (p1, p2, sv)."""
    seen = set()
    rows = []
    for y in range(b.nrow):
        nz = [x + 1 for x in range(b.ncol) if b.m[y, x] != 0]
        if len(nz) < 2:
            continue
        for a in range(len(nz)):          # combn(): lexicographic pairs
            for c in range(a + 1, len(nz)):
                p1, p2 = nz[a], nz[c]
                ori = int(np.sign(b.m[y, p1 - 1]) * np.sign(b.m[y, p2 - 1]))
                key = (p1, p2, ori)
                if key not in seen:       # unique() keeps first occurrences in order
                    seen.add(key)
                    rows.append(key)
    out = [(p1, p2, "inv" if o == -1 else "del") for p1, p2, o in rows]
    out += [(p1, p2, "dup") for p1, p2, o in rows if o == 1]          # rbind(all, pluspairs)
    return [t for t in out if not (t[1] - t[0] == 1 and t[2] == "inv")]


def carry_out_compressed_sv(b, p1, p2, sv):
    """R/seqmodder.R:10-109.  Lengths follow the colnames as R leaves them (including the manual repairs)."""
    n = b.ncol
    nan = float("nan")

    def part(idx, neg=False, drop=True):
        m, names = b.cols(idx, neg)
        if drop and len(idx) == 1:
            names = [nan]                 # vector argument of cbind(): column name ""
        return m, names

    def rng(a, z):
        return list(range(a, z + 1)) if a <= z else list(range(a, z - 1, -1))

    if sv == "del":
        keep = [i for i in range(1, n + 1) if not (p1 <= i <= p2 - 1)]
        m, names = b.cols(keep)
        if len(keep) == 1:                # vector -> as.matrix(); colnames(bitl)[1]   (:93-97)
            names = [b.cn[0]]
    elif sv == "dup":
        m1, n1 = part(rng(1, p2))
        m2, n2 = part(rng(p1 + 1, n))
        m, names = np.concatenate([m1, m2], axis=1), n1 + n2
        if p2 - p1 == 1:
            if p1 == 1:
                names[-1] = names[1]      # :25-28
            elif p2 == n:
                names[-1] = names[-2]     # :29-31
    elif sv == "inv":
        if p2 - p1 == 1:
            return Bitl(b.m.copy(), b.rn, b.cn)
        if p1 == 1 and p2 != n:           # border case 1 (:45-53): the flanks themselves are inverted
            m1, n1 = part(rng(p2, p1), neg=True)
            m2, n2 = part(rng(p2 + 1, n))
            m, names = np.concatenate([m1, m2], axis=1), n1 + n2
            names[-1] = b.cn[len(names) - 1]
        elif p1 > 1 and p2 == n:          # border case 2 (:56-64), drop = F
            m1, n1 = part(rng(1, p1 - 1), drop=False)
            m2, n2 = part(rng(p2, p1), neg=True, drop=False)
            m, names = np.concatenate([m1, m2], axis=1), n1 + n2
            names[0] = b.cn[0]
        elif p1 == 1 and p2 == n:         # border case 3 (:67-68)
            m, names = part(rng(p2, p1), neg=True)
        else:                             # standard case (:71-79)
            m1, n1 = part(rng(1, p1))
            m2, n2 = part(rng(p2 - 1, p1 + 1), neg=True)
            m3, n3 = part(rng(p2, n))
            m, names = np.concatenate([m1, m2, m3], axis=1), n1 + n2 + n3
        if p2 - p1 == 2:
            names = list(b.cn)            # :83-84
        elif p2 == n:
            names[p2 - 1] = names[p1 - 1]  # :85-87
    elif sv == "ref":
        return Bitl(b.m.copy(), b.rn, b.cn)
    else:
        raise ValueError(sv)
    m = m + 0.0                           # :106 minus-zero -> zero
    return Bitl(m, b.rn, names)


def calc_symm(b):
    su, sr = sum(b.rn), sum(b.cn)
    return min(su, sr) / max(su, sr)


def diagonalize(m, factor):
    nr, nc = m.shape
    r = np.arange(1, nr + 1)[:, None]
    c = np.arange(1, nc + 1)[None, :]
    return m * (np.abs(r - c) < (nc * factor))


def calculate_estimated_aln_score(b):
    """R/evaltools.R:51-71."""
    fact = 0.25
    if b.nrow < 10:
        fact = 1
    d = diagonalize(b.m, fact)
    s_rows = float(d.max(axis=1).sum())
    s_cols = float(d.max(axis=0).sum())
    return ((s_rows + s_cols) / (sum(b.rn) + sum(b.cn))) * 100


class ScorerState:
    """The globals of R/explore_mutation_space_v2.R:20-24."""

    def __init__(self):
        self.est_highest = 0.0
        self.n_eval_total = 0
        self.n_eval_calc = 0
        self.n_dp = 0


def calc_coarse_grained_aln_score(b, st, forcecalc=False, orig_symm=1.0, detail=None):
    """R/evaltools.R:87-303.  `detail` (dict) receives cost/total of the DP when it ran."""
    st.n_eval_total += 1
    nr, nc = b.nrow, b.ncol
    if nr == 1 and nc == 1:
        return 100.0
    if nr == 1 and nc > 1:
        return r_round3(float(b.m[b.m > 0].sum()) / sum(b.cn) * 100)
    if nr > 1 and nc == 1:
        return r_round3(float(b.m[b.m > 0].sum()) / sum(b.rn) * 100)
    row, col = nr, nc
    up, rt = b.rn, b.cn
    symmetry = min(sum(up), sum(rt)) / max(sum(up), sum(rt))
    if (symmetry < orig_symm * 1 and row > 5) and not forcecalc:
        return 1.0
    est = calculate_estimated_aln_score(b)
    if forcecalc:
        st.est_highest = est
    thr = 0.9 if nr < 10 else 1
    if est < st.est_highest * thr and not forcecalc:
        return 1.0
    if est > st.est_highest:
        st.est_highest = est
    st.n_eval_calc += 1
    inf = math.inf
    cum_u = np.cumsum(up)
    cum_r = np.cumsum(rt)
    D = np.where(b.m > 0, 0.0, inf)        # cost_d <- -mat; >= 0 -> Inf; < 0 -> 0
    C = np.full((row, col), 0.0)
    if row < 10:
        pct = 1
    elif row < 30:
        pct = 0.3
    else:
        pct = 0.2
    if forcecalc:
        pct = 1
    C[0, 0] = min(up[0] + rt[0], D[0, 0])
    for i in range(2, row + 1):
        if i / row > pct:
            C[i - 1, 0] = inf
        else:
            C[i - 1, 0] = min(up[i - 1] + C[i - 2, 0], cum_u[i - 2] + D[i - 1, 0])
    for j in range(2, col + 1):
        if j / col > pct:
            C[0, j - 1] = inf
        else:
            C[0, j - 1] = min(rt[j - 1] + C[0, j - 2], cum_r[j - 2] + D[0, j - 1])
    C[1:, 1:] = inf
    # which(t(middleval_mat) == 1, arr.ind = T) walks t(M) column-major == M row-major: i outer, j inner
    for i in range(2, row + 1):
        im = i / row
        for j in range(2, col + 1):
            jm = j / col
            if (im > (jm + pct)) or (jm > (im + pct)):
                continue
            C[i - 1, j - 1] = min(C[i - 2, j - 2] + D[i - 1, j - 1], C[i - 2, j - 1] + up[i - 1], C[i - 1, j - 2] + rt[j - 1])
    st.n_dp += 1
    if C[row - 1, col - 1] == inf:
        return 1.0
    total = sum(up) + sum(rt)
    if detail is not None:
        detail["cost"] = float(C[row - 1, col - 1])
        detail["total"] = float(total)
    return r_round3((1 - C[row - 1, col - 1] / total) * 100)


def return_diag_values_new(b, fraction=0.05):
    """R/explore_mutation_space_v2_helpers.R:98-113 -> tuple of values (what rlang::hash() is fed)."""
    nr, nc = b.nrow, b.ncol
    mn = min(nr, nc)
    max_dist = int(r_round0(max(5, fraction * nr)))
    last = (nr + 1) * (mn - 1) + 1
    diag_idx = np.arange(1, last + 1, nr + 1)
    idxs = diag_idx[:, None] + np.arange(-max_dist, max_dist + 1)[None, :]
    flat = idxs.flatten(order="F")               # R flattens column-major
    clean = flat[(flat > 0) & (flat <= last)]
    clean2 = nr * nc - clean
    allidx = np.concatenate([clean, clean2])
    allidx = allidx[allidx != 0]                 # a zero subscript selects nothing
    v = b.m.flatten(order="F")
    if allidx.size and (allidx.min() < 0 or allidx.max() > v.size):
        raise IndexError("return_diag_values_new: index outside the matrix")
    return tuple((v[allidx - 1] + 0.0).tolist())


def is_cluttered(b, limit=5):
    """R/explore_mutation_space_v2_helpers.R:209-235."""
    nr, nc = b.nrow, b.ncol
    if not (nr > 5 and nc > 5):
        return False
    c = [int((b.m[0, 1:nc - 1] != 0).sum()), int((b.m[nr - 1, 1:nc - 1] != 0).sum()),
         int((b.m[1:nr - 1, 0] != 0).sum()), int((b.m[1:nr - 1, nc - 1] != 0).sum())]
    return any(v >= limit for v in c)


def reduce_depth_if_needed(b, maxdepth):
    """R/explore_mutation_space_v2_helpers.R:318-347 (increase_only = F)."""
    bf, _, _ = flip_bitl_y_if_needed(b)
    n_pairs = len(find_sv_opportunities(bf))
    if n_pairs > 200 and maxdepth == 3:
        maxdepth = 2
    if n_pairs > 600:
        maxdepth = 1
    if n_pairs > 2000:
        maxdepth = 0
    return maxdepth


def dfs(b, maxdepth, st):
    """R/dfs_machinery.R:16-246 for maxdepth in {0, 1}.  Returns rows (hash key, depth, mut_path, eval).
    st.flip_unsure receives the side effect of flip_bitl_y_if_needed on log_collection (R/flip_bitl_y_if_needed.R:56-58)."""
    bitl, _, unsure = flip_bitl_y_if_needed(b)
    if unsure:
        st.flip_unsure = True
    orig_symm = calc_symm(bitl)
    visited = {}
    ref_hash = return_diag_values_new(carry_out_compressed_sv(bitl, 1, 1, "ref"))
    rows = []
    # dfsutil on the ref node: forcecalc (pair$sv == "ref")
    s_ref = calc_coarse_grained_aln_score(bitl, st, forcecalc=True, orig_symm=orig_symm)
    rows.append((ref_hash, 0, "1_1_ref", s_ref))
    visited[ref_hash] = s_ref
    if not (s_ref >= 0 and 0 < maxdepth):
        return rows
    for (p1, p2, sv) in find_sv_opportunities(bitl):
        mut = carry_out_compressed_sv(bitl, p1, p2, sv)
        h = return_diag_values_new(mut)
        path = "1_1_ref+%d_%d_%s" % (p1, p2, sv)
        if h in visited:
            if visited[h] > 1:
                rows.append((h, 1, path, visited[h]))
            continue
        s = calc_coarse_grained_aln_score(mut, st, forcecalc=False, orig_symm=orig_symm)
        if not math.isnan(s):
            if s > 90:
                rows.append((h, 1, path, s))
        visited[h] = s
    return rows


def annotate_positions(rows, b):
    """annotate_res_out_with_positions_lens (R/explore_mutation_space_v2_helpers.R:412-485) for tables whose only
    mutation column is mut1.  rows: list of dicts with eval, mut1; adds mut1_start/end/pos_pm/len/len_pm."""
    cn = b.cn
    count = 1
    stopped = False
    for r in rows:
        for k in ("mut1_start", "mut1_end", "mut1_pos_pm", "mut1_len", "mut1_len_pm", "mut2_len", "mut2_len_pm",
                  "mut3_len", "mut3_len_pm"):
            r.setdefault(k, NA)
    for r in rows:
        count += 1
        if count > 2000:
            stopped = True
            break
        if r["mut1"] == "ref":
            continue
        a, z = [int(v) for v in r["mut1"].split("_")[:2]]

        def lead(p):  # sum(colnumbers[1:(p - 1)]): 1:0 selects element 1
            return sum(cn[:p - 1]) if p >= 2 else cn[0]
        start_mean = r_round0(lead(a) + cn[a - 1] / 2)
        end_mean = r_round0(lead(z) + cn[z - 1] / 2)
        start_pm = r_round0(cn[a - 1] / 2)
        end_pm = r_round0(cn[z - 1] / 2)
        lr = ((end_mean - end_pm) - (start_mean + start_pm), (end_mean + end_pm) - (start_mean - start_pm))
        mean = (lr[0] + lr[1]) / 2
        r["mut1_start"], r["mut1_end"], r["mut1_pos_pm"] = start_mean, end_mean, start_pm
        r["mut1_len"], r["mut1_len_pm"] = mean, lr[1] - mean
    return rows, stopped


def solve_mutation_depth1(b, compression=1000):
    """solve_mutation(bitlocus, maxdepth = 1, ...) (R/explore_mutation_space_v2.R:14-112).
    Returns dict(res = rows sorted as R returns them, log = {...})."""
    st = ScorerState()
    log = {}
    if is_cluttered(b):
        rows = [dict(eval=0.0, mut1="ref")]   # annotate_res_out_with_positions_lens(res_out, bitlocus): ref row, all NA
        for k in ("mut1_start", "mut1_end", "mut1_pos_pm", "mut1_len", "mut1_len_pm", "mut2_len", "mut2_len_pm",
                  "mut3_len", "mut3_len_pm"):
            rows[0][k] = NA
        return dict(res=rows, log=dict(cluttered_boundaries=True), state=st)
    maxdepth = reduce_depth_if_needed(b, 1)
    rows = dfs(b, maxdepth, st)
    # sort_new_by_penalised: bitlocus_f is the un-flipped matrix; colnames are the same either way
    step_penalty = (compression / sum(b.cn) * 100) * 1.5
    pen = [r[3] - r[1] * step_penalty for r in rows]
    order = sorted(range(len(rows)), key=lambda k: -pen[k])     # stable, decreasing
    rows = [rows[k] for k in order]
    # transform_res_new_to_old
    res = []
    for (_, depth, path, ev) in rows:
        parts = path.split("+")
        res.append(dict(eval=float(ev), mut1=(parts[1] if len(parts) > 1 else "ref")))
    res, _ = annotate_positions(res, b)
    log.update(depth=maxdepth, mut_simulated=len(rows), mut_tested=len(rows))
    if getattr(st, "flip_unsure", False):
        log["flip_unsure"] = True
    return dict(res=res, log=log, state=st)


# ---------------------------------------------------------------- downstream of the solver (R/juliasolver.R:56-263)

def turn_mut_max_into_svdf(t):
    """R/juliasolver.R:138-245 (correction = T).  t: 'p1_p2_sv+...'.  Returns list of [start, end, sv]."""
    svs = []
    for s in t.split("+"):
        a, z, k = s.split("_")
        svs.append([float(a), float(z), k])
    n = len(svs)
    if n == 1:
        return svs
    for i in range(n - 1):
        sv = svs[i][2]
        s0, e0 = svs[i][0], svs[i][1]
        starts = [v[0] for v in svs]
        ends = [v[1] for v in svs]
        later = [k > i for k in range(n)]
        if sv == "inv":
            for k in range(n):
                ns = ne = None
                if starts[k] > s0 and starts[k] < e0 and later[k]:
                    ns = e0 - (svs[k][0] - s0)
                if ends[k] > s0 and ends[k] < e0 and later[k]:
                    ne = e0 - (svs[k][1] - s0)
                if ne is not None:
                    svs[k][1] = ne
                if ns is not None:
                    svs[k][0] = ns
            for k in range(n):
                if svs[k][0] > svs[k][1]:
                    svs[k][0], svs[k][1] = svs[k][1], svs[k][0]
        if sv == "del":
            for k in range(n):
                if starts[k] > s0 and later[k]:
                    svs[k][0] = svs[k][0] + (e0 - s0)
                if ends[k] > s0 and later[k]:
                    svs[k][1] = svs[k][1] + (e0 - s0)
        elif sv == "dup":
            for k in range(n):
                if starts[k] > e0 and later[k]:
                    svs[k][0] = svs[k][0] - (e0 - s0)
                if ends[k] > e0 and later[k]:
                    svs[k][1] = svs[k][1] - (e0 - s0)
    return svs


def format_julia_output(tsv_text, gridlines_x, depth):
    """R/juliasolver.R:56-120.  Returns dict(rows = [...], n_mut, log)."""
    lines = [ln for ln in tsv_text.split("\n") if ln != ""]
    recs = [ln.split("\t") for ln in lines]
    ncols = max(len(r) for r in recs)
    n_mut = (ncols - 2) // 3
    if n_mut == 0:
        # data.frame(eval = jout[1], nmut = jout[2], mut1 = 'ref'): one row per line, 3 columns
        rows = [dict(eval=float(r[0]), nmut=int(r[1]), mut1="ref") for r in recs]
        return dict(rows=rows, n_mut=0, only_ref=True, log=dict(depth=depth, mut_simulated=0, mut_tested=0))
    rows = []
    for r in recs:
        ev = float(r[0])
        nm = int(r[1])
        muts = []
        for i in range(n_mut):
            tri = r[2 + 3 * i: 5 + 3 * i]
            if len(tri) < 3 or tri[0] == "":
                muts.append(NA)
            else:
                muts.append("%s_%s_%s" % (tri[1].strip(), tri[2].strip(), tri[0].replace("move_", "")))
        row = dict(eval=ev, nmoves=nm)        # R column order: eval, mut_max (= V2), mut1..mutN, then the coordinates
        for i in range(n_mut):
            row["mut%d" % (i + 1)] = muts[i]
        for i in range(n_mut):
            for suffix in ("start", "end", "len"):
                row["mut%d_%s" % (i + 1, suffix)] = NA
        if nm != 0:
            present = [m for m in muts if m is not NA]
            df = turn_mut_max_into_svdf("+".join(present))

            def gl(v):
                k = int(v)
                return float(gridlines_x[k - 1]) if 1 <= k <= len(gridlines_x) else NA
            for i, (s, e, _) in enumerate(df):
                a, z = gl(s), gl(e)
                row["mut%d_start" % (i + 1)] = a
                row["mut%d_end" % (i + 1)] = z
                row["mut%d_len" % (i + 1)] = (z - a) if (a is not NA and z is not NA) else NA
        rows.append(row)

    def n_na(row):  # rowSums(is.na(jout_combine)): eval, mut_max (= V2), mut*, mut*_start/end/len
        return sum(1 for k, v in row.items() if v is NA)
    order = sorted(range(len(rows)), key=lambda k: (-rows[k]["eval"], -n_na(rows[k])))
    rows = [rows[k] for k in order]
    for row in rows:
        if row["nmoves"] == 0:
            row["mut1"] = "ref"
        del row["nmoves"]
    return dict(rows=rows, n_mut=n_mut, only_ref=False,
                log=dict(depth=depth, mut_simulated=len(rows), mut_tested=len(rows)))


def merge_results(res_r, fj):
    """R/wrapper_aln_and_analyse.R:177-191: res (R depth-1 table) and res_julia (format_julia_output)."""
    if fj["only_ref"] and len(fj["rows"]) == 1:   # dim(res_julia) == c(1, 3)
        return res_r
    rows = [dict(r) for r in fj["rows"]]
    if not any(r.get("mut1") == "ref" for r in rows):
        ref_eval = [r["eval"] for r in res_r if r["mut1"] == "ref"][0]
        keys = list(rows[0].keys())
        new = {k: NA for k in keys}
        new[keys[0]] = ref_eval
        new[keys[1]] = "ref"
        rows.append(new)

        def n_na(row):
            return sum(1 for v in row.values() if v is NA)
        order = sorted(range(len(rows)), key=lambda k: (-rows[k]["eval"], -n_na(rows[k])))
        rows = [rows[k] for k in order]
    return rows


def logfile_summary(res):
    """save_to_logfile (R/save_logfile.R:17-51): res_ref, res_max, n_res_max, mut_max and the maxres row."""
    refs = [r["eval"] for r in res if r.get("mut1") == "ref"]
    res_ref = refs[0] if len(refs) == 1 else (refs if refs else NA)
    res_max = res[0]["eval"]
    mx = max(r["eval"] for r in res)
    top = [r for r in res if r["eval"] == mx]
    n_res_max = len(top)
    mcols = [k for k in res[0].keys() if k in ("mut1", "mut2", "mut3")]
    na_counts = [sum(1 for k in mcols if r.get(k) is NA) for r in top]
    pos = na_counts.index(max(na_counts))        # which.max: first maximum, used as a row number of the FULL table
    maxres = res[pos]
    if len(maxres) > 2:
        nmut = sum(1 for k in ("mut1", "mut2", "mut3") if k in maxres and maxres[k] is not NA)
        mut_max = "+".join(str(maxres["mut%d" % (i + 1)]) for i in range(nmut))
    elif len(maxres) == 2:
        mut_max = str(maxres["mut1"])
    else:
        mut_max = "ref"
    return dict(res_ref=res_ref, res_max=res_max, n_res_max=n_res_max, mut_max=mut_max, maxres=maxres)


def berzerk_width(bf, init_width):
    """R/juliasolver.R:9-30 on the flipped matrix."""
    moves = find_sv_opportunities(bf)
    dups = [m for m in moves if m[2] == "dup"]
    longest = max([m[1] - m[0] for m in dups]) if dups else -math.inf
    if len(moves) > 400 or (len(moves) > 200 and longest > 40 and len(dups) > 40):
        return init_width / 10
    return init_width


def analyse_region(mat, len_y, len_x, gridlines_x, solver, depth=3, maxdup=2, init_width=1000, minreport=0.98,
                   compression=1000):
    """R/wrapper_aln_and_analyse.R:163-191 + save_to_logfile's summary.  `solver(aln_sign [n_x, n_y] int8, len_x,
    len_y, maxdepth, maxdup, maxwidth, minreport, maxreport) -> TSV text` is the Julia solver stand-in."""
    b = Bitl(mat, len_y, len_x)
    pre = solve_mutation_depth1(b, compression)
    res = pre["res"]
    log = dict(pre["log"])
    julia_called = False
    if pre["log"].get("cluttered_boundaries"):
        # the wrapper returns before the solver (R/wrapper_aln_and_analyse.R:164-172)
        summ = logfile_summary(res)
        return dict(res=res, log=log, julia_called=False, **summ)
    if max(r["eval"] for r in res) < 99.8:
        julia_called = True
        bf, flipped, unsure = flip_bitl_y_if_needed(b)
        log["flip_unsure"] = unsure or log.get("flip_unsure", False)   # the flag is only ever set, never cleared
        width = berzerk_width(bf, init_width)
        aln = np.sign(bf.m).T.astype(np.int8)
        tsv = solver(aln, np.array(bf.cn), np.array(bf.rn), depth, maxdup, width, minreport, 1000)
        fj = format_julia_output(tsv, gridlines_x, depth)
        log.update(fj["log"])
        res = merge_results(res, fj)
    summ = logfile_summary(res)
    return dict(res=res, log=log, julia_called=julia_called, **summ)


# ---------------------------------------------------------------------------------------------------------------
# save_to_logfile: the row of nahrwhals_res.tsv (R/save_logfile.R:112-232)
# ---------------------------------------------------------------------------------------------------------------
LOG_COLUMNS = ["seqname", "start", "end", "sample", "y_seqname", "y_start", "y_end", "width_orig", "xpad", "res_ref", "res_max",
               "n_res_max", "mut_max", "mut_simulated", "mut_tested", "search_depth", "grid_compression", "exceeds_x",
               "exceeds_y", "grid_inconsistency", "flip_unsure", "cluttered_boundaries"] + \
              ["mut%d_%s" % (i, s) for i in (1, 2, 3) for s in ("start", "end", "pos_pm", "len", "len_pm")]


def r_fixed15(v):
    """A double as write.table prints it under options(scipen = 999): <= 15 significant digits, fixed notation."""
    if v is NA or v != v:
        return "NA"
    v = float(v)
    if v == 0:
        return "0"
    target = float("%.14e" % v)
    nsig = 15
    for d in range(1, 16):
        if float("%.*e" % (d - 1, v)) == target:
            nsig = d
            break
    txt = "%.*e" % (nsig - 1, v)
    kp = int(txt.split("e")[1])
    return "%.*f" % (max(0, nsig - kp - 1), float(txt))


def logfile_row(summ, log):
    """summ: logfile_summary(); log: the log_collection entries (chr / start / end / samplename / xpad as the CHARACTER
    values of R/wrapper_aln_and_analyse.R:13-24, y_* or None, mut_simulated, mut_tested, depth, compression or None,
    exceeds_x, exceeds_y, grid_inconsistency, flip_unsure, cluttered_boundaries)."""
    txt = lambda v: "NA" if v is None else str(v)                      # noqa: E731
    lgl = lambda v: "NA" if v is None else ("TRUE" if v else "FALSE")  # noqa: E731
    start, end = float(log["start"]), float(log["end"])
    mr = summ["maxres"]
    num = lambda k: mr[k] if (k in mr and mr[k] is not NA and mr[k] is not None) else float("nan")  # noqa: E731
    comp = log.get("compression")
    f = [txt(log["chr"]), txt(log["start"]), txt(log["end"]), txt(log["samplename"]), txt(log.get("y_seqname")),
         txt(log.get("y_start")), txt(log.get("y_end")), r_fixed15(end - start), txt(log["xpad"]), r_fixed15(summ["res_ref"]),
         r_fixed15(summ["res_max"]), str(int(summ["n_res_max"])), summ["mut_max"], str(int(log["mut_simulated"])),
         str(int(log["mut_tested"])), str(int(log["depth"])), r_fixed15(0 if comp is None else comp), lgl(log["exceeds_x"]),
         lgl(log["exceeds_y"]), lgl(log["grid_inconsistency"]), lgl(log["flip_unsure"]), lgl(log["cluttered_boundaries"])]
    for i in (1, 2, 3):
        f.append(r_fixed15(num("mut%d_start" % i) + start))
        f.append(r_fixed15(num("mut%d_end" % i) + start))
        for s_ in ("pos_pm", "len", "len_pm"):
            f.append(r_fixed15(num("mut%d_%s" % (i, s_))))
    return "\t".join(f)
