"""Second, independent CPU restatement of solver.jl in pure Python (small cases only).

TEST INFRASTRUCTURE ONLY -- imported by tests/ to cross-check oracle/nw_oracle.cpp; never by the product.
Written directly against /root/reference/inst/extdata/scripts/scripts_nw_main/solver.jl (line refs below),
deliberately structured differently from the C++ oracle (dict-of-tuples, row-major DP, numpy float64) so that
a transcription slip in one of them shows up as a disagreement.

PARITY PIN STATUS: "parity unpinned" for the order of tied trajectories (Julia Dict order, solver.jl:731);
ids are iterated in discovery order instead.
"""
import math

import numpy as np

MOVE_NAMES = ("move_dup", "move_del", "move_inv")  # solver.jl:34


def corridor_pct(row):  # solver.jl:378-386
    return 1.0 if row < 10 else (0.3 if row < 30 else 0.2)


def julia_round3(x):  # Base.round(x; digits=3) -> round(x*1000)/1000 (ties to even), x if not finite
    if not math.isfinite(x):
        return x
    y = float(np.rint(np.float64(x) * 1000.0)) / 1000.0
    return y if math.isfinite(y) else x


def slow_score(index, aln, len_y, len_x, force=False):
    """solver.jl:329-376.  aln: [n_x][n_y] of -1/0/1.  Returns (score, cost, total)."""
    n_y = len(len_y)
    L = len(index)
    if (not force) and n_y > 10 and abs(L - n_y) > n_y * 0.5:  # :332-334
        return 0.0, float("nan"), 0.0
    inf = float("inf")
    row, col = L, n_y
    up = [float(len_x[abs(v) - 1]) for v in index]  # :344
    rt = [float(v) for v in len_y]  # :343
    d = [[0.0 if (aln[abs(v) - 1][j] * (1 if v > 0 else -1)) > 0 else inf for j in range(col)] for v in index]
    pct = 1.0 if force else corridor_pct(row)
    C = [[inf] * col for _ in range(row)]
    C[0][0] = min(up[0] + rt[0], d[0][0])  # :362
    cs = 0.0
    for i in range(2, row + 1):  # :388-396
        cs += up[i - 2]
        if i / row > pct:
            C[i - 1][0] = inf
        else:
            C[i - 1][0] = min(up[i - 1] + C[i - 2][0], cs + d[i - 1][0])
    cs = 0.0
    for j in range(2, col + 1):  # :398-406
        cs += rt[j - 2]
        if j / col > pct:
            C[0][j - 1] = inf
        else:
            C[0][j - 1] = min(rt[j - 1] + C[0][j - 2], cs + d[0][j - 1])
    for i in range(2, row + 1):  # :408-423 (any topological order gives the same minima)
        for j in range(2, col + 1):
            if (i / row > j / col + pct) or (j / col > i / row + pct):
                continue
            C[i - 1][j - 1] = min(C[i - 2][j - 2] + d[i - 1][j - 1], C[i - 2][j - 1] + up[i - 1],
                                  C[i - 1][j - 2] + rt[j - 1])
    total = sum(up) + sum(rt)
    cost = C[row - 1][col - 1]
    return julia_round3((1 - cost / total) * 100), cost, total


def to_moveset(aln):  # solver.jl:217-231
    n = len(aln)
    dd = [[False] * n for _ in range(n)]
    inv = [[False] * n for _ in range(n)]
    for a in range(n):
        for b in range(a + 1, n):
            prods = [p * q for p, q in zip(aln[a], aln[b])]
            dd[a][b] = dd[b][a] = any(p == 1 for p in prods)
            inv[a][b] = inv[b][a] = any(p == -1 for p in prods)
    return dd, inv


def apply_move(kind, d, p, q):  # solver.jl:244-278, 1-based p<q
    if kind == 0:
        return d[:q] + d[p:]
    if kind == 1:
        return d[:p - 1] + d[q - 1:]
    return d[:p] + tuple(-v for v in reversed(d[p:q - 1])) + d[q - 1:]


def solve(aln, len_x, len_y, maxdepth=3, maxdup=3, maxwidth=None, minreport=0.95, maxreport=None):
    """solver.jl:644-686 + :727-745.  Returns dict with ids/scores/edges/trajectories/tsv."""
    n_x = len(aln)
    dd, inv = to_moveset(aln)
    ids = {}
    seqs, scores, costs, totals, edges = [], [], [], [], []
    f32 = np.float32
    stats = []

    def insert(seq):
        s, c, t = slow_score(seq, aln, len_y, len_x)
        ids[seq] = len(seqs) + 1
        seqs.append(seq)
        scores.append(f32(s))
        costs.append(c)
        totals.append(t)
        edges.append([])
        return s

    root = tuple(range(1, n_x + 1))
    insert(root)
    best = 0.0
    frontier = [root]
    for depth in range(1, maxdepth + 1):
        if depth > 1 and maxwidth is not None:  # :669-672 stable sort by -Float32 score
            frontier = sorted(frontier, key=lambda s: -scores[ids[s] - 1])[:maxwidth]
        nxt = []
        gen = scored = 0
        for par in frontier:
            pid = ids[par]
            mult = {}
            for v in par:
                mult[abs(v)] = mult.get(abs(v), 0) + 1
            dup_ok = max(mult.values()) <= maxdup  # :292-298, :611
            L = len(par)
            for i in range(1, L + 1):
                for j in range(i + 1, L + 1):
                    a, b = par[i - 1], par[j - 1]
                    flip = a * b < 0
                    pdd, pinv = dd[abs(a) - 1][abs(b) - 1], inv[abs(a) - 1][abs(b) - 1]
                    can_dd = (flip and pinv) or ((not flip) and pdd)
                    can_inv = (flip and pdd) or ((not flip) and pinv)
                    kinds = []
                    if can_dd:
                        if dup_ok:
                            kinds.append(0)
                        kinds.append(1)
                    if can_inv:
                        kinds.append(2)
                    for kind in kinds:
                        child = apply_move(kind, par, i, j)
                        gen += 1
                        if child not in ids:
                            s = insert(child)
                            scored += 1
                            if s > best:
                                best = s
                            edges[-1].append((pid, kind, i, j))
                            if s >= 0.0:
                                nxt.append(child)
                        else:
                            edges[ids[child] - 1].append((pid, kind, i, j))
        stats.append((len(frontier), gen, scored))
        frontier = nxt

    def trajectories(cur, depth):  # :688-705
        if cur == 1:
            return [[]], True
        if depth > maxdepth:
            return [[]], False
        res = []
        for (pid, kind, i, j) in edges[cur - 1]:
            sub, valid = trajectories(pid, depth + 1)
            if valid:
                for t in sub:
                    res.append(t + [(kind, i, j)])
        return res, True

    trajs = []
    for k in range(1, len(seqs) + 1):  # canonical: ascending id
        if float(scores[k - 1]) >= minreport * best:
            for t in trajectories(k, 0)[0]:
                trajs.append((scores[k - 1], t))
    trajs.sort(key=lambda t: (-t[0], len(t[1])))  # python sort is stable
    if maxreport is not None:
        trajs = trajs[:maxreport]
    lines = []
    for s, t in trajs:
        lines.append(f32_str(s) + "\t" + str(len(t)) + "\t" +
                     "\t".join("%s\t%d\t%d" % (MOVE_NAMES[k], a, b) for (k, a, b) in t))
    return dict(seqs=seqs, scores=scores, costs=costs, totals=totals, edges=edges, best=best, stats=stats,
                tsv="".join(l + "\n" for l in lines))


def f32_str(v):
    """Julia print(::Float32) for the 3-decimal values in [0, 100] the solver emits."""
    v = np.float32(v)
    for prec in range(1, 10):
        s = "%.*g" % (prec, float(v))
        if np.float32(float(s)) == v:
            break
    if "e" in s or "E" in s:
        s = np.format_float_positional(float(s), trim="-")
    if "." not in s:
        s += ".0"
    return s
