"""Synthetic segmental-duplication gridmatrices (the `sdgrid` generator of SURVEY.md section 8d).

Produces the two inputs the solver reads (R/juliasolver.R:5-7 in the reference): a signed gridmatrix with
n_y rows (target blocks) x n_x columns (reference blocks), cell = sign * x-block length, and the two block
length vectors.  Reference blocks are grouped into repeat *families* (members share a length); the target
arrangement is the reference arrangement after a hidden chain of NAHR-style moves, so a perfect solution of
that depth exists.  numpy default_rng(seed) only -- deterministic across machines.
"""
import numpy as np


def _apply(kind, d, p, q):
    if kind == 0:  # dup
        return d[:q] + d[p:]
    if kind == 1:  # del
        return d[:p - 1] + d[q - 1:]
    return d[:p] + [-v for v in reversed(d[p:q - 1])] + d[q - 1:]  # inv


def sdgrid(n_x, families, p_inv=0.5, chain_depth=3, seed=1, max_ny=None, len_tol=None):
    """Returns (mat[n_y, n_x] int64, len_x[n_x] int64, len_y[n_y] int64, hidden_moves).
    len_tol: if given, every step of the hidden chain keeps the target length within n_x*(1 +- len_tol)."""
    rng = np.random.default_rng(seed)
    fam = np.zeros(n_x, np.int64)  # 0 = unique block
    slots = rng.permutation(n_x)
    pos = 0
    for f, m in enumerate(families, start=1):
        fam[slots[pos:pos + m]] = f
        pos += m
    orient = np.where((fam > 0) & (rng.random(n_x) < p_inv), -1, 1)
    base_len = 10000 * rng.integers(1, 21, size=n_x)
    fam_len = 10000 * rng.integers(1, 21, size=len(families) + 1)
    len_x = np.where(fam > 0, fam_len[fam], base_len).astype(np.int64)
    seq = list(range(1, n_x + 1))
    hidden = []
    for _ in range(chain_depth):
        L = len(seq)
        cands = []
        for p in range(1, L + 1):
            for q in range(p + 1, L + 1):
                a, b = seq[p - 1], seq[q - 1]
                fa, fb = fam[abs(a) - 1], fam[abs(b) - 1]
                if fa == 0 or fa != fb:
                    continue
                rel = orient[abs(a) - 1] * orient[abs(b) - 1] * (1 if a * b > 0 else -1)
                if rel > 0:
                    hi = max_ny if len_tol is None else int(n_x * (1 + len_tol))
                    lo = max(2, n_x // 2) if len_tol is None else int(np.ceil(n_x * (1 - len_tol)))
                    if hi is None or L + (q - p) <= hi:
                        cands.append((0, p, q))
                    if L - (q - p) >= lo:
                        cands.append((1, p, q))
                elif q > p + 1:
                    cands.append((2, p, q))
        if not cands:
            break
        mv = cands[int(rng.integers(len(cands)))]
        hidden.append(mv)
        seq = _apply(mv[0], seq, mv[1], mv[2])
    n_y = len(seq)
    mat = np.zeros((n_y, n_x), np.int64)
    len_y = np.zeros(n_y, np.int64)
    for k, t in enumerate(seq):
        xb = abs(t) - 1
        len_y[k] = len_x[xb]
        if fam[xb] == 0:
            mat[k, xb] = (1 if t > 0 else -1) * len_x[xb]
        else:
            for x in np.nonzero(fam == fam[xb])[0]:
                s = orient[x] * orient[xb] * (1 if t > 0 else -1)
                mat[k, x] = s * len_x[x]
    return mat, len_x, len_y, hidden


def to_solver_inputs(mat):
    """(aln_sign int8 [n_x, n_y]) as solver.jl:103-106 derives it from the TSV."""
    return np.ascontiguousarray(np.sign(mat).T.astype(np.int8))


def write_tsvs(mat, len_x, len_y, inmat_path, inlens_path):
    """Byte layout of R/juliasolver.R:5-7 (write.table, tab separated, no names)."""
    with open(inmat_path, "w") as f:
        for row in mat:
            f.write("\t".join(str(int(v)) for v in row) + "\n")
    with open(inlens_path, "w") as f:
        f.write("\t".join(str(int(v)) for v in len_y) + "\n")
        f.write("\t".join(str(int(v)) for v in len_x) + "\n")


def random_grid(n_x, n_y, density, seed, len_choices=(1, 2, 3, 5, 8, 13)):
    """Unstructured random signed grid (parity fuzzing; not a benchmark workload)."""
    rng = np.random.default_rng(seed)
    len_x = 1000 * rng.choice(len_choices, size=n_x)
    len_y = 1000 * rng.choice(len_choices, size=n_y)
    sign = rng.choice([-1, 0, 1], size=(n_y, n_x), p=[density / 2, 1 - density, density / 2])
    for k in range(min(n_x, n_y)):  # keep a noisy main diagonal so scores are not all zero
        if rng.random() < 0.8:
            sign[k * n_y // max(n_x, n_y) if n_y < n_x else k, k if n_y >= n_x else k] = 1
    mat = sign * len_x[None, :]
    return mat.astype(np.int64), len_x.astype(np.int64), len_y.astype(np.int64)


def matrix_to_alignments(mat, len_x, len_y):
    """The alignments a signed gridmatrix stands for, in the reference's paf columns (tstart, tend, qstart, qend):
    one slope +-1 segment per non-zero cell (reverse strand: qstart > qend).  Input of the gridding step."""
    ox = np.concatenate([[0], np.cumsum(len_x)]).astype(np.float64)
    oy = np.concatenate([[0], np.cumsum(len_y)]).astype(np.float64)
    ys, xs = np.nonzero(mat)
    pos = mat[ys, xs] > 0
    ts, te = ox[xs], ox[xs + 1]
    qs = np.where(pos, oy[ys], oy[ys + 1])
    qe = np.where(pos, oy[ys + 1], oy[ys])
    return ts, te, qs, qe


def matrix_to_chunked_paf(mat, len_x, len_y, chunk=10000):
    """The PAF table chunked minimap2 would report for a signed gridmatrix (input of the PAF preparation): every
    non-zero cell is one slope +-1 alignment cut into chunk-long pieces, columns as minimap2 writes them
    (qstart < qend, strand +1 / -1).  Returns the dict of columns nahrwhals_b200.paf expects."""
    ts, te, qs, qe = matrix_to_alignments(mat, len_x, len_y)
    span = float(max(np.sum(len_x), np.sum(len_y)))
    rows = []
    for a in range(len(ts)):
        length = te[a] - ts[a]
        neg = qs[a] > qe[a]
        for k in np.arange(0, length, chunk):
            ln = min(chunk, length - k)
            q0, q1 = (qs[a] - k - ln, qs[a] - k) if neg else (qs[a] + k, qs[a] + k + ln)
            rows.append((chunk, q0, q1, -1.0 if neg else 1.0, span, ts[a] + k, ts[a] + k + ln, ln - 3, ln, 60))
    arr = np.array(rows, dtype=np.float64).reshape(-1, 10)
    cols = ("qlen", "qstart", "qend", "strand", "tlen", "tstart", "tend", "nmatch", "alen", "mapq")
    return {k: arr[:, i].copy() for i, k in enumerate(cols)}
