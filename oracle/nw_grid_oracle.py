"""CPU restatement of the reference's gridding step: alignments -> gridlines -> grid list -> gridmatrix.

TEST INFRASTRUCTURE ONLY (imported by tests/, never by nahrwhals_b200/).

    make_xy_grid_fast        R/tsv_to_bitlocus.R:290-384   fix point of gridline propagation along the alignments
    bressiwrap / bresenham   R/tsv_to_bitlocus.R:415-... , R/bresenham.R:2-129   walk of every alignment over the grid
    remove_duplicates_triple R/remove_duplicates_triple.R:11-32
    gridlist_to_gridmatrix   R/gridlist_to_gridmatrix.R:8-51

An alignment is (tstart, tend, qstart, qend) as in the reference's paf data frame (target = x, query = y; tstart < tend;
qstart > qend for a reverse-strand alignment).  All arithmetic is IEEE double exactly as R does it.
PARITY UNPINNED: R is not installed here; the reference ships no gridding fixture.  R semantics relied upon:
`v[1:0]` is v[1] (so a generation that produced no candidate on an axis contributes the 0 of the pre-allocated
buffer, R/tsv_to_bitlocus.R:356-357), unique()/duplicated() keep first occurrences, order() is stable.
"""
import numpy as np


def make_xy_grid_fast(ts, te, qs, qe, generations=100000):
    """Returns (gridlines_x, gridlines_y, converged)."""
    ts, te, qs, qe = [np.asarray(v, dtype=np.float64) for v in (ts, te, qs, qe)]
    n = len(ts)
    all_x = sorted(set(ts.tolist()) | set(te.tolist()))
    all_y = sorted(set(qs.tolist()) | set(qe.tolist()))
    with np.errstate(divide="ignore", invalid="ignore"):
        slopes = (qe - qs) / (te - ts)
    converged = False
    new_x, new_y = [], []
    for gen in range(1, generations + 1):
        px, py = (all_x, all_y) if gen == 1 else (new_x, new_y)
        max_size = n * max(len(px), len(py))
        gx, gy = [], []
        pxa, pya = np.asarray(px, dtype=np.float64), np.asarray(py, dtype=np.float64)
        for r in range(n):
            cx = pxa[(pxa > ts[r]) & (pxa < te[r])]
            lo, hi = min(qs[r], qe[r]), max(qs[r], qe[r])
            cy = pya[(pya > lo) & (pya < hi)]
            gy.extend((qs[r] + slopes[r] * (cx - ts[r])).tolist())
            gx.extend((ts[r] + (cy - qs[r]) / slopes[r]).tolist())
        # new_x[1:(idx_x - 1)] with idx_x == 1 is new_x[c(1, 0)]: the first element of the zero-filled buffer
        if not gx:
            gx = [0.0] if max_size > 0 else [float("nan")]
        if not gy:
            gy = [0.0] if max_size > 0 else [float("nan")]
        sx, sy = set(all_x), set(all_y)
        new_x = list(dict.fromkeys(v for v in gx if v not in sx))
        new_y = list(dict.fromkeys(v for v in gy if v not in sy))
        if not new_x and not new_y:
            converged = True
            break
        all_x = all_x + new_x
        all_y = all_y + new_y
    return sorted(set(all_x)), sorted(set(all_y)), converged


def bresenham(x0, x1, y0, y1, gx, gy, dx, dy, xi, yi, max_steps=1 << 20):
    """R/bresenham.R:18-129.  xi / yi: 1-based grid index of the start point.  Returns [(x, y, z)]."""
    x, y = x0, y0
    out = []
    pos = y < y1
    steps = 0
    while x != x1 and y != y1:
        if not (1 <= xi <= len(dx)):
            raise IndexError("bresenham: x grid index outside the grid")
        sx = dx[xi - 1]
        if pos:
            if not (1 <= yi <= len(dy)):
                raise IndexError("bresenham: y grid index outside the grid")
            sy = dy[yi - 1]
        else:
            if not (1 <= yi - 1 <= len(dy)):
                raise IndexError("bresenham: y grid index outside the grid")
            sy = -dy[yi - 2]
        x = x + sx
        y = y + sy
        if pos:
            out.append((xi, yi, sx))
        else:
            out.append((xi, yi - 1, -sx))
        xi += 1
        yi += int(np.sign(sy))
        steps += 1
        if steps > max_steps:
            raise RuntimeError("bresenham: walk does not terminate (inconsistent grid)")
    return out


def bressiwrap(ts, te, qs, qe, gx, gy):
    """-> unique (x, y, z) rows in order of first appearance."""
    dx = np.diff(np.asarray(gx, dtype=np.float64)).tolist()
    dy = np.diff(np.asarray(gy, dtype=np.float64)).tolist()
    xpos = {v: k + 1 for k, v in reversed(list(enumerate(gx)))}
    ypos = {v: k + 1 for k, v in reversed(list(enumerate(gy)))}
    rows = []
    for r in range(len(ts)):
        rows += bresenham(float(ts[r]), float(te[r]), float(qs[r]), float(qe[r]), gx, gy, dx, dy, xpos[float(ts[r])],
                          ypos[float(qs[r])])
    return list(dict.fromkeys(rows))


def remove_duplicates_triple(gl):
    gl = list(dict.fromkeys(gl))
    seen = set()
    out = []
    for (x, y, z) in gl:
        if (x, y) in seen:
            z = 0.0
        seen.add((x, y))
        out.append((x, y, z))
    last = {}
    for k, (x, y, z) in enumerate(out):
        last[(x, y)] = k
    return [t for k, t in enumerate(out) if last[(t[0], t[1])] == k]


def gridlist_to_gridmatrix(gx, gy, gl):
    """-> (mat[dim_y, dim_x], row.names = diff(gy), colnames = diff(gx))."""
    dim_x, dim_y = len(gx) - 1, len(gy) - 1
    gl = sorted(gl, key=lambda t: t[0])            # order(grid_list$x), stable
    gl = remove_duplicates_triple(gl)
    xs = {t[0] for t in gl}
    ys = {t[1] for t in gl}
    for xm in range(1, dim_x + 1):
        if xm not in xs:
            gl.append((xm, 1, 0.0))
    for ym in range(1, dim_y + 1):
        if ym not in ys:
            gl.append((1, ym, 0.0))
    gl = remove_duplicates_triple(gl)
    mat = np.zeros((dim_y, dim_x), dtype=np.float64)
    for (x, y, z) in gl:
        mat[y - 1, x - 1] = z
    return mat, np.diff(np.asarray(gy, dtype=np.float64)), np.diff(np.asarray(gx, dtype=np.float64))


def paf_to_gridmatrix(ts, te, qs, qe):
    """make_xy_grid_fast -> bressiwrap -> order by x -> gridlist_to_gridmatrix (R/tsv_to_bitlocus.R:186-215)."""
    gx, gy, conv = make_xy_grid_fast(ts, te, qs, qe)
    gl = bressiwrap(ts, te, qs, qe, gx, gy)
    mat, ly, lx = gridlist_to_gridmatrix(gx, gy, gl)
    return dict(gridlines_x=gx, gridlines_y=gy, mat=mat, len_y=ly, len_x=lx, converged=conv)


def matrix_to_paf(mat, len_x, len_y):
    """Test helper (not reference code): the alignments a signed gridmatrix stands for -- one slope +-1 segment per
    non-zero cell."""
    ox = np.concatenate([[0], np.cumsum(len_x)]).astype(np.float64)
    oy = np.concatenate([[0], np.cumsum(len_y)]).astype(np.float64)
    ts, te, qs, qe = [], [], [], []
    for y in range(mat.shape[0]):
        for x in range(mat.shape[1]):
            if mat[y, x] == 0:
                continue
            ts.append(ox[x])
            te.append(ox[x + 1])
            if mat[y, x] > 0:
                qs.append(oy[y])
                qe.append(oy[y + 1])
            else:
                qs.append(oy[y + 1])
                qe.append(oy[y])
    return [np.array(v, dtype=np.float64) for v in (ts, te, qs, qe)]
