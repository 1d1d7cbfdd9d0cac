"""CPU restatement of wrapper_paf_to_bitlocus, the retry loop joining the PAF preparation and the gridding.

TEST INFRASTRUCTURE ONLY (imported by tests/, never by nahrwhals_b200/).  PARITY UNPINNED: R is not installed here.

    wrapper_paf_to_bitlocus   R/tsv_to_bitlocus.R:151-241
    update_params             R/tsv_to_bitlocus.R:111-129
    is_cluttered_paf          R/explore_mutation_space_v2_helpers.R:231-256
    determine_chunklen / determine_compression   R/wrapper_aln_and_analyse.R:269-307
The pieces it joins are the other restatements: nw_paf_oracle.load_and_prep_paf, nw_grid_oracle.make_xy_grid_fast /
bressiwrap / gridlist_to_gridmatrix, nw_region_oracle.find_sv_opportunities.
"""
import math

import numpy as np

from . import nw_grid_oracle as go
from . import nw_paf_oracle as po
from . import nw_region_oracle as ro


def update_params(minlen, compression):
    next_power_of_10 = 10.0 ** (math.floor(math.log10(minlen)) + 1)
    if minlen >= next_power_of_10:
        new_val = next_power_of_10 * 2
    else:
        new_val = minlen * 2
    if new_val >= next_power_of_10:
        new_val = next_power_of_10
    return new_val, new_val


def is_cluttered_paf(p, clutter_limit_per_border=5):
    n = len(p["qstart"])
    if n == 0:
        return False
    clutter = [int(np.sum(p["qstart"] == 0)), int(np.sum(p["qend"] == np.max(p["qend"]))),
               int(np.sum(p["tstart"] == 0)), int(np.sum(p["tend"] == np.max(p["tend"])))]
    seems_simplistic = (len(np.unique(p["qstart"])) < n / 2) or (len(np.unique(p["tstart"])) < n / 2) or \
                       (len(np.unique(p["qend"])) < n / 2) or (len(np.unique(p["qstart"])) < n / 2)
    return any(c >= clutter_limit_per_border for c in clutter) and seems_simplistic


def _which_min_value(interval_len, cutoffs, values):
    dists = np.array(cutoffs, dtype=np.float64) - interval_len
    dists[dists < 0] = np.inf
    kept = dists[dists > 0]                      # which.min indexes the FILTERED vector, values the unfiltered one
    return values[int(np.argmin(kept))] if len(kept) else None


def determine_chunklen(start, end):
    return _which_min_value(end - start, [10000, 500000, 1000000, 2000000, 5000000, 1e10],
                            [100, 1000, 10000, 10000, 10000, 50000])


def determine_compression(start, end):
    return _which_min_value(end - start, [50000, 500000, 5000000, 1e10], [100, 1000, 10000, 20000])


def wrapper_paf_to_bitlocus(paf, minlen, compression, chunklen=10000.0, max_n_alns=150, max_size_col_plus_rows=250,
                            max_pairs=2000, max_passes=64):
    """-> dict(status = 'ok' | 'cluttered' | 'empty' | 'passes', minlen, compression, passes, n_alns, n_pairs, paf, grid)"""
    info = dict(status=None, passes=0, n_alns=0, n_pairs=-1, paf=None, grid=None)
    while True:
        if info["passes"] >= max_passes:
            info["status"] = "passes"
            return info
        info["passes"] += 1
        info["minlen"], info["compression"] = minlen, compression
        p = po.load_and_prep_paf(paf, minlen, compression, chunklen)
        n = len(p["qstart"])
        info["n_alns"] = n
        if n > max_n_alns:
            minlen, compression = update_params(minlen, compression)
            continue
        if n == 0:
            info["status"] = "empty"
            return info
        if is_cluttered_paf(p):
            info["status"] = "cluttered"
            return info
        gx, gy, conv = go.make_xy_grid_fast(p["tstart"], p["tend"], p["qstart"], p["qend"])
        if len(gx) + len(gy) > max_size_col_plus_rows:
            minlen, compression = update_params(minlen, compression)
            continue
        gl = go.bressiwrap(p["tstart"], p["tend"], p["qstart"], p["qend"], gx, gy)
        mat, ly, lx = go.gridlist_to_gridmatrix(gx, gy, gl)
        info["n_pairs"] = len(ro.find_sv_opportunities(ro.Bitl(mat, ly, lx)))
        if info["n_pairs"] > max_pairs:
            minlen, compression = update_params(minlen, compression)
            continue
        info.update(status="ok", paf=p, grid=dict(gridlines_x=gx, gridlines_y=gy, mat=mat, len_y=ly, len_x=lx, converged=conv))
        return info
