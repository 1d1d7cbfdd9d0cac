/* nwgrid.h -- C ABI (libnwsolve.so) of the gridding step upstream of the solver.
 *
 * Reference: wrapper_paf_to_bitlocus (R/tsv_to_bitlocus.R:151-241) turns the alignments of a region into the solver's
 * wire format in three steps, all interpreted R loops:
 *     make_xy_grid_fast(paf)            R/tsv_to_bitlocus.R:290-384   gridlines = fix point of projecting every known
 *                                                                      gridline through every alignment it crosses
 *     bressiwrap(paf, gxy) / bresenham  R/tsv_to_bitlocus.R:415-..., R/bresenham.R:18-129   grid cells every alignment
 *                                                                      walks through, z = +-(x spacing of the cell)
 *     gridlist_to_gridmatrix            R/gridlist_to_gridmatrix.R:8-51 (+ remove_duplicates_triple): the n_y x n_x
 *                                                                      signed matrix; a cell hit with two different z is 0
 * nw_grid_build does the three steps for a batch of regions on the GPU: one CTA per region keeps the sorted gridline
 * sets in shared memory and iterates the fix point; one thread per alignment then walks the grid and merges its cells
 * into the region's matrix with compare-and-swap (order independent, like the reference's duplicate rule).
 * An alignment is (tstart, tend, qstart, qend) in the reference's paf columns: target = x, query = y, tstart < tend,
 * qstart > qend for a reverse-strand alignment.  All coordinates are IEEE doubles and every expression is evaluated
 * as R evaluates it (no FMA contraction).  No CPU fallback.
 */
#ifndef NWGRID_H
#define NWGRID_H

#include <stdint.h>

#include "nwsolve.h"

#ifdef __cplusplus
extern "C" {
#endif

#define NW_GRID_MAX_LINES 1024 /* gridlines per axis a region may reach (the reference gives up at 250 in total,
                                  R/tsv_to_bitlocus.R:192 params$max_size_col_plus_rows, and retries coarser) */
#define NW_GRID_MAX_ALNS 512   /* alignments per region (params$max_n_alns is 150 by default) */

typedef struct nw_grid nw_grid;

typedef struct nw_paf {
  const double* tstart; /* [n] */
  const double* tend;
  const double* qstart;
  const double* qend;
  int64_t n;
} nw_paf;

/* n_regions alignment sets -> n_regions grids.  rcs[k] (may be NULL) is the status of region k: NW_OK, NW_ERR_RANGE
 * (more than NW_GRID_MAX_LINES gridlines on an axis or more than NW_GRID_MAX_ALNS alignments: the caller coarsens the
 * alignments and retries, as the reference's loop does) or NW_ERR_ARG (degenerate alignment, or a walk that leaves the
 * grid -- the reference stops with an error there); outs[k] is NULL for a failed region.  The call itself returns the
 * first CUDA / argument error, NW_OK otherwise. */
int nw_grid_build(nw_ctx* ctx, const nw_paf* pafs, int64_t n_regions, nw_grid** outs, int32_t* rcs);

/* n_x / n_y = blocks per axis (gridlines - 1); converged = make_xy_grid_fast reached its fix point
 * (log_collection$grid_inconsistency == F); generations = iterations it took. */
int nw_grid_dims(const nw_grid* g, int32_t* n_x, int32_t* n_y, int32_t* converged, int32_t* generations);
/* gridlines: n_x + 1 and n_y + 1 sorted coordinates (gxy[[1]], gxy[[2]]) */
int nw_grid_lines(const nw_grid* g, double* gridlines_x, double* gridlines_y);
/* gridmatrix as R holds it: mat[n_y * n_x] row-major by y, len_x = colnames = diff(gridlines_x), len_y = row.names */
int nw_grid_matrix(const nw_grid* g, double* mat, double* len_x, double* len_y);
void nw_grid_free(nw_grid* g);

#ifdef __cplusplus
}
#endif
#endif /* NWGRID_H */
