/* nwbitlocus.h -- C ABI (libnwsolve.so) of wrapper_paf_to_bitlocus: the retry loop that joins the PAF preparation
 * (include/nwpaf.h) and the gridding (include/nwgrid.h) (SURVEY.md 8f row 2 "recompression loop").
 *
 * Reference: wrapper_paf_to_bitlocus (R/tsv_to_bitlocus.R:151-241) repeats, with (minlen, compression) made coarser
 * by update_params (:111-129) after every rejected pass:
 *     paf <- load_and_prep_paf_for_gridplot_transform(inpaf, minlen, compression, chunklen)       :168
 *     more than params$max_n_alns alignments                                   -> coarser, again  :171-176
 *     is_cluttered_paf(paf) (R/explore_mutation_space_v2_helpers.R:231-256)    -> give up (NULL)  :179-182
 *     gxy <- make_xy_grid_fast(paf); more than params$max_size_col_plus_rows gridlines in total
 *                                                                              -> coarser, again  :185-197
 *     grid_list <- bressiwrap(paf, gxy); more than 2000 rows in find_sv_opportunities(gridlist_to_gridmatrix(...))
 *     (R/find_sv_opportunities.R:12-98)                                        -> coarser, again  :210-222
 *     accepted: list(gridlines.x, gridlines.y, grid_list)
 * nw_bitlocus_build runs that loop for a batch of regions: every pass prepares the pending regions' alignment tables
 * (k_paf_pairs), grids all of them in ONE nw_grid_build batch, counts the SV opportunities of the matrices and
 * requeues the rejected regions with their own coarser parameters.  From the parsed PAF table on (columns as minimap2
 * writes them, qstart < qend); plots are not produced.  No CPU fallback.
 */
#ifndef NWBITLOCUS_H
#define NWBITLOCUS_H

#include <stdint.h>

#include "nwgrid.h"
#include "nwpaf.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nw_bitlocus_params {
  double minlen;                  /* params$minlen */
  double compression;             /* params$compression */
  double chunklen;                /* params$chunklen */
  int32_t max_n_alns;             /* params$max_n_alns (R/nahrwhals.R:55: 150) */
  int32_t max_size_col_plus_rows; /* params$max_size_col_plus_rows (250) */
  int32_t max_pairs;              /* 2000 in the reference (R/tsv_to_bitlocus.R:216) */
  int32_t max_passes;             /* safety bound on the loop the reference leaves unbounded; <= 0: 64 */
} nw_bitlocus_params;

typedef struct nw_bitlocus_info {
  double minlen;       /* parameters of the last pass (the accepted one on success: log_collection$compression) */
  double compression;
  int32_t passes;      /* times the loop body ran */
  int32_t cluttered;   /* 1: is_cluttered_paf stopped the region (the reference returns NULL); no grid */
  int32_t n_alns;      /* rows of the prepared table of the last pass */
  int32_t n_pairs;     /* nrow(find_sv_opportunities(gridmatrix)) of the last pass that got that far, else -1 */
} nw_bitlocus_info;

void nw_bitlocus_default_params(nw_bitlocus_params* p);
/* chunklen / minlen / compression = "default" (R/nahrwhals_coremodule.R:120-128): determine_chunklen and
 * determine_compression (R/wrapper_aln_and_analyse.R:269-307) of the region's reference interval */
void nw_bitlocus_auto_params(double start_x, double end_x, nw_bitlocus_params* p);

/* n_regions parsed PAF tables -> grids (and, if tables != NULL, the prepared tables of the accepted pass).
 * rcs[k] (may be NULL): NW_OK -- grids[k] is set, or infos[k].cluttered == 1 and grids[k] is NULL; NW_ERR_ARG -- no
 * alignment survives (the reference stops in make_xy_grid_fast) or a degenerate table; NW_ERR_RANGE -- max_passes
 * reached or the region exceeds the gridding limits at every scale.  infos may be NULL.  The call returns the first
 * CUDA / argument error of the batch itself, NW_OK otherwise. */
int nw_bitlocus_build(nw_ctx* ctx, const nw_paf_cols* pafs, int64_t n_regions, const nw_bitlocus_params* params,
                      nw_grid** grids, nw_paf_table** tables, nw_bitlocus_info* infos, int32_t* rcs);

#ifdef __cplusplus
}
#endif
#endif /* NWBITLOCUS_H */
