/* nwpipeline.h -- C ABI (libnwsolve.so): regionfile mode from the alignments on, in one call.
 *
 * Reference: the part of wrapper_aln_and_analyse that follows minimap2 (R/wrapper_aln_and_analyse.R:130-203), once per
 * region from the PSOCK workers of regionfile / whole-genome mode (R/nahrwhals.R:171-215, :240-290):
 *     grid_xy <- wrapper_condense_paf(params, outlinks)          wrapper_paf_to_bitlocus        -> nw_bitlocus_build
 *     is.null(grid_xy) (cluttered PAF): res_empty, write_results (:132-140)                     -> the empty row
 *     gridmatrix; is_cluttered; solve_mutation(maxdepth = 1); solver below 99.8; merge (:160-191) -> nw_region_analyse_many
 *     write_results -> save_to_logfile (:201, R/save_logfile.R)                                 -> nw_region_logrow
 * nw_regions_run composes those entry points for a batch of regions (nothing here that they do not do; it exists so that a
 * host in any language pays one call per regionfile instead of three per region).  Plots are not produced.
 */
#ifndef NWPIPELINE_H
#define NWPIPELINE_H

#include <stdint.h>

#include "nwbitlocus.h"
#include "nwregion.h"

#ifdef __cplusplus
extern "C" {
#endif

/* pafs[k] / metas[k]: the parsed PAF table and the identity of region k (metas[k].grid_inconsistency is ignored: it is
 * taken from the gridding, log_collection$grid_inconsistency, R/tsv_to_bitlocus.R:292-376; metas[k].compression is kept
 * as given).  ropts->compression is ignored: the step penalty of the pre-pass is params$compression = bp->compression.
 * status[k] (may be NULL): 0 = row of the analysed region, 1 = row of a region whose PAF is cluttered (res_empty),
 * negative = the region was dropped (error code; the reference loses such a region in its tryCatch).
 * *rows: malloc'ed text (release with nw_free), one '\n'-terminated line of nahrwhals_res.tsv per region with
 * status >= 0, in region order; row_off (may be NULL): n + 1 offsets into it (row k = [row_off[k], row_off[k+1])).
 * logfile (may be NULL): the rows are also appended there (header first if the file is empty), under a lock. */
int nw_regions_run(nw_ctx* ctx, const nw_paf_cols* pafs, const nw_log_meta* metas, int64_t n,
                   const nw_bitlocus_params* bp, const nw_region_opts* ropts, const char* logfile, char** rows,
                   int64_t* rows_len, int64_t* row_off, int32_t* status);

/* nw_regions_run_ctxs with the alignments read from PAF files (12 columns + optional tags, as minimap2 / the melt step
 * write them): paf_paths[k] is read like the read.table(inpaf) of load_and_prep_paf_for_gridplot_transform
 * (R/tsv_to_bitlocus.R:17).  A region whose file cannot be read or parsed is dropped (status[k] = NW_ERR_IO /
 * NW_ERR_PARSE), the others run. */
int nw_regions_run_files(nw_ctx** ctxs, int32_t n_ctx, const char* const* paf_paths, const nw_log_meta* metas, int64_t n,
                         const nw_bitlocus_params* bp, const nw_region_opts* ropts, int32_t regions_per_batch,
                         const char* logfile, char** rows, int64_t* rows_len, int64_t* row_off, int32_t* status);

/* The same over n_ctx contexts (one host thread + CUDA stream each; on one GPU -- the host phases of one batch overlap
 * the device phases of another -- or on several GPUs): batches of regions_per_batch (<= 0: 512) regions are handed out
 * dynamically; rows come back (and reach the logfile) in region order.  Results identical to nw_regions_run. */
int nw_regions_run_ctxs(nw_ctx** ctxs, int32_t n_ctx, const nw_paf_cols* pafs, const nw_log_meta* metas, int64_t n,
                        const nw_bitlocus_params* bp, const nw_region_opts* ropts, int32_t regions_per_batch,
                        const char* logfile, char** rows, int64_t* rows_len, int64_t* row_off, int32_t* status);

#ifdef __cplusplus
}
#endif
#endif /* NWPIPELINE_H */
