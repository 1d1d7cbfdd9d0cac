/* nwpaf.h -- C ABI (libnwsolve.so) of the PAF preparation upstream of the gridding.
 *
 * Reference: load_and_prep_paf_for_gridplot_transform (R/tsv_to_bitlocus.R:15-110) turns the chunked minimap2
 * alignments of a region into the table make_xy_grid_fast reads:
 *     strand swap (qstart > qend for '-' alignments), compress_paf_fnct(second_run = T) (R/compress_paf.R:173-406:
 *     merge_paf_entries_intraloop finds the pairs of alignments where one ends where the other starts -- an all-pairs
 *     test per target-axis chunk --, the best partner per row / column is kept, merge_rows stitches them),
 *     alen > minlen, rounding to the compression grid, enforce_slope_one (:258-275).
 * The all-pairs test runs on the GPU (one thread per ordered pair of alignments, every chunk both touch, the chunk's
 * own tolerance); the selection and the sequential stitching of merge_rows run on the host in the reference's order.
 * A table is passed column-wise as doubles; strand is +1 / -1.  The qname / tname strings influence no number
 * downstream and are not carried.  No CPU fallback.
 */
#ifndef NWPAF_H
#define NWPAF_H

#include <stdint.h>

#include "nwsolve.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nw_paf_cols {
  const double* qlen;
  const double* qstart;
  const double* qend;
  const double* strand; /* +1 or -1 */
  const double* tlen;
  const double* tstart;
  const double* tend;
  const double* nmatch;
  const double* alen;
  const double* mapq;
  int64_t n;
} nw_paf_cols;

typedef struct nw_paf_table nw_paf_table;

/* Replaces compress_paf_fnct(inpaf_df = paf, save_outpaf = F, second_run, inparam_chunklen, inparam_compression,
 * n_quadrants_per_axis) (R/compress_paf.R:173-406).  n_quadrants <= 0: the reference's choice (2 below 50 kb, else 10). */
int nw_paf_compress(nw_ctx* ctx, const nw_paf_cols* in, int32_t second_run, double chunklen, double compression,
                    int32_t n_quadrants, nw_paf_table** out);

/* Replaces load_and_prep_paf_for_gridplot_transform(inpaf, minlen, compression, inparam_chunklen) from the parsed
 * table on (R/tsv_to_bitlocus.R:38-82); `in` holds the PAF columns as minimap2 writes them (qstart < qend). */
int nw_paf_prepare(nw_ctx* ctx, const nw_paf_cols* in, double minlen, double compression, double chunklen,
                   nw_paf_table** out);

/* nw_paf_prepare for a batch of tables (the regions of a regionfile): the all-pairs tests of all tables run in ONE
 * kernel launch (one upload, one download).  rcs[k] (may be NULL) is the status of table k (NW_ERR_ARG for a malformed
 * table; outs[k] is NULL then); the call returns the first CUDA / argument error of the batch itself. */
int nw_paf_prepare_many(nw_ctx* ctx, const nw_paf_cols* ins, int64_t n_tables, double minlen, double compression,
                        double chunklen, nw_paf_table** outs, int32_t* rcs);

/* Replaces compress_paf_fnct(inpaf_link = inpaf, outpaf_link = outpaf, inparam_chunklen = chunklen)
 * (R/compress_paf.R:173-406 with save_outpaf = T: the "melt" step after the chunked minimap2 run,
 * R/seqbuilder_functions.R:107): read.table + unique(), strand swap, merge (first run: tolerance 0.05 * chunklen),
 * merged rows get the rebuilt qname of merge_rows (:49-59), write.table (numbers as R prints them; a table in which
 * rows were merged has the 12 PAF columns only, nmatch / alen as doubles).  The output file is written to a temporary
 * name and renamed.  n_quadrants <= 0: the reference's choice. */
int nw_paf_compress_file(nw_ctx* ctx, const char* inpaf_path, const char* outpaf_path, double chunklen, int32_t n_quadrants,
                         int64_t* rows_in, int64_t* rows_out);

/* Replaces the loop of whole-genome mode's all-vs-all melt (R/make_whole_genome_list.R:80-106): read.table, the chunk
 * suffix "_<start>-<end>" stripped from qname, strand swap, and for every query sequence in unique(paf$qname) order
 * compress_paf_fnct(inpaf_df = paf, outpaf_link, inparam_chunklen, n_quadrants_per_axis, qname = q): that sequence's
 * rows melted with reduced_qnames (a merged row keeps qname1) and APPENDED to outpaf_path (append = TRUE in the
 * reference; the file is not truncated).  The all-pairs tests of all query sequences run in one kernel launch.
 * The reference passes n_quadrants = 10. */
int nw_paf_compress_file_by_qname(nw_ctx* ctx, const char* inpaf_path, const char* outpaf_path, double chunklen,
                                  int32_t n_quadrants, int64_t* rows_in, int64_t* rows_out);

/* read.table(inpaf) of load_and_prep_paf_for_gridplot_transform (R/tsv_to_bitlocus.R:17-31): the numeric columns of a
 * PAF file (12 columns + optional tags) as a table; strand '+' / '-' becomes +1 / -1.  Host only. */
int nw_paf_read_file(const char* path, nw_paf_table** out);

int64_t nw_paf_rows(const nw_paf_table* t);
/* any pointer may be NULL; arrays of nw_paf_rows() doubles */
int nw_paf_columns(const nw_paf_table* t, double* qlen, double* qstart, double* qend, double* strand, double* tlen,
                   double* tstart, double* tend, double* nmatch, double* alen, double* mapq);
/* pairs the all-pairs test found (before selection) and kernels launched: diagnostics */
int nw_paf_stats(const nw_paf_table* t, int64_t* pairs_found, int32_t* gpu_launches);
void nw_paf_free(nw_paf_table* t);

#ifdef __cplusplus
}
#endif
#endif /* NWPAF_H */
