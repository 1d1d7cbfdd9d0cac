/* nwsolve.h -- C ABI of libnwsolve.so, the B200-native replacement for NAHRwhals' mutation-space solver.
 *
 * Each entry point names the reference interface it replaces.  Paths are relative to the reference tree;
 * "solver.jl" = inst/extdata/scripts/scripts_nw_main/solver.jl.
 *
 * Conventions: plain pointers + sizes, caller owns every input buffer, results are opaque handles freed by
 * the caller.  Return 0 on success, a negative NW_ERR_* otherwise; nw_last_error() gives the message of the
 * last failure on the calling thread.  A context is bound to one CUDA device and is single-threaded;
 * distinct contexts may be used from distinct threads or processes (one per GPU, or several per GPU).
 * There is no CPU fallback: every compute entry point fails with NW_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef NWSOLVE_H
#define NWSOLVE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NW_OK 0
#define NW_ERR_ARG (-1)    /* bad argument / shape */
#define NW_ERR_IO (-2)     /* cannot read or write a file */
#define NW_ERR_PARSE (-3)  /* malformed TSV (solver.jl:102, :128-130) */
#define NW_ERR_CUDA (-4)   /* CUDA runtime error or no device */
#define NW_ERR_RANGE (-5)  /* lengths not integral / exceed the int32 DP range */
#define NW_ERR_NOMEM (-6)  /* search does not fit device memory */
#define NW_ERR_COLLISION (-7) /* the dedup verifier found two different indices under one key, also with other hash bases */

#define NW_MOVE_DUP 0 /* MoveKind, solver.jl:34 */
#define NW_MOVE_DEL 1
#define NW_MOVE_INV 2

typedef struct nw_ctx nw_ctx;
typedef struct nw_result nw_result;

/* Command line of solver.jl (:756-791) as a struct.  maxwidth / maxreport < 0 mean `nothing`. */
typedef struct nw_opts {
  int32_t maxdepth;  /* -d, default 3 */
  int32_t maxdup;    /* -u, default 3 */
  int64_t maxwidth;  /* -w, default nothing (-1) */
  double minreport;  /* -r, default 0.95 */
  int64_t maxreport; /* -R, default nothing (-1) */
  int32_t keep_ids;  /* diagnostics: keep per-id cost/score arrays in the result (parity tests) */
  int32_t flags;     /* NW_FLAG_* */
} nw_opts;

#define NW_FLAG_GENERIC_DP 1 /* force the generic (non register-resident) DP kernel; parity cross-check */
#define NW_FLAG_NO_TRAJ 2    /* skip trajectory enumeration (benchmarking the search itself) */
#define NW_FLAG_PAIR_EXPAND 4 /* force the pair-testing expansion kernel (fallback for huge indices; parity cross-check) */
#define NW_FLAG_VERIFY_DEDUP 8 /* accepted for compatibility: verification is the default (see NW_FLAG_NO_VERIFY) */
#define NW_FLAG_DP32 128 /* keep every DP launch on the 32-bit register kernel (k_score_fast); without it a (region,
                            child length) bucket whose totals fit 15 bits runs on the packed 16-bit kernel (k_score_pair,
                            two candidates per thread).  Results are identical; parity cross-check */
#define NW_FLAG_DEBUG_WEAK_FP 16 /* TESTS ONLY: truncate fingerprints to 8 bits so that collisions occur */
/* Dedup exactness.  The reference dedups on the whole Vector{Int16} (solver.jl:81, :535).  Here a candidate is keyed in
 * the dedup table by 96 bits of a 124-bit polynomial fingerprint, and EVERY candidate that was merged into an already
 * known index is then compared with that index element by element on the device (both re-materialised through their
 * index maps).  A search whose comparison finds a difference is repeated with other hash bases (up to twice) and
 * fails with NW_ERR_COLLISION if that does not help, so a returned result is exact, not probabilistic.
 * NW_FLAG_NO_VERIFY skips the comparison (benchmarking the fingerprint-only path: two different indices would then be
 * merged silently with probability <= (4094 / 2^31)^3 per pair of distinct indices of a search). */
#define NW_FLAG_NO_VERIFY 32
#define NW_FLAG_DEBUG_WEAK_FP_ONCE 64 /* TESTS ONLY: as NW_FLAG_DEBUG_WEAK_FP, but only with the initial hash bases, so that
                                         the repeat of the search with other bases can be observed to succeed */

typedef struct nw_stats {
  int32_t n_levels;
  int32_t kernel_launches;  /* CUDA kernels launched by this solve */
  int64_t generated[16];    /* candidates emitted per level (every (parent,i,j,kind)) */
  int64_t scored[16];       /* unique new indices scored per level == reference slow_score calls */
  int64_t cells[16];        /* DP band cells evaluated per level (SURVEY 8d definition) */
  int64_t frontier[16];     /* parents expanded per level */
  float level_ms[16];       /* device time per level (CUDA events) */
  float total_ms;           /* device time of the whole search (events, excludes H2D of inputs) */
  int64_t n_ids;            /* unique indices incl. the root */
  double best;              /* best Float64 score over all children (solver.jl:660, 539-541) */
  int64_t h2d_bytes;        /* host->device bytes copied by this call */
  int64_t d2h_bytes;        /* device->host bytes copied by this call */
  float score_ms;           /* device time inside the banded-DP kernels (events on the launch stream) */
  int32_t score_launches;   /* number of banded-DP kernel launches */
  float host_ms[4];         /* host wall time: [0] region setup + upload, [1] search, [2] report, [3] keep_ids dump */
} nw_stats;

void nw_default_opts(nw_opts* o);
const char* nw_last_error(void);
const char* nw_version(void);

/* CUDA devices visible to the process (0 when there is none or the driver is missing). */
int nw_device_count(void);
/* One context per (process, GPU).  device < 0 picks LOCAL_RANK % device_count. */
int nw_ctx_create(int device, nw_ctx** out);
void nw_ctx_destroy(nw_ctx* ctx);

/* Replaces aln_bfs_scored + get_good_trajectories (solver.jl:644-686, :727-745) on an in-memory problem.
 * aln_sign: n_x*n_y int8 in {-1,0,1}, x-major (aln[x*n_y+y]) == transpose(sign(raw)) of solver.jl:103-106.
 * len_x / len_y: block lengths in bp (lengths[2] / lengths[1] of solver.jl:121-138). */
int nw_solve(nw_ctx* ctx, const int8_t* aln_sign, int32_t n_x, int32_t n_y, const int64_t* len_x,
             const int64_t* len_y, const nw_opts* opts, nw_result** out);

/* Replaces main() of solver.jl (:755-812) minus argument parsing: read_tsv + read_length_tsv + search +
 * write_trajectories.  This is what `julia solver.jl ... inmat inlens out` does for R/juliasolver.R:34-53.
 * On any failure no output file is left behind (written to out+".tmp" then renamed). */
int nw_solve_files(nw_ctx* ctx, const char* inmat_path, const char* inlens_path, const char* out_path,
                   const nw_opts* opts, nw_stats* stats_or_null);

/* Result accessors (trajectories in report order, solver.jl:739-743). */
int64_t nw_result_count(const nw_result* r);
/* score as Float32 (solver.jl:714), number of moves, pointer to n_moves*3 int32 (kind,start,stop; :44-48) */
int nw_result_get(const nw_result* r, int64_t k, float* score, int32_t* n_moves, const int32_t** moves3);
/* DP terminal cost / total in bp of the index trajectory k leads to (cost < 0: early exit, no DP ran) */
int nw_result_cost(const nw_result* r, int64_t k, int64_t* cost_bp, int64_t* total_bp, int64_t* index_id);
/* Bulk form of nw_result_get / nw_result_cost: arrays of nw_result_count() entries (any pointer may be NULL);
 * moves3 holds nw_result_move_count() (kind,start,stop) triples, move_off[k] is the first triple of trajectory k. */
int nw_result_arrays(const nw_result* r, float* score, int32_t* n_moves, int64_t* move_off, int64_t* cost_bp,
                     int64_t* total_bp, int64_t* index_id, int32_t* moves3);
int64_t nw_result_move_count(const nw_result* r);
/* The results of a whole regionfile run in two calls (nw_solve_batch / nw_solve_many hand back n handles):
 * per-result sizes, then every array of nw_result_arrays / nw_result_tsv / nw_result_stats concatenated in result
 * order into caller buffers of the summed sizes (move_off stays relative to each result's own first triple). */
int nw_results_sizes(nw_result* const* rs, int64_t n, int64_t* n_traj, int64_t* n_moves, int64_t* tsv_len);
int nw_results_gather(nw_result* const* rs, int64_t n, float* score, int32_t* n_moves, int64_t* move_off,
                      int64_t* cost_bp, int64_t* total_bp, int64_t* index_id, int32_t* moves3, char* tsv,
                      nw_stats* stats);
/* The exact bytes write_trajectories (solver.jl:747-753) would write. */
int nw_result_tsv(const nw_result* r, const char** data, int64_t* len);
int nw_result_stats(const nw_result* r, nw_stats* out);
/* keep_ids diagnostics: per id (1-based, root = 1; arrays of n_ids): Float32 score (-inf if unreachable),
 * DP cost in bp (-1 early exit, -2 unreachable), total in bp (0 on early exit), discovery level. */
int nw_result_ids(const nw_result* r, float* score, int64_t* cost_bp, int64_t* total_bp, int32_t* level);
/* keep_ids diagnostics: all provenance edges in discovery order: child id, parent id, (kind,start,stop). */
int64_t nw_result_edge_count(const nw_result* r);
int nw_result_edges(const nw_result* r, int32_t* child, int32_t* parent, int32_t* moves3);
void nw_result_free(nw_result* r);

/* Replaces slow_score (solver.jl:329-376) for a raw list of candidate indices (flat int16 + offsets).
 * force == 1 selects the R twin's forcecalc branch (R/evaltools.R:203-205; corridor 1.0, no early exit) that
 * produces res_ref; force == 2 the R twin without forcecalc (R/evaltools.R:193-201: corridor chosen by n_y, the row
 * count of R's matrix, and no early exit) that the depth-1 pre-pass uses (nwregion.h).  Outputs per candidate: cost_bp (-1 early exit, -2 unreachable), total_bp, Float64 score. */
int nw_score_batch(nw_ctx* ctx, const int8_t* aln_sign, int32_t n_x, int32_t n_y, const int64_t* len_x,
                   const int64_t* len_y, const int16_t* flat, const int64_t* off, int64_t n, int32_t force,
                   int32_t flags, int64_t* cost_bp, int64_t* total_bp, double* score);

/* Replaces to_moveset (solver.jl:217-231): two n_x*n_x uint8 matrices (duplicate_delete, invert). */
int nw_moveset(nw_ctx* ctx, const int8_t* aln_sign, int32_t n_x, int32_t n_y, uint8_t* dup_del, uint8_t* invert);

/* Regionfile / whole-genome modes (R/nahrwhals.R:171-215, :240-290 run one solver process per region from PSOCK
 * workers): n independent problems are solved by n_ctx contexts concurrently, one host thread + CUDA stream per
 * context, regions handed out dynamically.  Contexts may sit on one GPU (small regions overlap on the device) or on
 * several GPUs of the process.  outs[k] / rcs[k] receive the result handle / return code of problem k. */
typedef struct nw_problem {
  const int8_t* aln_sign; /* n_x*n_y, x-major */
  int32_t n_x, n_y;
  const int64_t* len_x;
  const int64_t* len_y;
} nw_problem;
int nw_solve_batch(nw_ctx** ctxs, int32_t n_ctx, const nw_problem* probs, int64_t n, const nw_opts* opts,
                   nw_result** outs, int32_t* rcs);

/* The same n independent problems on ONE context, max_regions_per_pass (<= 0: 256) regions at a time searched
 * TOGETHER: their frontiers are concatenated, every kernel launch of a BFS level covers all regions of the pass
 * (candidates carry their region through the parent; the dedup table is shared, fingerprints are salted with the
 * region id; ids, beam selection, best score and report are per region).  Small regions are launch-bound when
 * solved one by one; this path amortises every launch over the pass.  Results are identical, region by region, to
 * nw_solve.  On error no result is returned (all outs[k] are NULL). */
int nw_solve_many(nw_ctx* ctx, const nw_problem* probs, int64_t n, const nw_opts* opts, int32_t max_regions_per_pass,
                  nw_result** outs);

/* nw_solve_many over n_ctx contexts (one host thread + CUDA stream each, on one GPU or several): the passes are
 * handed out dynamically, so the host-side preparation and report of one pass overlap the device work of another.
 * Results are identical to nw_solve_many / nw_solve; on error no result is returned. */
int nw_solve_many_ctxs(nw_ctx** ctxs, int32_t n_ctx, const nw_problem* probs, int64_t n, const nw_opts* opts,
                       int32_t max_regions_per_pass, nw_result** outs);

/* File form of a regionfile run: region k is read from inmat[k] / inlens[k] (as nw_solve_files), the parsed regions
 * go through nw_solve_many_ctxs, and report k is written to outp[k] (complete or not at all).  rcs (may be NULL)
 * receives one return code per region: a region that does not parse, is out of range or cannot be written fails
 * alone -- if the batch as a whole is refused, its regions are solved one by one.  Returns the first error. */
int nw_solve_files_many(nw_ctx** ctxs, int32_t n_ctx, const char* const* inmat, const char* const* inlens,
                        const char* const* outp, int64_t n, const nw_opts* opts, int32_t max_regions_per_pass,
                        int32_t* rcs);

/* ---- one deep search sharded over `world` ranks (config 3 of BASELINE.json) -------------------------
 * The reference's level loop (solver.jl:667-684) is serial and its cost is the scorer (slow_score, :329-426, called
 * once per new index at :536).  Every rank runs the SAME deterministic expansion, dedup, id assignment and frontier
 * selection (replicated: no exchange, identical state everywhere); the banded DP of a level's fresh candidates is
 * divided between the ranks by contiguous ranges of discovery rank, and ONE all-gather per level over the score
 * arrays (12 bytes per fresh candidate, in place, on the context's stream, no host synchronisation) completes them.
 * Results -- every id, score, edge and the report -- are bit-identical to nw_solve on every rank.
 *
 * Multi-process (one rank per GPU, e.g. under torchrun / mpirun): rank 0 calls nw_dist_unique_id and hands the
 * NW_DIST_ID_BYTES bytes to the other ranks by any means (torch.distributed broadcast, MPI_Bcast, a file); every rank
 * then calls nw_dist_init (collective: ncclCommInitRank on the context's device) once, and nw_solve_sharded
 * (collective: the same problem and options on every rank) per search.  NCCL is bound at run time (libnccl.so.2);
 * without it these calls fail with NW_ERR_CUDA and nw_last_error() says why.  A rank that fails inside a collective
 * search leaves its peers waiting in NCCL -- run sharded jobs under the launcher's timeout.
 * nw_stats of a sharded search: counters are those of the whole search (cells summed over the ranks); times,
 * launches and byte counts are this rank's. */
#define NW_DIST_ID_BYTES 128
int nw_dist_unique_id(void* id_out /* NW_DIST_ID_BYTES */);
int nw_dist_init(nw_ctx* ctx, const void* id /* NW_DIST_ID_BYTES */, int32_t rank, int32_t world);
int nw_dist_info(nw_ctx* ctx, int32_t* rank, int32_t* world); /* 0 / 1 without a communicator */
void nw_dist_shutdown(nw_ctx* ctx);
int nw_solve_sharded(nw_ctx* ctx, const int8_t* aln_sign, int32_t n_x, int32_t n_y, const int64_t* len_x,
                     const int64_t* len_y, const nw_opts* opts, nw_result** out);
/* The same inside ONE process: ctxs[r] is rank r (one host thread each).  The contexts may sit on different GPUs
 * (the exchange is a set of peer copies over NVLink) or share a GPU (how the parity tests emulate several ranks on one
 * device).  outs[r] receives rank r's result (all identical); on error no result is returned.  This is what
 * `nwsolve --devices 0,1,...` runs for a single search, so the file + command-line boundary R uses
 * (R/juliasolver.R:34-53) reaches several GPUs without a launcher. */
int nw_solve_sharded_ctxs(nw_ctx** ctxs, int32_t world, const int8_t* aln_sign, int32_t n_x, int32_t n_y,
                          const int64_t* len_x, const int64_t* len_y, const nw_opts* opts, nw_result** outs);
/* File form (main() of solver.jl, :755-812, over several GPUs): TSVs in, rank 0's report out, complete or not at all. */
int nw_solve_files_sharded(nw_ctx** ctxs, int32_t world, const char* inmat_path, const char* inlens_path,
                           const char* out_path, const nw_opts* opts, nw_stats* stats_or_null);

/* Roofline denominators measured on this device with dependency-free micro-kernels:
 * out[0] VIADDMNMX, out[1] IADD3, out[2] VIADDMNMX+IMAD interleaved -- thread-instructions per second;
 * out[3] device-to-device copy bandwidth in bytes/s (read + write). */
int nw_bench_peaks(nw_ctx* ctx, double out[4]);

/* TSV readers of solver.jl:101-107 / :121-138 (host side; exposed for the bindings). Caller frees with nw_free. */
int nw_read_problem(const char* inmat_path, const char* inlens_path, int8_t** aln_sign, int32_t* n_x,
                    int32_t* n_y, int64_t** len_x, int64_t** len_y);
void nw_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* NWSOLVE_H */
