// CPU harness of the region stage's host logic (nahrwhals_b200/csrc/nw_region_host.h compiled without CUDA).
// TEST INFRASTRUCTURE: the per-child numbers the GPU phase would deliver (estimate sums, DP cost / total / score) are
// passed in by the test, which takes them from the R restatement, so that everything around the kernels -- flip,
// find_sv_opportunities, carry_out_compressed_sv, the visited set, the sequential filters, sorting, annotation,
// format_julia_output, merge, logfile summary -- is checked on the CPU.
#include "../../nahrwhals_b200/csrc/nw_region_host.h"

extern "C" {

// children of the (flipped) reference arrangement.  meta: 4 int32 per child (p1, p2, sv, length); blk / len: the
// positions of all children, flattened.  flags: cluttered, flipped, flip_unsure, maxdepth.  Returns the child count or -1.
int64_t hr_candidates(const int64_t* grid, int nx, int ny, const int64_t* lx, const int64_t* ly, int32_t* meta,
                      int64_t cap_meta, int16_t* blk, int64_t* len, int64_t cap_flat, int32_t* flags) {
  try {
    PrepassOut P;
    prepass_prepare(P, make_grid(grid, nx, ny, lx, ly));
    flags[0] = P.cluttered;
    flags[1] = P.did_flip;
    flags[2] = P.unsure;
    flags[3] = P.maxdepth;
    int64_t f = 0;
    if ((int64_t)P.cand.size() > cap_meta) return -1;
    for (size_t k = 0; k < P.cand.size(); ++k) {
      meta[k * 4 + 0] = P.moves[k].p1;
      meta[k * 4 + 1] = P.moves[k].p2;
      meta[k * 4 + 2] = P.moves[k].sv;
      meta[k * 4 + 3] = (int32_t)P.cand[k].size();
      for (const Pos& p : P.cand[k]) {
        if (f >= cap_flat) return -1;
        blk[f] = p.blk;
        len[f] = p.len;
        ++f;
      }
    }
    return (int64_t)P.cand.size();
  } catch (const RErr&) {
    return -2;
  }
}

// the whole host flow of nw_region_analyse with the GPU numbers supplied: entry 0 = the reference arrangement (forced
// corridor), entry k = child k.  solver_tsv: the report the solver would write for this region (NULL: fail with 2 if
// the region needs the solver).  Returns 0, or 1 = buffer too small, 2 = solver report needed, 3 = bad entry count,
// negative = NW_ERR_*.
int hr_region(const int64_t* grid, int nx, int ny, const int64_t* lx, const int64_t* ly, const double* gridlines, int ngl,
              int depth, int maxdup, double init_width, double minreport, double compression, const int64_t* srow,
              const int64_t* scol, const int64_t* cost, const int64_t* total, const double* score, int64_t n_entries,
              const char* solver_tsv, int64_t tsv_len, char* out, int64_t cap, char* mut_max, int64_t cap_mm,
              nw_region_summary* summ) {
  try {
    nw_region_opts o{depth, maxdup, init_width, minreport, compression, 0, 0};
    PrepassOut P;
    prepass_prepare(P, make_grid(grid, nx, ny, lx, ly));
    if (!P.cluttered && !P.trivial) {
      if (n_entries != (int64_t)P.cand.size() + 1) return 3;
      for (int64_t k = 0; k < n_entries; ++k) {
        P.srow[k] = srow[k];
        P.scol[k] = scol[k];
        P.ccost[k] = cost[k];
        P.ctotal[k] = total[k];
        P.cscore[k] = score[k];
      }
    }
    prepass_replay(P, o);
    nw_region_summary S;
    memset(&S, 0, sizeof S);
    fill_prepass_summary(S, P);
    if (needs_solver(P)) {
      S.solver_called = 1;
      guarded_width(P, init_width, S.width_reduced);
      if (!solver_tsv) return 2;
      std::unique_ptr<nw_table> J(format_solver_output(solver_tsv, tsv_len, gridlines, ngl));
      apply_solver_table(P, S, J, o);
    }
    finish_region(P.table, S);
    std::string t;
    render_tsv(*P.table, t);
    if ((int64_t)t.size() + 1 > cap || (int64_t)P.table->mut_max.size() + 1 > cap_mm) return 1;
    memcpy(out, t.c_str(), t.size() + 1);
    memcpy(mut_max, P.table->mut_max.c_str(), P.table->mut_max.size() + 1);
    *summ = S;
    return 0;
  } catch (const RErr& e) {
    return e.code;
  }
}

// round(x, 3) as the region stage computes it (nw_region_host.h round3)
double hr_round3(double x) { return round3(x); }

}  // extern "C"
