// nw_pipeline.cu -- regionfile mode from the alignments on (include/nwpipeline.h): a composition of the library's own C
// entry points (nw_bitlocus_build -> nw_region_analyse_many -> nw_region_logrow), written against the public headers only.
#include <atomic>
#include <cmath>
#include <mutex>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/nwpipeline.h"

namespace nw {
void set_last_error(const std::string& m);
}
extern "C" const char* nw_last_error(void);

extern "C" int nw_regions_run(nw_ctx* ctx, const nw_paf_cols* pafs, const nw_log_meta* metas, int64_t n,
                              const nw_bitlocus_params* bp, const nw_region_opts* ropts, const char* logfile, char** rows,
                              int64_t* rows_len, int64_t* row_off, int32_t* status) {
  if (rows) *rows = nullptr;
  if (rows_len) *rows_len = 0;
  if (!ctx || !bp || !ropts || !rows || !rows_len || n < 0 || (n > 0 && (!pafs || !metas))) {
    nw::set_last_error("null argument");
    return NW_ERR_ARG;
  }
  std::vector<nw_grid*> grids((size_t)n, nullptr);
  std::vector<nw_bitlocus_info> infos((size_t)n);
  std::vector<int32_t> st((size_t)n, NW_OK);
  std::vector<nw_table*> tables;
  auto cleanup = [&] {
    for (nw_grid* g : grids)
      if (g) nw_grid_free(g);
    for (nw_table* t : tables)
      if (t) nw_table_free(t);
  };
  int rc = nw_bitlocus_build(ctx, pafs, n, bp, grids.data(), nullptr, infos.data(), st.data());
  if (rc != NW_OK) {
    cleanup();
    return rc;
  }
  // gridmatrix <- gridlist_to_gridmatrix(grid_xy): the accepted grids as the solver stage's problems
  struct Prob {
    std::vector<int64_t> mat, lx, ly;
    std::vector<double> gx, gy;
    int32_t converged = 1;
  };
  std::vector<Prob> P((size_t)n);
  std::vector<int64_t> idx;
  std::vector<nw_region_problem> probs;
  for (int64_t k = 0; k < n; ++k) {
    if (st[k] != NW_OK || !grids[k]) continue;
    int32_t nx = 0, ny = 0, cv = 0, gen = 0;
    nw_grid_dims(grids[k], &nx, &ny, &cv, &gen);
    Prob& p = P[k];
    p.converged = cv;
    p.gx.resize((size_t)nx + 1);
    p.gy.resize((size_t)ny + 1);
    nw_grid_lines(grids[k], p.gx.data(), p.gy.data());
    std::vector<double> m((size_t)nx * ny), lx((size_t)nx), ly((size_t)ny);
    nw_grid_matrix(grids[k], m.data(), lx.data(), ly.data());
    nw_grid_free(grids[k]);
    grids[k] = nullptr;
    bool integral = true;
    auto conv = [&](const std::vector<double>& src, std::vector<int64_t>& dst) {
      dst.resize(src.size());
      for (size_t q = 0; q < src.size(); ++q) {
        dst[q] = (int64_t)std::llround(src[q]);
        if ((double)dst[q] != src[q]) integral = false;
      }
    };
    conv(m, p.mat);
    conv(lx, p.lx);
    conv(ly, p.ly);
    if (!integral) {  // the solver's wire format is integer bp (R/juliasolver.R:5-7)
      st[k] = NW_ERR_RANGE;
      nw::set_last_error("gridmatrix with non-integer lengths");
      continue;
    }
    idx.push_back(k);
    probs.push_back(nw_region_problem{p.mat.data(), nx, ny, p.lx.data(), p.ly.data(), p.gx.data(), nx + 1});
  }
  tables.assign(idx.size(), nullptr);
  std::vector<nw_region_summary> summ(idx.size());
  std::vector<int32_t> rrc(idx.size(), NW_OK);
  if (!idx.empty()) {
    nw_region_opts o = *ropts;
    o.compression = bp->compression;  // solve_mutation(..., compression = params$compression)
    rc = nw_region_analyse_many(ctx, probs.data(), (int64_t)probs.size(), &o, tables.data(), summ.data(), rrc.data());
    if (rc != NW_OK) {
      cleanup();
      return rc;
    }
  }
  std::vector<int64_t> slot((size_t)n, -1);
  for (size_t q = 0; q < idx.size(); ++q) {
    slot[idx[q]] = (int64_t)q;
    if (rrc[q] != NW_OK || !tables[q]) st[idx[q]] = rrc[q] != NW_OK ? rrc[q] : NW_ERR_ARG;
  }
  std::string text;
  std::vector<char> buf(8192);
  for (int64_t k = 0; k < n; ++k) {
    if (row_off) row_off[k] = (int64_t)text.size();
    int32_t out_status = st[k];
    if (st[k] == NW_OK) {
      nw_log_meta m = metas[k];
      nw_region_summary s;
      const char* mut_max = "ref";
      if (infos[k].cluttered) {  // res_empty <- data.frame(eval = 0, mut1 = "ref") (R/wrapper_aln_and_analyse.R:132-140)
        memset(&s, 0, sizeof s);
        s.res_ref = s.res_max = 0.0;
        s.n_res_max = 1;
        s.search_depth = ropts->depth;
        s.cluttered = 1;
        for (double& v : s.maxres) v = std::numeric_limits<double>::quiet_NaN();
        m.grid_inconsistency = 0;
        out_status = 1;
      } else {
        s = summ[(size_t)slot[k]];
        mut_max = nw_table_mut_max(tables[(size_t)slot[k]]);
        m.grid_inconsistency = P[k].converged ? 0 : 1;
      }
      int64_t len = 0;
      int r = nw_region_logrow(&s, mut_max, &m, buf.data(), (int64_t)buf.size(), &len);
      if (r == NW_ERR_RANGE) {
        buf.resize((size_t)len + 1);
        r = nw_region_logrow(&s, mut_max, &m, buf.data(), (int64_t)buf.size(), &len);
      }
      if (r != NW_OK) {
        out_status = r;
      } else {
        if (logfile) {
          const int a = nw_region_log_append(logfile, buf.data());
          if (a != NW_OK) {
            cleanup();
            return a;
          }
        }
        text.append(buf.data(), (size_t)len);
        text += '\n';
      }
    }
    if (status) status[k] = out_status;
  }
  if (row_off) row_off[n] = (int64_t)text.size();
  cleanup();
  char* out = (char*)malloc(text.size() + 1);
  if (!out) {
    nw::set_last_error("host out of memory");
    return NW_ERR_NOMEM;
  }
  memcpy(out, text.c_str(), text.size() + 1);
  *rows = out;
  *rows_len = (int64_t)text.size();
  return NW_OK;
}

extern "C" int nw_regions_run_ctxs(nw_ctx** ctxs, int32_t n_ctx, const nw_paf_cols* pafs, const nw_log_meta* metas, int64_t n,
                                   const nw_bitlocus_params* bp, const nw_region_opts* ropts, int32_t regions_per_batch,
                                   const char* logfile, char** rows, int64_t* rows_len, int64_t* row_off, int32_t* status) {
  if (rows) *rows = nullptr;
  if (rows_len) *rows_len = 0;
  if (!ctxs || n_ctx < 1 || !rows || !rows_len || n < 0) {
    nw::set_last_error("null argument");
    return NW_ERR_ARG;
  }
  for (int t = 0; t < n_ctx; ++t)
    if (!ctxs[t]) {
      nw::set_last_error("null context");
      return NW_ERR_ARG;
    }
  const int64_t per = regions_per_batch > 0 ? regions_per_batch : 512;
  const int64_t nb = (n + per - 1) / per;
  if (n_ctx == 1 || nb <= 1) return nw_regions_run(ctxs[0], pafs, metas, n, bp, ropts, logfile, rows, rows_len, row_off, status);
  struct Batch {
    char* text = nullptr;
    int64_t len = 0;
    std::vector<int64_t> off;
  };
  std::vector<Batch> B((size_t)nb);
  std::vector<int32_t> st((size_t)n, NW_OK);
  std::atomic<int64_t> next{0};
  std::mutex mu;
  int first_rc = NW_OK;
  std::string first_msg;
  auto work = [&](int t) {
    for (;;) {
      const int64_t b = next.fetch_add(1);
      if (b >= nb) return;
      const int64_t a = b * per, m = std::min<int64_t>(per, n - a);
      B[b].off.assign((size_t)m + 1, 0);
      const int rc = nw_regions_run(ctxs[t], pafs + a, metas + a, m, bp, ropts, nullptr, &B[b].text, &B[b].len, B[b].off.data(),
                                    st.data() + a);
      if (rc != NW_OK) {
        std::lock_guard<std::mutex> lk(mu);
        if (first_rc == NW_OK) {
          first_rc = rc;
          first_msg = nw_last_error();  // thread-local: carried to the caller's thread below
        }
        next.store(nb);
        return;
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_ctx; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  auto release = [&] {
    for (Batch& b : B) free(b.text);
  };
  if (first_rc != NW_OK) {
    release();
    nw::set_last_error(first_msg);
    return first_rc;
  }
  std::string text;
  for (int64_t b = 0; b < nb; ++b) {
    const int64_t a = b * per;
    if (row_off)
      for (size_t q = 0; q + 1 < B[b].off.size(); ++q) row_off[a + (int64_t)q] = (int64_t)text.size() + B[b].off[q];
    if (B[b].text) text.append(B[b].text, (size_t)B[b].len);
  }
  if (row_off) row_off[n] = (int64_t)text.size();
  release();
  if (status) memcpy(status, st.data(), (size_t)n * sizeof(int32_t));
  if (logfile) {  // in region order, one locked append per row (as the one-context form does)
    size_t p0 = 0;
    while (p0 < text.size()) {
      const size_t e = text.find('\n', p0);
      const std::string line = text.substr(p0, e - p0);
      const int a = nw_region_log_append(logfile, line.c_str());
      if (a != NW_OK) return a;
      p0 = e + 1;
    }
  }
  char* out = (char*)malloc(text.size() + 1);
  if (!out) {
    nw::set_last_error("host out of memory");
    return NW_ERR_NOMEM;
  }
  memcpy(out, text.c_str(), text.size() + 1);
  *rows = out;
  *rows_len = (int64_t)text.size();
  return NW_OK;
}

extern "C" int nw_regions_run_files(nw_ctx** ctxs, int32_t n_ctx, const char* const* paf_paths, const nw_log_meta* metas,
                                    int64_t n, const nw_bitlocus_params* bp, const nw_region_opts* ropts,
                                    int32_t regions_per_batch, const char* logfile, char** rows, int64_t* rows_len,
                                    int64_t* row_off, int32_t* status) {
  if (rows) *rows = nullptr;
  if (rows_len) *rows_len = 0;
  if (!ctxs || n_ctx < 1 || !rows || !rows_len || n < 0 || (n > 0 && (!paf_paths || !metas))) {
    nw::set_last_error("null argument");
    return NW_ERR_ARG;
  }
  // read.table(inpaf) of every region; the regions that could be read run as one batch
  struct Tab {
    std::vector<double> c[10];
  };
  std::vector<Tab> tabs((size_t)n);
  std::vector<int32_t> st((size_t)n, NW_OK);
  std::vector<int64_t> idx;
  std::vector<nw_paf_cols> cols;
  std::vector<nw_log_meta> ms;
  for (int64_t k = 0; k < n; ++k) {
    nw_paf_table* t = nullptr;
    const int rc = nw_paf_read_file(paf_paths[k], &t);
    if (rc != NW_OK) {
      st[k] = rc;
      continue;
    }
    const size_t m = (size_t)nw_paf_rows(t);
    Tab& T = tabs[k];
    for (auto& v : T.c) v.resize(m);
    nw_paf_columns(t, T.c[0].data(), T.c[1].data(), T.c[2].data(), T.c[3].data(), T.c[4].data(), T.c[5].data(), T.c[6].data(),
                   T.c[7].data(), T.c[8].data(), T.c[9].data());
    nw_paf_free(t);
    idx.push_back(k);
    cols.push_back(nw_paf_cols{T.c[0].data(), T.c[1].data(), T.c[2].data(), T.c[3].data(), T.c[4].data(), T.c[5].data(),
                               T.c[6].data(), T.c[7].data(), T.c[8].data(), T.c[9].data(), (int64_t)m});
    ms.push_back(metas[k]);
  }
  const int64_t m = (int64_t)idx.size();
  std::vector<int64_t> off((size_t)m + 1, 0);
  std::vector<int32_t> sub((size_t)std::max<int64_t>(m, 1), NW_OK);
  char* text = nullptr;
  int64_t len = 0;
  const int rc = nw_regions_run_ctxs(ctxs, n_ctx, cols.data(), ms.data(), m, bp, ropts, regions_per_batch, logfile, &text, &len,
                                     off.data(), sub.data());
  if (rc != NW_OK) return rc;
  for (int64_t q = 0; q < m; ++q) st[idx[q]] = sub[q];
  if (row_off) {  // the dropped regions own an empty range
    int64_t q = 0;
    for (int64_t k = 0; k < n; ++k) {
      row_off[k] = off[q];
      if (q < m && idx[q] == k) ++q;
    }
    row_off[n] = off[m];
  }
  if (status) memcpy(status, st.data(), (size_t)n * sizeof(int32_t));
  *rows = text;
  *rows_len = len;
  return NW_OK;
}
