// nw_paf.cu -- PAF preparation upstream of the gridding (include/nwpaf.h).
//
// compress_paf_fnct cuts the target axis into chunks, and inside every chunk tests ALL ordered pairs of alignments
// for "i ends where j starts" (merge_paf_entries_intraloop: outer() differences against a tolerance, same strand).
// k_paf_pairs does that test for the whole table in one launch: one thread per ordered pair, which walks the chunks
// both alignments touch (a bit mask per alignment) with each chunk's own tolerance and reports the first chunk in
// which the pair qualifies -- exactly the information needed to rebuild the reference's pair order
// (order(col), first appearance, row).  Selection (dplyr top_n per row, then per column) and merge_rows are short
// and strictly sequential: they run on the host in the reference's order.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "../../include/nwbitlocus.h"
#include "../../include/nwpaf.h"

struct nw_ctx;
namespace nw {
void set_last_error(const std::string& m);
int ctx_device(nw_ctx* c);
cudaStream_t ctx_stream(nw_ctx* c);
}  // namespace nw

namespace {

struct PErr {
  int code;
  std::string msg;
};
[[noreturn]] void pfail(int code, const std::string& m) { throw PErr{code, m}; }
#define PCK(call)                                                                                 \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) pfail(NW_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

struct PairRec {
  int32_t col, step, row, pad;
};

// one thread per ordered pair (i = row: the alignment that ends, j = col: the alignment that starts)
__global__ void k_paf_pairs(const double* __restrict__ ts, const double* __restrict__ te, const double* __restrict__ qs,
                            const double* __restrict__ qe, const double* __restrict__ strand, const double* __restrict__ alen,
                            const unsigned long long* __restrict__ mask, const double* __restrict__ tol, int n,
                            PairRec* __restrict__ out, unsigned long long cap, unsigned long long* __restrict__ count) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= (long long)n * n) return;
  const int j = (int)(w / n), i = (int)(w % n);  // column-major like which(, arr.ind = T): consecutive threads share j
  unsigned long long common = mask[i] & mask[j];
  if (!common || strand[i] != strand[j]) return;
  const double dt = fabs(te[i] - ts[j]), dq = fabs(qe[i] - qs[j]);
  while (common) {
    const int s = __ffsll((long long)common) - 1;
    common &= common - 1;
    const double t = tol[s];
    if (dt < t && dq < t && (!(t > 100) || (alen[i] > t && alen[j] > t))) {
      const unsigned long long k = atomicAdd(count, 1ull);
      if (k < cap) out[k] = PairRec{j, s, i, 0};
      return;  // later chunks only repeat the pair (unique() drops the repeats)
    }
  }
}

struct Table {
  std::vector<double> c[10];  // qlen qstart qend strand tlen tstart tend nmatch alen mapq
  size_t n() const { return c[0].size(); }
};
enum { QLEN = 0, QSTART, QEND, STRAND, TLEN, TSTART, TEND, NMATCH, ALEN, MAPQ };

Table take(const Table& t, const std::vector<size_t>& idx) {
  Table o;
  for (int k = 0; k < 10; ++k) {
    o.c[k].resize(idx.size());
    for (size_t q = 0; q < idx.size(); ++q) o.c[k][q] = t.c[k][idx[q]];
  }
  return o;
}
Table from_cols(const nw_paf_cols* in) {
  if (!in || in->n < 0) pfail(NW_ERR_ARG, "null argument");
  const double* p[10] = {in->qlen, in->qstart, in->qend, in->strand, in->tlen, in->tstart, in->tend, in->nmatch, in->alen, in->mapq};
  Table t;
  for (int k = 0; k < 10; ++k) {
    if (in->n > 0 && !p[k]) pfail(NW_ERR_ARG, "null column");
    t.c[k].assign(p[k], p[k] + in->n);
    for (double v : t.c[k])
      if (!std::isfinite(v)) pfail(NW_ERR_ARG, "non-finite value in the PAF table");
  }
  for (double v : t.c[STRAND])
    if (v != 1.0 && v != -1.0) pfail(NW_ERR_ARG, "strand must be +1 or -1");
  return t;
}
void swap_minus(Table& t) {  // transform(paf, qend = ifelse(strand == '-', qstart, qend), qstart = ifelse(..., qend, qstart))
  for (size_t k = 0; k < t.n(); ++k)
    if (t.c[STRAND][k] < 0) std::swap(t.c[QSTART][k], t.c[QEND][k]);
}
std::vector<size_t> stable_order(const std::vector<double>& key) {
  std::vector<size_t> o(key.size());
  for (size_t k = 0; k < o.size(); ++k) o[k] = k;
  std::stable_sort(o.begin(), o.end(), [&](size_t a, size_t b) { return key[a] < key[b]; });
  return o;
}

// merge_rows (R/compress_paf.R:13-138); pairs are (row, col), 0-based here
Table merge_rows(const Table& p, const std::vector<std::pair<int, int>>& pairs) {
  Table o = p;
  std::vector<double>&qs = o.c[QSTART], &qe = o.c[QEND], &ts = o.c[TSTART], &te = o.c[TEND], &nm = o.c[NMATCH], &al = o.c[ALEN];
  const std::vector<double>& st = o.c[STRAND];
  for (size_t k = pairs.size(); k-- > 0;) {
    int nl1 = pairs[k].first, nl2 = pairs[k].second;
    const int bu = nl1;
    if (st[nl1] > 0) {
      const double q1 = qs[nl1], q2 = qs[nl2], e1 = qe[nl1], e2 = qe[nl2], t1 = ts[nl1], t2 = ts[nl2], f1 = te[nl1], f2 = te[nl2];
      qe[bu] = std::max(e1, e2);
      qs[bu] = std::min(q1, q2);
      te[bu] = std::max(f1, f2);
      ts[bu] = std::min(t1, t2);
    } else {
      if (ts[nl1] > ts[nl2]) {
        nl1 = nl2;
        nl2 = bu;
      }
      const double q1 = qs[nl1], q2 = qs[nl2], e1 = qe[nl1], e2 = qe[nl2], t1 = ts[nl1], t2 = ts[nl2], f1 = te[nl1], f2 = te[nl2];
      qe[bu] = std::min(e1, e2);
      qs[bu] = std::max(q1, q2);
      te[bu] = std::max(f1, f2);
      ts[bu] = std::min(t1, t2);
    }
    const double nmv = nm[nl1] + nm[nl2], alv = al[nl1] + al[nl2];
    nm[bu] = nmv;
    al[bu] = alv;
  }
  std::vector<char> is_col(p.n(), 0);
  for (const auto& pr : pairs) is_col[pr.second] = 1;
  std::vector<size_t> keep;
  for (size_t k = 0; k < o.n(); ++k) {
    if (is_col[k]) continue;
    const double ratio = std::fabs(te[k] - ts[k]) / std::fabs(qe[k] - qs[k]);
    if (ratio >= 0.5 && ratio <= 2) keep.push_back(k);  // dplyr::between; NaN fails both comparisons
  }
  return take(o, keep);
}

struct CompressStats {
  int64_t pairs_found = 0;
  int launches = 0;
};

// compress_paf_fnct(inpaf_df = , save_outpaf = F, qname = NULL)
Table compress(nw_ctx* ctx, const Table& in, bool second_run, double chunklen, double compression, int n_quadrants,
               CompressStats& stats) {
  Table p = take(in, stable_order(in.c[QSTART]));
  const size_t n0 = p.n();
  if (n0 <= 1) return p;
  p = take(p, stable_order(p.c[TSTART]));
  const double range_t = *std::max_element(p.c[TEND].begin(), p.c[TEND].end());
  const double minlen_for_merge = n0 < 15000 ? 0.0 : chunklen * 0.9;
  int nq = n_quadrants > 0 ? n_quadrants : (range_t < 50000 ? 2 : 10);
  if (nq > 64) pfail(NW_ERR_ARG, "at most 64 quadrants per axis");
  Table shortt;
  {
    std::vector<size_t> a, b;
    for (size_t k = 0; k < p.n(); ++k) (p.c[ALEN][k] <= minlen_for_merge ? a : b).push_back(k);
    shortt = take(p, a);
    p = take(p, b);
  }
  const int n = (int)p.n();
  std::vector<double> tsteps(nq);
  {
    const double by = nq > 1 ? (range_t - 1) / (nq - 1) : 0.0;  // seq(1, range, length.out = nq): from + k * by
    for (int k = 0; k < nq; ++k) tsteps[k] = std::nearbyint(1 + k * by);
  }
  double tstepsize = nq > 1 ? tsteps[1] - tsteps[0] : NAN;
  if (nq == 1) tstepsize = 1e10;
  const int nsteps = nq - 1;
  auto append_short = [&](Table& t) {
    for (int k = 0; k < 10; ++k) t.c[k].insert(t.c[k].end(), shortt.c[k].begin(), shortt.c[k].end());
  };
  if (n < 2 || nsteps < 1) return p;  // no chunk can hold two alignments: rowpairs stays empty (short rows are dropped too)
  // chunk membership and per-chunk tolerance
  std::vector<unsigned long long> mask(n, 0ull);
  std::vector<double> tol(std::max(nsteps, 1), 0.0);
  bool any = false;
  for (int s = 0; s < nsteps; ++s) {
    const double lo = tsteps[s], hi = tsteps[s] + tstepsize;
    int cnt = 0;
    double minq = INFINITY;
    for (int k = 0; k < n; ++k) {
      const double a = p.c[TSTART][k], b = p.c[TEND][k];
      if ((a >= lo && a <= hi) || (b >= lo && b <= hi)) {
        mask[k] |= 1ull << s;
        ++cnt;
        minq = std::min(minq, p.c[QLEN][k]);
      }
    }
    if (cnt > 1) {
      tol[s] = second_run ? std::min(0.5 * compression, minq * 0.9) : 0.05 * chunklen;
      any = true;
    } else {
      for (int k = 0; k < n; ++k) mask[k] &= ~(1ull << s);  // nrow(inpaf_q) <= 1: the chunk is skipped
    }
  }
  if (!any) return p;
  // ---- GPU: all ordered pairs ----
  std::vector<PairRec> recs;
  {
    PCK(cudaSetDevice(nw::ctx_device(ctx)));
    cudaStream_t st = nw::ctx_stream(ctx);
    const size_t nb = (size_t)n * 8;
    unsigned char* d = nullptr;
    unsigned long long cap = std::max<unsigned long long>(4096, (unsigned long long)n * 8);
    for (;;) {
      const size_t total = 7 * nb + (size_t)nsteps * 8 + 16 + cap * sizeof(PairRec);
      PCK(cudaMalloc(&d, total));
      try {
        double* dts = (double*)d;
        double *dte = dts + n, *dqs = dte + n, *dqe = dqs + n, *dst = dqe + n, *dal = dst + n;
        unsigned long long* dmask = (unsigned long long*)(dal + n);
        double* dtol = (double*)(dmask + n);
        unsigned long long* dcount = (unsigned long long*)(dtol + nsteps);
        PairRec* dout = (PairRec*)(dcount + 2);
        PCK(cudaMemcpyAsync(dts, p.c[TSTART].data(), nb, cudaMemcpyHostToDevice, st));
        PCK(cudaMemcpyAsync(dte, p.c[TEND].data(), nb, cudaMemcpyHostToDevice, st));
        PCK(cudaMemcpyAsync(dqs, p.c[QSTART].data(), nb, cudaMemcpyHostToDevice, st));
        PCK(cudaMemcpyAsync(dqe, p.c[QEND].data(), nb, cudaMemcpyHostToDevice, st));
        PCK(cudaMemcpyAsync(dst, p.c[STRAND].data(), nb, cudaMemcpyHostToDevice, st));
        PCK(cudaMemcpyAsync(dal, p.c[ALEN].data(), nb, cudaMemcpyHostToDevice, st));
        PCK(cudaMemcpyAsync(dmask, mask.data(), nb, cudaMemcpyHostToDevice, st));
        PCK(cudaMemcpyAsync(dtol, tol.data(), (size_t)nsteps * 8, cudaMemcpyHostToDevice, st));
        PCK(cudaMemsetAsync(dcount, 0, 16, st));
        const long long work = (long long)n * n;
        k_paf_pairs<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(dts, dte, dqs, dqe, dst, dal, dmask, dtol, n, dout, cap, dcount);
        PCK(cudaGetLastError());
        ++stats.launches;
        unsigned long long found = 0;
        PCK(cudaMemcpyAsync(&found, dcount, 8, cudaMemcpyDeviceToHost, st));
        PCK(cudaStreamSynchronize(st));
        if (found <= cap) {
          recs.resize((size_t)found);
          if (found) PCK(cudaMemcpy(recs.data(), dout, (size_t)found * sizeof(PairRec), cudaMemcpyDeviceToHost));
          cudaFree(d);
          d = nullptr;
          break;
        }
        cap = found + found / 8;
      } catch (...) {
        cudaFree(d);
        throw;
      }
      cudaFree(d);
      d = nullptr;
    }
  }
  stats.pairs_found = (int64_t)recs.size();
  if (recs.empty()) return p;
  // rowpairs[order(col), ] (stable over chunk order, then column-major inside a chunk), unique(): the kernel reports a
  // pair once, with its first chunk, so the reference's order is (col, first chunk, row)
  std::sort(recs.begin(), recs.end(), [](const PairRec& a, const PairRec& b) {
    if (a.col != b.col) return a.col < b.col;
    if (a.step != b.step) return a.step < b.step;
    return a.row < b.row;
  });
  struct RP {
    int row, col;
    double ml;
  };
  std::vector<RP> rp;
  for (const PairRec& r : recs) {
    if (r.row == r.col) continue;
    const double ml = p.c[NMATCH][r.row] + p.c[NMATCH][r.col];
    if (ml > chunklen / 5) rp.push_back(RP{r.row, r.col, ml});
  }
  // group_by(row) %>% top_n(1, combined_matchlen) %>% group_by(col) %>% top_n(1, combined_matchlen): ties are kept
  {
    std::map<int, double> best;
    for (const RP& r : rp) {
      auto it = best.find(r.row);
      if (it == best.end() || r.ml > it->second) best[r.row] = r.ml;
    }
    std::vector<RP> k1;
    for (const RP& r : rp)
      if (r.ml == best[r.row]) k1.push_back(r);
    best.clear();
    for (const RP& r : k1) {
      auto it = best.find(r.col);
      if (it == best.end() || r.ml > it->second) best[r.col] = r.ml;
    }
    rp.clear();
    for (const RP& r : k1)
      if (r.ml == best[r.col]) rp.push_back(r);
  }
  if (!rp.empty()) {
    std::vector<std::pair<int, int>> pairs;
    for (const RP& r : rp) pairs.emplace_back(r.row, r.col);
    p = merge_rows(p, pairs);
  }
  append_short(p);
  return p;
}

Table enforce_slope_one(const Table& p) {  // R/tsv_to_bitlocus.R:258-275
  Table o = p;
  std::vector<size_t> keep;
  for (size_t k = 0; k < o.n(); ++k) {
    const double ql = std::fabs(o.c[QEND][k] - o.c[QSTART][k]), tl = std::fabs(o.c[TEND][k] - o.c[TSTART][k]);
    const double sq = -std::max(-ql, -tl);
    o.c[QEND][k] = o.c[QSTART][k] + sq;
    o.c[TEND][k] = o.c[TSTART][k] + sq;
    if (o.c[QSTART][k] != o.c[QEND][k] && o.c[TSTART][k] != o.c[TEND][k]) keep.push_back(k);
  }
  return take(o, keep);
}

// load_and_prep_paf_for_gridplot_transform from the parsed table on (R/tsv_to_bitlocus.R:38-82)
Table prepare(nw_ctx* ctx, Table p, double minlen, double compression, double chunklen, CompressStats& stats) {
  if (!(compression > 0)) pfail(NW_ERR_ARG, "compression must be positive");
  swap_minus(p);
  p = compress(ctx, p, true, chunklen, compression, 0, stats);
  {
    std::vector<size_t> keep;
    for (size_t k = 0; k < p.n(); ++k)
      if (p.c[ALEN][k] > minlen) keep.push_back(k);
    p = take(p, keep);
  }
  if (p.n() > 0) {
    for (int col : {QSTART, QEND, TSTART, TEND})
      for (double& v : p.c[col]) v = std::nearbyint(v / compression) * compression;  // round(x / compression, 0) * compression
    swap_minus(p);
    p = enforce_slope_one(p);
    swap_minus(p);
  }
  return p;
}

}  // namespace

struct nw_paf_table {
  Table t;
  CompressStats stats;
};

namespace {
int pguard(const std::function<void()>& fn) {
  try {
    fn();
    return NW_OK;
  } catch (const PErr& e) {
    nw::set_last_error(e.msg);
    cudaGetLastError();
    return e.code;
  } catch (const std::bad_alloc&) {
    nw::set_last_error("host out of memory");
    return NW_ERR_NOMEM;
  } catch (const std::exception& e) {
    nw::set_last_error(e.what());
    return NW_ERR_ARG;
  }
}
}  // namespace

extern "C" {

int nw_paf_compress(nw_ctx* ctx, const nw_paf_cols* in, int32_t second_run, double chunklen, double compression,
                    int32_t n_quadrants, nw_paf_table** out) {
  if (out) *out = nullptr;
  return pguard([&] {
    if (!ctx || !out) pfail(NW_ERR_ARG, "null argument");
    std::unique_ptr<nw_paf_table> T(new nw_paf_table);
    T->t = compress(ctx, from_cols(in), second_run != 0, chunklen, compression, n_quadrants, T->stats);
    *out = T.release();
  });
}

int nw_paf_prepare(nw_ctx* ctx, const nw_paf_cols* in, double minlen, double compression, double chunklen,
                   nw_paf_table** out) {
  if (out) *out = nullptr;
  return pguard([&] {
    if (!ctx || !out) pfail(NW_ERR_ARG, "null argument");
    std::unique_ptr<nw_paf_table> T(new nw_paf_table);
    T->t = prepare(ctx, from_cols(in), minlen, compression, chunklen, T->stats);
    *out = T.release();
  });
}

int64_t nw_paf_rows(const nw_paf_table* t) { return t ? (int64_t)t->t.n() : 0; }
int nw_paf_columns(const nw_paf_table* t, double* qlen, double* qstart, double* qend, double* strand, double* tlen,
                   double* tstart, double* tend, double* nmatch, double* alen, double* mapq) {
  if (!t) return NW_ERR_ARG;
  double* p[10] = {qlen, qstart, qend, strand, tlen, tstart, tend, nmatch, alen, mapq};
  for (int k = 0; k < 10; ++k)
    if (p[k] && t->t.n()) memcpy(p[k], t->t.c[k].data(), t->t.n() * 8);
  return NW_OK;
}
int nw_paf_stats(const nw_paf_table* t, int64_t* pairs_found, int32_t* gpu_launches) {
  if (!t) return NW_ERR_ARG;
  if (pairs_found) *pairs_found = t->stats.pairs_found;
  if (gpu_launches) *gpu_launches = t->stats.launches;
  return NW_OK;
}
void nw_paf_free(nw_paf_table* t) { delete t; }

}  // extern "C"

// =====================================================================================================
// wrapper_paf_to_bitlocus (R/tsv_to_bitlocus.R:151-241): prepare -> checks -> grid -> checks, coarser on rejection
// =====================================================================================================
namespace {

// update_params (R/tsv_to_bitlocus.R:111-129): both minlen and compression become the returned value
double update_params(double minlen) {
  const double next_power_of_10 = std::pow(10.0, std::floor(std::log10(minlen)) + 1.0);
  double new_val = minlen >= next_power_of_10 ? next_power_of_10 * 2 : minlen * 2;
  if (new_val >= next_power_of_10) new_val = next_power_of_10;
  return new_val;
}

// is_cluttered_paf (R/explore_mutation_space_v2_helpers.R:231-256); the fourth term of seems_simplistic repeats qstart
bool is_cluttered_paf(const Table& p, int limit = 5) {
  const size_t n = p.n();
  if (n == 0) return false;
  auto count_eq = [&](int col, double v) {
    int c = 0;
    for (double x : p.c[col]) c += x == v ? 1 : 0;
    return c;
  };
  auto n_unique = [&](int col) { return std::set<double>(p.c[col].begin(), p.c[col].end()).size(); };
  const double max_qend = *std::max_element(p.c[QEND].begin(), p.c[QEND].end());
  const double max_tend = *std::max_element(p.c[TEND].begin(), p.c[TEND].end());
  const int clutter[4] = {count_eq(QSTART, 0.0), count_eq(QEND, max_qend), count_eq(TSTART, 0.0), count_eq(TEND, max_tend)};
  const double half = (double)n / 2.0;
  const bool seems_simplistic = (double)n_unique(QSTART) < half || (double)n_unique(TSTART) < half || (double)n_unique(QEND) < half;
  bool any = false;
  for (int c : clutter) any = any || c >= limit;
  return any && seems_simplistic;
}

// nrow(find_sv_opportunities(mat)) (R/find_sv_opportunities.R:12-98): unique (p1, p2, orientation) over the rows with
// two or more non-zero cells; same-orientation pairs count twice (del + dup), adjacent inverted pairs are dropped
int64_t count_sv_opportunities(const std::vector<double>& mat, int ny, int nx) {
  std::set<std::pair<std::pair<int, int>, int>> seen;
  std::vector<int> nz;
  for (int y = 0; y < ny; ++y) {
    nz.clear();
    for (int x = 0; x < nx; ++x)
      if (mat[(size_t)y * nx + x] != 0) nz.push_back(x);
    for (size_t a = 0; a < nz.size(); ++a)
      for (size_t b = a + 1; b < nz.size(); ++b) {
        const int ori = ((mat[(size_t)y * nx + nz[a]] > 0) == (mat[(size_t)y * nx + nz[b]] > 0)) ? 1 : -1;
        seen.insert({{nz[a], nz[b]}, ori});
      }
  }
  int64_t n = 0;
  for (const auto& k : seen) {
    if (k.second == 1)
      n += 2;
    else if (k.first.second - k.first.first != 1)
      n += 1;
  }
  return n;
}

}  // namespace

extern "C" {

void nw_bitlocus_default_params(nw_bitlocus_params* p) {
  if (!p) return;
  p->minlen = 10000;  // "default" of a 0.5 - 5 Mb region (nw_bitlocus_auto_params)
  p->compression = 10000;
  p->chunklen = 10000;
  p->max_n_alns = 150;
  p->max_size_col_plus_rows = 250;
  p->max_pairs = 2000;
  p->max_passes = 64;
}

// determine_chunklen / determine_compression (R/wrapper_aln_and_analyse.R:269-307): the value of the first cutoff that
// is >= the interval length (which.min over dists[dists > 0]); at or beyond the last cutoff which.min of all-Inf is 1
static double first_cutoff_value(double interval_len, const double* cutoffs, const double* values, int n) {
  for (int k = 0; k < n; ++k) {
    const double d = cutoffs[k] - interval_len;
    if (d > 0) return values[k];
    if (d == 0) return k + 1 < n ? values[k] : values[0];  // the zero is filtered out: the index shifts by one
  }
  return values[0];
}
void nw_bitlocus_auto_params(double start_x, double end_x, nw_bitlocus_params* p) {
  if (!p) return;
  const double len = end_x - start_x;
  static const double chunk_cut[6] = {10000, 500000, 1000000, 2000000, 5000000, 1e10};
  static const double chunk_val[6] = {100, 1000, 10000, 10000, 10000, 50000};
  static const double comp_cut[4] = {50000, 500000, 5000000, 1e10};
  static const double comp_val[4] = {100, 1000, 10000, 20000};
  p->chunklen = first_cutoff_value(len, chunk_cut, chunk_val, 6);
  p->minlen = p->compression = first_cutoff_value(len, comp_cut, comp_val, 4);
}

int nw_bitlocus_build(nw_ctx* ctx, const nw_paf_cols* pafs, int64_t n_regions, const nw_bitlocus_params* params,
                      nw_grid** grids, nw_paf_table** tables, nw_bitlocus_info* infos, int32_t* rcs) {
  if (grids)
    for (int64_t k = 0; k < n_regions; ++k) grids[k] = nullptr;
  if (tables)
    for (int64_t k = 0; k < n_regions; ++k) tables[k] = nullptr;
  std::vector<nw_grid*> pass_grids;  // grids of the pass in flight (freed if the pass throws)
  const int rc = pguard([&] {
    if (!ctx || !params || !grids || n_regions < 0 || (n_regions > 0 && !pafs)) pfail(NW_ERR_ARG, "null argument");
    if (!(params->minlen > 0) || !(params->compression > 0) || !std::isfinite(params->minlen) ||
        !std::isfinite(params->compression))
      pfail(NW_ERR_ARG, "minlen and compression must be positive");
    const int max_passes = params->max_passes > 0 ? params->max_passes : 64;
    struct Region {
      double minlen, compression;
      nw_bitlocus_info info;
      std::unique_ptr<nw_paf_table> tab;
      int32_t rc = NW_OK;
    };
    std::vector<Region> R((size_t)n_regions);
    std::vector<int64_t> pending;
    for (int64_t k = 0; k < n_regions; ++k) {
      R[k].minlen = params->minlen;
      R[k].compression = params->compression;
      R[k].info = nw_bitlocus_info{params->minlen, params->compression, 0, 0, 0, -1};
      pending.push_back(k);
    }
    auto coarser = [&](Region& r, std::vector<int64_t>& next, int64_t k) {
      r.minlen = r.compression = update_params(r.minlen);
      r.tab.reset();
      if (r.info.passes >= max_passes) {
        r.rc = NW_ERR_RANGE;
        nw::set_last_error("wrapper_paf_to_bitlocus: no acceptable grid within max_passes");
      } else {
        next.push_back(k);
      }
    };
    while (!pending.empty()) {
      std::vector<int64_t> next, togrid;
      for (int64_t k : pending) {
        Region& r = R[k];
        r.info.passes += 1;
        r.info.minlen = r.minlen;
        r.info.compression = r.compression;
        r.tab.reset(new nw_paf_table);
        try {
          r.tab->t = prepare(ctx, from_cols(&pafs[k]), r.minlen, r.compression, params->chunklen, r.tab->stats);
        } catch (const PErr& e) {
          if (e.code == NW_ERR_CUDA) throw;
          r.rc = e.code;
          nw::set_last_error(e.msg);
          r.tab.reset();
          continue;
        }
        const size_t n = r.tab->t.n();
        r.info.n_alns = (int32_t)n;
        if ((int64_t)n > (int64_t)params->max_n_alns) {  // :171-176
          coarser(r, next, k);
          continue;
        }
        if (n == 0) {  // make_xy_grid_fast stops on an empty table
          r.rc = NW_ERR_ARG;
          nw::set_last_error("no alignment longer than minlen");
          r.tab.reset();
          continue;
        }
        if (is_cluttered_paf(r.tab->t)) {  // :179-182: return() -- no grid
          r.info.cluttered = 1;
          r.tab.reset();
          continue;
        }
        togrid.push_back(k);
      }
      if (!togrid.empty()) {
        std::vector<nw_paf> gp(togrid.size());
        for (size_t q = 0; q < togrid.size(); ++q) {
          const Table& t = R[togrid[q]].tab->t;
          gp[q] = nw_paf{t.c[TSTART].data(), t.c[TEND].data(), t.c[QSTART].data(), t.c[QEND].data(), (int64_t)t.n()};
        }
        pass_grids.assign(togrid.size(), nullptr);
        std::vector<int32_t> grc(togrid.size(), NW_OK);
        const int brc = nw_grid_build(ctx, gp.data(), (int64_t)gp.size(), pass_grids.data(), grc.data());
        if (brc != NW_OK) pfail(brc, std::string("nw_grid_build: ") + nw_last_error());
        for (size_t q = 0; q < togrid.size(); ++q) {
          const int64_t k = togrid[q];
          Region& r = R[k];
          nw_grid*& g = pass_grids[q];
          if (grc[q] == NW_ERR_RANGE && r.info.n_alns <= NW_GRID_MAX_ALNS &&
              params->max_size_col_plus_rows <= NW_GRID_MAX_LINES) {
            coarser(r, next, k);  // more than NW_GRID_MAX_LINES gridlines on an axis: certainly above the limit of :192
            continue;
          }
          if (grc[q] != NW_OK) {
            r.rc = grc[q];
            r.tab.reset();
            continue;
          }
          int32_t nx = 0, ny = 0, cv = 0, gen = 0;
          nw_grid_dims(g, &nx, &ny, &cv, &gen);
          if ((int64_t)nx + 1 + ny + 1 > (int64_t)params->max_size_col_plus_rows) {  // :192-197
            nw_grid_free(g);
            g = nullptr;
            coarser(r, next, k);
            continue;
          }
          std::vector<double> mat((size_t)nx * ny), lx((size_t)nx), ly((size_t)ny);
          nw_grid_matrix(g, mat.data(), lx.data(), ly.data());
          r.info.n_pairs = (int32_t)std::min<int64_t>(count_sv_opportunities(mat, ny, nx), INT32_MAX);
          if (r.info.n_pairs > params->max_pairs) {  // :215-222
            nw_grid_free(g);
            g = nullptr;
            coarser(r, next, k);
            continue;
          }
          grids[k] = g;
          g = nullptr;
        }
        pass_grids.clear();
      }
      pending.swap(next);
    }
    for (int64_t k = 0; k < n_regions; ++k) {
      if (rcs) rcs[k] = R[k].rc;
      if (infos) infos[k] = R[k].info;
      if (tables && grids[k]) tables[k] = R[k].tab.release();
    }
  });
  if (rc != NW_OK) {
    for (nw_grid* g : pass_grids)
      if (g) nw_grid_free(g);
    if (grids)
      for (int64_t k = 0; k < n_regions; ++k) {
        if (grids[k]) nw_grid_free(grids[k]);
        grids[k] = nullptr;
      }
    if (tables)
      for (int64_t k = 0; k < n_regions; ++k) {
        delete tables[k];
        tables[k] = nullptr;
      }
  }
  return rc;
}

}  // extern "C"
