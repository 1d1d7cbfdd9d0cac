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
#include <atomic>
#include <cctype>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <unordered_set>
#include <mutex>
#include <thread>
#include <cmath>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "nw_host.h"
#include "../../include/nwbitlocus.h"
#include "../../include/nwpaf.h"

struct nw_ctx;
namespace nw {
void set_last_error(const std::string& m);
int ctx_device(nw_ctx* c);
cudaStream_t ctx_stream(nw_ctx* c);
cudaError_t ctx_region_staging(nw_ctx* c, size_t bytes, void** host_pinned, void** dev);
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

// tables are independent on the host side: the pre / post phases of a batch run on a few threads.  fn must not throw
// anything but PErr; the first error is rethrown on the calling thread.
void paf_parallel_for(int64_t n, const std::function<void(int64_t)>& fn) {
  const unsigned hw = (unsigned)nw::host_cores();
  const int nt = (int)std::min<int64_t>(std::max<int64_t>(1, n / 2), std::min<unsigned>(hw, 8u));
  if (nt <= 1) {
    for (int64_t k = 0; k < n; ++k) fn(k);
    return;
  }
  std::atomic<int64_t> next{0};
  std::mutex mu;
  bool failed = false;
  PErr err{NW_OK, ""};
  auto work = [&] {
    for (;;) {
      const int64_t k = next.fetch_add(1);
      if (k >= n) return;
      try {
        fn(k);
      } catch (const PErr& e) {
        std::lock_guard<std::mutex> lk(mu);
        if (!failed) err = e;
        failed = true;
        next.store(n);
      } catch (const std::exception& e) {
        std::lock_guard<std::mutex> lk(mu);
        if (!failed) err = PErr{NW_ERR_NOMEM, e.what()};
        failed = true;
        next.store(n);
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  if (failed) throw err;
}
double paf_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct PairRec {
  int32_t col, step, row, tab;
};

// one thread per ordered pair (i = row: the alignment that ends, j = col: the alignment that starts) of one table of a
// batch; the column arrays of all tables are concatenated, a CTA works on one table (tables start at CTA boundaries)
struct PafTab {
  uint32_t off;      // first element of the table in the concatenated columns
  uint32_t n;        // alignments
  uint32_t tol_off;  // first tolerance (one per chunk)
  uint32_t blk0;     // first CTA of the table
  uint32_t out_off;  // first record of the table's output segment
  uint32_t out_cap;  // records the segment holds
  uint32_t shift;    // n < 256: a CTA covers 256 >> shift columns x (1 << shift) rows, 1 << shift >= n
  uint32_t strips;   // n >= 256: a CTA covers 256 rows of one column, `strips` CTAs per column
};
// CTAs a table needs (no per-thread division: the pair of a thread follows from its CTA's column / row origin)
inline void paf_tab_layout(uint32_t n, uint32_t& shift, uint32_t& strips, size_t& blocks) {
  if (n >= 256) {
    shift = 0;
    strips = (n + 255) / 256;
    blocks = (size_t)n * strips;
  } else {
    shift = 1;
    while ((1u << shift) < n) ++shift;
    strips = 0;
    const uint32_t cols = 256u >> shift;
    blocks = (n + cols - 1) / cols;
  }
}
// A CTA works through a contiguous range of 256-pair tiles: it finds the table of its first tile by a binary search over
// the tables' first tiles ONCE and walks forward from there (a search per tile cost more than the tile's pairs).
__global__ void __launch_bounds__(256) k_paf_pairs(const PafTab* __restrict__ tabs, int n_tabs, uint32_t n_tiles,
                                                   const double* __restrict__ ts, const double* __restrict__ te,
                                                   const double* __restrict__ qs, const double* __restrict__ qe,
                                                   const double* __restrict__ strand, const double* __restrict__ alen,
                                                   const unsigned long long* __restrict__ mask, const double* __restrict__ tol,
                                                   PairRec* __restrict__ out, unsigned int* __restrict__ count /* [n_tabs] */) {
  const uint32_t per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
  uint32_t b = blockIdx.x * per_cta;
  const uint32_t bend = min(b + per_cta, n_tiles);
  if (b >= bend) return;
  int tab;
  {  // last table with blk0 <= b (uniform over the CTA: broadcast loads)
    int lo = 0, hi = n_tabs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tabs[mid].blk0 <= b) lo = mid; else hi = mid - 1;
    }
    tab = lo;
  }
  PafTab T = tabs[tab];
  uint32_t next0 = tab + 1 < n_tabs ? tabs[tab + 1].blk0 : 0xffffffffu;
  for (; b < bend; ++b) {
    while (b >= next0) {
      ++tab;
      T = tabs[tab];
      next0 = tab + 1 < n_tabs ? tabs[tab + 1].blk0 : 0xffffffffu;
    }
    const uint32_t n = T.n, lb = b - T.blk0;
    // i = row: the alignment that ends, j = col: the alignment that starts; consecutive threads share j
    uint32_t i, j;
    if (T.strips) {
      j = lb / T.strips;
      i = (lb - j * T.strips) * 256u + threadIdx.x;
    } else {
      j = lb * (256u >> T.shift) + (threadIdx.x >> T.shift);
      i = threadIdx.x & ((1u << T.shift) - 1u);
    }
    if (i >= n || j >= n) continue;
    const size_t gi = (size_t)T.off + i, gj = (size_t)T.off + j;
    unsigned long long common = mask[gi] & mask[gj];
    if (!common || strand[gi] != strand[gj]) continue;
    const double dt = fabs(te[gi] - ts[gj]), dq = fabs(qe[gi] - qs[gj]);
    while (common) {
      const int s = __ffsll((long long)common) - 1;
      common &= common - 1;
      const double t = tol[T.tol_off + s];
      if (dt < t && dq < t && (!(t > 100) || (alen[gi] > t && alen[gj] > t))) {
        const unsigned int k = atomicAdd(&count[tab], 1u);  // the table's own cursor and output segment
        if (k < T.out_cap) out[(size_t)T.out_off + k] = PairRec{(int32_t)j, s, (int32_t)i, tab};
        break;  // later chunks only repeat the pair (unique() drops the repeats)
      }
    }
  }
}

struct Table {
  std::vector<double> c[11];  // qlen qstart qend strand tlen tstart tend nmatch alen mapq + the row of the input table
  size_t n() const { return c[0].size(); }
};
enum { QLEN = 0, QSTART, QEND, STRAND, TLEN, TSTART, TEND, NMATCH, ALEN, MAPQ, ROWID, NCOL };

Table take(const Table& t, const std::vector<size_t>& idx) {
  Table o;
  for (int k = 0; k < NCOL; ++k) {
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
  t.c[ROWID].resize((size_t)in->n);
  for (size_t k = 0; k < t.c[ROWID].size(); ++k) t.c[ROWID][k] = (double)k;
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

// as.character() / paste0() / write.table() of an R double: the fewest significant digits (at most 15) that reproduce
// the value at 15 digits, fixed notation unless scientific is narrower (R: formatReal + scientific(), scipen = 0)
std::string r_format_double(double v) {
  if (v == 0) return "0";
  if (!std::isfinite(v)) return std::isnan(v) ? "NA" : (v > 0 ? "Inf" : "-Inf");
  char ref[64], buf[64];
  snprintf(ref, sizeof ref, "%.14e", v);
  const double target = strtod(ref, nullptr);
  int nsig = 15;
  for (int d = 1; d <= 15; ++d) {
    snprintf(buf, sizeof buf, "%.*e", d - 1, v);
    if (strtod(buf, nullptr) == target) {
      nsig = d;
      break;
    }
  }
  snprintf(buf, sizeof buf, "%.*e", nsig - 1, v);
  const char* e = strchr(buf, 'e');
  const int kp = atoi(e + 1);
  const int neg = v < 0 ? 1 : 0;
  const int rgt = std::max(0, nsig - kp - 1);
  const int left = kp >= 0 ? kp + 1 : 1;
  const int wF = neg + left + (rgt > 0 ? rgt + 1 : 0);
  const int wE = neg + (nsig > 1 ? nsig + 1 : 1) + (std::abs(kp) >= 100 ? 5 : 4);
  if (wF <= wE) {
    snprintf(buf, sizeof buf, "%.*f", rgt, v);
    return buf;
  }
  std::string mant(buf, e - buf);
  char ex[16];
  snprintf(ex, sizeof ex, "e%c%02d", kp < 0 ? '-' : '+', std::abs(kp));
  return mant + ex;
}
// a value of a column read.table typed integer prints as an integer, of a double column like a double
std::string r_format(double v, bool integer_typed) {
  if (integer_typed) {
    char buf[32];
    snprintf(buf, sizeof buf, "%lld", (long long)v);
    return buf;
  }
  return r_format_double(v);
}

// the names merge_rows rewrites (R/compress_paf.R:49-59, :86-97): aligned with the rows of the table
struct MergeNames {
  std::vector<std::string> qname;
  bool reduced = false;     // reduced_qnames
  bool q_integer = true;    // the qstart / qend columns are integer typed (formatting of the rebuilt name)
};

// merge_rows (R/compress_paf.R:13-138); pairs are (row, col), 0-based here
Table merge_rows(const Table& p, const std::vector<std::pair<int, int>>& pairs, MergeNames* names = nullptr) {
  Table o = p;
  std::vector<double>&qs = o.c[QSTART], &qe = o.c[QEND], &ts = o.c[TSTART], &te = o.c[TEND], &nm = o.c[NMATCH], &al = o.c[ALEN];
  const std::vector<double>& st = o.c[STRAND];
  for (size_t k = pairs.size(); k-- > 0;) {
    int nl1 = pairs[k].first, nl2 = pairs[k].second;
    const int bu = nl1;
    if (st[nl1] > 0) {
      const double q1 = qs[nl1], q2 = qs[nl2], e1 = qe[nl1], e2 = qe[nl2], t1 = ts[nl1], t2 = ts[nl2], f1 = te[nl1], f2 = te[nl2];
      if (names && !names->reduced) {  // paste0(sub("_[^_]*$", "", qname1), "_", qstart1, "-", qend2 - 1)
        std::string base = names->qname[nl1];
        const size_t us = base.rfind('_');
        if (us != std::string::npos) base.erase(us);
        names->qname[bu] = base + "_" + r_format(q1, names->q_integer) + "-" + r_format_double(e2 - 1);
      }
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
      if (names) {  // :97 overrides the rebuilt name: qnames[nl1_bu] <- qname1
        const std::string q = names->qname[nl1];
        names->qname[bu] = q;
      }
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
  if (names) {
    std::vector<std::string> kept;
    for (size_t k : keep) kept.push_back(names->qname[k]);
    names->qname.swap(kept);
  }
  return take(o, keep);
}

struct CompressStats {
  int64_t pairs_found = 0;
  int launches = 0;
};

// compress_paf_fnct(inpaf_df = , save_outpaf = F, qname = NULL) in three phases, so that the all-pairs tests of many
// tables share one upload, one launch and one download: host (sorting, chunk masks, tolerances) -> GPU -> host
// (selection, merge_rows)
struct CompressJob {
  bool second_run = true;
  double chunklen = 10000, compression = 1000;
  int n_quadrants = 0;
  Table p, shortt;
  int n = 0, nsteps = 0;
  std::vector<unsigned long long> mask;
  std::vector<double> tol;
  bool need_gpu = false;
  std::vector<PairRec> recs;
  CompressStats stats;
  // file mode: the qname strings of the input rows (indexed by ROWID); out: names of the result rows, merged = merge_rows ran
  const std::vector<std::string>* in_qname = nullptr;
  bool q_integer = true;
  bool reduced_qnames = false;
  std::vector<std::string> out_qname;
  bool merged = false;
};

void compress_pre(CompressJob& J, const Table& in) {
  J.need_gpu = false;
  J.recs.clear();
  J.shortt = Table();
  J.p = take(in, stable_order(in.c[QSTART]));
  Table& p = J.p;
  const size_t n0 = p.n();
  if (n0 <= 1) return;
  p = take(p, stable_order(p.c[TSTART]));
  const double range_t = *std::max_element(p.c[TEND].begin(), p.c[TEND].end());
  const double minlen_for_merge = n0 < 15000 ? 0.0 : J.chunklen * 0.9;
  const int nq = J.n_quadrants > 0 ? J.n_quadrants : (range_t < 50000 ? 2 : 10);
  if (nq > 64) pfail(NW_ERR_ARG, "at most 64 quadrants per axis");
  {
    std::vector<size_t> a, b;
    for (size_t k = 0; k < p.n(); ++k) (p.c[ALEN][k] <= minlen_for_merge ? a : b).push_back(k);
    J.shortt = take(p, a);
    p = take(p, b);
  }
  const int n = J.n = (int)p.n();
  std::vector<double> tsteps(nq);
  {
    const double by = nq > 1 ? (range_t - 1) / (nq - 1) : 0.0;  // seq(1, range, length.out = nq): from + k * by
    for (int k = 0; k < nq; ++k) tsteps[k] = std::nearbyint(1 + k * by);
  }
  double tstepsize = nq > 1 ? tsteps[1] - tsteps[0] : NAN;
  if (nq == 1) tstepsize = 1e10;
  const int nsteps = J.nsteps = nq - 1;
  if (n < 2 || nsteps < 1) {  // no chunk can hold two alignments: rowpairs stays empty (short rows are dropped too)
    J.shortt = Table();
    return;
  }
  // chunk membership and per-chunk tolerance
  J.mask.assign(n, 0ull);
  J.tol.assign(std::max(nsteps, 1), 0.0);
  bool any = false;
  for (int s = 0; s < nsteps; ++s) {
    const double lo = tsteps[s], hi = tsteps[s] + tstepsize;
    int cnt = 0;
    double minq = INFINITY;
    for (int k = 0; k < n; ++k) {
      const double a = p.c[TSTART][k], b = p.c[TEND][k];
      if ((a >= lo && a <= hi) || (b >= lo && b <= hi)) {
        J.mask[k] |= 1ull << s;
        ++cnt;
        minq = std::min(minq, p.c[QLEN][k]);
      }
    }
    if (cnt > 1) {
      J.tol[s] = J.second_run ? std::min(0.5 * J.compression, minq * 0.9) : 0.05 * J.chunklen;
      any = true;
    } else {
      for (int k = 0; k < n; ++k) J.mask[k] &= ~(1ull << s);  // nrow(inpaf_q) <= 1: the chunk is skipped
    }
  }
  if (!any) {
    J.shortt = Table();
    return;
  }
  J.need_gpu = true;
}

// ---- GPU: all ordered pairs of every job that needs them, one launch per sub-batch ----
constexpr size_t PAF_BATCH_ALNS = 1u << 20;     // alignments per launch (56 B each in the staging buffer)
constexpr size_t PAF_BATCH_BLOCKS = 1u << 30;   // CTAs per launch

void compress_gpu_batch(nw_ctx* ctx, CompressJob** jobs, size_t nj, int& launches) {
  PCK(cudaSetDevice(nw::ctx_device(ctx)));
  cudaStream_t st = nw::ctx_stream(ctx);
  size_t N = 0, NT = 0, blocks = 0;
  std::vector<PafTab> tabs(nj);
  std::vector<size_t> cap(nj);
  for (size_t t = 0; t < nj; ++t) {
    const CompressJob& J = *jobs[t];
    uint32_t shift = 0, strips = 0;
    size_t nb = 0;
    paf_tab_layout((uint32_t)J.n, shift, strips, nb);
    tabs[t] = PafTab{(uint32_t)N, (uint32_t)J.n, (uint32_t)NT, (uint32_t)blocks, 0u, 0u, shift, strips};
    N += (size_t)J.n;
    NT += (size_t)J.nsteps;
    blocks += nb;
    cap[t] = 2 * (size_t)J.n + 64;  // chunks of one alignment pair up with their neighbours: ~n pairs
  }
  auto al16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
  const size_t o_tabs = 0, o_cols = al16(nj * sizeof(PafTab)), o_tol = o_cols + 7 * N * 8, o_cnt = al16(o_tol + NT * 8),
               o_out = al16(o_cnt + nj * 4);
  for (;;) {
    size_t ncap = 0;
    for (size_t t = 0; t < nj; ++t) {
      tabs[t].out_off = (uint32_t)ncap;
      tabs[t].out_cap = (uint32_t)cap[t];
      ncap += cap[t];
    }
    if (ncap > 0xffffffffull) pfail(NW_ERR_RANGE, "PAF batch too large");
    const size_t total = o_out + ncap * sizeof(PairRec);
    void *hp = nullptr, *dp = nullptr;
    PCK(nw::ctx_region_staging(ctx, total, &hp, &dp));
    unsigned char* h = (unsigned char*)hp;
    unsigned char* d = (unsigned char*)dp;
    memcpy(h + o_tabs, tabs.data(), nj * sizeof(PafTab));
    static const int colsel[6] = {TSTART, TEND, QSTART, QEND, STRAND, ALEN};
    paf_parallel_for((int64_t)nj, [&](int64_t t) {
      const CompressJob& J = *jobs[t];
      for (int c = 0; c < 6; ++c) memcpy(h + o_cols + ((size_t)c * N + tabs[t].off) * 8, J.p.c[colsel[c]].data(), (size_t)J.n * 8);
      memcpy(h + o_cols + ((size_t)6 * N + tabs[t].off) * 8, J.mask.data(), (size_t)J.n * 8);
      memcpy(h + o_tol + (size_t)tabs[t].tol_off * 8, J.tol.data(), (size_t)J.nsteps * 8);
    });
    memset(h + o_cnt, 0, o_out - o_cnt);
    PCK(cudaMemcpyAsync(d, h, o_out, cudaMemcpyHostToDevice, st));
    const double* dc = (const double*)(d + o_cols);
    const unsigned grid = (unsigned)std::min<size_t>(blocks, (size_t)148 * 32);  // 148 SMs x 8 resident CTAs x 4 waves
    k_paf_pairs<<<grid, 256, 0, st>>>((const PafTab*)(d + o_tabs), (int)nj, (uint32_t)blocks, dc, dc + N, dc + 2 * N, dc + 3 * N,
                                                  dc + 4 * N, dc + 5 * N, (const unsigned long long*)(dc + 6 * N),
                                                  (const double*)(d + o_tol), (PairRec*)(d + o_out),
                                                  (unsigned int*)(d + o_cnt));
    PCK(cudaGetLastError());
    ++launches;
    PCK(cudaMemcpyAsync(h + o_cnt, d + o_cnt, nj * 4, cudaMemcpyDeviceToHost, st));
    PCK(cudaStreamSynchronize(st));
    const unsigned int* found = (const unsigned int*)(h + o_cnt);
    bool overflow = false;
    size_t last = 0;  // records up to the end of the last non-empty segment are fetched
    for (size_t t = 0; t < nj; ++t) {
      if (found[t] > cap[t]) {
        cap[t] = (size_t)found[t] + found[t] / 8;
        overflow = true;
      }
      if (found[t]) last = (size_t)tabs[t].out_off + std::min<size_t>(found[t], tabs[t].out_cap);
    }
    if (overflow) continue;  // rare: a segment was too small -- again with the exact sizes
    if (last) {
      PCK(cudaMemcpyAsync(h + o_out, d + o_out, last * sizeof(PairRec), cudaMemcpyDeviceToHost, st));
      PCK(cudaStreamSynchronize(st));
      const PairRec* r = (const PairRec*)(h + o_out);
      paf_parallel_for((int64_t)nj, [&](int64_t t) {
        jobs[t]->recs.assign(r + tabs[t].out_off, r + tabs[t].out_off + found[t]);
      });
    }
    return;
  }
}

void compress_gpu(nw_ctx* ctx, const std::vector<CompressJob*>& all) {
  std::vector<CompressJob*> need;
  for (CompressJob* J : all)
    if (J->need_gpu) need.push_back(J);
  size_t a = 0;
  while (a < need.size()) {
    size_t b = a, N = 0, blocks = 0;
    while (b < need.size()) {
      const size_t n = (size_t)need[b]->n;
      uint32_t sh_ = 0, st_ = 0;
      size_t bl = 0;
      paf_tab_layout((uint32_t)n, sh_, st_, bl);
      if (b > a && (N + n > PAF_BATCH_ALNS || blocks + bl > PAF_BATCH_BLOCKS)) break;
      N += n;
      blocks += bl;
      ++b;
    }
    if (N > 0xffffffffull || blocks > 0x7fffffffull) pfail(NW_ERR_RANGE, "PAF table too large");
    int launches = 0;
    compress_gpu_batch(ctx, need.data() + a, b - a, launches);
    for (size_t t = a; t < b; ++t) need[t]->stats.launches += launches;  // launches the table took part in
    a = b;
  }
}

Table compress_post(CompressJob& J) {
  if (!J.need_gpu) return std::move(J.p);
  Table& p = J.p;
  const double chunklen = J.chunklen;
  auto append_short = [&](Table& t) {
    for (int k = 0; k < NCOL; ++k) t.c[k].insert(t.c[k].end(), J.shortt.c[k].begin(), J.shortt.c[k].end());
  };
  std::vector<PairRec>& recs = J.recs;
  J.stats.pairs_found = (int64_t)recs.size();
  if (recs.empty()) return std::move(p);
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
    const double none = -INFINITY;
    std::vector<double> best(p.n(), none);
    for (const RP& r : rp) best[r.row] = std::max(best[r.row], r.ml);
    std::vector<RP> k1;
    for (const RP& r : rp)
      if (r.ml == best[r.row]) k1.push_back(r);
    std::fill(best.begin(), best.end(), none);
    for (const RP& r : k1) best[r.col] = std::max(best[r.col], r.ml);
    rp.clear();
    for (const RP& r : k1)
      if (r.ml == best[r.col]) rp.push_back(r);
  }
  if (!rp.empty()) {
    std::vector<std::pair<int, int>> pairs;
    for (const RP& r : rp) pairs.emplace_back(r.row, r.col);
    if (J.in_qname) {
      MergeNames nm;
      nm.q_integer = J.q_integer;
      nm.reduced = J.reduced_qnames;
      for (size_t k = 0; k < p.n(); ++k) nm.qname.push_back((*J.in_qname)[(size_t)p.c[ROWID][k]]);
      p = merge_rows(p, pairs, &nm);
      J.out_qname.swap(nm.qname);
      for (size_t k = 0; k < J.shortt.n(); ++k) J.out_qname.push_back((*J.in_qname)[(size_t)J.shortt.c[ROWID][k]]);
    } else {
      p = merge_rows(p, pairs);
    }
    J.merged = true;
  }
  append_short(p);
  return std::move(p);
}

Table compress(nw_ctx* ctx, const Table& in, bool second_run, double chunklen, double compression, int n_quadrants,
               CompressStats& stats) {
  CompressJob J;
  J.second_run = second_run;
  J.chunklen = chunklen;
  J.compression = compression;
  J.n_quadrants = n_quadrants;
  compress_pre(J, in);
  compress_gpu(ctx, {&J});
  Table out = compress_post(J);
  stats = J.stats;
  return out;
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

// load_and_prep_paf_for_gridplot_transform from the parsed table on (R/tsv_to_bitlocus.R:38-82), in the phases of
// compress: prepare_pre (host) -> compress_gpu (one launch for all jobs of a batch) -> prepare_post (host)
struct PrepJob {
  CompressJob cj;
  double minlen = 0, compression = 0;
};
void prepare_pre(PrepJob& J, Table p, double minlen, double compression, double chunklen) {
  if (!(compression > 0)) pfail(NW_ERR_ARG, "compression must be positive");
  J.minlen = minlen;
  J.compression = compression;
  J.cj = CompressJob();
  J.cj.second_run = true;
  J.cj.chunklen = chunklen;
  J.cj.compression = compression;
  swap_minus(p);
  compress_pre(J.cj, p);
}
Table prepare_post(PrepJob& J, CompressStats& stats) {
  Table p = compress_post(J.cj);
  stats = J.cj.stats;
  const double minlen = J.minlen, compression = J.compression;
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
Table prepare(nw_ctx* ctx, Table p, double minlen, double compression, double chunklen, CompressStats& stats) {
  PrepJob J;
  prepare_pre(J, std::move(p), minlen, compression, chunklen);
  compress_gpu(ctx, {&J.cj});
  return prepare_post(J, stats);
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

int nw_paf_prepare_many(nw_ctx* ctx, const nw_paf_cols* ins, int64_t n_tables, double minlen, double compression,
                        double chunklen, nw_paf_table** outs, int32_t* rcs) {
  if (outs)
    for (int64_t k = 0; k < n_tables; ++k) outs[k] = nullptr;
  const int rc = pguard([&] {
    if (!ctx || !outs || n_tables < 0 || (n_tables > 0 && !ins)) pfail(NW_ERR_ARG, "null argument");
    static const bool dbg = getenv("NW_DEBUG_TIMING") != nullptr;
    const double t0 = paf_now();
    std::vector<PrepJob> jobs((size_t)n_tables);
    std::vector<int32_t> st((size_t)n_tables, NW_OK);
    std::vector<std::string> msg((size_t)n_tables);
    paf_parallel_for(n_tables, [&](int64_t k) {
      try {
        prepare_pre(jobs[k], from_cols(&ins[k]), minlen, compression, chunklen);
      } catch (const PErr& e) {
        st[k] = e.code;
        msg[k] = e.msg;
      }
    });
    std::vector<CompressJob*> run;
    for (int64_t k = 0; k < n_tables; ++k) {
      if (st[k] == NW_OK)
        run.push_back(&jobs[k].cj);
      else
        nw::set_last_error(msg[k]);
    }
    const double t1 = paf_now();
    compress_gpu(ctx, run);
    const double t2 = paf_now();
    std::vector<std::unique_ptr<nw_paf_table>> res((size_t)n_tables);
    paf_parallel_for(n_tables, [&](int64_t k) {
      if (st[k] != NW_OK) return;
      res[k].reset(new nw_paf_table);
      res[k]->t = prepare_post(jobs[k], res[k]->stats);
    });
    for (int64_t k = 0; k < n_tables; ++k) outs[k] = res[k].release();
    if (dbg)
      fprintf(stderr, "[nw_paf_prepare_many] %lld tables: pre %.2f ms, gpu %.2f ms, post %.2f ms\n", (long long)n_tables,
              1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (paf_now() - t2));
    if (rcs)
      for (int64_t k = 0; k < n_tables; ++k) rcs[k] = st[k];
  });
  if (rc != NW_OK && outs)
    for (int64_t k = 0; k < n_tables; ++k) {
      delete outs[k];
      outs[k] = nullptr;
    }
  return rc;
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
// PAF files: read.table(inpaf) and compress_paf_fnct(inpaf_link, outpaf_link) (save_outpaf = T)
// =====================================================================================================
namespace {

struct PafFile {
  Table t;                                  // numeric columns (+ ROWID = row of this struct's string vectors)
  std::vector<std::string> qname, tname, extra;  // extra: columns 13.. joined by tabs (carried through, never parsed)
  bool integer_typed[10];                   // read.table types a column integer if every value is an integer
};

// read.table(path, sep = "\t") of a PAF: 12 columns (qname qlen qstart qend strand tname tlen tstart tend nmatch alen
// mapq) + optional tags; dedupe = unique(inpaf) (first occurrences, compared as text)
PafFile read_paf_file(const char* path, bool dedupe) {
  if (!path) pfail(NW_ERR_ARG, "null path");
  std::ifstream f(path);
  if (!f) pfail(NW_ERR_IO, std::string("cannot open ") + path);
  PafFile P;
  for (bool& b : P.integer_typed) b = true;
  std::unordered_set<std::string> seen;
  std::string line;
  size_t lineno = 0;
  static const int numcol[12] = {-1, QLEN, QSTART, QEND, -2, -3, TLEN, TSTART, TEND, NMATCH, ALEN, MAPQ};
  while (std::getline(f, line)) {
    ++lineno;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty()) continue;  // blank.lines.skip = TRUE
    if (dedupe && !seen.insert(line).second) continue;
    std::vector<std::string> fld;
    size_t a = 0;
    for (int k = 0; k < 12; ++k) {
      const size_t b = line.find('\t', a);
      if (b == std::string::npos && k < 11) pfail(NW_ERR_PARSE, std::string(path) + ": line " + std::to_string(lineno) + " has fewer than 12 columns");
      fld.push_back(line.substr(a, b == std::string::npos ? std::string::npos : b - a));
      a = b == std::string::npos ? line.size() : b + 1;
    }
    const size_t row = P.qname.size();
    P.qname.push_back(fld[0]);
    P.tname.push_back(fld[5]);
    P.extra.push_back(a < line.size() ? line.substr(a) : std::string());
    for (int k = 0; k < 12; ++k) {
      if (numcol[k] == -1 || numcol[k] == -3) continue;
      if (numcol[k] == -2) {
        if (fld[k] != "+" && fld[k] != "-") pfail(NW_ERR_PARSE, std::string(path) + ": line " + std::to_string(lineno) + ": strand must be + or -");
        P.t.c[STRAND].push_back(fld[k] == "+" ? 1.0 : -1.0);
        continue;
      }
      char* end = nullptr;
      const double v = strtod(fld[k].c_str(), &end);
      if (end == fld[k].c_str() || *end != 0 || !std::isfinite(v))
        pfail(NW_ERR_PARSE, std::string(path) + ": line " + std::to_string(lineno) + ": column " + std::to_string(k + 1) + " is not a number");
      P.t.c[numcol[k]].push_back(v);
      if (v != std::floor(v) || std::fabs(v) > 2147483647.0 || fld[k].find_first_of(".eE") != std::string::npos)
        P.integer_typed[numcol[k]] = false;
    }
    P.t.c[ROWID].push_back((double)row);
  }
  return P;
}

// write.table(inpaf, quote = F, col.names = F, row.names = F, sep = "\t") after the final strand swap
void write_paf_file(const char* path, const Table& t, const std::vector<std::string>& qname, const PafFile& src, bool merged,
                    bool append = false) {
  const std::string tmp = append ? std::string(path) : std::string(path) + ".tmp";
  {
    std::ofstream o(tmp, append ? std::ios::app : std::ios::trunc);
    if (!o) pfail(NW_ERR_IO, std::string("cannot write ") + tmp);
    bool it[10];
    for (int k = 0; k < 10; ++k) it[k] = src.integer_typed[k];
    if (merged) it[NMATCH] = it[ALEN] = false;  // merge_rows: as.numeric(paffile$nmatch), as.numeric(paffile$alen)
    for (size_t k = 0; k < t.n(); ++k) {
      const size_t row = (size_t)t.c[ROWID][k];
      const bool neg = t.c[STRAND][k] < 0;
      const double qs = neg ? t.c[QEND][k] : t.c[QSTART][k], qe = neg ? t.c[QSTART][k] : t.c[QEND][k];
      o << qname[k] << '\t' << r_format(t.c[QLEN][k], it[QLEN]) << '\t' << r_format(qs, it[QSTART] && it[QEND]) << '\t'
        << r_format(qe, it[QSTART] && it[QEND]) << '\t' << (neg ? '-' : '+') << '\t' << src.tname[row] << '\t'
        << r_format(t.c[TLEN][k], it[TLEN]) << '\t' << r_format(t.c[TSTART][k], it[TSTART]) << '\t'
        << r_format(t.c[TEND][k], it[TEND]) << '\t' << r_format(t.c[NMATCH][k], it[NMATCH]) << '\t'
        << r_format(t.c[ALEN][k], it[ALEN]) << '\t' << r_format(t.c[MAPQ][k], it[MAPQ]);
      if (!merged && !src.extra[row].empty()) o << '\t' << src.extra[row];  // merge_rows rebuilds the 12 columns only
      o << '\n';
    }
    o.flush();
    if (!o) pfail(NW_ERR_IO, std::string("cannot write ") + tmp);
  }
  if (!append && rename(tmp.c_str(), path) != 0) pfail(NW_ERR_IO, std::string("cannot rename to ") + path);
}

}  // namespace

extern "C" {

int nw_paf_read_file(const char* path, nw_paf_table** out) {
  if (out) *out = nullptr;
  return pguard([&] {
    if (!out) pfail(NW_ERR_ARG, "null argument");
    PafFile P = read_paf_file(path, false);
    std::unique_ptr<nw_paf_table> T(new nw_paf_table);
    T->t = std::move(P.t);
    *out = T.release();
  });
}

int nw_paf_compress_file(nw_ctx* ctx, const char* inpaf_path, const char* outpaf_path, double chunklen, int32_t n_quadrants,
                         int64_t* rows_in, int64_t* rows_out) {
  return pguard([&] {
    if (!ctx || !outpaf_path) pfail(NW_ERR_ARG, "null argument");
    PafFile P = read_paf_file(inpaf_path, true);  // read.table + unique()
    if (rows_in) *rows_in = (int64_t)P.t.n();
    swap_minus(P.t);  // :196-200
    CompressJob J;
    J.second_run = false;
    J.chunklen = chunklen;
    J.n_quadrants = n_quadrants;
    J.in_qname = &P.qname;
    J.q_integer = P.integer_typed[QSTART] && P.integer_typed[QEND];
    compress_pre(J, P.t);
    compress_gpu(ctx, {&J});
    Table out = compress_post(J);
    if (!J.merged) {
      J.out_qname.clear();
      for (size_t k = 0; k < out.n(); ++k) J.out_qname.push_back(P.qname[(size_t)out.c[ROWID][k]]);
    }
    write_paf_file(outpaf_path, out, J.out_qname, P, J.merged);
    if (rows_out) *rows_out = (int64_t)out.n();
  });
}

int nw_paf_compress_file_by_qname(nw_ctx* ctx, const char* inpaf_path, const char* outpaf_path, double chunklen,
                                  int32_t n_quadrants, int64_t* rows_in, int64_t* rows_out) {
  return pguard([&] {
    if (!ctx || !outpaf_path) pfail(NW_ERR_ARG, "null argument");
    PafFile P = read_paf_file(inpaf_path, false);
    if (rows_in) *rows_in = (int64_t)P.t.n();
    // paf$V1 = sub("_[0-9]+-[0-9]+$", "", paf$V1) (R/make_whole_genome_list.R:81)
    for (std::string& q : P.qname) {
      size_t e = q.size(), d2 = e;
      while (d2 > 0 && isdigit((unsigned char)q[d2 - 1])) --d2;
      if (d2 == e || d2 == 0 || q[d2 - 1] != '-') continue;
      size_t d1 = d2 - 1;
      while (d1 > 0 && isdigit((unsigned char)q[d1 - 1])) --d1;
      if (d1 == d2 - 1 || d1 == 0 || q[d1 - 1] != '_') continue;
      q.erase(d1 - 1);
    }
    swap_minus(P.t);  // :97-101
    // one table per query sequence, in the order of unique(paf$qname)
    std::vector<std::string> order;
    std::map<std::string, std::vector<size_t>> rows_of;
    for (size_t k = 0; k < P.t.n(); ++k) {
      auto it = rows_of.find(P.qname[k]);
      if (it == rows_of.end()) {
        order.push_back(P.qname[k]);
        it = rows_of.emplace(P.qname[k], std::vector<size_t>()).first;
      }
      it->second.push_back(k);
    }
    std::vector<CompressJob> jobs(order.size());
    std::vector<CompressJob*> run;
    for (size_t g = 0; g < order.size(); ++g) {
      CompressJob& J = jobs[g];
      J.second_run = false;
      J.chunklen = chunklen;
      J.n_quadrants = n_quadrants;
      J.in_qname = &P.qname;
      J.q_integer = P.integer_typed[QSTART] && P.integer_typed[QEND];
      J.reduced_qnames = true;  // merge_rows(inpaf, rowpairs_singular, !is.null(qname))
      compress_pre(J, take(P.t, rows_of[order[g]]));
      run.push_back(&J);
    }
    compress_gpu(ctx, run);  // the all-pairs tests of every query sequence in one launch
    int64_t total = 0;
    for (size_t g = 0; g < order.size(); ++g) {
      CompressJob& J = jobs[g];
      Table out = compress_post(J);
      if (!J.merged) {
        J.out_qname.clear();
        for (size_t k = 0; k < out.n(); ++k) J.out_qname.push_back(P.qname[(size_t)out.c[ROWID][k]]);
      }
      write_paf_file(outpaf_path, out, J.out_qname, P, J.merged, true);  // append = !is.null(qname)
      total += (int64_t)out.n();
    }
    if (rows_out) *rows_out = total;
  });
}

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
    static const bool dbg = getenv("NW_DEBUG_TIMING") != nullptr;
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
    std::vector<PrepJob> jobs((size_t)n_regions);
    while (!pending.empty()) {
      std::vector<int64_t> next, togrid, prepared;
      std::vector<CompressJob*> run;
      std::vector<std::string> msg(pending.size());
      const double t0 = paf_now();
      paf_parallel_for((int64_t)pending.size(), [&](int64_t q) {  // host part of the preparation, tables in parallel
        const int64_t k = pending[q];
        Region& r = R[k];
        r.info.passes += 1;
        r.info.minlen = r.minlen;
        r.info.compression = r.compression;
        try {
          prepare_pre(jobs[k], from_cols(&pafs[k]), r.minlen, r.compression, params->chunklen);
        } catch (const PErr& e) {
          r.rc = e.code;
          msg[q] = e.msg;
        }
      });
      for (size_t q = 0; q < pending.size(); ++q) {
        const int64_t k = pending[q];
        if (R[k].rc == NW_OK) {
          run.push_back(&jobs[k].cj);
          prepared.push_back(k);
        } else {
          nw::set_last_error(msg[q]);
        }
      }
      const double t1 = paf_now();
      compress_gpu(ctx, run);  // one all-pairs launch for the whole pass
      const double t2 = paf_now();
      paf_parallel_for((int64_t)prepared.size(), [&](int64_t q) {
        const int64_t k = prepared[q];
        R[k].tab.reset(new nw_paf_table);
        R[k].tab->t = prepare_post(jobs[k], R[k].tab->stats);
        jobs[k] = PrepJob();
      });
      const double t3 = paf_now();
      for (int64_t k : prepared) {
        Region& r = R[k];
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
      const double t4 = paf_now();
      double t5 = t4;
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
        t5 = paf_now();
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
      if (dbg)
        fprintf(stderr, "[nw_bitlocus_build] pass of %zu regions: paf pre %.2f ms, gpu %.2f ms, post %.2f ms, checks %.2f ms, "
                        "grid %.2f ms, matrices + pairs %.2f ms\n", pending.size(), 1e3 * (t1 - t0), 1e3 * (t2 - t1),
                1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t5 - t4), 1e3 * (paf_now() - t5));
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
