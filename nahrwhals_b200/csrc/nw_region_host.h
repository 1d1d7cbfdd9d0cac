// nw_region_host.h -- host-only part of the region stage (include/nwregion.h): the model of the
// reference's matrices as position lists, flip / find_sv_opportunities / carry_out_compressed_sv, the visited-set key,
// the result tables, format_julia_output, and the three host phases of the pre-pass (prepare, replay of dfsutil's
// sequential part over the per-child numbers, summary).  Pure C++17, no CUDA: nw_region.cu adds the GPU phase and the
// C ABI; tests/harness/nw_region_harness.cpp compiles this header alone so that `pytest -m "not gpu"` checks the host
// logic against the R restatement (oracle/nw_region_oracle.py) with the oracle's per-child numbers standing in for
// the GPU's.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/nwregion.h"

namespace {


const double kNA = std::numeric_limits<double>::quiet_NaN();

struct RErr {
  int code;
  std::string msg;
};
[[noreturn]] void rfail(int code, const std::string& m) { throw RErr{code, m}; }

// ---------------------------------------------------------------------------------------------------------------
// host model of the reference's matrices
// ---------------------------------------------------------------------------------------------------------------
struct Grid {
  int ny = 0, nx = 0;
  std::vector<int64_t> v;       // v[y*nx + x], signed bp
  std::vector<int64_t> rn, cn;  // row.names (y lengths), colnames (x lengths)
  int64_t at(int y, int x) const { return v[(size_t)y * nx + x]; }
};

struct Pos {
  int16_t blk;  // signed 1-based x block of the (flipped) gridmatrix
  int64_t len;  // colname of the position as R leaves it
};
typedef std::vector<Pos> Arr;

struct Move {
  int p1, p2;
  int sv;  // 0 del, 1 dup, 2 inv   (names below)
};
const char* const kSvName[3] = {"del", "dup", "inv"};

double r_round0(double x) { return std::nearbyint(x); }  // R round(x): IEC 60559 half-even (default rounding mode)

// R/flip_bitl_y_if_needed.R:12-76
Grid flip_if_needed(const Grid& g, bool& flipped, bool& unsure) {
  std::vector<int64_t> f = g.v;
  const int nr = g.ny, nc = g.nx;
  if ((nr >= 4 || nc >= 4) && (nr >= 2 && nc >= 2)) {
    const int y_min = (int)std::max(r_round0(nr * (1.0 / 4.0)), 2.0), y_max = (int)std::min(r_round0(nr * (3.0 / 4.0)), nr - 2.0);
    const int x_min = (int)std::max(r_round0(nc * (1.0 / 4.0)), 2.0), x_max = (int)std::min(r_round0(nc * (3.0 / 4.0)), nc - 2.0);
    // a:z counts down when a > z; subscript 0 selects nothing
    for (int y = std::min(y_min, y_max); y <= std::max(y_min, y_max); ++y)
      if (y >= 1 && y <= nr)
        for (int x = 0; x < nc; ++x) f[(size_t)(y - 1) * nc + x] = 0;
    for (int x = std::min(x_min, x_max); x <= std::max(x_min, x_max); ++x)
      if (x >= 1 && x <= nc)
        for (int y = 0; y < nr; ++y) f[(size_t)y * nc + (x - 1)] = 0;
  }
  double pos = 0, neg = 0;
  for (int64_t t : f) {
    if (t > 0) pos += (double)t;
    if (t < 0) neg += (double)t;
  }
  neg = std::fabs(neg);
  if (pos + neg == 0) pos = 1;
  unsure = (std::max(pos, neg) / std::min(pos, neg)) < 2;  // x/0 = Inf
  flipped = neg > pos;
  if (!flipped) return g;
  Grid o = g;
  if (nr == 1) {
    for (auto& t : o.v) t = -t;
  } else {  // -apply(bitl, 2, rev): rows and their names reversed, signs negated
    for (int y = 0; y < nr; ++y) {
      o.rn[y] = g.rn[nr - 1 - y];
      for (int x = 0; x < nc; ++x) o.v[(size_t)y * nc + x] = -g.at(nr - 1 - y, x);
    }
  }
  return o;
}

// R/find_sv_opportunities.R:12-98
std::vector<Move> find_sv_opportunities(const Grid& g) {
  struct Key {
    int p1, p2, ori;
  };
  std::vector<Key> rows;
  std::map<std::pair<std::pair<int, int>, int>, int> seen;
  std::vector<int> nz;
  for (int y = 0; y < g.ny; ++y) {
    nz.clear();
    for (int x = 0; x < g.nx; ++x)
      if (g.at(y, x) != 0) nz.push_back(x + 1);
    if (nz.size() < 2) continue;
    for (size_t a = 0; a < nz.size(); ++a)
      for (size_t c = a + 1; c < nz.size(); ++c) {
        const int p1 = nz[a], p2 = nz[c];
        const int ori = ((g.at(y, p1 - 1) > 0) == (g.at(y, p2 - 1) > 0)) ? 1 : -1;
        auto k = std::make_pair(std::make_pair(p1, p2), ori);
        if (seen.emplace(k, 1).second) rows.push_back(Key{p1, p2, ori});
      }
  }
  std::vector<Move> out;
  for (const Key& k : rows) out.push_back(Move{k.p1, k.p2, k.ori == -1 ? 2 : 0});
  for (const Key& k : rows)
    if (k.ori == 1) out.push_back(Move{k.p1, k.p2, 1});
  std::vector<Move> keep;
  for (const Move& m : out)
    if (!(m.p2 - m.p1 == 1 && m.sv == 2)) keep.push_back(m);
  return keep;
}

// carry_out_compressed_sv (R/seqmodder.R:10-109) on the reference arrangement `a` (positions 1..n), including the
// manual colname repairs of :25-31, :52, :63, :83-87, :93-97, which move LENGTHS between positions.
Arr carry_out(const Arr& a, const Move& mv) {
  const int n = (int)a.size(), p1 = mv.p1, p2 = mv.p2;
  Arr o;
  auto put = [&](int i, bool neg) {
    Pos t = a[i - 1];
    if (neg) t.blk = (int16_t)-t.blk;
    o.push_back(t);
  };
  if (mv.sv == 0) {  // del: bitl[, -c(p1:(p2-1))]
    for (int i = 1; i <= n; ++i)
      if (!(i >= p1 && i <= p2 - 1)) put(i, false);
    if (o.size() == 1) o[0].len = a[0].len;  // one column left: colnames(bitl)[1]
  } else if (mv.sv == 1) {  // dup: cbind(bitl[, 1:p2], bitl[, (p1+1):n])
    for (int i = 1; i <= p2; ++i) put(i, false);
    for (int i = p1 + 1; i <= n; ++i) put(i, false);
    if (p2 - p1 == 1) {
      if (p1 == 1)
        o.back().len = o[1].len;
      else if (p2 == n)
        o.back().len = o[o.size() - 2].len;
    }
  } else {  // inv
    if (p2 - p1 == 1) return a;
    if (p1 == 1 && p2 != n) {  // border case 1: the flanks themselves are inverted
      for (int i = p2; i >= p1; --i) put(i, true);
      for (int i = p2 + 1; i <= n; ++i) put(i, false);
      o.back().len = a[o.size() - 1].len;
    } else if (p1 > 1 && p2 == n) {  // border case 2
      for (int i = 1; i <= p1 - 1; ++i) put(i, false);
      for (int i = p2; i >= p1; --i) put(i, true);
      o[0].len = a[0].len;
    } else if (p1 == 1 && p2 == n) {  // border case 3
      for (int i = p2; i >= p1; --i) put(i, true);
    } else {  // the standard case keeps both flanks
      for (int i = 1; i <= p1; ++i) put(i, false);
      for (int i = p2 - 1; i >= p1 + 1; --i) put(i, true);
      for (int i = p2; i <= n; ++i) put(i, false);
    }
    if (p2 - p1 == 2) {
      for (int i = 0; i < n; ++i) o[i].len = a[i].len;  // colnames(bitl_mut) <- colnames(bitl)
    } else if (p2 == n) {
      o[p2 - 1].len = o[p1 - 1].len;
    }
  }
  return o;
}

inline int64_t cell(const Grid& g, const Arr& a, int y, int p) {
  const int v = a[p].blk;
  const int64_t t = g.at(y, std::abs(v) - 1);
  return v < 0 ? -t : t;
}

// return_diag_values_new(matrix, fraction = 0.05) (helpers.R:98-113): the values rlang::hash() is fed
void diag_values(const Grid& g, const Arr& a, std::vector<int64_t>& out) {
  out.clear();
  const int64_t nr = g.ny, nc = (int64_t)a.size(), mn = std::min(nr, nc);
  const int64_t md = (int64_t)r_round0(std::max(5.0, 0.05 * (double)nr));
  const int64_t last = (nr + 1) * (mn - 1) + 1;
  auto value = [&](int64_t idx) { return cell(g, a, (int)((idx - 1) % nr), (int)((idx - 1) / nr)); };
  std::vector<int64_t> clean;
  for (int64_t o = -md; o <= md; ++o)  // outer() is flattened column-major: offsets outer, diagonal cells inner
    for (int64_t k = 0; k < mn; ++k) {
      const int64_t idx = 1 + k * (nr + 1) + o;
      if (idx > 0 && idx <= last) clean.push_back(idx);
    }
  for (int64_t idx : clean) out.push_back(value(idx));
  for (int64_t idx : clean) {
    const int64_t j = nr * nc - idx;
    if (j == 0) continue;  // zero subscript selects nothing
    if (j < 0 || j > nr * nc) rfail(NW_ERR_ARG, "return_diag_values_new: subscript outside the matrix");
    out.push_back(value(j));
  }
}

// 128-bit hash of the same value sequence without materialising it (the visited set of dfsutil is keyed by
// rlang::hash of that vector); equal hashes are verified element-wise by the caller.
struct DiagKey {
  uint64_t h1, h2, n;
  bool operator==(const DiagKey& o) const { return h1 == o.h1 && h2 == o.h2 && n == o.n; }
};
struct DiagKeyHash {
  size_t operator()(const DiagKey& k) const { return (size_t)(k.h1 ^ (k.h2 * 0x9e3779b97f4a7c15ull)); }
};
DiagKey diag_key(const Grid& g, const Arr& a) {
  const int64_t nr = g.ny, nc = (int64_t)a.size(), mn = std::min(nr, nc);
  const int64_t md = (int64_t)r_round0(std::max(5.0, 0.05 * (double)nr));
  const int64_t last = (nr + 1) * (mn - 1) + 1;
  DiagKey k{0x243f6a8885a308d3ull, 0x13198a2e03707344ull, 0};
  auto mix = [&](int64_t v) {
    k.h1 = (k.h1 ^ (uint64_t)v) * 0x100000001b3ull;
    k.h1 ^= k.h1 >> 29;
    k.h2 = (k.h2 + (uint64_t)v + 0x9e3779b97f4a7c15ull) * 0xbf58476d1ce4e5b9ull;
    k.h2 ^= k.h2 >> 31;
    ++k.n;
  };
  auto value = [&](int64_t idx) { return cell(g, a, (int)((idx - 1) % nr), (int)((idx - 1) / nr)); };
  for (int pass = 0; pass < 2; ++pass)
    for (int64_t o = -md; o <= md; ++o)
      for (int64_t q = 0; q < mn; ++q) {
        const int64_t idx = 1 + q * (nr + 1) + o;
        if (!(idx > 0 && idx <= last)) continue;
        if (pass == 0) {
          mix(value(idx));
        } else {
          const int64_t j = nr * nc - idx;
          if (j == 0) continue;
          if (j < 0 || j > nr * nc) rfail(NW_ERR_ARG, "return_diag_values_new: subscript outside the matrix");
          mix(value(j));
        }
      }
  return k;
}

// is_cluttered (helpers.R:209-235)
bool is_cluttered(const Grid& g) {
  const int nr = g.ny, nc = g.nx;
  if (!(nr > 5 && nc > 5)) return false;
  int c1 = 0, c2 = 0, c3 = 0, c4 = 0;
  for (int x = 1; x <= nc - 2; ++x) {
    c1 += g.at(0, x) != 0;
    c2 += g.at(nr - 1, x) != 0;
  }
  for (int y = 1; y <= nr - 2; ++y) {
    c3 += g.at(y, 0) != 0;
    c4 += g.at(y, nc - 1) != 0;
  }
  return c1 >= 5 || c2 >= 5 || c3 >= 5 || c4 >= 5;
}

double round3(double x) {  // round(x, 3) for the non-tie values that occur; identical to the solver's rounding
  const double y = std::nearbyint(x * 1000.0) / 1000.0;
  return std::isfinite(y) ? y : x;
}


}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// result tables
// ---------------------------------------------------------------------------------------------------------------
struct nw_table {
  // kind 0: table of solve_mutation (eval, mut1, mut1_start, mut1_end, mut1_pos_pm, mut1_len, mut1_len_pm, mut2_len,
  //         mut2_len_pm, mut3_len, mut3_len_pm);  kind 1: table of format_julia_output (eval, mut1..mutN,
  //         mut{i}_start, mut{i}_end, mut{i}_len for i = 1..N);  kind 2: format_julia_output's (eval, nmut, mut1 = ref)
  int kind = 0;
  int n_mut = 1;
  struct Row {
    double eval = 0;
    int nmoves = 0;
    std::vector<std::string> mut;  // "" = NA
    std::vector<double> num;       // NaN = NA; kind 0: 9 values, kind 1: 3 per mutation
  };
  std::vector<Row> rows;
  mutable std::string tsv;  // rendered on first use
  mutable bool tsv_ready = false;
  std::string mut_max;
  nw_region_summary summ{};
  std::vector<nw_candidate> cands;  // pre-pass diagnostics
};


namespace {


int na_count(const nw_table& t, const nw_table::Row& r) {
  int k = 0;
  for (const std::string& m : r.mut) k += m.empty();
  for (double v : r.num) k += std::isnan(v);
  (void)t;
  return k;
}

void sort_eval_na(nw_table& t) {  // order(-eval, -rowSums(is.na(.))), stable
  std::vector<int> na(t.rows.size());
  std::vector<size_t> ord(t.rows.size());
  for (size_t k = 0; k < t.rows.size(); ++k) {
    na[k] = na_count(t, t.rows[k]);
    ord[k] = k;
  }
  std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) {
    if (t.rows[a].eval != t.rows[b].eval) return t.rows[a].eval > t.rows[b].eval;
    return na[a] > na[b];
  });
  std::vector<nw_table::Row> r2;
  r2.reserve(ord.size());
  for (size_t k : ord) r2.push_back(std::move(t.rows[k]));
  t.rows.swap(r2);
}

void fmt_num(std::string& s, double v) {  // shortest text that reads back as v
  if (std::isnan(v)) {
    s += "NA";
    return;
  }
  char buf[40];
  if (v == std::floor(v) && std::fabs(v) < 1e15) {
    snprintf(buf, sizeof buf, "%lld", (long long)v);
  } else {
    snprintf(buf, sizeof buf, "%.15g", v);  // 3-decimal scores and x.5 lengths come out shortest already
    if (strtod(buf, nullptr) != v) snprintf(buf, sizeof buf, "%.17g", v);
  }
  s += buf;
}

// save_to_logfile (R/save_logfile.R:17-51) + text form
void render_tsv(const nw_table& t, std::string& s) {
  s.clear();
  if (t.kind == 0) {
    s = "eval\tmut1\tmut1_start\tmut1_end\tmut1_pos_pm\tmut1_len\tmut1_len_pm\tmut2_len\tmut2_len_pm\tmut3_len\tmut3_len_pm\n";
  } else if (t.kind == 2) {
    s = "eval\tnmut\tmut1\n";
  } else {
    s = "eval";
    for (int i = 1; i <= t.n_mut; ++i) s += "\tmut" + std::to_string(i);
    for (int i = 1; i <= t.n_mut; ++i)
      for (const char* suf : {"_start", "_end", "_len"}) s += "\tmut" + std::to_string(i) + suf;
    s += "\n";
  }
  s.reserve(s.size() + t.rows.size() * (size_t)(16 + 40 * t.n_mut));
  for (const auto& r : t.rows) {
    fmt_num(s, r.eval);
    if (t.kind == 2) s += "\t" + std::to_string(r.nmoves);
    for (const std::string& m : r.mut) {
      s += '\t';
      if (m.empty())
        s += "NA";
      else
        s += m;
    }
    for (double v : r.num) {
      s += '\t';
      fmt_num(s, v);
    }
    s += '\n';
  }
}

// the logfile summary of a table (the text form is rendered on demand by nw_table_tsv)
void finish_table(nw_table& t) {
  t.tsv.clear();
  t.tsv_ready = false;
  nw_region_summary& S = t.summ;
  S.n_rows = (int32_t)t.rows.size();
  S.n_mut = t.n_mut;
  for (double& v : S.maxres) v = kNA;
  S.res_ref = S.res_max = kNA;
  S.n_res_max = 0;
  t.mut_max = "ref";
  if (t.rows.empty()) return;
  int nref = 0;
  for (const auto& r : t.rows)
    if (!r.mut.empty() && r.mut[0] == "ref") {
      ++nref;
      S.res_ref = r.eval;
    }
  if (nref != 1) S.res_ref = kNA;
  S.res_max = t.rows[0].eval;
  double mx = t.rows[0].eval;
  for (const auto& r : t.rows) mx = std::max(mx, r.eval);
  // which.max over the max-eval subset of the NA count in mut1..mut3; the position is then used as a row number of
  // the full table (identical when the table is sorted by eval, which it is on this path)
  int pos = 0, best = -1, sub = 0;
  for (const auto& r : t.rows) {
    if (r.eval != mx) continue;
    ++S.n_res_max;
    int k = 0;
    for (int i = 0; i < std::min<int>(3, (int)r.mut.size()); ++i) k += r.mut[i].empty();
    if (k > best) {
      best = k;
      pos = sub;
    }
    ++sub;
  }
  const auto& m = t.rows[std::min<size_t>(pos, t.rows.size() - 1)];
  std::string mm;
  int nmut = 0;
  for (int i = 0; i < std::min<int>(3, (int)m.mut.size()); ++i) nmut += !m.mut[i].empty();
  for (int i = 0; i < nmut; ++i) mm += (i ? "+" : "") + m.mut[i];
  t.mut_max = mm;
  if (t.kind == 0) {  // mut1_start, mut1_end, mut1_pos_pm, mut1_len, mut1_len_pm, mut2_len, mut2_len_pm, mut3_len, mut3_len_pm
    for (int k = 0; k < 5; ++k) S.maxres[k] = m.num[k];
    S.maxres[8] = m.num[5];
    S.maxres[9] = m.num[6];
    S.maxres[13] = m.num[7];
    S.maxres[14] = m.num[8];
  } else if (t.kind == 1) {
    for (int i = 0; i < std::min(3, t.n_mut); ++i) {
      S.maxres[i * 5 + 0] = m.num[i * 3 + 0];
      S.maxres[i * 5 + 1] = m.num[i * 3 + 1];
      S.maxres[i * 5 + 3] = m.num[i * 3 + 2];
    }
  }
}

// turn_mut_max_into_svdf(t, correction = T) (R/juliasolver.R:138-245) on the moves of one trajectory
struct Sv {
  double start, end;
  int kind;  // 0 del, 1 dup, 2 inv
};
void correct_coordinates(std::vector<Sv>& v) {
  const int n = (int)v.size();
  if (n <= 1) return;
  std::vector<double> st(n), en(n);
  for (int i = 0; i < n - 1; ++i) {
    const double s0 = v[i].start, e0 = v[i].end;
    for (int k = 0; k < n; ++k) {
      st[k] = v[k].start;
      en[k] = v[k].end;
    }
    if (v[i].kind == 2) {
      for (int k = i + 1; k < n; ++k) {
        const bool cs = st[k] > s0 && st[k] < e0, ce = en[k] > s0 && en[k] < e0;
        if (ce) v[k].end = e0 - (en[k] - s0);
        if (cs) v[k].start = e0 - (st[k] - s0);
      }
      for (int k = 0; k < n; ++k)
        if (v[k].start > v[k].end) std::swap(v[k].start, v[k].end);
    }
    if (v[i].kind == 0) {
      for (int k = i + 1; k < n; ++k) {
        if (st[k] > s0) v[k].start += (e0 - s0);
        if (en[k] > s0) v[k].end += (e0 - s0);
      }
    } else if (v[i].kind == 1) {
      for (int k = i + 1; k < n; ++k) {
        if (st[k] > e0) v[k].start -= (e0 - s0);
        if (en[k] > e0) v[k].end -= (e0 - s0);
      }
    }
  }
}

// One reported trajectory as format_julia_output sees it: the line's eval, its move count (column V2) and the moves.
struct TrajIn {
  double eval;
  int nmoves;
  uint32_t mv_off, mv_n;
};

// Block 1 + 2 + sort of format_julia_output (R/juliasolver.R:68-117) on parsed trajectories.  ncols = the widest line
// of the report in fields (count.fields): 3 for a 0-move line ("score<TAB>0<TAB>"), 2 + 3m for m moves.
nw_table* table_from_trajectories(const std::vector<TrajIn>& tr, const std::vector<Sv>& mv, size_t ncols, const double* gl,
                                  int n_gl) {
  if (tr.empty()) rfail(NW_ERR_PARSE, "solver report is empty");  // read.table stops: no lines available
  if (ncols < 2) rfail(NW_ERR_PARSE, "solver report needs at least two columns");
  const int n_mut = (int)((ncols - 2) / 3);
  std::unique_ptr<nw_table> T(new nw_table);
  if (n_mut == 0) {  // data.frame(eval = jout[1], nmut = jout[2], mut1 = 'ref')
    T->kind = 2;
    T->n_mut = 1;
    for (const TrajIn& t : tr) {
      nw_table::Row r;
      r.eval = t.eval;
      r.nmoves = t.nmoves;
      r.mut = {"ref"};
      T->rows.push_back(std::move(r));
    }
    finish_table(*T);
    return T.release();
  }
  T->kind = 1;
  T->n_mut = n_mut;
  T->rows.reserve(tr.size() + 1);
  static const char* const names[3] = {"del", "dup", "inv"};
  std::vector<Sv> svs;
  for (const TrajIn& t : tr) {
    nw_table::Row r;
    r.eval = t.eval;
    r.nmoves = t.nmoves;
    r.mut.assign(n_mut, "");
    r.num.assign((size_t)3 * n_mut, kNA);
    svs.assign(mv.begin() + t.mv_off, mv.begin() + t.mv_off + t.mv_n);
    for (size_t i = 0; i < svs.size() && (int)i < n_mut; ++i) {
      char buf[96];
      snprintf(buf, sizeof buf, "%.15g_%.15g_%s", svs[i].start, svs[i].end, names[svs[i].kind]);  // concat_triplet
      r.mut[i] = buf;
    }
    if (r.nmoves != 0) {
      correct_coordinates(svs);  // process_row + turn_mut_max_into_svdf
      for (size_t i = 0; i < svs.size() && (int)i < n_mut; ++i) {
        auto at = [&](double p) {
          const long long k = (long long)p;  // R truncates a double subscript
          return (k >= 1 && k <= n_gl) ? gl[k - 1] : kNA;
        };
        const double a = at(svs[i].start), z = at(svs[i].end);
        r.num[i * 3 + 0] = a;
        r.num[i * 3 + 1] = z;
        r.num[i * 3 + 2] = z - a;  // NA propagates
      }
    }
    T->rows.push_back(std::move(r));
  }
  sort_eval_na(*T);
  for (auto& r : T->rows)
    if (r.nmoves == 0) r.mut[0] = "ref";
  finish_table(*T);
  return T.release();
}

// the bytes of a solver report (what read.table parses, R/juliasolver.R:58-59)
nw_table* format_solver_output(const char* tsv, int64_t len, const double* gl, int n_gl) {
  std::vector<TrajIn> tr;
  std::vector<Sv> mv;
  size_t ncols = 0;
  std::vector<std::pair<int64_t, int64_t>> f;  // [begin, end) of the fields of one line
  auto num = [&](std::pair<int64_t, int64_t> q) {
    const std::string t(tsv + q.first, tsv + q.second);
    char* e = nullptr;
    const double v = strtod(t.c_str(), &e);
    if (e == t.c_str()) rfail(NW_ERR_PARSE, "solver report: not a number: '" + t + "'");
    return v;
  };
  int64_t a = 0;
  while (a < len) {
    int64_t b = a;
    while (b < len && tsv[b] != '\n') ++b;
    if (b > a) {
      f.clear();
      int64_t s0 = a;
      for (int64_t k = a; k <= b; ++k)
        if (k == b || tsv[k] == '\t') {
          f.emplace_back(s0, k);
          s0 = k + 1;
        }
      ncols = std::max(ncols, f.size());
      TrajIn t;
      t.eval = num(f[0]);
      t.nmoves = f.size() > 1 ? (int)num(f[1]) : 0;
      t.mv_off = (uint32_t)mv.size();
      for (size_t c = 2; c + 2 < f.size(); c += 3) {  // read.table(fill = TRUE): short lines are padded with NA
        if (f[c].first == f[c].second) break;
        std::string kind(tsv + f[c].first, tsv + f[c].second);
        if (kind.compare(0, 5, "move_") == 0) kind = kind.substr(5);
        const int k = kind == "del" ? 0 : kind == "dup" ? 1 : kind == "inv" ? 2 : -1;
        if (k < 0) rfail(NW_ERR_PARSE, "solver report: unknown move '" + kind + "'");
        mv.push_back(Sv{num(f[c + 1]), num(f[c + 2]), k});
      }
      t.mv_n = (uint32_t)mv.size() - t.mv_off;
      tr.push_back(t);
    }
    a = b + 1;
  }
  return table_from_trajectories(tr, mv, ncols, gl, n_gl);
}

// ---------------------------------------------------------------------------------------------------------------
// the depth-1 pre-pass
// ---------------------------------------------------------------------------------------------------------------
struct PrepassOut {
  std::unique_ptr<nw_table> table;
  Grid g0, flipped;
  bool did_flip = false, unsure = false, cluttered = false;
  bool trivial = false;  // 1 x 1 gridmatrix: calc_coarse_grained_aln_score returns 100, there is nothing to search
  std::vector<Move> moves;  // find_sv_opportunities of the flipped matrix
  int64_t n_cand = 0, n_dp = 0;
  int maxdepth = 1;
  int launches = 0;
  // --- work of one region between its three phases (host prepare -> GPU batch -> host replay) ---
  std::vector<Arr> cand;                   // children of the reference arrangement (cand[k-1] <-> entry k)
  std::vector<std::vector<int16_t>> seqs;  // entry 0 = the reference arrangement; block ids incl. virtual blocks
  std::vector<int64_t> lenx;               // x lengths incl. one virtual block per (block, length) a colname repair made
  std::vector<int8_t> aln;                 // sign matrix of those blocks, x-major
  int nxa = 0;
  std::vector<long long> srow, scol;       // estimate sums per entry
  std::vector<int64_t> ccost, ctotal;      // DP per entry (entry 0: forced, full corridor)
  std::vector<double> cscore;
};

Grid make_grid(const int64_t* grid, int n_x, int n_y, const int64_t* len_x, const int64_t* len_y) {
  if (!grid || !len_x || !len_y) rfail(NW_ERR_ARG, "null argument");
  // 1 x 1 (one clean alignment: the region carries no SD) is the reference's "report 100%" case (R/evaltools.R:100-101);
  // a single row or column with several blocks is not reproduced (R drops such matrices to vectors inside
  // carry_out_compressed_sv, R/seqmodder.R, and the outcome depends on cbind's recycling)
  if ((n_x < 2 || n_y < 2) && !(n_x == 1 && n_y == 1)) rfail(NW_ERR_ARG, "the gridmatrix needs at least 2 x 2 blocks (or 1 x 1)");
  if (n_x > 16000 || n_y > 16000) rfail(NW_ERR_ARG, "gridmatrix too large");
  Grid g;
  g.nx = n_x;
  g.ny = n_y;
  g.v.assign(grid, grid + (size_t)n_x * n_y);
  g.rn.assign(len_y, len_y + n_y);
  g.cn.assign(len_x, len_x + n_x);
  for (int64_t t : g.rn)
    if (t <= 0) rfail(NW_ERR_RANGE, "non-positive y block length");
  for (int64_t t : g.cn)
    if (t <= 0) rfail(NW_ERR_RANGE, "non-positive x block length");
  return g;
}

void annotate_row(nw_table::Row& r, const Grid& g, int p1, int p2) {
  // annotate_res_out_with_positions_lens (helpers.R:412-485), mut1 only.  sum(colnumbers[1:(p-1)]): 1:0 selects element 1
  auto lead = [&](int p) {
    double s = 0;
    if (p >= 2)
      for (int i = 0; i < p - 1; ++i) s += (double)g.cn[i];
    else
      s = (double)g.cn[0];
    return s;
  };
  const double start_mean = r_round0(lead(p1) + (double)g.cn[p1 - 1] / 2), end_mean = r_round0(lead(p2) + (double)g.cn[p2 - 1] / 2);
  const double start_pm = r_round0((double)g.cn[p1 - 1] / 2), end_pm = r_round0((double)g.cn[p2 - 1] / 2);
  const double lo = (end_mean - end_pm) - (start_mean + start_pm), hi = (end_mean + end_pm) - (start_mean - start_pm);
  const double mean = (lo + hi) / 2;
  r.num[0] = start_mean;
  r.num[1] = end_mean;
  r.num[2] = start_pm;
  r.num[3] = mean;
  r.num[4] = hi - mean;
}

nw_table::Row& add_row(nw_table& T, double ev, const std::string& mut1) {
  nw_table::Row r;
  r.eval = ev;
  r.mut = {mut1};
  r.num.assign(9, kNA);
  T.rows.push_back(std::move(r));
  return T.rows.back();
}

// Phase 1 (host): clutter test, flip, moves, children, and the problem handed to the DP kernels: the x blocks of the
// flipped matrix plus one virtual block per (block, length) pair that a colname repair created.
void prepass_prepare(PrepassOut& P, Grid&& g0in) {
  P.g0 = std::move(g0in);
  P.table.reset(new nw_table);
  nw_table& T = *P.table;
  T.kind = 0;
  T.n_mut = 1;
  if (is_cluttered(P.g0)) {  // explore_mutation_space_v2.R:27-33
    P.cluttered = true;
    add_row(T, 0.0, "ref");
    P.flipped = P.g0;
    finish_table(T);
    return;
  }
  if (P.g0.nx == 1 && P.g0.ny == 1) {  // dfs: ref node only, eval 100 (R/evaltools.R:100-101); no SV opportunities
    P.trivial = true;
    P.flipped = P.g0;
    if (P.flipped.v[0] < 0) {  // flip_bitl_y_if_needed: one row, negative -> -bitl (:68-69)
      P.flipped.v[0] = -P.flipped.v[0];
      P.did_flip = true;
    }
    P.maxdepth = 1;
    add_row(T, 100.0, "ref");
    finish_table(T);
    return;
  }
  P.flipped = flip_if_needed(P.g0, P.did_flip, P.unsure);
  const Grid& g = P.flipped;
  P.moves = find_sv_opportunities(g);
  P.maxdepth = P.moves.size() > 2000 ? 0 : 1;  // reduce_depth_if_needed (helpers.R:318-347), maxdepth_input = 1
  const int n = g.nx;
  Arr root(n);
  for (int i = 0; i < n; ++i) root[i] = Pos{(int16_t)(i + 1), g.cn[i]};
  if (P.maxdepth >= 1)
    for (const Move& m : P.moves) P.cand.push_back(carry_out(root, m));
  P.n_cand = (int64_t)P.cand.size();
  P.lenx = g.cn;
  std::vector<int> vsrc;  // source block of every virtual block
  std::map<std::pair<int, int64_t>, int> vmap;
  P.seqs.assign(P.cand.size() + 1, {});
  auto to_seq = [&](const Arr& a, std::vector<int16_t>& s) {
    s.resize(a.size());
    for (size_t p = 0; p < a.size(); ++p) {
      const int b = std::abs(a[p].blk);
      int id = b;
      if (a[p].len != g.cn[b - 1]) {
        auto it = vmap.find({b, a[p].len});
        if (it == vmap.end()) {
          P.lenx.push_back(a[p].len);
          vsrc.push_back(b);
          it = vmap.emplace(std::make_pair(b, a[p].len), (int)P.lenx.size()).first;
        }
        id = it->second;
      }
      if (id > 32000) rfail(NW_ERR_ARG, "too many blocks");
      s[p] = (int16_t)(a[p].blk < 0 ? -id : id);
    }
  };
  to_seq(root, P.seqs[0]);
  for (size_t k = 0; k < P.cand.size(); ++k) to_seq(P.cand[k], P.seqs[k + 1]);
  P.nxa = (int)P.lenx.size();
  P.aln.resize((size_t)P.nxa * g.ny);
  for (int x = 0; x < P.nxa; ++x) {
    const int src = x < n ? x : vsrc[x - n] - 1;
    for (int y = 0; y < g.ny; ++y) {
      const int64_t t = g.at(y, src);
      P.aln[(size_t)x * g.ny + y] = (int8_t)(t > 0 ? 1 : t < 0 ? -1 : 0);
    }
  }
  const size_t ne = P.cand.size() + 1;
  P.srow.assign(ne, 0);
  P.scol.assign(ne, 0);
  P.ccost.assign(ne, -1);
  P.ctotal.assign(ne, 0);
  P.cscore.assign(ne, 0.0);
}

// Phase 3 (host): replay of dfsutil's sequential part (R/dfs_machinery.R:79-246) over the per-child numbers, then
// sort_new_by_penalised, transform_res_new_to_old, annotate_res_out_with_positions_lens.
void prepass_replay(PrepassOut& P, const nw_region_opts& o) {
  if (P.cluttered || P.trivial) return;
  nw_table& T = *P.table;
  const Grid& g = P.flipped;
  const size_t nc = P.cand.size();
  Arr root(g.nx);
  for (int i = 0; i < g.nx; ++i) root[i] = Pos{(int16_t)(i + 1), g.cn[i]};
  for (size_t k = 1; k <= nc; ++k) {
    const Move& m = P.moves[k - 1];
    T.cands.push_back(nw_candidate{m.p1, m.p2, m.sv, (int32_t)P.cand[k - 1].size(), P.srow[k], P.scol[k], P.ccost[k],
                                   P.ctotal[k], P.cscore[k]});
  }
  double s_ref = P.cscore[0];
  if (P.ccost[0] == -2) s_ref = 1.0;  // R/evaltools.R:259-261 (cannot happen with the full corridor)
  double sum_rn = 0, sum_cn0 = 0;
  for (int64_t t : g.rn) sum_rn += (double)t;
  for (int64_t t : g.cn) sum_cn0 += (double)t;
  const double orig_symm = std::min(sum_rn, sum_cn0) / std::max(sum_rn, sum_cn0);  // calc_symm
  auto est_of = [&](size_t k, double sum_cn) { return (((double)P.srow[k] + (double)P.scol[k]) / (sum_rn + sum_cn)) * 100; };
  double est_highest = est_of(0, sum_cn0);  // forcecalc: est_highest <<- est
  struct Rec {
    int depth;
    int mv;  // -1 = ref
    double eval;
  };
  std::vector<Rec> recs;
  recs.push_back(Rec{0, -1, s_ref});
  struct Seen {
    double eval;
    size_t entry;  // 0 = the reference arrangement, k = child k
  };
  std::unordered_map<DiagKey, Seen, DiagKeyHash> visited;
  std::vector<int64_t> dv, dv2;
  visited.emplace(diag_key(g, root), Seen{s_ref, 0});
  if (s_ref >= 0 && P.maxdepth > 0) {
    for (size_t k = 1; k <= nc; ++k) {
      const Arr& a = P.cand[k - 1];
      const DiagKey key = diag_key(g, a);
      auto it = visited.find(key);
      if (it != visited.end()) {
        diag_values(g, a, dv);  // equal hash: compare the value vectors themselves
        diag_values(g, it->second.entry ? P.cand[it->second.entry - 1] : root, dv2);
        if (dv != dv2) rfail(NW_ERR_COLLISION, "pre-pass: two different diagonal bands with one 128-bit hash");
        if (it->second.eval > 1) recs.push_back(Rec{1, (int)k - 1, it->second.eval});
        continue;
      }
      double s;
      double sum_cn = 0;
      for (const Pos& p : a) sum_cn += (double)p.len;
      if (a.size() == 1) {  // dim(mat)[2] == 1: pos_aln / sum(row.names)
        double pos_aln = 0;
        for (int y = 0; y < g.ny; ++y) {
          const int64_t t = cell(g, a, y, 0);
          if (t > 0) pos_aln += (double)t;
        }
        s = round3((pos_aln / sum_rn) * 100);
      } else {
        const double symmetry = std::min(sum_rn, sum_cn) / std::max(sum_rn, sum_cn);
        if (symmetry < orig_symm * 1 && g.ny > 5) {
          s = 1;
        } else {
          const double est = est_of(k, sum_cn);
          const double thr = g.ny < 10 ? 0.9 : 1.0;
          if (est < est_highest * thr) {
            s = 1;
          } else {
            if (est > est_highest) est_highest = est;
            ++P.n_dp;
            s = P.ccost[k] == -2 ? 1.0 : P.cscore[k];  // unreachable terminal cell: R/evaltools.R:259-261
          }
        }
      }
      if (s > 90) recs.push_back(Rec{1, (int)k - 1, s});
      visited.emplace(key, Seen{s, k});
    }
  }
  const double step_penalty = (o.compression / sum_cn0 * 100) * 1.5;
  std::vector<size_t> ord(recs.size());
  std::vector<double> pen(recs.size());
  for (size_t k = 0; k < recs.size(); ++k) {
    ord[k] = k;
    pen[k] = recs[k].eval - recs[k].depth * step_penalty;
  }
  std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return pen[a] > pen[b]; });
  for (size_t k : ord) {
    const Rec& rc = recs[k];
    if (rc.mv < 0) {
      add_row(T, rc.eval, "ref");
    } else {
      const Move& m = P.moves[rc.mv];
      add_row(T, rc.eval, std::to_string(m.p1) + "_" + std::to_string(m.p2) + "_" + kSvName[m.sv]);
    }
  }
  int count = 1;
  for (size_t q = 0; q < T.rows.size(); ++q) {
    ++count;
    if (count > 2000) break;  // "Stopping to annotate results to save some time."
    const Rec& rc = recs[ord[q]];
    if (rc.mv < 0) continue;
    annotate_row(T.rows[q], P.g0, P.moves[rc.mv].p1, P.moves[rc.mv].p2);
  }
  finish_table(T);
  // the work arrays are not needed any more
  std::vector<Arr>().swap(P.cand);
  std::vector<std::vector<int16_t>>().swap(P.seqs);
  std::vector<int8_t>().swap(P.aln);
}

void fill_prepass_summary(nw_region_summary& S, const PrepassOut& P) {
  S.search_depth = P.maxdepth;
  S.mut_simulated = S.mut_tested = (int64_t)P.table->rows.size();
  S.flipped = P.did_flip;
  S.flip_unsure = P.unsure;
  S.cluttered = P.cluttered;
  S.prepass_candidates = P.n_cand;
  S.prepass_dp = P.n_dp;
  S.gpu_launches = P.launches;
  if (P.cluttered) {
    S.search_depth = 0;
    S.mut_simulated = S.mut_tested = 0;
  }
}

// berzerk guard (R/juliasolver.R:9-30) on the flipped matrix
double guarded_width(const PrepassOut& P, double width, int32_t& reduced) {
  int64_t ndup = 0, longest = std::numeric_limits<int64_t>::min();
  for (const Move& m : P.moves)
    if (m.sv == 1) {
      ++ndup;
      longest = std::max<int64_t>(longest, m.p2 - m.p1);
    }
  const size_t nm = P.moves.size();
  reduced = 0;
  if (nm > 400 || (nm > 200 && longest > 40 && ndup > 40)) {
    width = width / 10;
    reduced = 1;
  }
  return width;
}

// R/wrapper_aln_and_analyse.R:180-190: merge of the pre-pass table T and the formatted solver table J
void merge_tables(std::unique_ptr<nw_table>& T, std::unique_ptr<nw_table>& J) {
  if (J->kind == 2 && J->rows.size() == 1) return;  // all(dim(res_julia) == c(1, 3)): res stays the pre-pass table
  bool has_ref = false;
  for (const auto& r : J->rows) has_ref |= (!r.mut.empty() && r.mut[0] == "ref");
  if (!has_ref) {
    double ref_eval = kNA;
    for (const auto& r : T->rows)
      if (r.mut[0] == "ref") ref_eval = r.eval;
    nw_table::Row r;
    r.eval = ref_eval;
    r.mut.assign(J->n_mut, "");
    r.mut[0] = "ref";
    r.num.assign(J->rows.empty() ? 0 : J->rows[0].num.size(), kNA);
    J->rows.push_back(std::move(r));
    sort_eval_na(*J);
  }
  T = std::move(J);
}

void finish_region(std::unique_ptr<nw_table>& T, nw_region_summary& S) {
  finish_table(*T);
  const nw_region_summary F = T->summ;  // res_ref, res_max, n_res_max, maxres, shape
  S.res_ref = F.res_ref;
  S.res_max = F.res_max;
  S.n_res_max = F.n_res_max;
  S.n_rows = F.n_rows;
  S.n_mut = F.n_mut;
  memcpy(S.maxres, F.maxres, sizeof S.maxres);
  T->summ = S;
}

// wrapper :177 -- does the region go to the solver after the pre-pass?
bool needs_solver(const PrepassOut& P) {
  if (P.cluttered) return false;
  double mx = -INFINITY;
  for (const auto& r : P.table->rows) mx = std::max(mx, r.eval);
  return mx < 99.8;
}

// the formatted solver table J replaces / extends the pre-pass table; format_julia_output overwrites the log counters
// (R/juliasolver.R:63-65, :107-111)
void apply_solver_table(PrepassOut& P, nw_region_summary& S, std::unique_ptr<nw_table>& J, const nw_region_opts& o) {
  S.search_depth = o.depth;
  S.mut_simulated = S.mut_tested = J->kind == 2 ? 0 : (int64_t)J->rows.size();
  merge_tables(P.table, J);
}

std::vector<int8_t> sign_matrix(const Grid& g) {
  std::vector<int8_t> aln((size_t)g.nx * g.ny);
  for (int x = 0; x < g.nx; ++x)
    for (int y = 0; y < g.ny; ++y) {
      const int64_t t = g.at(y, x);
      aln[(size_t)x * g.ny + y] = (int8_t)(t > 0 ? 1 : t < 0 ? -1 : 0);
    }
  return aln;
}

}  // namespace
