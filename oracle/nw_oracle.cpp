// nw_oracle.cpp -- CPU restatement of NAHRwhals' Julia mutation-space solver.
//
// TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA path: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
// Nothing under nahrwhals_b200/ links, imports or executes it.
//
// PARITY PIN STATUS: "parity unpinned" for tie order.  The reference ships no tests, no stored
// res.tsv and no gridmatrix, and neither julia nor R exist in this image, so this restatement is
// pinned only by (a) the two README known answers (README.md:135-139, reproduced in
// tests/test_oracle.py from the grid recovered from inst/extdata/output_png/...-6.png) and (b) an
// independent pure-Python restatement (oracle/nw_oracle_py.py) that must agree on every id.
// The one deliberately *canonicalised* behaviour is the iteration order of Julia's Dict in
// get_good_trajectories (solver.jl:731): we iterate in discovery-id order.
//
// Every function cites the solver.jl lines it follows
// (inst/extdata/scripts/scripts_nw_main/solver.jl in the reference tree).
//
// Build: see oracle/Makefile  (g++ -O2 -std=c++17 -shared -fPIC).

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

using Index = std::vector<int16_t>;  // solver.jl:81  Vector{Int16}

enum MoveKind : int { MOVE_DUP = 0, MOVE_DEL = 1, MOVE_INV = 2 };  // solver.jl:34
static const char* kMoveName[3] = {"move_dup", "move_del", "move_inv"};

struct MoveTaken {  // solver.jl:44-48
  int kind;
  int64_t start, stop;
};

struct IndexHash {
  size_t operator()(const Index& v) const noexcept {
    uint64_t h = 0xcbf29ce484222325ull ^ (uint64_t)v.size();
    for (int16_t e : v) {
      h ^= (uint16_t)e;
      h *= 0x100000001b3ull;
    }
    return (size_t)(h ^ (h >> 29));
  }
};

struct Problem {
  int n_x = 0, n_y = 0;
  std::vector<int8_t> aln;    // aln[x*n_y + y]  (solver.jl:106: transpose(sign(raw)) -> n_x x n_y)
  std::vector<double> len_y;  // lengths[1]  (solver.jl:343 walk_right_cost)
  std::vector<double> len_x;  // lengths[2]  (solver.jl:344 climb_up_cost source)
};

struct Opts {
  int maxdepth = 3;
  int maxdup = 3;
  int64_t maxwidth = -1;  // -1 == nothing
  double minreport = 0.95;
  int64_t maxreport = -1;  // -1 == nothing
  int64_t max_scored = -1; // bench only: stop the search after this many slow_score calls (bounded CPU sample)
};

struct ScoreOut {
  double score;  // Float64 result of slow_score (may be -Inf)
  double cost;   // cost_res[row,col] (integer valued or +Inf); NaN if early exit
  double total;  // sum of climb_up + walk_right
  int64_t cells; // DP cells evaluated (0 on early exit)
};

// round(x, digits=3) as Base._round_invstep: round(x*1000)/1000, ties-to-even, x if result not finite.
static double julia_round3(double x) {
  double y = std::nearbyint(x * 1000.0) / 1000.0;  // default FE_TONEAREST == RoundNearest
  if (!std::isfinite(y)) return x;
  return y;
}

// solver.jl:378-386
static double determine_corridor_pct(int row) {
  if (row < 10) return 1.0;
  if (row < 30) return 0.3;
  return 0.2;
}

struct ScoreScratch {
  std::vector<double> cost_res, cost_d, up, cumsum_u, cumsum_r;
};

// solver.jl:329-376 (slow_score) with fill_first_column! :388-396, fill_first_row! :398-406,
// fill_matrix! :408-426.  Matrices are stored column-major like Julia (i + j*row, 0-based).
// `force` reproduces the R twin's forcecalc branch (R/evaltools.R:203-205: corridor 1.0, no early exit), the
// path that yields res_ref (R/wrapper_aln_and_analyse.R:176).  Julia's own scorer always has force == false.
// force: 0 = slow_score as Julia runs it; 1 = the R twin's forcecalc branch (corridor 1.0, no early exit);
// 2 = the R twin without forcecalc (R/evaltools.R:193-201: no early exit, corridor chosen by dim(mat)[1], which is the
// y dimension of R's matrix; the predicates are symmetric in (i/row, j/col), so the transposed matrix Julia-style
// gives the same cell set).
static ScoreOut slow_score(const Problem& P, const int16_t* index, int L, ScoreScratch& ws, int force = 0) {
  const double inf = std::numeric_limits<double>::infinity();
  ScoreOut out{0.0, std::nan(""), 0.0, 0};
  const int ny = P.n_y;
  // :332-334 easy exit
  if (!force && ny > 10 && (double)std::abs(L - ny) > (double)ny * 0.5) {
    out.score = 0.0;
    return out;
  }
  const int row = L, col = ny;
  ws.cost_res.assign((size_t)row * col, 0.0);  // :356
  ws.cost_d.resize((size_t)row * col);
  ws.up.resize(row);
  ws.cumsum_u.resize(row);
  ws.cumsum_r.resize(col);
  auto C = [&](int i, int j) -> double& { return ws.cost_res[(size_t)(i - 1) + (size_t)(j - 1) * row]; };
  auto D = [&](int i, int j) -> double& { return ws.cost_d[(size_t)(i - 1) + (size_t)(j - 1) * row]; };
  const std::vector<double>& rt = P.len_y;  // :343
  double su = 0.0;
  for (int i = 1; i <= row; ++i) {
    int idx = index[i - 1];
    int ax = idx < 0 ? -idx : idx;
    ws.up[i - 1] = P.len_x[ax - 1];  // :344
    su += ws.up[i - 1];
    ws.cumsum_u[i - 1] = su;  // :349
    for (int j = 1; j <= col; ++j) {
      int a = P.aln[(size_t)(ax - 1) * ny + (j - 1)];
      int v = idx < 0 ? -a : a;  // :337 selected_rows
      // :352-354  cost_d = -aln; >=0 -> Inf ; <0 -> 0   i.e. 0 iff aln_to_use > 0
      D(i, j) = (v > 0) ? 0.0 : inf;
    }
  }
  double sr = 0.0;
  for (int j = 1; j <= col; ++j) {
    sr += rt[j - 1];
    ws.cumsum_r[j - 1] = sr;  // :350
  }
  const double pct = force == 1 ? 1.0 : determine_corridor_pct(force == 2 ? col : row);  // :359
  int64_t cells = 1;
  C(1, 1) = std::min(ws.up[0] + rt[0], D(1, 1));  // :362
  for (int i = 2; i <= row; ++i) {                // :388-396
    if ((double)i / (double)row > pct) {
      C(i, 1) = inf;
    } else {
      C(i, 1) = std::min(ws.up[i - 1] + C(i - 1, 1), ws.cumsum_u[i - 2] + D(i, 1));
      ++cells;
    }
  }
  for (int j = 2; j <= col; ++j) {  // :398-406
    if ((double)j / (double)col > pct) {
      C(1, j) = inf;
    } else {
      C(1, j) = std::min(rt[j - 1] + C(1, j - 1), ws.cumsum_r[j - 2] + D(1, j));
      ++cells;
    }
  }
  for (int j = 2; j <= col; ++j)  // :368
    for (int i = 2; i <= row; ++i) C(i, j) = inf;
  // :408-426  findall is column-major: j outer, i inner.
  for (int j = 2; j <= col; ++j) {
    const double jm = (double)j / (double)col;
    for (int i = 2; i <= row; ++i) {
      const double im = (double)i / (double)row;
      const int middle = 1 - ((im > (jm + pct) ? 1 : 0) + (jm > (im + pct) ? 1 : 0));  // :412-413
      if (middle != 1) continue;
      double a = C(i - 1, j - 1) + D(i, j);
      double b = C(i - 1, j) + ws.up[i - 1];
      double c = C(i, j - 1) + rt[j - 1];
      C(i, j) = std::min(a, std::min(b, c));  // :420-422
      ++cells;
    }
  }
  const double total = su + sr;  // :372
  out.cost = C(row, col);
  out.total = total;
  out.cells = cells;
  out.score = julia_round3((1.0 - out.cost / total) * 100.0);  // :374
  return out;
}

struct Moveset {  // solver.jl:11-14
  int n = 0;
  std::vector<uint8_t> dd, inv;
};

// solver.jl:217-231 (to_moveset) with compute_possible_moves :159-165
static Moveset to_moveset(const Problem& P) {
  Moveset m;
  m.n = P.n_x;
  m.dd.assign((size_t)P.n_x * P.n_x, 0);
  m.inv.assign((size_t)P.n_x * P.n_x, 0);
  for (int i = 0; i < P.n_x; ++i)
    for (int j = i + 1; j < P.n_x; ++j) {
      bool dd = false, iv = false;
      for (int y = 0; y < P.n_y; ++y) {
        int prod = (int)P.aln[(size_t)i * P.n_y + y] * (int)P.aln[(size_t)j * P.n_y + y];
        if (prod == 1) dd = true;
        if (prod == -1) iv = true;
      }
      m.dd[(size_t)i * P.n_x + j] = m.dd[(size_t)j * P.n_x + i] = dd;
      m.inv[(size_t)i * P.n_x + j] = m.inv[(size_t)j * P.n_x + i] = iv;
    }
  return m;
}

// solver.jl:244-246 / 259-261 / 276-278 (1-based positions p<q)
static Index apply_duplication(const Index& d, int p, int q) {
  Index r(d.begin(), d.begin() + q);
  r.insert(r.end(), d.begin() + p, d.end());
  return r;
}
static Index apply_deletion(const Index& d, int p, int q) {
  Index r(d.begin(), d.begin() + (p - 1));
  r.insert(r.end(), d.begin() + (q - 1), d.end());
  return r;
}
static Index apply_inversion(const Index& d, int p, int q) {
  Index r(d.begin(), d.begin() + p);
  for (int k = q - 1; k >= p + 1; --k) r.push_back((int16_t)(-d[k - 1]));
  r.insert(r.end(), d.begin() + (q - 1), d.end());
  return r;
}

// solver.jl:292-298
static int duplication_count(const Index& idx, int max_size) {
  std::vector<int64_t> values(max_size, 0);
  for (int16_t e : idx) values[(e < 0 ? -e : e) - 1] += 1;
  return (int)*std::max_element(values.begin(), values.end());
}

struct IndexMap {  // solver.jl:80-87 ; ids are 1-based, root == 1
  std::unordered_map<Index, int64_t, IndexHash> map;
  std::vector<Index> by_id;                  // by_id[id-1] (the reference's `indices` misses the root; unused there)
  std::vector<std::vector<int64_t>> parents; // [id-1]
  std::vector<float> scores;                 // Float32 (solver.jl:84)
  std::vector<std::vector<MoveTaken>> moves; // [id-1]
  std::vector<double> cost, total;           // diagnostics for parity tests: DP terminal cost / total
  std::vector<int> level;                    // discovery level (0 root)
  int64_t current = 0;
};

struct Stats {
  std::vector<int64_t> generated, scored, cells, frontier;  // per level (level 1..maxdepth)
  std::vector<double> level_seconds;
  double seconds = 0.0;
};

struct Search {
  const Problem& P;
  const Opts& O;
  Moveset ms;
  IndexMap im;
  ScoreScratch ws;
  Stats st;
  int cur_level = 0;
  int64_t scored_total = 0;
  bool stop = false;
  explicit Search(const Problem& p, const Opts& o) : P(p), O(o) { ms = to_moveset(p); }

  void record(const Index& index, double score, const ScoreOut& so) {
    im.map.emplace(index, im.current + 1);
    im.by_id.push_back(index);
    im.scores.push_back((float)score);
    im.cost.push_back(so.cost);
    im.total.push_back(so.total);
    im.level.push_back(cur_level);
    im.current += 1;
  }

  // solver.jl:532-564
  double update_index(std::vector<Index>& indices, Index&& index, int64_t parent_id, MoveTaken move,
                      double threshold, double best) {
    st.generated.back() += 1;
    auto it = im.map.find(index);
    if (it == im.map.end()) {
      ScoreOut so = slow_score(P, index.data(), (int)index.size(), ws);
      st.scored.back() += 1;
      if (O.max_scored >= 0 && ++scored_total >= O.max_scored) stop = true;
      st.cells.back() += so.cells;
      double score = so.score;
      if (score > best) best = score;
      record(index, score, so);
      im.parents.push_back({parent_id});
      im.moves.push_back({move});
      if (score >= threshold) indices.push_back(std::move(index));
    } else {
      int64_t pos = it->second;
      im.parents[pos - 1].push_back(parent_id);
      im.moves[pos - 1].push_back(move);
    }
    return best;
  }

  // solver.jl:589-642
  double all_moves_scored(const Index& index, double best, std::vector<Index>& indices, double suboptimality) {
    auto it = im.map.find(index);
    if (it == im.map.end()) {  // :598-604 root insert (score is Float32 start_score)
      ScoreOut so = slow_score(P, index.data(), (int)index.size(), ws);
      record(index, so.score, so);
      im.parents.push_back({});
      im.moves.push_back({});
      it = im.map.find(index);
    }
    const int64_t parent_id = it->second;
    const double current_best = best;
    const int L = (int)index.size();
    for (int i = 1; i <= L && !stop; ++i) {
      for (int j = i + 1; j <= L && !stop; ++j) {
        // :180-190 retrieve_possible_moves
        int i1 = index[i - 1], i2 = index[j - 1];
        bool flip = (i1 * i2) < 0;
        int a1 = std::abs(i1) - 1, a2 = std::abs(i2) - 1;
        bool dup_del = ms.dd[(size_t)a1 * ms.n + a2];
        bool inv = ms.inv[(size_t)a1 * ms.n + a2];
        bool pm_dd = (flip && inv) || (!flip && dup_del);
        bool pm_inv = (flip && dup_del) || (!flip && inv);
        if (pm_dd) {
          if (duplication_count(index, P.n_x) <= O.maxdup) {  // :611
            best = update_index(indices, apply_duplication(index, i, j), parent_id, {MOVE_DUP, i, j},
                                suboptimality * current_best, best);
          }
          best = update_index(indices, apply_deletion(index, i, j), parent_id, {MOVE_DEL, i, j},
                              suboptimality * current_best, best);
        }
        if (pm_inv) {
          best = update_index(indices, apply_inversion(index, i, j), parent_id, {MOVE_INV, i, j},
                              suboptimality * current_best, best);
        }
      }
    }
    return best;
  }

  // solver.jl:644-686
  double run(int max_depth) {
    auto t0 = std::chrono::steady_clock::now();
    Index root(P.n_x);
    for (int i = 0; i < P.n_x; ++i) root[i] = (int16_t)(i + 1);
    std::vector<Index> indices;
    double best = 0.0;
    auto open_level = [&](size_t nfront) {
      ++cur_level;
      st.generated.push_back(0);
      st.scored.push_back(0);
      st.cells.push_back(0);
      st.frontier.push_back((int64_t)nfront);
    };
    auto tl = std::chrono::steady_clock::now();
    cur_level = 0;
    // root is recorded at level 0 inside all_moves_scored, children at level 1
    {
      ScoreOut so = slow_score(P, root.data(), (int)root.size(), ws);
      record(root, so.score, so);
      im.parents.push_back({});
      im.moves.push_back({});
    }
    open_level(1);
    best = all_moves_scored(root, best, indices, 0.0);
    auto now = std::chrono::steady_clock::now();
    st.level_seconds.push_back(std::chrono::duration<double>(now - tl).count());
    for (int depth = 1; depth <= max_depth - 1; ++depth) {
      tl = std::chrono::steady_clock::now();
      const double subopt = 0.0;  // :668 suboptimality === nothing from main()
      if (O.maxwidth >= 0) {      // :669-672
        std::stable_sort(indices.begin(), indices.end(), [&](const Index& a, const Index& b) {
          float ka = -im.scores[im.map.find(a)->second - 1];
          float kb = -im.scores[im.map.find(b)->second - 1];
          return ka < kb;
        });
        if ((int64_t)indices.size() > O.maxwidth) indices.resize((size_t)O.maxwidth);
      }
      std::vector<Index> new_indices;
      open_level(indices.size());
      for (const Index& cur : indices) {
        if (stop) break;
        best = all_moves_scored(cur, best, new_indices, subopt);
      }
      indices.swap(new_indices);
      now = std::chrono::steady_clock::now();
      st.level_seconds.push_back(std::chrono::duration<double>(now - tl).count());
    }
    st.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return best;
  }
};

struct Trajectory {  // solver.jl:713-716
  float score;
  std::vector<MoveTaken> moves;
  int64_t id;  // index id it leads to (diagnostic)
};

// solver.jl:688-705.  Returns (result, valid)
static bool concatenate_moves_aux(const IndexMap& im, int64_t cur, int depth, int maxdepth,
                                  std::vector<std::vector<MoveTaken>>& result) {
  result.clear();
  if (cur == 1) {
    result.emplace_back();
    return true;
  } else if (depth > maxdepth) {
    result.emplace_back();
    return false;
  }
  const auto& ps = im.parents[cur - 1];
  const auto& mv = im.moves[cur - 1];
  std::vector<std::vector<MoveTaken>> traj;
  for (size_t e = 0; e < ps.size(); ++e) {
    bool valid = concatenate_moves_aux(im, ps[e], depth + 1, maxdepth, traj);
    if (valid) {
      for (auto& t : traj) {
        std::vector<MoveTaken> nt(t);
        nt.push_back(mv[e]);
        result.push_back(std::move(nt));
      }
    }
  }
  return true;
}

// solver.jl:727-745.  Dict iteration order is canonicalised to ascending id (see header).
static std::vector<Trajectory> get_good_trajectories(const IndexMap& im, double best, const Opts& O) {
  std::vector<Trajectory> trajs;
  std::vector<std::vector<MoveTaken>> res;
  for (int64_t id = 1; id <= im.current; ++id) {
    float score = im.scores[id - 1];
    if ((double)score >= O.minreport * best) {
      concatenate_moves_aux(im, id, 0, O.maxdepth, res);
      for (auto& t : res) trajs.push_back(Trajectory{score, t, id});
    }
  }
  std::stable_sort(trajs.begin(), trajs.end(), [](const Trajectory& a, const Trajectory& b) {
    float ka = -a.score, kb = -b.score;
    if (ka < kb) return true;
    if (kb < ka) return false;
    return a.moves.size() < b.moves.size();
  });
  if (O.maxreport >= 0 && (int64_t)trajs.size() > O.maxreport) trajs.resize((size_t)O.maxreport);
  return trajs;
}

// Julia print(::Float32): shortest round-trip digits, fixed notation in our range, >=1 decimal.
static std::string julia_float32_string(float v) {
  if (std::isnan(v)) return "NaN";
  if (std::isinf(v)) return v > 0 ? "Inf" : "-Inf";
  if (v == 0.0f) return std::signbit(v) ? "-0.0" : "0.0";
  char buf[64];
  int prec = 1;
  for (; prec <= 9; ++prec) {
    snprintf(buf, sizeof buf, "%.*e", prec - 1, (double)v);
    if (strtof(buf, nullptr) == v) break;
  }
  // buf = d.ddddde[+-]XX
  std::string s(buf);
  size_t epos = s.find('e');
  int ex = atoi(s.c_str() + epos + 1);
  std::string mant = s.substr(0, epos);
  bool neg = mant[0] == '-';
  if (neg) mant = mant.substr(1);
  std::string digits;
  for (char c : mant)
    if (c != '.') digits.push_back(c);
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  std::string out;
  if (ex >= -5 && ex < 6) {  // Julia uses fixed notation for 1e-5 <= |x| < 1e6 (Float32)
    if (ex >= 0) {
      std::string ip, fp;
      for (int k = 0; k <= ex; ++k) ip.push_back(k < (int)digits.size() ? digits[k] : '0');
      if ((int)digits.size() > ex + 1) fp = digits.substr(ex + 1);
      if (fp.empty()) fp = "0";
      out = ip + "." + fp;
    } else {
      out = "0.";
      for (int k = 0; k < -ex - 1; ++k) out.push_back('0');
      out += digits;
    }
  } else {
    out = digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : std::string("0")) + "e" +
          std::to_string(ex);
  }
  return neg ? "-" + out : out;
}

// solver.jl:718-725
static std::string trajectory_to_string(const Trajectory& t) {
  std::string s = julia_float32_string(t.score) + "\t" + std::to_string(t.moves.size()) + "\t";
  for (size_t k = 0; k < t.moves.size(); ++k) {
    if (k) s += "\t";
    s += std::string(kMoveName[t.moves[k].kind]) + "\t" + std::to_string(t.moves[k].start) + "\t" +
         std::to_string(t.moves[k].stop);
  }
  return s;
}

static bool split_parse_line(const std::string& line, std::vector<double>& out) {
  out.clear();
  size_t pos = 0;
  while (pos <= line.size()) {
    size_t tab = line.find('\t', pos);
    if (tab == std::string::npos) tab = line.size();
    std::string f = line.substr(pos, tab - pos);
    char* end = nullptr;
    double v = strtod(f.c_str(), &end);
    if (end == f.c_str()) return false;
    out.push_back(v);
    pos = tab + 1;
  }
  return true;
}

// solver.jl:101-107 (read_tsv; the log-length second output is dead when the lens file exists)
static bool read_tsv(const char* path, Problem& P) {
  std::ifstream f(path);
  if (!f) return false;
  std::string line;
  std::vector<std::vector<double>> rows;
  std::vector<double> vals;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty()) continue;
    if (!split_parse_line(line, vals)) return false;
    rows.push_back(vals);
  }
  if (rows.empty()) return false;
  P.n_y = (int)rows.size();
  P.n_x = (int)rows[0].size();
  for (auto& r : rows)
    if ((int)r.size() != P.n_x) return false;
  P.aln.assign((size_t)P.n_x * P.n_y, 0);
  for (int y = 0; y < P.n_y; ++y)
    for (int x = 0; x < P.n_x; ++x) {
      double v = rows[y][x];
      P.aln[(size_t)x * P.n_y + y] = (int8_t)((v > 0) - (v < 0));
    }
  return true;
}

// solver.jl:121-138
static bool read_length_tsv(const char* path, Problem& P) {
  std::ifstream f(path);
  if (!f) return false;
  std::vector<std::string> lines;
  std::string line;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    lines.push_back(line);
  }
  if (lines.size() != 2) return false;  // :128-130
  if (!split_parse_line(lines[0], P.len_y)) return false;
  if (!split_parse_line(lines[1], P.len_x)) return false;
  return (int)P.len_y.size() == P.n_y && (int)P.len_x.size() == P.n_x;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C API (ctypes) -- used by tests/ and bench.py only.
// ------------------------------------------------------------------------------------------------
extern "C" {

struct nwo_opts {
  int32_t maxdepth, maxdup;
  int64_t maxwidth;  // -1 none
  double minreport;
  int64_t maxreport;  // -1 none
  int64_t max_scored; // -1 none (bench-only bounded sample)
};

struct nwo_result {
  Problem P;
  Opts O;
  IndexMap im;
  Stats st;
  double best = 0.0;
  std::vector<Trajectory> trajs;
  std::string tsv;
  std::vector<int16_t> flat_index;  // concatenated indices by id
  std::vector<int64_t> index_off;
  std::vector<int64_t> edge_off, edge_parent;
  std::vector<int32_t> edge_move;  // kind,start,stop triples
};

static void fill_problem(Problem& P, const int8_t* aln, int n_x, int n_y, const double* len_x, const double* len_y) {
  P.n_x = n_x;
  P.n_y = n_y;
  P.aln.assign(aln, aln + (size_t)n_x * n_y);
  P.len_x.assign(len_x, len_x + n_x);
  P.len_y.assign(len_y, len_y + n_y);
}

static nwo_result* solve_impl(Problem&& P, const nwo_opts* o, int want_trajs) {
  nwo_result* R = new nwo_result();
  R->P = std::move(P);
  R->O.maxdepth = o->maxdepth;
  R->O.maxdup = o->maxdup;
  R->O.maxwidth = o->maxwidth;
  R->O.minreport = o->minreport;
  R->O.maxreport = o->maxreport;
  R->O.max_scored = o->max_scored;
  Search S(R->P, R->O);
  R->best = S.run(R->O.maxdepth);
  R->im = std::move(S.im);
  R->st = std::move(S.st);
  if (want_trajs) {
    R->trajs = get_good_trajectories(R->im, R->best, R->O);
    for (auto& t : R->trajs) R->tsv += trajectory_to_string(t) + "\n";
  }
  return R;
}

// Full search on an in-memory problem.  aln is n_x*n_y (x-major), values in {-1,0,1}.
nwo_result* nwo_solve(const int8_t* aln, int n_x, int n_y, const double* len_x, const double* len_y,
                      const nwo_opts* o, int want_trajs) {
  Problem P;
  fill_problem(P, aln, n_x, n_y, len_x, len_y);
  return solve_impl(std::move(P), o, want_trajs);
}

// solver.jl main() :755-812 minus the JIT warm-up search (:798).  Returns 0 on success.
int nwo_solve_files(const char* inmat, const char* inlens, const char* out, const nwo_opts* o) {
  Problem P;
  if (!read_tsv(inmat, P)) return -1;
  if (!read_length_tsv(inlens, P)) return -2;
  nwo_result* R = solve_impl(std::move(P), o, 1);
  std::string tmp = std::string(out) + ".tmp";
  FILE* f = fopen(tmp.c_str(), "w");
  if (!f) {
    delete R;
    return -3;
  }
  fwrite(R->tsv.data(), 1, R->tsv.size(), f);
  fclose(f);
  rename(tmp.c_str(), out);
  delete R;
  return 0;
}

void nwo_free(nwo_result* R) { delete R; }
int64_t nwo_n_ids(const nwo_result* R) { return R->im.current; }
double nwo_best(const nwo_result* R) { return R->best; }
double nwo_seconds(const nwo_result* R) { return R->st.seconds; }
int nwo_n_levels(const nwo_result* R) { return (int)R->st.generated.size(); }
void nwo_level_stats(const nwo_result* R, int64_t* generated, int64_t* scored, int64_t* cells, int64_t* frontier,
                     double* seconds) {
  for (size_t l = 0; l < R->st.generated.size(); ++l) {
    generated[l] = R->st.generated[l];
    scored[l] = R->st.scored[l];
    cells[l] = R->st.cells[l];
    frontier[l] = R->st.frontier[l];
    seconds[l] = R->st.level_seconds[l];
  }
}
// per-id arrays (id-1 indexed): Float32 score, DP cost (NaN = early exit, Inf = unreachable), total, level, length
void nwo_id_arrays(const nwo_result* R, float* score, double* cost, double* total, int32_t* level, int32_t* length) {
  for (int64_t k = 0; k < R->im.current; ++k) {
    score[k] = R->im.scores[k];
    cost[k] = R->im.cost[k];
    total[k] = R->im.total[k];
    level[k] = R->im.level[k];
    length[k] = (int32_t)R->im.by_id[k].size();
  }
}
int64_t nwo_index_total_len(const nwo_result* R) {
  int64_t n = 0;
  for (auto& v : R->im.by_id) n += (int64_t)v.size();
  return n;
}
void nwo_indices(const nwo_result* R, int16_t* flat, int64_t* off) {
  int64_t n = 0;
  for (int64_t k = 0; k < R->im.current; ++k) {
    off[k] = n;
    for (int16_t e : R->im.by_id[k]) flat[n++] = e;
  }
  off[R->im.current] = n;
}
int64_t nwo_n_edges(const nwo_result* R) {
  int64_t n = 0;
  for (auto& v : R->im.parents) n += (int64_t)v.size();
  return n;
}
// edges in (child id ascending, insertion order): off[id-1]..off[id]; parent id; move = kind,start,stop
void nwo_edges(const nwo_result* R, int64_t* off, int64_t* parent, int32_t* move3) {
  int64_t n = 0;
  for (int64_t k = 0; k < R->im.current; ++k) {
    off[k] = n;
    for (size_t e = 0; e < R->im.parents[k].size(); ++e) {
      parent[n] = R->im.parents[k][e];
      move3[3 * n + 0] = R->im.moves[k][e].kind;
      move3[3 * n + 1] = (int32_t)R->im.moves[k][e].start;
      move3[3 * n + 2] = (int32_t)R->im.moves[k][e].stop;
      ++n;
    }
  }
  off[R->im.current] = n;
}
int64_t nwo_n_trajs(const nwo_result* R) { return (int64_t)R->trajs.size(); }
const char* nwo_tsv(const nwo_result* R) { return R->tsv.c_str(); }
int64_t nwo_tsv_len(const nwo_result* R) { return (int64_t)R->tsv.size(); }

// Score raw candidate index vectors (flat + offsets).  Outputs per candidate: Float64 score, cost, total, cells.
void nwo_score_batch(const int8_t* aln, int n_x, int n_y, const double* len_x, const double* len_y,
                     const int16_t* flat, const int64_t* off, int64_t n, int force, double* score, double* cost,
                     double* total, int64_t* cells) {
  Problem P;
  fill_problem(P, aln, n_x, n_y, len_x, len_y);
  ScoreScratch ws;
  for (int64_t k = 0; k < n; ++k) {
    ScoreOut so = slow_score(P, flat + off[k], (int)(off[k + 1] - off[k]), ws, force);
    score[k] = so.score;
    cost[k] = so.cost;
    total[k] = so.total;
    cells[k] = so.cells;
  }
}

// Float32 -> Julia print string (for tests of the TSV writer)
int nwo_f32_string(float v, char* buf, int cap) {
  std::string s = julia_float32_string(v);
  snprintf(buf, cap, "%s", s.c_str());
  return (int)s.size();
}

}  // extern "C"

#ifdef NWO_MAIN
// CLI with solver.jl's argument table (:756-791).
int main(int argc, char** argv) {
  nwo_opts o{3, 3, -1, 0.95, -1, -1};
  std::vector<const char*> pos;
  for (int a = 1; a < argc; ++a) {
    std::string s = argv[a];
    auto need = [&](const char* nm) -> const char* {
      if (a + 1 >= argc) {
        fprintf(stderr, "missing value for %s\n", nm);
        exit(2);
      }
      return argv[++a];
    };
    if (s == "-d" || s == "--maxdepth") o.maxdepth = atoi(need("-d"));
    else if (s == "-u" || s == "--maxdup") o.maxdup = atoi(need("-u"));
    else if (s == "-w" || s == "--maxwidth") o.maxwidth = (int64_t)strtod(need("-w"), nullptr);
    else if (s == "-O" || s == "--maxoffset") (void)need("-O");
    else if (s == "-r" || s == "--minreport") o.minreport = strtod(need("-r"), nullptr);
    else if (s == "-R" || s == "--maxreport") o.maxreport = (int64_t)strtod(need("-R"), nullptr);
    else pos.push_back(argv[a]);
  }
  if (pos.size() != 3) {
    fprintf(stderr, "usage: nw_oracle [-d D] [-u U] [-w W] [-O O] [-r r] [-R R] inmat lens out\n");
    return 2;
  }
  int rc = nwo_solve_files(pos[0], pos[1], pos[2], &o);
  if (rc != 0) fprintf(stderr, "nw_oracle: error %d\n", rc);
  return rc == 0 ? 0 : 1;
}
#endif
