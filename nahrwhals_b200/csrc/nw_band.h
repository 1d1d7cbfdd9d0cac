// nw_band.h -- corridor geometry of slow_score, evaluated on the host in IEEE double exactly as written in
// solver.jl:378-386 (determine_corridor_pct), :390 / :400 (first column / first row) and :412-413 (interior):
//     first column  i>=2 : valid iff !(i/row > pct)
//     first row     j>=2 : valid iff !(j/col > pct)
//     interior           : valid iff !(i/row > j/col + pct) && !(j/col > i/row + pct)
// No integer cross-multiplication: the predicates are evaluated with the same divisions and additions as Julia.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace nw {

struct BandRowHost {
  int a, b;  // valid real columns [a, b], 1-based; a > b == empty row
};

// Corridor modes: NW_BAND_JULIA = slow_score (pct chosen by the child length, solver.jl:378-386);
// NW_BAND_FORCE = the R twin's forcecalc branch (pct 1.0, R/evaltools.R:203-205);
// NW_BAND_RTWIN = the R twin without forcecalc: pct chosen by dim(mat)[1] == n_y (R/evaltools.R:193-201; R's matrix
// has y as rows).  The predicates themselves are symmetric in (i/row, j/col), so the R cell set is the transpose of
// the Julia one for the same pct.
enum { NW_BAND_JULIA = 0, NW_BAND_FORCE = 1, NW_BAND_RTWIN = 2 };

struct BandTable {
  int L = 0, n_y = 0;
  int mode = NW_BAND_JULIA;
  bool contiguous = true;        // every row's valid set is one interval
  std::vector<BandRowHost> row;  // row[1..L]; row[0] unused
  int64_t cells = 0;             // sum of valid cells == SURVEY 8d "cells(L, n_y)"
  // --- derived data for the register-resident kernel (extended with virtual row 0 / column 0) ---
  bool fast_ok = false;
  int max_width = 0;             // max over rows of eb - ea + 1
  std::vector<int16_t> eb;       // eb[0..L]
  std::vector<int16_t> ea;       // ea[0..L]
  std::vector<uint8_t> sb;       // sb[1..L] = eb[i] - eb[i-1]
  std::vector<uint32_t> mask;    // mask[i*4 + w]: diag-allowed bits of row i, bit k <-> column eb[i]-k (k < 128)
};

inline double corridor_pct(int row, int mode, int n_y = 0) {
  if (mode == NW_BAND_FORCE) return 1.0;  // R/evaltools.R:203-205
  if (mode == NW_BAND_RTWIN) row = n_y;   // R/evaltools.R:193: dim_[1] is the y dimension
  if (row < 10) return 1.0;
  if (row < 30) return 0.3;
  return 0.2;
}

inline bool band_valid(int i, int j, int row, int col, double pct) {
  if (i == 1 && j == 1) return true;
  if (j == 1) return !((double)i / (double)row > pct);
  if (i == 1) return !((double)j / (double)col > pct);
  const double im = (double)i / (double)row, jm = (double)j / (double)col;
  return !(im > (jm + pct)) && !(jm > (im + pct));
}

// exhaustive form (tests only): every (i, j) is evaluated with band_valid()
inline void band_rows_bruteforce(int L, int n_y, int mode, std::vector<BandRowHost>& row, int64_t& cells,
                                 bool& contiguous) {
  const double pct = corridor_pct(L, mode, n_y);
  row.assign(L + 1, BandRowHost{1, 0});
  cells = 0;
  contiguous = true;
  for (int i = 1; i <= L; ++i) {
    int a = 0, b = -1, cnt = 0;
    for (int j = 1; j <= n_y; ++j)
      if (band_valid(i, j, L, n_y, pct)) {
        if (cnt == 0) a = j;
        b = j;
        ++cnt;
      }
    if (cnt) {
      row[i] = BandRowHost{a, b};
      if (b - a + 1 != cnt) contiguous = false;
    }
    cells += cnt;
  }
}

// Row intervals in O(L + n_y) predicate evaluations.  For a fixed row, A(j) = (i/row > j/col + pct) is non-increasing
// in j and B(j) = (j/col > i/row + pct) non-decreasing (IEEE division and addition are monotone), so the interior
// valid set is the interval [lo_i, hi_i]; both ends are non-decreasing in i and are tracked with forward-moving
// pointers that evaluate exactly the Float64 expressions of solver.jl:412-413 (no algebraic rewriting).
inline void band_rows_fast(int L, int n_y, double pct, std::vector<BandRowHost>& row, int64_t& cells,
                           bool& contiguous) {
  row.assign(L + 1, BandRowHost{1, 0});
  cells = 0;
  contiguous = true;
  {  // first row (solver.jl:400): columns 1..b with !(j/col > pct)
    int b = 1;
    while (b + 1 <= n_y && !((double)(b + 1) / (double)n_y > pct)) ++b;
    if (L >= 1) {
      row[1] = BandRowHost{1, b};
      cells += b;
    }
  }
  int lo = 2, hi = 1;
  for (int i = 2; i <= L; ++i) {
    const double im = (double)i / (double)L;
    while (lo <= n_y && (im > ((double)lo / (double)n_y + pct))) ++lo;                 // A(lo) still true
    while (hi + 1 <= n_y && !((double)(hi + 1) / (double)n_y > (im + pct))) ++hi;      // B(hi+1) false
    const bool col1 = !(im > pct);                                                     // solver.jl:390
    const bool interior = lo <= hi && lo <= n_y && hi >= 2;
    int a = 1, b = 0;
    if (interior) {
      a = col1 ? 1 : lo;
      b = hi;
      if (col1 && lo > 2) contiguous = false;
      cells += (hi - lo + 1) + (col1 ? 1 : 0);
    } else if (col1) {
      a = b = 1;
      cells += 1;
    }
    row[i] = BandRowHost{a, b};
  }
}

inline BandTable build_band(int L, int n_y, int mode, int max_window = 128) {
  BandTable T;
  T.L = L;
  T.n_y = n_y;
  T.mode = mode;
  const double pct = corridor_pct(L, mode, n_y);
  band_rows_fast(L, n_y, pct, T.row, T.cells, T.contiguous);
  // ---- fast-path derivation ----
  T.eb.assign(L + 1, 0);
  T.ea.assign(L + 1, 0);
  T.sb.assign(L + 1, 0);
  T.mask.assign((size_t)(L + 1) * 4, 0u);
  bool ok = T.contiguous && L >= 1 && n_y >= 1;
  for (int i = 1; i <= L && ok; ++i)
    if (T.row[i].a > T.row[i].b) ok = false;
  if (ok && T.row[L].b != n_y) ok = false;
  if (ok) {
    T.ea[0] = 0;
    T.eb[0] = (int16_t)T.row[1].b;
    for (int i = 1; i <= L; ++i) {
      T.ea[i] = (int16_t)(T.row[i].a == 1 ? 0 : T.row[i].a);
      T.eb[i] = (int16_t)T.row[i].b;
    }
    for (int i = 1; i <= L && ok; ++i) {
      const int s = T.eb[i] - T.eb[i - 1];
      if (s < 0 || s > 255) ok = false;
      if (T.ea[i] < T.ea[i - 1]) ok = false;
      if (T.row[i].a > T.eb[i - 1]) ok = false;  // the new row must overlap the old one (entering cells are fed by right moves)
      T.sb[i] = (uint8_t)(s < 0 ? 0 : s);
      const int w = T.eb[i] - T.ea[i] + 1;
      if (w > T.max_width) T.max_width = w;
    }
    if (T.eb[0] - T.ea[0] + 1 > T.max_width) T.max_width = T.eb[0] - T.ea[0] + 1;
    if (T.max_width > max_window) ok = false;
  }
  if (ok) {
    for (int i = 1; i <= L; ++i) {
      // diag move into (i,j) needs (i-1, j-1) valid in the extended grid: j-1 in [ea[i-1], eb[i-1]]
      const int j0 = std::max(T.row[i].a, (int)T.ea[i - 1] + 1), j1 = std::min(T.row[i].b, (int)T.eb[i - 1] + 1);
      if (j0 > j1) continue;
      const int k0 = T.eb[i] - j1, k1 = T.eb[i] - j0;  // bits k0..k1 (k1 < max_window <= 128)
      for (int w = k0 >> 5; w <= (k1 >> 5); ++w) {
        const int lo = std::max(k0 - w * 32, 0), hi = std::min(k1 - w * 32, 31);
        const uint32_t bits = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
        T.mask[(size_t)i * 4 + w] |= bits;
      }
    }
  }
  T.fast_ok = ok;
  return T;
}

}  // namespace nw
