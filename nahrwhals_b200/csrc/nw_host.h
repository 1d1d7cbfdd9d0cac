// nw_host.h -- host-side preparation shared by the driver and the CPU unit-test harness (pure C++17, no CUDA):
// packed region context (bit-planes, unit lengths), fingerprint power tables, root blob, TSV I/O.
// Reference: inst/extdata/scripts/scripts_nw_main/solver.jl (read_tsv :101-107, read_length_tsv :121-138).
#pragma once
#include <stdint.h>

#include <cmath>
#include <map>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "nw_device.h"

namespace nw {

constexpr int LMAX = FP_EMAX - 2;  // longest index (child) the search supports
constexpr int WINDOW_MAX = 128;    // widest register window compiled in nw_score_fast.cu
constexpr int RT_PAD = 160;        // zero padding of rt_rev beyond NYP (>= WINDOW_MAX + 4)

// Host cores this PROCESS may count on for its short parallel phases (report enumeration, region prepare / replay /
// format, PAF pre / post): the machine's hardware threads divided by the ranks a launcher started on this node
// (torchrun / mpirun export LOCAL_WORLD_SIZE / OMPI_COMM_WORLD_LOCAL_SIZE), NW_HOST_THREADS overrides.  Eight ranks with
// several contexts each would otherwise start hundreds of threads on a 32-core host.
inline int host_cores() {
  static const int n = [] {
    if (const char* e = getenv("NW_HOST_THREADS")) return std::max(1, atoi(e));
    unsigned hw = std::thread::hardware_concurrency();
    int cores = hw ? (int)hw : 4;
    for (const char* name : {"LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "SLURM_NTASKS_PER_NODE"})
      if (const char* e = getenv(name)) {
        const int w = atoi(e);
        if (w > 1) cores = std::max(1, cores / w);
        break;
      }
    return cores;
  }();
  return n;
}

struct HostError : std::runtime_error {
  int code;
  HostError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

struct RegionHost {
  int n_x = 0, n_y = 0, NYP = 0, PW = 0, PWS = 0, MW = 0;
  int64_t gcd = 1;
  int64_t sum_rt = 0, sum_up = 0;  // units
  int64_t max_up = 0;              // longest x block in units
  std::vector<unsigned char> ctx;  // [planes | upx (per plane row: index x*2+strand) | rt_rev x 4 shifted copies]
  uint32_t off_upx = 0, off_rtrev = 0;
  int rt_copy_stride = 0;
  std::vector<int32_t> rt, cumr;  // 1-based
  uint32_t* planes() { return reinterpret_cast<uint32_t*>(ctx.data()); }
  const uint32_t* planes() const { return reinterpret_cast<const uint32_t*>(ctx.data()); }
  const int32_t* upx() const { return reinterpret_cast<const int32_t*>(ctx.data() + off_upx); }
  const int32_t* rtrev() const { return reinterpret_cast<const int32_t*>(ctx.data() + off_rtrev); }
};

// aln: n_x*n_y int8, x-major.  Lengths in bp; all DP arithmetic is done in units of gcd(lengths) so that
// int32 suffices; dividing numerator and denominator of cost/total by the same integer leaves the correctly
// rounded IEEE quotient of solver.jl:374 unchanged.
inline RegionHost build_region(const int8_t* aln, int n_x, int n_y, const int64_t* len_x, const int64_t* len_y,
                               int maxdepth) {
  if (n_x < 1 || n_y < 1) throw HostError(-1, "empty gridmatrix");
  if (n_x > 32000 || n_y > 32000) throw HostError(-1, "gridmatrix too large");
  RegionHost R;
  R.n_x = n_x;
  R.n_y = n_y;
  int64_t g = 0;
  for (int x = 0; x < n_x; ++x) {
    if (len_x[x] <= 0) throw HostError(-5, "non-positive x block length");
    g = std::gcd(g, len_x[x]);
  }
  for (int y = 0; y < n_y; ++y) {
    if (len_y[y] <= 0) throw HostError(-5, "non-positive y block length");
    g = std::gcd(g, len_y[y]);
  }
  R.gcd = g;
  for (int x = 0; x < n_x; ++x) {
    R.sum_up += len_x[x] / g;
    R.max_up = std::max<int64_t>(R.max_up, len_x[x] / g);
  }
  for (int y = 0; y < n_y; ++y) R.sum_rt += len_y[y] / g;
  // a child can at most double per level (a duplication copies a segment): bound the int32 DP range
  const double worst = (double)R.sum_up * std::pow(2.0, std::max(0, maxdepth)) + (double)R.sum_rt;
  if (worst >= (double)(1 << 29)) throw HostError(-5, "block lengths exceed the int32 DP range");
  R.NYP = ((n_y + 1) + 31) / 32 * 32;
  R.PW = R.NYP / 32;
  R.PWS = R.PW + WINDOW_MAX / 32 + 1;
  R.MW = (n_x + 1 + 31) / 32;
  const size_t planes_bytes = (((size_t)(n_x + 1) * 2 * R.PWS * 4) + 15) & ~(size_t)15;
  const size_t upx_bytes = (((size_t)(n_x + 1) * 2 * 4) + 15) & ~(size_t)15;
  R.rt_copy_stride = (R.NYP + RT_PAD + 3) & ~3;
  const size_t rt_bytes = (size_t)R.rt_copy_stride * 4 * 4;
  R.off_upx = (uint32_t)planes_bytes;
  R.off_rtrev = (uint32_t)(planes_bytes + upx_bytes);
  R.ctx.assign(planes_bytes + upx_bytes + rt_bytes, 0);
  uint32_t* pl = R.planes();
  for (int x = 1; x <= n_x; ++x)
    for (int j = 1; j <= n_y; ++j) {
      const int a = aln[(size_t)(x - 1) * n_y + (j - 1)];
      if (a == 0) continue;
      const int r = R.NYP - 1 - j;
      pl[(size_t)(x * 2 + (a < 0 ? 1 : 0)) * R.PWS + (r >> 5)] |= 1u << (r & 31);
    }
  int32_t* upx = reinterpret_cast<int32_t*>(R.ctx.data() + R.off_upx);
  for (int x = 1; x <= n_x; ++x) upx[2 * x] = upx[2 * x + 1] = (int32_t)(len_x[x - 1] / g);  // same index as the plane row
  int32_t* rr = reinterpret_cast<int32_t*>(R.ctx.data() + R.off_rtrev);
  R.rt.assign(n_y + 1, 0);
  R.cumr.assign(n_y + 1, 0);
  for (int j = 1; j <= n_y; ++j) {
    R.rt[j] = (int32_t)(len_y[j - 1] / g);
    R.cumr[j] = R.cumr[j - 1] + R.rt[j];
  }
  for (int c = 0; c < 4; ++c)
    for (int t = 0; t < R.rt_copy_stride; ++t) {
      const int r = t + c;  // copy c holds rt_rev[t + c]
      const int j = R.NYP - 1 - r;
      rr[(size_t)c * R.rt_copy_stride + t] = (j >= 1 && j <= n_y) ? R.rt[j] : 0;
    }
  return R;
}

// Packed 16-bit DP (dp_row_step2 / k_score_pair): a child of Lc blocks has total <= sum_rt + Lc * max_up; when that
// bound fits 15 bits no half-register of the packed row step can overflow (nw_device.h).
inline bool pair_range_ok(const RegionHost& R, int Lc) {
  return R.sum_rt + (int64_t)Lc * R.max_up <= (int64_t)PAIR_TOTAL_MAX;
}

// Folding of DP window classes (run_scoring).  `slots_of`: kernel class -> work slots of a level, class = window W, with
// CLS_PAIR set for the packed kernel.  A class is folded into the next wider class that is present (same kernel kind)
// when walking the wider window costs less than a launch of its own: n CTAs pay (B - A) / A more each, a separate
// launch ends in a partly filled wave and costs at least the latency of one thread's DP.  Rule: fold iff
// n * (B - A) / A < wave_ctas / 2.  Folds chain (A into B, then B with A's slots into C).  Returns class -> target.
constexpr int CLS_PAIR = 0x1000;
inline std::map<int, int> fold_window_classes(std::map<int, uint64_t> slots_of, uint64_t wave_ctas) {
  std::map<int, int> into;
  for (auto it = slots_of.begin(); it != slots_of.end(); ++it) {
    auto nx = std::next(it);
    if (nx == slots_of.end()) break;
    if (it->first == 0 || ((it->first ^ nx->first) & CLS_PAIR) != 0) continue;
    const uint64_t wa = (uint64_t)(it->first & ~CLS_PAIR), wb = (uint64_t)(nx->first & ~CLS_PAIR);
    if (((it->second + 127) / 128) * (wb - wa) * 2 < wave_ctas * wa) {
      into[it->first] = nx->first;
      nx->second += it->second;
    }
  }
  return into;
}

// ---- fingerprint power tables: pw[e + FP_EMAX] = B_k^e mod p for four fixed (seeded) bases ----
inline uint32_t powmod(uint32_t b, uint64_t e) {
  uint32_t r = 1;
  while (e) {
    if (e & 1) r = mulmod(r, b);
    b = mulmod(b, b);
    e >>= 1;
  }
  return r;
}
inline std::vector<U4> build_pow_table(uint64_t reseed = 0) {
  std::vector<U4> pw(2 * FP_EMAX + 1);
  // splitmix64, fixed seed -> reproducible fingerprints; `reseed` > 0 picks other bases (a search whose dedup verifier
  // found two different indices under one key is repeated with new bases)
  uint64_t s = 0x9E3779B97F4A7C15ull + reseed * 0xD1B54A32D192ED03ull;
  for (int k = 0; k < 4; ++k) {
    uint32_t B;
    do {
      s += 0x9E3779B97F4A7C15ull;
      uint64_t z = s;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z ^= z >> 31;
      B = (uint32_t)(z % P31);
    } while (B < (1u << 20));
    const uint32_t Binv = powmod(B, (uint64_t)P31 - 2);
    uint32_t f = 1, b = 1;
    for (int e = 0; e <= FP_EMAX; ++e) {
      pw[FP_EMAX + e].v[k] = f;
      pw[FP_EMAX - e].v[k] = b;
      f = mulmod(f, B);
      b = mulmod(b, Binv);
    }
  }
  return pw;
}

// parent blob: [seq int16 x LS | F U4 x (LS+1) | G U4 x (LS+1)]
inline uint32_t blob_stride(int LS) { return (uint32_t)LS * 2 + 2u * (uint32_t)(LS + 1) * 16u; }
inline int round_ls(int L) { return (L + 7) & ~7; }
inline void fill_blob(unsigned char* out, int LS, const int16_t* seq, int L, const U4* pw, uint32_t rid1) {
  int16_t* s = reinterpret_cast<int16_t*>(out);
  for (int t = 0; t < LS; ++t) s[t] = t < L ? seq[t] : 0;
  U4* F = reinterpret_cast<U4*>(out + (size_t)LS * 2);
  U4* G = F + (LS + 1);
  for (int k = 0; k < 4; ++k) {
    F[0].v[k] = rid1;  // F'[n] = rid + B * F[n]
    G[0].v[k] = 0;     // G'[n] = B * G[n]
    for (int t = 0; t < L; ++t) {
      F[t + 1].v[k] = addmod(F[t].v[k], mulmod(elem_val(seq[t]), pw[FP_EMAX + t + 1].v[k]));
      G[t + 1].v[k] = addmod(G[t].v[k], mulmod(elem_val(seq[t]), pw[FP_EMAX + 1 - t].v[k]));
    }
    for (int t = L + 1; t <= LS; ++t) {
      F[t].v[k] = F[L].v[k];
      G[t].v[k] = G[L].v[k];
    }
  }
}

// ---- TSV input (solver.jl:101-107, :121-138; values may be written like "1e+05") ----
inline bool parse_fields(const std::string& line, std::vector<double>& out) {
  out.clear();
  size_t pos = 0;
  while (pos <= line.size()) {
    size_t tab = line.find('\t', pos);
    if (tab == std::string::npos) tab = line.size();
    const std::string f = line.substr(pos, tab - pos);
    char* end = nullptr;
    const double v = strtod(f.c_str(), &end);
    if (end == f.c_str()) return false;
    while (*end == ' ') ++end;
    if (*end != '\0') return false;
    out.push_back(v);
    pos = tab + 1;
  }
  return true;
}
inline std::vector<std::string> read_lines(const char* path, bool& ok) {
  std::vector<std::string> lines;
  FILE* f = fopen(path, "rb");
  ok = f != nullptr;
  if (!f) return lines;
  std::string cur;
  int ch;
  while ((ch = fgetc(f)) != EOF) {
    if (ch == '\n') {
      if (!cur.empty() && cur.back() == '\r') cur.pop_back();
      lines.push_back(cur);
      cur.clear();
    } else {
      cur.push_back((char)ch);
    }
  }
  if (!cur.empty()) lines.push_back(cur);
  fclose(f);
  return lines;
}
struct ProblemHost {
  int n_x = 0, n_y = 0;
  std::vector<int8_t> aln;  // x-major
  std::vector<int64_t> len_x, len_y;
};
inline int64_t integral_length(double v) {
  if (!(v > 0) || v != std::floor(v) || v > 9.0e15) throw HostError(-5, "block length is not a positive integer");
  return (int64_t)v;
}
inline ProblemHost read_problem(const char* inmat, const char* inlens) {
  ProblemHost P;
  bool ok;
  std::vector<std::string> ml = read_lines(inmat, ok);
  if (!ok) throw HostError(-2, std::string("cannot open ") + inmat);
  std::vector<std::vector<double>> rows;
  std::vector<double> vals;
  for (auto& l : ml) {
    if (l.empty()) continue;
    if (!parse_fields(l, vals)) throw HostError(-3, std::string("malformed number in ") + inmat);
    rows.push_back(vals);
  }
  if (rows.empty()) throw HostError(-3, std::string("empty gridmatrix ") + inmat);
  P.n_y = (int)rows.size();
  P.n_x = (int)rows[0].size();
  for (auto& r : rows)
    if ((int)r.size() != P.n_x) throw HostError(-3, std::string("ragged gridmatrix ") + inmat);
  P.aln.assign((size_t)P.n_x * P.n_y, 0);
  for (int y = 0; y < P.n_y; ++y)
    for (int x = 0; x < P.n_x; ++x) {
      const double v = rows[y][x];
      if (std::isnan(v)) throw HostError(-3, "NaN in gridmatrix");
      P.aln[(size_t)x * P.n_y + y] = (int8_t)((v > 0) - (v < 0));  // :103 sign, :106 transpose
    }
  std::vector<std::string> ll = read_lines(inlens, ok);
  if (!ok) throw HostError(-2, std::string("cannot open ") + inlens);
  if (ll.size() != 2) throw HostError(-3, "The TSV file must have exactly two rows.");  // :128-130
  std::vector<double> ly, lx;
  if (!parse_fields(ll[0], ly) || !parse_fields(ll[1], lx)) throw HostError(-3, "malformed number in lengths file");
  if ((int)ly.size() != P.n_y || (int)lx.size() != P.n_x)
    throw HostError(-3, "lengths file does not match the gridmatrix shape");
  for (double v : lx) P.len_x.push_back(integral_length(v));
  for (double v : ly) P.len_y.push_back(integral_length(v));
  return P;
}

// "<Float32 score>" as Julia prints it (solver.jl:723).  The score is k3/1000 with k3 an integer in
// [0, 100000]; its shortest round-trip Float32 repr is the 3-decimal string with trailing zeros trimmed
// (>= 1 decimal kept) -- checked for all 100001 values against a generic shortest-repr printer in tests.
inline std::string score_string(int k3) {
  char buf[32];
  snprintf(buf, sizeof buf, "%d.%03d", k3 / 1000, k3 % 1000);
  std::string s(buf);
  while (s.back() == '0' && s[s.size() - 2] != '.') s.pop_back();
  return s;
}

}  // namespace nw
