// CPU unit-test harness for the host/device-shared logic of nahrwhals_b200/csrc (nw_device.h, nw_band.h,
// nw_host.h).  TEST INFRASTRUCTURE ONLY: it lets `pytest -m "not gpu"` check, without a GPU, the exact integer
// code the kernels execute per item (index maps, fingerprint composition, the register-DP row step and the band
// tables) against the oracle.  It is never loaded by the product and is not a CPU fallback: it cannot run a search.
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "../../nahrwhals_b200/csrc/nw_band.h"
#include "../../nahrwhals_b200/csrc/nw_device.h"
#include "../../nahrwhals_b200/csrc/nw_host.h"

using namespace nw;

template <int W>
static int run_fast(const RegionHost& R, const BandTable& T, const int16_t* parent, int kind, int p, int q, int Lc,
                    int* sumup_out) {
  constexpr int NWORD = (W + 31) / 32;
  int S[W];
  for (int k = 0; k < W; ++k) S[k] = 0;
  int sumup = 0;
  const uint32_t* planes = R.planes();
  const int32_t* upx = R.upx();
  const int32_t* rtrev = R.rtrev();
  for (int i = 1; i <= Lc; ++i) {
    const int c = child_elem(parent, kind, p, q, i);
    const int x = c < 0 ? -c : c;
    const int up = upx[2 * x];
    sumup += up;
    const int r0 = R.NYP - 1 - T.eb[i];
    int sb = T.sb[i];
    const uint32_t* pl = planes + (size_t)(x * 2 + (c < 0 ? 1 : 0)) * R.PWS + (r0 >> 5);
    const int sh = r0 & 31;
    uint32_t m[NWORD];
    for (int t = 0; t < NWORD; ++t) {
      const uint64_t two = (uint64_t)pl[t] | ((uint64_t)pl[t + 1] << 32);
      m[t] = (uint32_t)(two >> sh) & T.mask[(size_t)i * 4 + t];
    }
    const int* rtw = rtrev + (size_t)(r0 & 3) * R.rt_copy_stride + (r0 & ~3);
    while (sb > 2) {
      dp_row_shift1<W>(S);
      --sb;
    }
    if (sb == 0) dp_row_step<W, 0>(S, m, up, rtw);
    else if (sb == 1) dp_row_step<W, 1>(S, m, up, rtw);
    else dp_row_step<W, 2>(S, m, up, rtw);
  }
  *sumup_out = sumup;
  return S[0];
}


// packed form (two candidates of one child length per register, dp_row_step2): the code k_score_pair runs per row
template <int W>
static void run_pair(const RegionHost& R, const BandTable& T, const int16_t* parent, const int* kind, const int* p,
                     const int* q, int Lc, int* s_out, int* sumup_out) {
  constexpr int NWORD = (W + 31) / 32;
  uint32_t S[W];
  for (int k = 0; k < W; ++k) S[k] = 0;
  uint32_t sumup2 = 0;
  const uint32_t* planes = R.planes();
  const int32_t* upx = R.upx();
  const int32_t* rtrev = R.rtrev();
  for (int i = 1; i <= Lc; ++i) {
    uint32_t m[2][NWORD];
    uint32_t up2 = 0;
    const int r0 = R.NYP - 1 - T.eb[i];
    for (int h = 0; h < 2; ++h) {
      const int c = child_elem(parent, kind[h], p[h], q[h], i);
      const int x = c < 0 ? -c : c;
      up2 |= (uint32_t)upx[2 * x] << (16 * h);
      const uint32_t* pl = planes + (size_t)(x * 2 + (c < 0 ? 1 : 0)) * R.PWS + (r0 >> 5);
      const int sh = r0 & 31;
      for (int t = 0; t < NWORD; ++t) {
        const uint64_t two = (uint64_t)pl[t] | ((uint64_t)pl[t + 1] << 32);
        m[h][t] = (uint32_t)(two >> sh) & T.mask[(size_t)i * 4 + t];
      }
    }
    sumup2 += up2;
    uint32_t Mq[2 * NWORD];
    for (int t = 0; t < NWORD; ++t) {
      Mq[2 * t] = (m[0][t] & 0xffffu) | (m[1][t] << 16);
      Mq[2 * t + 1] = (m[0][t] >> 16) | (m[1][t] & 0xffff0000u);
    }
    const int* rtw = rtrev + (size_t)(r0 & 3) * R.rt_copy_stride + (r0 & ~3);
    uint32_t rt2[W];
    for (int k = 0; k < W; ++k) rt2[k] = (uint32_t)rtw[k] * 0x10001u;
    int sb = T.sb[i];
    while (sb > 2) {
      dp_row_shift1<W>(S);
      --sb;
    }
    if (sb == 0) dp_row_step2<W, 0>(S, Mq, up2, rt2);
    else if (sb == 1) dp_row_step2<W, 1>(S, Mq, up2, rt2);
    else dp_row_step2<W, 2>(S, Mq, up2, rt2);
  }
  for (int h = 0; h < 2; ++h) {
    s_out[h] = (int)((S[0] >> (16 * h)) & 0xffffu);
    sumup_out[h] = (int)((sumup2 >> (16 * h)) & 0xffffu);
  }
}

extern "C" {

// Scores child(parent, kind, p, q) with the register-DP code path.  Returns 1 if the fast path applies
// (cost/total in bp written), 0 if the geometry needs the generic kernel, -1 on early exit.
int h_score_fast(const int8_t* aln, int n_x, int n_y, const int64_t* len_x, const int64_t* len_y, const int16_t* parent,
                 int L, int kind, int p, int q, int force, int64_t* cost_bp, int64_t* total_bp, int* window) {
  RegionHost R = build_region(aln, n_x, n_y, len_x, len_y, 4);
  const int Lc = child_len(L, kind, p, q);
  if (!force && early_exit(Lc, n_y)) return -1;
  BandTable T = build_band(Lc, n_y, force, WINDOW_MAX);
  if (!T.fast_ok) return 0;
  static const int wins[] = {8, 12, 16, 20, 24, 28, 32, 36, 40, 44, 48, 52, 56, 60, 64, 72, 80, 96, 112, 128};
  int W = 0;
  for (int w : wins)
    if (w >= T.max_width) {
      W = w;
      break;
    }
  if (!W) return 0;
  int sumup = 0, s = 0;
  switch (W) {
    case 8: s = run_fast<8>(R, T, parent, kind, p, q, Lc, &sumup); break;
    case 12: s = run_fast<12>(R, T, parent, kind, p, q, Lc, &sumup); break;
    case 16: s = run_fast<16>(R, T, parent, kind, p, q, Lc, &sumup); break;
    case 20: s = run_fast<20>(R, T, parent, kind, p, q, Lc, &sumup); break;
    case 24: s = run_fast<24>(R, T, parent, kind, p, q, Lc, &sumup); break;
    case 28: s = run_fast<28>(R, T, parent, kind, p, q, Lc, &sumup); break;
    case 32: s = run_fast<32>(R, T, parent, kind, p, q, Lc, &sumup); break;
    case 36: s = run_fast<36>(R, T, parent, kind, p, q, Lc, &sumup); break;
    case 40: s = run_fast<40>(R, T, parent, kind, p, q, Lc, &sumup); break;
    case 44: s = run_fast<44>(R, T, parent, kind, p, q, Lc, &sumup); break;
    case 48: s = run_fast<48>(R, T, parent, kind, p, q, Lc, &sumup); break;
    case 52: s = run_fast<52>(R, T, parent, kind, p, q, Lc, &sumup); break;
    case 56: s = run_fast<56>(R, T, parent, kind, p, q, Lc, &sumup); break;
    case 60: s = run_fast<60>(R, T, parent, kind, p, q, Lc, &sumup); break;
    case 64: s = run_fast<64>(R, T, parent, kind, p, q, Lc, &sumup); break;
    case 72: s = run_fast<72>(R, T, parent, kind, p, q, Lc, &sumup); break;
    case 80: s = run_fast<80>(R, T, parent, kind, p, q, Lc, &sumup); break;
    case 96: s = run_fast<96>(R, T, parent, kind, p, q, Lc, &sumup); break;
    case 112: s = run_fast<112>(R, T, parent, kind, p, q, Lc, &sumup); break;
    default: s = run_fast<128>(R, T, parent, kind, p, q, Lc, &sumup); break;
  }
  const int64_t total = (int64_t)sumup + R.sum_rt;
  *cost_bp = (total - s) * R.gcd;
  *total_bp = total * R.gcd;
  *window = W;
  return 1;
}


// Scores child(parent, move A) and child(parent, move B) -- two children of ONE length -- with the packed row step.
// Returns 1 if the packed path applies (cost / total of both in bp written), 0 if it does not (geometry needs the
// generic kernel, unequal lengths, or a total outside the 16-bit range), -1 on early exit.
int h_score_pair(const int8_t* aln, int n_x, int n_y, const int64_t* len_x, const int64_t* len_y, const int16_t* parent,
                 int L, const int* kind, const int* p, const int* q, int force, int64_t* cost_bp, int64_t* total_bp) {
  RegionHost R = build_region(aln, n_x, n_y, len_x, len_y, 4);
  const int Lc = child_len(L, kind[0], p[0], q[0]);
  if (Lc != child_len(L, kind[1], p[1], q[1])) return 0;
  if (!force && early_exit(Lc, n_y)) return -1;
  BandTable T = build_band(Lc, n_y, force, WINDOW_MAX);
  if (!T.fast_ok) return 0;
  if (!pair_range_ok(R, Lc)) return 0;
  static const int wins[] = {8, 12, 16, 20, 24, 28, 32, 36, 40, 44, 48, 52, 56, 60, 64, 72, 80, 96, 112, 128};
  int W = 0;
  for (int w : wins)
    if (w >= T.max_width) {
      W = w;
      break;
    }
  if (!W) return 0;
  int s[2], sumup[2];
  switch (W) {
    case 8: run_pair<8>(R, T, parent, kind, p, q, Lc, s, sumup); break;
    case 12: run_pair<12>(R, T, parent, kind, p, q, Lc, s, sumup); break;
    case 16: run_pair<16>(R, T, parent, kind, p, q, Lc, s, sumup); break;
    case 20: run_pair<20>(R, T, parent, kind, p, q, Lc, s, sumup); break;
    case 24: run_pair<24>(R, T, parent, kind, p, q, Lc, s, sumup); break;
    case 28: run_pair<28>(R, T, parent, kind, p, q, Lc, s, sumup); break;
    case 32: run_pair<32>(R, T, parent, kind, p, q, Lc, s, sumup); break;
    case 36: run_pair<36>(R, T, parent, kind, p, q, Lc, s, sumup); break;
    case 40: run_pair<40>(R, T, parent, kind, p, q, Lc, s, sumup); break;
    case 44: run_pair<44>(R, T, parent, kind, p, q, Lc, s, sumup); break;
    case 48: run_pair<48>(R, T, parent, kind, p, q, Lc, s, sumup); break;
    case 52: run_pair<52>(R, T, parent, kind, p, q, Lc, s, sumup); break;
    case 56: run_pair<56>(R, T, parent, kind, p, q, Lc, s, sumup); break;
    case 60: run_pair<60>(R, T, parent, kind, p, q, Lc, s, sumup); break;
    case 64: run_pair<64>(R, T, parent, kind, p, q, Lc, s, sumup); break;
    case 72: run_pair<72>(R, T, parent, kind, p, q, Lc, s, sumup); break;
    case 80: run_pair<80>(R, T, parent, kind, p, q, Lc, s, sumup); break;
    case 96: run_pair<96>(R, T, parent, kind, p, q, Lc, s, sumup); break;
    case 112: run_pair<112>(R, T, parent, kind, p, q, Lc, s, sumup); break;
    default: run_pair<128>(R, T, parent, kind, p, q, Lc, s, sumup); break;
  }
  for (int h = 0; h < 2; ++h) {
    const int64_t total = (int64_t)sumup[h] + R.sum_rt;
    cost_bp[h] = (total - s[h]) * R.gcd;
    total_bp[h] = total * R.gcd;
  }
  return 1;
}

// fold_window_classes (nw_host.h): n classes (cls[k], slots[k]) -> out[k] = class cls[k] is folded into (itself if not)
void h_fold_classes(int n, const int* cls, const uint64_t* slots, uint64_t wave_ctas, int* out) {
  std::map<int, uint64_t> m;
  for (int k = 0; k < n; ++k) m[cls[k]] += slots[k];
  const std::map<int, int> into = fold_window_classes(m, wave_ctas);
  for (int k = 0; k < n; ++k) {
    int c = cls[k];
    while (into.count(c)) c = into.at(c);
    out[k] = c;
  }
}

// 16-bit additions of the packed row step that wrapped since the library was loaded (must stay 0: pair_range_ok)
long h_pair_wraps() { return pair_wraps; }

// k3 (score*1000) from integer cost/total, the product's scoring arithmetic
int h_score_k3(int64_t cost, int64_t total) { return score_k3(cost, total); }

// materialises the child through the index map; returns its length
int h_child(const int16_t* parent, int L, int kind, int p, int q, int16_t* out) {
  const int Lc = child_len(L, kind, p, q);
  for (int t = 1; t <= Lc; ++t) out[t - 1] = (int16_t)child_elem(parent, kind, p, q, t);
  return Lc;
}

// 1 iff the O(1) composed fingerprint of child(parent, move) equals the direct fingerprint of the materialised child
int h_fp_check(const int16_t* parent, int L, int kind, int p, int q) {
  static std::vector<U4> pw = build_pow_table();
  const int LS = round_ls(L);
  std::vector<unsigned char> blob(blob_stride(LS));
  fill_blob(blob.data(), LS, parent, L, pw.data(), 1);
  const U4* F = reinterpret_cast<const U4*>(blob.data() + (size_t)LS * 2);
  const U4* G = F + (LS + 1);
  const U4 a = child_fingerprint(F, G, pw.data(), L, kind, p, q);
  std::vector<int16_t> ch(2 * L + 2);
  const int Lc = h_child(parent, L, kind, p, q, ch.data());
  const U4 b = seq_fingerprint(ch.data(), Lc, pw.data(), 1);
  return memcmp(&a, &b, sizeof(U4)) == 0;
}
void h_fp(const int16_t* seq, int L, uint32_t* out4) {
  static std::vector<U4> pw = build_pow_table();
  const U4 f = seq_fingerprint(seq, L, pw.data(), 1);
  memcpy(out4, &f, 16);
}

// band geometry: cells, fast_ok, max_width
void h_band(int L, int n_y, int force, int64_t* cells, int* fast_ok, int* max_width, int* contiguous) {
  BandTable T = build_band(L, n_y, force, WINDOW_MAX);
  *cells = T.cells;
  *fast_ok = T.fast_ok;
  *max_width = T.max_width;
  *contiguous = T.contiguous;
}

// 1 iff the O(L + n_y) pointer construction of the corridor rows equals the exhaustive evaluation of every cell
int h_band_check(int L, int n_y, int force) {
  std::vector<BandRowHost> r1, r2;
  int64_t c1 = 0, c2 = 0;
  bool k1 = true, k2 = true;
  band_rows_bruteforce(L, n_y, force, r1, c1, k1);
  band_rows_fast(L, n_y, corridor_pct(L, force, n_y), r2, c2, k2);
  if (c1 != c2 || k1 != k2) return 0;
  if (!k1) return 1;  // non-contiguous rows go to the generic kernel, which uses only (a, b) of contiguous tables
  for (int i = 1; i <= L; ++i)
    if (r1[i].a != r2[i].a || r1[i].b != r2[i].b) return 0;
  return 1;
}

// Julia-style Float32 score string of the product's TSV writer
int h_score_string(int k3, char* buf, int cap) {
  std::string s = score_string(k3);
  snprintf(buf, cap, "%s", s.c_str());
  return (int)s.size();
}

}  // extern "C"
