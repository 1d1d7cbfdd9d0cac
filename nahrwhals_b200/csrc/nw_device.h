// nw_device.h -- per-item logic shared by the CUDA kernels and (via NW_HD) host unit tests.
//
// Everything here is integer arithmetic restating pieces of the reference Julia solver
// (inst/extdata/scripts/scripts_nw_main/solver.jl, cited as solver.jl:LINE):
//   * the child index of a move as an *index map* over its parent (apply_* :244-278), never materialised;
//   * a 4 x 31-bit polynomial fingerprint of that child computed in O(1) from the parent's prefix tables
//     (stands in for the exact Dict{Vector{Int16}} key of :81/:535, collision bound in DESIGN.md);
//   * the corridor-banded min-plus DP of slow_score (:329-426) in its "savings" form
//       S[i][j] = max(S[i-1][j], S[i][j-1], match(i,j) ? S[i-1][j-1] + up_i + rt_j : -inf),  cost = total - S,
//     which is algebraically identical on integers (derivation in DESIGN.md) and makes two of the three moves free.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define NW_HD __host__ __device__ __forceinline__
#else
#define NW_HD inline
#endif

namespace nw {

// ---------------------------------------------------------------- candidate descriptors
// desc = parent_slot:32 | i:15 | j:15 | kind:2   (i<j are 1-based positions in the parent, solver.jl:607-608)
enum : int { KIND_DUP = 0, KIND_DEL = 1, KIND_INV = 2, KIND_ID = 3 };

NW_HD uint64_t pack_desc(uint32_t parent_slot, int i, int j, int kind) {
  return ((uint64_t)parent_slot << 32) | ((uint64_t)(uint32_t)i << 17) | ((uint64_t)(uint32_t)j << 2) |
         (uint64_t)(uint32_t)kind;
}
NW_HD uint32_t desc_parent(uint64_t d) { return (uint32_t)(d >> 32); }
NW_HD int desc_i(uint64_t d) { return (int)((d >> 17) & 0x7fff); }
NW_HD int desc_j(uint64_t d) { return (int)((d >> 2) & 0x7fff); }
NW_HD int desc_kind(uint64_t d) { return (int)(d & 3); }

// length of the child (apply_duplication :245, apply_deletion :260, apply_inversion :277)
NW_HD int child_len(int L, int kind, int p, int q) {
  return kind == KIND_DUP ? L + (q - p) : (kind == KIND_DEL ? L - (q - p) : L);
}

// element t (1-based) of the child, read through the index map from the parent sequence d[0..L-1]
NW_HD int child_elem(const int16_t* d, int kind, int p, int q, int t) {
  if (kind == KIND_DUP) return t <= q ? d[t - 1] : d[t - q + p - 1];
  if (kind == KIND_DEL) return t <= p - 1 ? d[t - 1] : d[t + (q - p) - 1];
  if (kind == KIND_INV) return (t <= p || t >= q) ? d[t - 1] : -d[p + q - t - 1];
  return d[t - 1];
}

// ---------------------------------------------------------------- fingerprints (mod p = 2^31-1, four bases)
constexpr uint32_t P31 = 0x7fffffffu;
constexpr int FP_EMAX = 4096;  // exponents in [-FP_EMAX, FP_EMAX]; sequences are capped at FP_EMAX-2 elements

NW_HD uint32_t addmod(uint32_t a, uint32_t b) {
  uint32_t r = a + b;
  return r >= P31 ? r - P31 : r;
}
NW_HD uint32_t submod(uint32_t a, uint32_t b) {
  uint32_t r = a + P31 - b;
  return r >= P31 ? r - P31 : r;
}
NW_HD uint32_t mulmod(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a * (uint64_t)b;
  uint32_t r = (uint32_t)(t & P31) + (uint32_t)(t >> 31);
  r = (r & P31) + (r >> 31);
  return r >= P31 ? r - P31 : r;
}
NW_HD uint32_t elem_val(int x) { return x > 0 ? (uint32_t)x : P31 - (uint32_t)(-x); }

struct U4 {
  uint32_t v[4];
};

// Prefix tables of a parent d (0-based t), already "finalised" with the region id as leading element:
//   F'[n] = val(rid+1) + B * sum_{t<n} val(d_t) B^t ,   G'[n] = B * sum_{t<n} val(d_t) B^-t ,   pw[e + FP_EMAX] = B^e.
// The child's fingerprint  H = val(rid+1) + B * sum_t val(c_t) B^t  is composed from differences of these tables in
// O(1) (the rid term cancels in every difference and survives exactly once through the leading F' term).
NW_HD U4 child_fingerprint(const U4* F, const U4* G, const U4* pw, int L, int kind, int p, int q) {
  U4 out;
  const U4 FLv = F[L];
  if (kind == KIND_DUP) {  // d[0..q-1] ++ d[p..L-1] placed at q
    const U4 a = F[q], b = F[p], w = pw[FP_EMAX + (q - p)];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) out.v[k] = addmod(a.v[k], mulmod(submod(FLv.v[k], b.v[k]), w.v[k]));
  } else if (kind == KIND_DEL) {  // d[0..p-2] ++ d[q-1..L-1] placed at p-1
    const U4 a = F[p - 1], b = F[q - 1], w = pw[FP_EMAX - (q - p)];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) out.v[k] = addmod(a.v[k], mulmod(submod(FLv.v[k], b.v[k]), w.v[k]));
  } else if (kind == KIND_INV) {  // d[0..p-1] ++ -rev(d[p..q-2]) ++ d[q-1..L-1]
    const U4 a = F[p], b = F[q - 1], g1 = G[q - 1], g0 = G[p], w = pw[FP_EMAX + (p + q - 2)];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) {
      const uint32_t mid = mulmod(submod(g1.v[k], g0.v[k]), w.v[k]);
      out.v[k] = addmod(submod(a.v[k], mid), submod(FLv.v[k], b.v[k]));
    }
  } else {
    out = FLv;
  }
  return out;
}

// direct fingerprint of a materialised sequence (root insertion; unit tests of the composition identities)
NW_HD U4 seq_fingerprint(const int16_t* s, int L, const U4* pw, uint32_t rid1) {
  U4 out;
  for (int k = 0; k < 4; ++k) {
    uint32_t h = 0;
    for (int t = 0; t < L; ++t) h = addmod(h, mulmod(elem_val(s[t]), pw[FP_EMAX + t].v[k]));
    out.v[k] = addmod(rid1, mulmod(h, pw[FP_EMAX + 1].v[k]));
  }
  return out;
}

NW_HD uint64_t fp_key0(const U4& f) { return (uint64_t)f.v[0] | ((uint64_t)f.v[1] << 32); }
NW_HD uint64_t fp_key1(const U4& f) { return (uint64_t)f.v[2] | ((uint64_t)f.v[3] << 32); }
// the 32 key bits stored next to the position in the second table word: base 2 (31 bits) and the top bit of base 3
NW_HD uint32_t fp_key_hi32(const U4& f) { return f.v[2] | ((f.v[3] & 0x40000000u) << 1); }
NW_HD uint64_t fp_slot_hash(uint64_t k0, uint64_t k1) {
  uint64_t h = k0 * 0x9E3779B97F4A7C15ull ^ (k1 + 0xD6E8FEB86659FD93ull);
  h ^= h >> 32;
  h *= 0xD6E8FEB86659FD93ull;
  h ^= h >> 29;
  return h;
}

// ---------------------------------------------------------------- scoring arithmetic
constexpr int DP_INF = 1 << 30;     // "+Inf" of the cost form (generic kernel)
constexpr int DP_NEG = -(1 << 30);  // "-Inf" of the savings form (register kernel)
constexpr int K3_NEGINF = -1;       // score == -Inf (unreachable terminal cell, solver.jl:374 with cost Inf)

// score*1000 as an exact integer:  k3 = round((1 - cost/total)*100 * 1000)  (solver.jl:374; Base.round digits=3
// is round(x*1000)/1000 with ties-to-even).  IEEE double ops, no FMA contraction possible (explicit _rn on device).
NW_HD int score_k3(int64_t cost_units, int64_t total_units) {
#if defined(__CUDA_ARCH__)
  double q = __ddiv_rn((double)cost_units, (double)total_units);
  double x = __dmul_rn(__dsub_rn(1.0, q), 100.0);
  return (int)__double2ll_rn(__dmul_rn(x, 1000.0));
#else
  volatile double q = (double)cost_units / (double)total_units;
  volatile double x1 = 1.0 - q;
  volatile double x = x1 * 100.0;
  volatile double y = x * 1000.0;
  return (int)__builtin_nearbyint(y);
#endif
}
// the Float64 score the reference would hold: round(...)/1000
NW_HD double k3_to_score(int k3) {
#if defined(__CUDA_ARCH__)
  return __ddiv_rn((double)k3, 1000.0);
#else
  return (double)k3 / 1000.0;
#endif
}

// early exit of slow_score (solver.jl:332-334): n_y > 10 && abs(L - n_y) > n_y*0.5  (exact in integers: 2|L-n_y| > n_y)
NW_HD bool early_exit(int L, int n_y) {
  int d = L - n_y;
  if (d < 0) d = -d;
  return n_y > 10 && 2 * d > n_y;
}

// ---------------------------------------------------------------- register-resident DP row step
// S[k] holds the savings of column (eb_i - k) of the current row (window right-aligned at the row's last valid
// column eb_i).  SB = eb_i - eb_{i-1}.  m = match bits of this row aligned to the window and already ANDed with
// the row's diag-allowed mask.  rtw[k] = rt of column (eb_i - k).  See DESIGN.md "fast DP kernel" for the proof
// that this equals the banded recurrence of solver.jl:388-423 whenever the band table passes fast_path_ok().
// max(a + b, c): one DPX instruction (VIADDMNMX) on sm_90+/sm_100a
NW_HD int viaddmax(int a, int b, int c) {
#if defined(__CUDA_ARCH__)
  return __viaddmax_s32(a, b, c);
#else
  const int t = a + b;
  return t > c ? t : c;
#endif
}

// `runp` carries (best diagonal candidate so far in this row) - up, so that both the predicated diagonal update
// and the cell update are a single add+max:   runp = max(S_old[ds] + rt_k, runp) ;  S[k] = max(runp + up, S_old[us]).
template <int W, int SB, typename RT>
NW_HD void dp_row_step(int (&S)[W], const uint32_t* m, int up, const RT& rtw) {
  int runp = DP_NEG;
  int saved = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = W - 1; k >= 0; --k) {
    const int us = (k - SB >= 0) ? (k - SB) : 0;
    const int ds = k - SB + 1;
    const int upv = S[us];
    if (ds >= 0 && ds <= W - 1) {
      const int dv = (SB == 0) ? saved : S[ds];
      if ((m[k >> 5] >> (k & 31)) & 1u) runp = viaddmax(dv, rtw[k], runp);
    }
    if (SB == 0) saved = S[k];
    S[k] = viaddmax(runp, up, upv);
  }
}

// pure shift by one column (no matches), used to decompose SB > 2
template <int W, typename T>
NW_HD void dp_row_shift1(T (&S)[W]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = W - 1; k >= 1; --k) S[k] = S[k - 1];
}

// ---------------------------------------------------------------- packed row step: two candidates per register
// The same recurrence on two candidates of one child length at once, one in each 16-bit half of a register, with the
// packed DPX instruction (VIADDMNMX.S16x2: per-half max(a + b, c)).  Valid when every candidate's total is below
// PAIR_TOTAL_MAX units: every intermediate (a savings value, or savings + rt + up of a path prefix) is bounded by the
// total, so no half ever overflows into its neighbour.
// The diagonal update cannot be predicated per half, so the match bit is folded into the addend instead:
//     addend_half = rt | (match ? 0 : 0x8000)          (= rt - 32768 when the cell does not match)
// a masked candidate lands in [-32768, -1] and loses every max against the savings (>= 0) it competes with, also
// after `up` is added (dv + rt + up <= total < 32768).  Mq[g] holds the match bits of window cells 16g .. 16g+15:
// candidate A's in bits 0..15, candidate B's in bits 16..31, so `Mq[g] << (15 - c)` (an IMAD.SHL on the FMA pipe)
// brings both bits of cell 16g + c to the sign positions and ONE LOP3 builds the addend: three ALU-pipe
// instructions per cell PAIR against two per cell in dp_row_step.
constexpr int PAIR_TOTAL_MAX = 32767;
constexpr uint32_t PAIR_NEG = 0x80008000u;  // (-32768, -32768)

#if !defined(__CUDACC__)
inline long pair_wraps = 0;  // host restatement only: 16-bit additions that wrapped (the range argument says: none, ever)
#endif
NW_HD uint32_t viaddmax2(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  return __viaddmax_s16x2(a, b, c);
#else
  uint32_t out = 0;
  for (int h = 0; h < 2; ++h) {
#if !defined(__CUDACC__)
    const int wide = (int)(int16_t)(uint16_t)((a >> (16 * h)) & 0xffffu) + (int)(int16_t)(uint16_t)((b >> (16 * h)) & 0xffffu);
    if (wide < -32768 || wide > 32767) ++pair_wraps;
#endif
    const int16_t x = (int16_t)(uint16_t)(((a >> (16 * h)) + (b >> (16 * h))) & 0xffffu);  // wrapping 16-bit add
    const int16_t y = (int16_t)(uint16_t)((c >> (16 * h)) & 0xffffu);
    out |= (uint32_t)(uint16_t)(x > y ? x : y) << (16 * h);
  }
  return out;
#endif
}

template <int W, int SB, typename RT>
NW_HD void dp_row_step2(uint32_t (&S)[W], const uint32_t* Mq, uint32_t up2, const RT& rt2) {
  uint32_t runp = PAIR_NEG;
  uint32_t saved = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = W - 1; k >= 0; --k) {
    const int us = (k - SB >= 0) ? (k - SB) : 0;
    const int ds = k - SB + 1;
    const uint32_t upv = S[us];
    if (ds >= 0 && ds <= W - 1) {
      const uint32_t dv = (SB == 0) ? saved : S[ds];
      const uint32_t v = Mq[k >> 4] << (15 - (k & 15));
      const uint32_t a2 = (uint32_t)rt2[k] | (~v & PAIR_NEG);
      runp = viaddmax2(dv, a2, runp);
    }
    if (SB == 0) saved = S[k];
    S[k] = viaddmax2(runp, up2, upv);
  }
}

}  // namespace nw
