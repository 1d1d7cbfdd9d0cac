// nw_score_pair.cu -- register-resident banded DP (slow_score, solver.jl:329-426), TWO candidates per thread.
//
// Same recurrence, window and band tables as k_score_fast (nw_score_fast.cu), but every register of the DP row holds
// the savings of two candidates of one child length, one per 16-bit half, and the row step runs on the packed DPX
// instruction (VIADDMNMX.S16x2): dp_row_step2, nw_device.h.  Per window cell PAIR the ALU pipe sees one LOP3 (the
// match bit of either candidate folded into the addend) and two packed DPX instructions -- 1.5 per cell against 2 in
// the one-candidate form -- and the per-row overhead (band row, window shift, rt window) is paid once for two
// candidates.  The host routes a (region, length) bucket here only when every total fits 15 bits
// (pair_range_ok, nw_host.h); everything else keeps the 32-bit kernel.
//
// CTA = 64 threads = 128 work slots (the same slot -> CTA -> region mapping as k_score_fast, so the work list and
// cta_rid need no second layout); a warp owns 64 consecutive slots of ONE length (the host pads the buckets of a
// packed launch to 64): lane l scores slots l and 32 + l.
#include "nw_kernels.cuh"

namespace nw {

constexpr int PAIR_THREADS = 64;
// Build knobs, both measured on the B200 (profiles/README.md): the defaults are the fastest.
#ifndef NW_PAIR_RTREG_MAX
#define NW_PAIR_RTREG_MAX 64  // widest window whose packed rt values are held in registers (LDS.128) instead of read per cell
#endif
#ifndef NW_PAIR_MINB_WIDE
#define NW_PAIR_MINB_WIDE 4   // CTAs per SM for 40 < W <= 64 (6 = 168 registers, by spilling: 2 x slower)
#endif
// CTAs per SM the register allocation has to leave room for (64 threads each): 128 registers up to W = 28, 168 up to 40
constexpr int pair_min_ctas(int W) { return W <= 28 ? 8 : (W <= 40 ? 6 : (W <= 64 ? NW_PAIR_MINB_WIDE : 4)); }
constexpr int PAIR_BROW = 8;  // uint32 per staged band row: {r0, plane word offset, rt byte offset, sb, mask[0..3]} = two LDS.128

// one candidate's index map over its parent (apply_* of solver.jl:244-278, never materialised)
struct PairMap {
  const int16_t* seq;
  int lo, hi, shift, pq;
  // child[t] = sgn(t) * seq[src(t)]:  dup: t<=q ? t : t-q+p ; del: t<p ? t : t+q-p ; inv: p<t<q ? p+q-t (negated) : t
  __device__ __forceinline__ void init(const ScoreArgs& A, uint64_t d, int Lc) {
    const uint32_t ps = desc_parent(d);
    const int kind = desc_kind(d), p = desc_i(d), q = desc_j(d);
    seq = reinterpret_cast<const int16_t*>(A.F.blob + (size_t)ps * A.F.pstride);
    lo = kind == KIND_DEL ? p - 1 : (kind == KIND_DUP ? q : (kind == KIND_INV ? p : Lc));
    hi = kind == KIND_INV ? q : 0;  // inversion window is (lo, hi); empty for the other kinds
    shift = kind == KIND_DUP ? p - q : (kind == KIND_DEL ? q - p : 0);
    pq = p + q;
  }
  // element t (1-based) of the child as (parent element, negate?): the sign flip of an inverted segment is applied by
  // the consumer one row later, so nothing waits for the load in the iteration that issues it
  __device__ __forceinline__ void elem(int t, int& v, uint32_t& neg) const {
    const bool seg2 = t > lo;
    int src = seg2 ? t + shift : t;
    const bool flip = seg2 && t < hi;
    src = flip ? pq - t : src;
    neg = flip ? 1u : 0u;
    v = seq[src - 1];
  }
};
// plane row of a block id: block * 2 + strand (addresses both the match bit-plane and the up cost)
__device__ __forceinline__ uint32_t plane_row(int v, uint32_t neg) {
  const uint32_t a = (uint32_t)(v < 0 ? -v : v);
  return __funnelshift_l((uint32_t)v, a, 1) ^ neg;  // |v| * 2 + sign bit, strand flipped inside an inversion
}

// Shared-memory layout per CTA:  [ region context (TMA bulk copy; rt_rev packed in place) | band rows uint32 [2][Lmax+2][8] ]
// The child sequences are NOT staged: a row's two block ids are read through the index maps two rows ahead (one LDG
// each, L1-resident parents) and its match words / up costs one row ahead, so the only dependent chain of a row is
// the DPX chain itself and a CTA needs no per-candidate tile (S150: 84 KB -> 33 KB of shared memory per CTA).
template <int W>
__global__ void __launch_bounds__(PAIR_THREADS, pair_min_ctas(W)) k_score_pair(ScoreArgs A) {
  constexpr int NWORD = (W + 31) / 32;
  constexpr int BROW = PAIR_BROW;
  constexpr bool RTREG = (W <= NW_PAIR_RTREG_MAX);  // rt window in registers (LDS.128) vs one LDS per cell (wide windows: register budget)
  constexpr int RTW = RTREG ? W : 4;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const RegionDev& R = A.F.regions[A.cta_rid ? (int)A.cta_rid[blockIdx.x] : 0];
  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, R.ctx_bytes);
    tma_bulk_g2s(smem, R.ctx, R.ctx_bytes, &bar);
  }
  const uint32_t* planes = reinterpret_cast<const uint32_t*>(smem);
  const int32_t* upx = reinterpret_cast<const int32_t*>(smem + R.off_upx);
  uint32_t* rtrev = reinterpret_cast<uint32_t*>(smem + R.off_rtrev);
  const int Lmax = A.lmax;
  uint32_t* brow = reinterpret_cast<uint32_t*>(smem + A.ctx_max) + (size_t)warp * (Lmax + 2) * BROW;

  const uint32_t w0 = blockIdx.x * 128u + warp * 64u + lane;
  uint32_t rA = w0 < A.n_work ? A.work[w0] : WORK_PAD;
  uint32_t rB = w0 + 32u < A.n_work ? A.work[w0 + 32u] : WORK_PAD;
  const uint32_t actA = __ballot_sync(0xffffffffu, rA != WORK_PAD);
  // slots fill from the front of a 64-padded bucket: an idle first half means the whole 64-slot group is padding
  const bool warp_idle = actA == 0;
  const bool mineA = rA != WORK_PAD, mineB = rB != WORK_PAD;
  int Lc = 0;
  PairMap mapA, mapB;
  if (!warp_idle) {
    const uint32_t rlead = __shfl_sync(0xffffffffu, rA, __ffs(actA) - 1);
    if (!mineA) rA = rlead;  // idle halves shadow a real candidate of the same length (results are not written)
    if (!mineB) rB = rA;
    const uint64_t dA = A.desc[A.newlist[rA]], dB = A.desc[A.newlist[rB]];
    Lc = child_len(A.F.len[desc_parent(dA)], desc_kind(dA), desc_i(dA), desc_j(dA));  // warp-uniform (buckets)
    mapA.init(A, dA, Lc);
    mapB.init(A, dB, Lc);
    // ---- stage the warp's band rows (row Lc + 1 = zeros: the look-ahead of the last row reads it) ----
    const int boff = R.band_off[Lc];
    const int16_t* eb_t = R.band_eb + boff;
    const uint8_t* sb_t = R.band_sb + boff;
    const uint4* mk_t = reinterpret_cast<const uint4*>(R.band_mask + (size_t)boff * 4);
    const int NYP1 = R.NYP - 1;
    const int rt_stride = R.rt_copy_stride;
    for (int i = 1 + lane; i <= Lc + 1; i += 32) {
      uint4 g = make_uint4(0u, 0u, 0u, 0u), mk = g;
      if (i <= Lc) {
        const int r0 = NYP1 - (int)eb_t[i];
        g = make_uint4((uint32_t)r0, (uint32_t)(r0 >> 5), (uint32_t)(((r0 & 3) * rt_stride + (r0 & ~3)) * 4), (uint32_t)sb_t[i]);
        mk = mk_t[i];
      }
      *reinterpret_cast<uint4*>(brow + i * BROW) = g;
      *reinterpret_cast<uint4*>(brow + i * BROW + 4) = mk;
    }
  }
  const int PWS = R.PWS;
  const int sum_rt = R.sum_rt;
  const int n_rt = R.rt_copy_stride * 4;
  mbar_wait(&bar, 0);
  // rt of both halves: rt | rt << 16, in place over the four shifted copies (every rt < 2^15 on this path)
  for (int t = tid; t < n_rt; t += PAIR_THREADS) rtrev[t] *= 0x10001u;
  __syncthreads();
  if (warp_idle) return;

  // Row inputs in two halves so that the loads of row t + 1 are in flight during the DP step of row t:
  // issue = band row (two LDS.128), plane words and up costs of both candidates; finish = match words of both
  // candidates interleaved by 16-cell groups, packed up cost, (rt byte offset, sb) of the row.
  struct RowLoads {
    uint4 g, mk4;
    uint32_t pa[NWORD + 1], pb[NWORD + 1];
    uint32_t ua, ub;
  };
  auto issue = [&](const uint32_t* bw, int vA, uint32_t nA, int vB, uint32_t nB, RowLoads& L) {
    L.g = *reinterpret_cast<const uint4*>(bw);
    L.mk4 = *reinterpret_cast<const uint4*>(bw + 4);
    const uint32_t ciA = plane_row(vA, nA), ciB = plane_row(vB, nB);
    L.ua = (uint32_t)upx[ciA];
    L.ub = (uint32_t)upx[ciB];
    const uint32_t* plA = planes + ciA * PWS + L.g.y;
    const uint32_t* plB = planes + ciB * PWS + L.g.y;
#pragma unroll
    for (int t = 0; t <= NWORD; ++t) {
      L.pa[t] = plA[t];
      L.pb[t] = plB[t];
    }
  };
  auto finish = [&](const RowLoads& L, uint32_t (&Mq)[2 * NWORD], uint32_t& up2) {
    const uint32_t mk[4] = {L.mk4.x, L.mk4.y, L.mk4.z, L.mk4.w};
    up2 = L.ua + (L.ub << 16);
#pragma unroll
    for (int t = 0; t < NWORD; ++t) {
      const uint32_t mA = __funnelshift_r(L.pa[t], L.pa[t + 1], L.g.x) & mk[t];  // shift = r0 mod 32
      const uint32_t mB = __funnelshift_r(L.pb[t], L.pb[t + 1], L.g.x) & mk[t];
      Mq[2 * t] = __byte_perm(mA, mB, 0x5410);      // cells 32t .. 32t+15: A in the low half, B in the high half
      Mq[2 * t + 1] = __byte_perm(mA, mB, 0x7632);  // cells 32t+16 .. 32t+31
    }
    return make_uint2(L.g.z, L.g.w);  // rt byte offset, sb
  };

  uint32_t S[W];
#pragma unroll
  for (int k = 0; k < W; ++k) S[k] = 0;
  uint32_t sumup2 = 0;
  const uint32_t* bw = brow + BROW;  // band row of the current DP row
  // software pipeline: elements two rows ahead (LDG), plane words / up / band row one row ahead (LDS)
  int vA, vB, vA2, vB2;
  uint32_t nA, nB, nA2, nB2;
  mapA.elem(1, vA, nA);
  mapB.elem(1, vB, nB);
  mapA.elem(min(2, Lc), vA2, nA2);
  mapB.elem(min(2, Lc), vB2, nB2);
  uint32_t Mq[2 * NWORD], up2;
  uint2 geo;
  {
    RowLoads L0;
    issue(bw, vA, nA, vB, nB, L0);
    geo = finish(L0, Mq, up2);
  }
  for (int t = 1; t <= Lc; ++t, bw += BROW) {
    const uint32_t* rtw = reinterpret_cast<const uint32_t*>(reinterpret_cast<const unsigned char*>(rtrev) + geo.x);
    const uint32_t sbw = geo.y;
    uint32_t rtr[RTW];
    if (RTREG) {
#pragma unroll
      for (int g = 0; g < RTW / 4; ++g) {
        const uint4 v = *reinterpret_cast<const uint4*>(rtw + 4 * g);
        rtr[4 * g] = v.x;
        rtr[4 * g + 1] = v.y;
        rtr[4 * g + 2] = v.z;
        rtr[4 * g + 3] = v.w;
      }
    }
    const int tn = min(t + 2, Lc);
    int vA3, vB3;
    uint32_t nA3, nB3;
    mapA.elem(tn, vA3, nA3);  // LDG, consumed in the next iteration
    mapB.elem(tn, vB3, nB3);
    RowLoads Ln;
    issue(bw + BROW, vA2, nA2, vB2, nB2, Ln);  // row t + 1 (row Lc + 1: zeros, unused)
    sumup2 += up2;
    if (RTREG) {
      if (sbw == 1) {
        dp_row_step2<W, 1>(S, Mq, up2, rtr);
      } else if (sbw == 0) {
        dp_row_step2<W, 0>(S, Mq, up2, rtr);
      } else {
        for (int e = (int)sbw - 2; e > 0; --e) dp_row_shift1<W>(S);  // rare: first rows of steep corridors
        dp_row_step2<W, 2>(S, Mq, up2, rtr);
      }
    } else {
      if (sbw == 1) {
        dp_row_step2<W, 1>(S, Mq, up2, rtw);
      } else if (sbw == 0) {
        dp_row_step2<W, 0>(S, Mq, up2, rtw);
      } else {
        for (int e = (int)sbw - 2; e > 0; --e) dp_row_shift1<W>(S);
        dp_row_step2<W, 2>(S, Mq, up2, rtw);
      }
    }
    geo = finish(Ln, Mq, up2);
    vA2 = vA3;
    vB2 = vB3;
    nA2 = nA3;
    nB2 = nB3;
  }
  const int totalA = (int)(sumup2 & 0xffffu) + sum_rt, totalB = (int)(sumup2 >> 16) + sum_rt;
  const int costA = totalA - (int)(S[0] & 0xffffu), costB = totalB - (int)(S[0] >> 16);
  const int kA = score_k3(costA, totalA), kB = score_k3(costB, totalB);
  if (mineA) {
    A.total[rA] = totalA;
    A.cost[rA] = costA;
    A.k3[rA] = kA;
  }
  if (mineB) {
    A.total[rB] = totalB;
    A.cost[rB] = costB;
    A.k3[rB] = kB;
  }
  int b = max(kA, kB);  // warp max -> one atomic (shadow halves carry a duplicate of a real value)
  for (int o = 16; o; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
  int* best_slot = A.best_k3 + (A.cta_rid ? (int)A.cta_rid[blockIdx.x] : 0);
  if (lane == 0 && b > 0) atomicMax(best_slot, b);
}

size_t score_pair_smem(int W, uint32_t ctx_bytes, int lmax) {
  (void)W;
  return (size_t)ctx_bytes + (size_t)2 * (lmax + 2) * PAIR_BROW * 4;
}

template <int W>
static int launch_w(const ScoreArgs& A, cudaStream_t st) {
  const size_t sm = score_pair_smem(W, A.ctx_max, A.lmax);
  k_score_pair<W><<<(A.n_work + 127) / 128, PAIR_THREADS, sm, st>>>(A);
  return 1;
}

#define NW_PAIR_WINDOWS(X) X(8) X(12) X(16) X(20) X(24) X(28) X(32) X(36) X(40) X(44) X(48) X(52) X(56) X(60) X(64) X(72) X(80) X(96) X(112) X(128)

void init_score_pair_attrs() {
  int dev = 0, optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const int mx = optin - 2048;  // the limit covers static + dynamic shared memory of a block
#define X(Wv) cudaFuncSetAttribute(k_score_pair<Wv>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  NW_PAIR_WINDOWS(X)
#undef X
}

int launch_score_pair(int W, const ScoreArgs& A, cudaStream_t st) {
  if (A.n_work == 0) return 0;
  switch (W) {
#define X(Wv) \
  case Wv:    \
    return launch_w<Wv>(A, st);
    NW_PAIR_WINDOWS(X)
#undef X
    default:
      return -1;
  }
}

}  // namespace nw
