// nw_score_coop.cu -- warp-cooperative form of the banded DP (slow_score, solver.jl:329-426) for sm_100a.
//
// The register-resident kernel of nw_score_fast.cu gives one candidate to one thread: W window columns in W registers
// and a dependent chain of W add+max steps per row.  That is the throughput form -- but one launch can never be faster
// than ONE thread walking L rows x W columns (~70 us at 150 blocks), and at W = 64 its 247 registers leave 8 warps per
// SM.  Here FOUR lanes share a candidate: lane q owns the window columns [q*WL, (q+1)*WL), WL = W / 4, and a row is
//   1. neighbour cells of the previous row by shuffle (the window moves SB columns per row: SB = 1 needs the right
//      neighbour's top cell, SB = 2 its top two, SB = 0 the left neighbour's bottom cell);
//   2. the lane's own chain over WL columns (same two DPX instructions per cell as the thread kernel), which leaves the
//      best diagonal candidate of the lane's columns in `run`;
//   3. an exclusive suffix maximum of `run` over the lanes to the left (three shuffles inside the 4-lane group): the
//      savings form makes right moves free, so a row is a prefix maximum along the columns and the carry is one number;
//   4. one more add+max per cell folds the carry in.
// Three DPX instructions per cell instead of two and four times the row prologue, but a quarter of the chain, a third of
// the registers and 8 candidates per warp: the latency floor of a launch drops ~3.5x and wide windows gain occupancy.
// Results are bit-identical to the thread kernel (same recurrence, integers).  The driver picks the form per launch
// (run_scoring); NW_DP_COOP=0/1 forces one of them for benchmarks and parity runs.
#include "nw_kernels.cuh"

namespace nw {

// Shared memory per CTA (128 threads = 4 warps x 8 candidates = one 32-slot bucket segment of the work list):
//   [ region context (TMA bulk copy) | child rows uint16 [Lmax][32] | band rows (per warp) uint32 [4][Lmax+1][3+NWORD] ]
template <int W>
__global__ void __launch_bounds__(128) k_score_coop(ScoreArgs A) {
  constexpr int WL = W / 4;                 // columns per lane
  constexpr int NWORD = (W + 31) / 32;
  constexpr int BROW = 3 + NWORD;
  static_assert(W % 16 == 0, "the rt window of a lane is read as 16-byte vectors");
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = lane & 3;                   // position of this lane inside its candidate's group
  const int grp = tid >> 2;                 // candidate slot of the CTA (0..31)
  const RegionDev& R = A.F.regions[A.cta_rid ? (int)A.cta_rid[blockIdx.x >> 2] : 0];  // cta_rid is per 128 work slots
  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, R.ctx_bytes);
    tma_bulk_g2s(smem, R.ctx, R.ctx_bytes, &bar);
  }
  const uint32_t* planes = reinterpret_cast<const uint32_t*>(smem);
  const int32_t* upx = reinterpret_cast<const int32_t*>(smem + R.off_upx);
  const int32_t* rtrev = reinterpret_cast<const int32_t*>(smem + R.off_rtrev);
  const int Lmax = A.lmax;
  uint16_t* crow = reinterpret_cast<uint16_t*>(smem + A.ctx_max);
  uint32_t* brow = reinterpret_cast<uint32_t*>(smem + A.ctx_max + (size_t)Lmax * 64) + (size_t)warp * (Lmax + 1) * BROW;

  const uint32_t w = blockIdx.x * 32u + (uint32_t)grp;
  uint32_t r = w < A.n_work ? A.work[w] : WORK_PAD;
  const uint32_t act = __ballot_sync(0xffffffffu, r != WORK_PAD);
  if (act == 0) {  // whole warp idle (bucket padding) -- still has to observe the barrier before exiting
    mbar_wait(&bar, 0);
    return;
  }
  const bool mine = r != WORK_PAD;
  const uint32_t rlead = __shfl_sync(0xffffffffu, r, __ffs(act) - 1);
  if (!mine) r = rlead;  // idle groups shadow the first active one (results are not written)
  const uint64_t d = A.desc[A.newlist[r]];
  const uint32_t ps = desc_parent(d);
  const int kind = desc_kind(d), p = desc_i(d), qq = desc_j(d);
  const int L = A.F.len[ps];
  const int16_t* seq = reinterpret_cast<const int16_t*>(A.F.blob + (size_t)ps * A.F.pstride);
  const int Lc = child_len(L, kind, p, qq);  // warp-uniform by construction of the buckets
  // ---- stage the candidate's child sequence: its four lanes take every fourth element ----
  {
    const int lo = kind == KIND_DEL ? p - 1 : (kind == KIND_DUP ? qq : (kind == KIND_INV ? p : Lc));
    const int hi = kind == KIND_INV ? qq : 0x7fffffff;
    const int shift = kind == KIND_DUP ? p - qq : (kind == KIND_DEL ? qq - p : 0);
    for (int t = 1 + q; t <= Lc; t += 4) {
      int src = t > lo ? t + shift : t;
      const bool flip = kind == KIND_INV && t > lo && t < hi;
      src = flip ? p + qq - t : src;
      const int v = seq[src - 1];
      const int a = v < 0 ? -v : v;
      crow[(size_t)(t - 1) * 32 + grp] = (uint16_t)(a * 2 + ((v < 0) != flip ? 1 : 0));
    }
  }
  // ---- stage the warp's band rows ----
  {
    const int boff = R.band_off[Lc];
    const int16_t* eb_t = R.band_eb + boff;
    const uint8_t* sb_t = R.band_sb + boff;
    const uint32_t* mk_t = R.band_mask + (size_t)boff * 4;
    const int NYP1 = R.NYP - 1;
    const int rt_stride = R.rt_copy_stride;
    for (int i = 1 + lane; i <= Lc; i += 32) {
      const int r0 = NYP1 - (int)eb_t[i];
      brow[i * BROW] = (uint32_t)r0;
      brow[i * BROW + 1] = (uint32_t)(((r0 & 3) * rt_stride + (r0 & ~3)) * 4);
      brow[i * BROW + 2] = (uint32_t)sb_t[i];
#pragma unroll
      for (int t = 0; t < NWORD; ++t) brow[i * BROW + 3 + t] = mk_t[(size_t)i * 4 + t];
    }
  }
  const int PWS = R.PWS;
  const int sum_rt = R.sum_rt;
  int* best_slot = A.best_k3 + (A.cta_rid ? (int)A.cta_rid[blockIdx.x >> 2] : 0);
  __syncwarp();
  mbar_wait(&bar, 0);

  int S[WL];
#pragma unroll
  for (int k = 0; k < WL; ++k) S[k] = 0;
  int sumup = 0;
  const uint16_t* cp = crow + grp;
  const uint32_t* bw = brow + BROW;
  const int bit0 = q * WL;  // first window index of this lane
  for (int i = 0; i < Lc; ++i, cp += 32, bw += BROW) {
    const uint32_t ci = *cp;
    const uint32_t r0 = bw[0];
    const int sbw = (int)bw[2];
    const int up = upx[ci];
    sumup += up;
    // match bits of the lane's WL columns (window bit k <-> column eb_i - k), already ANDed with the diag-allowed mask
    uint32_t mb;
    {
      const uint32_t* pl = planes + ci * PWS + (r0 >> 5);
      const int wd = bit0 >> 5, sh = bit0 & 31;
      const uint32_t m0 = __funnelshift_r(pl[wd], pl[wd + 1], r0) & bw[3 + wd];
      if (WL == 32) {
        mb = m0;
      } else {
        uint32_t m1 = 0;
        if (sh + WL > 32) m1 = __funnelshift_r(pl[wd + 1], pl[wd + 2], r0) & bw[3 + ((wd + 1) < NWORD ? wd + 1 : wd)] &
                               (((wd + 1) < NWORD) ? 0xffffffffu : 0u);
        mb = __funnelshift_r(m0, m1, sh) & ((1u << (WL & 31)) - 1u);
      }
    }
    // rt of the lane's columns: 16-byte vectors from the aligned copy of the reversed y lengths
    const int* rtw = reinterpret_cast<const int*>(reinterpret_cast<const unsigned char*>(rtrev) + bw[1]) + bit0;
    int rtr[WL];
#pragma unroll
    for (int g = 0; g < WL / 4; ++g) {
      const int4 v = *reinterpret_cast<const int4*>(rtw + 4 * g);
      rtr[4 * g] = v.x;
      rtr[4 * g + 1] = v.y;
      rtr[4 * g + 2] = v.z;
      rtr[4 * g + 3] = v.w;
    }
    // rare: the window jumps by more than two columns (first rows of steep corridors): single shifts first
    int sb = sbw;
    for (; sb > 2; --sb) {
      const int in = __shfl_up_sync(0xffffffffu, S[WL - 1], 1, 4);  // right neighbour's top cell
#pragma unroll
      for (int k = WL - 1; k >= 1; --k) S[k] = S[k - 1];
      if (q > 0) S[0] = in;  // window index 0 keeps its value (dp_row_shift1)
    }
    // X[m] = previous row at window index bit0 + m - sb, m = 0..WL (clamped at the window's right edge)
    int X[WL + 1];
    if (sb == 1) {
      const int in = __shfl_up_sync(0xffffffffu, S[WL - 1], 1, 4);
      X[0] = q > 0 ? in : S[0];
#pragma unroll
      for (int m = 1; m <= WL; ++m) X[m] = S[m - 1];
    } else if (sb == 0) {
      const int in = __shfl_down_sync(0xffffffffu, S[0], 1, 4);  // left neighbour's bottom cell
#pragma unroll
      for (int m = 0; m < WL; ++m) X[m] = S[m];
      X[WL] = in;
      if (q == 3) mb &= ~(1u << (WL - 1));  // no cell left of the window (the mask never allows it; guard)
    } else {
      const int in0 = __shfl_up_sync(0xffffffffu, S[WL - 2], 1, 4);
      const int in1 = __shfl_up_sync(0xffffffffu, S[WL - 1], 1, 4);
      X[0] = q > 0 ? in0 : S[0];
      X[1] = q > 0 ? in1 : S[0];
#pragma unroll
      for (int m = 2; m <= WL; ++m) X[m] = S[m - 2];
      if (q == 0) mb &= ~1u;  // window index 0 has no diagonal source when the window moved by two (guard)
    }
    // the lane's own chain, left to right (window index WL-1 down to 0)
    int run = DP_NEG;
#pragma unroll
    for (int k = WL - 1; k >= 0; --k) {
      if ((mb >> k) & 1u) run = viaddmax(X[k + 1], rtr[k], run);
      S[k] = viaddmax(run, up, X[k]);
    }
    // best diagonal candidate of the columns to the left of this lane's (lanes q+1..3 of the group)
    int inc = run;
    int t1 = __shfl_down_sync(0xffffffffu, inc, 1, 4);
    if (q < 3) inc = max(inc, t1);
    t1 = __shfl_down_sync(0xffffffffu, inc, 2, 4);
    if (q < 2) inc = max(inc, t1);
    int carry = __shfl_down_sync(0xffffffffu, inc, 1, 4);
    if (q == 3) carry = DP_NEG;
#pragma unroll
    for (int k = 0; k < WL; ++k) S[k] = viaddmax(carry, up, S[k]);
  }
  const int total = sumup + sum_rt;
  const int cost = total - S[0];  // meaningful in lane 0 of the group (window index 0 = last column)
  const int k3 = score_k3(cost, total);
  if (mine && q == 0) {
    A.total[r] = total;
    A.cost[r] = cost;
    A.k3[r] = k3;
  }
  int b = q == 0 ? k3 : 0;
  for (int o = 16; o; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
  if (lane == 0 && b > 0) atomicMax(best_slot, b);
}

static const int kCoopWindows[] = {16, 32, 48, 64, 80, 96, 112, 128};

int score_coop_window(int width) {
  for (int w : kCoopWindows)
    if (w >= width) return w;
  return 0;
}
size_t score_coop_smem(int W, uint32_t ctx_bytes, int lmax) {
  const int nword = (W + 31) / 32;
  return (size_t)ctx_bytes + (size_t)lmax * 64 + (size_t)4 * (lmax + 1) * (3 + nword) * 4;
}

template <int W>
static int launch_cw(const ScoreArgs& A, cudaStream_t st) {
  const size_t sm = score_coop_smem(W, A.ctx_max, A.lmax);
  k_score_coop<W><<<(A.n_work + 31) / 32, 128, sm, st>>>(A);
  return 1;
}

void init_score_coop_attrs() {
  int dev = 0, optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const int mx = optin - 2048;
  cudaFuncSetAttribute(k_score_coop<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_coop<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_coop<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_coop<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_coop<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_coop<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_coop<112>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_coop<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
}

int launch_score_coop(int W, const ScoreArgs& A, cudaStream_t st) {
  if (A.n_work == 0) return 0;
  switch (W) {
    case 16: return launch_cw<16>(A, st);
    case 32: return launch_cw<32>(A, st);
    case 48: return launch_cw<48>(A, st);
    case 64: return launch_cw<64>(A, st);
    case 80: return launch_cw<80>(A, st);
    case 96: return launch_cw<96>(A, st);
    case 112: return launch_cw<112>(A, st);
    case 128: return launch_cw<128>(A, st);
    default: return -1;
  }
}

}  // namespace nw
