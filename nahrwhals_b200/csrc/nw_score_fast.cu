// nw_score_fast.cu -- register-resident banded DP (slow_score, solver.jl:329-426) for sm_100a.
//
// One thread scores one candidate; the 32 candidates of a warp have the same child length, hence the same
// corridor geometry, so all control flow (row count, window shifts) is warp-uniform.  The DP row lives in W
// registers as a window right-aligned at the row's last valid column; the child's block permutation and
// strand flips are applied as an index map over the parent sequence; the match bits of a row come from the
// region's bit-planes staged once per CTA into shared memory with a TMA bulk copy (cp.async.bulk -> UBLKCP).
// Savings form: S = max(up, run) per cell + a predicated diagonal update (see nw_device.h / DESIGN.md).
#include "nw_kernels.cuh"

namespace nw {

template <int W>
__global__ void __launch_bounds__(128) k_score_fast(ScoreArgs A) {
  constexpr int NWORD = (W + 31) / 32;
  constexpr bool RTREG = (W <= 64);         // rt window in registers (LDS.128) vs predicated scalar LDS per cell
  constexpr int RTW = RTREG ? W : 4;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, A.R.ctx_bytes);
    tma_bulk_g2s(smem, A.R.ctx, A.R.ctx_bytes, &bar);
  }
  const uint32_t* planes = reinterpret_cast<const uint32_t*>(smem);
  const int32_t* upx = reinterpret_cast<const int32_t*>(smem + A.R.off_upx);
  const int32_t* rtrev = reinterpret_cast<const int32_t*>(smem + A.R.off_rtrev);

  const uint32_t w = blockIdx.x * 128u + tid;
  uint32_t r = w < A.n_work ? A.work[w] : WORK_PAD;
  const uint32_t act = __ballot_sync(0xffffffffu, r != WORK_PAD);
  if (act == 0) {  // whole warp idle (bucket padding) -- still has to observe the barrier before exiting
    mbar_wait(&bar, 0);
    return;
  }
  const bool mine = r != WORK_PAD;
  const uint32_t rlead = __shfl_sync(0xffffffffu, r, __ffs(act) - 1);
  if (!mine) r = rlead;  // idle lanes shadow the first active lane (results are not written)
  const uint64_t d = A.desc[A.newlist[r]];
  const uint32_t ps = desc_parent(d);
  const int kind = desc_kind(d), p = desc_i(d), q = desc_j(d);
  const int L = A.F.len[ps];
  const int16_t* seq = reinterpret_cast<const int16_t*>(A.F.blob + (size_t)ps * A.F.pstride);
  const int Lc = child_len(L, kind, p, q);  // warp-uniform by construction of the buckets
  const int boff = A.band_off[Lc];
  const int16_t* eb_t = A.band_eb + boff;
  const uint8_t* sb_t = A.band_sb + boff;
  const uint32_t* mk_t = A.band_mask + (size_t)boff * 4;
  const int NYP1 = A.R.NYP - 1, PWS = A.R.PWS;

  int S[W];
#pragma unroll
  for (int k = 0; k < W; ++k) S[k] = 0;
  int sumup = 0;
  int c = child_elem(seq, kind, p, q, 1);
  mbar_wait(&bar, 0);
  for (int i = 1; i <= Lc; ++i) {
    const int cn = (i < Lc) ? child_elem(seq, kind, p, q, i + 1) : 0;  // software prefetch of the next row's block
    const int x = c < 0 ? -c : c;
    const int up = upx[x];
    sumup += up;
    const int r0 = NYP1 - (int)eb_t[i];
    int sb = sb_t[i];
    const uint32_t* pl = planes + (size_t)(x * 2 + (c < 0 ? 1 : 0)) * PWS + (r0 >> 5);
    const int sh = r0 & 31;
    uint32_t m[NWORD];
#pragma unroll
    for (int t = 0; t < NWORD; ++t) m[t] = __funnelshift_r(pl[t], pl[t + 1], sh) & mk_t[(size_t)i * 4 + t];
    // rt of the window columns: 4 shifted copies of rt_rev make the window start 16-B aligned for LDS.128
    const int* rtw = rtrev + (size_t)(r0 & 3) * A.R.rt_copy_stride + (r0 & ~3);
    int rtr[RTW];
    if (RTREG) {
#pragma unroll
      for (int g = 0; g < RTW / 4; ++g) {
        const int4 v = *reinterpret_cast<const int4*>(rtw + 4 * g);
        rtr[4 * g] = v.x;
        rtr[4 * g + 1] = v.y;
        rtr[4 * g + 2] = v.z;
        rtr[4 * g + 3] = v.w;
      }
    }
    while (sb > 2) {
      dp_row_shift1<W>(S);
      --sb;
    }
    if (RTREG) {
      if (sb == 0)
        dp_row_step<W, 0>(S, m, up, rtr);
      else if (sb == 1)
        dp_row_step<W, 1>(S, m, up, rtr);
      else
        dp_row_step<W, 2>(S, m, up, rtr);
    } else {
      if (sb == 0)
        dp_row_step<W, 0>(S, m, up, rtw);
      else if (sb == 1)
        dp_row_step<W, 1>(S, m, up, rtw);
      else
        dp_row_step<W, 2>(S, m, up, rtw);
    }
    c = cn;
  }
  const int total = sumup + A.R.sum_rt;
  const int cost = total - S[0];
  const int k3 = score_k3(cost, total);
  if (mine) {
    A.total[r] = total;
    A.cost[r] = cost;
    A.k3[r] = k3;
  }
  int b = k3;  // warp max -> one atomic (shadow lanes carry a duplicate of a real value)
  for (int o = 16; o; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
  if (lane == 0) atomicMax(A.best_k3, b);
}

static const int kWindows[] = {8, 16, 24, 32, 40, 48, 64, 80, 96, 128};

int score_fast_window(int width) {
  for (int w : kWindows)
    if (w >= width) return w;
  return 0;
}
bool score_fast_supports(int width) { return score_fast_window(width) != 0; }

template <int W>
static int launch_w(const ScoreArgs& A, cudaStream_t st) {
  const size_t sm = A.R.ctx_bytes;
  cudaFuncSetAttribute(k_score_fast<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  k_score_fast<W><<<(A.n_work + 127) / 128, 128, sm, st>>>(A);
  return 1;
}

int launch_score_fast(int W, const ScoreArgs& A, cudaStream_t st) {
  if (A.n_work == 0) return 0;
  switch (W) {
    case 8: return launch_w<8>(A, st);
    case 16: return launch_w<16>(A, st);
    case 24: return launch_w<24>(A, st);
    case 32: return launch_w<32>(A, st);
    case 40: return launch_w<40>(A, st);
    case 48: return launch_w<48>(A, st);
    case 64: return launch_w<64>(A, st);
    case 80: return launch_w<80>(A, st);
    case 96: return launch_w<96>(A, st);
    case 128: return launch_w<128>(A, st);
    default: return -1;
  }
}

}  // namespace nw
