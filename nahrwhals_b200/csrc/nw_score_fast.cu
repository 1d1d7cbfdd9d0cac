// nw_score_fast.cu -- register-resident banded DP (slow_score, solver.jl:329-426) for sm_100a.
//
// One thread scores one candidate; the 32 candidates of a warp have the same child length, hence the same
// corridor geometry, so all control flow (row count, window shifts) is warp-uniform.  The DP row lives in W
// registers as a window right-aligned at the row's last valid column; the child's block permutation and
// strand flips are applied as an index map over the parent sequence; the match bits of a row come from the
// region's bit-planes staged once per CTA into shared memory with a TMA bulk copy (cp.async.bulk -> UBLKCP).
// Savings form: S = max(up, run) per cell + a predicated diagonal update (see nw_device.h / DESIGN.md).
#include "nw_kernels.cuh"

#ifndef NW_DP_MINB
#define NW_DP_MINB 1
#endif
#ifndef NW_RTREG_MAX
#define NW_RTREG_MAX 64  // widest window whose rt values are held in registers (LDS.128) instead of read per cell
#endif

namespace nw {

// Shared-memory layout per CTA (128 threads = 4 warps):
//   [ region context (TMA bulk copy) | child rows uint16 [Lmax][128] | band rows (per warp) uint32 [4][Lmax+1][3+NWORD] ]
// The child's blocks are materialised once through the index map into a lane-interleaved smem tile -- stored as the
// plane-row index (block * 2 + strand) that addresses both the match bit-plane and the up cost -- so the row
// loop never waits on global memory; the warp's band table (one length per warp) is copied next to it.
template <int W>
__global__ void __launch_bounds__(128, NW_DP_MINB) k_score_fast(ScoreArgs A) {
  constexpr int NWORD = (W + 31) / 32;
  constexpr bool RTREG = (W <= NW_RTREG_MAX);  // rt window in registers (LDS.128) vs predicated scalar LDS per cell
  constexpr int RTW = RTREG ? W : 4;
  // uint32 words per band row, everything the row loop needs ready-made (no unpacking on the ALU pipe):
  //   r0 (window start in reversed columns), byte offset of the 16-B aligned rt window, sb, masks
  constexpr int BROW = 3 + NWORD;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // all candidates of a CTA belong to one region (the host pads region segments to CTA boundaries)
  const RegionDev& R = A.F.regions[A.cta_rid ? (int)A.cta_rid[blockIdx.x] : 0];
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
  uint32_t* brow = reinterpret_cast<uint32_t*>(smem + A.ctx_max + (size_t)Lmax * 256) + (size_t)warp * (Lmax + 1) * BROW;

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
  // ---- stage this lane's child sequence (index map over the parent, branch-free) ----
  {
    // child[t] = sgn(t) * seq[src(t)]:  dup: t<=q ? t : t-q+p ; del: t<p ? t : t+q-p ; inv: p<t<q ? p+q-t (negated) : t
    const int lo = kind == KIND_DEL ? p - 1 : (kind == KIND_DUP ? q : (kind == KIND_INV ? p : Lc));
    const int hi = kind == KIND_INV ? q : 0x7fffffff;  // inversion window is (lo, hi)
    const int shift = kind == KIND_DUP ? p - q : (kind == KIND_DEL ? q - p : 0);
#pragma unroll 4
    for (int t = 1; t <= Lc; ++t) {
      int src = t > lo ? t + shift : t;
      const bool flip = kind == KIND_INV && t > lo && t < hi;
      src = flip ? p + q - t : src;
      const int v = seq[src - 1];
      const int a = v < 0 ? -v : v;
      crow[(size_t)(t - 1) * 128 + tid] = (uint16_t)(a * 2 + ((v < 0) != flip ? 1 : 0));
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
      const int sb = sb_t[i];
      brow[i * BROW] = (uint32_t)r0;
      // rt of the window columns: 4 shifted copies of rt_rev make the window start 16-B aligned for LDS.128
      brow[i * BROW + 1] = (uint32_t)(((r0 & 3) * rt_stride + (r0 & ~3)) * 4);
      brow[i * BROW + 2] = (uint32_t)sb;  // 0 / 1 / 2 = the row step's shift; > 2: sb - 2 single shifts first, then step<2>
#pragma unroll
      for (int t = 0; t < NWORD; ++t) brow[i * BROW + 3 + t] = mk_t[(size_t)i * 4 + t];
    }
  }
  const int PWS = R.PWS;
  const int sum_rt = R.sum_rt;
  int* best_slot = A.best_k3 + (A.cta_rid ? (int)A.cta_rid[blockIdx.x] : 0);
  __syncwarp();
  mbar_wait(&bar, 0);

  int S[W];
#pragma unroll
  for (int k = 0; k < W; ++k) S[k] = 0;
  int sumup = 0;
  // running pointers (one add per row each) instead of re-deriving both addresses from the row number
  const uint16_t* cp = crow + tid;
  const uint32_t* bw = brow + BROW;
  const uint16_t* const cend = cp + (size_t)Lc * 128;
  for (; cp != cend; cp += 128, bw += BROW) {
    const uint32_t ci = *cp;  // plane row = block * 2 + strand; bw: warp-uniform broadcast loads
    const uint32_t r0 = bw[0];
    const uint32_t sbw = bw[2];
    const int up = upx[ci];
    sumup += up;
    const uint32_t* pl = planes + ci * PWS + (r0 >> 5);
    uint32_t m[NWORD];
#pragma unroll
    for (int t = 0; t < NWORD; ++t) m[t] = __funnelshift_r(pl[t], pl[t + 1], r0) & bw[3 + t];  // shift = r0 mod 32
    const int* rtw = reinterpret_cast<const int*>(reinterpret_cast<const unsigned char*>(rtrev) + bw[1]);
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
    // the common shifts first: one compare + branch each, no unpacking of the band word
    if (RTREG) {
      if (sbw == 1) {
        dp_row_step<W, 1>(S, m, up, rtr);
      } else if (sbw == 0) {
        dp_row_step<W, 0>(S, m, up, rtr);
      } else {
        for (int e = (int)sbw - 2; e > 0; --e) dp_row_shift1<W>(S);  // rare: first rows of steep corridors
        dp_row_step<W, 2>(S, m, up, rtr);
      }
    } else {
      if (sbw == 1) {
        dp_row_step<W, 1>(S, m, up, rtw);
      } else if (sbw == 0) {
        dp_row_step<W, 0>(S, m, up, rtw);
      } else {
        for (int e = (int)sbw - 2; e > 0; --e) dp_row_shift1<W>(S);
        dp_row_step<W, 2>(S, m, up, rtw);
      }
    }
  }
  const int total = sumup + sum_rt;
  const int cost = total - S[0];
  const int k3 = score_k3(cost, total);
  if (mine) {
    A.total[r] = total;
    A.cost[r] = cost;
    A.k3[r] = k3;
  }
  int b = k3;  // warp max -> one atomic (shadow lanes carry a duplicate of a real value)
  for (int o = 16; o; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
  if (lane == 0 && b > 0) atomicMax(best_slot, b);
}

static const int kWindows[] = {8, 12, 16, 20, 24, 28, 32, 36, 40, 44, 48, 52, 56, 60, 64, 72, 80, 96, 112, 128};

int score_fast_window(int width) {
  for (int w : kWindows)
    if (w >= width) return w;
  return 0;
}
bool score_fast_supports(int width) { return score_fast_window(width) != 0; }

size_t score_fast_smem(int W, uint32_t ctx_bytes /* largest region context of the launch */, int lmax) {
  const int nword = (W + 31) / 32;
  return (size_t)ctx_bytes + (size_t)lmax * 256 + (size_t)4 * (lmax + 1) * (3 + nword) * 4;
}

template <int W>
static int launch_w(const ScoreArgs& A, cudaStream_t st) {
  const size_t sm = score_fast_smem(W, A.ctx_max, A.lmax);
  k_score_fast<W><<<(A.n_work + 127) / 128, 128, sm, st>>>(A);
  return 1;
}

void init_score_fast_attrs() {
  int dev = 0, optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const int mx = optin - 2048;  // the limit covers static + dynamic shared memory of a block
  cudaFuncSetAttribute(k_score_fast<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_fast<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_fast<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_fast<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_fast<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_fast<28>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_fast<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_fast<36>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_fast<40>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_fast<44>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_fast<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_fast<52>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_fast<56>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_fast<60>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_fast<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_fast<72>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_fast<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_fast<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_fast<112>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_score_fast<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
}

int launch_score_fast(int W, const ScoreArgs& A, cudaStream_t st) {
  if (A.n_work == 0) return 0;
  switch (W) {
    case 8: return launch_w<8>(A, st);
    case 12: return launch_w<12>(A, st);
    case 16: return launch_w<16>(A, st);
    case 20: return launch_w<20>(A, st);
    case 24: return launch_w<24>(A, st);
    case 28: return launch_w<28>(A, st);
    case 32: return launch_w<32>(A, st);
    case 36: return launch_w<36>(A, st);
    case 40: return launch_w<40>(A, st);
    case 44: return launch_w<44>(A, st);
    case 48: return launch_w<48>(A, st);
    case 52: return launch_w<52>(A, st);
    case 56: return launch_w<56>(A, st);
    case 60: return launch_w<60>(A, st);
    case 64: return launch_w<64>(A, st);
    case 72: return launch_w<72>(A, st);
    case 80: return launch_w<80>(A, st);
    case 96: return launch_w<96>(A, st);
    case 112: return launch_w<112>(A, st);
    case 128: return launch_w<128>(A, st);
    default: return -1;
  }
}

}  // namespace nw
