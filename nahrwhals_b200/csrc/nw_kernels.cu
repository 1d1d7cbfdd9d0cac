// nw_kernels.cu -- sm_100a kernels of the mutation-space search (everything except the register-resident DP,
// which lives in nw_score_fast.cu).  Reference semantics: inst/extdata/scripts/scripts_nw_main/solver.jl.
#include "nw_kernels.cuh"
#include "nw_launch.h"

namespace nw {

// =====================================================================================================
// to_moveset  (solver.jl:217-231, compute_possible_moves :159-165)
// One thread per (a, b): dup_del = any_y aln[a,y]*aln[b,y] == +1 ; invert = any_y product == -1, on bit-planes.
// =====================================================================================================
__global__ void k_moveset(const RegionDev* regions) {
  const RegionDev& R = regions[blockIdx.y];
  const int n = R.n_x;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)n * n) return;
  const int a = (int)(t / n) + 1, b = (int)(t % n) + 1;
  if (a == b) return;
  const uint32_t* planes = reinterpret_cast<const uint32_t*>(R.ctx);
  const uint32_t* pa = planes + (size_t)(a * 2) * R.PWS;
  const uint32_t* pb = planes + (size_t)(b * 2) * R.PWS;
  uint32_t dd = 0, iv = 0;
  for (int w = 0; w < R.PWS; ++w) {
    const uint32_t ap = pa[w], an = pa[R.PWS + w], bp = pb[w], bn = pb[R.PWS + w];
    dd |= (ap & bp) | (an & bn);
    iv |= (ap & bn) | (an & bp);
  }
  if (dd) atomicOr(&R.ms_dd[(size_t)a * R.MW + (b >> 5)], 1u << (b & 31));
  if (iv) atomicOr(&R.ms_inv[(size_t)a * R.MW + (b >> 5)], 1u << (b & 31));
  if (dd | iv) R.ms_any[a] = 1u;
}

// =====================================================================================================
// hash table: exact-dedup stand-in (solver.jl:535 haskey / :544 insert).  First discoverer = min gpos.
// =====================================================================================================
// 128-bit compare-and-swap (SASS: ATOMG.E.CAS.128); addr must be 16-B aligned
__device__ __forceinline__ void cas128(uint64_t* addr, uint64_t cl, uint64_t ch, uint64_t vl, uint64_t vh, uint64_t& ol,
                                       uint64_t& oh) {
  asm volatile(
      "{\n .reg .b128 c, v, o;\n mov.b128 c, {%3, %4};\n mov.b128 v, {%5, %6};\n"
      " atom.global.cas.b128 o, [%2], c, v;\n mov.b128 {%0, %1}, o;\n}"
      : "=l"(ol), "=l"(oh)
      : "l"(addr), "l"(cl), "l"(ch), "l"(vl), "l"(vh)
      : "memory");
}

// Insert the key (k0, kh) with discovery position gpos; returns the slot.  Slot = {k0, kh << 32 | min gpos}, all-ones =
// empty.  Reads come first: a slot's key never changes once set and its position only decreases, so a duplicate of an
// index discovered earlier needs no atomic at all, and a stale read can only cause a redundant compare-and-swap.
__device__ __forceinline__ uint32_t table_insert(uint64_t* T, uint64_t mask, uint64_t k0, uint32_t kh, uint32_t gpos) {
  uint64_t s = fp_slot_hash(k0, kh) & mask;
  const uint64_t mine = ((uint64_t)kh << 32) | gpos;
  while (true) {
    uint64_t* e = T + s * TBL_WORDS;
    const ulonglong2 cur = __ldcg(reinterpret_cast<const ulonglong2*>(e));  // key and position arrive together
    uint64_t o0 = cur.x, o1 = cur.y;
    if ((o0 == TBL_EMPTY) != (o1 == TBL_EMPTY)) continue;  // a 16-byte read torn by a concurrent claim: read again
    if (o0 == TBL_EMPTY) {
      cas128(e, TBL_EMPTY, TBL_EMPTY, k0, mine, o0, o1);
      if (o0 == TBL_EMPTY && o1 == TBL_EMPTY) return (uint32_t)s;  // claimed
    }
    if (o0 == k0 && (uint32_t)(o1 >> 32) == kh) {
      // same index: keep the smaller discovery position (the first discoverer of solver.jl:535 in serial order)
      while ((uint32_t)o1 > gpos) {
        uint64_t n0, n1;
        cas128(e, o0, o1, k0, mine, n0, n1);
        if (n0 == o0 && n1 == o1) break;
        o0 = n0;
        o1 = n1;
      }
      return (uint32_t)s;
    }
    s = (s + 1) & mask;
  }
}

__global__ void k_table_rehash(const uint64_t* oldT, uint64_t old_cap, uint64_t* newT, uint64_t new_mask) {
  for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < old_cap; s += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k0 = oldT[s * TBL_WORDS];
    if (k0 == TBL_EMPTY) continue;
    const uint64_t w1 = oldT[s * TBL_WORDS + 1];
    table_insert(newT, new_mask, k0, (uint32_t)(w1 >> 32), (uint32_t)w1);
  }
}

// insert the roots (region r: global position r) -- solver.jl:596-604.  The root's fingerprint is the last entry of the
// prefix table F' the host wrote into its frontier blob (fill_blob), so this is one table insertion per region.
__global__ void k_insert_roots(const unsigned char* blob, uint32_t pstride, int LS, const int32_t* len, int n_regions,
                               uint64_t* table, uint64_t tmask) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_regions) return;
  const U4* Ft = reinterpret_cast<const U4*>(blob + (size_t)r * pstride + (size_t)LS * 2);
  const U4 f = Ft[len[r]];
  table_insert(table, tmask, fp_key0(f), fp_key_hi32(f), (uint32_t)r);
}

// =====================================================================================================
// expand: one CTA per parent.  Pair loop + retrieve_possible_moves + emission order dup, del, inv
// (solver.jl:607-638, :180-190).  EMIT=false counts, EMIT=true writes the candidate descriptors in discovery
// order.  The parent sequence is staged into shared memory with one TMA bulk copy.
// =====================================================================================================
__device__ __forceinline__ int pair_moves(const RegionDev& R, int a, int b, int dupok, int& can_dd, int& can_inv) {
  const int aa = a < 0 ? -a : a, bb = b < 0 ? -b : b;
  const bool flip = (a ^ b) < 0;
  const bool dd = (__ldg(&R.ms_dd[(size_t)aa * R.MW + (bb >> 5)]) >> (bb & 31)) & 1u;
  const bool iv = (__ldg(&R.ms_inv[(size_t)aa * R.MW + (bb >> 5)]) >> (bb & 31)) & 1u;
  can_dd = flip ? iv : dd;   // :187
  can_inv = flip ? dd : iv;  // :188
  return can_dd * (1 + dupok) + can_inv;
}

template <bool EMIT>
__global__ void __launch_bounds__(128) k_expand_pairs(ExpandArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_warp_tot[4];
  const int parent = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = A.F.len[parent];
  const int LS = A.F.LS;
  const int dupok = A.F.dupok[parent];
  const RegionDev& R = A.F.regions[region_index(A.F, parent)];
  int16_t* seq = reinterpret_cast<int16_t*>(smem);
  const uint32_t blob_bytes = (uint32_t)LS * 2;
  uint32_t* rowcnt = reinterpret_cast<uint32_t*>(smem + blob_bytes);

  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, blob_bytes);
    tma_bulk_g2s(smem, A.F.blob + (size_t)parent * A.F.pstride, blob_bytes, &bar);
  }
  mbar_wait(&bar, 0);

  // ---- pass 1: candidates per row i (warp per row, lanes over j; counts from ballots, no shuffles) ----
  int maxlen = 0;
  const uint32_t lt = (1u << lane) - 1u;
  for (int i = warp + 1; i <= L; i += 4) {
    const int a = seq[i - 1];
    uint32_t n = 0;
    if (__ldg(&R.ms_any[a < 0 ? -a : a])) {  // a block without any partner starts no move (warp-uniform test)
      for (int j0 = i + 1; j0 <= L; j0 += 32) {
        const int j = j0 + lane;
        int cdd = 0, civ = 0;
        if (j <= L) pair_moves(R, a, seq[j - 1], dupok, cdd, civ);
        const uint32_t mdd = __ballot_sync(0xffffffffu, cdd), minv = __ballot_sync(0xffffffffu, civ);
        n += (uint32_t)(1 + dupok) * __popc(mdd) + __popc(minv);
        if (mdd && dupok) maxlen = max(maxlen, L + (j0 + 31 - __clz(mdd) - i));
      }
    }
    if (lane == 0) rowcnt[i - 1] = n;
  }
  __syncthreads();
  // ---- exclusive scan of rowcnt over rows (warp 0) ----
  if (warp == 0) {
    uint32_t carry = 0;
    for (int r0 = 0; r0 < L; r0 += 32) {
      const int r = r0 + lane;
      const uint32_t v = r < L ? rowcnt[r] : 0;
      uint32_t inc = v;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (r < L) rowcnt[r] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) s_warp_tot[0] = carry;
  }
  __syncthreads();
  if (!EMIT) {
    if (tid == 0) A.cnt[parent] = s_warp_tot[0];
    if (lane == 0 && maxlen > 0) atomicMax(A.maxlen, maxlen);  // maxlen is warp-uniform
    return;
  }
  // ---- pass 2: emit in discovery order (i, then j, then dup < del < inv; solver.jl:607-638) ----
  const uint32_t base = A.off[parent];
  for (int i = warp + 1; i <= L; i += 4) {
    const int a = seq[i - 1];
    if (!__ldg(&R.ms_any[a < 0 ? -a : a])) continue;
    uint32_t rowpos = base + rowcnt[i - 1];
    for (int j0 = i + 1; j0 <= L; j0 += 32) {
      const int j = j0 + lane;
      int cdd = 0, civ = 0;
      if (j <= L) pair_moves(R, a, seq[j - 1], dupok, cdd, civ);
      const uint32_t mdd = __ballot_sync(0xffffffffu, cdd), minv = __ballot_sync(0xffffffffu, civ);
      if ((mdd | minv) == 0) continue;
      uint32_t pos = rowpos + (uint32_t)(1 + dupok) * __popc(mdd & lt) + __popc(minv & lt);
      rowpos += (uint32_t)(1 + dupok) * __popc(mdd) + __popc(minv);
      if (cdd && dupok) A.desc[pos++] = pack_desc((uint32_t)parent, i, j, KIND_DUP);
      if (cdd) A.desc[pos++] = pack_desc((uint32_t)parent, i, j, KIND_DEL);
      if (civ) A.desc[pos++] = pack_desc((uint32_t)parent, i, j, KIND_INV);
    }
  }
}

// -----------------------------------------------------------------------------------------------------
// k_expand: same enumeration, one THREAD per row.  Instead of testing all L-i pairs of a row, the row's partner
// positions are assembled as bit masks over positions: occ[s][b] = positions where block b occurs with strand s
// (built once per parent in shared memory); for row i with block a the duplication/deletion partners are
//   OR_{b in dup_del(a)} occ[same strand][b]  |  OR_{b in invert(a)} occ[other strand][b]      (solver.jl:183-188)
// and the inversion partners the mirror image.  Candidates of the row are the set bits above i, in bit order
// (= j order, the discovery order of solver.jl:608).  Work per row ~ #partner blocks instead of L.  The masks
// live in LWT registers per strand class (LWT = words covering the longest parent of the level).
// Shared memory: seq int16[LS] | occ u32[2][n_x+1][LWT] | rowcnt u32[LS].
// -----------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t expand_rows_smem(int LS, int n_x, int LW) {
  return (((size_t)LS * 2 + 15) & ~(size_t)15) + (size_t)2 * (n_x + 1) * LW * 4 + (size_t)LS * 4;
}

// LWT = mask words of the launch (longest parent of the level), LWA <= LWT = mask words of THIS parent's size class:
// the loops over mask words and the occupancy layout `occ[(strand * (n_x + 1) + block) * LWA + word]` follow the parent
// (a 20-block region searched together with 120-block ones pays for 1 word per block, not 16); words >= LWA stay 0.
template <int LWT, int LWA>
__device__ __forceinline__ uint32_t row_masks(const RegionDev& R, int n_x, const int16_t* seq, const uint32_t* occ, int i,
                                              int dupok, uint32_t (&ddm)[LWT], uint32_t (&ivm)[LWT], int& top) {
  const int MW = R.MW;
  const int a = seq[i - 1];
  const int aa = a < 0 ? -a : a, sa = a < 0 ? 1 : 0;
#pragma unroll
  for (int x = 0; x < LWT; ++x) ddm[x] = ivm[x] = 0;
  top = -1;
  if (!__ldg(&R.ms_any[aa])) return 0;
  const uint32_t* same = occ + (size_t)sa * (n_x + 1) * LWA;
  const uint32_t* oppo = occ + (size_t)(1 - sa) * (n_x + 1) * LWA;
  for (int w = 0; w < MW; ++w) {
    uint32_t bits = __ldg(&R.ms_dd[(size_t)aa * MW + w]);  // dup_del table true (solver.jl:184)
    while (bits) {
      const int b = w * 32 + __ffs(bits) - 1;
      bits &= bits - 1;
#pragma unroll
      for (int x = 0; x < LWA; ++x) {
        ddm[x] |= same[(size_t)b * LWA + x];  // not flipped: dup/del
        ivm[x] |= oppo[(size_t)b * LWA + x];  // flipped:     inversion   (:187-188)
      }
    }
    bits = __ldg(&R.ms_inv[(size_t)aa * MW + w]);  // invert table true (:185)
    while (bits) {
      const int b = w * 32 + __ffs(bits) - 1;
      bits &= bits - 1;
#pragma unroll
      for (int x = 0; x < LWA; ++x) {
        ivm[x] |= same[(size_t)b * LWA + x];
        ddm[x] |= oppo[(size_t)b * LWA + x];
      }
    }
  }
  uint32_t n = 0;
#pragma unroll
  for (int x = 0; x < LWA; ++x) {  // keep positions j > i only: bit index t = j-1 >= i
    const int lo = x * 32;
    uint32_t keep = 0xffffffffu;
    if (lo + 32 <= i) keep = 0;
    else if (lo < i) keep = ~((1u << (i - lo)) - 1u);
    ddm[x] &= keep;
    ivm[x] &= keep;
    n += (uint32_t)(1 + dupok) * __popc(ddm[x]) + __popc(ivm[x]);
    if (ddm[x]) top = lo + 31 - __clz(ddm[x]);
  }
  return n;
}

// One parent, mask words of its own size class (LWA <= LWT): see row_masks.
template <bool EMIT, int LWT, int LWA>
__device__ __forceinline__ void expand_parent(const ExpandArgs& A, unsigned char* smem, uint64_t* bar, uint32_t* s_tot,
                                              int* s_maxlen, int parent, int L) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int LS = A.F.LS;
  const int dupok = A.F.dupok[parent];
  const RegionDev& R = A.F.regions[region_index(A.F, parent)];
  const int n_x = R.n_x;  // layout of the occurrence masks: this region's blocks, this parent's mask words
  int16_t* seq = reinterpret_cast<int16_t*>(smem);
  uint32_t* occ = reinterpret_cast<uint32_t*>(smem + (((size_t)LS * 2 + 15) & ~(size_t)15));
  uint32_t* rowcnt = occ + (size_t)2 * (A.nx_max + 1) * LWT;  // fixed by the launch (expand_rows_smem)
  for (int t = tid; t < 2 * (n_x + 1) * LWA; t += blockDim.x) occ[t] = 0;
  mbar_wait(bar, 0);
  __syncthreads();
  for (int t = tid; t < L; t += blockDim.x) {
    const int v = seq[t];
    const int b = v < 0 ? -v : v;
    atomicOr(&occ[((size_t)(v < 0 ? 1 : 0) * (n_x + 1) + b) * LWA + (t >> 5)], 1u << (t & 31));
  }
  __syncthreads();
  // ---- pass 1: candidate count of every row (the masks of a thread's first row stay in registers for pass 2) ----
  int maxlen = 0;
  uint32_t ddm0[LWA], ivm0[LWA];
  uint32_t cnt0 = 0;
  for (int i = tid + 1; i <= L; i += blockDim.x) {
    uint32_t ddm[LWT], ivm[LWT];
    int top;
    const uint32_t cnt = row_masks<LWT, LWA>(R, n_x, seq, occ, i, dupok, ddm, ivm, top);
    rowcnt[i - 1] = cnt;
    if (dupok && top >= 0) maxlen = max(maxlen, L + (top + 1 - i));
    if (EMIT && i == tid + 1) {
      cnt0 = cnt;
#pragma unroll
      for (int x = 0; x < LWA; ++x) {
        ddm0[x] = ddm[x];
        ivm0[x] = ivm[x];
      }
    }
  }
  if (!EMIT && maxlen > 0) atomicMax(s_maxlen, maxlen);
  __syncthreads();
  // ---- exclusive scan of rowcnt over rows (warp 0) ----
  if (warp == 0) {
    uint32_t carry = 0;
    for (int r0 = 0; r0 < L; r0 += 32) {
      const int r = r0 + lane;
      const uint32_t v = r < L ? rowcnt[r] : 0;
      uint32_t inc = v;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (r < L) rowcnt[r] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) *s_tot = carry;
  }
  __syncthreads();
  if (!EMIT) {
    if (tid == 0) {
      A.cnt[parent] = *s_tot;
      if (*s_maxlen > 0) atomicMax(A.maxlen, *s_maxlen);
    }
    return;
  }
  // ---- pass 2: emit in discovery order (i, then j, then dup < del < inv; solver.jl:607-638) ----
  const uint32_t base = A.off[parent];
  for (int i = tid + 1; i <= L; i += blockDim.x) {
    uint32_t ddm[LWT], ivm[LWT];
    if (i == tid + 1) {
      if (cnt0 == 0) continue;
#pragma unroll
      for (int x = 0; x < LWA; ++x) {
        ddm[x] = ddm0[x];
        ivm[x] = ivm0[x];
      }
    } else {  // indices longer than the CTA: the further rows of a thread are recomputed
      int top;
      if (row_masks<LWT, LWA>(R, n_x, seq, occ, i, dupok, ddm, ivm, top) == 0) continue;
    }
    uint32_t pos = base + rowcnt[i - 1];
#pragma unroll
    for (int x = 0; x < LWA; ++x) {
      const uint32_t dd = ddm[x], iv = ivm[x];
      uint32_t any = dd | iv;
      while (any) {
        const int bit = __ffs(any) - 1;
        any &= any - 1;
        const int j = x * 32 + bit + 1;
        if ((dd >> bit) & 1u) {
          if (dupok) A.desc[pos++] = pack_desc((uint32_t)parent, i, j, KIND_DUP);
          A.desc[pos++] = pack_desc((uint32_t)parent, i, j, KIND_DEL);
        }
        if ((iv >> bit) & 1u) A.desc[pos++] = pack_desc((uint32_t)parent, i, j, KIND_INV);
      }
    }
  }
}

template <bool EMIT, int LWT>
__global__ void __launch_bounds__(128) k_expand(ExpandArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tot;
  __shared__ int s_maxlen;
  const int parent = blockIdx.x;
  const int tid = threadIdx.x;
  const int L = A.F.len[parent];
  const uint32_t blob_bytes = (uint32_t)A.F.LS * 2;
  if (tid == 0) {
    mbar_init(&bar, 1);
    s_maxlen = 0;
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, blob_bytes);
    tma_bulk_g2s(smem, A.F.blob + (size_t)parent * A.F.pstride, blob_bytes, &bar);
  }
  // size class of this parent (CTA-uniform): regions of very different sizes share one launch in regionfile mode
  const int lw = (L + 31) >> 5;
  if (LWT > 1 && lw <= 1) expand_parent<EMIT, LWT, 1>(A, smem, &bar, &s_tot, &s_maxlen, parent, L);
  else if (LWT > 2 && lw <= 2) expand_parent<EMIT, LWT, (LWT > 2 ? 2 : LWT)>(A, smem, &bar, &s_tot, &s_maxlen, parent, L);
  else if (LWT > 4 && lw <= 4) expand_parent<EMIT, LWT, (LWT > 4 ? 4 : LWT)>(A, smem, &bar, &s_tot, &s_maxlen, parent, L);
  else if (LWT > 8 && lw <= 8) expand_parent<EMIT, LWT, (LWT > 8 ? 8 : LWT)>(A, smem, &bar, &s_tot, &s_maxlen, parent, L);
  else if (LWT > 16 && lw <= 16) expand_parent<EMIT, LWT, (LWT > 16 ? 16 : LWT)>(A, smem, &bar, &s_tot, &s_maxlen, parent, L);
  else expand_parent<EMIT, LWT, LWT>(A, smem, &bar, &s_tot, &s_maxlen, parent, L);
}

template <int LWT>
static void launch_expand_rows(bool emit, const ExpandArgs& A, cudaStream_t st) {
  const size_t sm = expand_rows_smem(A.F.LS, A.nx_max, LWT);
  if (emit)
    k_expand<true, LWT><<<A.F.P, 128, sm, st>>>(A);
  else
    k_expand<false, LWT><<<A.F.P, 128, sm, st>>>(A);
}

int launch_expand(bool emit, const ExpandArgs& A, cudaStream_t st) {
  if (A.F.P <= 0) return 0;
  const int LW = (A.F.LS + 31) / 32;
  const int LWT = LW <= 1 ? 1 : LW <= 2 ? 2 : LW <= 4 ? 4 : LW <= 8 ? 8 : LW <= 16 ? 16 : LW <= 32 ? 32 : 0;
  if (LWT && !A.force_pairs && expand_rows_smem(A.F.LS, A.nx_max, LWT) <= 160 * 1024) {
    switch (LWT) {
      case 1: launch_expand_rows<1>(emit, A, st); break;
      case 2: launch_expand_rows<2>(emit, A, st); break;
      case 4: launch_expand_rows<4>(emit, A, st); break;
      case 8: launch_expand_rows<8>(emit, A, st); break;
      case 16: launch_expand_rows<16>(emit, A, st); break;
      default: launch_expand_rows<32>(emit, A, st); break;
    }
    return 1;
  }
  // very long indices / very many blocks: pair-testing kernel (the occurrence masks would not fit)
  const size_t rows = (size_t)A.F.LS * 4;
  const size_t sm = (size_t)A.F.LS * 2 + rows;
  if (emit)
    k_expand_pairs<true><<<A.F.P, 128, sm, st>>>(A);
  else
    k_expand_pairs<false><<<A.F.P, 128, sm, st>>>(A);
  return 1;
}

// =====================================================================================================
// dedup insert: one thread per candidate (full lanes, maximal memory-level parallelism for the table atomics).
// The child is fingerprinted through the index map from its parent's prefix tables (read via the read-only path;
// candidates of one parent are adjacent, so the tables stay in L1) -- the stand-in for the exact Dict key of
// solver.jl:535/:544.  The first discoverer of an index is the candidate with the smallest discovery position.
// =====================================================================================================
__global__ void __launch_bounds__(256) k_hash_insert(ExpandArgs A, uint32_t G) {
  const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= G) return;
  const uint64_t d = A.desc[pos];
  const uint32_t ps = desc_parent(d);
  const unsigned char* blob = A.F.blob + (size_t)ps * A.F.pstride;
  const U4* Ft = reinterpret_cast<const U4*>(blob + (size_t)A.F.LS * 2);
  const U4* Gt = Ft + (A.F.LS + 1);
  U4 f = child_fingerprint(Ft, Gt, A.pw, A.F.len[ps], desc_kind(d), desc_i(d), desc_j(d));
  if (A.weak_fp) {  // tests only: provoke collisions to exercise the verifier
    f.v[0] &= 0xffu;
    f.v[1] = f.v[2] = f.v[3] = 0;
  }
  A.slot[pos] = table_insert(A.table, A.tmask, fp_key0(f), fp_key_hi32(f), A.gbase + pos);
}
int launch_hash_insert(const ExpandArgs& A, uint32_t G, cudaStream_t st) {
  if (G == 0) return 0;
  k_hash_insert<<<(G + 255) / 256, 256, 0, st>>>(A, G);
  return 1;
}

// =====================================================================================================
// exclusive scan (uint32) -- three kernels, 4096 items per block
// =====================================================================================================
constexpr int SCAN_T = 256, SCAN_I = 16, SCAN_B = SCAN_T * SCAN_I;

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* sh /*[33]*/, uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint32_t inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) sh[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < nw ? sh[lane] : 0;
    uint32_t wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    sh[lane] = wi - w;
    if (lane == 31) sh[32] = wi;
  }
  __syncthreads();
  const uint32_t r = sh[warp] + inc - v;
  total = sh[32];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_partial(const uint32_t* in, uint64_t n, uint32_t* bsum) {
  __shared__ uint32_t sh[33];
  const uint64_t b0 = (uint64_t)blockIdx.x * SCAN_B;
  uint32_t s = 0;
  for (int k = 0; k < SCAN_I; ++k) {
    const uint64_t idx = b0 + (uint64_t)k * SCAN_T + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  uint32_t tot;
  block_excl_scan(s, sh, tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(1024) k_scan_bsums(uint32_t* bsum, uint32_t nb, uint32_t* total_out) {
  __shared__ uint32_t sh[33];
  uint32_t carry = 0;
  for (uint32_t c0 = 0; c0 < nb; c0 += 1024) {
    const uint32_t idx = c0 + threadIdx.x;
    const uint32_t v = idx < nb ? bsum[idx] : 0;
    uint32_t tot;
    const uint32_t e = block_excl_scan(v, sh, tot);
    if (idx < nb) bsum[idx] = carry + e;
    carry += tot;
  }
  if (threadIdx.x == 0) *total_out = carry;
}
__global__ void __launch_bounds__(1024) k_scan_final(const uint32_t* in, uint32_t* out, uint64_t n, const uint32_t* bsum) {
  __shared__ uint32_t sh[33];
  const uint64_t b0 = (uint64_t)blockIdx.x * SCAN_B + (uint64_t)threadIdx.x * 4;  // one coalesced 16-B load per thread
  uint32_t v[4] = {0, 0, 0, 0};
  if (b0 + 3 < n) {
    const uint4 q = *reinterpret_cast<const uint4*>(in + b0);
    v[0] = q.x;
    v[1] = q.y;
    v[2] = q.z;
    v[3] = q.w;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (b0 + k < n) v[k] = in[b0 + k];
  }
  uint32_t tot;
  uint32_t e = block_excl_scan(v[0] + v[1] + v[2] + v[3], sh, tot) + bsum[blockIdx.x];
  if (b0 + 3 < n) {
    uint4 o;
    o.x = e;
    o.y = e + v[0];
    o.z = e + v[0] + v[1];
    o.w = e + v[0] + v[1] + v[2];
    *reinterpret_cast<uint4*>(out + b0) = o;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (b0 + k < n) out[b0 + k] = e;
      e += v[k];
    }
  }
}

// small inputs (the common case in regionfile mode): one CTA, one launch, coalesced tiles of 1024
constexpr uint32_t SCAN_SMALL = 8192;
__global__ void __launch_bounds__(1024) k_scan_small(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* total_out) {
  __shared__ uint32_t sh[33];
  uint32_t carry = 0;
  for (uint32_t t0 = 0; t0 < n; t0 += 1024) {
    const uint32_t idx = t0 + threadIdx.x;
    const uint32_t v = idx < n ? in[idx] : 0;
    uint32_t tot;
    const uint32_t e = block_excl_scan(v, sh, tot);
    if (idx < n) out[idx] = carry + e;
    carry += tot;
  }
  if (threadIdx.x == 0) *total_out = carry;
}

int launch_scan(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* bsum_scratch, uint32_t* total_out,
                cudaStream_t st) {
  if (n == 0) {
    cudaMemsetAsync(total_out, 0, sizeof(uint32_t), st);
    return 0;
  }
  if (n <= SCAN_SMALL && in != out) {
    k_scan_small<<<1, 1024, 0, st>>>(in, out, (uint32_t)n, total_out);
    return 1;
  }
  const uint32_t nb = (uint32_t)((n + SCAN_B - 1) / SCAN_B);
  k_scan_partial<<<nb, SCAN_T, 0, st>>>(in, n, bsum_scratch);
  k_scan_bsums<<<1, 1024, 0, st>>>(bsum_scratch, nb, total_out);
  k_scan_final<<<nb, 1024, 0, st>>>(in, out, n, bsum_scratch);
  return 3;
}
uint64_t scan_scratch_items(uint64_t n) { return (n + SCAN_B - 1) / SCAN_B + 1; }

// =====================================================================================================
// resolve / assign: which candidate is the first discoverer of its index (solver.jl:535-561)
// =====================================================================================================
// four candidates per thread: four independent random table reads in flight.  A warp covers 128 consecutive
// candidates = four words of the first-discoverer bit mask, assembled from the lanes' nibbles by three shuffles.
__global__ void __launch_bounds__(256) k_resolve(ExpandArgs A, uint32_t G, uint32_t* fresh_mask, uint32_t* wcount,
                                                 uint32_t* owner) {
  const uint32_t pos0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
  uint32_t nib = 0;
  if (pos0 + 3 < G) {
    const uint4 sl = *reinterpret_cast<const uint4*>(A.slot + pos0);
    const uint32_t s4[4] = {sl.x, sl.y, sl.z, sl.w};
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) o[q] = (uint32_t)__ldg(&A.table[(uint64_t)s4[q] * TBL_WORDS + 1]);  // the random reads
    *reinterpret_cast<uint4*>(owner + pos0) = make_uint4(o[0], o[1], o[2], o[3]);
#pragma unroll
    for (int q = 0; q < 4; ++q) nib |= (o[q] == A.gbase + pos0 + q ? 1u : 0u) << q;
  } else {
    for (uint32_t q = 0; q < 4 && pos0 + q < G; ++q) {
      const uint32_t o = (uint32_t)A.table[(uint64_t)A.slot[pos0 + q] * TBL_WORDS + 1];
      owner[pos0 + q] = o;
      nib |= (o == A.gbase + pos0 + q ? 1u : 0u) << q;
    }
  }
  const int lane = threadIdx.x & 31;
  uint32_t w = nib << (4 * (lane & 7));
  w |= __shfl_xor_sync(0xffffffffu, w, 1);
  w |= __shfl_xor_sync(0xffffffffu, w, 2);
  w |= __shfl_xor_sync(0xffffffffu, w, 4);
  if ((lane & 7) == 0 && pos0 < G) {
    fresh_mask[pos0 >> 5] = w;
    wcount[pos0 >> 5] = (uint32_t)__popc(w);
  }
}

__device__ __forceinline__ int32_t owner_id(const LevelTable& LT, uint32_t g) {
  int l = 0;
  while (l + 1 < LT.n_levels && g >= LT.gbase[l + 1]) ++l;
  return LT.cand_id[l][g - LT.gbase[l]];
}

// Four consecutive candidates per thread: the per-candidate records arrive as 16-byte vectors and the four dependent
// chains (parent length / region, owner's id) are in flight together -- the kernel is bound by memory latency.
// Returns the child length of a fresh candidate of this rank that needs a DP (its histogram key), else -1.
__device__ __forceinline__ int assign_one(const AssignArgs& A, uint32_t pos, uint32_t fresh, uint32_t r, uint32_t own,
                                          uint64_t d, int& rid_out) {
  int len_key = -1;
  if (fresh) {
    A.cand_id[pos] = A.id_base + (int32_t)r;
    A.newlist[r] = pos;
    const uint32_t ps = desc_parent(d);
    const int len = child_len(A.front_len[ps], desc_kind(d), desc_i(d), desc_j(d));
    const int rid = A.front_rid ? (int)A.front_rid[ps] : 0;
    rid_out = rid;
    A.new_len[r] = (uint16_t)len;
    A.new_rid[r] = (uint16_t)rid;
    if (!A.force && early_exit(len, A.regions[rid].n_y)) {  // solver.jl:332-334 -> score 0.0, no DP
      A.k3[r] = 0;
      A.cost[r] = -1;
      A.total[r] = 0;
    } else if (owns(A.own, r)) {  // sharded search: another rank scores the rest (exchange after the DP)
      len_key = len;
    }
  } else {
    A.cand_id[pos] = own >= A.gbase ? A.id_base + (int32_t)mask_rank(A.fresh_mask, A.wprefix, own - A.gbase) : owner_id(A.LT, own);
  }
  return len_key;
}
// ONE = a single region: the (length, sub-counter) histogram of the CTA is built with shared-memory atomics and flushed
// once per non-empty bin (the fresh ranks of a CTA's 1024 candidates span at most two sub-counters); several regions:
// warp-aggregated global atomics keyed by (region, length, sub-counter).
template <bool ONE>
__global__ void __launch_bounds__(256) k_assign(AssignArgs A) {
  extern __shared__ uint32_t s_hist[];  // ONE: [2][lcap]
  const uint32_t pos0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
  uint32_t r_cta = 0;
  if (ONE) {
    for (int b = threadIdx.x; b < 2 * A.lcap; b += blockDim.x) s_hist[b] = 0;
    r_cta = A.wprefix[(blockIdx.x * blockDim.x * 4u) >> 5];  // rank of the CTA's first candidate
    __syncthreads();
  }
  int len[4] = {-1, -1, -1, -1}, rid[4] = {0, 0, 0, 0};
  uint32_t rk[4] = {0, 0, 0, 0};
  if (pos0 < A.G) {
    const uint32_t m = A.fresh_mask[pos0 >> 5], base = A.wprefix[pos0 >> 5], bit = pos0 & 31;
    if (pos0 + 3 < A.G) {
      const uint4 ow = *reinterpret_cast<const uint4*>(A.owner + pos0);
      const ulonglong2 d01 = *reinterpret_cast<const ulonglong2*>(A.desc + pos0);
      const ulonglong2 d23 = *reinterpret_cast<const ulonglong2*>(A.desc + pos0 + 2);
      const uint32_t o4[4] = {ow.x, ow.y, ow.z, ow.w};
      const uint64_t d4[4] = {d01.x, d01.y, d23.x, d23.y};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        rk[q] = base + __popc(m & ((1u << (bit + q)) - 1u));
        len[q] = assign_one(A, pos0 + q, (m >> (bit + q)) & 1u, rk[q], o4[q], d4[q], rid[q]);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t pos = pos0 + q;
        if (pos >= A.G) break;
        rk[q] = base + __popc(m & ((1u << (bit + q)) - 1u));
        len[q] = assign_one(A, pos, (m >> (bit + q)) & 1u, rk[q], A.owner[pos], A.desc[pos], rid[q]);
      }
    }
  }
  if (ONE) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (len[q] >= 0) atomicAdd(&s_hist[(((rk[q] >> 10) != (r_cta >> 10)) ? A.lcap : 0) + len[q]], 1u);
    __syncthreads();
    for (int b = threadIdx.x; b < 2 * A.lcap; b += blockDim.x) {
      const uint32_t n = s_hist[b];
      if (n) {
        const int l = b < A.lcap ? b : b - A.lcap;
        const uint32_t rr = b < A.lcap ? r_cta : r_cta + 1024u;  // any rank of that sub-counter's 1024-block
        atomicAdd(&A.hist[bucket_key(0, A.lcap, l, rr)], n);
      }
    }
  } else {
    // warp-aggregated histogram (all 32 lanes converge here)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int key = len[q] >= 0 ? bucket_key(rid[q], A.lcap, len[q], rk[q]) : -1;
      const uint32_t peers = __match_any_sync(0xffffffffu, key);
      if (key >= 0 && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&A.hist[key], (uint32_t)__popc(peers));
    }
  }
}

int launch_assign(const AssignArgs& A, cudaStream_t st) {
  if (A.G == 0) return 0;
  const uint32_t threads = (A.G + 3) / 4;
  if (!A.front_rid && (size_t)A.lcap * 8 <= 40 * 1024)
    k_assign<true><<<(threads + 255) / 256, 256, (size_t)A.lcap * 8, st>>>(A);
  else
    k_assign<false><<<(threads + 255) / 256, 256, 0, st>>>(A);
  return 1;
}
int launch_resolve(const ExpandArgs& A, uint32_t G, uint32_t* fresh_mask, uint32_t* wcount, uint32_t* owner,
                   cudaStream_t st) {
  if (G == 0) return 0;
  k_resolve<<<((G + 3) / 4 + 255) / 256, 256, 0, st>>>(A, G, fresh_mask, wcount, owner);
  return 1;
}
// new-rank of the candidates at positions idx[i] (idx[i] < limit, else fill): region boundaries of a forest level
__global__ void k_gather_rank(const uint32_t* mask, const uint32_t* wprefix, const uint32_t* idx, uint32_t n, uint32_t limit,
                              uint32_t fill, uint32_t* dst) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = idx[i] < limit ? mask_rank(mask, wprefix, idx[i]) : fill;
}
int launch_gather_rank(const uint32_t* mask, const uint32_t* wprefix, const uint32_t* idx, uint32_t n, uint32_t limit,
                       uint32_t fill, uint32_t* dst, cudaStream_t st) {
  if (!n) return 0;
  k_gather_rank<<<(n + 255) / 256, 256, 0, st>>>(mask, wprefix, idx, n, limit, fill, dst);
  return 1;
}

// scatter the DP-needing new candidates into per-(region, length) buckets (each padded to a multiple of 32 lanes).
// Positions inside a bucket come from a cursor per bucket; the few hot cursors are hit once per (CTA, bucket):
// warps aggregate with match_any, the CTA aggregates in a small shared-memory table keyed by bucket.
constexpr int SCAT_SLOTS = 64;
__global__ void __launch_bounds__(1024) k_bucket_scatter(const uint16_t* new_len, const uint16_t* new_rid,
                                                         const RegionDev* regions, int lcap, uint32_t n_new, int force,
                                                         OwnRange own, const uint32_t* bstart, uint32_t* cursor,
                                                         uint32_t* work) {
  __shared__ int s_key[SCAT_SLOTS];
  __shared__ uint32_t s_cnt[SCAT_SLOTS], s_base[SCAT_SLOTS];
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < SCAT_SLOTS) {
    s_key[tid] = -1;
    s_cnt[tid] = 0;
  }
  __syncthreads();
  const uint32_t r = blockIdx.x * blockDim.x + tid;
  int key = -1;
  if (r < n_new && owns(own, r)) {
    const int len = new_len[r], rid = new_rid ? new_rid[r] : 0;
    if (force || !early_exit(len, regions[rid].n_y)) key = bucket_key(rid, lcap, len, r);
  }
  const uint32_t peers = __match_any_sync(0xffffffffu, key);
  const int leader = __ffs(peers) - 1;
  const uint32_t mine = __popc(peers & ((1u << lane) - 1u));
  int slot = -1;
  uint32_t wbase = 0;
  if (key >= 0 && lane == leader) {
    int h = (int)(((uint32_t)key * 2654435761u) >> 26);  // 6 bits
    for (int probe = 0; probe < SCAT_SLOTS; ++probe) {
      const int old = atomicCAS(&s_key[h], -1, key);
      if (old == -1 || old == key) {
        slot = h;
        break;
      }
      h = (h + 1) & (SCAT_SLOTS - 1);
    }
    if (slot >= 0) wbase = atomicAdd(&s_cnt[slot], (uint32_t)__popc(peers));
  }
  __syncthreads();
  if (tid < SCAT_SLOTS && s_key[tid] >= 0) s_base[tid] = atomicAdd(&cursor[s_key[tid]], s_cnt[tid]);
  __syncthreads();
  // more distinct buckets in this CTA than table slots: that warp goes to the global cursor directly
  if (key >= 0 && lane == leader) wbase = slot >= 0 ? s_base[slot] + wbase : atomicAdd(&cursor[key], (uint32_t)__popc(peers));
  wbase = __shfl_sync(0xffffffffu, wbase, leader);
  if (key >= 0) work[bstart[key] + wbase + mine] = r;
}
// a single region: the CTA's 1024 consecutive new-ranks share one sub-counter; a shared-memory counter per length hands
// every candidate its place inside the CTA's share of the bucket, one global cursor atomic per (CTA, non-empty length)
__global__ void __launch_bounds__(1024) k_bucket_scatter_one(const uint16_t* new_len, const RegionDev* regions, int lcap,
                                                             uint32_t n_new, int force, OwnRange own, const uint32_t* bstart,
                                                             uint32_t* cursor, uint32_t* work) {
  extern __shared__ uint32_t s_cb[];  // [lcap] counts, then [lcap] bases
  uint32_t* s_cnt = s_cb;
  uint32_t* s_base = s_cb + lcap;
  for (int b = threadIdx.x; b < lcap; b += blockDim.x) s_cnt[b] = 0;
  __syncthreads();
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  int len = -1;
  uint32_t local = 0;
  if (r < n_new && owns(own, r)) {
    const int l = new_len[r];
    if (force || !early_exit(l, regions[0].n_y)) {
      len = l;
      local = atomicAdd(&s_cnt[l], 1u);
    }
  }
  __syncthreads();
  const uint32_t r_cta = blockIdx.x * blockDim.x;
  for (int b = threadIdx.x; b < lcap; b += blockDim.x)
    if (s_cnt[b]) s_base[b] = atomicAdd(&cursor[bucket_key(0, lcap, b, r_cta)], s_cnt[b]);
  __syncthreads();
  if (len >= 0) work[bstart[bucket_key(0, lcap, len, r_cta)] + s_base[len] + local] = r;
}
int launch_bucket_scatter(const uint16_t* new_len, const uint16_t* new_rid, const RegionDev* regions, int lcap,
                          uint32_t n_new, int force, const OwnRange& own, const uint32_t* bstart, uint32_t* cursor,
                          uint32_t* work, cudaStream_t st) {
  if (n_new == 0) return 0;
  if (!new_rid && (size_t)lcap * 8 <= 40 * 1024)
    k_bucket_scatter_one<<<(n_new + 1023) / 1024, 1024, (size_t)lcap * 8, st>>>(new_len, regions, lcap, n_new, force, own,
                                                                                bstart, cursor, work);
  else
    k_bucket_scatter<<<(n_new + 1023) / 1024, 1024, 0, st>>>(new_len, new_rid, regions, lcap, n_new, force, own, bstart,
                                                            cursor, work);
  return 1;
}
// best score of a level over ALL fresh candidates (sharded search: the DP kernels only saw this rank's share)
__global__ void k_max_k3(const int32_t* k3, uint32_t n, int* best) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  int b = r < n ? k3[r] : 0;
  for (int o = 16; o; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
  if ((threadIdx.x & 31) == 0 && b > 0) atomicMax(best, b);
}
int launch_max_k3(const int32_t* k3, uint32_t n, int* best, cudaStream_t st) {
  if (!n) return 0;
  k_max_k3<<<(n + 255) / 256, 256, 0, st>>>(k3, n, best);
  return 1;
}

// ---- bucket bookkeeping on the device: the host only ever sees the non-empty (region, length) bins ----
static_assert(NSUB == 8, "bin kernels read the sub-counters of a bin as two uint4");
// histogram of (region, child length, sub-counter) from explicit lengths (roots, nw_score_batch)
__global__ void k_len_hist(const uint16_t* new_len, const uint16_t* new_rid, const RegionDev* regions, int lcap, uint32_t n,
                           int force, uint32_t* hist) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int len = new_len[r], rid = new_rid ? new_rid[r] : 0;
  if (force || !early_exit(len, regions[rid].n_y)) atomicAdd(&hist[bucket_key(rid, lcap, len, r)], 1u);
}
__global__ void k_bin_flags(const uint32_t* hist, uint32_t nb, uint32_t* flag, uint32_t* nsum) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const uint4* h = reinterpret_cast<const uint4*>(hist + (size_t)i * NSUB);
  const uint4 a = h[0], b = h[1];
  const uint32_t n = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
  nsum[i] = n;
  flag[i] = n ? 1u : 0u;
}
__global__ void k_bin_compact(const uint32_t* flag, const uint32_t* rank, const uint32_t* nsum, uint32_t nb, BinRec* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb || !flag[i]) return;
  out[rank[i]] = BinRec{i, nsum[i]};
}
// ent[k] = {bin, first work slot of the bin}: start of every sub-counter's run, and its cursor
__global__ void k_bucket_fill(const BinRec* ent, uint32_t n, const uint32_t* hist, uint32_t* bstart, uint32_t* cursor) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const size_t k0 = (size_t)ent[k].bin * NSUB;
  uint32_t at = ent[k].n;
#pragma unroll
  for (int sb = 0; sb < NSUB; ++sb) {
    bstart[k0 + sb] = at;
    cursor[k0 + sb] = 0;
    at += hist[k0 + sb];
  }
}
// small histograms (single search: one region): flags + scan + compaction in ONE block; out[0] = {count, 0},
// records from out[1] on, in (region, length) order
__global__ void __launch_bounds__(1024) k_bins_small(const uint32_t* hist, uint32_t nb, BinRec* out) {
  __shared__ uint32_t s_warp[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t n[8];
  uint32_t cnt = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint32_t i = (uint32_t)tid * 8 + q;
    n[q] = 0;
    if (i < nb) {
      const uint4* h = reinterpret_cast<const uint4*>(hist + (size_t)i * NSUB);
      const uint4 a = h[0], b = h[1];
      n[q] = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    }
    cnt += n[q] ? 1u : 0u;
  }
  uint32_t incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_warp[lane], wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += v;
    }
    s_warp[lane] = wi - w;  // exclusive
    if (lane == 31) out[0] = BinRec{wi, 0u};
  }
  __syncthreads();
  uint32_t at = s_warp[warp] + incl - cnt;
#pragma unroll
  for (int q = 0; q < 8; ++q)
    if (n[q]) out[1 + at++] = BinRec{(uint32_t)tid * 8 + q, n[q]};
}
int launch_bins_small(const uint32_t* hist, uint32_t nb, BinRec* out, cudaStream_t st) {
  k_bins_small<<<1, 1024, 0, st>>>(hist, nb, out);
  return 1;
}
int launch_len_hist(const uint16_t* new_len, const uint16_t* new_rid, const RegionDev* regions, int lcap, uint32_t n,
                    int force, uint32_t* hist, cudaStream_t st) {
  if (!n) return 0;
  k_len_hist<<<(n + 255) / 256, 256, 0, st>>>(new_len, new_rid, regions, lcap, n, force, hist);
  return 1;
}
int launch_bin_flags(const uint32_t* hist, uint32_t nb, uint32_t* flag, uint32_t* nsum, cudaStream_t st) {
  if (!nb) return 0;
  k_bin_flags<<<(nb + 255) / 256, 256, 0, st>>>(hist, nb, flag, nsum);
  return 1;
}
int launch_bin_compact(const uint32_t* flag, const uint32_t* rank, const uint32_t* nsum, uint32_t nb, BinRec* out,
                       cudaStream_t st) {
  if (!nb) return 0;
  k_bin_compact<<<(nb + 255) / 256, 256, 0, st>>>(flag, rank, nsum, nb, out);
  return 1;
}
int launch_bucket_fill(const BinRec* ent, uint32_t n, const uint32_t* hist, uint32_t* bstart, uint32_t* cursor,
                       cudaStream_t st) {
  if (!n) return 0;
  k_bucket_fill<<<(n + 255) / 256, 256, 0, st>>>(ent, n, hist, bstart, cursor);
  return 1;
}

// =====================================================================================================
// generic DP: literal cost-form transcription of slow_score (solver.jl:362-423) in int32 units, any geometry.
// Thread per candidate; the two DP rows live in a coalesced global scratch [2][n_y+2][threads].
// =====================================================================================================
__device__ __forceinline__ int sat_add(int a, int b) { return a >= DP_INF ? DP_INF : min(a + b, DP_INF); }

__global__ void __launch_bounds__(128) k_score_generic(ScoreArgs A) {
  const uint32_t T = gridDim.x * blockDim.x;
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  int32_t* row0 = A.scratch + tid;
  const size_t rstride = (size_t)(A.ny_max + 2) * T;
  unsigned long long cells = 0;
  for (uint32_t w = tid; w < A.n_work; w += T) {
    const uint32_t r = A.work[w];
    if (r == WORK_PAD) continue;
    const uint64_t d = A.desc[A.newlist[r]];
    const uint32_t ps = desc_parent(d);
    const int rid = region_index(A.F, ps);
    const RegionDev& R = A.F.regions[rid];
    const int ny = R.n_y;
    const uint32_t* planes = reinterpret_cast<const uint32_t*>(R.ctx);
    const int32_t* upx = reinterpret_cast<const int32_t*>(R.ctx + R.off_upx);
    const int kind = desc_kind(d), p = desc_i(d), q = desc_j(d);
    const int L = A.F.len[ps];
    const int16_t* seq = reinterpret_cast<const int16_t*>(A.F.blob + (size_t)ps * A.F.pstride);
    const int Lc = child_len(L, kind, p, q);
    const int16_t* ab = R.band_ab + (size_t)R.band_off[Lc] * 2;
    int cur = 0;
    int cumU = 0;
    for (int i = 1; i <= Lc; ++i) {
      const int c = child_elem(seq, kind, p, q, i);
      const int x = c < 0 ? -c : c;
      const uint32_t* pl = planes + (size_t)(x * 2 + (c < 0 ? 1 : 0)) * R.PWS;
      const int up = upx[2 * x];
      const int a = ab[i * 2], b = ab[i * 2 + 1];
      int32_t* rc = row0 + (size_t)cur * rstride;
      const int32_t* rp = row0 + (size_t)(cur ^ 1) * rstride;
      int left = DP_INF;  // C[i][j-1]
      for (int j = 1; j <= ny; ++j) {
        int v = DP_INF;
        if (j >= a && j <= b) {
          const int rr = R.NYP - 1 - j;
          const bool match = (pl[rr >> 5] >> (rr & 31)) & 1u;
          const int dcost = match ? 0 : DP_INF;
          const int rt = R.rt[j];
          if (i == 1 && j == 1) {
            v = min(up + rt, dcost);  // :362
          } else if (j == 1) {
            v = min(sat_add(rp[(size_t)1 * T], up), sat_add(dcost, cumU));  // :393 (cumU = cumsum_u[i-1])
          } else if (i == 1) {
            v = min(sat_add(left, rt), sat_add(dcost, R.cumr[j - 1]));  // :403
          } else {
            v = min(sat_add(rp[(size_t)(j - 1) * T], dcost),
                    min(sat_add(rp[(size_t)j * T], up), sat_add(left, rt)));  // :420-422
          }
          ++cells;
        }
        rc[(size_t)j * T] = v;
        left = v;
      }
      cumU += up;
      cur ^= 1;
    }
    const int32_t* rl = row0 + (size_t)(cur ^ 1) * rstride;
    const int fin = rl[(size_t)ny * T];
    const int total = cumU + R.sum_rt;
    A.total[r] = total;
    if (fin >= DP_INF) {
      A.cost[r] = -2;
      A.k3[r] = K3_NEGINF;
    } else {
      A.cost[r] = fin;
      const int k3 = score_k3(fin, total);
      A.k3[r] = k3;
      if (k3 > 0) atomicMax(&A.best_k3[rid], k3);
    }
  }
  if (cells && A.cells) atomicAdd(A.cells, cells);
}

int generic_threads() { return 148 * 4 * 128; }
int launch_score_generic(const ScoreArgs& A, cudaStream_t st) {
  if (A.n_work == 0) return 0;
  k_score_generic<<<148 * 4, 128, 0, st>>>(A);
  return 1;
}

// =====================================================================================================
// frontier selection (solver.jl:551-553 push if score >= 0 ; :669-672 stable sort by -score, keep max_count)
// =====================================================================================================
__global__ void k_valid_flags(const int32_t* k3, uint32_t n, uint32_t* flag) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) flag[r] = k3[r] >= 0 ? 1u : 0u;
}
__global__ void k_compact_ranks(const uint32_t* flag, const uint32_t* pre, uint32_t n, uint32_t* sel) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n && flag[r]) sel[pre[r]] = r;
}
// sort key: region (16 bits) | 100000 - k3 (17 bits) | discovery rank (31 bits).  Ascending key == region-major,
// then descending Float32 score, ties by discovery rank (the stable order of solver.jl:670).
__global__ void k_make_keys(const int32_t* k3, const uint16_t* new_rid, const uint32_t* flag, const uint32_t* pre, uint32_t n,
                            uint64_t* keys) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n && flag[r])
    keys[pre[r]] = ((uint64_t)(new_rid ? new_rid[r] : 0) << 48) | ((uint64_t)(uint32_t)(100000 - k3[r]) << 31) | r;
}
// first take[rid] keys of every region's sorted segment -> sel (new-ranks in next-frontier order)
__global__ void k_take_sorted(const uint64_t* keys, uint32_t n, const uint32_t* seg_start, const uint32_t* take,
                              const uint32_t* out_start, uint32_t* sel) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t key = keys[i];
  const uint32_t rid = (uint32_t)(key >> 48);
  const uint32_t j = i - seg_start[rid];
  if (j < take[rid]) sel[out_start[rid] + j] = (uint32_t)(key & 0x7fffffffu);
}
// dst[i] = src[idx[i]] (idx[i] < limit) else fill   -- region boundaries of the scanned per-candidate arrays
__global__ void k_gather_u32(const uint32_t* src, const uint32_t* idx, uint32_t n, uint32_t limit, uint32_t fill,
                             uint32_t* dst) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = idx[i] < limit ? src[idx[i]] : fill;
}
__global__ void k_fill_u64(uint64_t* p, uint64_t n0, uint64_t n1, uint64_t v) {
  const uint64_t i = n0 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n1) p[i] = v;
}
// bitonic sort: local (shared memory, 2048 keys / CTA) + global steps
constexpr int BIT_CHUNK = 2048;
__global__ void __launch_bounds__(1024) k_bitonic_local(uint64_t* keys, uint32_t kstart) {
  // kstart == 0: full local sort of each chunk (k = 2..BIT_CHUNK); otherwise finish stage k=kstart for j < BIT_CHUNK
  __shared__ uint64_t s[BIT_CHUNK];
  const uint64_t base = (uint64_t)blockIdx.x * BIT_CHUNK;
  s[threadIdx.x] = keys[base + threadIdx.x];
  s[threadIdx.x + 1024] = keys[base + threadIdx.x + 1024];
  __syncthreads();
  const uint32_t k0 = kstart ? kstart : 2;
  const uint32_t k1 = kstart ? kstart : BIT_CHUNK;
  for (uint32_t k = k0; k <= k1; k <<= 1) {
    for (uint32_t j = min(k >> 1, (uint32_t)BIT_CHUNK >> 1); j > 0; j >>= 1) {
      const uint32_t t = threadIdx.x;
      const uint32_t lo = ((t / j) * 2 * j) + (t % j);
      const uint32_t hi = lo + j;
      const uint64_t gi = base + lo;
      const bool asc = ((gi & k) == 0);
      const uint64_t a = s[lo], b = s[hi];
      if ((a > b) == asc) {
        s[lo] = b;
        s[hi] = a;
      }
      __syncthreads();
    }
  }
  keys[base + threadIdx.x] = s[threadIdx.x];
  keys[base + threadIdx.x + 1024] = s[threadIdx.x + 1024];
}
__global__ void k_bitonic_global(uint64_t* keys, uint64_t n, uint32_t j, uint32_t k) {
  const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (t >= n / 2) return;
  const uint64_t lo = ((t / j) * 2 * j) + (t % j);
  const uint64_t hi = lo + j;
  const bool asc = ((lo & k) == 0);
  const uint64_t a = keys[lo], b = keys[hi];
  if ((a > b) == asc) {
    keys[lo] = b;
    keys[hi] = a;
  }
}
// sorts keys[0..npad) ascending; npad power of two >= BIT_CHUNK
int launch_sort_u64(uint64_t* keys, uint64_t npad, cudaStream_t st) {
  int launches = 0;
  k_bitonic_local<<<(uint32_t)(npad / BIT_CHUNK), 1024, 0, st>>>(keys, 0);
  ++launches;
  for (uint64_t k = (uint64_t)BIT_CHUNK * 2; k <= npad; k <<= 1) {
    for (uint64_t j = k >> 1; j >= BIT_CHUNK; j >>= 1) {
      k_bitonic_global<<<(uint32_t)((npad / 2 + 255) / 256), 256, 0, st>>>(keys, npad, (uint32_t)j, (uint32_t)k);
      ++launches;
    }
    k_bitonic_local<<<(uint32_t)(npad / BIT_CHUNK), 1024, 0, st>>>(keys, (uint32_t)k);
    ++launches;
  }
  return launches;
}
__global__ void k_keys_to_sel(const uint64_t* keys, uint32_t n, uint32_t* sel) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) sel[t] = (uint32_t)(keys[t] & 0xffffffffu);
}

__global__ void k_k3_hist(const int32_t* k3, uint32_t n, int k3_floor, uint32_t* hist);
// ---- top-w of one region without sorting everything: the w-th best k3 from a histogram, then only the w winners are
// sorted.  The stable order of solver.jl:669-672 is (score descending, discovery rank ascending), so the frontier is
// every candidate above the threshold score plus the first `quota` candidates (in discovery order) that tie with it.
constexpr int TOPK_CH = 98;  // bins per thread of the single-CTA threshold search (1024 x 98 >= 100001)
// thr[0] = threshold k3, thr[1] = quota of the tie bin, thr[2] = candidates above the threshold, thr[3] = 0 (append cursor)
__global__ void __launch_bounds__(1024) k_topk_threshold(const uint32_t* __restrict__ hist, uint32_t w, uint32_t* __restrict__ thr) {
  __shared__ uint32_t sh[33];
  const int t = threadIdx.x;
  const int hi = 100000 - t * TOPK_CH;  // this thread's bins, descending: hi .. max(hi - TOPK_CH + 1, 0)
  uint32_t sum = 0;
#pragma unroll 14
  for (int i = 0; i < TOPK_CH; ++i) {  // branch-free so that the 98 independent loads overlap
    const int b = hi - i;
    const uint32_t h = hist[b > 0 ? b : 0];
    sum += b >= 0 ? h : 0u;
  }
  uint32_t total;
  const uint32_t before = block_excl_scan(sum, sh, total);
  if (before < w && w <= before + sum) {
    uint32_t cum = before;
    for (int b = hi; b > hi - TOPK_CH && b >= 0; --b) {
      const uint32_t h = hist[b];
      if (cum + h >= w) {
        thr[0] = (uint32_t)b;
        thr[1] = w - cum;
        thr[2] = cum;
        thr[3] = 0;
        break;
      }
      cum += h;
    }
  }
}
__global__ void k_topk_tie_flags(const int32_t* __restrict__ k3, uint32_t n, const uint32_t* __restrict__ thr,
                                 uint32_t* __restrict__ flag) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) flag[r] = k3[r] == (int32_t)thr[0] ? 1u : 0u;
}
// keys of the w winners (same key as k_make_keys, region 0) + padding up to npad; winners above the threshold are
// appended in any order (the sort orders them), the ties take their discovery rank inside the tie bin
__global__ void k_topk_keys(const int32_t* __restrict__ k3, uint32_t n, uint32_t* __restrict__ thr,
                            const uint32_t* __restrict__ tie_rank, uint32_t w, uint64_t* __restrict__ keys, uint32_t npad) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < npad - w) keys[w + r] = ~0ull;
  if (r >= n) return;
  const int32_t v = k3[r], T = (int32_t)thr[0];
  if (v < T) return;
  uint32_t pos;
  if (v > T) {
    pos = atomicAdd(&thr[3], 1u);
  } else {
    if (tie_rank[r] >= thr[1]) return;
    pos = thr[2] + tie_rank[r];
  }
  keys[pos] = ((uint64_t)(uint32_t)(100000 - v) << 31) | r;
}
// The same threshold search over many CTAs: CTA b sums the 1024 bins [100000 - 1024 b - 1023, 100000 - 1024 b]; the CTA
// that finishes last walks the 98 block sums from the best score down, then the 1024 bins of the block in which the
// running count crosses w.  thr[4] counts finished CTAs and bsum[] holds the block sums (both zeroed by the caller).
constexpr int TOPK_BLOCKS = 98;  // 98 x 1024 >= 100001
__global__ void __launch_bounds__(256) k_topk_threshold_mc(const uint32_t* __restrict__ hist, uint32_t w, uint32_t* __restrict__ thr,
                                                           uint32_t* __restrict__ bsum) {
  __shared__ uint32_t sh[33];
  __shared__ int s_last;
  const int t = threadIdx.x;
  auto block_sum = [&](int b, uint32_t (&v)[4]) {  // thread t holds bins hi - 4t .. hi - 4t - 3 of block b (descending)
    const int hi = 100000 - b * 1024 - t * 4;
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int bin = hi - i;
      v[i] = bin >= 0 ? hist[bin] : 0u;
      sum += v[i];
    }
    return sum;
  };
  uint32_t v[4];
  uint32_t total;
  const uint32_t mine = block_sum(blockIdx.x, v);
  block_excl_scan(mine, sh, total);
  if (t == 0) {
    bsum[blockIdx.x] = total;
    __threadfence();
    s_last = atomicAdd(&thr[4], 1u) == (uint32_t)gridDim.x - 1 ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  __shared__ uint32_t s_blk, s_before;
  if (t == 0) {
    uint32_t cum = 0;
    int b = 0;
    for (; b < (int)gridDim.x; ++b) {
      const uint32_t x = __ldcg(&bsum[b]);
      if (cum + x >= w) break;
      cum += x;
    }
    s_blk = (uint32_t)b;
    s_before = cum;
  }
  __syncthreads();
  if (s_blk >= gridDim.x) return;  // fewer than w candidates: the caller never asks for that
  const uint32_t sum = block_sum((int)s_blk, v);
  const uint32_t before = s_before + block_excl_scan(sum, sh, total);
  if (before < w && w <= before + sum) {
    uint32_t cum = before;
    const int hi = 100000 - (int)s_blk * 1024 - t * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (cum + v[i] >= w) {
        thr[0] = (uint32_t)(hi - i);
        thr[1] = w - cum;
        thr[2] = cum;
        thr[3] = 0;
        break;
      }
      cum += v[i];
    }
  }
}
// Order of up to RANK_MAX unique keys by counting: every thread finds how many keys are smaller than its own (all keys in
// shared memory) and writes its payload (the low 31 bits: the discovery rank) at that place -- a sort of the beam's few
// winners without the 66 barrier-separated steps of a single-CTA bitonic network.
constexpr uint32_t RANK_MAX = 4096;
__global__ void __launch_bounds__(256) k_rank_select(const uint64_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ sel) {
  extern __shared__ uint64_t s_keys[];
  for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) s_keys[j] = keys[j];
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t k = s_keys[i];
  uint32_t rank = 0;
#pragma unroll 8
  for (uint32_t j = 0; j < n; ++j) rank += s_keys[j] < k ? 1u : 0u;
  sel[rank] = (uint32_t)(k & 0x7fffffffu);
}
int launch_rank_select(const uint64_t* keys, uint32_t n, uint32_t* sel, cudaStream_t st) {
  if (!n) return 0;
  k_rank_select<<<(n + 255) / 256, 256, (size_t)n * 8, st>>>(keys, n, sel);
  return 1;
}
bool rank_select_fits(uint64_t n) { return n <= RANK_MAX; }

// keys[0..w) <- the w best of k3[0..n) (unordered; ascending key == frontier order), keys[w..npad) padding; needs more than
// w valid candidates.  hist: 100008 words (the last 7 are thr), flag / tie_rank: n words
int launch_topk_keys(const int32_t* k3, uint32_t n, uint32_t w, uint32_t* hist, uint32_t* thr, uint32_t* flag, uint32_t* tie_rank,
                     uint32_t* bsum_scratch, uint32_t* total_out, uint64_t* keys, uint32_t npad, cudaStream_t st) {
  int l = 0;
  cudaMemsetAsync(hist, 0, (size_t)100008 * 4, st);
  k_k3_hist<<<(n + 255) / 256, 256, 0, st>>>(k3, n, 0, hist);
  k_topk_threshold_mc<<<TOPK_BLOCKS, 256, 0, st>>>(hist, w, thr, bsum_scratch);
  k_topk_tie_flags<<<(n + 255) / 256, 256, 0, st>>>(k3, n, thr, flag);
  l += 3;
  l += launch_scan(flag, tie_rank, n, bsum_scratch, total_out, st);
  const uint32_t cover = n > npad ? n : npad;
  k_topk_keys<<<(cover + 255) / 256, 256, 0, st>>>(k3, n, thr, tie_rank, w, keys, npad);
  ++l;
  return l;
}

int launch_valid_flags(const int32_t* k3, uint32_t n, uint32_t* flag, cudaStream_t st) {
  if (!n) return 0;
  k_valid_flags<<<(n + 255) / 256, 256, 0, st>>>(k3, n, flag);
  return 1;
}
int launch_compact_ranks(const uint32_t* flag, const uint32_t* pre, uint32_t n, uint32_t* sel, cudaStream_t st) {
  if (!n) return 0;
  k_compact_ranks<<<(n + 255) / 256, 256, 0, st>>>(flag, pre, n, sel);
  return 1;
}
int launch_take_sorted(const uint64_t* keys, uint32_t n, const uint32_t* seg_start, const uint32_t* take,
                       const uint32_t* out_start, uint32_t* sel, cudaStream_t st) {
  if (!n) return 0;
  k_take_sorted<<<(n + 255) / 256, 256, 0, st>>>(keys, n, seg_start, take, out_start, sel);
  return 1;
}
int launch_gather_u32(const uint32_t* src, const uint32_t* idx, uint32_t n, uint32_t limit, uint32_t fill, uint32_t* dst,
                      cudaStream_t st) {
  if (!n) return 0;
  k_gather_u32<<<(n + 255) / 256, 256, 0, st>>>(src, idx, n, limit, fill, dst);
  return 1;
}
int launch_make_keys(const int32_t* k3, const uint16_t* new_rid, const uint32_t* flag, const uint32_t* pre, uint32_t n,
                     uint64_t* keys, uint64_t nvalid, uint64_t npad, cudaStream_t st) {
  int l = 0;
  if (npad > nvalid) {
    k_fill_u64<<<(uint32_t)((npad - nvalid + 255) / 256), 256, 0, st>>>(keys, nvalid, npad, ~0ull);
    ++l;
  }
  if (n) {
    k_make_keys<<<(n + 255) / 256, 256, 0, st>>>(k3, new_rid, flag, pre, n, keys);
    ++l;
  }
  return l;
}
int launch_keys_to_sel(const uint64_t* keys, uint32_t n, uint32_t* sel, cudaStream_t st) {
  if (!n) return 0;
  k_keys_to_sel<<<(n + 255) / 256, 256, 0, st>>>(keys, n, sel);
  return 1;
}

// =====================================================================================================
// materialise the next frontier: child sequences (apply_* :244-278 through the index map), duplication_count
// guard of the *next* expansion (:292-298, :611) and the prefix-hash tables used by k_expand.
// One CTA (256 threads = 8 warps: one per (base, F|G) table) per selected child.
// =====================================================================================================
__global__ void __launch_bounds__(256) k_materialise(MaterialiseArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  int16_t* cs = reinterpret_cast<int16_t*>(smem);                                        // [LS2]
  int* cntx = reinterpret_cast<int*>(smem + (((size_t)A.LS2 * 2 + 15) & ~(size_t)15));  // [n_x+1]
  __shared__ int s_max;
  const int e = blockIdx.x;
  const uint32_t r = A.sel[e];
  const uint64_t d = A.newlist ? A.desc[A.newlist[r]] : A.desc[r];
  const uint32_t ps = desc_parent(d);
  const int kind = desc_kind(d), p = desc_i(d), q = desc_j(d);
  const int L = A.F.len[ps];
  const int16_t* seq = reinterpret_cast<const int16_t*>(A.F.blob + (size_t)ps * A.F.pstride);
  const int Lc = child_len(L, kind, p, q);
  const int rid = region_index(A.F, ps);
  const uint32_t rid1 = A.F.regions[rid].rid1;
  unsigned char* out = A.blob2 + (size_t)e * A.pstride2;
  int16_t* oseq = reinterpret_cast<int16_t*>(out);
  if (threadIdx.x == 0) s_max = 0;
  for (int x = threadIdx.x; x <= A.nx_max; x += blockDim.x) cntx[x] = 0;
  __syncthreads();
  for (int t = threadIdx.x; t < A.LS2; t += blockDim.x) {
    int c = 0;
    if (t < Lc) {
      c = child_elem(seq, kind, p, q, t + 1);
      const int m = atomicAdd(&cntx[c < 0 ? -c : c], 1) + 1;
      atomicMax(&s_max, m);
    }
    cs[t] = (int16_t)c;
    oseq[t] = (int16_t)c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    A.len2[e] = Lc;
    A.id2[e] = A.id_base + (int32_t)r;
    A.dupok2[e] = s_max <= A.maxdup ? 1 : 0;
    A.rid2[e] = (uint16_t)rid;
  }
  // prefix tables: warp w -> base (w & 3), table (w >> 2): 0 = F (B^t), 1 = G (B^-t)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = warp & 3, which = warp >> 2;
  U4* tab = reinterpret_cast<U4*>(out + (size_t)A.LS2 * 2) + (which ? (A.LS2 + 1) : 0);
  // F'[n] = rid + B * sum_{t<n} val B^t ;  G'[n] = B * sum_{t<n} val B^-t  (see child_fingerprint)
  uint32_t carry = which ? 0u : rid1;
  if (lane == 0) tab[0].v[k] = carry;
  for (int t0 = 0; t0 < Lc; t0 += 32) {
    const int t = t0 + lane;
    uint32_t term = 0;
    if (t < Lc) term = mulmod(elem_val(cs[t]), A.pw[FP_EMAX + (which ? 1 - t : t + 1)].v[k]);
    uint32_t inc = term;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc = addmod(inc, u);
    }
    if (t < Lc) tab[t + 1].v[k] = addmod(carry, inc);
    carry = addmod(carry, __shfl_sync(0xffffffffu, inc, 31));
  }
  // entries beyond Lc are never read (k_expand indexes F[<=L], G[<=L])
}

int launch_materialise(const MaterialiseArgs& A, cudaStream_t st) {
  if (A.P2 <= 0) return 0;
  const size_t sm = (((size_t)A.LS2 * 2 + 15) & ~(size_t)15) + (size_t)(A.nx_max + 1) * 4;
  k_materialise<<<A.P2, 256, sm, st>>>(A);
  return 1;
}

// =====================================================================================================
// report extraction: which ids are good (solver.jl:733), which ancestors/edges are needed (:688-705)
// =====================================================================================================
// k3_min is the integer image of solver.jl:733 `score >= suboptimality * best` (the Float32 score is monotone in
// k3; the host finds the smallest passing k3 with the reference's own Float32/Float64 comparison).
// With a report limit R only the first R ids in (score desc, id asc) order can own one of the first R trajectories
// (every id has a trajectory as short as its discovery level and ids ascend with level), so ids below the k3 of
// the R-th best id (k3_star) are skipped, and of the ids tied at k3_star only the first `quota` in id order kept.
__global__ void k_good_flags(const int32_t* k3, uint32_t n, int32_t id_base, int k3_min, int k3_star, uint32_t quota,
                             const uint32_t* tie_rank, uint8_t* needed) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int v = k3[r];
  if (v < 0 || v < k3_min || v < k3_star) return;
  if (v == k3_star && tie_rank[r] >= quota) return;
  needed[id_base + r] = 1;
}
__global__ void k_tie_flags(const int32_t* k3, uint32_t n, int k3_star, uint32_t* flag) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) flag[r] = k3[r] == k3_star ? 1u : 0u;
}
// histogram of k3 over the fresh ids of a level (warp-aggregated atomics; k3 == 0 is a very hot bin)
__global__ void k_k3_hist(const int32_t* k3, uint32_t n, int k3_floor, uint32_t* hist) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  int key = -1;
  if (r < n) key = k3[r];
  if (key < k3_floor) key = -1;  // ids below the minreport threshold can never be reported: no atomics for them
  const uint32_t peers = __match_any_sync(0xffffffffu, key);
  if (key >= 0 && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[key], (uint32_t)__popc(peers));
}
// needed[id] = 1 + (fewest moves from id to a reported id), 0 = not needed.  A node `dist` moves before a reported
// id can only sit on a trajectory if level(node) + dist <= maxdepth + 1 (concatenate_moves_aux cuts longer paths,
// solver.jl:688-693), which bounds the ancestor closure.  Candidates generated at BFS level `lvl` have parents of
// level lvl-1, so an edge into a child with needed == v is useful iff lvl + v <= maxdepth + 2.
__global__ void k_mark_parents(const int32_t* cand_id, const uint64_t* desc, const int32_t* front_id, uint32_t G,
                               uint8_t* needed, int pass, int lvl, int maxdepth) {
  const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= G) return;
  const int v = needed[cand_id[pos]];
  if (v == pass && lvl + v <= maxdepth + 2) {
    const int32_t pid = front_id[desc_parent(desc[pos])];
    if (!needed[pid]) needed[pid] = (uint8_t)(pass + 1);
  }
}
__global__ void k_edge_flags(const int32_t* cand_id, uint32_t G, const uint8_t* needed, uint32_t* flag, int lvl,
                             int maxdepth) {
  const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= G) return;
  const int v = needed[cand_id[pos]];
  flag[pos] = (v && lvl + v <= maxdepth + 2) ? 1u : 0u;
}
__global__ void k_edge_compact(const int32_t* cand_id, const uint64_t* desc, const int32_t* front_id, uint32_t G,
                               const uint32_t* flag, const uint32_t* pre, uint32_t gbase, EdgeRec* out) {
  const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= G || !flag[pos]) return;
  const uint64_t d = desc[pos];
  const uint32_t ps = desc_parent(d);
  EdgeRec e;
  e.child = cand_id[pos];
  e.parent = front_id[ps];
  e.move = (uint32_t)desc_kind(d) | ((uint32_t)desc_i(d) << 2) | ((uint32_t)desc_j(d) << 17);
  e.gpos = gbase + pos;
  out[pre[pos]] = e;
}
__global__ void k_id_flags(const uint8_t* needed, int32_t id_base, uint32_t n, uint32_t* flag) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) flag[r] = needed[id_base + r] ? 1u : 0u;
}
__global__ void k_id_compact(const int32_t* k3, const int32_t* cost, const int32_t* total, int32_t id_base, uint32_t n,
                             const uint32_t* flag, const uint32_t* pre, IdRec* out) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n || !flag[r]) return;
  IdRec o;
  o.id = id_base + (int32_t)r;
  o.k3 = k3[r];
  o.cost = cost[r];
  o.total = total[r];
  out[pre[r]] = o;
}

// single-pass extraction: records are appended through a warp-aggregated atomic counter (order restored on the
// host by sorting on gpos / id); *counter keeps counting past `cap` so the host can detect an overflow and retry.
__device__ __forceinline__ uint32_t warp_append(bool take, uint32_t* counter) {
  const uint32_t m = __ballot_sync(0xffffffffu, take);
  if (!m) return 0;
  const int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
}
// child / parent are written as NODE INDICES (rank of the id among the needed ids, from an exclusive scan of the
// needed flags; -1 = parent not needed), so the host builds its graph without any search
__global__ void k_edge_append(const int32_t* cand_id, const uint64_t* desc, const int32_t* front_id, uint32_t G,
                              const uint8_t* needed, const uint32_t* nrank, int lvl, int maxdepth, uint32_t gbase,
                              EdgeRec* out, uint32_t cap, uint32_t* counter) {
  const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
  bool take = false;
  int32_t child = 0;
  if (pos < G) {
    child = cand_id[pos];
    const int v = needed[child];
    take = v && lvl + v <= maxdepth + 2;
  }
  const uint32_t at = warp_append(take, counter);
  if (!take || at >= cap) return;
  const uint64_t d = desc[pos];
  const uint32_t ps = desc_parent(d);
  const int32_t pid = front_id[ps];
  EdgeRec e;
  e.child = (int32_t)nrank[child];
  e.parent = needed[pid] ? (int32_t)nrank[pid] : -1;
  e.move = (uint32_t)desc_kind(d) | ((uint32_t)desc_i(d) << 2) | ((uint32_t)desc_j(d) << 17);
  e.gpos = gbase + pos;
  out[at] = e;
}
// record of every needed id of a level, written at its node index (ascending id order, no sort on the host)
__global__ void k_id_write(const int32_t* k3, const int32_t* cost, const int32_t* total, int32_t id_base, uint32_t n,
                           const uint8_t* needed, const uint32_t* nrank, IdRec* out) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n || !needed[id_base + r]) return;
  IdRec o;
  o.id = id_base + (int32_t)r;
  o.k3 = k3[r];
  o.cost = cost[r];
  o.total = total[r];
  out[nrank[id_base + r]] = o;
}
__global__ void k_needed_flags(const uint8_t* needed, uint32_t n, uint32_t* flag) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = needed[i] ? 1u : 0u;
}
__global__ void k_id_append(const int32_t* k3, const int32_t* cost, const int32_t* total, int32_t id_base, uint32_t n,
                            const uint8_t* needed, IdRec* out, uint32_t cap, uint32_t* counter) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  const bool take = r < n && needed[id_base + r];
  const uint32_t at = warp_append(take, counter);
  if (!take || at >= cap) return;
  IdRec o;
  o.id = id_base + (int32_t)r;
  o.k3 = k3[r];
  o.cost = cost[r];
  o.total = total[r];
  out[at] = o;
}
int launch_edge_append(const int32_t* cand_id, const uint64_t* desc, const int32_t* front_id, uint32_t G,
                       const uint8_t* needed, const uint32_t* nrank, int lvl, int maxdepth, uint32_t gbase, void* out,
                       uint32_t cap, uint32_t* counter, cudaStream_t st) {
  if (!G) return 0;
  k_edge_append<<<(G + 255) / 256, 256, 0, st>>>(cand_id, desc, front_id, G, needed, nrank, lvl, maxdepth, gbase,
                                                (EdgeRec*)out, cap, counter);
  return 1;
}
int launch_id_write(const int32_t* k3, const int32_t* cost, const int32_t* total, int32_t id_base, uint32_t n,
                    const uint8_t* needed, const uint32_t* nrank, void* out, cudaStream_t st) {
  if (!n) return 0;
  k_id_write<<<(n + 255) / 256, 256, 0, st>>>(k3, cost, total, id_base, n, needed, nrank, (IdRec*)out);
  return 1;
}
int launch_needed_flags(const uint8_t* needed, uint32_t n, uint32_t* flag, cudaStream_t st) {
  if (!n) return 0;
  k_needed_flags<<<(n + 255) / 256, 256, 0, st>>>(needed, n, flag);
  return 1;
}
int launch_id_append(const int32_t* k3, const int32_t* cost, const int32_t* total, int32_t id_base, uint32_t n,
                     const uint8_t* needed, void* out, uint32_t cap, uint32_t* counter, cudaStream_t st) {
  if (!n) return 0;
  k_id_append<<<(n + 255) / 256, 256, 0, st>>>(k3, cost, total, id_base, n, needed, (IdRec*)out, cap, counter);
  return 1;
}

// regionfile mode: ids whose score passes their region's minreport threshold, appended as (gid, k3, rid) records
struct QualRec {
  int32_t gid, k3, rid, pad;
};
__global__ void k_qual_append(const int32_t* k3, const uint16_t* nrid, const int32_t* k3_min, int32_t id_base, uint32_t n,
                              QualRec* out, uint32_t cap, uint32_t* counter) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  bool take = false;
  int v = -1, rid = 0;
  if (r < n) {
    v = k3[r];
    rid = nrid ? (int)nrid[r] : 0;
    take = v >= 0 && v >= k3_min[rid];
  }
  const uint32_t at = warp_append(take, counter);
  if (!take || at >= cap) return;
  out[at] = QualRec{id_base + (int32_t)r, v, rid, 0};
}
__global__ void k_set_needed(const int32_t* gids, uint32_t n, uint8_t* needed) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) needed[gids[i]] = 1;
}
int launch_qual_append(const int32_t* k3, const uint16_t* nrid, const int32_t* k3_min, int32_t id_base, uint32_t n, void* out,
                       uint32_t cap, uint32_t* counter, cudaStream_t st) {
  if (!n) return 0;
  k_qual_append<<<(n + 255) / 256, 256, 0, st>>>(k3, nrid, k3_min, id_base, n, (QualRec*)out, cap, counter);
  return 1;
}
int launch_set_needed(const int32_t* gids, uint32_t n, uint8_t* needed, cudaStream_t st) {
  if (!n) return 0;
  k_set_needed<<<(n + 255) / 256, 256, 0, st>>>(gids, n, needed);
  return 1;
}

int launch_good_flags(const int32_t* k3, uint32_t n, int32_t id_base, int k3_min, int k3_star, uint32_t quota,
                      const uint32_t* tie_rank, uint8_t* needed, cudaStream_t st) {
  if (!n) return 0;
  k_good_flags<<<(n + 255) / 256, 256, 0, st>>>(k3, n, id_base, k3_min, k3_star, quota, tie_rank, needed);
  return 1;
}
int launch_tie_flags(const int32_t* k3, uint32_t n, int k3_star, uint32_t* flag, cudaStream_t st) {
  if (!n) return 0;
  k_tie_flags<<<(n + 255) / 256, 256, 0, st>>>(k3, n, k3_star, flag);
  return 1;
}
int launch_k3_hist(const int32_t* k3, uint32_t n, int k3_floor, uint32_t* hist, cudaStream_t st) {
  if (!n) return 0;
  k_k3_hist<<<(n + 255) / 256, 256, 0, st>>>(k3, n, k3_floor, hist);
  return 1;
}
int launch_mark_parents(const int32_t* cand_id, const uint64_t* desc, const int32_t* front_id, uint32_t G,
                        uint8_t* needed, int pass, int lvl, int maxdepth, cudaStream_t st) {
  if (!G) return 0;
  k_mark_parents<<<(G + 255) / 256, 256, 0, st>>>(cand_id, desc, front_id, G, needed, pass, lvl, maxdepth);
  return 1;
}
int launch_edge_flags(const int32_t* cand_id, uint32_t G, const uint8_t* needed, uint32_t* flag, int lvl, int maxdepth,
                      cudaStream_t st) {
  if (!G) return 0;
  k_edge_flags<<<(G + 255) / 256, 256, 0, st>>>(cand_id, G, needed, flag, lvl, maxdepth);
  return 1;
}
int launch_edge_compact(const int32_t* cand_id, const uint64_t* desc, const int32_t* front_id, uint32_t G,
                        const uint32_t* flag, const uint32_t* pre, uint32_t gbase, void* out, cudaStream_t st) {
  if (!G) return 0;
  k_edge_compact<<<(G + 255) / 256, 256, 0, st>>>(cand_id, desc, front_id, G, flag, pre, gbase, (EdgeRec*)out);
  return 1;
}
int launch_id_flags(const uint8_t* needed, int32_t id_base, uint32_t n, uint32_t* flag, cudaStream_t st) {
  if (!n) return 0;
  k_id_flags<<<(n + 255) / 256, 256, 0, st>>>(needed, id_base, n, flag);
  return 1;
}
int launch_id_compact(const int32_t* k3, const int32_t* cost, const int32_t* total, int32_t id_base, uint32_t n,
                      const uint32_t* flag, const uint32_t* pre, void* out, cudaStream_t st) {
  if (!n) return 0;
  k_id_compact<<<(n + 255) / 256, 256, 0, st>>>(k3, cost, total, id_base, n, flag, pre, (IdRec*)out);
  return 1;
}

// =====================================================================================================
// dedup verifier (NW_FLAG_VERIFY_DEDUP): every candidate that was merged into an already known index is compared
// element by element with that index, both re-materialised through their index maps.  Equal sequences always have
// equal fingerprints, so a clean run proves the dedup of this search exact (solver.jl:535 semantics).
// =====================================================================================================
// Element t (1-based) of a child through the index map of its move over the parent sequence `seq`:
//   dup / del / identity:  seq[t - 1 + (t > lo ? shift : 0)]          (lo = q / p-1 / never, shift = p-q / q-p / 0)
//   inversion:             lo < t < hi ? -seq[lo + hi - t - 1] : seq[t - 1]           (lo = p, hi = q)
// (apply_duplication / apply_deletion / apply_inversion, solver.jl:244-278).
struct VerItem {            // 32 bytes in shared memory, one per merged candidate of the CTA
  uint32_t psA, psB;        // parent slots (A in this level's frontier, B in the frontier of level lvB)
  uint32_t lvB_n;           // owner's level | child length << 8
  uint32_t loA_hiA;         // lo | hi << 16   (hi = 0: no inversion window)
  int32_t shA;              // shift of the non-inversion map
  uint32_t loB_hiB;
  int32_t shB;
  uint32_t pad;
};
__device__ __forceinline__ void move_params(int L, int kind, int p, int q, uint32_t& lo_hi, int32_t& sh) {
  int lo = kind == KIND_DEL ? p - 1 : (kind == KIND_DUP ? q : (kind == KIND_INV ? p : 0x7fff));
  int hi = kind == KIND_INV ? q : 0;
  sh = kind == KIND_DUP ? p - q : (kind == KIND_DEL ? q - p : 0);
  lo_hi = (uint32_t)lo | ((uint32_t)hi << 16);
}
// element t of a child as its low 16 bits (upper bits undefined): equality of two int16 values is equality of their low
// 16 bits, so neither the load nor the inversion's negation needs a sign extension
template <bool INV>
__device__ __forceinline__ uint32_t ver_elem(const uint16_t* seq, int lo, int hi, int sh, int t) {
  if (INV) {
    const bool flip = t > lo && t < hi;
    const uint32_t v = seq[(unsigned)(flip ? lo + hi - t - 1 : t - 1)];
    return flip ? 0u - v : v;
  }
  return seq[(unsigned)(t - 1 + (t > lo ? sh : 0))];
}
template <bool INVA, bool INVB>
__device__ __forceinline__ bool ver_compare(const int16_t* sa_, int loA, int hiA, int shA, const int16_t* sb_, int loB, int hiB,
                                            int shB, int n, int gl) {
  const uint16_t* sa = reinterpret_cast<const uint16_t*>(sa_);
  const uint16_t* sb = reinterpret_cast<const uint16_t*>(sb_);
  uint32_t acc = 0;
  int t = 1 + gl;
  for (; t + 8 <= n; t += 16) {  // two elements per lane and step: four loads in flight
    const uint32_t a0 = ver_elem<INVA>(sa, loA, hiA, shA, t), b0 = ver_elem<INVB>(sb, loB, hiB, shB, t);
    const uint32_t a1 = ver_elem<INVA>(sa, loA, hiA, shA, t + 8), b1 = ver_elem<INVB>(sb, loB, hiB, shB, t + 8);
    acc |= (a0 ^ b0) | (a1 ^ b1);
  }
  if (t <= n) acc |= ver_elem<INVA>(sa, loA, hiA, shA, t) ^ ver_elem<INVB>(sb, loB, hiB, shB, t);
  return (acc & 0xffffu) != 0;
}
// Element-wise dedup verifier.  A CTA takes 1024 consecutive candidates.  Phase 1 (4 candidates per thread): every merged
// candidate reads its own move and -- one random read -- its owner's, and leaves a 32-byte item in shared memory, filed
// under its class: no inversion on either side / on one side (filed with that side as B: equality is symmetric) / on
// both.  Phase 2: groups of 8 lanes take one item each and stride over the elements (coalesced reads of the parent
// sequences in the frontier blobs, L2-resident at these sizes).  The items are walked in (class, length) order -- a
// counting sort in shared memory --, so the four groups of a warp run the same specialised loop the same number of times:
// no per-element kind tests and next to no divergence between the groups.
// No host synchronisation: mismatches accumulate in a device counter that the search reads once, at its end.
constexpr int VER_CAND = 1024;
constexpr int VER_LBINS = 32;            // length bins of 8 elements (longer children share the last bin)
constexpr int VER_BINS = 3 * VER_LBINS;  // (class, length bin)
__global__ void __launch_bounds__(256) k_verify_dups(ExpandArgs A, VerifyTable V, const uint32_t* fresh_mask, const uint32_t* owner,
                                                     uint32_t first, uint32_t G, unsigned int* mismatches) {
  __shared__ __align__(16) VerItem s_item[VER_CAND];
  __shared__ uint16_t s_order[VER_CAND];
  __shared__ uint32_t s_bin[VER_BINS + 1];  // items per (class, length bin); [VER_BINS] = items filed
  __shared__ uint32_t s_base[VER_BINS];
  const int tid = threadIdx.x;
  if (tid <= VER_BINS) s_bin[tid] = 0;
  __syncthreads();
  bool bad = false;
  int my_slot[4] = {-1, -1, -1, -1}, my_key[4] = {0, 0, 0, 0};
  const uint32_t pos0 = first + blockIdx.x * (uint32_t)VER_CAND + (uint32_t)tid * 4u;  // first: multiple of 32
  if (pos0 < G) {
    const uint32_t word = fresh_mask[pos0 >> 5] >> (pos0 & 31);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t pos = pos0 + u;
      if (pos >= G || ((word >> u) & 1u)) continue;
      const uint64_t d = A.desc[pos];
      const uint32_t ps = desc_parent(d);
      const int L = A.F.len[ps];
      const int kind = desc_kind(d), p = desc_i(d), q = desc_j(d);
      const int n = child_len(L, kind, p, q);
      const uint32_t g = owner[pos];
      int l = 0;
      while (l + 1 < V.n_levels && g >= V.gbase[l + 1]) ++l;
      const uint64_t od = V.desc[l][g - V.gbase[l]];
      const uint32_t ops = desc_parent(od);
      const int Lo = V.len[l][ops];
      const int okind = desc_kind(od), op = desc_i(od), oq = desc_j(od);
      if (n != child_len(Lo, okind, op, oq)) {
        bad = true;
        continue;
      }
      VerItem it;
      uint32_t lhA, lhB;
      int32_t shA, shB;
      move_params(L, kind, p, q, lhA, shA);
      move_params(Lo, okind, op, oq, lhB, shB);
      // parent A lives in this level's frontier (level index V.n_levels - 1), parent B in level l's; store both as
      // (level, slot) so that the two sides can be swapped
      const int lvA = V.n_levels - 1;
      const bool invA = kind == KIND_INV, invB = okind == KIND_INV;
      const bool swap = invA && !invB;  // a single inversion is filed as side B
      it.psA = swap ? ops : ps;
      it.psB = swap ? ps : ops;
      const int cls = (invA ? 1 : 0) + (invB ? 1 : 0);
      it.lvB_n = (uint32_t)(swap ? lvA : l) | ((uint32_t)(swap ? l : lvA) << 4) | ((uint32_t)n << 8) | ((uint32_t)cls << 24);
      it.loA_hiA = swap ? lhB : lhA;
      it.shA = swap ? shB : shA;
      it.loB_hiB = swap ? lhA : lhB;
      it.shB = swap ? shA : shB;
      it.pad = 0;
      const uint32_t slot = atomicAdd(&s_bin[VER_BINS], 1u);
      s_item[slot] = it;
      // key = class, then length in steps of 8 elements (one stride of a lane group): neighbours in the processing order
      // run the same loop the same number of times
      const int key = cls * VER_LBINS + min(VER_LBINS - 1, n >> 3);
      atomicAdd(&s_bin[key], 1u);
      my_slot[u] = (int)slot;
      my_key[u] = key;
    }
  }
  __syncthreads();
  if (tid < 32) {  // exclusive scan of the bins (one warp, VER_BINS <= 96)
    uint32_t carry = 0;
    for (int b0 = 0; b0 < VER_BINS; b0 += 32) {
      const int b = b0 + tid;
      const uint32_t v = b < VER_BINS ? s_bin[b] : 0u;
      uint32_t inc = v;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (tid >= o) inc += t;
      }
      if (b < VER_BINS) {
        s_base[b] = carry + inc - v;
        s_bin[b] = 0;  // becomes the placement cursor
      }
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (my_slot[u] >= 0) s_order[s_base[my_key[u]] + atomicAdd(&s_bin[my_key[u]], 1u)] = (uint16_t)my_slot[u];
  __syncthreads();
  const int grp = tid >> 3, gl = tid & 7;  // 32 groups of 8 lanes
  const uint32_t n_items = s_bin[VER_BINS];
  for (uint32_t i = grp; i < n_items; i += 32) {
    const VerItem it = s_item[s_order[i]];
    const int lvB = (int)(it.lvB_n & 15u), lvA = (int)((it.lvB_n >> 4) & 15u), n = (int)((it.lvB_n >> 8) & 0xffffu);
    const int cls = (int)(it.lvB_n >> 24);
    // level 0 of the table is the root pseudo-level: its "frontier" is level 1's (V.blob[0] is set accordingly)
    const int16_t* sa = reinterpret_cast<const int16_t*>(V.blob[lvA] + (size_t)it.psA * V.pstride[lvA]);
    const int16_t* sb = reinterpret_cast<const int16_t*>(V.blob[lvB] + (size_t)it.psB * V.pstride[lvB]);
    const int loA = (int)(it.loA_hiA & 0xffffu), hiA = (int)(it.loA_hiA >> 16);
    const int loB = (int)(it.loB_hiB & 0xffffu), hiB = (int)(it.loB_hiB >> 16);
    bool diff;
    if (cls == 0) diff = ver_compare<false, false>(sa, loA, hiA, it.shA, sb, loB, hiB, it.shB, n, gl);
    else if (cls == 1) diff = ver_compare<false, true>(sa, loA, hiA, it.shA, sb, loB, hiB, it.shB, n, gl);
    else diff = ver_compare<true, true>(sa, loA, hiA, it.shA, sb, loB, hiB, it.shB, n, gl);
    bad |= diff;
  }
  if (__syncthreads_or(bad ? 1 : 0) && tid == 0) atomicAdd(mismatches, 1u);  // any non-zero count condemns the search
}
// verifies the candidates at positions [first, end) of the level (first a multiple of 32)
int launch_verify_dups(const ExpandArgs& A, const VerifyTable& V, const uint32_t* fresh_mask, const uint32_t* owner, uint32_t first,
                       uint32_t end, unsigned int* mismatches, cudaStream_t st) {
  if (end <= first) return 0;
  k_verify_dups<<<(end - first + VER_CAND - 1) / VER_CAND, 256, 0, st>>>(A, V, fresh_mask, owner, first, end, mismatches);
  return 1;
}

// =====================================================================================================
// integer-pipe roofline denominators: dependency-free micro-kernels, 8 independent chains / thread.
// MODE 0: VIADDMNMX (DPX add+max, alu pipe)   MODE 1: IADD3 (alu pipe)   MODE 2: VIADDMNMX + IMAD interleaved
// (alu + fma pipes).  Reported as thread-instructions per second.
// =====================================================================================================
template <int MODE>
__global__ void __launch_bounds__(256) k_int_peak(int* out, int iters, int a, int b, int m) {
  int x[8], y[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    x[k] = threadIdx.x + k;
    y[k] = blockIdx.x - k;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (MODE == 0) x[k] = __viaddmax_s32(x[k], a, b + k);
        if (MODE == 1) x[k] = x[k] + x[(k + 1) & 7] + a;  // IADD3; operands change every step, nothing is loop-invariant
        if (MODE == 2) {
          x[k] = __viaddmax_s32(x[k], a, b + k);
          y[k] = y[k] * m + a;
        }
      }
    }
  }
  int s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k] ^ y[k];
  if (s == 0x7fffffff) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// returns thread-instructions issued per launch
double launch_int_peak(int mode, int* out, int iters, cudaStream_t st) {
  const int grid = 148 * 8, block = 256;
  if (mode == 0) k_int_peak<0><<<grid, block, 0, st>>>(out, iters, 3, 1000, 1);
  if (mode == 1) k_int_peak<1><<<grid, block, 0, st>>>(out, iters, 3, 1000, 1);
  if (mode == 2) k_int_peak<2><<<grid, block, 0, st>>>(out, iters, 3, 1000, 1);
  const double per_thread = (double)iters * 4 * 8 * (mode == 2 ? 2 : 1);
  return per_thread * grid * block;
}

// Opt-in dynamic shared memory limits are per (device, function) state: they are raised ONCE per device here, never
// per launch (a per-launch cudaFuncSetAttribute races between contexts running on different host threads).
void init_score_fast_attrs();
void init_score_coop_attrs();
void init_score_pair_attrs();
void init_kernel_attrs() {
  int dev = 0, optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const int mx = optin - 2048;  // the limit covers static + dynamic shared memory of a block
  cudaFuncSetAttribute(k_expand<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_expand<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_expand<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_expand<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_expand<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_expand<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_expand<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_expand<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_expand<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_expand<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_expand<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_expand<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_expand_pairs<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_expand_pairs<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_materialise, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  init_score_fast_attrs();
  init_score_coop_attrs();
  init_score_pair_attrs();
}

int launch_moveset(const RegionDev* regions_dev, int n_regions, int nx_max, cudaStream_t st) {
  const long long n = (long long)nx_max * nx_max;
  k_moveset<<<dim3((unsigned)((n + 255) / 256), (unsigned)n_regions), 256, 0, st>>>(regions_dev);
  return 1;
}
int launch_insert_roots(const unsigned char* blob, uint32_t pstride, int LS, const int32_t* len, int n_regions,
                        uint64_t* table, uint64_t tmask, cudaStream_t st) {
  k_insert_roots<<<(n_regions + 127) / 128, 128, 0, st>>>(blob, pstride, LS, len, n_regions, table, tmask);
  return 1;
}
int launch_rehash(const uint64_t* oldT, uint64_t old_cap, uint64_t* newT, uint64_t new_mask, cudaStream_t st) {
  k_table_rehash<<<148 * 8, 256, 0, st>>>(oldT, old_cap, newT, new_mask);
  return 1;
}

}  // namespace nw
