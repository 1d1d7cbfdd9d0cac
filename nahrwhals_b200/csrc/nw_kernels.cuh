// nw_kernels.cuh -- device-side structures and launch prototypes shared by nw_kernels.cu, nw_score_fast.cu
// and the host driver (nw_driver.cu).  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nw_device.h"

namespace nw {

constexpr uint64_t TBL_EMPTY = ~0ull;
// Dedup table slot = 16 bytes, claimed and updated with ONE 128-bit compare-and-swap:
//   word 0 = fingerprint bits 0..63, word 1 = fingerprint bits 64..95 << 32 | first-discoverer position (32 bits).
// Two slots share a 32-byte DRAM sector, so a probe that steps to the neighbour slot usually stays in the sector it
// already fetched, and the table (and its per-level clear) is half the size of a 32-byte-slot layout.  The table keys
// on 96 of the fingerprint's 124 bits; exactness does not rest on the key width: the search verifies every merged
// candidate element-wise against its owner (k_verify_dups) unless the caller opts out.
constexpr int TBL_WORDS = 2;
constexpr int MAX_LEVELS = 16;
constexpr uint32_t WORK_PAD = 0xffffffffu;
// The per-length counters used to bucket fresh candidates are split NSUB ways (by new-rank / 256) to spread the
// atomics of k_assign / k_bucket_scatter over more addresses.
constexpr int NSUB = 8;
// bins are (region, child length, sub-counter); lcap = (longest child length of the level) + 1
__host__ __device__ inline int bucket_key(int rid, int lcap, int len, uint32_t new_rank) {
  return (rid * lcap + len) * NSUB + (int)((new_rank >> 10) & (NSUB - 1));
}

// Per-region read-only context in device memory.  [planes | upx | rt_rev] is one contiguous, 16-B aligned
// blob of ctx_bytes so the scoring kernels can stage it with a single TMA bulk copy.
struct RegionDev {
  int n_x, n_y;
  int NYP;     // padded number of (reversed) columns incl. the virtual column 0; multiple of 32
  int PWS;     // 32-bit words per (block, strand) bit-plane row, incl. zero padding
  int MW;      // 32-bit words per moveset row
  int sum_rt;  // sum of y lengths in units
  uint32_t rid1;
  uint32_t ctx_bytes;
  const unsigned char* ctx;  // blob base
  uint32_t off_upx, off_rtrev;  // byte offsets inside the blob (planes at 0)
  int rt_copy_stride;           // ints between the 4 shifted copies of rt_rev (copy c holds rt_rev[t + c])
  const int32_t* rt;     // rt[1..n_y] units
  const int32_t* cumr;   // cumr[j] = sum_{j'<=j} rt
  uint32_t* ms_dd;       // [(n_x+1) * MW] bit rows, 1-based block ids
  uint32_t* ms_inv;
  uint32_t* ms_any;      // [n_x+1] non-zero iff the block has any dup/del or inversion partner
  // corridor tables of this region's n_y (device mirror of BandStore; refreshed by the host before every DP launch)
  const int32_t* band_off;   // [LMAX+2] row offset of the table of child length L, -1 if absent
  const int16_t* band_ab;    // generic kernel: {a,b} per row
  const int16_t* band_eb;    // register kernel: eb per row (row 0 = virtual)
  const uint8_t* band_sb;    // register kernel: window shift per row
  const uint32_t* band_mask; // register kernel: 4 mask words per row
};

// One BFS level's frontier: P parents, each a blob [seq int16 x LS | F U4 x (LS+1) | G U4 x (LS+1)].
struct FrontierDev {
  int P, LS;
  uint32_t pstride;  // bytes per parent blob (multiple of 16)
  const unsigned char* blob;
  const int32_t* len;
  const int32_t* id;
  const uint8_t* dupok;
  // several independent regions in one search (regionfile mode): region of every parent, region table
  const uint16_t* rid;        // nullptr: every parent belongs to region 0
  const RegionDev* regions;   // device array
};
__device__ __forceinline__ int region_index(const FrontierDev& F, uint32_t parent_slot) {
  return F.rid ? (int)F.rid[parent_slot] : 0;
}

struct LevelTable {  // candidate -> id lookup across levels (owner resolution)
  int n_levels;                      // levels already finished (level 0 = root pseudo-level)
  uint32_t gbase[MAX_LEVELS + 1];    // first global position of each finished level; gbase[n_levels] = current base
  const int32_t* cand_id[MAX_LEVELS];
};

struct ExpandArgs {
  FrontierDev F;
  int nx_max;            // largest n_x over the regions of this launch (shared-memory layout of k_expand)
  uint32_t* cnt;         // [P] candidates per parent (count pass out)
  const uint32_t* off;   // [P] exclusive offsets (emit pass in)
  int* maxlen;           // max child length (count pass out)
  uint64_t* desc;        // [G] (emit)
  uint32_t* slot;        // [G] hash-table slot of each candidate (emit)
  uint64_t* table;
  uint64_t tmask;
  uint32_t gbase;
  const U4* pw;
  int force_pairs;           // NW_FLAG_PAIR_EXPAND
  int weak_fp;               // NW_FLAG_DEBUG_WEAK_FP (tests only)
};

// Frontier-sharded search (one deep search over `world` ranks): expansion, dedup and selection are
// replicated (deterministic, identical on every rank); only the banded DP -- the dominant cost -- is divided: rank r
// scores the fresh candidates whose new-rank lies in [r * chunk, (r + 1) * chunk), chunk = own_chunk(n_new, world),
// and one all-gather per level over the k3 / cost / total arrays (in place, `chunk` items per rank) completes them.
// n_dev points at the level's fresh-candidate count in device memory (the host may not know it yet at launch time).
struct OwnRange {
  int world, rank;
  const uint32_t* n_dev;
};
__host__ __device__ inline uint32_t own_chunk(uint32_t n, int world) {
  const uint32_t c = (((n + (uint32_t)world - 1u) / (uint32_t)world) + 3u) & ~3u;  // multiple of 4: 16-B aligned segments
  return c < 4u ? 4u : c;
}
__device__ __forceinline__ bool owns(const OwnRange& o, uint32_t r) {
  return o.world <= 1 || r / own_chunk(*o.n_dev, o.world) == (uint32_t)o.rank;
}

// Everything needed to re-materialise the index a candidate of any finished level stands for (dedup verifier)
struct VerifyTable {
  int n_levels;  // levels 0..n_levels-1 are valid; level 0 is the root pseudo-level
  uint32_t gbase[MAX_LEVELS + 1];
  const uint64_t* desc[MAX_LEVELS];
  const unsigned char* blob[MAX_LEVELS];
  uint32_t pstride[MAX_LEVELS];
  const int32_t* len[MAX_LEVELS];
};

struct ScoreArgs {
  FrontierDev F;
  uint32_t ctx_max;         // largest region context (shared-memory layout of k_score_fast)
  int ny_max;               // largest n_y (scratch layout of k_score_generic)
  const uint16_t* cta_rid;  // k_score_fast: region of every CTA of this launch (nullptr: region 0)
  const uint32_t* work;     // [n_work] new-rank per lane slot, WORK_PAD = idle
  uint32_t n_work;
  int lmax;                 // longest child in this launch (sizes the shared-memory tiles of k_score_fast)
  const uint32_t* newlist;  // rank -> candidate position in this level
  const uint64_t* desc;
  int32_t* k3;              // [n_new] out
  int32_t* cost;            // [n_new] out (units; -2 unreachable)
  int32_t* total;           // [n_new] out (units)
  int* best_k3;             // [n_regions] best score*1000 per region
  unsigned long long* cells;  // DP cells evaluated (stat)
  int32_t* scratch;         // generic kernel rows
  int force;
};

struct AssignArgs {
  LevelTable LT;
  const uint32_t* slot;
  const uint64_t* table;
  const uint32_t* owner;    // first-discoverer position of each candidate's index (k_resolve)
  // first-discoverer flags as a bit mask (bit pos & 31 of word pos >> 5) and the exclusive scan of the words' popcounts:
  // the new-rank of ANY candidate of the level is wprefix[w] + popc(mask[w] & below(pos)), read from two arrays that
  // stay L2-resident (1/16 of a byte-per-flag + 4-byte-per-rank layout) -- duplicates look their owner's rank up there
  const uint32_t* fresh_mask;
  const uint32_t* wprefix;
  const uint64_t* desc;
  const int32_t* front_len;
  uint32_t gbase, G;
  int32_t id_base;  // id of new-rank 0
  const uint16_t* front_rid;  // region of every parent (nullptr: region 0)
  const RegionDev* regions;
  int lcap;                   // bins per region = lcap * NSUB
  uint16_t* new_rid;          // [n_new] out: region of every fresh candidate
  int force;
  OwnRange own;     // sharded search: only this rank's share of the fresh candidates enters the DP histogram
  int32_t* cand_id;   // [G] out
  uint32_t* newlist;  // [n_new] out: rank -> pos
  uint16_t* new_len;  // [n_new] out: child length
  int32_t* k3;        // [n_new] (early exits are settled here)
  int32_t* cost;
  int32_t* total;
  uint32_t* hist;     // [Lcap] count of DP-needing new candidates per child length
};

struct MaterialiseArgs {
  FrontierDev F;           // parents (current level)
  const uint32_t* sel;     // [P2] selected new-ranks, in frontier order
  const uint32_t* newlist;
  const uint64_t* desc;
  int32_t id_base;
  int maxdup, nx_max;
  int P2, LS2;
  uint32_t pstride2;
  unsigned char* blob2;
  int32_t* len2;
  int32_t* id2;
  uint8_t* dupok2;
  uint16_t* rid2;  // region of every new parent
  const U4* pw;
};

// non-empty (region, length) bin of the bucket histogram: bin = rid * lcap + len; n = candidates (device -> host) or
// first work slot (host -> device)
struct BinRec {
  uint32_t bin, n;
};
struct EdgeRec {
  int32_t child, parent;
  uint32_t move;  // kind | i << 2 | j << 17
  uint32_t gpos;  // global discovery position (edge order of solver.jl:559-560)
};
__host__ __device__ inline int edge_kind(const EdgeRec& e) { return (int)(e.move & 3u); }
__host__ __device__ inline int edge_i(const EdgeRec& e) { return (int)((e.move >> 2) & 0x7fffu); }
__host__ __device__ inline int edge_j(const EdgeRec& e) { return (int)((e.move >> 17) & 0x7fffu); }
struct IdRec {
  int32_t id, k3, cost, total;
};
__device__ __forceinline__ uint32_t mask_rank(const uint32_t* mask, const uint32_t* wprefix, uint32_t pos) {
  return wprefix[pos >> 5] + __popc(mask[pos >> 5] & ((1u << (pos & 31)) - 1u));
}
__device__ __forceinline__ bool mask_bit(const uint32_t* mask, uint32_t pos) { return (mask[pos >> 5] >> (pos & 31)) & 1u; }
// sub-counter of a fresh candidate's (region, length) bin: a function of its new-rank only, constant over 1024
// consecutive ranks (one CTA of the bucket scatter)
__host__ __device__ inline int bucket_sub(uint32_t new_rank) { return (int)((new_rank >> 10) & (NSUB - 1)); }
// ---- launchers (return the number of kernels launched) ----
int launch_score_fast(int W, const ScoreArgs& A, cudaStream_t st);
bool score_fast_supports(int width);  // true if some compiled window >= width
int score_fast_window(int width);     // smallest compiled window >= width (0 if none)
size_t score_fast_smem(int W, uint32_t ctx_bytes, int lmax);  // dynamic shared memory of one CTA
// packed form (nw_score_pair.cu): two candidates of one length per thread in the 16-bit halves of the DP registers;
// CTA = 64 threads = 128 work slots, buckets padded to 64 slots; only for totals below PAIR_TOTAL_MAX units
int launch_score_pair(int W, const ScoreArgs& A, cudaStream_t st);
size_t score_pair_smem(int W, uint32_t ctx_bytes, int lmax);
// warp-cooperative form (nw_score_coop.cu): four lanes per candidate, CTA = 32 work slots
int launch_score_coop(int W, const ScoreArgs& A, cudaStream_t st);
int score_coop_window(int width);  // smallest compiled cooperative window >= width (0 if none)
size_t score_coop_smem(int W, uint32_t ctx_bytes, int lmax);

// ---- small device helpers ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (SASS: UBLKCP).  dst/src 16-B aligned, bytes % 16 == 0.
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

}  // namespace nw
