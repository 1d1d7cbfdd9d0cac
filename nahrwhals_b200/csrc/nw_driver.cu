// nw_driver.cu -- host frontier driver + C ABI of libnwsolve.so.
//
// Restates the control flow of aln_bfs_scored / get_good_trajectories / main of the reference
// (inst/extdata/scripts/scripts_nw_main/solver.jl:644-686, :727-745, :755-812) around the CUDA kernels of
// nw_kernels.cu / nw_score_fast.cu.  Per BFS level: count -> scan -> emit+fingerprint+dedup -> resolve first
// discoverers -> assign ids -> bucket by child length -> banded DP -> select the next frontier -> materialise.
// There is no CPU fallback anywhere in this file: every entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <functional>
#include <map>
#include <memory>
#include <atomic>
#include <mutex>
#include <thread>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/nwsolve.h"
#include "nw_band.h"
#include "nw_exchange.h"
#include "nw_host.h"
#include "nw_kernels.cuh"
#include "nw_launch.h"

using namespace nw;

static thread_local std::string g_last_error;
static std::atomic<int> g_live_ctxs{0};  // contexts alive in this process: they share the process's host cores
struct nw_ctx;

namespace {

struct Err {
  int code;
  std::string msg;
};
[[noreturn]] void fail(int code, const std::string& m) { throw Err{code, m}; }

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      fail(e_ == cudaErrorMemoryAllocation ? NW_ERR_NOMEM : NW_ERR_CUDA,                           \
           std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

// Device memory pool: a solve allocates ~2 GB of per-level arrays; cudaMalloc/cudaFree of that size costs tens of
// milliseconds and synchronises the device, so freed blocks are kept per context and reused by later solves.
// Stream synchronisation by polling: the search makes several short host round trips per level, and a blocking
// wait adds a scheduler wake-up (tens of microseconds, far more when ranks outnumber idle cores) to each of them.
inline cudaError_t spin_sync(cudaStream_t s) {
  cudaError_t e = cudaSuccess;
  for (int it = 0; it < 64; ++it) {  // ~50-100 us of polling covers the short round trips ...
    e = cudaStreamQuery(s);
    if (e != cudaErrorNotReady) return e;
  }
  return cudaStreamSynchronize(s);   // ... long kernels block instead of burning a core other contexts need
}

struct DevPool {
  std::multimap<size_t, void*> free_blocks;
  size_t pooled_bytes = 0;
  void* take(size_t bytes, size_t* cap_out) {
    auto it = free_blocks.lower_bound(bytes);
    if (it != free_blocks.end() && it->first <= bytes * 2) {
      void* p = it->second;
      *cap_out = it->first;
      pooled_bytes -= it->first;
      free_blocks.erase(it);
      return p;
    }
    return nullptr;
  }
  void give(void* p, size_t cap) {
    free_blocks.emplace(cap, p);
    pooled_bytes += cap;
  }
  void trim() {
    for (auto& kv : free_blocks) cudaFree(kv.second);
    free_blocks.clear();
    pooled_bytes = 0;
  }
};
static thread_local DevPool* g_pool = nullptr;  // set by the ABI entry points to the calling context's pool

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), cap(o.cap) {
    o.p = nullptr;
    o.cap = 0;
  }
  DevBuf& operator=(DevBuf&& o) noexcept {
    release();
    p = o.p;
    cap = o.cap;
    o.p = nullptr;
    o.cap = 0;
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) {
      if (g_pool)
        g_pool->give(p, cap);
      else
        cudaFree(p);
    }
    p = nullptr;
    cap = 0;
  }
  void* ensure(size_t bytes, bool grow_slack = false) {
    if (bytes == 0) bytes = 16;
    if (bytes > cap) {
      release();
      (void)grow_slack;
      // size classes (powers of two up to 256 MiB, then multiples of 64 MiB) make pooled blocks reusable across
      // regions of different shapes: a cudaMalloc in the middle of a batch stalls every context on the device
      size_t want = 256;
      if (bytes <= ((size_t)256 << 20)) {
        while (want < bytes) want <<= 1;
      } else {
        const size_t q = (size_t)64 << 20;
        want = (bytes + q - 1) / q * q;
      }
      if (g_pool) {
        p = g_pool->take(want, &cap);
        if (p) return p;
      }
      cudaError_t e = cudaMalloc(&p, want);
      if (e == cudaErrorMemoryAllocation && g_pool && g_pool->pooled_bytes) {
        cudaGetLastError();
        g_pool->trim();
        e = cudaMalloc(&p, want);
      }
      if (e != cudaSuccess) p = nullptr;
      CK(e);
      cap = want;
    }
    return p;
  }
  template <typename T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

// counted copies (nw_stats.h2d_bytes / d2h_bytes)
void h2d(nw_ctx* c, void* dst, const void* src, size_t n);
void d2h(nw_ctx* c, void* dst, const void* src, size_t n);

// Corridor tables depend only on (child length, n_y, force): one process-wide host cache shared by all contexts
// (a regionfile run meets the same few thousand shapes over and over, and every context would otherwise rebuild
// them).  Tables are immutable once published; the packed arrays the device mirrors copy from are append-only and
// guarded by the mutex.
struct BandHost {
  int n_y = 0;
  int force = 0;  // corridor mode (NW_BAND_*)
  std::mutex mu;
  std::unique_ptr<std::atomic<BandTable*>[]> tabs;  // [LMAX+2], built on first use, never freed
  std::vector<int32_t> off;                         // [LMAX+2] row offset or -1
  std::vector<int16_t> ab, eb;
  std::vector<uint8_t> sb;
  std::vector<uint32_t> mask;
  BandHost(int ny, int f) : n_y(ny), force(f), tabs(new std::atomic<BandTable*>[LMAX + 2]), off(LMAX + 2, -1) {
    for (int i = 0; i < LMAX + 2; ++i) tabs[i].store(nullptr, std::memory_order_relaxed);
  }
  const BandTable& get(int L) {
    BandTable* t = tabs[L].load(std::memory_order_acquire);
    if (t) return *t;
    std::lock_guard<std::mutex> lk(mu);
    t = tabs[L].load(std::memory_order_relaxed);
    if (t) return *t;
    BandTable* T = new BandTable(build_band(L, n_y, force, WINDOW_MAX));
    off[L] = (int32_t)eb.size();
    for (int i = 0; i <= L; ++i) {
      ab.push_back((int16_t)(i ? T->row[i].a : 1));
      ab.push_back((int16_t)(i ? T->row[i].b : 0));
      eb.push_back(T->eb[i]);
      sb.push_back(T->sb[i]);
      for (int w = 0; w < 4; ++w) mask.push_back(T->mask[(size_t)i * 4 + w]);
    }
    tabs[L].store(T, std::memory_order_release);
    return *T;
  }
};
BandHost& global_bands(int n_y, int force) {
  static std::mutex mu;
  static std::map<int, std::unique_ptr<BandHost>> all;  // entries live for the life of the process
  std::lock_guard<std::mutex> lk(mu);
  std::unique_ptr<BandHost>& e = all[n_y * 3 + force];
  if (!e) e.reset(new BandHost(n_y, force));
  return *e;
}

struct BandStore {  // one context's device mirror of the tables of one (n_y, force)
  int n_y = -1;
  int force = 0;
  BandHost* H = nullptr;
  DevBuf d_off, d_ab, d_eb, d_sb, d_mask;
  size_t up_rows = 0;  // table rows already on the device (tables are append-only)
  void reset(int ny, int f) {
    n_y = ny;
    force = f;
    H = &global_bands(ny, f);
    up_rows = 0;
  }
  const BandTable& get(int L) { return H->get(L); }
  // copies the rows added (by any context) since the last upload, everything after a reallocation.  The pageable
  // sources are consumed when cudaMemcpyAsync returns, so the lock is held only for the duration of the calls.
  bool upload(nw_ctx* c, cudaStream_t st) {
    (void)st;
    std::lock_guard<std::mutex> lk(H->mu);
    const size_t rows = H->eb.size();
    if (rows == up_rows) return false;
    if (rows * 4 > d_ab.cap || rows * 2 > d_eb.cap || rows > d_sb.cap || rows * 16 > d_mask.cap) {
      d_ab.ensure(rows * 4 * 2);  // slack: the next few tables append in place
      d_eb.ensure(rows * 2 * 2);
      d_sb.ensure(rows * 2);
      d_mask.ensure(rows * 16 * 2);
      up_rows = 0;
    }
    h2d(c, d_off.ensure(H->off.size() * 4), H->off.data(), H->off.size() * 4);
    const size_t r0 = up_rows, n = rows - up_rows;
    h2d(c, d_ab.as<int16_t>() + r0 * 2, H->ab.data() + r0 * 2, n * 4);
    h2d(c, d_eb.as<int16_t>() + r0, H->eb.data() + r0, n * 2);
    h2d(c, d_sb.as<uint8_t>() + r0, H->sb.data() + r0, n);
    h2d(c, d_mask.as<uint32_t>() + r0 * 4, H->mask.data() + r0 * 4, n * 16);
    up_rows = rows;
    return true;
  }
};

}  // namespace

struct nw_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;        // the root's own DP runs here, beside level 1 (nothing waits for it until the report)
  cudaEvent_t side_go = nullptr, side_done = nullptr;
  cudaEvent_t ver_go = nullptr, ver_done = nullptr;  // dedup verifier of a level, on the side stream beside the level's DP
  std::vector<U4> h_pow;
  DevBuf d_pow;
  uint64_t fp_reseed = 0;  // hash bases in use (bumped when a search's dedup verifier reports a collision)
  // regions of the current search (regionfile mode runs many independent regions through one pipeline)
  std::vector<RegionHost> RH;
  std::vector<RegionDev> RD;          // host copies; d_regions is the device array the kernels index by region id
  std::vector<BandStore*> rbands;     // corridor tables of every region's n_y
  DevBuf d_regions, d_ctx_all, d_small_all, d_ms_all;
  int n_regions = 0, nx_max = 0, ny_max = 0;
  uint32_t ctx_max = 0;
  bool regions_dirty = true;          // RD changed (band pointers) and must be re-uploaded
  std::map<int, BandStore> band_cache;  // corridor tables per (n_y, force), kept across regions
  // dedup table
  DevBuf table;
  uint64_t table_cap = 0;
  DevBuf table_prev;                 // the table a level outgrew: released when the next one is retired (no host sync)
  // per-level scratch
  DevBuf cnt, off, bsum, slot, owner, is_new, newrank, nrank, new_len, hist, bstart, cursor, work, scratch, flag, pre, sel, keys,
      misc, edgebuf, cta_rid, new_rid, best, gather, binflag, binrank, binsum, binrec, rootwork, rflag;
  nw::Exchange* xch = nullptr;  // NCCL communicator of the frontier-sharded search (nw_dist_init), owned by the context
  void* pinned = nullptr;  // pinned staging area for D2H scalars / histograms (grown on demand)
  size_t pinned_bytes = 0;
  // staging of the region pre-pass (nw_region.cu): pinned host buffer + device buffer, grown on demand, kept
  void* rg_pinned = nullptr;
  size_t rg_pinned_bytes = 0;
  void* rg_dev = nullptr;
  size_t rg_dev_bytes = 0;
  DevPool pool;
  int launches = 0;
  int64_t h2d_bytes = 0, d2h_bytes = 0;
  // device time of the dominant (DP) kernels, measured with events on the launching stream
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> score_events;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> event_pool;
  int64_t score_work_items = 0;
  nw_ctx() = default;
};

namespace {
void h2d(nw_ctx* c, void* dst, const void* src, size_t n) {
  CK(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, c->stream));
  c->h2d_bytes += (int64_t)n;
}
void d2h(nw_ctx* c, void* dst, const void* src, size_t n) {
  CK(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, c->stream));
  c->d2h_bytes += (int64_t)n;
}
}  // namespace

struct nw_result {
  nw_stats stats{};
  std::vector<float> t_score;
  std::vector<int32_t> t_k3;
  std::vector<int64_t> t_cost, t_total, t_id;
  std::vector<int64_t> t_off;  // into t_moves (triples)
  std::vector<int32_t> t_moves;
  std::string tsv;
  // keep_ids
  std::vector<float> id_score;
  std::vector<int64_t> id_cost, id_total;
  std::vector<int32_t> id_level;
  std::vector<int32_t> e_child, e_parent, e_moves;
};

namespace {

constexpr size_t PINNED_BYTES = 512 * 1024;

struct Level {
  // frontier expanded at this level
  int P = 0, LS = 0;
  uint32_t pstride = 0;
  DevBuf fblob, flen, fid, fdup, frid;  // frid: region of every parent (uint16)
  // candidates generated at this level
  uint32_t G = 0, gbase = 0, n_new = 0;
  int32_t id_base = 0;
  DevBuf desc, cand_id, newlist, k3, cost, total, nrid;  // nrid: region of every fresh id (uint16)
  // region boundaries (R+1 entries each): parents, candidate positions, new ranks of every region at this level
  std::vector<uint32_t> fstart, cstart, nstart;
  FrontierDev frontier() const {
    FrontierDev F;
    F.P = P;
    F.LS = LS;
    F.pstride = pstride;
    F.blob = fblob.as<unsigned char>();
    F.len = flen.as<int32_t>();
    F.id = fid.as<int32_t>();
    F.dupok = fdup.as<uint8_t>();
    F.rid = frid.p ? frid.as<uint16_t>() : nullptr;
    F.regions = nullptr;  // filled by Search::frontier_of (device region table of the context)
    return F;
  }
};

struct ProblemRef {
  const int8_t* aln;
  int n_x, n_y;
  const int64_t* len_x;
  const int64_t* len_y;
};

// Builds and uploads the context of every region of a search (packed into a few buffers: one copy each).
void setup_regions(nw_ctx* c, const ProblemRef* probs, int R, int maxdepth, int force) {
  if (R < 1 || R > 65535) fail(NW_ERR_ARG, "a batch holds between 1 and 65535 regions");
  cudaStream_t st = c->stream;
  c->RH.clear();
  c->RH.reserve(R);
  c->RD.assign(R, RegionDev{});
  c->rbands.assign(R, nullptr);
  c->n_regions = R;
  c->nx_max = c->ny_max = 0;
  c->ctx_max = 0;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t ctx_bytes = 0, small_bytes = 0, ms_bytes = 0;
  std::vector<size_t> o_ctx(R), o_rt(R), o_cumr(R), o_dd(R), o_inv(R), o_any(R);
  for (int r = 0; r < R; ++r) {
    try {
      c->RH.push_back(build_region(probs[r].aln, probs[r].n_x, probs[r].n_y, probs[r].len_x, probs[r].len_y, maxdepth));
    } catch (const HostError& e) {
      fail(e.code, (R > 1 ? "region " + std::to_string(r) + ": " : std::string()) + e.what());
    }
    const RegionHost& H = c->RH.back();
    if (H.ctx.size() > 200 * 1024) fail(NW_ERR_RANGE, "region context exceeds shared memory");
    o_ctx[r] = ctx_bytes;
    ctx_bytes += al(H.ctx.size());
    o_rt[r] = small_bytes;
    small_bytes += al(H.rt.size() * 4);
    o_cumr[r] = small_bytes;
    small_bytes += al(H.cumr.size() * 4);
    const size_t msb = al((size_t)(H.n_x + 1) * H.MW * 4);
    o_dd[r] = ms_bytes;
    ms_bytes += msb;
    o_inv[r] = ms_bytes;
    ms_bytes += msb;
    o_any[r] = ms_bytes;
    ms_bytes += al((size_t)(H.n_x + 1) * 4);
    c->nx_max = std::max(c->nx_max, H.n_x);
    c->ny_max = std::max(c->ny_max, H.n_y);
    c->ctx_max = std::max<uint32_t>(c->ctx_max, (uint32_t)H.ctx.size());
  }
  std::vector<unsigned char> h_ctx(ctx_bytes, 0), h_small(small_bytes, 0);
  for (int r = 0; r < R; ++r) {
    const RegionHost& H = c->RH[r];
    memcpy(h_ctx.data() + o_ctx[r], H.ctx.data(), H.ctx.size());
    memcpy(h_small.data() + o_rt[r], H.rt.data(), H.rt.size() * 4);
    memcpy(h_small.data() + o_cumr[r], H.cumr.data(), H.cumr.size() * 4);
  }
  unsigned char* dctx = (unsigned char*)c->d_ctx_all.ensure(ctx_bytes);
  unsigned char* dsm = (unsigned char*)c->d_small_all.ensure(small_bytes);
  unsigned char* dms = (unsigned char*)c->d_ms_all.ensure(ms_bytes);
  h2d(c, dctx, h_ctx.data(), ctx_bytes);
  h2d(c, dsm, h_small.data(), small_bytes);
  CK(cudaMemsetAsync(dms, 0, ms_bytes, st));
  for (int r = 0; r < R; ++r) {
    const RegionHost& H = c->RH[r];
    RegionDev& D = c->RD[r];
    D.n_x = H.n_x;
    D.n_y = H.n_y;
    D.NYP = H.NYP;
    D.PWS = H.PWS;
    D.MW = H.MW;
    D.sum_rt = (int)H.sum_rt;
    D.rid1 = (uint32_t)r + 1;  // leading element of every fingerprint of this region
    D.ctx_bytes = (uint32_t)H.ctx.size();
    D.ctx = dctx + o_ctx[r];
    D.off_upx = H.off_upx;
    D.off_rtrev = H.off_rtrev;
    D.rt_copy_stride = H.rt_copy_stride;
    D.rt = reinterpret_cast<const int32_t*>(dsm + o_rt[r]);
    D.cumr = reinterpret_cast<const int32_t*>(dsm + o_cumr[r]);
    D.ms_dd = reinterpret_cast<uint32_t*>(dms + o_dd[r]);
    D.ms_inv = reinterpret_cast<uint32_t*>(dms + o_inv[r]);
    D.ms_any = reinterpret_cast<uint32_t*>(dms + o_any[r]);
    BandStore& bs = c->band_cache[H.n_y * 3 + force];
    if (bs.n_y != H.n_y || bs.force != force) bs.reset(H.n_y, force);
    c->rbands[r] = &bs;
  }
  c->d_regions.ensure((size_t)R * sizeof(RegionDev));
  h2d(c, c->d_regions.p, c->RD.data(), (size_t)R * sizeof(RegionDev));
  CK(spin_sync(st));  // the packed host buffers go out of scope
  c->regions_dirty = false;
  c->launches += launch_moveset(c->d_regions.as<RegionDev>(), R, c->nx_max, st);
}

// corridor tables may have been (re)allocated: refresh the device pointers of every region
void refresh_region_bands(nw_ctx* c) {
  bool changed = false;
  for (int r = 0; r < c->n_regions; ++r) {
    BandStore* b = c->rbands[r];
    RegionDev& D = c->RD[r];
    if (D.band_off != b->d_off.as<int32_t>() || D.band_ab != b->d_ab.as<int16_t>() || D.band_eb != b->d_eb.as<int16_t>() ||
        D.band_sb != b->d_sb.as<uint8_t>() || D.band_mask != b->d_mask.as<uint32_t>()) {
      D.band_off = b->d_off.as<int32_t>();
      D.band_ab = b->d_ab.as<int16_t>();
      D.band_eb = b->d_eb.as<int16_t>();
      D.band_sb = b->d_sb.as<uint8_t>();
      D.band_mask = b->d_mask.as<uint32_t>();
      changed = true;
    }
  }
  if (changed) {
    h2d(c, c->d_regions.p, c->RD.data(), (size_t)c->n_regions * sizeof(RegionDev));
    CK(spin_sync(c->stream));
  }
}

void table_reserve(nw_ctx* c, uint64_t need_entries) {
  uint64_t want = 1024;
  // slots >= 1.5 x (ids so far + candidates of the level), rounded up to a power of two: the level's real load stays
  // below ~0.4 because a third or more of the candidates are duplicates.  The table is cleared (0xFF) for every level
  // that outgrows it, and at 2 x the 2 GB clear of the S64 level-3 table alone was 6 % of the search (measured:
  // 5.69 ms at 2 x, 5.53 ms at 1.5 x; NW_TABLE_SLACK_PCT overrides for experiments).
  static const int load_pct = getenv("NW_TABLE_SLACK_PCT") ? std::max(110, atoi(getenv("NW_TABLE_SLACK_PCT"))) : 150;
  while (want * 100 < need_entries * (uint64_t)load_pct) want <<= 1;
  if (want <= c->table_cap) return;
  if (want > (1ull << 32)) fail(NW_ERR_NOMEM, "dedup table would exceed 2^32 slots");
  const size_t bytes = want * TBL_WORDS * 8;
  if (c->table_cap == 0) {  // fresh search: reuse the physical buffer when it is large enough
    c->table_prev.release();
    c->table.ensure(bytes);
    CK(cudaMemsetAsync(c->table.p, 0xFF, bytes, c->stream));
    c->table_cap = want;
    return;
  }
  DevBuf nt;
  nt.ensure(bytes);
  CK(cudaMemsetAsync(nt.p, 0xFF, bytes, c->stream));
  c->launches += launch_rehash(c->table.as<uint64_t>(), c->table_cap, nt.as<uint64_t>(), want - 1, c->stream);
  // the outgrown table is still being read by the rehash: it is kept until the next growth (or the next search)
  // retires it, by which time the host has synchronised with the stream several times -- no sync here (clearing
  // the new table on a second stream under the emit kernel was measured as well: no further gain)
  c->table_prev = std::move(c->table);
  c->table = std::move(nt);
  c->table_cap = want;
}

// Buckets the `n_new` fresh candidates of a level by child length and runs the DP kernels.
// hist (host copy) gives the number of DP-needing candidates per length.
struct ScoreView {
  FrontierDev F;
  uint32_t n_new;
  const uint32_t* newlist;
  const uint64_t* desc;
  int32_t *k3, *cost, *total;
};
ScoreView view_of(const Level& L) {
  return ScoreView{L.frontier(),          L.n_new,           L.newlist.as<uint32_t>(), L.desc.as<uint64_t>(),
                   L.k3.as<int32_t>(),    L.cost.as<int32_t>(), L.total.as<int32_t>()};
}
// Buckets the fresh candidates by (region, child length) and runs the DP kernels.  hist_dev[bucket_key(...)] is the
// device histogram of (region, length, sub-counter) written by k_assign / k_len_hist; lcap = longest child length + 1.
// Only the non-empty bins travel to the host (compacted on the device, in (region, length) order).
// Work-list layout: kernel class (register window W ascending, generic last) > region > length; every bucket is
// padded to a warp, every region segment of a register-kernel class to a CTA (a CTA stages one region context).
// n_new_deferred != nullptr: the caller has not fetched the number of fresh candidates yet (misc[5], written by the
// scan of the first-discoverer flags); it travels with the bins in the same round trip and is handed back.  Only for
// launches that take the single-copy path below (one region, few bins) -- saves one host round trip per level.
void run_scoring(nw_ctx* c, const ScoreView& Lv, const uint32_t* hist_dev, int lcap, const uint16_t* new_rid_dev, int force,
                 int flags, int* best_dev, int64_t* cells_out /* [n_regions] or nullptr */,
                 uint32_t* n_new_deferred = nullptr, OwnRange own = OwnRange{1, 0, nullptr}) {
  const int R = c->n_regions;
  uint32_t n_new = Lv.n_new;
  const size_t nbins = (size_t)R * lcap * NSUB;
  const uint32_t nb = (uint32_t)((size_t)R * lcap);
  cudaStream_t st = c->stream;
  static const bool dbg = getenv("NW_DEBUG_TIMING") != nullptr;
  auto tnow = [] { return std::chrono::steady_clock::now(); };
  auto tms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
  };
  const auto T0 = tnow();
  if (cells_out) std::fill(cells_out, cells_out + R, (int64_t)0);
  // non-empty bins -> host
  std::vector<BinRec> bins;
  if (nb <= BINS_SMALL_LIMIT) {  // single searches: one kernel, one copy, one sync
    BinRec* brec = (BinRec*)c->binrec.ensure((size_t)(nb + 1) * sizeof(BinRec), true);
    c->launches += launch_bins_small(hist_dev, nb, brec, st);
    BinRec* pin = reinterpret_cast<BinRec*>(static_cast<unsigned char*>(c->pinned) + 4096);
    if (n_new_deferred) d2h(c, c->pinned, c->misc.p, 32);
    d2h(c, pin, brec, (size_t)(nb + 1) * sizeof(BinRec));
    CK(spin_sync(st));
    if (n_new_deferred) {
      n_new = reinterpret_cast<const uint32_t*>(c->pinned)[5];
      *n_new_deferred = n_new;
    }
    const uint32_t n = pin[0].bin;
    if (!n) return;
    bins.assign(pin + 1, pin + 1 + n);
  } else {
    if (n_new_deferred) fail(NW_ERR_ARG, "internal: deferred fresh count on the multi-copy bucket path");
    uint32_t* bflag = (uint32_t*)c->binflag.ensure((size_t)nb * 4, true);
    uint32_t* brank = (uint32_t*)c->binrank.ensure((size_t)nb * 4, true);
    uint32_t* bsumn = (uint32_t*)c->binsum.ensure((size_t)nb * 4, true);
    uint32_t* bsum = (uint32_t*)c->bsum.ensure(scan_scratch_items(nb) * 4, true);
    int32_t* misc = (int32_t*)c->misc.p;
    c->launches += launch_bin_flags(hist_dev, nb, bflag, bsumn, st);
    c->launches += launch_scan(bflag, brank, nb, bsum, (uint32_t*)(misc + 13), st);
    BinRec* brec = (BinRec*)c->binrec.ensure((size_t)std::min<uint64_t>(nb, (uint64_t)n_new + 1) * sizeof(BinRec), true);
    c->launches += launch_bin_compact(bflag, brank, bsumn, nb, brec, st);
    uint32_t* pin = (uint32_t*)c->pinned;
    d2h(c, pin, misc, 64);
    CK(spin_sync(st));
    const uint32_t n = pin[13];
    if (!n) return;
    bins.resize(n);
    d2h(c, bins.data(), brec, (size_t)n * sizeof(BinRec));
    CK(spin_sync(st));
  }
  struct Bucket {
    int rid, L, cls;  // cls: window W (32-bit register kernel), W | CLS_PAIR (packed register kernel) or 0 (generic)
    uint32_t n, bin;
  };
  // NW_DP_PAIR=0 keeps every bucket on the 32-bit register kernel (experiments, A/B parity tests)
  static const bool pair_on = !getenv("NW_DP_PAIR") || atoi(getenv("NW_DP_PAIR")) != 0;
  auto cls_smem = [&](int cls, int L) {
    return (cls & CLS_PAIR) ? score_pair_smem(cls & ~CLS_PAIR, c->ctx_max, L) : score_fast_smem(cls, c->ctx_max, L);
  };
  std::vector<Bucket> bk(bins.size());
  std::vector<uint32_t> cls_count;  // counting sort by kernel class: register windows ascending, generic (0) last
  auto cls_key = [](int cls) { return cls ? (uint32_t)cls : 0x2001u; };
  cls_count.assign(0x2003, 0);
  for (size_t k = 0; k < bins.size(); ++k) {
    const int r = (int)(bins[k].bin / (uint32_t)lcap), L = (int)(bins[k].bin % (uint32_t)lcap);
    const BandTable& T = c->rbands[r]->get(L);
    if (!T.contiguous) fail(NW_ERR_RANGE, "corridor rows are not intervals for this shape (unsupported)");
    int cls = 0;
    if (!(flags & NW_FLAG_GENERIC_DP) && T.fast_ok) cls = score_fast_window(T.max_width);
    if (cls && score_fast_smem(cls, c->ctx_max, L) > 200 * 1024) cls = 0;  // tiles would not fit: generic kernel
    // packed form (two candidates per thread, 16-bit halves): every total of the bucket must fit 15 bits
    if (cls && pair_on && !(flags & NW_FLAG_DP32) && pair_range_ok(c->RH[r], L) && score_pair_smem(cls, c->ctx_max, L) <= 200 * 1024) cls |= CLS_PAIR;
    bk[k] = Bucket{r, L, cls, bins[k].n, bins[k].bin};
    if (cells_out) cells_out[r] += (int64_t)bins[k].n * T.cells;
  }
  {
    // window classes that would not fill a launch of their own are folded into the next wider class present
    // (fold_window_classes, nw_host.h: S150 windows 60 and 64 become one launch per level, regionfile passes need fewer
    // launches, the full-corridor pre-pass of the region stage keeps its narrow classes out of the 112 / 128 windows).
    // NW_CLASS_MERGE_CTAS overrides the wave size (148 x 6 CTAs), NW_NO_CLASS_MERGE disables.
    static const bool no_merge = getenv("NW_NO_CLASS_MERGE") != nullptr;
    static const int merge_ctas = getenv("NW_CLASS_MERGE_CTAS") ? atoi(getenv("NW_CLASS_MERGE_CTAS")) : 148 * 6;
    std::map<int, uint64_t> cta_of;  // fast classes present -> work slots
    for (const Bucket& b : bk)
      if (b.cls) cta_of[b.cls] += (b.n + 31) & ~31u;
    std::map<int, int> into;
    if (!no_merge) into = fold_window_classes(cta_of, (uint64_t)merge_ctas);
    for (Bucket& b : bk) {
      int cls = b.cls;
      while (cls && into.count(cls)) {
        const int up = into[cls];
        if (cls_smem(up, b.L) > 200 * 1024) break;
        cls = up;
      }
      b.cls = cls;
      cls_count[cls_key(cls) + 1]++;
    }
  }
  const auto T1 = tnow();
  {
    bool any = false;
    for (int r = 0; r < R; ++r) any |= c->rbands[r]->upload(c, st);
    if (any) CK(spin_sync(st));
  }
  refresh_region_bands(c);
  const auto T2 = tnow();
  for (size_t k = 1; k < cls_count.size(); ++k) cls_count[k] += cls_count[k - 1];
  std::vector<uint32_t> order(bk.size());  // class-major; (region, length) ascending inside a class (stable)
  for (size_t k = 0; k < bk.size(); ++k) order[cls_count[cls_key(bk[k].cls)]++] = (uint32_t)k;
  struct Range {
    int cls;
    uint32_t begin, end;
    int lmax;
  };
  std::vector<Range> ranges;
  std::vector<uint16_t> cta_rid;  // region of every CTA, over the whole work list (CTA = 128 slots)
  uint32_t pos = 0;
  int prev_rid = -1;
  for (uint32_t idx : order) {
    const Bucket& b = bk[idx];
    const bool new_class = ranges.empty() || ranges.back().cls != b.cls;
    if (new_class || (b.cls && b.rid != prev_rid)) pos = (pos + 127) & ~127u;  // CTA boundary
    if (new_class) ranges.push_back(Range{b.cls, pos, pos, 0});
    prev_rid = b.rid;
    bins[idx].n = pos;  // the sub-counters of one length share one contiguous, warp-padded bucket (k_bucket_fill)
    const uint32_t pad = (b.cls & CLS_PAIR) ? 63u : 31u;  // a warp of the packed kernel owns 64 slots of one length
    pos += (b.n + pad) & ~pad;
    ranges.back().end = pos;
    ranges.back().lmax = std::max(ranges.back().lmax, b.L);
    if (R > 1) {
      const size_t last_cta = (pos + 127) / 128;
      if (cta_rid.size() < last_cta) cta_rid.resize(last_cta, (uint16_t)b.rid);
    }
  }
  const auto T3 = tnow();
  const uint32_t n_work = pos;
  {
    BinRec* brec = c->binrec.as<BinRec>();
    h2d(c, brec, bins.data(), bins.size() * sizeof(BinRec));
    c->launches += launch_bucket_fill(brec, (uint32_t)bins.size(), hist_dev, (uint32_t*)c->bstart.ensure(nbins * 4, true),
                                      (uint32_t*)c->cursor.ensure(nbins * 4, true), st);
  }
  CK(cudaMemsetAsync(c->work.ensure((size_t)n_work * 4, true), 0xFF, (size_t)n_work * 4, st));
  uint16_t* d_cta_rid = nullptr;
  if (R > 1) {
    d_cta_rid = (uint16_t*)c->cta_rid.ensure(cta_rid.size() * 2, true);
    h2d(c, d_cta_rid, cta_rid.data(), cta_rid.size() * 2);
  }
  c->launches += launch_bucket_scatter(c->new_len.as<uint16_t>(), R > 1 ? new_rid_dev : nullptr, c->d_regions.as<RegionDev>(), lcap, n_new,
                                       force, own, c->bstart.as<uint32_t>(), c->cursor.as<uint32_t>(),
                                       c->work.as<uint32_t>(), st);
  // no sync: cudaMemcpyAsync from pageable memory has consumed the host vectors (bins, cta_rid) when it returns
  if (dbg)
    fprintf(stderr, "[nw scoring] buckets %zu, work %u for %u new, ranges %zu: bucket loop %.2f ms, bands %.2f, layout %.2f, scatter %.2f\n",
            bk.size(), n_work, n_new, ranges.size(), tms(T0, T1), tms(T1, T2), tms(T2, T3), tms(T3, tnow()));
  ScoreArgs A{};
  A.F = Lv.F;
  A.F.regions = c->d_regions.as<RegionDev>();
  A.ctx_max = c->ctx_max;
  A.ny_max = c->ny_max;
  A.newlist = Lv.newlist;
  A.desc = Lv.desc;
  A.k3 = Lv.k3;
  A.cost = Lv.cost;
  A.total = Lv.total;
  A.best_k3 = best_dev;
  A.cells = nullptr;
  A.force = force;
  for (const Range& r : ranges) {
    A.work = c->work.as<uint32_t>() + r.begin;
    A.n_work = r.end - r.begin;
    A.lmax = r.lmax;
    A.cta_rid = d_cta_rid ? d_cta_rid + r.begin / 128 : nullptr;
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    if (!c->event_pool.empty()) {
      ev = c->event_pool.back();
      c->event_pool.pop_back();
    } else {
      CK(cudaEventCreate(&ev.first));
      CK(cudaEventCreate(&ev.second));
    }
    CK(cudaEventRecord(ev.first, st));
    if (r.cls) {
      // Form of the register DP for this launch.  The cooperative form (four lanes per candidate) has a quarter of the
      // dependent chain per row, so a launch too small to fill the device finishes sooner (the latency of ONE thread's DP
      // is the floor of the thread form); it executes ~2.7x the instructions per cell, so everything larger stays with
      // one thread per candidate (measured on S150 / S64: profiles/README.md).  NW_DP_COOP=0 / 1 forces one form.
      static const int coop_mode = getenv("NW_DP_COOP") ? atoi(getenv("NW_DP_COOP")) : -1;
      static const int coop_max_ctas = getenv("NW_DP_COOP_MAX_CTAS") ? atoi(getenv("NW_DP_COOP_MAX_CTAS")) : 48;
      const int win = r.cls & ~CLS_PAIR;
      const int cw = score_coop_window(win);
      bool coop = cw && score_coop_smem(cw, c->ctx_max, r.lmax) <= 200 * 1024;
      if (coop && coop_mode == 0) coop = false;
      if (coop && coop_mode < 0) coop = (int)(A.n_work / 128) < coop_max_ctas;
      const int l = coop ? launch_score_coop(cw, A, st)
                         : ((r.cls & CLS_PAIR) ? launch_score_pair(win, A, st) : launch_score_fast(win, A, st));
      if (l < 0) fail(NW_ERR_CUDA, "no register-DP kernel for window " + std::to_string(win));
      c->launches += l;
    } else {
      const size_t sc = (size_t)2 * (c->ny_max + 2) * generic_threads() * 4;
      A.scratch = (int32_t*)c->scratch.ensure(sc);
      c->launches += launch_score_generic(A, st);
    }
    CK(cudaEventRecord(ev.second, st));
    c->score_events.push_back(ev);
  }
  CK(cudaGetLastError());
}

// sums and recycles the DP-kernel event pairs (stream must be idle)
void collect_score_time(nw_ctx* c, float* ms_out, int* launches_out) {
  float tot = 0;
  for (auto& ev : c->score_events) {
    float ms = 0;
    cudaEventElapsedTime(&ms, ev.first, ev.second);
    tot += ms;
    c->event_pool.push_back(ev);
  }
  if (ms_out) *ms_out = tot;
  if (launches_out) *launches_out = (int)c->score_events.size();
  c->score_events.clear();
}

struct Search {
  nw_ctx* c;
  nw_opts o;
  std::vector<Level> lv;  // lv[0] = root pseudo-level, lv[d] = BFS level d
  int32_t n_ids = 0;
  uint32_t gtotal = 0;
  int R = 1;  // regions searched together (regionfile mode); region r has root id r+1
  std::vector<int> root_k3, root_cost, root_total, best_k3;  // per region
  std::vector<nw_stats> rst;                                   // per region
  nw_stats st{};                                               // whole batch (timing, launches)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_ver = nullptr;
  std::vector<cudaEvent_t> lev;

  // frontier sharding (nw_solve_sharded*): every rank runs the same search, the DP of a level is divided by new-rank
  // ranges and completed by one all-gather (nw_exchange.h).  xch == nullptr is the plain single-GPU search.
  Exchange* xch = nullptr;
  int rank = 0, world = 1;
  uint64_t fp_reseed0 = 0;  // hash bases of the first attempt of this call (NW_FLAG_DEBUG_WEAK_FP_ONCE)
  uint32_t edge_cap_hint = 0, id_cap_hint = 0;
  int level_maxlen = 0;  // longest duplication child of the level in flight
  explicit Search(nw_ctx* ctx, const nw_opts& opts) : c(ctx), o(opts) {}
  ~Search() {
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (ev_ver) cudaEventDestroy(ev_ver);
    for (auto e : lev) cudaEventDestroy(e);
  }

  uint32_t* pin_u32() { return reinterpret_cast<uint32_t*>(c->pinned); }
  template <typename F>
  void xcall(F&& f) {  // transport errors of the sharded search surface like CUDA errors
    try {
      f();
    } catch (const ExchangeError& e) {
      fail(NW_ERR_CUDA, "sharded search exchange: " + e.msg);
    }
  }

  FrontierDev frontier_of(const Level& L) const {
    FrontierDev F = L.frontier();
    F.regions = c->d_regions.as<RegionDev>();
    if (R == 1) F.rid = nullptr;
    return F;
  }
  uint32_t* pinned(size_t bytes) {  // growable pinned staging area
    if (bytes > c->pinned_bytes) {
      if (c->pinned) cudaFreeHost(c->pinned);
      c->pinned = nullptr;
      c->pinned_bytes = 0;
      const size_t want = std::max(bytes + bytes / 2, (size_t)512 * 1024);
      CK(cudaMallocHost(&c->pinned, want));
      c->pinned_bytes = want;
    }
    return reinterpret_cast<uint32_t*>(c->pinned);
  }
  // dst_host[i] = src_dev[idx[i]] (idx[i] < limit) else fill; one tiny H2D + kernel; the D2H is left to the caller
  // (results land at gather_dev[0..n))
  void gather_async(const uint32_t* src_dev, const std::vector<uint32_t>& idx, uint32_t limit, uint32_t fill) {
    const uint32_t n = (uint32_t)idx.size();
    uint32_t* g = (uint32_t*)c->gather.ensure((size_t)n * 8, true);
    h2d(c, g + n, idx.data(), (size_t)n * 4);
    c->launches += launch_gather_u32(src_dev, g + n, n, limit, fill, g, c->stream);
  }

  void init_root() {
    cudaStream_t s = c->stream;
    R = c->n_regions;
    root_k3.assign(R, 0);
    root_cost.assign(R, -1);
    root_total.assign(R, 0);
    best_k3.assign(R, 0);
    rst.assign(R, nw_stats{});
    if (c->nx_max > LMAX) fail(NW_ERR_RANGE, "n_x exceeds the supported index length");
    lv.clear();
    lv.emplace_back();  // level 0 (root pseudo level)
    lv.emplace_back();  // level 1 frontier = the R roots
    Level& L0 = lv[0];
    Level& L1 = lv[1];
    L1.P = R;
    L1.LS = round_ls(c->nx_max);
    L1.pstride = blob_stride(L1.LS);
    std::vector<unsigned char> blob((size_t)R * L1.pstride);
    std::vector<int32_t> len(R), id(R);
    std::vector<uint8_t> dupok(R, (1 <= o.maxdup) ? 1 : 0);  // duplication_count(root) == 1 (solver.jl:611)
    std::vector<uint16_t> rid(R);
    std::vector<uint64_t> d0(R);
    std::vector<uint32_t> iota(R);
    std::vector<uint16_t> rl(R);
    std::vector<int16_t> root;
    for (int r = 0; r < R; ++r) {
      const int n_x = c->RH[r].n_x;
      root.resize(n_x);
      for (int i = 0; i < n_x; ++i) root[i] = (int16_t)(i + 1);
      fill_blob(blob.data() + (size_t)r * L1.pstride, L1.LS, root.data(), n_x, c->h_pow.data(), c->RD[r].rid1);
      len[r] = n_x;
      id[r] = r + 1;
      rid[r] = (uint16_t)r;
      d0[r] = pack_desc((uint32_t)r, 0, 0, KIND_ID);
      iota[r] = (uint32_t)r;
      rl[r] = (uint16_t)n_x;
    }
    h2d(c, L1.fblob.ensure(blob.size()), blob.data(), blob.size());
    h2d(c, L1.flen.ensure((size_t)R * 4), len.data(), (size_t)R * 4);
    h2d(c, L1.fid.ensure((size_t)R * 4), id.data(), (size_t)R * 4);
    h2d(c, L1.fdup.ensure((size_t)R), dupok.data(), (size_t)R);
    h2d(c, L1.frid.ensure((size_t)R * 2), rid.data(), (size_t)R * 2);
    L1.fstart.resize(R + 1);
    for (int r = 0; r <= R; ++r) L1.fstart[r] = (uint32_t)r;
    // level 0: one pseudo candidate per region (the root itself, KIND_ID over its level-1 frontier slot), ids 1..R
    L0.G = (uint32_t)R;
    L0.gbase = 0;
    L0.n_new = (uint32_t)R;
    L0.id_base = 1;
    L0.fstart = L0.cstart = L0.nstart = L1.fstart;
    h2d(c, L0.desc.ensure((size_t)R * 8), d0.data(), (size_t)R * 8);
    h2d(c, L0.cand_id.ensure((size_t)R * 4), id.data(), (size_t)R * 4);
    h2d(c, L0.newlist.ensure((size_t)R * 4), iota.data(), (size_t)R * 4);
    h2d(c, L0.nrid.ensure((size_t)R * 2), rid.data(), (size_t)R * 2);
    h2d(c, c->new_len.ensure((size_t)R * 2, true), rl.data(), (size_t)R * 2);
    std::vector<int32_t> k3i(R, 0), ci(R, -1), ti(R, 0);
    h2d(c, L0.k3.ensure((size_t)R * 4), k3i.data(), (size_t)R * 4);
    h2d(c, L0.cost.ensure((size_t)R * 4), ci.data(), (size_t)R * 4);
    h2d(c, L0.total.ensure((size_t)R * 4), ti.data(), (size_t)R * 4);
    n_ids = R;
    gtotal = (uint32_t)R;
    // dedup table
    c->table_cap = 0;  // logical capacity; the physical buffer is kept across solves
    table_reserve(c, (uint64_t)(1 << 15) + 4 * (uint64_t)R);
    c->launches += launch_insert_roots(L1.fblob.as<unsigned char>(), L1.pstride, L1.LS, L1.flen.as<int32_t>(), R,
                                       c->table.as<uint64_t>(), c->table_cap - 1, s);
    int32_t* misc = (int32_t*)c->misc.ensure(256);
    CK(cudaMemsetAsync(misc, 0, 256, s));
    CK(cudaMemsetAsync(c->best.ensure((size_t)(R + 1) * 4), 0, (size_t)(R + 1) * 4, s));
    CK(spin_sync(s));
    // score the roots (solver.jl:599) -- a root's score never updates `best` (:606): dummy best array
    const int lcap = c->nx_max + 1;
    root_pending = false;
    if (R == 1 && score_root_aside(L0, L1)) return;  // one region: the root's DP runs beside level 1 (collected by run())
    {
      const size_t nbins = (size_t)R * lcap * NSUB;
      uint32_t* hist = (uint32_t*)c->hist.ensure(nbins * 4, true);
      CK(cudaMemsetAsync(hist, 0, nbins * 4, s));
      c->launches += launch_len_hist(c->new_len.as<uint16_t>(), L0.nrid.as<uint16_t>(), c->d_regions.as<RegionDev>(), lcap,
                                     (uint32_t)R, 0, hist, s);
      ScoreView view = view_of(L0);  // candidates of level 0 against the frontier of level 1
      view.F = frontier_of(L1);
      int* dummy_best = (int*)c->flag.ensure((size_t)(R + 1) * 4, true);
      CK(cudaMemsetAsync(dummy_best, 0, (size_t)(R + 1) * 4, s));
      run_scoring(c, view, hist, lcap, L0.nrid.as<uint16_t>(), 0, o.flags, dummy_best, nullptr);
    }
    collect_roots();
  }
  void collect_roots() {
    cudaStream_t s = c->stream;
    Level& L0 = lv[0];
    std::vector<int32_t> k3i(R), ci(R), ti(R);
    d2h(c, k3i.data(), L0.k3.p, (size_t)R * 4);
    d2h(c, ci.data(), L0.cost.p, (size_t)R * 4);
    d2h(c, ti.data(), L0.total.p, (size_t)R * 4);
    CK(spin_sync(s));
    for (int r = 0; r < R; ++r) {
      root_k3[r] = k3i[r];
      root_cost[r] = ci[r];
      root_total[r] = ti[r];
    }
  }
  // One region: the root is a single candidate, and one DP of a 150-block index is ~70 us of pure latency that nothing
  // in the search depends on (the root's score enters the report only).  Its launch goes to the context's side stream
  // behind the uploads; run() joins it before the report.  Returns false when the root needs the generic kernel.
  bool root_pending = false;
  bool score_root_aside(Level& L0, Level& L1) {
    cudaStream_t s = c->stream;
    const int L = c->RH[0].n_x;
    if (early_exit(L, c->RH[0].n_y)) return false;  // settled without a DP (k3 0, cost -1: the arrays' initial values)
    const BandTable& T = c->rbands[0]->get(L);
    if (!T.contiguous) fail(NW_ERR_RANGE, "corridor rows are not intervals for this shape (unsupported)");
    int cls = 0;
    if (!(o.flags & NW_FLAG_GENERIC_DP) && T.fast_ok) cls = score_fast_window(T.max_width);
    if (!cls || score_fast_smem(cls, c->ctx_max, L) > 200 * 1024) return false;
    if (c->rbands[0]->upload(c, s)) CK(spin_sync(s));
    refresh_region_bands(c);
    uint32_t work[32];
    work[0] = 0;
    for (int k = 1; k < 32; ++k) work[k] = WORK_PAD;
    h2d(c, c->rootwork.ensure(sizeof work + 16), work, sizeof work);
    int* dummy_best = reinterpret_cast<int*>(c->rootwork.as<unsigned char>() + sizeof work);
    CK(cudaMemsetAsync(dummy_best, 0, 8, s));
    CK(cudaEventRecord(c->side_go, s));
    CK(cudaStreamWaitEvent(c->side, c->side_go, 0));
    ScoreArgs A{};
    A.F = frontier_of(L1);
    A.ctx_max = c->ctx_max;
    A.ny_max = c->ny_max;
    A.newlist = L0.newlist.as<uint32_t>();
    A.desc = L0.desc.as<uint64_t>();
    A.k3 = L0.k3.as<int32_t>();
    A.cost = L0.cost.as<int32_t>();
    A.total = L0.total.as<int32_t>();
    A.best_k3 = dummy_best;
    A.work = c->rootwork.as<uint32_t>();
    A.n_work = 32;
    A.lmax = L;
    A.cta_rid = nullptr;
    {
      // one candidate: the cooperative form has a quarter of the latency (NW_DP_COOP=0 keeps the thread form)
      static const int coop_mode = getenv("NW_DP_COOP") ? atoi(getenv("NW_DP_COOP")) : -1;
      const int cw = score_coop_window(cls);
      const bool coop = coop_mode != 0 && cw && score_coop_smem(cw, c->ctx_max, L) <= 200 * 1024;
      if ((coop ? launch_score_coop(cw, A, c->side) : launch_score_fast(cls, A, c->side)) < 0)
        fail(NW_ERR_CUDA, "no register-DP kernel for window " + std::to_string(cls));
    }
    c->launches += 1;
    CK(cudaEventRecord(c->side_done, c->side));
    CK(cudaGetLastError());
    root_pending = true;
    return true;
  }
  void join_root() {
    if (!root_pending) return;
    CK(cudaStreamWaitEvent(c->stream, c->side_done, 0));
    root_pending = false;
    collect_roots();
  }

  // dedup verifier: element-wise comparison of every merged candidate of level d with its owner.  No host wait: the
  // mismatch counter (misc[12], cleared by init_root) is read once when the search ends (check_dedup).
  bool verify() const {
    static const bool env_off = getenv("NW_NO_VERIFY") != nullptr;  // diagnostics: what the verifier costs in a whole run
    return !(o.flags & NW_FLAG_NO_VERIFY) && !env_off;
  }
  void verify_level(int d, const ExpandArgs& E, const uint32_t* fresh_mask, const uint32_t* owner, uint32_t G) {
    int32_t* misc = (int32_t*)c->misc.p;
    VerifyTable V{};
    V.n_levels = d + 1;
    for (int l = 0; l <= d; ++l) {
      const Level& fl = lv[l == 0 ? 1 : l];  // the root pseudo-candidate lives in level 1's frontier
      V.gbase[l] = lv[l].gbase;
      V.desc[l] = lv[l].desc.as<uint64_t>();
      V.blob[l] = fl.fblob.as<unsigned char>();
      V.pstride[l] = fl.pstride;
      V.len[l] = fl.flen.as<int32_t>();
    }
    V.gbase[d + 1] = lv[d].gbase + G;
    // beside the rest of the level (id assignment, bucketing, DP) on the side stream: the verifier is bound by memory
    // latency, the DP by the integer pipe.  It reads the level's descriptors, the first-discoverer mask and the owner
    // positions; the next level waits for it before it overwrites the latter two (verify_join).
    // (ver_go was recorded when the mask was complete; this launch is issued after the level's DP so that the DP is
    // already running when the verifier reaches the device)
    CK(cudaStreamWaitEvent(c->side, c->ver_go, 0));
    // sharded search: every rank checks its slice of the level's candidates; the counts are summed in check_dedup
    uint32_t first = 0, end = G;
    if (world > 1) {
      const uint64_t per = ((uint64_t)G + world - 1) / world;
      first = (uint32_t)std::min<uint64_t>(G, (per * rank) & ~31ull);
      end = rank + 1 == world ? G : (uint32_t)std::min<uint64_t>(G, (per * (rank + 1)) & ~31ull);
    }
    c->launches += launch_verify_dups(E, V, fresh_mask, owner, first, end, (unsigned int*)(misc + 12), c->side);
    CK(cudaEventRecord(c->ver_done, c->side));
    if (!ev_ver) CK(cudaEventCreate(&ev_ver));
    CK(cudaEventRecord(ev_ver, c->side));  // timed twin of ver_done: the device time of the search includes the verifier
    ver_pending = true;
  }
  bool ver_pending = false;
  // issue point of the verifier: right behind the scan (it then fills the gaps of id assignment, the bucket round trip and
  // the scatter, and waits out the DP) or behind the DP launch
  bool verify_early = getenv("NW_VERIFY_LATE") == nullptr;
  void verify_join() {  // orders the main stream behind the verifier in flight (no host wait)
    if (!ver_pending) return;
    CK(cudaStreamWaitEvent(c->stream, c->ver_done, 0));
    ver_pending = false;
  }
  // Called once per search, after the report has been extracted: the verifier of the last level runs on the side stream
  // beside the report's kernels and host work.  A search is only handed out when this passes; its device time is the
  // later of the two streams' ends.
  void check_dedup() {
    if (!verify()) return;
    verify_join();
    d2h(c, pin_u32(), c->misc.p, 64);
    CK(spin_sync(c->stream));
    if (ev_ver && ev0) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, ev0, ev_ver) == cudaSuccess && ms > st.total_ms) {
        st.total_ms = ms;
        for (auto& q : rst) q.total_ms = ms;
      }
      cudaGetLastError();
    }
    int64_t bad = pin_u32()[12];
    if (xch) xcall([&] { xch->allreduce_sum(&bad, 1, c->stream); });  // every rank reaches the same verdict
    if (bad)
      fail(NW_ERR_COLLISION, std::to_string(bad) + " groups of merged candidates differ from the index they were "
                             "merged into (fingerprint collision)");
  }

  // one BFS level: expand lv[d]'s frontier (all regions together).  Returns false when nothing new was found.
  bool run_level(int d) {
    cudaStream_t s = c->stream;
    Level& L = lv[d];
    int32_t* misc = (int32_t*)c->misc.p;  // [2]=maxlen, [4..]=totals
    for (int r = 0; r < R; ++r) rst[r].frontier[d - 1] = L.fstart[r + 1] - L.fstart[r];
    L.cstart.assign(R + 1, 0);
    L.nstart.assign(R + 1, 0);
    if (L.P == 0) return false;
    static const bool dbg = getenv("NW_DEBUG_TIMING") != nullptr;
    auto tnow = [] { return std::chrono::steady_clock::now(); };
    auto tms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return std::chrono::duration<double, std::milli>(b - a).count();
    };
    const auto T0 = tnow();
    ExpandArgs E{};
    E.F = frontier_of(L);
    E.nx_max = c->nx_max;
    E.cnt = (uint32_t*)c->cnt.ensure((size_t)L.P * 4, true);
    E.maxlen = misc + 2;
    E.force_pairs = (o.flags & NW_FLAG_PAIR_EXPAND) ? 1 : 0;
    E.weak_fp = ((o.flags & NW_FLAG_DEBUG_WEAK_FP) || ((o.flags & NW_FLAG_DEBUG_WEAK_FP_ONCE) && c->fp_reseed == fp_reseed0)) ? 1 : 0;
    CK(cudaMemsetAsync(misc + 2, 0, 4, s));
    c->launches += launch_expand(false, E, s);
    uint32_t* off = (uint32_t*)c->off.ensure((size_t)L.P * 4, true);
    uint32_t* bsum = (uint32_t*)c->bsum.ensure(scan_scratch_items((uint64_t)L.P) * 4, true);
    c->launches += launch_scan(E.cnt, off, (uint64_t)L.P, bsum, (uint32_t*)(misc + 4), s);
    if (R > 1) gather_async(off, L.fstart, (uint32_t)L.P, 0xffffffffu);  // first candidate of every region
    d2h(c, pin_u32(), misc, 32);
    if (R > 1) d2h(c, pinned((size_t)(R + 1) * 4 + 256) + 64, c->gather.p, (size_t)(R + 1) * 4);
    CK(spin_sync(s));
    CK(cudaGetLastError());
    const auto T1 = tnow();
    const uint32_t G = pin_u32()[4];
    const int maxlen = (int)pin_u32()[2];
    for (int r = 0; r <= R; ++r) {
      const uint32_t v = R > 1 ? pin_u32()[64 + r] : (r == 0 ? 0u : G);
      L.cstart[r] = v == 0xffffffffu ? G : v;
    }
    if (maxlen > LMAX) fail(NW_ERR_RANGE, "a child index exceeds the supported length");
    if ((uint64_t)gtotal + G >= 0xfffffff0ull || G >= 0x7ffffff0u) fail(NW_ERR_NOMEM, "more than 2^32 candidates");
    L.G = G;
    L.gbase = gtotal;
    L.id_base = n_ids + 1;
    for (int r = 0; r < R; ++r) rst[r].generated[d - 1] = L.cstart[r + 1] - L.cstart[r];
    level_maxlen = maxlen;
    if (G == 0) {
      L.n_new = 0;
      return false;
    }
    E.off = off;
    E.desc = (uint64_t*)L.desc.ensure((size_t)G * 8);
    E.slot = (uint32_t*)c->slot.ensure((size_t)G * 4, true);
    E.gbase = L.gbase;
    E.pw = c->d_pow.as<U4>();
    table_reserve(c, (uint64_t)n_ids + G);
    E.table = c->table.as<uint64_t>();
    E.tmask = c->table_cap - 1;
    c->launches += launch_expand(true, E, s);
    c->launches += launch_hash_insert(E, G, s);
    // first-discoverer flags as a bit mask + exclusive scan of the words' popcounts (new-rank of any candidate)
    verify_join();
    const uint32_t nwords = (G + 31) / 32;
    const uint32_t nwords4 = (nwords + 3u) & ~3u;  // the scan reads its input as 16-byte vectors
    uint32_t* is_new = (uint32_t*)c->is_new.ensure((size_t)nwords4 * 8 + 16, true);  // [mask words | word counts]
    uint32_t* wcount = is_new + nwords4;
    uint32_t* owner = (uint32_t*)c->owner.ensure((size_t)G * 4, true);
    c->launches += launch_resolve(E, G, is_new, wcount, owner, s);
    uint32_t* newrank = (uint32_t*)c->newrank.ensure((size_t)nwords * 4 + 16, true);  // word prefixes
    bsum = (uint32_t*)c->bsum.ensure(scan_scratch_items(nwords) * 4, true);
    c->launches += launch_scan(wcount, newrank, nwords, bsum, (uint32_t*)(misc + 5), s);
    if (verify()) {  // everything the verifier reads is complete here
      CK(cudaEventRecord(c->ver_go, s));
      if (verify_early) verify_level(d, E, is_new, owner, G);
    }
    const int lcap = std::min(std::max(maxlen, L.LS), LMAX) + 1;  // longest possible child of this level, + 1
    // single search with few (length) bins: the fresh count is fetched together with the DP buckets (one host round
    // trip less per level); the per-fresh arrays are then sized by the candidate count
    static const bool no_defer = getenv("NW_NO_DEFER") != nullptr;
    const bool defer = R == 1 && (size_t)lcap <= BINS_SMALL_LIMIT && !no_defer;
    uint32_t n_new = G;  // upper bound until known
    if (!defer) {
      if (R > 1) {  // first fresh rank of every region
        const uint32_t n = (uint32_t)L.cstart.size();
        uint32_t* g = (uint32_t*)c->gather.ensure((size_t)n * 8, true);
        h2d(c, g + n, L.cstart.data(), (size_t)n * 4);
        c->launches += launch_gather_rank(is_new, newrank, g + n, n, G, 0xffffffffu, g, s);
      }
      d2h(c, pin_u32(), misc, 32);
      if (R > 1) d2h(c, pin_u32() + 64, c->gather.p, (size_t)(R + 1) * 4);
      CK(spin_sync(s));
      CK(cudaGetLastError());
      n_new = pin_u32()[5];
      for (int r = 0; r <= R; ++r) {
        const uint32_t v = R > 1 ? pin_u32()[64 + r] : (r == 0 ? 0u : n_new);
        L.nstart[r] = v == 0xffffffffu ? n_new : v;
      }
      L.n_new = n_new;
      for (int r = 0; r < R; ++r) rst[r].scored[d - 1] = L.nstart[r + 1] - L.nstart[r];
    }
    const auto T2 = tnow();
    // ids / edges / histogram of (region, child length)
    const size_t nbins = (size_t)R * lcap * NSUB;
    if (nbins > ((size_t)64 << 20)) fail(NW_ERR_NOMEM, "too many (region, length) buckets in one batch");
    AssignArgs A{};
    A.LT.n_levels = d;
    for (int l = 0; l < d; ++l) {
      A.LT.gbase[l] = lv[l].gbase;
      A.LT.cand_id[l] = lv[l].cand_id.as<int32_t>();
    }
    A.LT.gbase[d] = L.gbase;
    A.slot = E.slot;
    A.table = E.table;
    A.owner = owner;
    A.fresh_mask = is_new;
    A.wprefix = newrank;
    A.desc = E.desc;
    A.front_len = L.flen.as<int32_t>();
    A.front_rid = R > 1 ? L.frid.as<uint16_t>() : nullptr;
    A.regions = c->d_regions.as<RegionDev>();
    A.lcap = lcap;
    A.gbase = L.gbase;
    A.G = G;
    A.id_base = L.id_base;
    A.force = 0;
    const OwnRange own{world, rank, (const uint32_t*)(misc + 5)};
    A.own = own;
    A.cand_id = (int32_t*)L.cand_id.ensure((size_t)G * 4);
    A.newlist = (uint32_t*)L.newlist.ensure((size_t)std::max<uint32_t>(n_new, 1) * 4);
    A.new_len = (uint16_t*)c->new_len.ensure((size_t)std::max<uint32_t>(n_new, 1) * 2, true);
    A.new_rid = (uint16_t*)L.nrid.ensure((size_t)std::max<uint32_t>(n_new, 1) * 2);
    // sharded search: the score arrays are exchanged in place as `world` segments of own_chunk items
    const size_t n_score = world > 1 ? (size_t)own_chunk(n_new, world) * world : std::max<uint32_t>(n_new, 1);
    A.k3 = (int32_t*)L.k3.ensure(n_score * 4);
    A.cost = (int32_t*)L.cost.ensure(n_score * 4);
    A.total = (int32_t*)L.total.ensure(n_score * 4);
    A.hist = (uint32_t*)c->hist.ensure(nbins * 4, true);
    CK(cudaMemsetAsync(A.hist, 0, nbins * 4, s));
    c->launches += launch_assign(A, s);
    CK(cudaGetLastError());
    const auto T3 = tnow();
    std::vector<int64_t> cells(R, 0);
    run_scoring(c, ScoreView{frontier_of(L), defer ? 0u : L.n_new, L.newlist.as<uint32_t>(), L.desc.as<uint64_t>(),
                             L.k3.as<int32_t>(), L.cost.as<int32_t>(), L.total.as<int32_t>()},
                A.hist, lcap, L.nrid.as<uint16_t>(), 0, o.flags, c->best.as<int>(), cells.data(), defer ? &n_new : nullptr, own);
    if (verify() && !verify_early) verify_level(d, E, is_new, owner, G);
    if (xch && n_new) {
      // the one exchange step of a level: every rank's share of the DP results, in place, on the search's stream
      xcall([&] { xch->allgather3(L.k3.as<int32_t>(), L.cost.as<int32_t>(), L.total.as<int32_t>(), own_chunk(n_new, world), s); });
      c->launches += launch_max_k3(L.k3.as<int32_t>(), n_new, c->best.as<int>(), s);
    }
    if (defer) {
      L.nstart[0] = 0;
      L.nstart[1] = n_new;
      L.n_new = n_new;
      rst[0].scored[d - 1] = n_new;
    }
    for (int r = 0; r < R; ++r) rst[r].cells[d - 1] = cells[r];
    if (dbg) {
      const auto T4 = tnow();
      CK(spin_sync(s));
      const auto T5 = tnow();
      fprintf(stderr, "[nw level %d] R=%d P=%d G=%u new=%u lcap=%d: count %.2f ms, expand+dedup %.2f, assign %.2f, score prep %.2f, "
                      "score wait %.2f\n", d, R, L.P, G, n_new, lcap, tms(T0, T1), tms(T1, T2), tms(T2, T3), tms(T3, T4), tms(T4, T5));
    }
    gtotal += G;
    n_ids += (int32_t)n_new;
    return n_new > 0;
  }

  // frontier of level d+1 from the fresh candidates of level d (solver.jl:551-553, :669-672), per region
  void select_next(int d) {
    cudaStream_t s = c->stream;
    lv.emplace_back();
    Level& N = lv[d + 1];
    Level& Lr = lv[d];
    const uint32_t n = Lr.n_new;
    int32_t* misc = (int32_t*)c->misc.p;
    N.fstart.assign(R + 1, 0);
    static const bool dbg = getenv("NW_DEBUG_TIMING") != nullptr;
    const auto S0 = std::chrono::steady_clock::now();
    if (n == 0) {
      N.P = 0;
      return;
    }
    uint32_t* flag = (uint32_t*)c->flag.ensure((size_t)n * 4, true);
    uint32_t* pre = (uint32_t*)c->pre.ensure((size_t)n * 4, true);
    uint32_t* bsum = (uint32_t*)c->bsum.ensure(scan_scratch_items(n) * 4, true);
    c->launches += launch_valid_flags(Lr.k3.as<int32_t>(), n, flag, s);
    c->launches += launch_scan(flag, pre, n, bsum, (uint32_t*)(misc + 6), s);
    if (R > 1) gather_async(pre, Lr.nstart, n, 0xffffffffu);  // first valid rank of every region
    d2h(c, pin_u32(), misc, 32);
    if (R > 1) d2h(c, pin_u32() + 64, c->gather.p, (size_t)(R + 1) * 4);
    CK(spin_sync(s));
    const uint32_t nvalid = pin_u32()[6];
    std::vector<uint32_t> vstart(R + 1);
    for (int r = 0; r <= R; ++r) {
      const uint32_t v = R > 1 ? pin_u32()[64 + r] : (r == 0 ? 0u : nvalid);
      vstart[r] = v == 0xffffffffu ? nvalid : v;
    }
    // how many children of every region enter the next frontier
    std::vector<uint32_t> take(R), ostart(R + 1, 0);
    for (int r = 0; r < R; ++r) {
      const uint32_t v = vstart[r + 1] - vstart[r];
      take[r] = (o.maxwidth >= 0 && (int64_t)v > o.maxwidth) ? (uint32_t)o.maxwidth : v;
      ostart[r + 1] = ostart[r] + take[r];
    }
    const uint32_t P2 = ostart[R];
    N.fstart = ostart;
    N.P = (int)P2;
    uint32_t* sel = (uint32_t*)c->sel.ensure((size_t)std::max<uint32_t>(nvalid, 1) * 4, true);
    if (o.maxwidth < 0) {
      c->launches += launch_compact_ranks(flag, pre, n, sel, s);  // region-major already (candidates follow parents)
    } else if (R == 1 && o.maxwidth >= 1 && (int64_t)nvalid > o.maxwidth && nvalid > 2048 && !getenv("NW_NO_TOPK")) {
      // beam search of one region: the width-th best score from a histogram, only the winners are sorted
      const uint32_t wsel = (uint32_t)o.maxwidth;
      uint32_t npad = 2048;
      while (npad < wsel) npad <<= 1;
      uint64_t* keys = (uint64_t*)c->keys.ensure((size_t)npad * 8, true);
      uint32_t* hist = (uint32_t*)c->hist.ensure((size_t)100008 * 4, true);
      // flag is consumed (the valid flags are no longer needed), pre becomes the tie ranks
      bsum = (uint32_t*)c->bsum.ensure(std::max<uint64_t>(scan_scratch_items(n), 128) * 4, true);
      c->launches += launch_topk_keys(Lr.k3.as<int32_t>(), n, wsel, hist, hist + 100001, flag, pre, bsum, (uint32_t*)(misc + 7),
                                      keys, npad, s);
      if (rank_select_fits(wsel)) {  // the winners in frontier order by counting (no sort network, no padding)
        c->launches += launch_rank_select(keys, wsel, sel, s);
      } else {
        c->launches += launch_sort_u64(keys, npad, s);
        uint32_t* seg = (uint32_t*)c->gather.ensure((size_t)(3 * R + 3) * 4, true);
        const uint32_t pack[3] = {0u, wsel, 0u};
        h2d(c, seg, pack, sizeof pack);
        c->launches += launch_take_sorted(keys, wsel, seg, seg + 1, seg + 2, sel, s);
      }
    } else if (nvalid) {
      uint64_t npad = 2048;
      while (npad < nvalid) npad <<= 1;
      uint64_t* keys = (uint64_t*)c->keys.ensure(npad * 8, true);
      c->launches += launch_make_keys(Lr.k3.as<int32_t>(), R > 1 ? Lr.nrid.as<uint16_t>() : nullptr, flag, pre, n, keys, nvalid,
                                      npad, s);
      if (R == 1 && rank_select_fits(nvalid) && (int64_t)take[0] == (int64_t)nvalid) {
        // one region keeping every valid child (small levels): order by counting, straight into sel
        c->launches += launch_rank_select(keys, nvalid, sel, s);
        goto selected;
      }
      c->launches += launch_sort_u64(keys, npad, s);
      // sorted keys are region-major: region r occupies [vstart[r], vstart[r+1]); keep its first take[r]
      uint32_t* seg = (uint32_t*)c->gather.ensure((size_t)(3 * R + 3) * 4, true);
      std::vector<uint32_t> pack(3 * (size_t)R);
      for (int r = 0; r < R; ++r) {
        pack[r] = vstart[r];
        pack[R + r] = take[r];
        pack[2 * (size_t)R + r] = ostart[r];
      }
      h2d(c, seg, pack.data(), pack.size() * 4);
      c->launches += launch_take_sorted(keys, nvalid, seg, seg + R, seg + 2 * (size_t)R, sel, s);
      // no sync: the pageable source of the H2D above is consumed when cudaMemcpyAsync returns
    }
  selected:
    if (P2 == 0) return;
    int maxlen = std::max(level_maxlen, c->nx_max);
    maxlen = std::max(maxlen, Lr.LS);  // inversions keep the parent's length
    N.LS = round_ls(maxlen);
    N.pstride = blob_stride(N.LS);
    MaterialiseArgs M{};
    M.F = frontier_of(Lr);
    M.sel = sel;
    M.newlist = Lr.newlist.as<uint32_t>();
    M.desc = Lr.desc.as<uint64_t>();
    M.id_base = Lr.id_base;
    M.maxdup = o.maxdup;
    M.nx_max = c->nx_max;
    M.P2 = (int)P2;
    M.LS2 = N.LS;
    M.pstride2 = N.pstride;
    M.blob2 = (unsigned char*)N.fblob.ensure((size_t)P2 * N.pstride);
    M.len2 = (int32_t*)N.flen.ensure((size_t)P2 * 4);
    M.id2 = (int32_t*)N.fid.ensure((size_t)P2 * 4);
    M.dupok2 = (uint8_t*)N.fdup.ensure((size_t)P2);
    M.rid2 = (uint16_t*)N.frid.ensure((size_t)P2 * 2);
    M.pw = c->d_pow.as<U4>();
    c->launches += launch_materialise(M, s);
    CK(cudaGetLastError());
    if (dbg) {
      CK(spin_sync(s));
      fprintf(stderr, "[nw select %d] valid=%u P2=%u: %.2f ms\n", d, nvalid, P2,
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - S0).count());
    }
  }

  void pad_levels(int D) {  // levels that were never reached still exist (empty) for the report loops
    for (int e = 0; e <= D; ++e) {
      if ((int)lv.size() <= e) lv.emplace_back();
      Level& L = lv[e];
      if ((int)L.fstart.size() != R + 1) L.fstart.assign(R + 1, 0);
      if ((int)L.cstart.size() != R + 1) L.cstart.assign(R + 1, 0);
      if ((int)L.nstart.size() != R + 1) L.nstart.assign(R + 1, 0);
    }
  }
  // per-region statistics once the search is over (best, ids, timing shared by the batch)
  void finish_stats(int n_levels) {
    cudaStream_t s = c->stream;
    std::vector<int32_t> best(R, 0);
    d2h(c, best.data(), c->best.p, (size_t)R * 4);
    CK(spin_sync(s));
    for (size_t k = 0; k + 1 < lev.size(); k += 2) {
      float ms = 0;
      cudaEventElapsedTime(&ms, lev[k], lev[k + 1]);
      st.level_ms[k / 2] = ms;
    }
    if (ev0 && ev1) cudaEventElapsedTime(&st.total_ms, ev0, ev1);
    st.n_levels = n_levels;
    st.n_ids = n_ids;
    st.kernel_launches = c->launches;
    if (xch) {  // DP cells are counted where they are evaluated: sum the ranks' shares (R == 1 in a sharded search)
      int64_t v[16];
      for (int l = 0; l < 16; ++l) v[l] = rst[0].cells[l];
      xcall([&] { xch->allreduce_sum(v, 16, s); });
      for (int l = 0; l < 16; ++l) rst[0].cells[l] = v[l];
    }
    for (int r = 0; r < R; ++r) {
      best_k3[r] = best[r];
      nw_stats& q = rst[r];
      q.n_levels = n_levels;
      q.total_ms = st.total_ms;
      for (int l = 0; l < 16; ++l) q.level_ms[l] = st.level_ms[l];
      q.n_ids = 1;
      for (int l = 1; l < (int)lv.size(); ++l) q.n_ids += lv[l].nstart[r + 1] - lv[l].nstart[r];
      q.best = k3_to_score(best_k3[r]);
      q.kernel_launches = c->launches;
    }
    st.best = rst[0].best;
  }

  void run() {
    cudaStream_t s = c->stream;
    CK(cudaEventCreate(&ev0));
    CK(cudaEventCreate(&ev1));
    CK(cudaEventRecord(ev0, s));
    init_root();
    const int D = std::min(o.maxdepth, MAX_LEVELS - 1);
    int n_levels = 0;
    for (int d = 1; d <= D; ++d) {
      cudaEvent_t a, b;
      CK(cudaEventCreate(&a));
      CK(cudaEventCreate(&b));
      lev.push_back(a);
      lev.push_back(b);
      CK(cudaEventRecord(a, s));
      const bool any = run_level(d);
      n_levels = d;
      if (d < D) select_next(d);
      CK(cudaEventRecord(b, s));
      // the frontier blob of level d is no longer needed once level d+1 is materialised (ids are kept); the root's own
      // DP on the side stream reads level 1's blob: whatever reuses that memory is ordered behind it (no host wait)
      if (d == 1 && root_pending) CK(cudaStreamWaitEvent(s, c->side_done, 0));
      if (!verify()) lv[d].fblob.release();  // the verifier re-reads the sequences of old frontiers
      if (d < D && (!any || lv[d + 1].P == 0)) {
        n_levels = D;  // solver.jl keeps looping over empty frontiers; nothing more can be generated
        break;
      }
    }
    pad_levels(D);
    join_root();
    CK(cudaEventRecord(ev1, s));
    CK(spin_sync(s));
    CK(cudaGetLastError());
    finish_stats(n_levels);
  }

  // ---- report: get_good_trajectories (solver.jl:727-745) ----
  // host graph of the needed ids: ids sorted ascending (binary search), incoming edges in CSR form, every
  // child's edges in discovery order (solver.jl:559-560 appends provenance in the order candidates are generated)
  struct NodeTable {
    std::vector<int32_t> id, k3, cost, total;
    std::vector<uint32_t> estart;
    std::vector<EdgeRec> edges;
    std::vector<int32_t> eparent;  // node index of every edge's parent (-1: not in the table)
    std::vector<signed char> memo; // reachability memo of the enumerator: [node][moves] 0 unknown, 1 yes, 2 no
    int memo_w = 0;
    int find(int32_t gid) const {
      auto it = std::lower_bound(id.begin(), id.end(), gid);
      return (it != id.end() && *it == gid) ? (int)(it - id.begin()) : -1;
    }
    size_t size() const { return id.size(); }
  };
  NodeTable nodes;

  static float score_f32(int k3) { return k3 < 0 ? -INFINITY : (float)k3_to_score(k3); }

  // ---- report extraction, in phases so that the frontier-sharded search can interleave collectives ----
  // phase 1: which ids are reported (identical on every rank: it only reads the replicated per-id arrays)
  void extract_good(uint8_t* needed, int k3_min, int k3_star, int64_t quota) {
    cudaStream_t s = c->stream;
    CK(cudaMemsetAsync(needed, 0, (size_t)n_ids + 16, s));
    const int nl = (int)lv.size();
    int32_t* misc0 = (int32_t*)c->misc.p;
    for (int l = 1; l < nl; ++l) {
      const Level& L = lv[l];
      if (!L.n_new) continue;
      uint32_t* tie_rank = nullptr;
      uint32_t q = 0xffffffffu;
      if (k3_star >= 0) {  // rank of every id tied at k3_star, in id order (levels ascend with id)
        // (not is_new / owner: the last level's dedup verifier may still be reading them on the side stream)
        uint32_t* tflag = (uint32_t*)c->rflag.ensure((size_t)L.n_new * 4, true);
        tie_rank = (uint32_t*)c->newrank.ensure((size_t)L.n_new * 4, true);
        uint32_t* bsum = (uint32_t*)c->bsum.ensure(scan_scratch_items(L.n_new) * 4, true);
        c->launches += launch_tie_flags(L.k3.as<int32_t>(), L.n_new, k3_star, tflag, s);
        c->launches += launch_scan(tflag, tie_rank, L.n_new, bsum, (uint32_t*)(misc0 + 10), s);
        q = (uint32_t)std::min<int64_t>(std::max<int64_t>(quota, 0), 0xfffffff0ll);
      }
      c->launches += launch_good_flags(L.k3.as<int32_t>(), L.n_new, L.id_base, k3_min, k3_star, q, tie_rank, needed, s);
      if (k3_star >= 0) {
        d2h(c, pin_u32(), misc0, 64);
        CK(spin_sync(s));
        quota -= (int64_t)pin_u32()[10];
      }
    }
  }
  // phase 2: one step of the depth-bounded ancestor closure over this rank's edges (pass = 1 + child distance)
  void extract_mark(uint8_t* needed, int pass) {
    const int nl = (int)lv.size();
    for (int l = nl - 1; l >= 1; --l)
      if (l + pass <= o.maxdepth + 2)
        c->launches += launch_mark_parents(lv[l].cand_id.as<int32_t>(), lv[l].desc.as<uint64_t>(),
                                           lv[l].fid.as<int32_t>(), lv[l].G, needed, pass, l, o.maxdepth, c->stream);
  }
  // phase 3: node index of every needed id (exclusive scan of the needed flags: ascending id order, roots first),
  // then this rank's edges into needed ids (all levels) compacted into c->edgebuf with node indices as end points;
  // returns their number.  The ranks are identical on every rank of a sharded search (needed is replicated).
  uint32_t n_nodes = 0;
  uint32_t extract_edges(uint8_t* needed) {
    cudaStream_t s = c->stream;
    int32_t* misc = (int32_t*)c->misc.p;
    const int nl = (int)lv.size();
    const uint32_t nflag = (uint32_t)n_ids + 1;
    CK(cudaMemsetAsync(needed + 1, 1, (size_t)R, s));  // the roots are always nodes (index r for region r)
    uint32_t* nflags = (uint32_t*)c->rflag.ensure((size_t)nflag * 4, true);
    uint32_t* nrank = (uint32_t*)c->nrank.ensure((size_t)nflag * 4, true);
    uint32_t* bsum = (uint32_t*)c->bsum.ensure(scan_scratch_items(nflag) * 4, true);
    c->launches += launch_needed_flags(needed, nflag, nflags, s);
    c->launches += launch_scan(nflags, nrank, nflag, bsum, (uint32_t*)(misc + 11), s);
    uint32_t cap = std::max<uint32_t>(edge_cap_hint, 1u << 16);
    for (;;) {  // one pass; repeated with a larger buffer only if the append counter overflowed the capacity
      EdgeRec* buf = (EdgeRec*)c->edgebuf.ensure((size_t)cap * sizeof(EdgeRec));
      CK(cudaMemsetAsync(misc + 8, 0, 4, s));
      for (int l = 1; l < nl; ++l) {
        const Level& L = lv[l];
        c->launches += launch_edge_append(L.cand_id.as<int32_t>(), L.desc.as<uint64_t>(), L.fid.as<int32_t>(), L.G, needed,
                                          nrank, l, o.maxdepth, L.gbase, buf, cap, (uint32_t*)(misc + 8), s);
      }
      d2h(c, pin_u32(), misc, 64);
      CK(spin_sync(s));
      CK(cudaGetLastError());
      const uint32_t n = pin_u32()[8];
      n_nodes = pin_u32()[11];
      if (n <= cap) {
        edge_cap_hint = std::max(edge_cap_hint, n);
        return n;
      }
      cap = n + n / 8;
    }
  }
  // phase 4: host graph of the needed ids from the edge records of all ranks.  Node records arrive in node order
  // and edges carry node indices, so this is one counting sort by child plus the discovery order inside every
  // child's short edge list -- no searching.
  void build_nodes(const uint8_t* needed, std::vector<EdgeRec>& edges) {
    cudaStream_t s = c->stream;
    const int nl = (int)lv.size();
    const size_t nn = n_nodes;
    std::vector<IdRec> hi(nn);
    if (nn) {
      IdRec* buf = (IdRec*)c->keys.ensure(nn * sizeof(IdRec));
      for (int l = 1; l < nl; ++l) {
        const Level& L = lv[l];
        c->launches += launch_id_write(L.k3.as<int32_t>(), L.cost.as<int32_t>(), L.total.as<int32_t>(), L.id_base, L.n_new,
                                       needed, c->nrank.as<uint32_t>(), buf, s);
      }
      if (nn > (size_t)R) d2h(c, hi.data() + R, buf + R, (nn - (size_t)R) * sizeof(IdRec));
      CK(spin_sync(s));
    }
    for (int r = 0; r < R && (size_t)r < nn; ++r) hi[r] = IdRec{r + 1, root_k3[r], root_cost[r], root_total[r]};
    NodeTable& T = nodes;
    T.id.resize(nn);
    T.k3.resize(nn);
    T.cost.resize(nn);
    T.total.resize(nn);
    for (size_t k = 0; k < nn; ++k) {
      T.id[k] = hi[k].id;
      T.k3[k] = hi[k].k3;
      T.cost[k] = hi[k].cost;
      T.total[k] = hi[k].total;
    }
    T.estart.assign(nn + 1, 0);
    for (const EdgeRec& e : edges) T.estart[(size_t)e.child + 1]++;
    for (size_t k = 0; k < nn; ++k) T.estart[k + 1] += T.estart[k];
    T.edges.resize(edges.size());
    std::vector<uint32_t> cur(T.estart.begin(), T.estart.end() - 1);
    for (const EdgeRec& e : edges) T.edges[cur[e.child]++] = e;
    for (size_t k = 0; k < nn; ++k) {
      const uint32_t e0 = T.estart[k], e1 = T.estart[k + 1];
      if (e1 - e0 == 2) {
        if (T.edges[e0 + 1].gpos < T.edges[e0].gpos) std::swap(T.edges[e0], T.edges[e0 + 1]);
      } else if (e1 - e0 > 2) {
        std::sort(T.edges.begin() + e0, T.edges.begin() + e1,
                  [](const EdgeRec& a, const EdgeRec& b) { return a.gpos < b.gpos; });
      }
    }
    T.eparent.resize(T.edges.size());
    for (size_t e = 0; e < T.edges.size(); ++e) T.eparent[e] = T.edges[e].parent;
    T.memo_w = o.maxdepth + 3;
    T.memo.assign(nn * (size_t)T.memo_w, 0);
    CK(cudaGetLastError());
  }
  void extract(int k3_min, int k3_star, int64_t quota) {  // single-rank composition of the phases
    uint8_t* needed = (uint8_t*)c->flag.ensure((size_t)n_ids + 16, true);
    extract_good(needed, k3_min, k3_star, quota);
    for (int pass = 1; pass <= o.maxdepth + 1; ++pass) extract_mark(needed, pass);
    const uint32_t ne = extract_edges(needed);
    std::vector<EdgeRec> he(ne);
    if (ne) {
      d2h(c, he.data(), c->edgebuf.p, (size_t)ne * sizeof(EdgeRec));
      CK(spin_sync(c->stream));
    }
    build_nodes(needed, he);
  }

  // ---- trajectory enumeration: concatenate_moves_aux + sort + truncate (solver.jl:688-705, :739-743) ----
  // The reference enumerates every backward path from a reported id to the root (first visit of the root ends the
  // path; at most maxdepth+1 moves), in edge order, then stable-sorts by (-score, n_moves) and keeps max_report.
  // We produce exactly that sequence directly: groups by score desc, then n_moves asc, then id asc, then the
  // reference's own enumeration order restricted to paths of that length -- and stop at max_report.
  struct Enumerator {
    Search& S;
    std::vector<EdgeRec> stack;  // edges from the target backwards
    int64_t limit = -1, emitted = 0;
    nw_result* R = nullptr;
    int cur_id = 0, cur_k3 = 0, rid = 0;
    explicit Enumerator(Search& s) : S(s) {}
    // can a root be reached from node `nd` in exactly r moves (first visit of the root ends the path)
    bool reach(int nd, int r) {
      if (nd < 0) return false;
      NodeTable& T = S.nodes;
      if (T.id[nd] <= S.R) return r == 0;  // ids 1..R are the roots of the R regions
      if (r <= 0) return false;
      signed char& m = T.memo[(size_t)nd * T.memo_w + std::min(r, T.memo_w - 1)];
      if (m) return m == 1;
      bool ok = false;
      for (uint32_t e = T.estart[nd]; e < T.estart[nd + 1] && !ok; ++e) ok = reach(T.eparent[e], r - 1);
      T.memo[(size_t)nd * T.memo_w + std::min(r, T.memo_w - 1)] = ok ? 1 : 2;
      return ok;
    }
    bool full() const { return limit >= 0 && emitted >= limit; }
    // append a non-negative integer without allocating
    static void put_int(std::string& s, long long v) {
      char buf[24];
      int n = 0;
      do {
        buf[n++] = (char)('0' + v % 10);
        v /= 10;
      } while (v);
      while (n) s.push_back(buf[--n]);
    }
    int emit_id = -1, emit_cost = -1, emit_total = 0;
    int64_t emit_local = 0;
    std::string emit_score;
    void emit() {
      static const char* names[3] = {"move_dup", "move_del", "move_inv"};
      if (emit_id != cur_id) {  // per-id fields are looked up once, not once per trajectory
        emit_id = cur_id;
        const int nd = S.nodes.find(cur_id);
        emit_cost = nd >= 0 ? S.nodes.cost[nd] : -1;
        emit_total = nd >= 0 ? S.nodes.total[nd] : 0;
        int lvl, rr;
        S.locate(cur_id, lvl, rr, emit_local);
        emit_score = score_string(cur_k3);
      }
      R->t_score.push_back(score_f32(cur_k3));
      R->t_k3.push_back(cur_k3);
      R->t_id.push_back(emit_local);
      R->t_cost.push_back(emit_cost < 0 ? emit_cost : (int64_t)emit_cost * S.c->RH[rid].gcd);
      R->t_total.push_back((int64_t)emit_total * S.c->RH[rid].gcd);
      std::string& line = R->tsv;
      line += emit_score;
      line += '\t';
      put_int(line, (long long)stack.size());
      line += '\t';  // :723
      bool first = true;
      for (size_t k = stack.size(); k-- > 0;) {  // stack holds the last move first; trajectories list moves root-first
        const EdgeRec& e = stack[k];
        const int kind = edge_kind(e), i = edge_i(e), j = edge_j(e);
        R->t_moves.push_back(kind);
        R->t_moves.push_back(i);
        R->t_moves.push_back(j);
        if (!first) line += '\t';
        first = false;
        line += names[kind];
        line += '\t';
        put_int(line, i);
        line += '\t';
        put_int(line, j);  // :719
      }
      line += '\n';
      R->t_off.push_back((int64_t)R->t_moves.size());
      ++emitted;
    }
    // all paths of exactly r more moves from node `nd` back to the root, in the reference's enumeration order
    void walk(int nd, int r) {
      if (nd < 0) return;
      const NodeTable& T = S.nodes;
      if (T.id[nd] <= S.R) {
        if (r == 0) emit();
        return;
      }
      for (uint32_t q = T.estart[nd]; q < T.estart[nd + 1]; ++q) {
        if (full()) return;
        if (!reach(T.eparent[q], r - 1)) continue;
        stack.push_back(T.edges[q]);
        walk(T.eparent[q], r - 1);
        stack.pop_back();
      }
    }
  };

  // thresholds of the report (identical on every rank): k3_min from minreport, k3_star/quota from maxreport
  // smallest k3 whose Float32 score passes `score >= suboptimality * best` (solver.jl:733; monotone in k3)
  int k3_min_of(int best) const {
    const double thr = o.minreport * k3_to_score(best);
    int lo = 0, hi = 100001;  // first k3 in [0, 100000] with (double)(float)(k3/1000) >= thr
    while (lo < hi) {
      const int mid = (lo + hi) / 2;
      if ((double)score_f32(mid) >= thr) hi = mid; else lo = mid + 1;
    }
    return lo;
  }
  // thresholds of a single-region report (identical on every rank): k3_min from minreport, k3_star/quota from maxreport
  void report_thresholds(int& k3_min, int& k3_star, int64_t& quota) {
    k3_min = k3_min_of(best_k3[0]);
    k3_star = -1;
    quota = 0;
    if (o.maxreport >= 0 && n_ids > o.maxreport) {
      cudaStream_t s = c->stream;
      uint32_t* hist = (uint32_t*)c->keys.ensure((size_t)100002 * 4, true);
      CK(cudaMemsetAsync(hist, 0, (size_t)100002 * 4, s));
      for (size_t l = 1; l < lv.size(); ++l)
        c->launches += launch_k3_hist(lv[l].k3.as<int32_t>(), lv[l].n_new, k3_min, hist, s);
      std::vector<uint32_t> hh(100002, 0);
      if (k3_min <= 100000) d2h(c, hh.data() + k3_min, hist + k3_min, (size_t)(100001 - k3_min) * 4);  // bins below are empty
      CK(spin_sync(s));
      if (root_k3[0] >= 0 && root_k3[0] >= k3_min) hh[root_k3[0]] += 1;
      int64_t cum = 0;
      for (int k = 100000; k >= 0; --k) {
        if (cum + hh[k] >= o.maxreport) {
          k3_star = k;
          quota = o.maxreport - cum;  // ids to take from the tie bin, in id order
          break;
        }
        cum += hh[k];
      }
      if (k3_star >= 0 && root_k3[0] == k3_star) quota -= 1;  // the root (id 1) is the first of its tie bin
    }
  }

  // ---- global id <-> (level, region, per-region id).  Ids are assigned level by level in global discovery order and
  // a level's candidates are region-major, so the fresh ids of one region at one level are contiguous. ----
  std::vector<std::vector<int64_t>> local_base;  // [level][region]: per-region id of the region's first fresh id
  void build_local_ids() {
    const int nl = (int)lv.size();
    local_base.assign(nl, std::vector<int64_t>(R, 2));
    for (int l = 2; l < nl; ++l)
      for (int r = 0; r < R; ++r) local_base[l][r] = local_base[l - 1][r] + (lv[l - 1].nstart[r + 1] - lv[l - 1].nstart[r]);
  }
  void locate(int32_t gid, int& level, int& rid, int64_t& local) const {
    if (gid <= R) {
      level = 0;
      rid = gid - 1;
      local = 1;
      return;
    }
    int l = 1;
    while (l + 1 < (int)lv.size() && lv[l + 1].n_new > 0 && gid >= lv[l + 1].id_base) ++l;
    const uint32_t rank = (uint32_t)(gid - lv[l].id_base);
    const auto& ns = lv[l].nstart;
    rid = (int)(std::upper_bound(ns.begin(), ns.end(), rank) - ns.begin()) - 1;
    if (rid < 0) rid = 0;
    if (rid >= R) rid = R - 1;
    level = l;
    local = local_base[l][rid] + (rank - ns[rid]);
  }

  // one search, one region (also every rank of a sharded search)
  void report(nw_result* out) {
    int k3_min, k3_star;
    int64_t quota;
    report_thresholds(k3_min, k3_star, quota);
    extract(k3_min, k3_star, quota);
    build_local_ids();
    enumerate(out, 0, k3_min);
  }

  // regionfile mode: several regions searched together; out[r] receives region r's report
  void report_many(nw_result** out) {
    if (R == 1) {
      report(out[0]);
      return;
    }
    cudaStream_t s = c->stream;
    const bool dbg = getenv("NW_DEBUG_TIMING") != nullptr;
    auto tnow = [] { return std::chrono::steady_clock::now(); };
    auto tms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return std::chrono::duration<double, std::milli>(b - a).count();
    };
    const auto T0 = tnow();
    int32_t* misc = (int32_t*)c->misc.p;
    build_local_ids();
    std::vector<int32_t> k3min(R);
    for (int r = 0; r < R; ++r) k3min[r] = k3_min_of(best_k3[r]);
    int32_t* d_k3min = (int32_t*)c->gather.ensure((size_t)R * 4, true);
    h2d(c, d_k3min, k3min.data(), (size_t)R * 4);
    // ids that pass their region's threshold (usually few): (gid, k3, rid) records, selected per region on the host
    struct Qual {
      int32_t gid, k3, rid, pad;
    };
    std::vector<Qual> q;
    uint32_t cap = 1u << 16;
    for (;;) {
      Qual* buf = (Qual*)c->keys.ensure((size_t)cap * sizeof(Qual));
      CK(cudaMemsetAsync(misc + 9, 0, 4, s));
      for (size_t l = 1; l < lv.size(); ++l)
        c->launches += launch_qual_append(lv[l].k3.as<int32_t>(), lv[l].nrid.as<uint16_t>(), d_k3min, lv[l].id_base,
                                          lv[l].n_new, buf, cap, (uint32_t*)(misc + 9), s);
      d2h(c, pin_u32(), misc, 64);
      CK(spin_sync(s));
      const uint32_t n = pin_u32()[9];
      if (n <= cap) {
        q.resize(n);
        if (n) {
          d2h(c, q.data(), buf, (size_t)n * sizeof(Qual));
          CK(spin_sync(s));
        }
        break;
      }
      cap = n + n / 8;
    }
    const auto T1 = tnow();
    std::vector<std::vector<std::pair<int32_t, int32_t>>> good(R);  // (-k3, gid)
    for (int r = 0; r < R; ++r)
      if (root_k3[r] >= 0 && root_k3[r] >= k3min[r]) good[r].emplace_back(-root_k3[r], r + 1);
    for (const Qual& e : q) good[e.rid].emplace_back(-e.k3, e.gid);
    std::vector<int32_t> gids;
    for (int r = 0; r < R; ++r) {
      std::sort(good[r].begin(), good[r].end());
      // only the first maxreport ids in (score desc, id asc) order can own one of the first maxreport trajectories
      if (o.maxreport >= 0 && (int64_t)good[r].size() > o.maxreport) good[r].resize((size_t)o.maxreport);
      for (auto& g : good[r]) gids.push_back(g.second);
    }
    uint8_t* needed = (uint8_t*)c->flag.ensure((size_t)n_ids + 16, true);
    CK(cudaMemsetAsync(needed, 0, (size_t)n_ids + 16, s));
    if (!gids.empty()) {
      int32_t* dg = (int32_t*)c->keys.ensure(gids.size() * 4, true);
      h2d(c, dg, gids.data(), gids.size() * 4);
      c->launches += launch_set_needed(dg, (uint32_t)gids.size(), needed, s);
      CK(spin_sync(s));
    }
    const auto T2 = tnow();
    for (int pass = 1; pass <= o.maxdepth + 1; ++pass) extract_mark(needed, pass);
    const uint32_t ne = extract_edges(needed);
    std::vector<EdgeRec> he(ne);
    if (ne) {
      d2h(c, he.data(), c->edgebuf.p, (size_t)ne * sizeof(EdgeRec));
      CK(spin_sync(s));
    }
    const auto T3 = tnow();
    build_nodes(needed, he);
    const auto T4 = tnow();
    {  // regions are independent (disjoint nodes, own result object): enumerate them on a few host threads
      const int nt = std::max(1, std::min({8, host_cores() / std::max(1, g_live_ctxs.load()), R / 32}));
      std::atomic<int> next{0};
      std::exception_ptr err;
      std::mutex err_mu;
      auto work = [&] {
        try {
          for (int r = next.fetch_add(1); r < R; r = next.fetch_add(1)) enumerate_list(out[r], r, good[r]);
        } catch (...) {
          std::lock_guard<std::mutex> g(err_mu);
          if (!err) err = std::current_exception();
          next.store(R);
        }
      };
      std::vector<std::thread> th;
      for (int t = 1; t < nt; ++t) th.emplace_back(work);
      work();
      for (auto& t : th) t.join();
      if (err) std::rethrow_exception(err);
    }
    const auto T5 = tnow();
    if (dbg)
      fprintf(stderr, "[nw report_many] R=%d qual %.2f ms (%zu ids), select %.2f, mark+edges %.2f (%u edges), nodes %.2f (%zu), "
                      "enumerate %.2f\n", R, tms(T0, T1), q.size(), tms(T1, T2), tms(T2, T3), ne, tms(T3, T4), nodes.size(),
              tms(T4, T5));
  }

  void enumerate(nw_result* out, int rid, int k3_min) {
    std::vector<std::pair<int32_t, int32_t>> good;  // (-k3, id): canonical Dict order = ascending discovery id
    for (size_t k = 0; k < nodes.size(); ++k) {
      const int k3 = nodes.k3[k];
      if (k3 >= 0 && k3 >= k3_min) good.emplace_back(-k3, nodes.id[k]);
    }
    std::sort(good.begin(), good.end());
    enumerate_list(out, rid, good);
  }
  // trajectories of the ids in `good` (sorted by score desc, id asc), in final report order, up to maxreport
  void enumerate_list(nw_result* out, int rid, const std::vector<std::pair<int32_t, int32_t>>& good) {
    Enumerator E(*this);
    E.limit = o.maxreport;
    E.R = out;
    E.rid = rid;
    if (o.maxreport >= 0) {  // the usual report is full: size the output once
      const size_t n = (size_t)std::min<int64_t>(o.maxreport, std::min<int64_t>(4096, 16 + 2 * (int64_t)good.size()));
      out->t_score.reserve(n);
      out->t_k3.reserve(n);
      out->t_id.reserve(n);
      out->t_cost.reserve(n);
      out->t_total.reserve(n);
      out->t_off.reserve(n + 1);
      out->t_moves.reserve(n * 3 * (size_t)(o.maxdepth + 1));
      out->tsv.reserve(n * (size_t)(16 + 20 * (o.maxdepth + 1)));
    }
    out->t_off.push_back(0);
    size_t g0 = 0;
    while (g0 < good.size() && !E.full()) {
      size_t g1 = g0;
      while (g1 < good.size() && good[g1].first == good[g0].first) ++g1;
      for (int m = 0; m <= o.maxdepth + 1 && !E.full(); ++m)
        for (size_t g = g0; g < g1 && !E.full(); ++g) {
          E.cur_id = good[g].second;
          E.cur_k3 = -good[g].first;
          const int nd = nodes.find(E.cur_id);
          if (E.reach(nd, m)) E.walk(nd, m);
        }
      g0 = g1;
    }
  }

  // keep_ids diagnostics: every id and every provenance edge, per region, with per-region (reference) ids
  void dump_ids(nw_result** out) {
    cudaStream_t s = c->stream;
    if (local_base.empty()) build_local_ids();
    for (int r = 0; r < R; ++r) {
      nw_result* Q = out[r];
      const size_t n = (size_t)rst[r].n_ids;
      Q->id_score.assign(n, 0.f);
      Q->id_cost.assign(n, -1);
      Q->id_total.assign(n, 0);
      Q->id_level.assign(n, 0);
    }
    auto put = [&](int r, int64_t id, int k3, int cost, int total, int level) {
      nw_result* Q = out[r];
      Q->id_score[id - 1] = score_f32(k3);
      Q->id_cost[id - 1] = cost < 0 ? cost : (int64_t)cost * c->RH[r].gcd;
      Q->id_total[id - 1] = (int64_t)total * c->RH[r].gcd;
      Q->id_level[id - 1] = level;
    };
    for (int r = 0; r < R; ++r) put(r, 1, root_k3[r], root_cost[r], root_total[r], 0);
    for (int l = 1; l < (int)lv.size(); ++l) {
      const Level& L = lv[l];
      if (L.n_new) {
        std::vector<int32_t> k3(L.n_new), cost(L.n_new), total(L.n_new);
        d2h(c, k3.data(), L.k3.p, (size_t)L.n_new * 4);
        d2h(c, cost.data(), L.cost.p, (size_t)L.n_new * 4);
        d2h(c, total.data(), L.total.p, (size_t)L.n_new * 4);
        CK(spin_sync(s));
        for (int r = 0; r < R; ++r)
          for (uint32_t q = L.nstart[r]; q < L.nstart[r + 1]; ++q)
            put(r, local_base[l][r] + (q - L.nstart[r]), k3[q], cost[q], total[q], l);
      }
      if (L.G) {
        std::vector<int32_t> cid(L.G), fid(L.P);
        std::vector<uint64_t> desc(L.G);
        d2h(c, cid.data(), L.cand_id.p, (size_t)L.G * 4);
        d2h(c, desc.data(), L.desc.p, (size_t)L.G * 8);
        d2h(c, fid.data(), L.fid.p, (size_t)L.P * 4);
        CK(spin_sync(s));
        for (int r = 0; r < R; ++r) {
          nw_result* Q = out[r];
          for (uint32_t g = L.cstart[r]; g < L.cstart[r + 1]; ++g) {
            int lvl, rr;
            int64_t lc, lp;
            locate(cid[g], lvl, rr, lc);
            locate(fid[desc_parent(desc[g])], lvl, rr, lp);
            Q->e_child.push_back((int32_t)lc);
            Q->e_parent.push_back((int32_t)lp);
            Q->e_moves.push_back(desc_kind(desc[g]));
            Q->e_moves.push_back(desc_i(desc[g]));
            Q->e_moves.push_back(desc_j(desc[g]));
          }
        }
      }
    }
  }
};

int guard(const std::function<void()>& fn) {
  try {
    fn();
    return NW_OK;
  } catch (const Err& e) {
    g_last_error = e.msg;
    return e.code;
  } catch (const HostError& e) {
    g_last_error = e.what();
    return e.code;
  } catch (const std::bad_alloc&) {
    g_last_error = "host out of memory";
    return NW_ERR_NOMEM;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return NW_ERR_ARG;
  }
}

struct PoolScope {
  DevPool* prev;
  explicit PoolScope(nw_ctx* c) : prev(g_pool) { g_pool = c ? &c->pool : nullptr; }
  ~PoolScope() { g_pool = prev; }
};

void check_opts(const nw_opts* o) {
  if (!o) fail(NW_ERR_ARG, "opts is null");
  if (o->maxdepth < 1 || o->maxdepth >= MAX_LEVELS) fail(NW_ERR_ARG, "maxdepth must be in [1, 15]");
  if (o->maxdup < 0) fail(NW_ERR_ARG, "maxdup must be >= 0");
}

// Searches R regions together through one pipeline instance (R == 1: the plain single search).  out[r] receives
// region r's report and statistics; timings / launch counts / byte counts are those of the whole batch.
void do_solve_many(nw_ctx* c, const ProblemRef* probs, int R, const nw_opts* o, nw_result** out, Exchange* xch = nullptr) {
  CK(cudaSetDevice(c->device));
  if (xch && R != 1) fail(NW_ERR_ARG, "a sharded search holds one region");
  c->launches = 0;
  c->h2d_bytes = c->d2h_bytes = 0;
  collect_score_time(c, nullptr, nullptr);
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<float, std::milli>(b - a).count();
  };
  const auto t0 = now();
  setup_regions(c, probs, R, o->maxdepth, NW_BAND_JULIA);
  std::unique_ptr<Search> Sp;
  const auto t1 = now();
  auto t2 = t1;
  float score_ms = 0;
  int score_launches = 0;
  const uint64_t reseed0 = c->fp_reseed;
  for (int attempt = 0;; ++attempt) {
    Sp.reset(new Search(c, *o));
    Sp->fp_reseed0 = reseed0;
    if (xch && xch->world() > 1) {
      Sp->xch = xch;
      Sp->rank = xch->rank();
      Sp->world = xch->world();
    }
    try {
      Sp->run();
      t2 = now();
      collect_score_time(c, &score_ms, &score_launches);
      if (!(o->flags & NW_FLAG_NO_TRAJ)) Sp->report_many(out);
      Sp->check_dedup();
      break;
    } catch (const Err& e) {
      // two different indices under one table key (found by the element-wise verifier): other hash bases, same search.
      // Every rank of a sharded search sees the same collision (the dedup is replicated), so they all come here.
      if (e.code != NW_ERR_COLLISION || attempt >= 2 || (o->flags & NW_FLAG_DEBUG_WEAK_FP)) throw;
      Sp.reset();
      for (int r = 0; r < R; ++r) *out[r] = nw_result();
      c->fp_reseed += 1;
      c->h_pow = build_pow_table(c->fp_reseed);
      h2d(c, c->d_pow.p, c->h_pow.data(), c->h_pow.size() * sizeof(U4));
      CK(spin_sync(c->stream));
      if (c->side) CK(cudaStreamSynchronize(c->side));
      collect_score_time(c, nullptr, nullptr);
    }
  }
  Search& S = *Sp;
  const auto t3 = now();
  if (o->keep_ids) S.dump_ids(out);
  const auto t4 = now();
  collect_score_time(c, nullptr, nullptr);
  for (int r = 0; r < R; ++r) {
    nw_stats q = S.rst[r];
    q.score_ms = score_ms;
    q.score_launches = score_launches;
    q.host_ms[0] = ms(t0, t1);
    q.host_ms[1] = ms(t1, t2);
    q.host_ms[2] = ms(t2, t3);
    q.host_ms[3] = ms(t3, t4);
    q.kernel_launches = c->launches;
    q.h2d_bytes = c->h2d_bytes;
    q.d2h_bytes = c->d2h_bytes;
    out[r]->stats = q;
  }
}

void do_solve(nw_ctx* c, const int8_t* aln, int n_x, int n_y, const int64_t* len_x, const int64_t* len_y,
              const nw_opts* o, nw_result* R) {
  const ProblemRef P{aln, n_x, n_y, len_x, len_y};
  do_solve_many(c, &P, 1, o, &R);
}

}  // namespace

// internal hooks for the other translation units of the library (nw_region.cu)
namespace nw {
struct ScoreProblem {
  const int8_t* aln;  // n_x*n_y, x-major
  int n_x, n_y;
  const int64_t* len_x;
  const int64_t* len_y;
};
void set_last_error(const std::string& m) { g_last_error = m; }
int ctx_device(nw_ctx* c) { return c->device; }
cudaStream_t ctx_stream(nw_ctx* c) { return c->stream; }
void ctx_add_launches(nw_ctx* c, int n) { c->launches += n; }
int ctx_launches(nw_ctx* c) { return c->launches; }
// growable staging pair of the region pre-pass; returns cudaSuccess or the allocation error
cudaError_t ctx_region_staging(nw_ctx* c, size_t bytes, void** host_pinned, void** dev) {
  if (bytes > c->rg_pinned_bytes) {
    if (c->rg_pinned) cudaFreeHost(c->rg_pinned);
    c->rg_pinned = nullptr;
    c->rg_pinned_bytes = 0;
    const size_t want = bytes + bytes / 2;
    cudaError_t e = cudaMallocHost(&c->rg_pinned, want);
    if (e != cudaSuccess) return e;
    c->rg_pinned_bytes = want;
  }
  if (bytes > c->rg_dev_bytes) {
    if (c->rg_dev) cudaFree(c->rg_dev);
    c->rg_dev = nullptr;
    c->rg_dev_bytes = 0;
    const size_t want = bytes + bytes / 2;
    cudaError_t e = cudaMalloc(&c->rg_dev, want);
    if (e != cudaSuccess) return e;
    c->rg_dev_bytes = want;
  }
  *host_pinned = c->rg_pinned;
  *dev = c->rg_dev;
  return cudaSuccess;
}
}  // namespace nw

extern "C" {

void nw_default_opts(nw_opts* o) {
  if (!o) return;
  o->maxdepth = 3;
  o->maxdup = 3;
  o->maxwidth = -1;
  o->minreport = 0.95;
  o->maxreport = -1;
  o->keep_ids = 0;
  o->flags = 0;
}
const char* nw_last_error(void) { return g_last_error.c_str(); }

const char* nw_version(void) { return "nwsolve-b200 0.1 (sm_100a)"; }

int nw_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int nw_ctx_create(int device, nw_ctx** out) {
  if (!out) return NW_ERR_ARG;
  *out = nullptr;
  return guard([&] {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      fail(NW_ERR_CUDA, std::string("no CUDA device: ") + (e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)));
    if (device < 0) {
      const char* lr = getenv("LOCAL_RANK");
      device = lr ? atoi(lr) % ndev : 0;
    }
    if (device >= ndev) fail(NW_ERR_ARG, "device index out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) fail(NW_ERR_CUDA, std::string("device is not sm_100 (Blackwell B200): ") + prop.name);
    {
      static std::mutex mu;
      static bool done[64] = {false};
      std::lock_guard<std::mutex> lk(mu);
      if (device < 64 && !done[device]) {
        init_kernel_attrs();
        CK(cudaGetLastError());
        done[device] = true;
      }
    }
    nw_ctx* c = new nw_ctx();
    g_live_ctxs.fetch_add(1);
    c->device = device;
    PoolScope pool_scope(c);
    try {
      {
        // the side stream ranks below the main one: its kernels (one root DP, the dedup verifier) fill what the critical
        // path leaves free -- the verifier of the last level runs under the report's short kernels and host phases
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, hi));
        CK(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, lo));
      }
      CK(cudaEventCreateWithFlags(&c->side_go, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&c->side_done, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&c->ver_go, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&c->ver_done, cudaEventDisableTiming));
      CK(cudaMallocHost(&c->pinned, PINNED_BYTES));
      c->pinned_bytes = PINNED_BYTES;
      c->h_pow = build_pow_table();
      CK(cudaMemcpy(c->d_pow.ensure(c->h_pow.size() * sizeof(U4)), c->h_pow.data(), c->h_pow.size() * sizeof(U4),
                    cudaMemcpyHostToDevice));
      c->misc.ensure(256);
    } catch (...) {
      nw_ctx_destroy(c);
      throw;
    }
    *out = c;
  });
}

void nw_ctx_destroy(nw_ctx* c) {
  if (!c) return;
  g_live_ctxs.fetch_sub(1);
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->side) cudaStreamSynchronize(c->side);
  if (c->side_go) cudaEventDestroy(c->side_go);
  if (c->side_done) cudaEventDestroy(c->side_done);
  if (c->ver_go) cudaEventDestroy(c->ver_go);
  if (c->ver_done) cudaEventDestroy(c->ver_done);
  if (c->side) cudaStreamDestroy(c->side);
  delete c->xch;
  c->xch = nullptr;
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->rg_pinned) cudaFreeHost(c->rg_pinned);
  if (c->rg_dev) cudaFree(c->rg_dev);
  for (auto& ev : c->score_events) c->event_pool.push_back(ev);
  for (auto& ev : c->event_pool) {
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  cudaStream_t s = c->stream;
  {
    // member buffers return to the pool while it is current; the pool itself is then freed explicitly
    c->pool.trim();
    DevPool keep;
    DevPool* prev = g_pool;
    g_pool = &keep;
    delete c;
    g_pool = prev;
    keep.trim();
  }
  if (s) cudaStreamDestroy(s);
}

int nw_solve(nw_ctx* c, const int8_t* aln, int32_t n_x, int32_t n_y, const int64_t* len_x, const int64_t* len_y,
             const nw_opts* o, nw_result** out) {
  if (!c || !aln || !len_x || !len_y || !out) {
    g_last_error = "null argument";
    return NW_ERR_ARG;
  }
  *out = nullptr;
  nw_result* R = new nw_result();
  PoolScope pool_scope(c);
  const int rc = guard([&] {
    check_opts(o);
    do_solve(c, aln, n_x, n_y, len_x, len_y, o, R);
  });
  if (rc != NW_OK) {
    delete R;
    cudaGetLastError();
    return rc;
  }
  *out = R;
  return NW_OK;
}

// the output appears complete or not at all: R/juliasolver.R:56-59 only checks that the file exists
static void write_file_atomic(const char* outp, const std::string& data) {
  const std::string tmp = std::string(outp) + ".tmp";
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) fail(NW_ERR_IO, std::string("cannot write ") + tmp);
  const size_t w = fwrite(data.data(), 1, data.size(), f);
  const int cl = fclose(f);
  if (w != data.size() || cl != 0) {
    remove(tmp.c_str());
    fail(NW_ERR_IO, std::string("short write to ") + tmp);
  }
  if (rename(tmp.c_str(), outp) != 0) {
    remove(tmp.c_str());
    fail(NW_ERR_IO, std::string("cannot rename to ") + outp);
  }
}

int nw_solve_files(nw_ctx* c, const char* inmat, const char* inlens, const char* outp, const nw_opts* o,
                   nw_stats* stats) {
  if (!c || !inmat || !inlens || !outp) {
    g_last_error = "null argument";
    return NW_ERR_ARG;
  }
  nw_result R;
  PoolScope pool_scope(c);
  const int rc = guard([&] {
    check_opts(o);
    ProblemHost P = read_problem(inmat, inlens);
    do_solve(c, P.aln.data(), P.n_x, P.n_y, P.len_x.data(), P.len_y.data(), o, &R);
    write_file_atomic(outp, R.tsv);
  });
  if (rc != NW_OK) cudaGetLastError();
  if (rc == NW_OK && stats) *stats = R.stats;
  return rc;
}

int nw_solve_files_many(nw_ctx** ctxs, int32_t n_ctx, const char* const* inmat, const char* const* inlens,
                        const char* const* outp, int64_t n, const nw_opts* opts, int32_t max_regions_per_pass,
                        int32_t* rcs) {
  if (!ctxs || n_ctx < 1 || !ctxs[0] || !inmat || !inlens || !outp || n < 0 || !opts) {
    g_last_error = "null argument";
    return NW_ERR_ARG;
  }
  std::vector<int32_t> rc_local((size_t)n, NW_OK);
  int32_t* rc_of = rcs ? rcs : rc_local.data();
  int first_rc = NW_OK;
  std::string first_msg;
  auto note = [&](int64_t k, int rc) {
    rc_of[k] = rc;
    if (rc != NW_OK && first_rc == NW_OK) {
      first_rc = rc;
      first_msg = "region " + std::to_string(k) + ": " + g_last_error;
    }
  };
  // 1. parse every region; a region that does not parse is reported and left out
  std::vector<ProblemHost> P((size_t)n);
  std::vector<int64_t> live;
  for (int64_t k = 0; k < n; ++k) {
    const int rc = guard([&] {
      if (!inmat[k] || !inlens[k] || !outp[k]) fail(NW_ERR_ARG, "null path");
      P[k] = read_problem(inmat[k], inlens[k]);
    });
    note(k, rc);
    if (rc == NW_OK) live.push_back(k);
  }
  // 2. forest search of the parsed regions; if the batch as a whole is refused (one region out of range, ...)
  //    the regions are solved one by one so that only the offending ones fail
  const int64_t m = (int64_t)live.size();
  std::vector<nw_problem> probs((size_t)m);
  for (int64_t q = 0; q < m; ++q) {
    const ProblemHost& H = P[live[q]];
    probs[q] = nw_problem{H.aln.data(), H.n_x, H.n_y, H.len_x.data(), H.len_y.data()};
  }
  std::vector<nw_result*> res((size_t)m, nullptr);
  int rc_all = m ? nw_solve_many_ctxs(ctxs, n_ctx, probs.data(), m, opts, max_regions_per_pass, res.data()) : NW_OK;
  if (rc_all != NW_OK)
    for (int64_t q = 0; q < m; ++q) {
      const nw_problem& B = probs[q];
      note(live[q], nw_solve(ctxs[0], B.aln_sign, B.n_x, B.n_y, B.len_x, B.len_y, opts, &res[q]));
    }
  // 3. reports
  for (int64_t q = 0; q < m; ++q) {
    if (!res[q]) continue;
    note(live[q], guard([&] { write_file_atomic(outp[live[q]], res[q]->tsv); }));
    delete res[q];
  }
  if (first_rc != NW_OK) g_last_error = first_msg;
  return first_rc;
}

int64_t nw_result_count(const nw_result* r) { return r ? (int64_t)r->t_score.size() : 0; }
int nw_result_get(const nw_result* r, int64_t k, float* score, int32_t* n_moves, const int32_t** moves3) {
  if (!r || k < 0 || k >= (int64_t)r->t_score.size()) return NW_ERR_ARG;
  if (score) *score = r->t_score[k];
  if (n_moves) *n_moves = (int32_t)((r->t_off[k + 1] - r->t_off[k]) / 3);
  if (moves3) *moves3 = r->t_moves.data() + r->t_off[k];
  return NW_OK;
}
int nw_result_cost(const nw_result* r, int64_t k, int64_t* cost_bp, int64_t* total_bp, int64_t* index_id) {
  if (!r || k < 0 || k >= (int64_t)r->t_score.size()) return NW_ERR_ARG;
  if (cost_bp) *cost_bp = r->t_cost[k];
  if (total_bp) *total_bp = r->t_total[k];
  if (index_id) *index_id = r->t_id[k];
  return NW_OK;
}
int nw_result_arrays(const nw_result* r, float* score, int32_t* n_moves, int64_t* move_off, int64_t* cost_bp,
                     int64_t* total_bp, int64_t* index_id, int32_t* moves3) {
  if (!r) return NW_ERR_ARG;
  const size_t n = r->t_score.size();
  for (size_t k = 0; k < n; ++k) {
    if (score) score[k] = r->t_score[k];
    if (n_moves) n_moves[k] = (int32_t)((r->t_off[k + 1] - r->t_off[k]) / 3);
    if (move_off) move_off[k] = r->t_off[k] / 3;
    if (cost_bp) cost_bp[k] = r->t_cost[k];
    if (total_bp) total_bp[k] = r->t_total[k];
    if (index_id) index_id[k] = r->t_id[k];
  }
  if (moves3 && !r->t_moves.empty()) memcpy(moves3, r->t_moves.data(), r->t_moves.size() * 4);
  return NW_OK;
}
// bulk accessors: the results of a whole regionfile run in two calls (an FFI host pays per call, not per region)
int nw_results_sizes(nw_result* const* rs, int64_t n, int64_t* n_traj, int64_t* n_moves, int64_t* tsv_len) {
  if (!rs || n < 0) return NW_ERR_ARG;
  for (int64_t k = 0; k < n; ++k) {
    const nw_result* r = rs[k];
    if (!r) return NW_ERR_ARG;
    if (n_traj) n_traj[k] = (int64_t)r->t_score.size();
    if (n_moves) n_moves[k] = (int64_t)(r->t_moves.size() / 3);
    if (tsv_len) tsv_len[k] = (int64_t)r->tsv.size();
  }
  return NW_OK;
}
int nw_results_gather(nw_result* const* rs, int64_t n, float* score, int32_t* n_moves, int64_t* move_off, int64_t* cost_bp,
                      int64_t* total_bp, int64_t* index_id, int32_t* moves3, char* tsv, nw_stats* stats) {
  if (!rs || n < 0) return NW_ERR_ARG;
  size_t at = 0, mt = 0, tt = 0;  // trajectories, move triples, TSV bytes written so far
  for (int64_t k = 0; k < n; ++k) {
    const nw_result* r = rs[k];
    if (!r) return NW_ERR_ARG;
    const int rc = nw_result_arrays(r, score ? score + at : nullptr, n_moves ? n_moves + at : nullptr,
                                    move_off ? move_off + at : nullptr, cost_bp ? cost_bp + at : nullptr,
                                    total_bp ? total_bp + at : nullptr, index_id ? index_id + at : nullptr,
                                    moves3 ? moves3 + mt * 3 : nullptr);
    if (rc != NW_OK) return rc;
    if (tsv && !r->tsv.empty()) memcpy(tsv + tt, r->tsv.data(), r->tsv.size());
    if (stats) stats[k] = r->stats;
    at += r->t_score.size();
    mt += r->t_moves.size() / 3;
    tt += r->tsv.size();
  }
  return NW_OK;
}
int64_t nw_result_move_count(const nw_result* r) { return r ? (int64_t)(r->t_moves.size() / 3) : 0; }
int nw_result_tsv(const nw_result* r, const char** data, int64_t* len) {
  if (!r || !data || !len) return NW_ERR_ARG;
  *data = r->tsv.data();
  *len = (int64_t)r->tsv.size();
  return NW_OK;
}
int nw_result_stats(const nw_result* r, nw_stats* out) {
  if (!r || !out) return NW_ERR_ARG;
  *out = r->stats;
  return NW_OK;
}
int nw_result_ids(const nw_result* r, float* score, int64_t* cost_bp, int64_t* total_bp, int32_t* level) {
  if (!r || r->id_score.empty()) return NW_ERR_ARG;
  const size_t n = r->id_score.size();
  if (score) memcpy(score, r->id_score.data(), n * 4);
  if (cost_bp) memcpy(cost_bp, r->id_cost.data(), n * 8);
  if (total_bp) memcpy(total_bp, r->id_total.data(), n * 8);
  if (level) memcpy(level, r->id_level.data(), n * 4);
  return NW_OK;
}
int64_t nw_result_edge_count(const nw_result* r) { return r ? (int64_t)r->e_child.size() : 0; }
int nw_result_edges(const nw_result* r, int32_t* child, int32_t* parent, int32_t* moves3) {
  if (!r) return NW_ERR_ARG;
  const size_t n = r->e_child.size();
  if (child) memcpy(child, r->e_child.data(), n * 4);
  if (parent) memcpy(parent, r->e_parent.data(), n * 4);
  if (moves3) memcpy(moves3, r->e_moves.data(), n * 12);
  return NW_OK;
}
void nw_result_free(nw_result* r) { delete r; }

}  // extern "C" (reopened below)

// Scores n raw candidate indices that belong to R regions (rid[k] = region of candidate k; nullptr: all region 0)
// through the banded-DP kernels in ONE bucketed pass.  nw_score_batch is the R == 1 case; the region pre-pass
// (nw_region.cu) hands in the depth-1 children of a whole batch of regions at once.
namespace nw {
// depth_floor: lower bound of the int32-range depth the region contexts are built for; reuse_regions != 0: the region
// contexts of the previous call on this context (same problems, depth_floor covering this call) are kept and only the
// corridor mode is switched -- the pre-pass scores the same regions twice (forced reference row, then the children).
int score_batch_regions(nw_ctx* c, const ScoreProblem* probs, int R, const uint16_t* rid, const int16_t* flat,
                        const int64_t* off, int64_t n, int force, int flags, int64_t* cost_bp, int64_t* total_bp,
                        double* score, int depth_floor, int reuse_regions) {
  if (!c || !probs || R < 1 || !flat || !off || n < 0) {
    g_last_error = "null argument";
    return NW_ERR_ARG;
  }
  PoolScope pool_scope(c);
  const int rc = guard([&] {
    CK(cudaSetDevice(c->device));
    if (n == 0) return;
    if (n >= (1ll << 31)) fail(NW_ERR_ARG, "too many candidates");
    if (force < 0 || force > 2) fail(NW_ERR_ARG, "force must be 0 (slow_score), 1 (forcecalc) or 2 (R twin corridor)");
    cudaStream_t s = c->stream;
    int maxlen = 1;
    int depth_equiv = std::max(0, depth_floor);  // depth bound for the int32 range: a block may occur up to len/n_x times
    for (int64_t k = 0; k < n; ++k) {
      const int r = rid ? rid[k] : 0;
      if (r >= R) fail(NW_ERR_ARG, "candidate region out of range");
      const int n_x = probs[r].n_x;
      const int64_t l = off[k + 1] - off[k];
      if (l < 1 || l > LMAX) fail(NW_ERR_ARG, "candidate length out of range");
      maxlen = std::max<int>(maxlen, (int)l);
      while (((int64_t)n_x << depth_equiv) < l) ++depth_equiv;
      for (int64_t t = off[k]; t < off[k + 1]; ++t) {
        const int v = flat[t];
        if (v == 0 || v > n_x || v < -n_x) fail(NW_ERR_ARG, "candidate element out of range");
      }
    }
    if (reuse_regions && c->n_regions == R && depth_equiv <= depth_floor) {
      for (int r = 0; r < R; ++r) {  // same contexts, other corridor tables
        const int ny = c->RH[r].n_y;
        BandStore& bs = c->band_cache[ny * 3 + force];
        if (bs.n_y != ny || bs.force != force) bs.reset(ny, force);
        c->rbands[r] = &bs;
      }
    } else {
      std::vector<ProblemRef> P(R);
      for (int r = 0; r < R; ++r) P[r] = ProblemRef{probs[r].aln, probs[r].n_x, probs[r].n_y, probs[r].len_x, probs[r].len_y};
      setup_regions(c, P.data(), R, depth_equiv, force);
    }
    // candidates as a "frontier" of n parents scored through KIND_ID descriptors
    Level L;
    L.P = (int)n;
    L.LS = round_ls(maxlen);
    L.pstride = (uint32_t)L.LS * 2;  // sequence only (no prefix tables needed for scoring)
    std::vector<int16_t> seqs((size_t)n * L.LS, 0);
    std::vector<int32_t> lens(n);
    std::vector<uint64_t> desc(n);
    std::vector<uint32_t> newlist(n);
    std::vector<uint16_t> nlen(n);
    const int lcap = maxlen + 1;
    std::vector<uint16_t> nrid(n, 0);
    std::vector<int32_t> k3(n, 0), cost(n, -1), total(n, 0);
    for (int64_t k = 0; k < n; ++k) {
      const int l = (int)(off[k + 1] - off[k]);
      memcpy(&seqs[(size_t)k * L.LS], flat + off[k], (size_t)l * 2);
      lens[k] = l;
      desc[k] = pack_desc((uint32_t)k, 0, 0, KIND_ID);
      newlist[k] = (uint32_t)k;
      nlen[k] = (uint16_t)l;
      if (rid) nrid[k] = rid[k];
    }
    L.G = (uint32_t)n;
    L.n_new = (uint32_t)n;
    h2d(c, L.fblob.ensure(seqs.size() * 2), seqs.data(), seqs.size() * 2);
    h2d(c, L.flen.ensure(n * 4), lens.data(), n * 4);
    h2d(c, L.desc.ensure(n * 8), desc.data(), n * 8);
    h2d(c, L.newlist.ensure(n * 4), newlist.data(), n * 4);
    h2d(c, c->new_len.ensure(n * 2, true), nlen.data(), n * 2);
    h2d(c, L.nrid.ensure(n * 2), nrid.data(), n * 2);
    h2d(c, L.k3.ensure(n * 4), k3.data(), n * 4);
    h2d(c, L.cost.ensure(n * 4), cost.data(), n * 4);
    h2d(c, L.total.ensure(n * 4), total.data(), n * 4);
    int32_t* misc = (int32_t*)c->misc.ensure(256);
    CK(cudaMemsetAsync(misc, 0, 256, s));
    CK(spin_sync(s));
    {
      ScoreView V = view_of(L);
      V.F.regions = c->d_regions.as<RegionDev>();
      V.F.rid = R > 1 ? L.nrid.as<uint16_t>() : nullptr;  // every candidate is its own "parent": same region array
      int* dummy_best = (int*)c->best.ensure((size_t)R * 8 + 16);
      CK(cudaMemsetAsync(dummy_best, 0, (size_t)R * 8 + 16, s));
      const size_t nbins = (size_t)R * lcap * NSUB;
      if (nbins > ((size_t)64 << 20)) fail(NW_ERR_NOMEM, "too many (region, length) buckets in one batch");
      uint32_t* hist = (uint32_t*)c->hist.ensure(nbins * 4, true);
      CK(cudaMemsetAsync(hist, 0, nbins * 4, s));
      c->launches += launch_len_hist(c->new_len.as<uint16_t>(), R > 1 ? L.nrid.as<uint16_t>() : nullptr,
                                     c->d_regions.as<RegionDev>(), lcap, (uint32_t)n, force, hist, s);
      run_scoring(c, V, hist, lcap, L.nrid.as<uint16_t>(), force, flags, dummy_best, nullptr);
    }
    d2h(c, k3.data(), L.k3.p, n * 4);
    d2h(c, cost.data(), L.cost.p, n * 4);
    d2h(c, total.data(), L.total.p, n * 4);
    CK(spin_sync(s));
    CK(cudaGetLastError());
    for (int64_t k = 0; k < n; ++k) {
      const int64_t g = c->RH[rid ? rid[k] : 0].gcd;
      if (cost_bp) cost_bp[k] = cost[k] < 0 ? cost[k] : (int64_t)cost[k] * g;
      if (total_bp) total_bp[k] = (int64_t)total[k] * g;
      if (score) score[k] = k3[k] < 0 ? -INFINITY : k3_to_score(k3[k]);
    }
  });
  if (rc != NW_OK) cudaGetLastError();
  return rc;
}
}  // namespace nw

extern "C" {

int nw_score_batch(nw_ctx* c, const int8_t* aln, int32_t n_x, int32_t n_y, const int64_t* len_x, const int64_t* len_y,
                   const int16_t* flat, const int64_t* off, int64_t n, int32_t force, int32_t flags, int64_t* cost_bp,
                   int64_t* total_bp, double* score) {
  if (!c || !aln || !len_x || !len_y || !flat || !off || n < 0) {
    g_last_error = "null argument";
    return NW_ERR_ARG;
  }
  const nw::ScoreProblem P{aln, n_x, n_y, len_x, len_y};
  return nw::score_batch_regions(c, &P, 1, nullptr, flat, off, n, force, flags, cost_bp, total_bp, score, 0, 0);
}

int nw_moveset(nw_ctx* c, const int8_t* aln, int32_t n_x, int32_t n_y, uint8_t* dup_del, uint8_t* invert) {
  if (!c || !aln || !dup_del || !invert) {
    g_last_error = "null argument";
    return NW_ERR_ARG;
  }
  PoolScope pool_scope(c);
  const int rc = guard([&] {
    CK(cudaSetDevice(c->device));
    std::vector<int64_t> lx(n_x, 1), ly(n_y, 1);
    {
      const ProblemRef P{aln, n_x, n_y, lx.data(), ly.data()};
      setup_regions(c, &P, 1, 1, NW_BAND_JULIA);
    }
    const int MW = c->RH[0].MW;
    std::vector<uint32_t> dd((size_t)(n_x + 1) * MW), iv((size_t)(n_x + 1) * MW);
    d2h(c, dd.data(), c->RD[0].ms_dd, dd.size() * 4);
    d2h(c, iv.data(), c->RD[0].ms_inv, iv.size() * 4);
    CK(spin_sync(c->stream));
    CK(cudaGetLastError());
    for (int a = 1; a <= n_x; ++a)
      for (int b = 1; b <= n_x; ++b) {
        dup_del[(size_t)(a - 1) * n_x + (b - 1)] = (dd[(size_t)a * MW + (b >> 5)] >> (b & 31)) & 1u;
        invert[(size_t)(a - 1) * n_x + (b - 1)] = (iv[(size_t)a * MW + (b >> 5)] >> (b & 31)) & 1u;
      }
  });
  if (rc != NW_OK) cudaGetLastError();
  return rc;
}

int nw_solve_many_ctxs(nw_ctx** ctxs, int32_t n_ctx, const nw_problem* probs, int64_t n, const nw_opts* opts,
                       int32_t max_regions_per_pass, nw_result** outs) {
  if (!ctxs || n_ctx < 1 || !probs || n < 0 || !opts || !outs) {
    g_last_error = "null argument";
    return NW_ERR_ARG;
  }
  for (int t = 0; t < n_ctx; ++t)
    if (!ctxs[t]) {
      g_last_error = "null context";
      return NW_ERR_ARG;
    }
  for (int64_t k = 0; k < n; ++k) outs[k] = nullptr;
  int per = max_regions_per_pass > 0 ? max_regions_per_pass : 256;
  per = std::min(per, 4096);
  const int64_t n_pass = (n + per - 1) / per;
  // Passes of equal estimated work when several contexts share the batch: regions are sorted by estimated work (cells of
  // the gridmatrix x off-diagonal alignments, the source of moves) and dealt round-robin into the passes -- the
  // longest-processing-time rule applied to the passes --, so every pass is the same mix of heavy and light regions:
  // contexts / GPUs finish together, and every pass needs the same buffer sizes (a pass far heavier than the ones a
  // context's memory pool has seen costs a cudaMalloc, which stalls every context on the device).  The reference starts
  // regions in file order from a PSOCK pool (R/nahrwhals.R:171-215).  Results go back to the caller's order.
  std::vector<int64_t> order;
  std::vector<nw_problem> sorted;
  if (n_ctx > 1 && n_pass > 1) {
    std::vector<double> cost((size_t)n, 0.0);
    for (int64_t k = 0; k < n; ++k) {
      const nw_problem& q = probs[k];
      if (!q.aln_sign || q.n_x < 1 || q.n_y < 1) continue;
      int64_t nnz = 0;
      const int64_t cells = (int64_t)q.n_x * q.n_y;
      for (int64_t t = 0; t < cells; ++t) nnz += q.aln_sign[t] != 0;
      const int64_t extra = std::max<int64_t>(0, nnz - std::min(q.n_x, q.n_y));
      cost[k] = (double)cells * (1.0 + (double)extra);
    }
    std::vector<int64_t> by_cost((size_t)n);
    for (int64_t k = 0; k < n; ++k) by_cost[k] = k;
    std::stable_sort(by_cost.begin(), by_cost.end(), [&](int64_t a, int64_t b) { return cost[a] > cost[b]; });
    // pass p takes the regions at sorted positions p, p + n_pass, p + 2 n_pass, ...; passes are contiguous runs of `per`
    order.assign((size_t)n, 0);
    int64_t at = 0;
    for (int64_t pass = 0; pass < n_pass; ++pass) {
      const int64_t m = std::min<int64_t>(per, n - pass * per);
      for (int64_t j = 0, src = pass; j < m; ++j, src += n_pass) {
        // the last passes may run out of their stride: fall back to the unused tail in order
        order[at++] = src < n ? by_cost[src] : -1;
      }
    }
    {  // repair the (rare) holes left by a short last pass
      std::vector<char> used((size_t)n, 0);
      for (int64_t k = 0; k < n; ++k)
        if (order[k] >= 0) used[order[k]] = 1;
      int64_t spare = 0;
      for (int64_t k = 0; k < n; ++k)
        if (order[k] < 0) {
          while (used[by_cost[spare]]) ++spare;
          order[k] = by_cost[spare];
          used[by_cost[spare]] = 1;
        }
    }
    sorted.resize((size_t)n);
    for (int64_t k = 0; k < n; ++k) sorted[k] = probs[order[k]];
    probs = sorted.data();
  }
  std::vector<nw_result*> outs_sorted;
  nw_result** outs_caller = outs;
  if (!order.empty()) {
    outs_sorted.assign((size_t)n, nullptr);
    outs = outs_sorted.data();
  }
  std::atomic<int64_t> next{0};
  std::atomic<int> first_err{NW_OK};
  std::mutex err_mu;
  std::string err_msg;
  // passes are handed out dynamically; while one context's pass is on the device the others prepare the next
  // regions / enumerate the previous report on the host
  auto worker = [&](int t) {
    nw_ctx* c = ctxs[t];
    PoolScope pool_scope(c);
    std::vector<ProblemRef> P;
    std::vector<nw_result*> R;
    for (;;) {
      const int64_t pass = next.fetch_add(1);
      if (pass >= n_pass || first_err.load() != NW_OK) return;
      const int64_t k0 = pass * per;
      const int m = (int)std::min<int64_t>(per, n - k0);
      const int rc = guard([&] {
        check_opts(opts);
        P.resize(m);
        R.resize(m);
        for (int k = 0; k < m; ++k) {
          const nw_problem& q = probs[k0 + k];
          if (!q.aln_sign || !q.len_x || !q.len_y) fail(NW_ERR_ARG, "null problem buffer");
          P[k] = ProblemRef{q.aln_sign, q.n_x, q.n_y, q.len_x, q.len_y};
        }
        for (int k = 0; k < m; ++k) {
          R[k] = new nw_result();
          outs[k0 + k] = R[k];
        }
        do_solve_many(c, P.data(), m, opts, R.data());
      });
      if (rc != NW_OK) {
        cudaGetLastError();
        std::lock_guard<std::mutex> lk(err_mu);
        if (first_err.load() == NW_OK) {
          err_msg = g_last_error;
          first_err.store(rc);
        }
        return;
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_ctx && t < n_pass; ++t) th.emplace_back(worker, t);
  worker(0);
  for (auto& x : th) x.join();
  const int rc = first_err.load();
  if (rc != NW_OK) {
    g_last_error = err_msg;
    for (int64_t k = 0; k < n; ++k) {
      delete outs[k];
      outs[k] = nullptr;
    }
  }
  if (!order.empty())
    for (int64_t k = 0; k < n; ++k) outs_caller[order[k]] = outs[k];
  return rc;
}

int nw_solve_many(nw_ctx* c, const nw_problem* probs, int64_t n, const nw_opts* opts, int32_t max_regions_per_pass,
                  nw_result** outs) {
  return nw_solve_many_ctxs(&c, 1, probs, n, opts, max_regions_per_pass, outs);
}

int nw_solve_batch(nw_ctx** ctxs, int32_t n_ctx, const nw_problem* probs, int64_t n, const nw_opts* opts,
                   nw_result** outs, int32_t* rcs) {
  if (!ctxs || n_ctx < 1 || !probs || n < 0 || !opts || !outs) {
    g_last_error = "null argument";
    return NW_ERR_ARG;
  }
  for (int t = 0; t < n_ctx; ++t)
    if (!ctxs[t]) {
      g_last_error = "null context";
      return NW_ERR_ARG;
    }
  std::atomic<int64_t> next{0};
  std::atomic<int> first_err{NW_OK};
  std::mutex err_mu;
  std::string err_msg;
  auto worker = [&](int t) {
    for (;;) {
      const int64_t k = next.fetch_add(1);
      if (k >= n) return;
      outs[k] = nullptr;
      const nw_problem& P = probs[k];
      const int rc = nw_solve(ctxs[t], P.aln_sign, P.n_x, P.n_y, P.len_x, P.len_y, opts, &outs[k]);
      if (rcs) rcs[k] = rc;
      if (rc != NW_OK) {
        std::lock_guard<std::mutex> lk(err_mu);
        if (first_err.load() == NW_OK) {
          first_err.store(rc);
          err_msg = "problem " + std::to_string(k) + ": " + nw_last_error();
        }
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_ctx; ++t) th.emplace_back(worker, t);
  worker(0);
  for (auto& x : th) x.join();
  if (first_err.load() != NW_OK) g_last_error = err_msg;
  return first_err.load();
}

// ---------------------------------------------------------------------------------------------------
// frontier-sharded single search (BASELINE config 3): see include/nwsolve.h
// ---------------------------------------------------------------------------------------------------
int nw_dist_unique_id(void* id128) {
  if (!id128) return NW_ERR_ARG;
  return guard([&] {
    try {
      nccl_unique_id(id128);
    } catch (const ExchangeError& e) {
      fail(NW_ERR_CUDA, e.msg);
    }
  });
}
int nw_dist_init(nw_ctx* c, const void* id128, int32_t rank, int32_t world) {
  if (!c || !id128) {
    g_last_error = "null argument";
    return NW_ERR_ARG;
  }
  return guard([&] {
    if (world < 1 || rank < 0 || rank >= world) fail(NW_ERR_ARG, "bad rank/world");
    CK(cudaSetDevice(c->device));
    delete c->xch;
    c->xch = nullptr;
    try {
      c->xch = nccl_exchange_create(id128, rank, world);
    } catch (const ExchangeError& e) {
      fail(NW_ERR_CUDA, e.msg);
    }
  });
}
void nw_dist_shutdown(nw_ctx* c) {
  if (!c || !c->xch) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  delete c->xch;
  c->xch = nullptr;
}
int nw_dist_info(nw_ctx* c, int32_t* rank, int32_t* world) {
  if (!c) return NW_ERR_ARG;
  if (rank) *rank = c->xch ? c->xch->rank() : 0;
  if (world) *world = c->xch ? c->xch->world() : 1;
  return NW_OK;
}
int nw_solve_sharded(nw_ctx* c, const int8_t* aln, int32_t n_x, int32_t n_y, const int64_t* len_x, const int64_t* len_y,
                     const nw_opts* o, nw_result** out) {
  if (!c || !aln || !len_x || !len_y || !out) {
    g_last_error = "null argument";
    return NW_ERR_ARG;
  }
  *out = nullptr;
  nw_result* R = new nw_result();
  PoolScope pool_scope(c);
  const int rc = guard([&] {
    check_opts(o);
    if (!c->xch) fail(NW_ERR_ARG, "nw_solve_sharded: the context has no communicator (call nw_dist_init first)");
    const ProblemRef P{aln, n_x, n_y, len_x, len_y};
    do_solve_many(c, &P, 1, o, &R, c->xch);
  });
  if (rc != NW_OK) {
    delete R;
    cudaGetLastError();
    return rc;
  }
  *out = R;
  return NW_OK;
}
int nw_solve_sharded_ctxs(nw_ctx** ctxs, int32_t world, const int8_t* aln, int32_t n_x, int32_t n_y, const int64_t* len_x,
                          const int64_t* len_y, const nw_opts* o, nw_result** outs) {
  if (!ctxs || world < 1 || !aln || !len_x || !len_y || !o || !outs) {
    g_last_error = "null argument";
    return NW_ERR_ARG;
  }
  for (int r = 0; r < world; ++r) {
    outs[r] = nullptr;
    if (!ctxs[r]) {
      g_last_error = "null context";
      return NW_ERR_ARG;
    }
    for (int q = 0; q < r; ++q)
      if (ctxs[q] == ctxs[r]) {
        g_last_error = "nw_solve_sharded_ctxs: every rank needs its own context";
        return NW_ERR_ARG;
      }
  }
  // direct peer copies between the ranks' devices (NVLink) where the contexts sit on different GPUs
  for (int r = 0; r < world; ++r)
    for (int q = 0; q < world; ++q)
      if (ctxs[r]->device != ctxs[q]->device) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, ctxs[r]->device, ctxs[q]->device);
        if (can) {
          cudaSetDevice(ctxs[r]->device);
          cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
          cudaGetLastError();  // already enabled is fine
        }
      }
  std::shared_ptr<LocalGroup> grp = local_group_create(world);
  std::vector<int> rcs(world, NW_OK);
  std::vector<std::string> msgs(world);
  std::vector<nw_result*> res(world, nullptr);
  auto worker = [&](int r) {
    nw_ctx* c = ctxs[r];
    PoolScope pool_scope(c);
    std::unique_ptr<Exchange> x(local_exchange_create(grp, r, c->device));
    res[r] = new nw_result();
    rcs[r] = guard([&] {
      check_opts(o);
      const ProblemRef P{aln, n_x, n_y, len_x, len_y};
      nw_result* R = res[r];
      do_solve_many(c, &P, 1, o, &R, x.get());
    });
    if (rcs[r] != NW_OK) {
      msgs[r] = g_last_error;
      x->abort();  // the other ranks stop at their next exchange instead of waiting for this one
      cudaGetLastError();
    }
  };
  std::vector<std::thread> th;
  for (int r = 1; r < world; ++r) th.emplace_back(worker, r);
  worker(0);
  for (auto& t : th) t.join();
  int rc = NW_OK;
  for (int r = 0; r < world && rc == NW_OK; ++r)
    if (rcs[r] != NW_OK) {
      rc = rcs[r];
      g_last_error = "rank " + std::to_string(r) + ": " + msgs[r];
    }
  for (int r = 0; r < world; ++r) {
    if (rc == NW_OK) {
      outs[r] = res[r];
    } else {
      delete res[r];
    }
  }
  return rc;
}

int nw_solve_files_sharded(nw_ctx** ctxs, int32_t world, const char* inmat, const char* inlens, const char* outp,
                           const nw_opts* o, nw_stats* stats) {
  if (!ctxs || world < 1 || !inmat || !inlens || !outp || !o) {
    g_last_error = "null argument";
    return NW_ERR_ARG;
  }
  ProblemHost P;
  int rc = guard([&] { P = read_problem(inmat, inlens); });
  if (rc != NW_OK) return rc;
  std::vector<nw_result*> res((size_t)world, nullptr);
  nw_opts oo = *o;
  rc = nw_solve_sharded_ctxs(ctxs, world, P.aln.data(), P.n_x, P.n_y, P.len_x.data(), P.len_y.data(), &oo, res.data());
  if (rc == NW_OK) {
    rc = guard([&] { write_file_atomic(outp, res[0]->tsv); });
    if (rc == NW_OK && stats) *stats = res[0]->stats;
  }
  for (nw_result* r : res) delete r;
  return rc;
}

int nw_bench_peaks(nw_ctx* c, double out[4]) {
  if (!c || !out) return NW_ERR_ARG;
  PoolScope pool_scope(c);
  const int rc = guard([&] {
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    int* sink = (int*)c->scratch.ensure((size_t)148 * 8 * 256 * 4);
    for (int mode = 0; mode < 3; ++mode) {
      double best = 0;
      for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(a, s));
        const double n = launch_int_peak(mode, sink, 4096, s);
        CK(cudaEventRecord(b, s));
        CK(cudaEventSynchronize(b));
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (rep) best = std::max(best, n / (ms * 1e-3));
      }
      out[mode] = best;
    }
    const size_t bytes = (size_t)1 << 30;
    DevBuf x, y;
    x.ensure(bytes);
    y.ensure(bytes);
    double bw = 0;
    for (int rep = 0; rep < 4; ++rep) {
      CK(cudaEventRecord(a, s));
      CK(cudaMemcpyAsync(y.p, x.p, bytes, cudaMemcpyDeviceToDevice, s));
      CK(cudaEventRecord(b, s));
      CK(cudaEventSynchronize(b));
      float ms = 0;
      cudaEventElapsedTime(&ms, a, b);
      if (rep) bw = std::max(bw, 2.0 * bytes / (ms * 1e-3));
    }
    out[3] = bw;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    CK(cudaGetLastError());
  });
  if (rc != NW_OK) cudaGetLastError();
  return rc;
}

int nw_read_problem(const char* inmat, const char* inlens, int8_t** aln, int32_t* n_x, int32_t* n_y, int64_t** len_x,
                    int64_t** len_y) {
  if (!inmat || !inlens || !aln || !n_x || !n_y || !len_x || !len_y) return NW_ERR_ARG;
  return guard([&] {
    ProblemHost P = read_problem(inmat, inlens);
    *n_x = P.n_x;
    *n_y = P.n_y;
    *aln = (int8_t*)malloc(P.aln.size());
    *len_x = (int64_t*)malloc(P.len_x.size() * 8);
    *len_y = (int64_t*)malloc(P.len_y.size() * 8);
    memcpy(*aln, P.aln.data(), P.aln.size());
    memcpy(*len_x, P.len_x.data(), P.len_x.size() * 8);
    memcpy(*len_y, P.len_y.data(), P.len_y.size() * 8);
  });
}
void nw_free(void* p) { free(p); }

}  // extern "C"
