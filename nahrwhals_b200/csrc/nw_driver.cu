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
#include <atomic>
#include <mutex>
#include <thread>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/nwsolve.h"
#include "nw_band.h"
#include "nw_host.h"
#include "nw_kernels.cuh"
#include "nw_launch.h"

using namespace nw;

static thread_local std::string g_last_error;
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

struct BandStore {  // host cache + device mirror of per-length corridor tables for one (n_y, force)
  int n_y = -1;
  bool force = false;
  std::map<int, BandTable> tabs;
  std::vector<int32_t> off;  // [LMAX+2] row offset or -1
  std::vector<int16_t> ab, eb;
  std::vector<uint8_t> sb;
  std::vector<uint32_t> mask;
  DevBuf d_off, d_ab, d_eb, d_sb, d_mask;
  bool dirty = true;
  void reset(int ny, bool f) {
    n_y = ny;
    force = f;
    tabs.clear();
    off.assign(LMAX + 2, -1);
    ab.clear();
    eb.clear();
    sb.clear();
    mask.clear();
    dirty = true;
  }
  const BandTable& get(int L) {
    auto it = tabs.find(L);
    if (it != tabs.end()) return it->second;
    BandTable T = build_band(L, n_y, force, WINDOW_MAX);
    off[L] = (int32_t)eb.size();
    for (int i = 0; i <= L; ++i) {
      ab.push_back((int16_t)(i ? T.row[i].a : 1));
      ab.push_back((int16_t)(i ? T.row[i].b : 0));
      eb.push_back(T.eb[i]);
      sb.push_back(T.sb[i]);
      for (int w = 0; w < 4; ++w) mask.push_back(T.mask[(size_t)i * 4 + w]);
    }
    dirty = true;
    return tabs.emplace(L, std::move(T)).first->second;
  }
  void upload(nw_ctx* c, cudaStream_t st) {
    if (!dirty) return;
    h2d(c, d_off.ensure(off.size() * 4), off.data(), off.size() * 4);
    h2d(c, d_ab.ensure(ab.size() * 2, true), ab.data(), ab.size() * 2);
    h2d(c, d_eb.ensure(eb.size() * 2, true), eb.data(), eb.size() * 2);
    h2d(c, d_sb.ensure(sb.size(), true), sb.data(), sb.size());
    h2d(c, d_mask.ensure(mask.size() * 4, true), mask.data(), mask.size() * 4);
    CK(spin_sync(st));  // host vectors may be reallocated by the next get()
    dirty = false;
  }
};

}  // namespace

struct nw_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::vector<U4> h_pow;
  DevBuf d_pow;
  // region
  RegionHost RH;
  RegionDev RD{};
  DevBuf d_ctx, d_rt, d_cumr, d_msdd, d_msinv, d_msany;
  std::map<int, BandStore> band_cache;  // corridor tables per (n_y, force), kept across regions
  BandStore* bands_p = nullptr;
  // dedup table
  DevBuf table;
  uint64_t table_cap = 0;
  // per-level scratch
  DevBuf cnt, off, bsum, slot, owner, is_new, newrank, new_len, hist, bstart, cursor, work, scratch, flag, pre, sel, keys,
      misc, rslot, newflag, grank, edgebuf;
  void* dist = nullptr;  // active frontier-sharded search (Search*, nw_dist_*)
  void* pinned = nullptr;  // small pinned staging area for D2H scalars / histograms
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
constexpr size_t HIST_N = (size_t)(LMAX + 2) * NSUB;  // (child length, sub-counter) bins

struct Level {
  // frontier expanded at this level
  int P = 0, LS = 0;
  uint32_t pstride = 0;
  DevBuf fblob, flen, fid, fdup;
  // candidates generated at this level
  uint32_t G = 0, gbase = 0, n_new = 0;
  int32_t id_base = 0;
  DevBuf desc, cand_id, newlist, k3, cost, total;
  // frontier-sharded search only: G counts this rank's candidates, G_all all ranks'; k3/cost/total are the merged
  // (replicated) per-id arrays, l_* the rank-local ones of the candidates this rank scored
  uint32_t G_all = 0, n_local_new = 0;
  DevBuf off_loc, off_glob, desc_new, gid_at, l_k3, l_cost, l_total;
  FrontierDev frontier() const {
    FrontierDev F;
    F.P = P;
    F.LS = LS;
    F.pstride = pstride;
    F.blob = fblob.as<unsigned char>();
    F.len = flen.as<int32_t>();
    F.id = fid.as<int32_t>();
    F.dupok = fdup.as<uint8_t>();
    return F;
  }
};

void setup_region(nw_ctx* c, const int8_t* aln, int n_x, int n_y, const int64_t* len_x, const int64_t* len_y,
                  int maxdepth, bool force) {
  try {
    c->RH = build_region(aln, n_x, n_y, len_x, len_y, maxdepth);
  } catch (const HostError& e) {
    fail(e.code, e.what());
  }
  RegionHost& H = c->RH;
  cudaStream_t st = c->stream;
  h2d(c, c->d_ctx.ensure(H.ctx.size()), H.ctx.data(), H.ctx.size());
  h2d(c, c->d_rt.ensure(H.rt.size() * 4), H.rt.data(), H.rt.size() * 4);
  h2d(c, c->d_cumr.ensure(H.cumr.size() * 4), H.cumr.data(), H.cumr.size() * 4);
  const size_t msb = (size_t)(n_x + 1) * H.MW * 4;
  CK(cudaMemsetAsync(c->d_msdd.ensure(msb), 0, msb, st));
  CK(cudaMemsetAsync(c->d_msinv.ensure(msb), 0, msb, st));
  CK(cudaMemsetAsync(c->d_msany.ensure((size_t)(n_x + 1) * 4), 0, (size_t)(n_x + 1) * 4, st));
  RegionDev& D = c->RD;
  D.n_x = n_x;
  D.n_y = n_y;
  D.NYP = H.NYP;
  D.PWS = H.PWS;
  D.MW = H.MW;
  D.sum_rt = (int)H.sum_rt;
  D.rid1 = 1;
  D.ctx_bytes = (uint32_t)H.ctx.size();
  D.ctx = c->d_ctx.as<unsigned char>();
  D.off_upx = H.off_upx;
  D.off_rtrev = H.off_rtrev;
  D.rt_copy_stride = H.rt_copy_stride;
  D.rt = c->d_rt.as<int32_t>();
  D.cumr = c->d_cumr.as<int32_t>();
  D.ms_dd = c->d_msdd.as<uint32_t>();
  D.ms_inv = c->d_msinv.as<uint32_t>();
  D.ms_any = c->d_msany.as<uint32_t>();
  if (D.ctx_bytes > 200 * 1024) fail(NW_ERR_RANGE, "region context exceeds shared memory");
  c->launches += launch_moveset(D, st);
  BandStore& bs = c->band_cache[n_y * 2 + (force ? 1 : 0)];
  if (bs.n_y != n_y || bs.force != force) bs.reset(n_y, force);
  c->bands_p = &bs;
}

void table_reserve(nw_ctx* c, uint64_t need_entries) {
  uint64_t want = 1024;
  while (want < need_entries * 2) want <<= 1;
  if (want <= c->table_cap) return;
  if (want > (1ull << 32)) fail(NW_ERR_NOMEM, "dedup table would exceed 2^32 slots");
  const size_t bytes = want * TBL_WORDS * 8;
  if (c->table_cap == 0) {  // fresh search: reuse the physical buffer when it is large enough
    c->table.ensure(bytes);
    CK(cudaMemsetAsync(c->table.p, 0xFF, bytes, c->stream));
    c->table_cap = want;
    return;
  }
  DevBuf nt;
  nt.ensure(bytes);
  CK(cudaMemsetAsync(nt.p, 0xFF, bytes, c->stream));
  c->launches += launch_rehash(c->table.as<uint64_t>(), c->table_cap, nt.as<uint64_t>(), want - 1, c->stream);
  CK(spin_sync(c->stream));
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
// `lmax` bounds the child lengths present (bins above it are untouched: keeps the per-level copies small).
void run_scoring(nw_ctx* c, const ScoreView& Lv, const uint32_t* hist_h, int lmax, int force, int flags, int* best_dev,
                 int64_t* cells_out) {
  lmax = std::min(std::max(lmax, 1), LMAX);
  const size_t nbins = (size_t)(lmax + 1) * NSUB;
  cudaStream_t st = c->stream;
  struct Bucket {
    int L, cls;  // cls: window W (fast) or 0 (generic)
    uint32_t n;
  };
  std::vector<Bucket> bk;
  int64_t cells = 0;
  for (int L = 1; L <= lmax; ++L) {
    uint32_t nL = 0;
    for (int sb = 0; sb < NSUB; ++sb) nL += hist_h[(size_t)L * NSUB + sb];
    if (!nL) continue;
    const BandTable& T = c->bands_p->get(L);
    if (!T.contiguous) fail(NW_ERR_RANGE, "corridor rows are not intervals for this shape (unsupported)");
    int cls = 0;
    if (!(flags & NW_FLAG_GENERIC_DP) && T.fast_ok) cls = score_fast_window(T.max_width);
    if (cls && score_fast_smem(cls, c->RD.ctx_bytes, L) > 200 * 1024) cls = 0;  // tiles would not fit: generic kernel
    bk.push_back(Bucket{L, cls, nL});
    cells += (int64_t)nL * T.cells;
  }
  if (cells_out) *cells_out = cells;
  if (bk.empty()) return;
  c->bands_p->upload(c, st);
  std::stable_sort(bk.begin(), bk.end(), [](const Bucket& a, const Bucket& b) {
    const int ka = a.cls ? a.cls : 1 << 20, kb = b.cls ? b.cls : 1 << 20;
    return ka < kb;
  });
  std::vector<uint32_t> bstart(nbins, 0);
  struct Range {
    int cls;
    uint32_t begin, end;
    int lmax;
  };
  std::vector<Range> ranges;
  uint32_t pos = 0;
  for (const Bucket& b : bk) {
    if (ranges.empty() || ranges.back().cls != b.cls) {
      pos = (pos + 127) & ~127u;  // class ranges start on a CTA boundary
      ranges.push_back(Range{b.cls, pos, pos, 0});
    }
    uint32_t at = pos;  // the sub-counters of one length share one contiguous, 32-padded bucket
    for (int sb = 0; sb < NSUB; ++sb) {
      bstart[(size_t)b.L * NSUB + sb] = at;
      at += hist_h[(size_t)b.L * NSUB + sb];
    }
    pos += (b.n + 31) & ~31u;
    ranges.back().end = pos;
    ranges.back().lmax = std::max(ranges.back().lmax, b.L);
  }
  const uint32_t n_work = pos;
  h2d(c, c->bstart.ensure(HIST_N * 4), bstart.data(), nbins * 4);
  CK(cudaMemsetAsync(c->cursor.ensure(HIST_N * 4), 0, nbins * 4, st));
  CK(cudaMemsetAsync(c->work.ensure((size_t)n_work * 4, true), 0xFF, (size_t)n_work * 4, st));
  c->launches += launch_bucket_scatter(c->new_len.as<uint16_t>(), Lv.n_new, c->RD.n_y, force, c->bstart.as<uint32_t>(),
                                       c->cursor.as<uint32_t>(), c->work.as<uint32_t>(), st);
  CK(spin_sync(st));  // bstart host vector goes out of scope
  ScoreArgs A{};
  A.R = c->RD;
  A.F = Lv.F;
  A.newlist = Lv.newlist;
  A.desc = Lv.desc;
  A.band_off = c->bands_p->d_off.as<int32_t>();
  A.band_ab = c->bands_p->d_ab.as<int16_t>();
  A.band_eb = c->bands_p->d_eb.as<int16_t>();
  A.band_sb = c->bands_p->d_sb.as<uint8_t>();
  A.band_mask = c->bands_p->d_mask.as<uint32_t>();
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
      const int l = launch_score_fast(r.cls, A, st);
      if (l < 0) fail(NW_ERR_CUDA, "no register-DP kernel for window " + std::to_string(r.cls));
      c->launches += l;
    } else {
      const size_t sc = (size_t)2 * (c->RD.n_y + 2) * generic_threads() * 4;
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
  int root_k3 = 0, root_cost = -1, root_total = 0;
  int best_k3 = 0;
  nw_stats st{};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<cudaEvent_t> lev;

  int rank = 0, world = 1;      // frontier sharding (nw_dist_*); world == 1 is the plain single-GPU search
  ExpandArgs curE{};            // expansion arguments of the level in flight (sharded search)
  int cur_level = 0;
  bool dist_more = false;
  int dist_k3_min = 0;
  int dist_maxlen = 0;
  uint32_t dist_edges = 0;
  uint32_t edge_cap_hint = 0, id_cap_hint = 0;
  explicit Search(nw_ctx* ctx, const nw_opts& opts) : c(ctx), o(opts) {}
  ~Search() {
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    for (auto e : lev) cudaEventDestroy(e);
  }

  uint32_t* pin_u32() { return reinterpret_cast<uint32_t*>(c->pinned); }

  void init_root() {
    cudaStream_t s = c->stream;
    const int n_x = c->RD.n_x;
    if (n_x > LMAX) fail(NW_ERR_RANGE, "n_x exceeds the supported index length");
    lv.clear();
    lv.emplace_back();  // level 0 (root pseudo level)
    lv.emplace_back();  // level 1 frontier = {root}
    Level& L0 = lv[0];
    Level& L1 = lv[1];
    std::vector<int16_t> root(n_x);
    for (int i = 0; i < n_x; ++i) root[i] = (int16_t)(i + 1);
    L1.P = 1;
    L1.LS = round_ls(n_x);
    L1.pstride = blob_stride(L1.LS);
    std::vector<unsigned char> blob(L1.pstride);
    fill_blob(blob.data(), L1.LS, root.data(), n_x, c->h_pow.data(), c->RD.rid1);
    h2d(c, L1.fblob.ensure(blob.size()), blob.data(), blob.size());
    const int32_t len = n_x, id = 1;
    const uint8_t dupok = (1 <= o.maxdup) ? 1 : 0;  // duplication_count(root) == 1 (solver.jl:611)
    h2d(c, L1.flen.ensure(4), &len, 4);
    h2d(c, L1.fid.ensure(4), &id, 4);
    h2d(c, L1.fdup.ensure(1), &dupok, 1);
    CK(spin_sync(s));
    // level 0: one pseudo candidate (the root itself, KIND_ID over frontier slot 0 of level 1), id 1
    L0.G = 1;
    L0.gbase = 0;
    L0.n_new = 1;
    L0.id_base = 1;
    const uint64_t d0 = pack_desc(0, 0, 0, KIND_ID);
    const int32_t one = 1;
    const uint32_t zero = 0;
    h2d(c, L0.desc.ensure(8), &d0, 8);
    h2d(c, L0.cand_id.ensure(4), &one, 4);
    h2d(c, L0.newlist.ensure(4), &zero, 4);
    L0.k3.ensure(4);
    L0.cost.ensure(4);
    L0.total.ensure(4);
    // frontier view for scoring the root: borrow level 1's frontier
    n_ids = 1;
    gtotal = 1;
    // dedup table
    c->table_cap = 0;  // logical capacity; the physical buffer is kept across solves
    table_reserve(c, 1 << 15);
    c->launches += launch_insert_root(L1.fblob.as<int16_t>(), n_x, c->d_pow.as<U4>(), c->RD.rid1, c->table.as<uint64_t>(),
                                      c->table_cap - 1, s);
    // score the root (solver.jl:599) -- its score never updates `best` (:606)
    int32_t* misc = (int32_t*)c->misc.ensure(256);
    CK(cudaMemsetAsync(misc, 0, 256, s));
    std::vector<uint32_t> hist(HIST_N, 0);
    const uint16_t rl = (uint16_t)n_x;
    h2d(c, c->new_len.ensure(2), &rl, 2);
    if (early_exit(n_x, c->RD.n_y)) {
      root_k3 = 0;
      root_cost = -1;
      root_total = 0;
      const int32_t v[3] = {0, -1, 0};
      h2d(c, L0.k3.p, &v[0], 4);
      h2d(c, L0.cost.p, &v[1], 4);
      h2d(c, L0.total.p, &v[2], 4);
      CK(spin_sync(s));
    } else {
      hist[bucket_key(n_x, 0)] = 1;
      ScoreView view = view_of(L0);  // candidates of level 0 against the frontier of level 1
      view.F = L1.frontier();
      run_scoring(c, view, hist.data(), n_x, 0, o.flags, misc + 1 /*dummy best*/, nullptr);
      int32_t v[3];
      d2h(c, &v[0], L0.k3.p, 4);
      d2h(c, &v[1], L0.cost.p, 4);
      d2h(c, &v[2], L0.total.p, 4);
      CK(spin_sync(s));
      root_k3 = v[0];
      root_cost = v[1];
      root_total = v[2];
    }
  }

  // NW_FLAG_VERIFY_DEDUP: element-wise comparison of every merged candidate of level d with its owner
  void verify_level(int d, const ExpandArgs& E, const uint32_t* is_new, const uint32_t* owner, uint32_t G) {
    cudaStream_t s = c->stream;
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
    CK(cudaMemsetAsync(misc + 12, 0, 4, s));
    c->launches += launch_verify_dups(E, V, is_new, owner, G, (unsigned int*)(misc + 12), s);
    d2h(c, pin_u32(), misc, 64);
    CK(spin_sync(s));
    CK(cudaGetLastError());
    if (pin_u32()[12])
      fail(NW_ERR_COLLISION, std::to_string(pin_u32()[12]) + " candidates of level " + std::to_string(d) +
                                 " were merged into a different index (fingerprint collision)");
  }

  // one BFS level: expand lv[d]'s frontier.  Returns false when the frontier is empty.
  bool run_level(int d) {
    cudaStream_t s = c->stream;
    Level& L = lv[d];
    int32_t* misc = (int32_t*)c->misc.p;  // [0]=best_k3, [1]=dummy, [2]=maxlen, [4..]=totals
    st.frontier[d - 1] = L.P;
    if (L.P == 0) return false;
    ExpandArgs E{};
    E.R = c->RD;
    E.F = L.frontier();
    E.cnt = (uint32_t*)c->cnt.ensure((size_t)L.P * 4, true);
    E.maxlen = misc + 2;
    E.force_pairs = (o.flags & NW_FLAG_PAIR_EXPAND) ? 1 : 0;
    E.weak_fp = (o.flags & NW_FLAG_DEBUG_WEAK_FP) ? 1 : 0;
    CK(cudaMemsetAsync(misc + 2, 0, 4, s));
    c->launches += launch_expand(false, E, s);
    uint32_t* off = (uint32_t*)c->off.ensure((size_t)L.P * 4, true);
    uint32_t* bsum = (uint32_t*)c->bsum.ensure(scan_scratch_items((uint64_t)L.P) * 4, true);
    c->launches += launch_scan(E.cnt, off, (uint64_t)L.P, bsum, (uint32_t*)(misc + 4), s);
    d2h(c, pin_u32(), misc, 32);
    CK(spin_sync(s));
    CK(cudaGetLastError());
    const uint32_t G = pin_u32()[4];
    const int maxlen = (int)pin_u32()[2];
    if (maxlen > LMAX) fail(NW_ERR_RANGE, "a child index exceeds the supported length");
    if ((uint64_t)gtotal + G >= 0xfffffff0ull || G >= 0x7ffffff0u) fail(NW_ERR_NOMEM, "more than 2^32 candidates");
    L.G = G;
    L.gbase = gtotal;
    L.id_base = n_ids + 1;
    st.generated[d - 1] = G;
    if (G == 0) {
      L.n_new = 0;
      return false;
    }
    table_reserve(c, (uint64_t)n_ids + G);
    E.off = off;
    E.desc = (uint64_t*)L.desc.ensure((size_t)G * 8);
    E.slot = (uint32_t*)c->slot.ensure((size_t)G * 4, true);
    E.table = c->table.as<uint64_t>();
    E.tmask = c->table_cap - 1;
    E.gbase = L.gbase;
    E.pw = c->d_pow.as<U4>();
    E.own_mod = 1;
    E.own_rem = 0;
    E.off_glob = nullptr;
    c->launches += launch_expand(true, E, s);
    c->launches += launch_hash_insert(E, G, s);
    uint32_t* is_new = (uint32_t*)c->is_new.ensure((size_t)G * 4, true);
    uint32_t* owner = (uint32_t*)c->owner.ensure((size_t)G * 4, true);
    c->launches += launch_resolve(E, G, is_new, owner, s);
    uint32_t* newrank = (uint32_t*)c->newrank.ensure((size_t)G * 4, true);
    bsum = (uint32_t*)c->bsum.ensure(scan_scratch_items(G) * 4, true);
    c->launches += launch_scan(is_new, newrank, G, bsum, (uint32_t*)(misc + 5), s);
    d2h(c, pin_u32(), misc, 32);
    CK(spin_sync(s));
    CK(cudaGetLastError());
    const uint32_t n_new = pin_u32()[5];
    L.n_new = n_new;
    st.scored[d - 1] = n_new;
    // ids / edges / histogram of child lengths
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
    A.is_new = is_new;
    A.newrank = newrank;
    A.desc = E.desc;
    A.front_len = L.flen.as<int32_t>();
    A.gbase = L.gbase;
    A.G = G;
    A.id_base = L.id_base;
    A.n_y = c->RD.n_y;
    A.force = 0;
    A.cand_id = (int32_t*)L.cand_id.ensure((size_t)G * 4);
    A.newlist = (uint32_t*)L.newlist.ensure((size_t)n_new * 4);
    A.new_len = (uint16_t*)c->new_len.ensure((size_t)n_new * 2, true);
    A.k3 = (int32_t*)L.k3.ensure((size_t)n_new * 4);
    A.cost = (int32_t*)L.cost.ensure((size_t)n_new * 4);
    A.total = (int32_t*)L.total.ensure((size_t)n_new * 4);
    const int lmax = std::min(std::max(maxlen, L.LS), LMAX);  // longest possible child of this level
    const size_t nbins = (size_t)(lmax + 1) * NSUB;
    A.hist = (uint32_t*)c->hist.ensure(HIST_N * 4);
    CK(cudaMemsetAsync(A.hist, 0, nbins * 4, s));
    c->launches += launch_assign(A, s);
    if (o.flags & NW_FLAG_VERIFY_DEDUP) verify_level(d, E, is_new, owner, G);
    uint32_t* hist_h = pin_u32() + 64;
    d2h(c, hist_h, A.hist, nbins * 4);
    CK(spin_sync(s));
    CK(cudaGetLastError());
    int64_t cells = 0;
    run_scoring(c, view_of(L), hist_h, lmax, 0, o.flags, misc + 0, &cells);
    st.cells[d - 1] = cells;
    gtotal += G;
    n_ids += (int32_t)n_new;
    return n_new > 0;
  }

  // frontier of level d+1 from the fresh candidates of level d (solver.jl:551-553, :669-672)
  void select_next(int d) {
    cudaStream_t s = c->stream;
    Level& L = lv[d];
    lv.emplace_back();
    Level& N = lv[d + 1];
    Level& Lr = lv[d];  // re-take reference (emplace_back may reallocate)
    (void)L;
    const uint32_t n = Lr.n_new;
    int32_t* misc = (int32_t*)c->misc.p;
    if (n == 0) {
      N.P = 0;
      return;
    }
    uint32_t* flag = (uint32_t*)c->flag.ensure((size_t)n * 4, true);
    uint32_t* pre = (uint32_t*)c->pre.ensure((size_t)n * 4, true);
    uint32_t* bsum = (uint32_t*)c->bsum.ensure(scan_scratch_items(n) * 4, true);
    c->launches += launch_valid_flags(Lr.k3.as<int32_t>(), n, flag, s);
    c->launches += launch_scan(flag, pre, n, bsum, (uint32_t*)(misc + 6), s);
    d2h(c, pin_u32(), misc, 32);
    CK(spin_sync(s));
    const uint32_t nvalid = pin_u32()[6];
    uint32_t P2 = nvalid;
    uint32_t* sel = (uint32_t*)c->sel.ensure((size_t)std::max<uint32_t>(nvalid, 1) * 4, true);
    if (o.maxwidth < 0) {
      c->launches += launch_compact_ranks(flag, pre, n, sel, s);
    } else {
      uint64_t npad = 2048;
      while (npad < nvalid) npad <<= 1;
      uint64_t* keys = (uint64_t*)c->keys.ensure(npad * 8, true);
      c->launches += launch_make_keys(Lr.k3.as<int32_t>(), flag, pre, n, keys, nvalid, npad, s);
      c->launches += launch_sort_u64(keys, npad, s);
      if ((int64_t)P2 > o.maxwidth) P2 = (uint32_t)o.maxwidth;
      c->launches += launch_keys_to_sel(keys, P2, sel, s);
    }
    N.P = (int)P2;
    if (P2 == 0) return;
    // child lengths are bounded by the level's maxlen (misc[2]) -- use it for the stride
    int maxlen = (int)pin_u32()[2];
    if (maxlen < c->RD.n_x) maxlen = c->RD.n_x;
    // inversions keep the parent's length: include the longest parent
    maxlen = std::max(maxlen, Lr.LS);
    N.LS = round_ls(maxlen);
    N.pstride = blob_stride(N.LS);
    MaterialiseArgs M{};
    M.F = Lr.frontier();
    M.sel = sel;
    M.newlist = world > 1 ? nullptr : Lr.newlist.as<uint32_t>();  // sharded: descriptors of the merged new ids
    M.desc = world > 1 ? Lr.desc_new.as<uint64_t>() : Lr.desc.as<uint64_t>();
    M.id_base = Lr.id_base;
    M.maxdup = o.maxdup;
    M.n_x = c->RD.n_x;
    M.P2 = (int)P2;
    M.LS2 = N.LS;
    M.pstride2 = N.pstride;
    M.blob2 = (unsigned char*)N.fblob.ensure((size_t)P2 * N.pstride);
    M.len2 = (int32_t*)N.flen.ensure((size_t)P2 * 4);
    M.id2 = (int32_t*)N.fid.ensure((size_t)P2 * 4);
    M.dupok2 = (uint8_t*)N.fdup.ensure((size_t)P2);
    M.pw = c->d_pow.as<U4>();
    M.rid1 = c->RD.rid1;
    c->launches += launch_materialise(M, s);
    CK(cudaGetLastError());
  }

  void run() {
    cudaStream_t s = c->stream;
    CK(cudaEventCreate(&ev0));
    CK(cudaEventCreate(&ev1));
    CK(cudaEventRecord(ev0, s));
    init_root();
    int32_t* misc = (int32_t*)c->misc.p;
    CK(cudaMemsetAsync(misc, 0, 4, s));  // best_k3 = 0  (best = 0.0, solver.jl:660)
    const int D = std::min(o.maxdepth, MAX_LEVELS - 1);
    st.n_levels = 0;
    for (int d = 1; d <= D; ++d) {
      cudaEvent_t a, b;
      CK(cudaEventCreate(&a));
      CK(cudaEventCreate(&b));
      lev.push_back(a);
      lev.push_back(b);
      CK(cudaEventRecord(a, s));
      const bool any = run_level(d);
      st.n_levels = d;
      if (d < D) {
        select_next(d);
      }
      CK(cudaEventRecord(b, s));
      // the frontier blob of level d is no longer needed once level d+1 is materialised (ids are kept)
      if (!(o.flags & NW_FLAG_VERIFY_DEDUP)) lv[d].fblob.release();  // the verifier re-reads old frontiers
      if (d < D && (!any || lv[d + 1].P == 0)) {
        // solver.jl keeps looping over empty frontiers; nothing more can be generated
        for (int e = d + 1; e <= D; ++e) {
          if ((int)lv.size() <= e) lv.emplace_back();
          st.n_levels = e;
        }
        break;
      }
    }
    d2h(c, pin_u32(), misc, 4);
    CK(cudaEventRecord(ev1, s));
    CK(spin_sync(s));
    CK(cudaGetLastError());
    best_k3 = (int)pin_u32()[0];
    for (size_t k = 0; k + 1 < lev.size(); k += 2) {
      float ms = 0;
      cudaEventElapsedTime(&ms, lev[k], lev[k + 1]);
      st.level_ms[k / 2] = ms;
    }
    cudaEventElapsedTime(&st.total_ms, ev0, ev1);
    st.n_ids = n_ids;
    st.best = k3_to_score(best_k3);
    st.kernel_launches = c->launches;
  }

  // =================================================================================================
  // frontier-sharded search: the level loop of run() cut at the point where ranks exchange records
  // =================================================================================================
  void dist_begin() {
    cudaStream_t s = c->stream;
    CK(cudaEventCreate(&ev0));
    CK(cudaEventCreate(&ev1));
    CK(cudaEventRecord(ev0, s));
    init_root();
    int32_t* misc = (int32_t*)c->misc.p;
    CK(cudaMemsetAsync(misc, 0, 4, s));
    cur_level = 1;
    dist_more = o.maxdepth >= 1;
    st.n_levels = 0;
  }

  // expand this rank's share of level cur_level; returns the number of records it contributes to the exchange
  int64_t dist_expand() {
    cudaStream_t s = c->stream;
    const int d = cur_level;
    Level& L = lv[d];
    int32_t* misc = (int32_t*)c->misc.p;
    st.frontier[d - 1] = L.P;
    st.n_levels = d;
    L.G = L.G_all = L.n_local_new = L.n_new = 0;
    L.gbase = gtotal;
    L.id_base = n_ids + 1;
    curE = ExpandArgs{};
    if (L.P == 0) return 0;
    ExpandArgs E{};
    E.R = c->RD;
    E.F = L.frontier();
    E.cnt = (uint32_t*)c->cnt.ensure((size_t)L.P * 4, true);
    E.maxlen = misc + 2;
    E.own_mod = 1;
    E.force_pairs = (o.flags & NW_FLAG_PAIR_EXPAND) ? 1 : 0;
    CK(cudaMemsetAsync(misc + 2, 0, 4, s));
    c->launches += launch_expand(false, E, s);  // every rank counts every parent: global offsets need no collective
    uint32_t* off_glob = (uint32_t*)L.off_glob.ensure((size_t)L.P * 4);
    uint32_t* off_loc = (uint32_t*)L.off_loc.ensure((size_t)L.P * 4);
    uint32_t* own_cnt = (uint32_t*)c->off.ensure((size_t)L.P * 4, true);
    uint32_t* bsum = (uint32_t*)c->bsum.ensure(scan_scratch_items((uint64_t)L.P) * 4, true);
    c->launches += launch_scan(E.cnt, off_glob, (uint64_t)L.P, bsum, (uint32_t*)(misc + 4), s);
    c->launches += launch_mask_counts(E.cnt, (uint32_t)L.P, world, rank, own_cnt, s);
    c->launches += launch_scan(own_cnt, off_loc, (uint64_t)L.P, bsum, (uint32_t*)(misc + 7), s);
    d2h(c, pin_u32(), misc, 32);
    CK(spin_sync(s));
    CK(cudaGetLastError());
    const uint32_t G_all = pin_u32()[4], G = pin_u32()[7];
    const int maxlen = (int)pin_u32()[2];
    dist_maxlen = maxlen;
    if (maxlen > LMAX) fail(NW_ERR_RANGE, "a child index exceeds the supported length");
    if ((uint64_t)gtotal + G_all >= 0xfffffff0ull || G_all >= 0x7ffffff0u) fail(NW_ERR_NOMEM, "more than 2^32 candidates");
    L.G = G;
    L.G_all = G_all;
    st.generated[d - 1] = G_all;
    if (G_all == 0) return 0;
    table_reserve(c, (uint64_t)n_ids + G_all);  // room for every rank's records: no rehash during the merge
    E.off = off_loc;
    E.off_glob = off_glob;
    E.own_mod = world;
    E.own_rem = rank;
    E.desc = (uint64_t*)L.desc.ensure((size_t)std::max<uint32_t>(G, 1) * 8);
    E.slot = (uint32_t*)c->slot.ensure((size_t)std::max<uint32_t>(G, 1) * 4, true);
    E.table = c->table.as<uint64_t>();
    E.tmask = c->table_cap - 1;
    E.gbase = L.gbase;
    E.pw = c->d_pow.as<U4>();
    curE = E;
    if (G == 0) return 0;
    c->launches += launch_expand(true, E, s);
    c->launches += launch_hash_insert(E, G, s);
    uint32_t* is_new = (uint32_t*)c->is_new.ensure((size_t)G * 4, true);
    uint32_t* owner = (uint32_t*)c->owner.ensure((size_t)G * 4, true);
    c->launches += launch_resolve(E, G, is_new, owner, s);
    uint32_t* newrank = (uint32_t*)c->newrank.ensure((size_t)G * 4, true);
    bsum = (uint32_t*)c->bsum.ensure(scan_scratch_items(G) * 4, true);
    c->launches += launch_scan(is_new, newrank, G, bsum, (uint32_t*)(misc + 5), s);
    d2h(c, pin_u32(), misc, 32);
    CK(spin_sync(s));
    CK(cudaGetLastError());
    const uint32_t n_loc = pin_u32()[5];
    L.n_local_new = n_loc;
    AssignArgs A{};
    A.LT.n_levels = 0;
    A.slot = E.slot;
    A.table = E.table;
    A.owner = owner;
    A.is_new = is_new;
    A.newrank = newrank;
    A.desc = E.desc;
    A.front_len = L.flen.as<int32_t>();
    A.gbase = L.gbase;
    A.G = G;
    A.id_base = 0;
    A.n_y = c->RD.n_y;
    A.force = 0;
    A.skip_ids = 1;
    A.cand_id = nullptr;
    A.newlist = (uint32_t*)L.newlist.ensure((size_t)std::max<uint32_t>(n_loc, 1) * 4);
    A.new_len = (uint16_t*)c->new_len.ensure((size_t)std::max<uint32_t>(n_loc, 1) * 2, true);
    A.k3 = (int32_t*)L.l_k3.ensure((size_t)std::max<uint32_t>(n_loc, 1) * 4);
    A.cost = (int32_t*)L.l_cost.ensure((size_t)std::max<uint32_t>(n_loc, 1) * 4);
    A.total = (int32_t*)L.l_total.ensure((size_t)std::max<uint32_t>(n_loc, 1) * 4);
    const int lmax = std::min(std::max(maxlen, L.LS), LMAX);
    const size_t nbins = (size_t)(lmax + 1) * NSUB;
    A.hist = (uint32_t*)c->hist.ensure(HIST_N * 4);
    CK(cudaMemsetAsync(A.hist, 0, nbins * 4, s));
    c->launches += launch_assign(A, s);
    uint32_t* hist_h = pin_u32() + 64;
    d2h(c, hist_h, A.hist, nbins * 4);
    CK(spin_sync(s));
    CK(cudaGetLastError());
    int64_t cells = 0;
    ScoreView V{L.frontier(), n_loc, L.newlist.as<uint32_t>(), L.desc.as<uint64_t>(), L.l_k3.as<int32_t>(),
                L.l_cost.as<int32_t>(), L.l_total.as<int32_t>()};
    run_scoring(c, V, hist_h, lmax, 0, o.flags, misc + 1 /*best is taken from the merged records*/, &cells);
    st.cells[d - 1] = cells;  // cells evaluated by THIS rank (includes candidates another rank discovered first)
    CK(spin_sync(s));
    return n_loc;
  }

  void dist_pack(void* out) {
    Level& L = lv[cur_level];
    if (!L.n_local_new) return;
    c->launches += launch_dist_pack(curE, L.newlist.as<uint32_t>(), L.n_local_new, L.l_k3.as<int32_t>(),
                                    L.l_cost.as<int32_t>(), L.l_total.as<int32_t>(), out, c->stream);
    CK(spin_sync(c->stream));
    CK(cudaGetLastError());
  }

  // all: world segments of `stride` records each, segment r holding counts[r] records.  Identical on every rank.
  void dist_merge(const void* all, const int64_t* counts, int64_t stride) {
    cudaStream_t s = c->stream;
    const int d = cur_level;
    const int D = std::min(o.maxdepth, MAX_LEVELS - 1);
    int32_t* misc = (int32_t*)c->misc.p;
    {
      Level& L = lv[d];
      const DistRec* rec = (const DistRec*)all;
      uint32_t n_new = 0;
      if (L.G_all) {
        uint32_t* rslot = (uint32_t*)c->rslot.ensure((size_t)std::max<int64_t>(world * stride, 1) * 4, true);
        for (int r = 0; r < world; ++r)
          c->launches += launch_dist_merge_insert(rec + (size_t)r * stride, (uint32_t)counts[r], c->table.as<uint64_t>(),
                                                  c->table_cap - 1, rslot + (size_t)r * stride, s);
        uint32_t* newflag = (uint32_t*)c->newflag.ensure((size_t)L.G_all * 4, true);
        uint32_t* grank = (uint32_t*)c->grank.ensure((size_t)L.G_all * 4, true);
        CK(cudaMemsetAsync(newflag, 0, (size_t)L.G_all * 4, s));
        for (int r = 0; r < world; ++r)
          c->launches += launch_dist_merge_flag(rec + (size_t)r * stride, (uint32_t)counts[r], rslot + (size_t)r * stride,
                                                c->table.as<uint64_t>(), L.gbase, newflag, s);
        uint32_t* bsum = (uint32_t*)c->bsum.ensure(scan_scratch_items(L.G_all) * 4, true);
        c->launches += launch_scan(newflag, grank, L.G_all, bsum, (uint32_t*)(misc + 5), s);
        d2h(c, pin_u32(), misc, 32);
        CK(spin_sync(s));
        n_new = pin_u32()[5];
        L.n_new = n_new;
        const size_t nn = std::max<uint32_t>(n_new, 1);
        L.k3.ensure(nn * 4);
        L.cost.ensure(nn * 4);
        L.total.ensure(nn * 4);
        L.desc_new.ensure(nn * 8);
        L.gid_at.ensure((size_t)L.G_all * 4);
        CK(cudaMemsetAsync(L.gid_at.p, 0, (size_t)L.G_all * 4, s));
        for (int r = 0; r < world; ++r)
          c->launches += launch_dist_scatter(rec + (size_t)r * stride, (uint32_t)counts[r], rslot + (size_t)r * stride,
                                             c->table.as<uint64_t>(), L.gbase, grank, L.id_base, L.k3.as<int32_t>(),
                                             L.cost.as<int32_t>(), L.total.as<int32_t>(), L.desc_new.as<uint64_t>(),
                                             L.gid_at.as<int32_t>(), misc + 0, s);
        if (L.G) {  // ids of the indices this rank's candidates produced (provenance edges)
          LevelTable LT{};
          LT.n_levels = d + 1;
          for (int l = 0; l <= d; ++l) {
            LT.gbase[l] = lv[l].gbase;
            LT.cand_id[l] = l == 0 ? lv[0].cand_id.as<int32_t>() : lv[l].gid_at.as<int32_t>();
          }
          LT.gbase[d + 1] = L.gbase + L.G_all;
          c->launches += launch_dist_cand_id(curE.slot, c->table.as<uint64_t>(), LT, L.G,
                                             (int32_t*)L.cand_id.ensure((size_t)L.G * 4), s);
        }
      }
      st.scored[d - 1] = n_new;
      gtotal += L.G_all;
      n_ids += (int32_t)n_new;
      // select_next reads the level's maxlen from the pinned copy of misc: restore what dist_expand saw
      CK(cudaMemcpyAsync(misc + 2, &dist_maxlen, 4, cudaMemcpyHostToDevice, s));
    }
    bool more = false;
    if (d < D) {
      select_next(d);
      more = lv[d].n_new > 0 && lv[d + 1].P > 0;
    }
    CK(spin_sync(s));
    CK(cudaGetLastError());
    lv[d].fblob.release();
    if (!more) {
      for (int e = d + 1; e <= D; ++e) {
        if ((int)lv.size() <= e) lv.emplace_back();
        st.n_levels = e;
      }
      d2h(c, pin_u32(), misc, 4);
      CK(cudaEventRecord(ev1, s));
      CK(spin_sync(s));
      best_k3 = (int)pin_u32()[0];
      cudaEventElapsedTime(&st.total_ms, ev0, ev1);
      st.n_ids = n_ids;
      st.best = k3_to_score(best_k3);
    }
    dist_more = more;
    cur_level = d + 1;
  }

  void dist_report_begin(uint8_t* needed) {
    int k3_star;
    int64_t quota;
    report_thresholds(dist_k3_min, k3_star, quota);
    extract_good(needed, dist_k3_min, k3_star, quota);
    CK(spin_sync(c->stream));
    CK(cudaGetLastError());
  }
  void dist_finish(const uint8_t* needed, const void* all, const int64_t* counts, int64_t stride, nw_result* R) {
    std::vector<EdgeRec> edges;
    for (int r = 0; r < world; ++r) {
      if (!counts[r]) continue;
      const size_t at = edges.size();
      edges.resize(at + (size_t)counts[r]);
      d2h(c, edges.data() + at, (const EdgeRec*)all + (size_t)r * stride, (size_t)counts[r] * sizeof(EdgeRec));
    }
    CK(spin_sync(c->stream));
    build_nodes(needed, edges);
    enumerate(R, dist_k3_min);
  }

  // ---- report: get_good_trajectories (solver.jl:727-745) ----
  struct Node {
    int32_t k3 = 0, cost = -1, total = 0;
    std::vector<EdgeRec> in;
  };
  std::unordered_map<int32_t, Node> nodes;

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
        uint32_t* tflag = (uint32_t*)c->is_new.ensure((size_t)L.n_new * 4, true);
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
  // phase 3: compact this rank's edges into needed ids (all levels) into c->edgebuf; returns their number
  uint32_t extract_edges(const uint8_t* needed) {
    cudaStream_t s = c->stream;
    int32_t* misc = (int32_t*)c->misc.p;
    const int nl = (int)lv.size();
    uint32_t cap = std::max<uint32_t>(edge_cap_hint, 1u << 16);
    for (;;) {  // one pass; repeated with a larger buffer only if the append counter overflowed the capacity
      EdgeRec* buf = (EdgeRec*)c->edgebuf.ensure((size_t)cap * sizeof(EdgeRec));
      CK(cudaMemsetAsync(misc + 8, 0, 4, s));
      for (int l = 1; l < nl; ++l) {
        const Level& L = lv[l];
        c->launches += launch_edge_append(L.cand_id.as<int32_t>(), L.desc.as<uint64_t>(), L.fid.as<int32_t>(), L.G, needed,
                                          l, o.maxdepth, L.gbase, world > 1 ? L.off_loc.as<uint32_t>() : nullptr,
                                          world > 1 ? L.off_glob.as<uint32_t>() : nullptr, buf, cap,
                                          (uint32_t*)(misc + 8), s);
      }
      d2h(c, pin_u32(), misc, 64);
      CK(spin_sync(s));
      CK(cudaGetLastError());
      const uint32_t n = pin_u32()[8];
      if (n <= cap) {
        edge_cap_hint = std::max(edge_cap_hint, n);
        return n;
      }
      cap = n + n / 8;
    }
  }
  // phase 4: host graph of the needed ids from the edge records of all ranks (sorted into discovery order)
  void build_nodes(const uint8_t* needed, std::vector<EdgeRec>& edges) {
    cudaStream_t s = c->stream;
    int32_t* misc = (int32_t*)c->misc.p;
    const int nl = (int)lv.size();
    std::sort(edges.begin(), edges.end(), [](const EdgeRec& a, const EdgeRec& b) { return a.gpos < b.gpos; });
    nodes.clear();
    nodes[1] = Node{root_k3, root_cost, root_total, {}};
    for (const EdgeRec& e : edges) nodes[e.child].in.push_back(e);
    uint32_t cap = std::max<uint32_t>(id_cap_hint, 1u << 14);
    std::vector<IdRec> hi;
    for (;;) {
      IdRec* buf = (IdRec*)c->keys.ensure((size_t)cap * sizeof(IdRec));
      CK(cudaMemsetAsync(misc + 9, 0, 4, s));
      for (int l = 1; l < nl; ++l) {
        const Level& L = lv[l];
        c->launches += launch_id_append(L.k3.as<int32_t>(), L.cost.as<int32_t>(), L.total.as<int32_t>(), L.id_base, L.n_new,
                                        needed, buf, cap, (uint32_t*)(misc + 9), s);
      }
      d2h(c, pin_u32(), misc, 64);
      CK(spin_sync(s));
      const uint32_t n = pin_u32()[9];
      if (n <= cap) {
        id_cap_hint = std::max(id_cap_hint, n);
        hi.resize(n);
        if (n) {
          d2h(c, hi.data(), buf, (size_t)n * sizeof(IdRec));
          CK(spin_sync(s));
        }
        break;
      }
      cap = n + n / 8;
    }
    for (const IdRec& r : hi) {
      Node& nd = nodes[r.id];
      nd.k3 = r.k3;
      nd.cost = r.cost;
      nd.total = r.total;
    }
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
    std::unordered_map<int64_t, char> memo;  // (id, r) -> can the root be reached from id in exactly r moves
    std::vector<EdgeRec> stack;              // edges from the target backwards
    int64_t limit = -1, emitted = 0;
    nw_result* R = nullptr;
    int cur_id = 0, cur_k3 = 0;
    explicit Enumerator(Search& s) : S(s) {}
    bool reach(int32_t id, int r) {
      if (id == 1) return r == 0;
      if (r <= 0) return false;
      const int64_t key = ((int64_t)id << 8) | r;
      auto it = memo.find(key);
      if (it != memo.end()) return it->second;
      bool ok = false;
      auto nd = S.nodes.find(id);
      if (nd != S.nodes.end())
        for (const EdgeRec& e : nd->second.in)
          if (reach(e.parent, r - 1)) {
            ok = true;
            break;
          }
      memo.emplace(key, ok);
      return ok;
    }
    bool full() const { return limit >= 0 && emitted >= limit; }
    void emit() {
      static const char* names[3] = {"move_dup", "move_del", "move_inv"};
      const Node& nd = S.nodes[cur_id];
      R->t_score.push_back(score_f32(cur_k3));
      R->t_k3.push_back(cur_k3);
      R->t_id.push_back(cur_id);
      R->t_cost.push_back(nd.cost < 0 ? nd.cost : (int64_t)nd.cost * S.c->RH.gcd);
      R->t_total.push_back((int64_t)nd.total * S.c->RH.gcd);
      std::string& line = R->tsv;
      line += score_string(cur_k3);
      line += '\t';
      line += std::to_string(stack.size());
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
        line += std::to_string(i);
        line += '\t';
        line += std::to_string(j);  // :719
      }
      line += '\n';
      R->t_off.push_back((int64_t)R->t_moves.size());
      ++emitted;
    }
    // all paths of exactly r more moves from `id` back to the root, in the reference's enumeration order
    void walk(int32_t id, int r) {
      if (id == 1) {
        if (r == 0) emit();
        return;
      }
      auto nd = S.nodes.find(id);
      if (nd == S.nodes.end()) return;
      for (const EdgeRec& e : nd->second.in) {
        if (full()) return;
        if (!reach(e.parent, r - 1)) continue;
        stack.push_back(e);
        walk(e.parent, r - 1);
        stack.pop_back();
      }
    }
  };

  // thresholds of the report (identical on every rank): k3_min from minreport, k3_star/quota from maxreport
  void report_thresholds(int& k3_min, int& k3_star, int64_t& quota) {
    const double best = k3_to_score(best_k3);
    const double thr = o.minreport * best;  // solver.jl:733
    // smallest k3 whose Float32 score passes `score >= suboptimality * best` (monotone in k3)
    k3_min = 100001;
    {
      int lo = 0, hi = 100001;  // first k3 in [0, 100000] with (double)(float)(k3/1000) >= thr
      while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if ((double)score_f32(mid) >= thr) hi = mid; else lo = mid + 1;
      }
      k3_min = lo;
    }
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
      if (root_k3 >= 0 && root_k3 >= k3_min) hh[root_k3] += 1;
      int64_t cum = 0;
      for (int k = 100000; k >= 0; --k) {
        if (cum + hh[k] >= o.maxreport) {
          k3_star = k;
          quota = o.maxreport - cum;  // ids to take from the tie bin, in id order
          break;
        }
        cum += hh[k];
      }
      if (k3_star >= 0 && root_k3 == k3_star) quota -= 1;  // the root (id 1) is the first of its tie bin
    }
  }
  void report(nw_result* R) {
    int k3_min, k3_star;
    int64_t quota;
    report_thresholds(k3_min, k3_star, quota);
    extract(k3_min, k3_star, quota);
    enumerate(R, k3_min);
  }
  void enumerate(nw_result* R, int k3_min) {
    std::vector<std::pair<int32_t, int32_t>> good;  // (-k3, id): canonical Dict order = ascending discovery id
    for (auto& kv : nodes) {
      const int k3 = kv.second.k3;
      if (k3 >= 0 && k3 >= k3_min) good.emplace_back(-k3, kv.first);
    }
    std::sort(good.begin(), good.end());
    Enumerator E(*this);
    E.limit = o.maxreport;
    E.R = R;
    R->t_off.push_back(0);
    size_t g0 = 0;
    while (g0 < good.size() && !E.full()) {
      size_t g1 = g0;
      while (g1 < good.size() && good[g1].first == good[g0].first) ++g1;
      for (int m = 0; m <= o.maxdepth + 1 && !E.full(); ++m)
        for (size_t g = g0; g < g1 && !E.full(); ++g) {
          E.cur_id = good[g].second;
          E.cur_k3 = -good[g].first;
          if (E.reach(E.cur_id, m)) E.walk(E.cur_id, m);
        }
      g0 = g1;
    }
  }

  void dump_ids(nw_result* R) {
    cudaStream_t s = c->stream;
    R->id_score.assign(n_ids, 0.f);
    R->id_cost.assign(n_ids, -1);
    R->id_total.assign(n_ids, 0);
    R->id_level.assign(n_ids, 0);
    auto put = [&](int32_t id, int k3, int cost, int total, int level) {
      R->id_score[id - 1] = score_f32(k3);
      R->id_cost[id - 1] = cost < 0 ? cost : (int64_t)cost * c->RH.gcd;
      R->id_total[id - 1] = (int64_t)total * c->RH.gcd;
      R->id_level[id - 1] = level;
    };
    put(1, root_k3, root_cost, root_total, 0);
    for (int l = 1; l < (int)lv.size(); ++l) {
      const Level& L = lv[l];
      if (L.n_new) {
        std::vector<int32_t> k3(L.n_new), cost(L.n_new), total(L.n_new);
        d2h(c, k3.data(), L.k3.p, (size_t)L.n_new * 4);
        d2h(c, cost.data(), L.cost.p, (size_t)L.n_new * 4);
        d2h(c, total.data(), L.total.p, (size_t)L.n_new * 4);
        CK(spin_sync(s));
        for (uint32_t r = 0; r < L.n_new; ++r) put(L.id_base + (int32_t)r, k3[r], cost[r], total[r], l);
      }
      if (L.G) {
        std::vector<int32_t> cid(L.G), fid(L.P);
        std::vector<uint64_t> desc(L.G);
        d2h(c, cid.data(), L.cand_id.p, (size_t)L.G * 4);
        d2h(c, desc.data(), L.desc.p, (size_t)L.G * 8);
        d2h(c, fid.data(), L.fid.p, (size_t)L.P * 4);
        CK(spin_sync(s));
        for (uint32_t g = 0; g < L.G; ++g) {
          R->e_child.push_back(cid[g]);
          R->e_parent.push_back(fid[desc_parent(desc[g])]);
          R->e_moves.push_back(desc_kind(desc[g]));
          R->e_moves.push_back(desc_i(desc[g]));
          R->e_moves.push_back(desc_j(desc[g]));
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

void do_solve(nw_ctx* c, const int8_t* aln, int n_x, int n_y, const int64_t* len_x, const int64_t* len_y,
              const nw_opts* o, nw_result* R) {
  CK(cudaSetDevice(c->device));
  c->launches = 0;
  c->h2d_bytes = c->d2h_bytes = 0;
  collect_score_time(c, nullptr, nullptr);
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<float, std::milli>(b - a).count();
  };
  const auto t0 = now();
  setup_region(c, aln, n_x, n_y, len_x, len_y, o->maxdepth, false);
  Search S(c, *o);
  const auto t1 = now();
  S.run();
  const auto t2 = now();
  collect_score_time(c, &S.st.score_ms, &S.st.score_launches);
  if (!(o->flags & NW_FLAG_NO_TRAJ)) S.report(R);
  const auto t3 = now();
  if (o->keep_ids) S.dump_ids(R);
  const auto t4 = now();
  S.st.host_ms[0] = ms(t0, t1);
  S.st.host_ms[1] = ms(t1, t2);
  S.st.host_ms[2] = ms(t2, t3);
  S.st.host_ms[3] = ms(t3, t4);
  collect_score_time(c, nullptr, nullptr);
  S.st.kernel_launches = c->launches;
  S.st.h2d_bytes = c->h2d_bytes;
  S.st.d2h_bytes = c->d2h_bytes;
  R->stats = S.st;
}

}  // namespace

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
    c->device = device;
    PoolScope pool_scope(c);
    try {
      CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
      CK(cudaMallocHost(&c->pinned, PINNED_BYTES));
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
  cudaSetDevice(c->device);
  nw_dist_abort(c);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->pinned) cudaFreeHost(c->pinned);
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
    const std::string tmp = std::string(outp) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) fail(NW_ERR_IO, std::string("cannot write ") + tmp);
    const size_t w = fwrite(R.tsv.data(), 1, R.tsv.size(), f);
    const int cl = fclose(f);
    if (w != R.tsv.size() || cl != 0) {
      remove(tmp.c_str());
      fail(NW_ERR_IO, std::string("short write to ") + tmp);
    }
    if (rename(tmp.c_str(), outp) != 0) {
      remove(tmp.c_str());
      fail(NW_ERR_IO, std::string("cannot rename to ") + outp);
    }
  });
  if (rc != NW_OK) cudaGetLastError();
  if (rc == NW_OK && stats) *stats = R.stats;
  return rc;
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

int nw_score_batch(nw_ctx* c, const int8_t* aln, int32_t n_x, int32_t n_y, const int64_t* len_x, const int64_t* len_y,
                   const int16_t* flat, const int64_t* off, int64_t n, int32_t force, int32_t flags, int64_t* cost_bp,
                   int64_t* total_bp, double* score) {
  if (!c || !aln || !len_x || !len_y || !flat || !off || n < 0) {
    g_last_error = "null argument";
    return NW_ERR_ARG;
  }
  PoolScope pool_scope(c);
  const int rc = guard([&] {
    CK(cudaSetDevice(c->device));
    if (n == 0) return;
    if (n >= (1ll << 31)) fail(NW_ERR_ARG, "too many candidates");
    cudaStream_t s = c->stream;
    int maxlen = 1;
    for (int64_t k = 0; k < n; ++k) {
      const int64_t l = off[k + 1] - off[k];
      if (l < 1 || l > LMAX) fail(NW_ERR_ARG, "candidate length out of range");
      maxlen = std::max<int>(maxlen, (int)l);
      for (int64_t t = off[k]; t < off[k + 1]; ++t) {
        const int v = flat[t];
        if (v == 0 || v > n_x || v < -n_x) fail(NW_ERR_ARG, "candidate element out of range");
      }
    }
    // depth bound for the int32 range: candidates may hold each block up to maxlen/n_x times
    int depth_equiv = 0;
    while (((int64_t)n_x << depth_equiv) < maxlen) ++depth_equiv;
    setup_region(c, aln, n_x, n_y, len_x, len_y, depth_equiv, force != 0);
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
    std::vector<uint32_t> hist(HIST_N, 0);
    std::vector<int32_t> k3(n, 0), cost(n, -1), total(n, 0);
    for (int64_t k = 0; k < n; ++k) {
      const int l = (int)(off[k + 1] - off[k]);
      memcpy(&seqs[(size_t)k * L.LS], flat + off[k], (size_t)l * 2);
      lens[k] = l;
      desc[k] = pack_desc((uint32_t)k, 0, 0, KIND_ID);
      newlist[k] = (uint32_t)k;
      nlen[k] = (uint16_t)l;
      if (force || !early_exit(l, n_y)) hist[bucket_key(l, (uint32_t)k)]++;
    }
    L.G = (uint32_t)n;
    L.n_new = (uint32_t)n;
    h2d(c, L.fblob.ensure(seqs.size() * 2), seqs.data(), seqs.size() * 2);
    h2d(c, L.flen.ensure(n * 4), lens.data(), n * 4);
    h2d(c, L.desc.ensure(n * 8), desc.data(), n * 8);
    h2d(c, L.newlist.ensure(n * 4), newlist.data(), n * 4);
    h2d(c, c->new_len.ensure(n * 2, true), nlen.data(), n * 2);
    h2d(c, L.k3.ensure(n * 4), k3.data(), n * 4);
    h2d(c, L.cost.ensure(n * 4), cost.data(), n * 4);
    h2d(c, L.total.ensure(n * 4), total.data(), n * 4);
    int32_t* misc = (int32_t*)c->misc.ensure(256);
    CK(cudaMemsetAsync(misc, 0, 256, s));
    CK(spin_sync(s));
    run_scoring(c, view_of(L), hist.data(), maxlen, force, flags, misc + 1, nullptr);
    d2h(c, k3.data(), L.k3.p, n * 4);
    d2h(c, cost.data(), L.cost.p, n * 4);
    d2h(c, total.data(), L.total.p, n * 4);
    CK(spin_sync(s));
    CK(cudaGetLastError());
    for (int64_t k = 0; k < n; ++k) {
      if (cost_bp) cost_bp[k] = cost[k] < 0 ? cost[k] : (int64_t)cost[k] * c->RH.gcd;
      if (total_bp) total_bp[k] = (int64_t)total[k] * c->RH.gcd;
      if (score) score[k] = k3[k] < 0 ? -INFINITY : k3_to_score(k3[k]);
    }
  });
  if (rc != NW_OK) cudaGetLastError();
  return rc;
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
    setup_region(c, aln, n_x, n_y, lx.data(), ly.data(), 1, false);
    const int MW = c->RH.MW;
    std::vector<uint32_t> dd((size_t)(n_x + 1) * MW), iv((size_t)(n_x + 1) * MW);
    d2h(c, dd.data(), c->d_msdd.p, dd.size() * 4);
    d2h(c, iv.data(), c->d_msinv.p, iv.size() * 4);
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
// frontier-sharded single search: see include/nwsolve.h for the protocol
// ---------------------------------------------------------------------------------------------------
static Search* dist_of(nw_ctx* c) {
  if (!c || !c->dist) fail(NW_ERR_ARG, "no sharded search in progress (call nw_dist_begin first)");
  return static_cast<Search*>(c->dist);
}
void nw_dist_abort(nw_ctx* c) {
  if (!c || !c->dist) return;
  PoolScope pool_scope(c);
  cudaSetDevice(c->device);
  delete static_cast<Search*>(c->dist);
  c->dist = nullptr;
}
int nw_dist_begin(nw_ctx* c, const int8_t* aln, int32_t n_x, int32_t n_y, const int64_t* len_x, const int64_t* len_y,
                  const nw_opts* o, int32_t rank, int32_t world) {
  if (!c || !aln || !len_x || !len_y) {
    g_last_error = "null argument";
    return NW_ERR_ARG;
  }
  nw_dist_abort(c);
  PoolScope pool_scope(c);
  const int rc = guard([&] {
    check_opts(o);
    if (world < 1 || rank < 0 || rank >= world) fail(NW_ERR_ARG, "bad rank/world");
    CK(cudaSetDevice(c->device));
    c->launches = 0;
    c->h2d_bytes = c->d2h_bytes = 0;
    collect_score_time(c, nullptr, nullptr);
    setup_region(c, aln, n_x, n_y, len_x, len_y, o->maxdepth, false);
    Search* S = new Search(c, *o);
    c->dist = S;
    S->rank = rank;
    S->world = world;
    S->dist_begin();
  });
  if (rc != NW_OK) {
    cudaGetLastError();
    nw_dist_abort(c);
  }
  return rc;
}
int nw_dist_expand(nw_ctx* c, int64_t* n_records) {
  PoolScope pool_scope(c);
  const int rc = guard([&] {
    Search* S = dist_of(c);
    CK(cudaSetDevice(c->device));
    if (!S->dist_more) fail(NW_ERR_ARG, "the sharded search has no further level");
    const int64_t n = S->dist_expand();
    if (n_records) *n_records = n;
  });
  if (rc != NW_OK) cudaGetLastError();
  return rc;
}
int nw_dist_pack(nw_ctx* c, void* dev_out) {
  PoolScope pool_scope(c);
  const int rc = guard([&] {
    Search* S = dist_of(c);
    CK(cudaSetDevice(c->device));
    S->dist_pack(dev_out);
  });
  if (rc != NW_OK) cudaGetLastError();
  return rc;
}
int nw_dist_merge(nw_ctx* c, const void* dev_all, const int64_t* counts, int64_t stride_records, int32_t* more) {
  PoolScope pool_scope(c);
  const int rc = guard([&] {
    Search* S = dist_of(c);
    CK(cudaSetDevice(c->device));
    if (!counts) fail(NW_ERR_ARG, "counts is null");
    S->dist_merge(dev_all, counts, stride_records);
    if (more) *more = S->dist_more ? 1 : 0;
  });
  if (rc != NW_OK) cudaGetLastError();
  return rc;
}
int64_t nw_dist_n_ids(nw_ctx* c) { return (c && c->dist) ? (int64_t) static_cast<Search*>(c->dist)->n_ids : -1; }
int nw_dist_report_begin(nw_ctx* c, void* needed_dev, int64_t nbytes) {
  PoolScope pool_scope(c);
  const int rc = guard([&] {
    Search* S = dist_of(c);
    CK(cudaSetDevice(c->device));
    if (!needed_dev || nbytes < (int64_t)S->n_ids + 16) fail(NW_ERR_ARG, "needed buffer must hold n_ids + 16 bytes");
    collect_score_time(c, &S->st.score_ms, &S->st.score_launches);
    S->dist_report_begin((uint8_t*)needed_dev);
  });
  if (rc != NW_OK) cudaGetLastError();
  return rc;
}
int nw_dist_report_mark(nw_ctx* c, void* needed_dev, int32_t pass) {
  PoolScope pool_scope(c);
  const int rc = guard([&] {
    Search* S = dist_of(c);
    CK(cudaSetDevice(c->device));
    S->extract_mark((uint8_t*)needed_dev, pass);
    CK(spin_sync(c->stream));
    CK(cudaGetLastError());
  });
  if (rc != NW_OK) cudaGetLastError();
  return rc;
}
int nw_dist_report_edges(nw_ctx* c, const void* needed_dev, int64_t* n_edges) {
  PoolScope pool_scope(c);
  const int rc = guard([&] {
    Search* S = dist_of(c);
    CK(cudaSetDevice(c->device));
    const uint32_t n = S->extract_edges((const uint8_t*)needed_dev);
    S->dist_edges = n;
    if (n_edges) *n_edges = n;
  });
  if (rc != NW_OK) cudaGetLastError();
  return rc;
}
int nw_dist_report_pack(nw_ctx* c, void* dev_out) {
  PoolScope pool_scope(c);
  const int rc = guard([&] {
    Search* S = dist_of(c);
    CK(cudaSetDevice(c->device));
    if (S->dist_edges) {
      CK(cudaMemcpyAsync(dev_out, c->edgebuf.p, (size_t)S->dist_edges * sizeof(EdgeRec), cudaMemcpyDeviceToDevice,
                         c->stream));
      CK(spin_sync(c->stream));
    }
  });
  if (rc != NW_OK) cudaGetLastError();
  return rc;
}
int nw_dist_finish(nw_ctx* c, const void* needed_dev, const void* dev_edges_all, const int64_t* counts,
                   int64_t stride_edges, nw_result** out) {
  if (!out) return NW_ERR_ARG;
  *out = nullptr;
  PoolScope pool_scope(c);
  nw_result* R = new nw_result();
  const int rc = guard([&] {
    Search* S = dist_of(c);
    CK(cudaSetDevice(c->device));
    if (!counts) fail(NW_ERR_ARG, "counts is null");
    S->dist_finish((const uint8_t*)needed_dev, dev_edges_all, counts, stride_edges, R);
    S->st.kernel_launches = c->launches;
    S->st.h2d_bytes = c->h2d_bytes;
    S->st.d2h_bytes = c->d2h_bytes;
    R->stats = S->st;
  });
  nw_dist_abort(c);
  if (rc != NW_OK) {
    delete R;
    cudaGetLastError();
    return rc;
  }
  *out = R;
  return NW_OK;
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
