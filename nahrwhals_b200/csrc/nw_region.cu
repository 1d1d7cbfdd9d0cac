// nw_region.cu -- the per-region steps either side of the solver call (include/nwregion.h).
//
// Reference control flow: R/wrapper_aln_and_analyse.R:176-191.  The reference manipulates the gridmatrix itself
// (cbind of column ranges, R/seqmodder.R); here an arrangement is a list of positions (signed x block, length of the
// position), the matrix is never materialised, and all children of the reference arrangement are evaluated in two
// launches: k_est_diag (estimate pre-filter, R/evaltools.R:51-71) and the banded-DP kernels behind nw_score_batch
// (corridor chosen the way the R twin chooses it, R/evaltools.R:193-205).  The sequential part of dfsutil
// (visited-hash, running est_highest, > 90 filter) is then replayed on the host over the per-child numbers in the
// reference's order.  No CPU fallback: without a device nw_region_prepass / nw_region_analyse fail.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/file.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "nw_host.h"
#include "../../include/nwregion.h"

struct nw_ctx;
namespace nw {
void set_last_error(const std::string& m);
int ctx_device(nw_ctx* c);
cudaStream_t ctx_stream(nw_ctx* c);
int ctx_launches(nw_ctx* c);
cudaError_t ctx_region_staging(nw_ctx* c, size_t bytes, void** host_pinned, void** dev);
struct ScoreProblem {
  const int8_t* aln;  // n_x*n_y, x-major
  int n_x, n_y;
  const int64_t* len_x;
  const int64_t* len_y;
};
int score_batch_regions(nw_ctx* c, const ScoreProblem* probs, int R, const uint16_t* rid, const int16_t* flat,
                        const int64_t* off, int64_t n, int force, int flags, int64_t* cost_bp, int64_t* total_bp,
                        double* score, int depth_floor, int reuse_regions);
}  // namespace nw

#include "nw_region_host.h"

namespace {

#define RCK(call)                                                                         \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) rfail(NW_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)
#define NWCK(call)                                       \
  do {                                                   \
    int rc_ = (call);                                    \
    if (rc_ != NW_OK) rfail(rc_, nw_last_error());       \
  } while (0)

// ---------------------------------------------------------------------------------------------------------------
// GPU: estimate pre-filter of every child (calculate_estimated_aln_score, R/evaltools.R:51-71).
//   matdiag = mat * (|row - col| < ncol * fact);  sum(rowMaxs(matdiag)) + sum(colMaxs(matdiag))
// One warp per entry (child or reference arrangement) of any region of the batch.  The child matrix is read through its position list: mat[y][p] = sign(blk_p) * grid[y][|blk_p|].
// fact = 1 when n_y < 10 (|d| < L), else 0.25 (|d| < L/4  <=>  4|d| < L, exact in integers).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_est_diag(const int64_t* __restrict__ grid_all, const int4* __restrict__ meta,
                                                  const int16_t* __restrict__ blk, const int64_t* __restrict__ off,
                                                  int n, long long* __restrict__ out /* [n][2]: rows, cols */) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  const int4 mt = meta[w];  // grid offset (lo, hi), n_x, n_y of the entry's region
  const int64_t* grid = grid_all + (((long long)(unsigned)mt.x) | ((long long)mt.y << 32));
  const int n_x = mt.z, n_y = mt.w;
  const int16_t* b = blk + off[w];
  const int L = (int)(off[w + 1] - off[w]);
  const int fm = n_y < 10 ? 1 : 4;  // diagonalize_fact 1 (dim(mat)[1] < 10) or 0.25
  auto inside = [&](int y, int p) { return fm * abs(y - p) < L; };
  const long long NEG = LLONG_MIN;
  long long sr = 0, sc = 0;
  for (int y = lane; y < n_y; y += 32) {  // rowMaxs: the row also holds the zeros of the cells outside the band
    long long m = NEG;
    bool any_out = false;
    for (int p = 0; p < L; ++p) {
      if (inside(y, p)) {
        const int v = b[p];
        const long long g = grid[(size_t)y * n_x + (abs(v) - 1)];
        m = max(m, v < 0 ? -g : g);
      } else {
        any_out = true;
      }
    }
    if (any_out || m == NEG) m = max(m, 0ll);
    sr += m;
  }
  for (int p = lane; p < L; p += 32) {  // colMaxs
    const int v = b[p];
    const int x = abs(v) - 1;
    long long m = NEG;
    bool any_out = false;
    for (int y = 0; y < n_y; ++y) {
      if (inside(y, p)) {
        const long long g = grid[(size_t)y * n_x + x];
        m = max(m, v < 0 ? -g : g);
      } else {
        any_out = true;
      }
    }
    if (any_out || m == NEG) m = max(m, 0ll);
    sc += m;
  }
  for (int o = 16; o; o >>= 1) {
    sr += __shfl_xor_sync(0xffffffffu, sr, o);
    sc += __shfl_xor_sync(0xffffffffu, sc, o);
  }
  if (lane == 0) {
    out[(size_t)w * 2] = sr;
    out[(size_t)w * 2 + 1] = sc;
  }
}

// the same table straight from a result handle: the trajectories are taken from the result's arrays instead of
// re-parsing its text.  A reported score is a 3-decimal number printed from Float32 (solver.jl:714, :719), so the
// double R reads from the text is round(f * 1000) / 1000.
nw_table* format_solver_result(const nw_result* res, const double* gl, int n_gl) {
  const int64_t n = nw_result_count(res), nm = nw_result_move_count(res);
  std::vector<float> sc((size_t)n);
  std::vector<int32_t> nmv((size_t)n), m3((size_t)nm * 3 + 3);
  std::vector<int64_t> off((size_t)n);
  NWCK(nw_result_arrays(res, sc.data(), nmv.data(), off.data(), nullptr, nullptr, nullptr, m3.data()));
  std::vector<TrajIn> tr((size_t)n);
  std::vector<Sv> mv((size_t)nm);
  static const int kmap[3] = {1, 0, 2};  // NW_MOVE_DUP, NW_MOVE_DEL, NW_MOVE_INV -> dup, del, inv
  for (int64_t k = 0; k < nm; ++k) mv[k] = Sv{(double)m3[k * 3 + 1], (double)m3[k * 3 + 2], kmap[m3[k * 3]]};
  size_t ncols = 0;
  for (int64_t k = 0; k < n; ++k) {
    tr[k] = TrajIn{std::nearbyint((double)sc[k] * 1000.0) / 1000.0, nmv[k], (uint32_t)off[k], (uint32_t)nmv[k]};
    ncols = std::max<size_t>(ncols, nmv[k] == 0 ? 3 : 2 + 3 * (size_t)nmv[k]);
  }
  return table_from_trajectories(tr, mv, ncols, gl, n_gl);
}

// Phase 2 (GPU), for a whole batch of regions: ONE estimate launch over every entry of every region, one DP pass over
// all reference arrangements (forced corridor) and one over all children (R-twin corridor).
void prepass_gpu(nw_ctx* ctx, std::vector<PrepassOut*>& Ps, int flags, int& launches) {
  std::vector<PrepassOut*> live;
  for (PrepassOut* P : Ps)
    if (!P->cluttered && !P->trivial) live.push_back(P);
  if (live.empty()) return;
  if (live.size() > 60000) rfail(NW_ERR_ARG, "too many regions in one pre-pass batch");
  const int L0 = nw::ctx_launches(ctx);
  static const bool dbg = getenv("NW_DEBUG_TIMING") != nullptr;
  auto tnow = [] { return std::chrono::steady_clock::now(); };
  auto tms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
  };
  const auto tp0 = tnow();
  // ---- estimate sums: inputs are packed straight into the context's pinned staging buffer (one H2D copy) ----
  {
    size_t n_grid = 0, n_blk = 0, ne = 0;
    for (PrepassOut* P : live) {
      n_grid += P->flipped.v.size();
      n_blk += (size_t)P->flipped.nx;
      for (const Arr& a : P->cand) n_blk += a.size();
      ne += P->cand.size() + 1;
    }
    if (ne >= ((size_t)1 << 26)) rfail(NW_ERR_ARG, "too many pre-pass candidates in one batch");
    auto up16 = [](size_t v) { return (v + 15) / 16 * 16; };  // the int4 metadata needs 16-byte alignment
    const size_t o_grid = 0, o_off = o_grid + up16(n_grid * 8), o_meta = o_off + up16((ne + 1) * 8), o_blk = o_meta + ne * 16,
                 o_out = o_blk + up16(n_blk * 2), total = o_out + ne * 16;
    RCK(cudaSetDevice(nw::ctx_device(ctx)));
    cudaStream_t st = nw::ctx_stream(ctx);
    void *hp = nullptr, *dp = nullptr;
    RCK(nw::ctx_region_staging(ctx, total, &hp, &dp));
    unsigned char *h = static_cast<unsigned char*>(hp), *d = static_cast<unsigned char*>(dp);
    int64_t* h_grid = reinterpret_cast<int64_t*>(h + o_grid);
    int64_t* h_off = reinterpret_cast<int64_t*>(h + o_off);
    int32_t* h_meta = reinterpret_cast<int32_t*>(h + o_meta);
    int16_t* h_blk = reinterpret_cast<int16_t*>(h + o_blk);
    size_t go = 0, bo = 0, e = 0;
    h_off[0] = 0;
    for (PrepassOut* P : live) {
      const Grid& g = P->flipped;
      memcpy(h_grid + go, g.v.data(), g.v.size() * 8);
      for (size_t k = 0; k <= P->cand.size(); ++k, ++e) {
        if (k == 0) {
          for (int i = 0; i < g.nx; ++i) h_blk[bo++] = (int16_t)(i + 1);
        } else {
          for (const Pos& p : P->cand[k - 1]) h_blk[bo++] = p.blk;
        }
        h_off[e + 1] = (int64_t)bo;
        h_meta[e * 4 + 0] = (int32_t)(go & 0xffffffffull);
        h_meta[e * 4 + 1] = (int32_t)(go >> 32);
        h_meta[e * 4 + 2] = g.nx;
        h_meta[e * 4 + 3] = g.ny;
      }
      go += g.v.size();
    }
    RCK(cudaMemcpyAsync(d, h, o_out, cudaMemcpyHostToDevice, st));
    k_est_diag<<<(unsigned)((ne * 32 + 255) / 256), 256, 0, st>>>((const int64_t*)(d + o_grid), (const int4*)(d + o_meta),
                                                                   (const int16_t*)(d + o_blk), (const int64_t*)(d + o_off),
                                                                   (int)ne, (long long*)(d + o_out));
    RCK(cudaGetLastError());
    ++launches;
    RCK(cudaMemcpyAsync(h + o_out, d + o_out, ne * 16, cudaMemcpyDeviceToHost, st));
    RCK(cudaStreamSynchronize(st));
    const long long* out = reinterpret_cast<const long long*>(h + o_out);
    e = 0;
    for (PrepassOut* P : live)
      for (size_t k = 0; k <= P->cand.size(); ++k, ++e) {
        P->srow[k] = out[e * 2];
        P->scol[k] = out[e * 2 + 1];
      }
  }
  // ---- DP: pass 0 = reference arrangements, forcecalc (dfsutil: forcecalc = (pair$sv == "ref")); pass 1 = children
  //      with more than one column (a single remaining column has its own formula, R/evaltools.R:109-114) ----
  std::vector<nw::ScoreProblem> probs(live.size());
  for (size_t r = 0; r < live.size(); ++r) {
    PrepassOut* P = live[r];
    probs[r] = nw::ScoreProblem{P->aln.data(), P->nxa, P->flipped.ny, P->lenx.data(), P->flipped.rn.data()};
  }
  const auto tp1 = tnow();
  double tpass[2] = {0, 0};
  int did_pass0 = 0;  // pass 1 reuses the region contexts pass 0 built (a child holds a block at most twice: depth 1)
  for (int pass = 0; pass < 2; ++pass) {
    const auto tq = tnow();
    std::vector<int16_t> flat;
    std::vector<int64_t> off(1, 0);
    std::vector<uint16_t> rid;
    std::vector<std::pair<uint32_t, uint32_t>> who;
    for (size_t r = 0; r < live.size(); ++r) {
      PrepassOut* P = live[r];
      const size_t k0 = pass ? 1 : 0, k1 = pass ? P->cand.size() : 0;
      for (size_t k = k0; k <= k1; ++k) {
        if (P->seqs[k].size() < 2) continue;
        flat.insert(flat.end(), P->seqs[k].begin(), P->seqs[k].end());
        off.push_back((int64_t)flat.size());
        rid.push_back((uint16_t)r);
        who.emplace_back((uint32_t)r, (uint32_t)k);
      }
    }
    if (who.empty()) continue;
    std::vector<int64_t> c2(who.size()), t2(who.size());
    std::vector<double> s2(who.size());
    NWCK(nw::score_batch_regions(ctx, probs.data(), (int)live.size(), rid.data(), flat.data(), off.data(), (int64_t)who.size(),
                                 pass ? 2 : 1, flags & NW_FLAG_GENERIC_DP, c2.data(), t2.data(), s2.data(), 1, pass && did_pass0));
    if (pass == 0) did_pass0 = 1;
    for (size_t q = 0; q < who.size(); ++q) {
      PrepassOut* P = live[who[q].first];
      P->ccost[who[q].second] = c2[q];
      P->ctotal[who[q].second] = t2[q];
      P->cscore[who[q].second] = s2[q];
    }
    tpass[pass] = tms(tq, tnow());
  }
  if (dbg) fprintf(stderr, "[nw_region prepass gpu] estimate %.1f ms, forced ref DP %.1f ms, child DP %.1f ms\n", tms(tp0, tp1), tpass[0], tpass[1]);
  launches += nw::ctx_launches(ctx) - L0;
}

nw_opts solver_opts(const nw_region_opts& o, double width) {
  nw_opts so;
  nw_default_opts(&so);
  so.maxdepth = o.depth;
  so.maxdup = o.maxdup;
  so.maxwidth = (int64_t)width;  // the command line carries paste0(width): "100", "1000"
  so.minreport = o.minreport;
  so.maxreport = 1000;  // -R 1000 (R/juliasolver.R:48)
  so.flags = o.flags;
  return so;
}

// Regions are independent on the host side too: the prepare / replay / format phases of a batch run on a few threads.
thread_local int g_host_threads = 12;  // budget of the calling thread (divided among contexts by the _ctxs entry point)
void parallel_for(int64_t n, const std::function<void(int64_t)>& fn) {
  const unsigned hw = (unsigned)nw::host_cores();
  const int nt = (int)std::min<int64_t>(std::max<int64_t>(1, n / 8), std::min<unsigned>(hw, (unsigned)std::max(1, g_host_threads)));
  if (nt <= 1) {
    for (int64_t k = 0; k < n; ++k) fn(k);
    return;
  }
  std::atomic<int64_t> next{0};
  std::mutex mu;
  bool failed = false;
  RErr err{NW_OK, ""};
  auto work = [&] {
    for (;;) {
      const int64_t k = next.fetch_add(1);
      if (k >= n) return;
      try {
        fn(k);
      } catch (const RErr& e) {
        std::lock_guard<std::mutex> lk(mu);
        if (!failed) err = e;
        failed = true;
        next.store(n);
      } catch (const std::exception& e) {
        std::lock_guard<std::mutex> lk(mu);
        if (!failed) err = RErr{NW_ERR_ARG, e.what()};
        failed = true;
        next.store(n);
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  if (failed) throw err;
}

int rguard(const std::function<void()>& fn) {
  try {
    fn();
    return NW_OK;
  } catch (const RErr& e) {
    nw::set_last_error(e.msg);
    return e.code;
  } catch (const std::bad_alloc&) {
    nw::set_last_error("host out of memory");
    return NW_ERR_NOMEM;
  } catch (const std::exception& e) {
    nw::set_last_error(e.what());
    return NW_ERR_ARG;
  }
}

// The solver stage of n regions: pre-pass of all of them in one GPU batch, then the regions below 99.8 through the
// forest search (nw_solve_many; regions of one width setting per call), format + merge + summary per region.
void analyse_many(nw_ctx* ctx, const nw_region_problem* regs, int64_t n, const nw_region_opts& o, nw_table** outs,
                  nw_region_summary* summaries) {
  static const bool dbg = getenv("NW_DEBUG_TIMING") != nullptr;
  auto tnow = [] { return std::chrono::steady_clock::now(); };
  auto tms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
  };
  const auto t0 = tnow();
  std::vector<PrepassOut> Ps((size_t)n);
  std::vector<PrepassOut*> ptrs;
  parallel_for(n, [&](int64_t k) {
    const nw_region_problem& r = regs[k];
    if (r.n_gridlines > 0 && !r.gridlines_x) rfail(NW_ERR_ARG, "null gridlines");
    prepass_prepare(Ps[k], make_grid(r.grid, r.n_x, r.n_y, r.len_x, r.len_y));
  });
  for (int64_t k = 0; k < n; ++k) ptrs.push_back(&Ps[k]);
  int launches = 0;
  const auto t1 = tnow();
  prepass_gpu(ctx, ptrs, o.flags, launches);
  const auto t2 = tnow();
  std::vector<nw_region_summary> S((size_t)n);
  std::vector<int64_t> need[2];  // regions that go to the solver, by width setting (plain / reduced by the guard)
  std::vector<std::vector<int8_t>> alns((size_t)n);
  parallel_for(n, [&](int64_t k) { prepass_replay(Ps[k], o); });
  for (int64_t k = 0; k < n; ++k) {
    PrepassOut& P = Ps[k];
    P.launches = k == 0 ? launches : 0;  // the launches are shared by the batch: reported once
    memset(&S[k], 0, sizeof S[k]);
    fill_prepass_summary(S[k], P);
    if (needs_solver(P)) {  // wrapper :177
      S[k].solver_called = 1;
      guarded_width(P, o.init_width, S[k].width_reduced);
      need[S[k].width_reduced].push_back(k);
    }
  }
  const auto t3 = tnow();
  double t_solve = 0, t_format = 0;
  for (int wset = 0; wset < 2; ++wset) {
    const std::vector<int64_t>& idx = need[wset];
    if (idx.empty()) continue;
    const nw_opts so = solver_opts(o, wset ? o.init_width / 10 : o.init_width);
    std::vector<nw_problem> probs(idx.size());
    parallel_for((int64_t)idx.size(), [&](int64_t q) { alns[idx[q]] = sign_matrix(Ps[idx[q]].flipped); });
    for (size_t q = 0; q < idx.size(); ++q) {
      const PrepassOut& P = Ps[idx[q]];
      probs[q] = nw_problem{alns[idx[q]].data(), P.flipped.nx, P.flipped.ny, P.flipped.cn.data(), P.flipped.rn.data()};
    }
    std::vector<nw_result*> res(idx.size(), nullptr);
    const auto ts0 = tnow();
    if (idx.size() == 1) {
      NWCK(nw_solve(ctx, probs[0].aln_sign, probs[0].n_x, probs[0].n_y, probs[0].len_x, probs[0].len_y, &so, &res[0]));
    } else {
      NWCK(nw_solve_many(ctx, probs.data(), (int64_t)idx.size(), &so, 512, res.data()));
    }
    const auto ts1 = tnow();
    t_solve += tms(ts0, ts1);
    try {
      parallel_for((int64_t)idx.size(), [&](int64_t q) {
        const int64_t k = idx[q];
        nw_stats st{};
        NWCK(nw_result_stats(res[q], &st));
        std::unique_ptr<nw_table> J(format_solver_result(res[q], regs[k].gridlines_x, regs[k].n_gridlines));
        if (idx.size() == 1 || q % 512 == 0) S[k].gpu_launches += st.kernel_launches;  // counters are per pass
        apply_solver_table(Ps[k], S[k], J, o);
      });
    } catch (...) {
      for (nw_result* r : res) nw_result_free(r);
      throw;
    }
    for (nw_result* r : res) nw_result_free(r);
    t_format += tms(ts1, tnow());
  }
  const auto t4 = tnow();
  parallel_for(n, [&](int64_t k) {
    finish_region(Ps[k].table, S[k]);
    if (summaries) summaries[k] = S[k];
  });
  for (int64_t k = 0; k < n; ++k) outs[k] = Ps[k].table.release();
  if (dbg)
    fprintf(stderr, "[nw_region] %lld regions: prepare %.1f ms, prepass gpu %.1f ms, replay %.1f ms, solver %.1f ms (%zu regions), "
                    "format+merge %.1f ms, finish %.1f ms\n", (long long)n, tms(t0, t1), tms(t1, t2), tms(t2, t3), t_solve,
            need[0].size() + need[1].size(), t_format, tms(t4, tnow()));
}

// A batch with per-region status: if the batch as a whole is refused (a gridmatrix that is not 2 x 2 at least, a
// length outside the int32 DP range, ...) its regions are analysed one by one and only the offending ones fail --
// the reference drops such a region in its tryCatch (R/nahrwhals.R:204, :274) and carries on.
void analyse_many_isolated(nw_ctx* ctx, const nw_region_problem* regs, int64_t n, const nw_region_opts& o, nw_table** outs,
                           nw_region_summary* summaries, int32_t* rcs) {
  if (!rcs) {
    analyse_many(ctx, regs, n, o, outs, summaries);
    return;
  }
  try {
    analyse_many(ctx, regs, n, o, outs, summaries);
    for (int64_t k = 0; k < n; ++k) rcs[k] = NW_OK;
    return;
  } catch (const RErr& e) {
    if (e.code == NW_ERR_CUDA || e.code == NW_ERR_NOMEM) throw;  // not a property of one region
    for (int64_t k = 0; k < n; ++k) {
      nw_table_free(outs[k]);
      outs[k] = nullptr;
    }
    cudaGetLastError();
  }
  for (int64_t k = 0; k < n; ++k) {
    try {
      analyse_many(ctx, regs + k, 1, o, outs + k, summaries ? summaries + k : nullptr);
      rcs[k] = NW_OK;
    } catch (const RErr& e) {
      if (e.code == NW_ERR_CUDA || e.code == NW_ERR_NOMEM) throw;
      nw::set_last_error("region " + std::to_string(k) + ": " + e.msg);
      rcs[k] = e.code;
      outs[k] = nullptr;
      if (summaries) memset(&summaries[k], 0, sizeof summaries[k]);
      cudaGetLastError();
    }
  }
}

}  // namespace

extern "C" {

void nw_region_default_opts(nw_region_opts* o) {
  if (!o) return;
  o->depth = 3;
  o->maxdup = 2;
  o->init_width = 1000;
  o->minreport = 0.98;
  o->compression = 1000;
  o->flags = 0;
  o->reserved = 0;
}

int nw_region_prepass(nw_ctx* ctx, const int64_t* grid, int32_t n_x, int32_t n_y, const int64_t* len_x,
                      const int64_t* len_y, const nw_region_opts* opts, nw_table** out, nw_region_summary* summary) {
  if (out) *out = nullptr;
  return rguard([&] {
    if (!ctx || !opts || !out) rfail(NW_ERR_ARG, "null argument");
    PrepassOut P;
    prepass_prepare(P, make_grid(grid, n_x, n_y, len_x, len_y));
    std::vector<PrepassOut*> one{&P};
    prepass_gpu(ctx, one, opts->flags, P.launches);
    prepass_replay(P, *opts);
    nw_region_summary S = P.table->summ;  // res_ref / res_max / n_res_max / maxres of the pre-pass table
    fill_prepass_summary(S, P);
    P.table->summ = S;
    if (summary) *summary = S;
    *out = P.table.release();
  });
}

int nw_region_format(const char* tsv, int64_t tsv_len, const double* gridlines_x, int32_t n_gridlines, nw_table** out) {
  if (out) *out = nullptr;
  return rguard([&] {
    if (!tsv || tsv_len < 0 || !out || (n_gridlines > 0 && !gridlines_x)) rfail(NW_ERR_ARG, "null argument");
    *out = format_solver_output(tsv, tsv_len, gridlines_x, n_gridlines);
  });
}

int nw_region_analyse(nw_ctx* ctx, const int64_t* grid, int32_t n_x, int32_t n_y, const int64_t* len_x,
                      const int64_t* len_y, const double* gridlines_x, int32_t n_gridlines,
                      const nw_region_opts* opts, nw_table** out, nw_region_summary* summary) {
  if (out) *out = nullptr;
  return rguard([&] {
    if (!ctx || !opts || !out) rfail(NW_ERR_ARG, "null argument");
    const nw_region_problem r{grid, n_x, n_y, len_x, len_y, gridlines_x, n_gridlines};
    analyse_many(ctx, &r, 1, *opts, out, summary);
  });
}

int nw_region_analyse_many(nw_ctx* ctx, const nw_region_problem* regions, int64_t n, const nw_region_opts* opts,
                           nw_table** outs, nw_region_summary* summaries, int32_t* rcs) {
  if (outs)
    for (int64_t k = 0; k < n; ++k) outs[k] = nullptr;
  return rguard([&] {
    if (!ctx || !opts || !outs || n < 0 || (n > 0 && !regions)) rfail(NW_ERR_ARG, "null argument");
    const int64_t PASS = 4096;  // regions per pre-pass batch (the solver batches 512 at a time below that)
    for (int64_t a = 0; a < n; a += PASS) {
      const int64_t m = std::min(PASS, n - a);
      try {
        analyse_many_isolated(ctx, regions + a, m, *opts, outs + a, summaries ? summaries + a : nullptr, rcs ? rcs + a : nullptr);
      } catch (...) {
        for (int64_t k = 0; k < a; ++k) {
          nw_table_free(outs[k]);
          outs[k] = nullptr;
        }
        throw;
      }
    }
  });
}

int nw_region_analyse_many_ctxs(nw_ctx** ctxs, int32_t n_ctx, const nw_region_problem* regions, int64_t n,
                                const nw_region_opts* opts, int32_t regions_per_batch, nw_table** outs,
                                nw_region_summary* summaries, int32_t* rcs) {
  if (outs)
    for (int64_t k = 0; k < n; ++k) outs[k] = nullptr;
  return rguard([&] {
    if (!ctxs || n_ctx < 1 || !opts || !outs || n < 0 || (n > 0 && !regions)) rfail(NW_ERR_ARG, "null argument");
    for (int t = 0; t < n_ctx; ++t)
      if (!ctxs[t]) rfail(NW_ERR_ARG, "null context");
    const int64_t B = regions_per_batch > 0 ? std::min<int64_t>(regions_per_batch, 4096) : 1024;
    const int64_t nb = (n + B - 1) / B;
    std::atomic<int64_t> next{0};
    std::mutex mu;
    bool failed = false;
    RErr err{NW_OK, ""};
    const int per_ctx = std::max(1, nw::host_cores() / n_ctx);
    auto work = [&](int t) {
      g_host_threads = per_ctx;
      for (;;) {
        const int64_t b = next.fetch_add(1);
        if (b >= nb) return;
        const int64_t a = b * B, m = std::min(B, n - a);
        try {
          analyse_many_isolated(ctxs[t], regions + a, m, *opts, outs + a, summaries ? summaries + a : nullptr,
                                rcs ? rcs + a : nullptr);
        } catch (const RErr& e) {
          std::lock_guard<std::mutex> lk(mu);
          if (!failed) err = e;
          failed = true;
          next.store(nb);
        } catch (const std::exception& e) {
          std::lock_guard<std::mutex> lk(mu);
          if (!failed) err = RErr{NW_ERR_ARG, e.what()};
          failed = true;
          next.store(nb);
        }
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_ctx; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    g_host_threads = 12;
    if (failed) {
      for (int64_t k = 0; k < n; ++k) {
        nw_table_free(outs[k]);
        outs[k] = nullptr;
      }
      throw err;
    }
  });
}

int nw_table_summary(const nw_table* t, nw_region_summary* out) {
  if (!t || !out) {
    nw::set_last_error("null argument");
    return NW_ERR_ARG;
  }
  *out = t->summ;
  return NW_OK;
}
const char* nw_table_mut_max(const nw_table* t) { return t ? t->mut_max.c_str() : ""; }
int nw_table_tsv(const nw_table* t, const char** data, int64_t* len) {
  if (!t || !data || !len) {
    nw::set_last_error("null argument");
    return NW_ERR_ARG;
  }
  if (!t->tsv_ready) {
    render_tsv(*t, t->tsv);
    t->tsv_ready = true;
  }
  *data = t->tsv.data();
  *len = (int64_t)t->tsv.size();
  return NW_OK;
}
int64_t nw_table_candidates(const nw_table* t, const nw_candidate** out) {
  if (!t || !out) return 0;
  *out = t->cands.data();
  return (int64_t)t->cands.size();
}
int32_t nw_table_rows(const nw_table* t) { return t ? (int32_t)t->rows.size() : 0; }
void nw_table_free(nw_table* t) { delete t; }

// ---- save_to_logfile (R/save_logfile.R:112-232): the row of nahrwhals_res.tsv ----
static const char* kLogHeader =
    "seqname\tstart\tend\tsample\ty_seqname\ty_start\ty_end\twidth_orig\txpad\tres_ref\tres_max\tn_res_max\tmut_max\t"
    "mut_simulated\tmut_tested\tsearch_depth\tgrid_compression\texceeds_x\texceeds_y\tgrid_inconsistency\tflip_unsure\t"
    "cluttered_boundaries\tmut1_start\tmut1_end\tmut1_pos_pm\tmut1_len\tmut1_len_pm\tmut2_start\tmut2_end\tmut2_pos_pm\t"
    "mut2_len\tmut2_len_pm\tmut3_start\tmut3_end\tmut3_pos_pm\tmut3_len\tmut3_len_pm";
const char* nw_region_logrow_header(void) { return kLogHeader; }

// a double as write.table / as.character print it under options(scipen = 999): the fewest significant digits (<= 15)
// that reproduce the value at 15 digits, fixed notation
static std::string r_fixed15(double v) {
  if (std::isnan(v)) return "NA";
  if (std::isinf(v)) return v > 0 ? "Inf" : "-Inf";
  if (v == 0) return "0";
  char ref[64], buf[400];
  snprintf(ref, sizeof ref, "%.14e", v);
  const double target = strtod(ref, nullptr);
  int nsig = 15;
  for (int d = 1; d <= 15; ++d) {
    snprintf(buf, sizeof buf, "%.*e", d - 1, v);
    if (strtod(buf, nullptr) == target) {
      nsig = d;
      break;
    }
  }
  snprintf(buf, sizeof buf, "%.*e", nsig - 1, v);
  const int kp = atoi(strchr(buf, 'e') + 1);
  snprintf(buf, sizeof buf, "%.*f", std::max(0, nsig - kp - 1), strtod(buf, nullptr));
  return buf;
}

int nw_region_logrow(const nw_region_summary* s, const char* mut_max, const nw_log_meta* m, char* buf, int64_t cap,
                     int64_t* len) {
  if (!s || !m || !len || (cap > 0 && !buf)) {
    nw::set_last_error("null argument");
    return NW_ERR_ARG;
  }
  auto txt = [](const char* t) { return std::string(t ? t : "NA"); };
  auto num = [](const char* t) { return t ? strtod(t, nullptr) : std::numeric_limits<double>::quiet_NaN(); };
  auto lgl = [](int32_t v) { return std::string(v < 0 ? "NA" : (v ? "TRUE" : "FALSE")); };
  const double start = num(m->start), end = num(m->end);
  std::vector<std::string> f = {txt(m->chr), txt(m->start), txt(m->end), txt(m->samplename), txt(m->y_seqname), txt(m->y_start),
                                txt(m->y_end), r_fixed15(end - start), txt(m->xpad), r_fixed15(s->res_ref), r_fixed15(s->res_max),
                                std::to_string((long long)s->n_res_max), txt(mut_max && *mut_max ? mut_max : nullptr),
                                std::to_string((long long)s->mut_simulated), std::to_string((long long)s->mut_tested),
                                std::to_string(s->search_depth), r_fixed15(std::isnan(m->compression) ? 0.0 : m->compression),
                                lgl(m->exceeds_x), lgl(m->exceeds_y), lgl(m->grid_inconsistency), lgl(s->flip_unsure),
                                lgl(s->cluttered)};
  for (int k = 0; k < 3; ++k) {
    f.push_back(r_fixed15(s->maxres[5 * k] + start));  // as.character(as.numeric(maxres$mutK_start) + as.numeric(log$start))
    f.push_back(r_fixed15(s->maxres[5 * k + 1] + start));
    for (int q = 2; q < 5; ++q) f.push_back(r_fixed15(s->maxres[5 * k + q]));
  }
  std::string row;
  for (size_t k = 0; k < f.size(); ++k) {
    if (k) row += '\t';
    row += f[k];
  }
  *len = (int64_t)row.size();
  if ((int64_t)row.size() + 1 > cap) {
    nw::set_last_error("nw_region_logrow: buffer too small");
    return NW_ERR_RANGE;
  }
  memcpy(buf, row.c_str(), row.size() + 1);
  return NW_OK;
}

int nw_region_log_append(const char* logfile, const char* row) {
  if (!logfile || !row) {
    nw::set_last_error("null argument");
    return NW_ERR_ARG;
  }
  const int fd = open(logfile, O_WRONLY | O_CREAT | O_APPEND, 0644);
  if (fd < 0) {
    nw::set_last_error(std::string("cannot open ") + logfile);
    return NW_ERR_IO;
  }
  int rc = NW_OK;
  if (flock(fd, LOCK_EX) != 0) {
    rc = NW_ERR_IO;
  } else {
    struct stat st;
    std::string out;
    if (fstat(fd, &st) == 0 && st.st_size == 0) out = std::string(kLogHeader) + "\n";
    out += row;
    out += '\n';
    size_t off = 0;
    while (off < out.size()) {
      const ssize_t w = write(fd, out.data() + off, out.size() - off);
      if (w <= 0) {
        rc = NW_ERR_IO;
        break;
      }
      off += (size_t)w;
    }
    flock(fd, LOCK_UN);
  }
  close(fd);
  if (rc != NW_OK) nw::set_last_error(std::string("cannot write ") + logfile);
  return rc;
}

}  // extern "C"
