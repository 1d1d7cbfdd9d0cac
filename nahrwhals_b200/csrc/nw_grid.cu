// nw_grid.cu -- gridding on the GPU (include/nwgrid.h): alignments -> gridlines -> gridmatrix.
//
// k_gridlines: one CTA per region.  The sorted gridline sets of both axes live in shared memory; a generation of
// make_xy_grid_fast (R/tsv_to_bitlocus.R:305-380) is one strided pass over (alignment, gridline-to-process) pairs with
// a binary search for `%in% all_x`, a bitonic sort + unique of the new lines, and a merge by re-sorting.  Set
// semantics make the result independent of the order in which threads append candidates.
// k_grid_fill: one thread per alignment walks the grid as bresenham() does (R/bresenham.R:45-124) and merges every cell
// into the region's matrix: EMPTY -> z by compare-and-swap, a second writer with a different z turns the cell into 0
// (remove_duplicates_triple: same (x, y) with different z -> 0).  k_grid_finish replaces EMPTY by 0.
// Every double expression uses the explicit round-to-nearest intrinsics so that no FMA is formed.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../../include/nwgrid.h"

struct nw_ctx;
namespace nw {
void set_last_error(const std::string& m);
int ctx_device(nw_ctx* c);
cudaStream_t ctx_stream(nw_ctx* c);
cudaError_t ctx_region_staging(nw_ctx* c, size_t bytes, void** host_pinned, void** dev);
}  // namespace nw

namespace {

constexpr int CAP = NW_GRID_MAX_LINES;
constexpr int GT = 256;  // threads per CTA of k_gridlines
constexpr size_t GRID_BATCH = 8192;  // regions per pass of nw_grid_build
constexpr unsigned long long EMPTY_BITS = 0x7ff8dead00000001ull;  // a NaN payload no computation produces

enum : int { ST_OK = 0, ST_OVERFLOW = 1, ST_WALK = 2 };

struct RegionIn {  // per region, in the packed device buffer
  long long aln_off;  // first alignment (index into ts/te/qs/qe)
  int n_aln;
  int pad;
};
struct RegionOut {
  int n_gx, n_gy, status, converged, generations, pad;
  long long mat_off;  // filled by the host between the two kernels
};

__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7ff0000000000000ll); }

// ascending bitonic sort of CAP doubles in shared memory (unused tail = +inf)
__device__ void sort_cap(double* a) {
  for (int k = 2; k <= CAP; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < CAP; i += GT) {
        const int l = i ^ j;
        if (l > i) {
          const double x = a[i], y = a[l];
          const bool up = (i & k) == 0;
          if ((x > y) == up) {
            a[i] = y;
            a[l] = x;
          }
        }
      }
    }
  __syncthreads();
}

// in-place unique of a sorted array of n valid entries (tail +inf); returns the new count to every thread
__device__ int unique_sorted(double* a, int n, double* tmp, int* cnt) {
  if (threadIdx.x == 0) *cnt = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < CAP; i += GT) tmp[i] = a[i];
  __syncthreads();
  // keep a[i] iff first or different from its predecessor; position = number of kept entries before it
  // (n <= 1024: a serial scan by one thread is a few microseconds and keeps the code obvious)
  if (threadIdx.x == 0) {
    int m = 0;
    for (int i = 0; i < n; ++i)
      if (i == 0 || tmp[i] != tmp[i - 1]) a[m++] = tmp[i];
    for (int i = m; i < CAP; ++i) a[i] = dinf();
    *cnt = m;
  }
  __syncthreads();
  return *cnt;
}

__device__ __forceinline__ bool contains(const double* a, int n, double v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo < n && a[lo] == v;
}
__device__ __forceinline__ int first_index(const double* a, int n, double v) {  // match(v, a) - 1, or -1
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v)
      lo = mid + 1;
    else
      hi = mid;
  }
  return (lo < n && a[lo] == v) ? lo : -1;
}

__global__ void __launch_bounds__(GT) k_gridlines(const RegionIn* __restrict__ rin, const double* __restrict__ ts,
                                                  const double* __restrict__ te, const double* __restrict__ qs,
                                                  const double* __restrict__ qe, RegionOut* __restrict__ rout,
                                                  double* __restrict__ gx_out, double* __restrict__ gy_out) {
  extern __shared__ __align__(16) unsigned char sm[];
  double* ax = reinterpret_cast<double*>(sm);  // all_x, sorted
  double* ay = ax + CAP;
  double* px = ay + CAP;  // lines to process in this generation
  double* py = px + CAP;
  double* nx = py + CAP;  // new candidates of this generation
  double* ny = nx + CAP;
  __shared__ int s_cnt, s_nnx, s_nny, s_genx, s_geny, s_over;
  const int reg = blockIdx.x;
  const RegionIn R = rin[reg];
  const double* Ts = ts + R.aln_off;
  const double* Te = te + R.aln_off;
  const double* Qs = qs + R.aln_off;
  const double* Qe = qe + R.aln_off;
  const int n = R.n_aln;
  RegionOut O{0, 0, ST_OK, 0, 0, 0, 0};
  // all_x <- sort(unique(c(tstart, tend))), all_y likewise (the host guarantees 2n <= CAP)
  for (int i = threadIdx.x; i < CAP; i += GT) {
    ax[i] = i < n ? Ts[i] : (i < 2 * n ? Te[i - n] : dinf());
    ay[i] = i < n ? Qs[i] : (i < 2 * n ? Qe[i - n] : dinf());
  }
  if (threadIdx.x == 0) s_over = 0;
  sort_cap(ax);
  int n_ax = unique_sorted(ax, 2 * n, nx, &s_cnt);
  sort_cap(ay);
  int n_ay = unique_sorted(ay, 2 * n, nx, &s_cnt);
  int n_px = n_ax, n_py = n_ay;
  for (int i = threadIdx.x; i < CAP; i += GT) {
    px[i] = ax[i];
    py[i] = ay[i];
  }
  __syncthreads();
  int gen = 0;
  bool converged = false;
  for (gen = 1; gen <= 100000; ++gen) {
    if (threadIdx.x == 0) s_nnx = s_nny = s_genx = s_geny = 0;
    for (int i = threadIdx.x; i < CAP; i += GT) nx[i] = ny[i] = dinf();
    __syncthreads();
    const int np = max(n_px, n_py);
    for (long long w = threadIdx.x; w < (long long)n * np; w += GT) {
      const int r = (int)(w / np), e = (int)(w % np);
      const double t0 = Ts[r], t1 = Te[r], q0 = Qs[r], q1 = Qe[r];
      const double slope = __ddiv_rn(__dsub_rn(q1, q0), __dsub_rn(t1, t0));
      if (e < n_px) {
        const double v = px[e];
        if (v > t0 && v < t1) {  // current_xs -> new_y_vals = qstart + slope * delta_x
          const double c = __dadd_rn(q0, __dmul_rn(slope, __dsub_rn(v, t0)));
          atomicAdd(&s_geny, 1);
          if (!contains(ay, n_ay, c)) {
            const int k = atomicAdd(&s_nny, 1);
            if (k < CAP)
              ny[k] = c;
            else
              s_over = 1;
          }
        }
      }
      if (e < n_py) {
        const double v = py[e];
        if (v > fmin(q0, q1) && v < fmax(q0, q1)) {  // current_ys -> new_x_vals = tstart + delta_y / slope
          const double c = __dadd_rn(t0, __ddiv_rn(__dsub_rn(v, q0), slope));
          atomicAdd(&s_genx, 1);
          if (!contains(ax, n_ax, c)) {
            const int k = atomicAdd(&s_nnx, 1);
            if (k < CAP)
              nx[k] = c;
            else
              s_over = 1;
          }
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      // new_x[1:(idx_x - 1)] with idx_x == 1 is the first (zero) element of the pre-allocated buffer
      // (R/tsv_to_bitlocus.R:356-357); max_size = n * max(length(process_x), length(process_y)) > 0 here
      if (s_genx == 0 && !contains(ax, n_ax, 0.0)) nx[s_nnx++] = 0.0;
      if (s_geny == 0 && !contains(ay, n_ay, 0.0)) ny[s_nny++] = 0.0;
    }
    __syncthreads();
    if (s_over) break;
    const int cx = min(s_nnx, CAP), cy = min(s_nny, CAP);
    sort_cap(nx);
    const int ux = unique_sorted(nx, cx, px, &s_cnt);
    sort_cap(ny);
    const int uy = unique_sorted(ny, cy, px, &s_cnt);
    if (ux == 0 && uy == 0) {
      converged = true;
      break;
    }
    if (n_ax + ux > CAP || n_ay + uy > CAP) {
      if (threadIdx.x == 0) s_over = 1;
      __syncthreads();
      break;
    }
    for (int i = threadIdx.x; i < CAP; i += GT) {  // process lists of the next generation = the new lines
      px[i] = nx[i];
      py[i] = ny[i];
    }
    for (int i = threadIdx.x; i < ux; i += GT) ax[n_ax + i] = nx[i];
    for (int i = threadIdx.x; i < uy; i += GT) ay[n_ay + i] = ny[i];
    n_ax += ux;
    n_ay += uy;
    n_px = ux;
    n_py = uy;
    __syncthreads();
    sort_cap(ax);  // new lines are not in the set, so sorting keeps it duplicate-free
    sort_cap(ay);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_ax; i += GT) gx_out[(size_t)reg * CAP + i] = ax[i];
  for (int i = threadIdx.x; i < n_ay; i += GT) gy_out[(size_t)reg * CAP + i] = ay[i];
  if (threadIdx.x == 0) {
    O.n_gx = n_ax;
    O.n_gy = n_ay;
    O.status = s_over ? ST_OVERFLOW : ST_OK;
    O.converged = converged ? 1 : 0;
    O.generations = min(gen, 100000);
    rout[reg] = O;
  }
}

// one thread per alignment: bresenham() of the reference over the finished grid
__global__ void k_grid_fill(const RegionIn* __restrict__ rin, const int* __restrict__ aln_region, long long n_aln_total,
                            const double* __restrict__ ts, const double* __restrict__ te, const double* __restrict__ qs,
                            const double* __restrict__ qe, RegionOut* __restrict__ rout, const double* __restrict__ gx_all,
                            const double* __restrict__ gy_all, unsigned long long* __restrict__ mats) {
  const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_aln_total) return;
  const int reg = aln_region[a];
  RegionOut& O = rout[reg];
  if (O.status != ST_OK) return;
  const double* gx = gx_all + (size_t)reg * CAP;
  const double* gy = gy_all + (size_t)reg * CAP;
  const int ngx = O.n_gx, ngy = O.n_gy, nxb = ngx - 1;
  unsigned long long* mat = mats + O.mat_off;
  double x = ts[a], y = qs[a];
  const double x1 = te[a], y1 = qe[a];
  int xi = first_index(gx, ngx, x), yi = first_index(gy, ngy, y);  // 0-based match()
  const bool pos = y < y1;
  if (xi < 0 || yi < 0) {
    O.status = ST_WALK;
    return;
  }
  for (int step = 0; x != x1 && y != y1; ++step) {
    if (step > ngx + ngy || xi + 1 >= ngx || (pos ? yi + 1 >= ngy : yi - 1 < 0)) {
      O.status = ST_WALK;  // the reference indexes past the grid here (NA in the loop condition -> error)
      return;
    }
    const double sx = __dsub_rn(gx[xi + 1], gx[xi]);
    const double sy = pos ? __dsub_rn(gy[yi + 1], gy[yi]) : -__dsub_rn(gy[yi], gy[yi - 1]);
    x = __dadd_rn(x, sx);
    y = __dadd_rn(y, sy);
    const int cy = pos ? yi : yi - 1;
    const double z = pos ? sx : -sx;
    unsigned long long* cell = mat + (size_t)cy * nxb + xi;
    const unsigned long long zb = (unsigned long long)__double_as_longlong(z);
    const unsigned long long old = atomicCAS(cell, EMPTY_BITS, zb);
    if (old != EMPTY_BITS && old != zb) atomicExch(cell, 0ull);  // two different z on one cell -> 0.0
    xi += 1;
    yi += pos ? 1 : -1;
  }
}

__global__ void k_grid_clear(unsigned long long* mats, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mats[i] = EMPTY_BITS;
}
__global__ void k_grid_finish(unsigned long long* mats, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && mats[i] == EMPTY_BITS) mats[i] = 0ull;
}

struct GErr {
  int code;
  std::string msg;
};
[[noreturn]] void gfail(int code, const std::string& m) { throw GErr{code, m}; }
#define GCK(call)                                                                                 \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) gfail(NW_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

}  // namespace

struct nw_grid {
  int n_x = 0, n_y = 0, converged = 0, generations = 0;
  std::vector<double> gx, gy, mat;
};

extern "C" {

int nw_grid_build(nw_ctx* ctx, const nw_paf* pafs, int64_t n_regions, nw_grid** outs, int32_t* rcs) {
  if (outs)
    for (int64_t k = 0; k < n_regions; ++k) outs[k] = nullptr;
  int code = NW_ERR_CUDA;
  try {
    if (!ctx || !outs || n_regions < 0 || (n_regions > 0 && !pafs)) gfail(NW_ERR_ARG, "null argument");
    if (n_regions == 0) return NW_OK;
    if (n_regions > (1 << 20)) gfail(NW_ERR_ARG, "too many regions in one call");
    std::vector<int32_t> rc((size_t)n_regions, NW_OK);
    std::vector<int64_t> live;  // regions that go to the device
    long long n_aln = 0;
    for (int64_t k = 0; k < n_regions; ++k) {
      const nw_paf& p = pafs[k];
      if (p.n < 1 || !p.tstart || !p.tend || !p.qstart || !p.qend) {
        rc[k] = NW_ERR_ARG;
        continue;
      }
      if (p.n > NW_GRID_MAX_ALNS) {
        rc[k] = NW_ERR_RANGE;
        continue;
      }
      bool ok = true;
      for (int64_t i = 0; i < p.n && ok; ++i)
        ok = std::isfinite(p.tstart[i]) && std::isfinite(p.tend[i]) && std::isfinite(p.qstart[i]) && std::isfinite(p.qend[i]) &&
             p.tstart[i] < p.tend[i] && p.qstart[i] != p.qend[i];
      if (!ok) {
        rc[k] = NW_ERR_ARG;  // enforce_slope_one (R/tsv_to_bitlocus.R:258-275) removes such alignments upstream
        continue;
      }
      live.push_back(k);
      n_aln += p.n;
    }
    // batches of at most GRID_BATCH regions: the staging area holds 16 KB of gridline buffers per region
    const std::vector<int64_t> live_all = std::move(live);
    for (size_t b0 = 0; b0 < live_all.size(); b0 += GRID_BATCH) {
      live.assign(live_all.begin() + b0, live_all.begin() + std::min(live_all.size(), b0 + GRID_BATCH));
      n_aln = 0;
      for (int64_t k : live) n_aln += pafs[k].n;
      const size_t R = live.size();
      GCK(cudaSetDevice(nw::ctx_device(ctx)));
      cudaStream_t st = nw::ctx_stream(ctx);
      auto up16 = [](size_t v) { return (v + 15) / 16 * 16; };
      // staging layout (same offsets on host and device): inputs | outputs of k_gridlines
      const size_t o_rin = 0, o_areg = o_rin + up16(R * sizeof(RegionIn)), o_ts = o_areg + up16((size_t)n_aln * 4),
                   o_te = o_ts + (size_t)n_aln * 8, o_qs = o_te + (size_t)n_aln * 8, o_qe = o_qs + (size_t)n_aln * 8,
                   o_rout = o_qe + (size_t)n_aln * 8, o_gx = o_rout + up16(R * sizeof(RegionOut)),
                   o_gy = o_gx + R * CAP * 8, total = o_gy + R * CAP * 8;
      void *hp = nullptr, *dp = nullptr;
      GCK(nw::ctx_region_staging(ctx, total, &hp, &dp));
      unsigned char *h = static_cast<unsigned char*>(hp), *d = static_cast<unsigned char*>(dp);
      RegionIn* rin = reinterpret_cast<RegionIn*>(h + o_rin);
      int* areg = reinterpret_cast<int*>(h + o_areg);
      double *hts = reinterpret_cast<double*>(h + o_ts), *hte = reinterpret_cast<double*>(h + o_te),
             *hqs = reinterpret_cast<double*>(h + o_qs), *hqe = reinterpret_cast<double*>(h + o_qe);
      long long off = 0;
      for (size_t r = 0; r < R; ++r) {
        const nw_paf& p = pafs[live[r]];
        rin[r] = RegionIn{off, (int)p.n, 0};
        memcpy(hts + off, p.tstart, (size_t)p.n * 8);
        memcpy(hte + off, p.tend, (size_t)p.n * 8);
        memcpy(hqs + off, p.qstart, (size_t)p.n * 8);
        memcpy(hqe + off, p.qend, (size_t)p.n * 8);
        for (int64_t i = 0; i < p.n; ++i) areg[off + i] = (int)r;
        off += p.n;
      }
      static bool attr_done = false;
      const int smem = 6 * CAP * 8;
      if (!attr_done) {
        GCK(cudaFuncSetAttribute(k_gridlines, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_done = true;
      }
      GCK(cudaMemcpyAsync(d, h, o_rout, cudaMemcpyHostToDevice, st));
      k_gridlines<<<(unsigned)R, GT, smem, st>>>((const RegionIn*)(d + o_rin), (const double*)(d + o_ts), (const double*)(d + o_te),
                                                 (const double*)(d + o_qs), (const double*)(d + o_qe), (RegionOut*)(d + o_rout),
                                                 (double*)(d + o_gx), (double*)(d + o_gy));
      GCK(cudaGetLastError());
      GCK(cudaMemcpyAsync(h + o_rout, d + o_rout, up16(R * sizeof(RegionOut)), cudaMemcpyDeviceToHost, st));
      GCK(cudaStreamSynchronize(st));
      RegionOut* rout = reinterpret_cast<RegionOut*>(h + o_rout);
      long long cells = 0;
      for (size_t r = 0; r < R; ++r) {
        rout[r].mat_off = cells;
        if (rout[r].status == ST_OK) cells += (long long)(rout[r].n_gx - 1) * (rout[r].n_gy - 1);
      }
      unsigned long long* dmat = nullptr;
      std::vector<unsigned long long> hmat((size_t)cells);
      if (cells > 0) {
        GCK(cudaMalloc(&dmat, (size_t)cells * 8));
        try {
          k_grid_clear<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(dmat, cells);
          GCK(cudaGetLastError());
          GCK(cudaMemcpyAsync(d + o_rout, h + o_rout, up16(R * sizeof(RegionOut)), cudaMemcpyHostToDevice, st));
          k_grid_fill<<<(unsigned)((n_aln + 127) / 128), 128, 0, st>>>(
              (const RegionIn*)(d + o_rin), (const int*)(d + o_areg), n_aln, (const double*)(d + o_ts), (const double*)(d + o_te),
              (const double*)(d + o_qs), (const double*)(d + o_qe), (RegionOut*)(d + o_rout), (const double*)(d + o_gx),
              (const double*)(d + o_gy), dmat);
          GCK(cudaGetLastError());
          k_grid_finish<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(dmat, cells);
          GCK(cudaGetLastError());
          GCK(cudaMemcpyAsync(hmat.data(), dmat, (size_t)cells * 8, cudaMemcpyDeviceToHost, st));
        } catch (...) {
          cudaFree(dmat);
          throw;
        }
      }
      {  // status words, and only the used prefix of every region's gridline rows
        int mx = 2;
        for (size_t r = 0; r < R; ++r) mx = std::max(mx, std::max(rout[r].n_gx, rout[r].n_gy));
        GCK(cudaMemcpyAsync(h + o_rout, d + o_rout, up16(R * sizeof(RegionOut)), cudaMemcpyDeviceToHost, st));
        GCK(cudaMemcpy2DAsync(h + o_gx, (size_t)CAP * 8, d + o_gx, (size_t)CAP * 8, (size_t)mx * 8, 2 * R, cudaMemcpyDeviceToHost, st));
      }
      GCK(cudaStreamSynchronize(st));
      if (dmat) cudaFree(dmat);
      const double* hgx = reinterpret_cast<const double*>(h + o_gx);
      const double* hgy = reinterpret_cast<const double*>(h + o_gy);
      for (size_t r = 0; r < R; ++r) {
        const RegionOut& O = rout[r];
        const int64_t k = live[r];
        if (O.status == ST_OVERFLOW) {
          rc[k] = NW_ERR_RANGE;
          continue;
        }
        if (O.status == ST_WALK || O.n_gx < 2 || O.n_gy < 2) {
          rc[k] = NW_ERR_ARG;
          continue;
        }
        std::unique_ptr<nw_grid> G(new nw_grid);
        G->n_x = O.n_gx - 1;
        G->n_y = O.n_gy - 1;
        G->converged = O.converged;
        G->generations = O.generations;
        G->gx.assign(hgx + r * CAP, hgx + r * CAP + O.n_gx);
        G->gy.assign(hgy + r * CAP, hgy + r * CAP + O.n_gy);
        G->mat.resize((size_t)G->n_x * G->n_y);
        memcpy(G->mat.data(), hmat.data() + O.mat_off, G->mat.size() * 8);
        outs[k] = G.release();
      }
    }
    if (rcs) memcpy(rcs, rc.data(), (size_t)n_regions * 4);
    return NW_OK;
  } catch (const GErr& e) {
    nw::set_last_error(e.msg);
    code = e.code;
  } catch (const std::exception& e) {
    nw::set_last_error(e.what());
    code = NW_ERR_ARG;
  }
  for (int64_t k = 0; k < n_regions && outs; ++k) {
    delete outs[k];
    outs[k] = nullptr;
  }
  cudaGetLastError();
  return code;
}

int nw_grid_dims(const nw_grid* g, int32_t* n_x, int32_t* n_y, int32_t* converged, int32_t* generations) {
  if (!g) return NW_ERR_ARG;
  if (n_x) *n_x = g->n_x;
  if (n_y) *n_y = g->n_y;
  if (converged) *converged = g->converged;
  if (generations) *generations = g->generations;
  return NW_OK;
}
int nw_grid_lines(const nw_grid* g, double* gx, double* gy) {
  if (!g) return NW_ERR_ARG;
  if (gx) memcpy(gx, g->gx.data(), g->gx.size() * 8);
  if (gy) memcpy(gy, g->gy.data(), g->gy.size() * 8);
  return NW_OK;
}
int nw_grid_matrix(const nw_grid* g, double* mat, double* len_x, double* len_y) {
  if (!g) return NW_ERR_ARG;
  if (mat) memcpy(mat, g->mat.data(), g->mat.size() * 8);
  if (len_x)
    for (int i = 0; i < g->n_x; ++i) len_x[i] = g->gx[i + 1] - g->gx[i];
  if (len_y)
    for (int i = 0; i < g->n_y; ++i) len_y[i] = g->gy[i + 1] - g->gy[i];
  return NW_OK;
}
void nw_grid_free(nw_grid* g) { delete g; }

}  // extern "C"
