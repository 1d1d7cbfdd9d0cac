// nwsolve -- command-line twin of `julia solver.jl` (argument table: solver.jl:756-791 in the reference).
//   nwsolve [-d D] [-u U] [-w W] [-O O] [-r r] [-R R] compressed_alignment compressed_lengths out_path
//   nwsolve [flags] --batch list.tsv [--contexts K] [--regions-per-pass N] [--devices 0,1,...|all]
//   nwsolve [flags] --devices 0,1,...|all compressed_alignment compressed_lengths out_path
// --devices spreads the work over several GPUs of the machine from this one process (no launcher): a batch hands its
// passes to `--contexts` contexts on every listed device (regions are independent, R/nahrwhals.R:171-172: no collective);
// a single search is sharded over one context per listed device (nw_solve_sharded_ctxs: replicated expansion, the DP of
// every level divided between the devices, one exchange per level).  A device may be listed twice (two ranks on it).
// --batch is the regionfile / whole-genome form (R/nahrwhals.R:171-215, :240-290 run one solver process per region):
// every line of list.tsv is "compressed_alignment<TAB>compressed_lengths<TAB>out_path"; all regions are searched by
// one process through nw_solve_files_many (forest search).  A region that fails leaves no out_path and a message.
// Exit status 0 and out_path written on success; non-zero, a message on stderr and NO out_path otherwise
// (R/juliasolver.R:56-59 treats a missing file as failure).  -O/--maxoffset is accepted and ignored, as in the
// reference where it lands in an unused positional of slow_score (solver.jl:329).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nwsolve.h"

static void usage() {
  fprintf(stderr,
          "usage: nwsolve [--maxdepth|-d D] [--maxdup|-u U] [--maxwidth|-w W] [--maxoffset|-O O]\n"
          "               [--minreport|-r r] [--maxreport|-R R] [--device N] [--stats]\n"
          "               compressed_alignment compressed_lengths out_path\n"
          "       nwsolve [flags] --batch list.tsv [--contexts K] [--regions-per-pass N] [--devices 0,1,...|all]\n"
          "               (list.tsv: compressed_alignment<TAB>compressed_lengths<TAB>out_path per line)\n"
          "       nwsolve [flags] --devices 0,1,...|all compressed_alignment compressed_lengths out_path\n"
          "               (one search sharded over the listed GPUs)\n");
}

// "0,1,3" or "all" -> device indices (empty on a malformed list)
static std::vector<int> parse_devices(const char* spec) {
  std::vector<int> out;
  if (!strcmp(spec, "all")) {
    const int n = nw_device_count();
    for (int k = 0; k < n; ++k) out.push_back(k);
    return out;
  }
  const char* p = spec;
  while (*p) {
    char* end = nullptr;
    const long v = strtol(p, &end, 10);
    if (end == p || v < 0 || v > 1023) return {};
    out.push_back((int)v);
    p = end;
    if (*p == ',') ++p;
    else if (*p) return {};
  }
  return out;
}

static int run_batch(const char* list, const nw_opts& o, const std::vector<int>& devices, int contexts, int per_pass,
                     bool stats) {
  FILE* f = fopen(list, "rb");
  if (!f) {
    fprintf(stderr, "nwsolve: cannot open %s\n", list);
    return 1;
  }
  std::vector<std::string> a, b, c;
  std::string line;
  int ch;
  auto flush = [&]() -> bool {
    while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
    if (line.empty()) return true;
    const size_t t1 = line.find('\t'), t2 = t1 == std::string::npos ? t1 : line.find('\t', t1 + 1);
    if (t2 == std::string::npos || line.find('\t', t2 + 1) != std::string::npos) return false;
    a.push_back(line.substr(0, t1));
    b.push_back(line.substr(t1 + 1, t2 - t1 - 1));
    c.push_back(line.substr(t2 + 1));
    line.clear();
    return true;
  };
  bool ok = true;
  while (ok && (ch = fgetc(f)) != EOF) {
    if (ch == '\n') ok = flush();
    else line.push_back((char)ch);
  }
  if (ok) ok = flush();
  fclose(f);
  if (!ok) {
    fprintf(stderr, "nwsolve: %s: every line must hold three tab-separated paths\n", list);
    return 2;
  }
  const int64_t n = (int64_t)a.size();
  std::vector<nw_ctx*> ctxs;
  for (int k = 0; k < contexts; ++k)
    for (int device : devices) {  // device-major order inside a round: the first contexts to ask for work sit on different GPUs
      nw_ctx* x = nullptr;
      if (nw_ctx_create(device, &x) != NW_OK) {
        fprintf(stderr, "nwsolve: %s\n", nw_last_error());
        for (nw_ctx* y : ctxs) nw_ctx_destroy(y);
        return 1;
      }
      ctxs.push_back(x);
    }
  std::vector<const char*> pa(n), pb(n), pc(n);
  for (int64_t k = 0; k < n; ++k) {
    pa[k] = a[k].c_str();
    pb[k] = b[k].c_str();
    pc[k] = c[k].c_str();
  }
  std::vector<int32_t> rcs(n, 0);
  const int rc = nw_solve_files_many(ctxs.data(), (int32_t)ctxs.size(), pa.data(), pb.data(), pc.data(), n, &o, per_pass,
                                     rcs.data());
  int64_t bad = 0;
  for (int64_t k = 0; k < n; ++k)
    if (rcs[k] != NW_OK) {
      ++bad;
      fprintf(stderr, "nwsolve: region %lld (%s) failed with code %d\n", (long long)k, pa[k], rcs[k]);
    }
  if (rc != NW_OK) fprintf(stderr, "nwsolve: %s\n", nw_last_error());
  if (stats)
    fprintf(stderr, "nwsolve: %lld regions, %lld failed, %zu contexts on %zu device(s)\n", (long long)n, (long long)bad,
            ctxs.size(), devices.size());
  for (nw_ctx* y : ctxs) nw_ctx_destroy(y);
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  nw_opts o;
  nw_default_opts(&o);
  int device = -1;
  std::vector<int> devices;
  bool stats = false;
  const char* batch = nullptr;
  int contexts = 3, per_pass = 512;
  std::vector<const char*> pos;
  for (int a = 1; a < argc; ++a) {
    const std::string s = argv[a];
    auto val = [&](const char* name) -> const char* {
      if (a + 1 >= argc) {
        fprintf(stderr, "nwsolve: option %s requires a value\n", name);
        exit(2);
      }
      return argv[++a];
    };
    if (s == "-d" || s == "--maxdepth") o.maxdepth = atoi(val("-d"));
    else if (s == "-u" || s == "--maxdup") o.maxdup = atoi(val("-u"));
    else if (s == "-w" || s == "--maxwidth") o.maxwidth = (int64_t)strtod(val("-w"), nullptr);
    else if (s == "-O" || s == "--maxoffset") (void)val("-O");
    else if (s == "-r" || s == "--minreport") o.minreport = strtod(val("-r"), nullptr);
    else if (s == "-R" || s == "--maxreport") o.maxreport = (int64_t)strtod(val("-R"), nullptr);
    else if (s == "--device") device = atoi(val("--device"));
    else if (s == "--devices") {
      devices = parse_devices(val("--devices"));
      if (devices.empty()) {
        fprintf(stderr, "nwsolve: --devices needs a comma-separated list of GPU indices or 'all'\n");
        return 2;
      }
    }
    else if (s == "--stats") stats = true;
    else if (s == "--batch") batch = val("--batch");
    else if (s == "--contexts") contexts = atoi(val("--contexts"));
    else if (s == "--regions-per-pass") per_pass = atoi(val("--regions-per-pass"));
    else if (s == "-h" || s == "--help") {
      usage();
      return 0;
    } else if (s.size() > 1 && s[0] == '-' && !(s[1] >= '0' && s[1] <= '9')) {
      fprintf(stderr, "nwsolve: unknown option %s\n", s.c_str());
      usage();
      return 2;
    } else pos.push_back(argv[a]);
  }
  if (batch) {
    if (!pos.empty() || contexts < 1 || contexts > 64) {
      usage();
      return 2;
    }
    if (devices.empty()) devices.push_back(device);
    return run_batch(batch, o, devices, contexts, per_pass, stats);
  }
  if (pos.size() != 3) {
    usage();
    return 2;
  }
  if (devices.size() > 1) {  // one search over several GPUs
    std::vector<nw_ctx*> ctxs;
    for (int dv : devices) {
      nw_ctx* x = nullptr;
      if (nw_ctx_create(dv, &x) != NW_OK) {
        fprintf(stderr, "nwsolve: %s\n", nw_last_error());
        for (nw_ctx* y : ctxs) nw_ctx_destroy(y);
        return 1;
      }
      ctxs.push_back(x);
    }
    nw_stats st;
    const int rc = nw_solve_files_sharded(ctxs.data(), (int32_t)ctxs.size(), pos[0], pos[1], pos[2], &o, &st);
    if (rc != NW_OK) fprintf(stderr, "nwsolve: %s\n", nw_last_error());
    else if (stats)
      fprintf(stderr, "nwsolve: levels=%d best=%.3f device_ms=%.3f ranks=%zu\n", st.n_levels, st.best, st.total_ms, ctxs.size());
    for (nw_ctx* y : ctxs) nw_ctx_destroy(y);
    return rc == NW_OK ? 0 : 1;
  }
  if (devices.size() == 1) device = devices[0];
  nw_ctx* ctx = nullptr;
  int rc = nw_ctx_create(device, &ctx);
  if (rc != NW_OK) {
    fprintf(stderr, "nwsolve: %s\n", nw_last_error());
    return 1;
  }
  nw_stats st;
  rc = nw_solve_files(ctx, pos[0], pos[1], pos[2], &o, &st);
  if (rc != NW_OK) {
    fprintf(stderr, "nwsolve: %s\n", nw_last_error());
    nw_ctx_destroy(ctx);
    return 1;
  }
  if (stats) {
    long long gen = 0, sc = 0;
    for (int l = 0; l < st.n_levels; ++l) {
      gen += st.generated[l];
      sc += st.scored[l];
    }
    fprintf(stderr, "nwsolve: levels=%d generated=%lld scored=%lld best=%.3f device_ms=%.3f kernels=%d\n", st.n_levels,
            gen, sc, st.best, st.total_ms, st.kernel_launches);
  }
  nw_ctx_destroy(ctx);
  return 0;
}
