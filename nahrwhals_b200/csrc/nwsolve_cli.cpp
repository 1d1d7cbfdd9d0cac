// nwsolve -- command-line twin of `julia solver.jl` (argument table: solver.jl:756-791 in the reference).
//   nwsolve [-d D] [-u U] [-w W] [-O O] [-r r] [-R R] compressed_alignment compressed_lengths out_path
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
          "               compressed_alignment compressed_lengths out_path\n");
}

int main(int argc, char** argv) {
  nw_opts o;
  nw_default_opts(&o);
  int device = -1;
  bool stats = false;
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
    else if (s == "--stats") stats = true;
    else if (s == "-h" || s == "--help") {
      usage();
      return 0;
    } else if (s.size() > 1 && s[0] == '-' && !(s[1] >= '0' && s[1] <= '9')) {
      fprintf(stderr, "nwsolve: unknown option %s\n", s.c_str());
      usage();
      return 2;
    } else pos.push_back(argv[a]);
  }
  if (pos.size() != 3) {
    usage();
    return 2;
  }
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
