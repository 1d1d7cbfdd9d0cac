python - <<'P'
import sys, os, time
sys.path.insert(0, os.getcwd())
from tools.region_batch import make_regions, run_many
regs = make_regions(4000)
for per, nctx in ((512, 1), (512, 2), (512, 4), (1024, 1)):
    r, sc, dt = run_many(regs, per, nctx=nctx)
    print("GPU, %d ctx, %d per pass: %.1f regions/s, %.3g cand/s (%.3f s)" % (nctx, per, r, sc / dt, dt), flush=True)
P
nproc
