import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nahrwhals_b200 as nb
from tools.region_batch import make_regions, FLAGS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
per = int(sys.argv[2]) if len(sys.argv) > 2 else 512
regs = make_regions(n)
s = nb.Solver(0)
nb.solve_many(s, regs[:per], regions_per_pass=per, **FLAGS)
t0 = time.perf_counter()
res = nb.solve_many(s, regs, regions_per_pass=per, **FLAGS)
dt = time.perf_counter() - t0
st = res[0].stats
print("wall %.1f ms; first pass device %.2f ms, levels %s, launches %d, DP %.2f ms in %d launches" % (
    dt * 1e3, st.total_ms, ["%.2f" % v for v in list(st.level_ms)[:st.n_levels]], st.kernel_launches, st.score_ms, st.score_launches))
print("scored total", sum(sum(R.scored) for R in res))
