#!/usr/bin/env python
"""Latency anatomy of one small region (regionfile mode): wall / device / host phases / launches."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import nahrwhals_b200 as nb
from region_batch import make_regions, FLAGS
regs = make_regions(40)
s = nb.Solver(0)
for r in regs[:5]:
    s.solve(*r, **FLAGS)
tot = 0
for k, r in enumerate(regs[5:25]):
    t0 = time.perf_counter(); R = s.solve(*r, **FLAGS); w = time.perf_counter() - t0
    st = R.stats; tot += w
    print("n_x %3d n_y %3d: wall %.2f ms device %.2f ms host[setup %.2f search %.2f report %.2f] launches %d gen %s scored %s" %
          (r[0].shape[0], r[0].shape[1], w * 1e3, st.total_ms, st.host_ms[0], st.host_ms[1], st.host_ms[2], st.kernel_launches, sum(R.generated), sum(R.scored)))
print("mean wall %.2f ms" % (tot / 20 * 1e3))
