#!/usr/bin/env python
"""Runs a few searches of a bench workload and prints per-level device times / counters (used under ncu for the
launch list committed in profiles/).  usage: python tools/profile_solve.py [s64|s150] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nahrwhals_b200 as nb  # noqa: E402
from bench import make_workload  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "s64"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
aln, lx, ly, flags, desc = make_workload(name, 1)
if os.environ.get("NW_PROFILE_NO_VERIFY"):   # fingerprint-only dedup: what the element-wise verifier costs
    flags = dict(flags, verify_dedup=False)
if os.environ.get("NW_PROFILE_DP32"):        # 32-bit register DP everywhere: what the packed 16-bit form buys
    flags = dict(flags, dp32=True)
s = nb.Solver(0)
for r in range(reps):
    t0 = time.perf_counter()
    R = s.solve(aln, lx, ly, **flags)
    wall = time.perf_counter() - t0
    st = R.stats
    nl = st.n_levels
    print("rep %d: wall %.2f ms, device %.2f ms, DP kernels %.3f ms in %d launches, %d kernel launches, h2d %d B, d2h %d B"
          % (r, wall * 1e3, st.total_ms, st.score_ms, st.score_launches, st.kernel_launches, st.h2d_bytes, st.d2h_bytes))
    for l in range(nl):
        print("   level %d: frontier %d generated %d scored %d cells %d  %.3f ms" %
              (l + 1, st.frontier[l], st.generated[l], st.scored[l], st.cells[l], st.level_ms[l]))
    print("   best %.3f, %d trajectories; host ms: setup %.2f search %.2f report %.2f dump %.2f" %
          (st.best, len(R.scores), st.host_ms[0], st.host_ms[1], st.host_ms[2], st.host_ms[3]))
s.close()
