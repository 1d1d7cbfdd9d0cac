#!/usr/bin/env python
"""Phase timing of the region stage (nw_region_analyse_many): NW_DEBUG_TIMING=1 python tools/stage_profile.py [n]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import nahrwhals_b200 as nb  # noqa: E402
from region_batch import make_stage_regions  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
regs = make_stage_regions(n, seed=0)
s = nb.Solver(0)
kw = dict(depth=3, maxdup=2, init_width=1000, minreport=0.98)
for it in range(3):
    tm = {}
    t0 = time.perf_counter()
    res = nb.analyse_regions(s, regs, timing=tm, **kw)
    print("pass %d: native %.1f ms, python total %.1f ms, %.0f regions/s native" % (
        it, tm["native_s"] * 1e3, (time.perf_counter() - t0) * 1e3, n / tm["native_s"]), flush=True)
s.close()
