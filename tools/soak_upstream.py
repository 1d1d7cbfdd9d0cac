#!/usr/bin/env python
"""Soak test of the upstream steps (PAF preparation batch, wrapper_paf_to_bitlocus loop, PAF file mode, the one-call
regionfile pipeline) against the R restatements on random inputs.  usage: python tools/soak_upstream.py [rounds] [seed]"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nahrwhals_b200 as nb  # noqa: E402
from oracle import nw_bitlocus_oracle as bo, nw_paf_oracle as po, oracle_lib  # noqa: E402
from tests.test_bitlocus import sd_paf  # noqa: E402
from tests.test_paf import chunked_paf, write_chunk_paf  # noqa: E402
from tests.test_pipeline import oracle_row  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
s = nb.Solver(0)
bad = 0
t0 = time.time()
tmp = tempfile.mkdtemp()
for rnd in range(rounds):
    rng = np.random.default_rng(1000 * seed0 + rnd)
    tabs = [chunked_paf(rng, n_segments=int(rng.integers(1, 30)), chunk=int(rng.choice([2000, 5000, 10000])),
                        jitter=int(rng.choice([0, 0, 2, 30, 400])), span=int(rng.choice([40000, 300000, 900000, 3000000])),
                        drop=float(rng.choice([0.0, 0.1, 0.3])), extra=int(rng.integers(0, 15))) for _ in range(30)]
    tabs += [sd_paf(int(rng.integers(0, 10 ** 6)), n_blocks=int(rng.integers(8, 40))) for _ in range(30)]
    # 1. batched preparation
    minlen, comp = float(rng.choice([200, 1000, 5000])), float(rng.choice([100, 1000, 10000]))
    got, rcs = nb.load_and_prep_pafs(s, tabs, minlen, comp, 10000.0)
    for k, t in enumerate(tabs):
        want = po.load_and_prep_paf(t, minlen, comp, 10000.0)
        if any(not np.array_equal(got[k][c], want[c]) for c in po.COLS):
            bad += 1
            print("MISMATCH prepare round %d table %d" % (rnd, k))
    # 2. the bitlocus loop with random limits
    kw = dict(max_n_alns=int(rng.choice([5, 20, 150])), max_size_col_plus_rows=int(rng.choice([16, 40, 250])),
              max_pairs=int(rng.choice([2, 50, 2000])))
    B = nb.paf_to_bitlocus(s, tabs, minlen, comp, 10000.0, **kw)
    for k, t in enumerate(tabs):
        O = bo.wrapper_paf_to_bitlocus(t, minlen, comp, 10000.0, **kw)
        st = {"ok": "ok", "cluttered": "cluttered", "empty": "failed", "passes": "failed"}[O["status"]]
        same = (B[k].status == st and B[k].passes == O["passes"] and B[k].minlen == O["minlen"] and B[k].n_alns == O["n_alns"])
        if same and st == "ok":
            same = (np.array_equal(B[k].grid.mat, O["grid"]["mat"]) and np.array_equal(B[k].grid.gridlines_x, O["grid"]["gridlines_x"])
                    and np.array_equal(B[k].grid.gridlines_y, O["grid"]["gridlines_y"]) and B[k].n_pairs == O["n_pairs"]
                    and all(np.array_equal(B[k].paf[c], O["paf"][c]) for c in po.COLS))
        if not same:
            bad += 1
            print("MISMATCH bitlocus round %d table %d %s %s" % (rnd, k, B[k].status, O["status"]))
    # 3. file mode
    for k in range(0, len(tabs), 3):
        fin, fo, fg = (os.path.join(tmp, n) for n in ("in.paf", "o.paf", "g.paf"))
        write_chunk_paf(fin, tabs[k], tags=bool(k % 2), dup_lines=k % 3)
        chunk = float(tabs[k]["qlen"][0]) if len(tabs[k]["qlen"]) else 10000.0
        a = po.compress_paf_file(fin, fo, chunk)
        b = nb.compress_paf_file(s, fin, fg, chunk)
        if a != b or open(fo).read() != open(fg).read():
            bad += 1
            print("MISMATCH file mode round %d table %d" % (rnd, k))
    # 4. the one-call pipeline on the SD regions
    sub = tabs[30:]
    metas = [dict(chr="chr%d" % (1 + k % 22), start=str(50000 * k), end=str(50000 * k + 700000), samplename="a_b") for k in range(len(sub))]
    depth = int(rng.choice([1, 2, 3]))
    rows, st = nb.run_regions_native(s, sub, metas, 1000.0, 1000.0, 10000.0, depth=depth)
    for k in range(len(sub)):
        want = oracle_row(oracle_lib, sub[k], metas[k], depth, 1000.0, 1000.0, 10000.0)
        if rows[k] != want:
            bad += 1
            print("MISMATCH pipeline round %d region %d\n  got  %s\n  want %s" % (rnd, k, rows[k], want))
    print("round %d done, %d mismatches so far, %.1f s" % (rnd, bad, time.time() - t0), flush=True)
s.close()
print("soak_upstream: %d rounds, %d mismatches" % (rounds, bad))
sys.exit(1 if bad else 0)
