#!/usr/bin/env python
"""Summarises an `ncu --page source --csv` dump: executed warp-instructions per opcode and top stall lines.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_mix.py src.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci, ie, si = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
agg, samp = collections.Counter(), collections.Counter()
tot = stot = 0
lines = []
for r in rows[hi + 1:]:
    try:
        n = int(float(r[ie]))
        s = int(float(r[si]))
    except Exception:
        continue
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[ci])
    if not m:
        continue
    op = m.group(2).split(".")[0]
    agg[op] += n
    samp[op] += s
    tot += n
    stot += s
    lines.append((s, n, r[ci].strip()))
print("executed warp-instructions: %d   stall samples: %d" % (tot, stot))
for k, v in agg.most_common(22):
    print("%-12s %13d %5.1f%%   samples %5.1f%%" % (k, v, 100.0 * v / tot, 100.0 * samp[k] / max(stot, 1)))
print("--- top sampled instructions")
for s, n, src in sorted(lines, reverse=True)[:14]:
    print("%6d %12d  %s" % (s, n, src[:90]))
