#!/usr/bin/env python
"""Group the SASS instructions of a kernel by their execution count (= by loop nest / code path) (developer).
usage: tools/sass_bycount.py rep [n]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if "Source" in r)
hdr = rows[h]
i_ex = hdr.index("Instructions Executed"); i_s = hdr.index("Warp Stall Sampling (All Samples)")
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in rows[h + 1:]:
    if len(r) > i_ex and r[i_ex].isdigit():
        e = int(r[i_ex]); a = agg[e]; a[0] += 1; a[1] += e; a[2] += int(r[i_s] or 0)
tot_e = sum(a[1] for a in agg.values()); tot_s = sum(a[2] for a in agg.values())
print("exec/instr   #instr   total exec        samples")
for e, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:n]:
    print("%9d  %7d  %10d %5.1f%%  %6d %5.1f%%" % (e, a[0], a[1], 100.0 * a[1] / tot_e, a[2], 100.0 * a[2] / tot_s))
