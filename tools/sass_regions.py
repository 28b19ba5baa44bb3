#!/usr/bin/env python
"""Stall samples and executed instructions by CUDA source line range (developer).  usage: tools/sass_regions.py rep file.cu a-b a-b ..."""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; fname = sys.argv[2]
ranges = [tuple(int(v) for v in r.split("-")) for r in sys.argv[3:]]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if "Source" in r)
hdr = rows[h]
i_ex = hdr.index("Instructions Executed"); i_s = hdr.index("Warp Stall Sampling (All Samples)")
src = open(fname).read().splitlines()
agg = collections.OrderedDict((r, [0, 0]) for r in ranges); other = [0, 0]
cur = None
for r in rows[h + 1:]:
    if len(r) <= i_ex: continue
    if r[0]:
        ln = int(r[0]); txt = r[1].strip()
        # only lines of the named file: match the text
        cur = ln if (0 < ln <= len(src) and src[ln - 1].strip()[:40] == txt[:40]) else -1
    elif r[3] and r[3] != "-":
        e = int(r[i_ex]) if r[i_ex].isdigit() else 0
        sm = int(r[i_s]) if r[i_s].isdigit() else 0
        tgt = other
        if cur and cur > 0:
            for rg in ranges:
                if rg[0] <= cur <= rg[1]: tgt = agg[rg]; break
        tgt[0] += e; tgt[1] += sm
tot_e = sum(v[0] for v in agg.values()) + other[0]; tot_s = sum(v[1] for v in agg.values()) + other[1]
for rg, v in agg.items():
    print("lines %4d-%4d  exec %9d (%4.1f%%)  samples %6d (%4.1f%%)" % (rg[0], rg[1], v[0], 100.0 * v[0] / tot_e, v[1], 100.0 * v[1] / tot_s))
print("other (inlined helpers / other files)  exec %9d (%4.1f%%)  samples %6d (%4.1f%%)" % (other[0], 100.0 * other[0] / tot_e, other[1], 100.0 * other[1] / tot_s))
