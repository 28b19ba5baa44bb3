#!/usr/bin/env python
"""Top-sampled SASS lines of an .ncu-rep (developer).  usage: tools/sass_top.py rep [n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if "Source" in r)
hdr = rows[h]
i_src = hdr.index("Source"); i_ex = hdr.index("Instructions Executed"); i_s = hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
data = [(i, r) for i, r in enumerate(rows[h + 1:]) if len(r) > i_ex and r[i_ex].isdigit()]
tot = sum(int(r[i_s] or 0) for _, r in data)
top = sorted(data, key=lambda ir: -int(ir[1][i_s] or 0))[:n]
for i, r in sorted(top):
    st = sorted([(int(r[k] or 0), hdr[k][6:]) for k in stall_cols], reverse=True)[:2]
    print("%5d %9s %5s  %-70s %s" % (i, r[i_ex], r[i_s], r[i_src][:70], " ".join("%s:%d" % (b, a) for a, b in st if a)))
print("total samples", tot)
