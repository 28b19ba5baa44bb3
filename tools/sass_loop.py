#!/usr/bin/env python
"""Print the SASS of the hottest loop of an .ncu-rep: instructions whose execution count lies in [lo, hi],
with stall samples.  usage: tools/sass_loop.py rep lo hi [start_index] [max_lines]"""
import csv, io, subprocess, sys, collections
rep, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
start = int(sys.argv[4]) if len(sys.argv) > 4 else 0
maxl = int(sys.argv[5]) if len(sys.argv) > 5 else 400
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if "Source" in r)
hdr = rows[h]
i_src = hdr.index("Source"); i_ex = hdr.index("Instructions Executed"); i_s = hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
data = [r for r in rows[h + 1:] if len(r) > i_ex and r[i_ex].isdigit()]
tot = sum(int(r[i_s] or 0) for r in data)
op = collections.Counter(); n = 0; samp = 0
for i, r in enumerate(data):
    ex = int(r[i_ex])
    if i < start or not (lo <= ex <= hi):
        continue
    n += 1
    if n > maxl:
        break
    s = int(r[i_s] or 0); samp += s
    m = r[i_src].split(); name = m[1] if m[0].startswith("@") else m[0]
    op[name.split(".")[0]] += 1
    st = sorted([(int(r[k] or 0), hdr[k][6:]) for k in stall_cols], reverse=True)[:1]
    print("%5d %8s %5d  %-64s %s" % (i, r[i_ex], s, r[i_src][:64], (st[0][1] + ":" + str(st[0][0])) if st and st[0][0] else ""))
print("lines", n, "samples", samp, "of", tot, "mix", op.most_common(14))
