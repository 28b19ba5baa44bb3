#!/usr/bin/env python
"""Instruction mix of the hot part of a kernel from an .ncu-rep (developer): executed warp-instructions by opcode, for the SASS
lines whose execution count is at least `frac` of the most executed one.  usage: tools/sass_hist.py rep [frac]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if "Source" in r)
hdr = rows[h]
i_src = hdr.index("Source"); i_ex = hdr.index("Instructions Executed")
data = [r for r in rows[h + 1:] if len(r) > i_ex and r[i_ex].isdigit()]
mx = max(int(r[i_ex]) for r in data)
tot = sum(int(r[i_ex]) for r in data)
ops = collections.Counter(); hot = 0
for r in data:
    e = int(r[i_ex])
    s = r[i_src].strip()
    if s.startswith("@"):
        s = s.split(None, 1)[1]
    op = s.split()[0].split(".")[0] if s else "?"
    ops[op] += e
print("total executed", tot, "max per line", mx)
for op, e in ops.most_common(40):
    print("%-12s %10d %5.1f%%" % (op, e, 100.0 * e / tot))
