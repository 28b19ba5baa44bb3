#!/usr/bin/env python
"""Per CUDA-source-line executed-instruction totals with the opcode mix (developer).  usage: tools/sass_lines.py rep [n]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if "Source" in r)
hdr = rows[h]
i_ex = hdr.index("Instructions Executed")
lines = []          # (lineno, text, total, Counter)
cur = None
for r in rows[h + 1:]:
    if len(r) <= i_ex: continue
    if r[0]:
        cur = [r[0], r[1], 0, collections.Counter()]
        lines.append(cur)
    elif cur is not None and r[3] and r[3] != "-":
        s = r[3].strip()
        if s.startswith("@"): s = s.split(None, 1)[1]
        op = s.split()[0].split(".")[0]
        e = int(r[i_ex]) if r[i_ex].isdigit() else 0
        cur[2] += e; cur[3][op] += e
tot = sum(l[2] for l in lines)
print("total", tot)
for l in sorted(lines, key=lambda l: -l[2])[:n]:
    print("%5s %8d %4.1f%%  %-60s %s" % (l[0], l[2], 100.0 * l[2] / tot, l[1].strip()[:60], " ".join("%s:%d" % (k, v * 1000 // max(l[2], 1)) for k, v in l[3].most_common(5))))
