#!/usr/bin/env python
"""Turn one tools/profile_round.sh run (files under gpurun_out/) into the tracked summaries under profiles/.
usage: tools/make_profiles.py <suffix of bench/launch files, e.g. r2b> <grad rep> <fcc rep> <fac rep> [round tag, default r2]"""
import collections, csv, io, json, re, subprocess, sys

suffix, grad_rep, fcc_rep, fac_rep = sys.argv[1:5]
RT = sys.argv[5] if len(sys.argv) > 5 else "r2"
G = "gpurun_out/"
for rep, out in ((grad_rep, "grad_kernel"), (fcc_rep, "fcc_scan"), (fac_rep, "fac_scan")):
    txt = subprocess.run([sys.executable, "tools/ncu_summary.py", G + rep, "45" if out == "grad_kernel" else "30"],
                         capture_output=True, text=True).stdout
    open("profiles/%s_%s_ncu_summary.txt" % (out, RT), "w").write(txt)

raw = subprocess.run(["ncu", "-i", G + grad_rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, vals = rows[0], rows[1], rows[2]
def g(k):
    i = hdr.index(k)
    return int(round(float(vals[i].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]))
r, w = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
json.dump({"kernel": "grad_kernel<float,40,1,1>", "source": "ncu --set full --clock-control none, %s%s (round %s)" % (G, grad_rep, RT[1:]),
           "dram_bytes_read": r, "dram_bytes_write": w, "dram_bytes_per_launch": r + w, "algorithmic_bytes": 143475200},
          open("profiles/grad_kernel_traffic.json", "w"), indent=1)

txt = open(G + "launches_%s.csv" % suffix).read()
txt = txt[txt.index('"ID"'):]
open("profiles/launches_%s.csv" % RT, "w").write(txt)
agg = collections.OrderedDict()
for x in csv.DictReader(io.StringIO(txt)):
    short = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "")
    if short.startswith("at::"):
        short = "torch glue: " + short.split("<")[0]
    a = agg.setdefault((short, x["Grid Size"], x["Block Size"]), [0, 0.0])
    a[0] += 1; a[1] += float(x["Metric Value"]) / 1e3
tot = sum(a[1] for a in agg.values())
b = json.loads(open(G + "bench_%s.json" % suffix).read().strip().splitlines()[-1])
json.dump(b, open("profiles/bench_%s_1gpu.json" % RT, "w"))
L = ["# ncu launch list, round %s (profiles/launches_%s.md)\n" % (RT[1:], RT),
     "Command (B200, via gpurun, `tools/profile_round.sh`): `ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --api eager-ws` (the workspace route issued eagerly: the same kernels the step graph replays, visible to ncu one by one)",
     "60 consecutive launches (~6.5 steps of cfg-2: T=1000, N=64, L=40, S=200, fp32, fwd+bwd); raw list: `profiles/launches_%s.csv`." % RT + "  Per-launch times under ncu are cold-cache and serialised;",
     "what must agree with bench.py's live CUDA-event numbers is each kernel's SHARE of the step, not the absolute.\n",
     "| kernel | launches | grid | block | total us | avg us | share of listed time |", "|---|---|---|---|---|---|---|"]
ncu_avg = {}
for (name, grid, block), a in agg.items():
    L.append("| `%s` | %d | %s | %s | %.1f | %.1f | %.1f%% |" % (name, a[0], grid, block, a[1], a[1] / a[0], 100 * a[1] / tot))
    for key in ("grad_kernel", "fcc_scan_kernel", "fac_scan_kernel", "grad_finalize_kernel"):
        if key in name:
            ncu_avg[key] = a[1] / a[0]
L.append("")
L.append("bench.py (same build, not under a profiler, CUDA events on the launching streams; `profiles/bench_%s_1gpu.json`):\n" % RT)
k = b["kernels"]
for n, v in k.items():
    L.append("* `%s`: %.1f us average over %d launches%s" % (n, v["avg_us"], v["launches"],
             " (separate pass: event pairs on the two scan streams would perturb the timed region)" if v.get("note") == "separate pass"
             else " (%s)" % v.get("note", "inside the timed region")))
live = sum(v["avg_us"] for v in k.values()); nt = sum(ncu_avg.values())
L.append("")
L.append("Shares of the four asg kernels, live vs ncu: " + ", ".join("%s %.1f%% vs %.1f%%" % (n.replace("_kernel", ""), 100 * k[n]["avg_us"] / live, 100 * ncu_avg[n] / nt) for n in k if n in ncu_avg) + ".")
L.append("")
L.append("The FCC and FAC scan kernels run concurrently on two streams, so a step's device time is max(fcc, fac) + grad + finalize (+ ~5 tiny torch glue kernels of the autograd surface).")
L.append("ms_per_step = %.4f (host issue time %.3f ms), value = %.0f utterances/s, e2e = %.0f utterances/s (8 host contexts in flight), cpu_baseline (compiled reference, %d cores) = %.1f utterances/s."
         % (b["ms_per_step"], b.get("host_issue_ms_per_step", 0), b["value"], b["e2e"]["value"], b["cpu_baseline"]["cores"], b["cpu_baseline"]["value"]))
L.append("roofline (grad_kernel): %.0f GB/s of %.1f measured peak = %.3f; DRAM traffic per launch %d B vs %d algorithmic."
         % (b["roofline"]["achieved"], b["roofline"]["peak"], b["roofline"]["frac"], r + w, 143475200))
open("profiles/launches_%s.md" % RT, "w").write("\n".join(L) + "\n")
print("\n".join(L[-6:]))
