#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics + the hottest source lines.
usage: tools/ncu_summary.py gpurun_out/x.ncu-rep [n_lines]"""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_xu.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        print("== kernel:", name[:120])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-78s %s %s" % (k, vals[i], units[i]))
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    # find the header row
    h = next(i for i, r in enumerate(rows) if "Source" in r and any(c.startswith("Warp Stall Sampling (All") for c in r))
    hdr = rows[h]
    i_src = hdr.index("Source"); i_s = hdr.index("Warp Stall Sampling (All Samples)"); i_ex = hdr.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    data = []
    for r in rows[h + 1:]:
        try:
            data.append((int(r[i_s]), r))
        except (ValueError, IndexError):
            pass
    tot = sum(d[0] for d in data) or 1
    print("== total stall samples", tot, " instructions executed", sum(int(d[1][i_ex] or 0) for d in data))
    agg = {}
    for s, r in data:
        for i in stall_cols:
            v = int(r[i] or 0)
            if v:
                agg[hdr[i]] = agg.get(hdr[i], 0) + v
    print("== stall reasons:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    top = sorted(range(len(data)), key=lambda i: -data[i][0])[:nlines]
    for idx in sorted(top):
        s, r = data[idx]
        st = sorted([(int(r[i] or 0), hdr[i]) for i in stall_cols], reverse=True)[:2]
        print("  %5d %5.1f%% exec=%-9s %-70s %s" % (idx, 100.0 * s / tot, r[i_ex], r[i_src][:70], " ".join("%s:%d" % (b[6:], a) for a, b in st if a)))


if __name__ == "__main__":
    main()
