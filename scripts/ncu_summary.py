#!/usr/bin/env python
"""Key metrics of an .ncu-rep (raw page) + the hottest source lines (source page)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
for r in rows[2:]:
    print("----")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("%-90s %s %s" % (w, r[i][:100], units[i]))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur_file, h = "", None
    lines = {}   # (file, line, text) -> [samples, instructions]
    tot = 0.0
    for r in csv.reader(src.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            h = r
            ci = h.index("# Samples")
            ii = h.index("Instructions Executed")
            continue
        if h is None or len(r) <= ci:
            continue
        if r[0] != "":      # a CUDA source line row (aggregated over its SASS)
            try:
                v = float(r[ci]); ins = float(r[ii])
            except ValueError:
                continue
            key = (cur_file, r[0], r[1].strip())
            e = lines.setdefault(key, [0.0, 0.0])
            e[0] += v; e[1] += ins
            tot += v
    print("hottest source lines (# warp-stall samples; warp instructions executed):")
    for (f, ln, text), (v, ins) in sorted(lines.items(), key=lambda x: -x[1][0])[:int(sys.argv[2])]:
        print("%5.1f%%  %10.0f inst  %s:%s  %s" % (100 * v / (tot or 1), ins, f, ln, text[:110]))
