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
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    # find header row
    for hi, r in enumerate(rows):
        if "Source" in r and any("Sampl" in c for c in r):
            break
    h = rows[hi]
    si = h.index("Source")
    ci = [i for i, c in enumerate(h) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or "Samples" in c][0]
    body = [r for r in rows[hi + 1:] if len(r) > ci and r[ci].replace(".", "").isdigit()]
    tot = sum(float(r[ci]) for r in body) or 1
    body.sort(key=lambda r: -float(r[ci]))
    print("hottest source lines (%s):" % h[ci])
    for r in body[:int(sys.argv[2])]:
        print("%6.1f%%  %s" % (100 * float(r[ci]) / tot, r[si].strip()[:140]))
