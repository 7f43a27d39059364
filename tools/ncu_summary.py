#!/usr/bin/env python
"""tools/ncu_summary.py <report.ncu-rep> -- the per-launch metrics of an `ncu --set full` capture that DESIGN.md quotes
(durations, DRAM bytes, issue utilisation, occupancy, stall ratios), one column per captured launch, plus the pass totals."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out)); hdr, units, data = rows[0], rows[1], rows[2:]
WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__waves_per_multiprocessor", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__sectors_read.sum", "dram__sectors_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_membar"]
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w} [{units[i]}]: {[r[i][:26] for r in data]}")
col = lambda n: [float(r[hdr.index(n)].replace(',', '')) for r in data]
names = [r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "") for r in data]
dur, rd, wr = col("gpu__time_duration.sum"), col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
first = {}
for k, n in enumerate(names):
    if n in first: break
    first[n] = k
tot = sum(dur[k] for k in first.values()); traffic = sum(rd[k] * scale[ur] + wr[k] * scale[uw] for k in first.values())
print(f"\nfirst pass totals: {tot:.1f} us (serialised under ncu), DRAM traffic {traffic / 1e6:.1f} MB")
print("shares of the serialised pass: " + ", ".join(f"{n} {100 * dur[k] / tot:.0f}%" for n, k in first.items()))
for n, k in first.items():
    b = rd[k] * scale[ur] + wr[k] * scale[uw]
    print(f"  {n}: {b / 1e6:.1f} MB in {dur[k]:.1f} us = {b / dur[k] / 1e3:.0f} GB/s")
