#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + SASS opcode mix + top stall lines) for profiles/."""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
for vals in rows[2:]:
    for h, u, v in zip(hdr, units, vals):
        if h in keys:
            print(f"{h:88s} {v:>22s} {u}")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                      capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
h = rows[1]
si, ie, ss = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
ops, samp = collections.Counter(), collections.Counter()
tot = tots = 0
for r in rows[2:]:
    try:
        n, s = int(r[ie]), int(r[ss])
    except Exception:
        continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[si])
    op = m.group(2).split('.')[0] if m else '?'
    ops[op] += n
    samp[op] += s
    tot += n
    tots += s
print("\nSASS opcode mix (warp instructions executed; share of stall samples)")
print("total", tot, "static instructions", len(rows) - 2)
for op, n in ops.most_common(22):
    print(f"{op:12s} {n:14d} {n / tot:.3f}  samples {samp[op] / max(tots, 1):.3f}")
