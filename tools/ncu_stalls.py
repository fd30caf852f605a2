#!/usr/bin/env python
"""Per-SASS-instruction stall attribution of an .ncu-rep (source page, --import-source on).
usage: ncu_stalls.py report.ncu-rep [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
print(rows[0][1])
hdr, data = rows[1], rows[2:]
col = hdr.index
S, I = col('# Samples'), col('Instructions Executed')
tot, toti = sum(int(r[S]) for r in data), sum(int(r[I]) for r in data)
kinds = ['stall_long_sb', 'stall_wait', 'stall_short_sb', 'stall_selected', 'stall_not_selected', 'stall_branch_resolving',
         'stall_no_inst', 'stall_barrier', 'stall_math', 'stall_mio', 'stall_lg', 'stall_dispatch']
print("samples %d, warp instructions %d, static instructions %d" % (tot, toti, len(data)))
print("share of samples by reason: " + ", ".join("%s %.1f%%" % (k[6:], 100 * sum(int(r[col(k)]) for r in data) / tot)
                                                  for k in kinds))
for k in kinds[:3]:
    c = col(k)
    tk = sum(int(r[c]) for r in data)
    print("\ntop instructions by %s (%.1f%% of all samples): share of this reason / of executed instructions" %
          (k, 100 * tk / tot))
    for i in sorted(range(len(data)), key=lambda i: -int(data[i][c]))[:top]:
        r = data[i]
        print("  #%-5d %-62s %5.1f%%  %5.2f%%" % (i, r[1].strip()[:62], 100 * int(r[c]) / tk, 100 * int(r[I]) / toti))
