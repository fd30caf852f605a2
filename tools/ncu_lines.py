#!/usr/bin/env python
"""Attribute an .ncu-rep's per-SASS-instruction counters to CUDA source lines.

usage: ncu_lines.py report.ncu-rep lib.so kernel_substring [min_share]
The CLI's CUDA source page carries no metrics, so the SASS listing of the report is matched
instruction-by-instruction with `nvdisasm -g` line info of the cubin embedded in the library.
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kname = sys.argv[1:4]
min_share = float(sys.argv[4]) if len(sys.argv) > 4 else 0.004
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines = None
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    if kname not in txt:
        continue
    cur, infn, out = None, False, []
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            infn = kname in m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            out.append(cur)
    if out:
        lines = out
        break
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                      capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
h = rows[1]
ie, ss, si = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
ig, iw = h.index("L1 Tag Requests Global"), h.index("L1 Wavefronts Shared")
body = [r for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
if lines is None or len(lines) != len(body):
    print("warning: %s SASS rows in report vs %s in cubin" % (len(body), None if lines is None else len(lines)))
n = min(len(body), len(lines or []))
inst, samp, lg, lw = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
num = lambda v: int(v) if v.strip().isdigit() else 0
for r, l in zip(body[:n], lines[:n]):
    inst[l] += int(r[ie])
    samp[l] += int(r[ss])
    lg[l] += num(r[ig])
    lw[l] += num(r[iw])
ti, ts = sum(inst.values()), max(sum(samp.values()), 1)
srcs = {}
tg, tw = max(sum(lg.values()), 1), max(sum(lw.values()), 1)
print("total warp instructions %d, samples %d, L1 tag requests global %d, L1 wavefronts shared %d" % (ti, ts, tg, tw))
for l in sorted(inst, key=lambda k: (k is None, k)):
    if l is None or (inst[l] / ti < min_share and samp[l] / ts < min_share and lg[l] / tg < min_share
                     and lw[l] / tw < min_share):
        continue
    f, no = l
    if f not in srcs:
        cand = glob.glob(os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f))
        srcs[f] = open(cand[0]).read().splitlines() if cand else []
    text = srcs[f][no - 1].strip()[:90] if no - 1 < len(srcs[f]) else ""
    print("%-18s:%4d inst %.3f samp %.3f gl1 %.3f shw %.3f | %s" % (f, no, inst[l] / ti, samp[l] / ts, lg[l] / tg,
                                                                     lw[l] / tw, text))
