#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel (nvdisasm -g line info).
usage: sass_lines.py lib.so kernel_substring [file_substring]"""
import collections, glob, os, re, subprocess, sys, tempfile
lib, kname = sys.argv[1:3]
fsub = sys.argv[3] if len(sys.argv) > 3 else ""
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    if kname not in txt:
        continue
    cur, infn = None, False
    cnt = collections.Counter(); ops = collections.defaultdict(collections.Counter)
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            infn = kname in m.group(1); continue
        if not infn: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            cnt[cur] += 1; ops[cur][m.group(1).split('.')[0]] += 1
    tot = sum(cnt.values())
    print("static instructions", tot)
    for k in sorted(cnt):
        if fsub in k[0]:
            print("%-20s:%4d %4d  %s" % (k[0], k[1], cnt[k], " ".join("%s:%d" % kv for kv in ops[k].most_common(6))))
    break
