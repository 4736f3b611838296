#!/usr/bin/env python
"""Per-source-line stall samples of one kernel: joins an ncu report's SASS page with nvdisasm line info.

usage: ncu_lines.py <report.ncu-rep> <libmodelardb_cuda.so> <kernel substring> [top N] [--outer FILE] [--by-line]

--outer FILE   attribute every instruction to the OUTERMOST inline frame that lies in FILE (e.g. mdb_fit_warp.cuh),
               so that inlined helpers (ddiv_fast, shuffles, keep_min ...) are charged to their call sites
--by-line      print in source-line order instead of by sample count (section accounting)
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

outer = sys.argv[sys.argv.index("--outer") + 1] if "--outer" in sys.argv else None
lo_hi = None
if "--range" in sys.argv:  # only frames of FILE with lo <= line <= hi count as "outer" (skips a dispatching wrapper)
    i = sys.argv.index("--range")
    lo_hi = (int(sys.argv[i + 1]), int(sys.argv[i + 2]))
    del sys.argv[i:i + 3]
args = [a for a in sys.argv[1:] if not a.startswith("--") and a != outer]
rep, lib, kern = args[0], args[1], args[2]
top = int(args[3]) if len(args) > 3 else 30
by_line = "--by-line" in sys.argv

d = tempfile.mkdtemp()
subprocess.run(f"cd {d} && cuobjdump -xelf all {os.path.abspath(lib)} > /dev/null", shell=True, check=True)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
sass = subprocess.run(f"nvdisasm -gi -c {d}/{cubin}", shell=True, capture_output=True, text=True).stdout.split("\n")
infn, frames, fresh, seq = False, [], True, []
for ln in sass:
    if ln.startswith(".text.") or ln.strip().startswith(".section"):
        infn = kern in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:
            frames = []
            fresh = False
        frames.append((m.group(1).split("/")[-1], int(m.group(2))))  # innermost first
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        fresh = True
        key = frames[0] if frames else None
        if outer:
            inside = [f for f in frames if f[0] == outer and (lo_hi is None or lo_hi[0] <= f[1] <= lo_hi[1])]
            if inside:
                key = inside[-1]
        seq.append((key, m.group(2).strip()))

out = subprocess.run(f"ncu -i {rep} --page source --csv --print-source sass", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr = rows[1]
col = {k: hdr.index(k) for k in ("# Samples", "Instructions Executed", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_math",
                                 "stall_branch_resolving")}
data = [r for r in rows[2:] if len(r) == len(hdr)]
if len(data) != len(seq):
    print("warning: instruction count mismatch", len(data), len(seq))
n = min(len(data), len(seq))
agg = collections.defaultdict(collections.Counter)
tot = collections.Counter()
for i in range(n):
    for k, c in col.items():
        v = int(data[i][c] or 0)
        agg[seq[i][0]][k] += v
        tot[k] += v
print("instructions", n, "samples", tot["# Samples"], "executed", tot["Instructions Executed"], {k: tot[k] for k in col if k.startswith("stall")})
items = sorted(agg.items(), key=(lambda kv: (str(kv[0][0]), kv[0][1]) if kv[0] else ("", 0)) if by_line else (lambda kv: -kv[1]["# Samples"]))
if not by_line:
    items = items[:top]
for line, c in items:
    print(f"{str(line):38s} samples {c['# Samples']:6d} {100 * c['# Samples'] / max(1, tot['# Samples']):5.1f}%  exec {c['Instructions Executed']:10d}"
          f" {100 * c['Instructions Executed'] / max(1, tot['Instructions Executed']):5.1f}%"
          f"  long {c['stall_long_sb']:5d} short {c['stall_short_sb']:5d} wait {c['stall_wait']:5d} math {c['stall_math']:5d}")
