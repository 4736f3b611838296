#!/usr/bin/env python
"""Per-source-line stall samples of one kernel: joins an ncu report's SASS page with nvdisasm line info.
usage: ncu_lines.py <report.ncu-rep> <libmodelardb_cuda.so> <kernel substring> [top N]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
d = tempfile.mkdtemp()
subprocess.run(f"cd {d} && cuobjdump -xelf all {os.path.abspath(lib)} > /dev/null", shell=True, check=True)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
sass = subprocess.run(f"nvdisasm -g -c {d}/{cubin}", shell=True, capture_output=True, text=True).stdout.split("\n")
infn, cur, seq = False, None, []
for ln in sass:
    if ln.startswith(".text.") or ln.strip().startswith(".section"):
        infn = kern in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        seq.append((cur, m.group(2).strip()))
out = subprocess.run(f"ncu -i {rep} --page source --csv --print-source sass", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr = rows[1]
col = {k: hdr.index(k) for k in ("# Samples", "Instructions Executed", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_branch_resolving")}
data = [r for r in rows[2:] if len(r) == len(hdr)]
if len(data) != len(seq):
    print("warning: instruction count mismatch", len(data), len(seq))
n = min(len(data), len(seq))
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for i in range(n):
    for k, c in col.items():
        v = int(data[i][c] or 0)
        agg[seq[i][0]][k] += v
        tot[k] += v
print("instructions", n, "samples", tot["# Samples"], "executed", tot["Instructions Executed"],
      {k: tot[k] for k in col if k.startswith("stall")})
for line, c in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    print(f"{str(line):38s} samples {c['# Samples']:5d} {100*c['# Samples']/max(1,tot['# Samples']):5.1f}%  exec {c['Instructions Executed']:8d}"
          f"  long {c['stall_long_sb']:4d} short {c['stall_short_sb']:4d} wait {c['stall_wait']:4d} math {c['stall_math']:4d}")
