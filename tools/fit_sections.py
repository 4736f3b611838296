"""Instruction and stall-sample shares of the sections of WarpFitT::fit_k inside a chain kernel, from an ncu --set full
report with source (--import-source on): every SASS instruction is charged to the OUTERMOST inline frame that lies in
fit_k, so inlined helpers (divisions, shuffles, scans) count for their call sites.  The section boundaries are found
from marker comments in mdb_fit_warp.cuh, so the tool follows the source.

usage: python tools/fit_sections.py report.ncu-rep libmodelardb_cuda.so kernel_substring
"""
import os
import re
import subprocess
import sys

rep, lib, kern = sys.argv[1:4]
here = os.path.dirname(os.path.abspath(__file__))
src = open(os.path.join(here, "..", "modelardb_rs_b200", "csrc", "mdb_fit_warp.cuh")).read().split("\n")


def line_of(marker, after=0):
    for i, text in enumerate(src):
        if i + 1 > after and marker in text:
            return i + 1
    raise SystemExit(f"marker not found: {marker}")


start = line_of("FittedModel fit_k(uint32_t start")
marks = [("fit prologue", start), ("loop top / loads / special values", line_of("while (pmc_ok || swing_ok)", start)),
         ("regularity", line_of("regularity of newly visited points", start)), ("PMC lane-local prefixes", line_of("float lmn[P], lmx[P];", start)),
         ("PMC warp scan", line_of("in-order warp scan of the lane aggregates", start)), ("PMC in-order sum", line_of("if (!exact) {", start)),
         ("PMC average + bound test", line_of("int fail_p = IDX_INF;", start)), ("PMC commit", line_of("if (accepted > 0) {", start)),
         ("Swing deviations", line_of("double dev[P];", start)), ("Swing quiet test", line_of("// Quiet step", start)),
         ("Swing candidate slopes", line_of("double cus[P], cls[P];", start)), ("Swing bound scan", line_of("while (lo < cnt && swing_ok)", start)),
         ("Swing verification walk", line_of("auto bounds_after", start)), ("Swing mismatch", line_of("if (first_mis < first_rej)", start)),
         ("Swing commit", line_of("const int stop = first_rej", start)), ("fit epilogue", line_of("        FittedModel m;", line_of("MDB_TICK(12)", start)))]
end = marks[-1][1] + 60
out = subprocess.run([sys.executable, os.path.join(here, "ncu_lines.py"), rep, lib, kern, "--outer", "mdb_fit_warp.cuh", "--range", str(start), str(end),
                      "--by-line"], capture_output=True, text=True).stdout
agg, tot = {}, [0, 0]
for ln in out.split("\n"):
    m = re.match(r"\('([^']+)', (\d+)\)\s+samples\s+(\d+)\s+[\d.]+%\s+exec\s+(\d+)", ln)
    if not m:
        continue
    f, line, samples, execd = m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4))
    name = "outside fit_k: " + f
    if f == "mdb_fit_warp.cuh" and start <= line <= end:
        name = [n for n, first in marks if first <= line][-1]
    a = agg.setdefault(name, [0, 0])
    a[0] += samples
    a[1] += execd
    tot[0] += samples
    tot[1] += execd
print(f"sections of fit_k in {kern} ({os.path.basename(rep)}): {tot[1] / 1e6:.0f} M warp instructions, {tot[0]} stall samples\n")
print(f"{'section':40s} {'instructions':>13s} {'samples':>9s}")
for name, (samples, execd) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if execd * 1000 >= tot[1]:
        print(f"{name:40s} {100 * execd / tot[1]:12.1f}% {100 * samples / tot[0]:8.1f}%")
