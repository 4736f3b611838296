"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list as a markdown table.

Usage: python tools/launch_summary.py launches.csv "command that produced it" > profiles/rNN_launch_list_summary.md
Kernels of torch (data generation in bench.py) are folded into one row.
"""
import csv
import sys
from collections import OrderedDict

path, command = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = val / 1000.0 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1000.0
    name = r["Kernel Name"]
    name = name.split("(")[0].strip()
    if "at::" in name or "elementwise" in name or "cub::" in name:
        name = "(torch: synthetic data generation / fills)"
    rows.append((name, us))

agg = OrderedDict()
for name, us in rows:
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += us
    a[2] = max(a[2], us)
total = sum(a[1] for k, a in agg.items() if not k.startswith("(torch"))
print("# ncu launch list (gpu__time_duration.sum, --clock-control none)\n")
if command:
    print(f"Command: `{command}`\n")
print("Times are per-launch, cold-cache and serialised by ncu: compare SHARES with bench.py's live CUDA-event numbers, not absolutes.")
print(f"Share is of this library's kernels only ({total / 1000.0:.2f} ms over {sum(a[0] for k, a in agg.items() if not k.startswith('(torch'))} launches).\n")
print("| kernel | launches | total us | max us | share |")
print("|---|---:|---:|---:|---:|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    share = "" if name.startswith("(torch") else f"{100.0 * a[1] / total:.1f} %"
    print(f"| {name} | {a[0]} | {a[1]:.1f} | {a[2]:.1f} | {share} |")

chain = [us for name, us in rows if name.startswith("k_spec_chain")]
if chain:
    print("\nChain kernel launches in order (us): " + ", ".join(f"{u:.0f}" for u in chain[:64]))
