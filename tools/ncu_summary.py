"""Text summary of one `ncu --set full` report for profiles/: launch, throughput, memory, pipes, stalls.

Usage: python tools/ncu_summary.py report.ncu-rep "command that produced it" > profiles/rNN_<kernel>_ncu_full.txt
"""
import csv
import subprocess
import sys

rep, command = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
print(f"ncu --set full summary of {rep.split('/')[-1]}")
if command:
    print(f"command: {command}")
print(f"kernel: {m.get('Kernel Name', ('', '?'))[1]}   grid {m.get('launch__grid_size', ('', '?'))[1]} x block {m.get('launch__block_size', ('', '?'))[1]}"
      f"   registers/thread {m.get('launch__registers_per_thread', ('', '?'))[1]}\n")
groups = [
    ("time / throughput", ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
                           "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
                           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
                           "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor"]),
    ("memory", ["dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
                "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum"]),
    ("pipes (% of peak while active)", ["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                                        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
                                        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active"]),
]
for title, keys in groups:
    print(f"[{title}]")
    for k in keys:
        if k in m:
            print(f"  {k:75s} {m[k][1]:>18s} {m[k][0]}")
    print()
print("[warp stalls per issued instruction]")
print(f"  {'smsp__average_warp_latency_per_inst_issued.ratio':75s} {m.get('smsp__average_warp_latency_per_inst_issued.ratio', ('', '?'))[1]:>18s} cycle")
stalls = sorted(((float(v.replace(',', '')), h) for h, (u, v) in m.items()
                 if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")), reverse=True)
for v, h in stalls[:10]:
    print(f"  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):75s} {v:18.3f}")
