"""Aggregate pinned host<->device bandwidth of the box with N GPUs copying AT THE SAME TIME.

bench.py's e2e leg moves 12 B/point in and 12 B/point out of every GPU; whether 8 GPUs can do that 8x as fast as one is a
property of the host (PCIe topology, IOMMU / virtualisation, host DRAM and NUMA), which this measures without any of
the library: one process per GPU (torchrun), every rank copies `GiB` of pinned memory H2D, D2H and both at once, all
ranks start each phase together (gloo barrier), and rank 0 prints per-rank and aggregate GB/s as one JSON line.

Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 \
           tools/pcie_probe_multi.py [GiB] [numa]      (plain `python tools/pcie_probe_multi.py` = one GPU)
`numa`: also report which NUMA node the GPU and the process's CPUs are on (sysfs), to explain asymmetries.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")

n = int(gib * (1 << 30))
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d()
    d2h()


def barrier():
    if world > 1:
        dist.barrier()


def timed(fn, reps=4):
    fn()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    barrier()
    return dt


res = {}
for name, fn, factor in (("h2d", h2d, 1), ("d2h", d2h, 1), ("both", both, 2)):
    res[name] = factor * n / timed(fn) / 1e9


def numa_info():
    info = {}
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
    except Exception:
        bus = None
    try:
        import subprocess
        q = subprocess.run(["nvidia-smi", f"--id={local}", "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True)
        bus = q.stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = "0000:" + bus.split(":", 1)[1]
        info["pci"] = bus
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            info["gpu_numa_node"] = int(f.read())
    except Exception as e:  # noqa: BLE001
        info["gpu_numa_node"] = f"unknown ({type(e).__name__})"
    try:
        info["cpus_allowed"] = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return info


row = {"rank": rank, **{k: round(v, 1) for k, v in res.items()}, **numa_info()}
rows = [None] * world
if world > 1:
    dist.all_gather_object(rows, row)
else:
    rows = [row]
if rank == 0:
    agg = {k: round(sum(r[k] for r in rows), 1) for k in ("h2d", "d2h", "both")}
    out = {"probe": "pinned host<->device copies, all GPUs at once", "n_gpus": world, "GiB_per_copy": gib, "unit": "GB/s",
           "aggregate": agg, "per_gpu_mean": {k: round(v / world, 1) for k, v in agg.items()}, "ranks": rows,
           "host": {"cpus": os.cpu_count(), "numa_nodes": len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
                    if os.path.isdir("/sys/devices/system/node") else None}}
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
