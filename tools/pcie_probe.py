"""Measure pinned host<->device copy bandwidth on the box (one direction, then both at once).

The e2e number of bench.py moves 12 bytes per point in and 12 bytes per point out, so these figures bound it.
Usage: python tools/pcie_probe.py [GiB]
"""
import sys
import time

import torch

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
n = int(gib * (1 << 30))
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d()
    d2h()


def small_d2h_latency(busy):
    """Wall time of an 8-byte D2H + stream sync on s1 while s2 is idle / busy with a bulk D2H."""
    small_d = torch.zeros(1, dtype=torch.int64, device="cuda")
    small_h = torch.zeros(1, dtype=torch.int64).pin_memory()
    torch.cuda.synchronize()
    if busy:
        d2h()
    t0 = time.perf_counter()
    with torch.cuda.stream(s1):
        small_h.copy_(small_d, non_blocking=True)
    s1.synchronize()
    dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    return dt


small_d2h_latency(False)
print(f"8-byte D2H + sync, other stream idle:          {small_d2h_latency(False) * 1e3:8.3f} ms")
print(f"8-byte D2H + sync, other stream in a bulk D2H: {small_d2h_latency(True) * 1e3:8.3f} ms")

t = timed(h2d)
print(f"H2D  {n / t / 1e9:7.1f} GB/s")
t = timed(d2h)
print(f"D2H  {n / t / 1e9:7.1f} GB/s")
t = timed(both)
print(f"both {2 * n / t / 1e9:7.1f} GB/s aggregate ({n / t / 1e9:.1f} each way)")
