"""Event counters of the warp-cooperative fit engine (a library built with -DMDB_FIT_COUNTERS; select it with
MODELARDB_CUDA_LIB) over one compress of bench-shaped data: how many fits, steps and one-thread fits the engine ran.
With the default engine (one lane per chain first) these are the events of the stitching phase alone.
Usage: MODELARDB_CUDA_LIB=... python tools/count_fit_events.py [engine] [series] [eb ...]"""
import ctypes as C
import sys
import time

import torch

sys.path.insert(0, '.')
import bench
from modelardb_rs_b200 import _native
from modelardb_rs_b200 import compression as mc

engine = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ebs = sys.argv[3:] or ["rel:1.0"]
ctx = mc.Context(0)
ctx.set_fit_engine(engine)
import os
if os.environ.get("WARMUP"): ctx.set_lane_warmup(int(os.environ["WARMUP"]))
names = ["fits", "scalar", "steps", "quiet", "spec", "mismatch", "pmc_inorder", "wide", "cyc_load", "cyc_pmc", "cyc_quiet", "cyc_cand", "cyc_scan", "items", "cyc_chain", "cyc_sched"]
for ebs_ in ebs:
    eb = mc.ErrorBound(*bench.parse_eb(ebs_))
    npnt = 1_000_000
    vals = bench.gen_values_device(torch, ns, npnt, 1000, "sine", "cuda:0")
    ts = (bench.EPOCH_US + bench.STEP_US * torch.arange(npnt, device="cuda:0", dtype=torch.int64)).repeat(ns)
    off = torch.arange(ns + 1, device="cuda:0", dtype=torch.int64) * npnt
    out = (C.c_uint64 * 16)()
    mc.compress(ts, vals, off, eb, ctx).free()
    _native.lib().mdbcu_debug_counters(ctx._h, out)
    torch.cuda.synchronize()
    t0 = time.time()
    seg = mc.compress(ts, vals, off, eb, ctx)
    dt = time.time() - t0
    _native.lib().mdbcu_debug_counters(ctx._h, out)
    o = list(out)
    print(ebs_, f"engine {engine}: compress {dt * 1e3:.1f} ms, rows {len(seg)},", dict(zip(names, o)))
    print("   points per fit (if every point were fitted once):", ns * npnt / max(1, o[0]), " steps*128/points:", o[2] * 128 / (ns * npnt))
    seg.free()
