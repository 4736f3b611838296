#!/usr/bin/env python
"""Benchmark of the ModelarDB hot path on B200: compress -> grid -> aggregate of synthetic multi-series data.

Workload (BASELINE.json configs[1]): time series of 1 M f32 points each (sine + noise, regular 1 ms
timestamps at epoch scale), 1 % relative error bound, compress + full grid decompression (+ the
model aggregates GROUP BY series).  The 10 000-series table (10^10 points, 120 GB raw) does not fit
one GPU next to its own reconstruction, so it is streamed in slabs of --series series; one "step" is
one slab through compress -> grid -> aggregate.  With --gpus N every rank processes its own slab
(weak scaling, series are independent; only the per-series aggregates are gathered over NCCL).

One JSON line is printed by rank 0 (see the driver contract in the task statement):
  value     data points/s with inputs resident in HBM (device space of the C-ABI)
  e2e       the same through the C-ABI with HOST buffers (pinned), H2D/D2H inside the timed region
  roofline  the dominant kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the oracle (a C++ port of the reference's Rust) on this box's host cores, bounded sample

`--impl reference` times that CPU port alone with all host threads on the same workload shape.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

EPOCH_US = 1_600_000_000_000_000
STEP_US = 1000


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--config", default="cfg2", choices=list(CONFIGS),
                   help="BASELINE.json config: cfg2 = configs[1] (default), cfg3 = configs[2], cfg4 = configs[3], cfg5 = configs[4]; "
                        "--series / --points / --eb / --kind given explicitly override the preset")
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--series", type=int, default=None, help="series per slab (per GPU; cfg4: in total)")
    p.add_argument("--points", type=int, default=None, help="points per series")
    p.add_argument("--eb", default=None, help="lossless | abs:X | rel:X")
    p.add_argument("--units", default="series", choices=["series", "buffers"],
                   help="series: one compress unit per series (bulk / embedded path); "
                        "buffers: 65 536-point buffers (server ingestion path)")
    p.add_argument("--kind", default=None, choices=["sine", "walk"])
    p.add_argument("--sine", default="50:150,1:20,500:2000", help="per-series ranges of the sine generator: base LO:HI, amplitude LO:HI, period LO:HI")
    p.add_argument("--e2e-series", type=int, default=100, help="series per slab of the host-buffer (e2e) measurement")
    p.add_argument("--e2e-steps", type=int, default=5, help="slabs per worker in the timed e2e region")
    p.add_argument("--e2e-workers", type=int, default=0,
                   help="host threads pipelining slabs (one context each); 0 = up to 8 on one GPU, 4 per rank otherwise (measured on B200, "
                        "round 2: 4 workers x 200 series 3.45 G points/s, 6 x 100 3.56, 8 x 100 3.60)")
    p.add_argument("--e2e-stages", default="all", choices=["all", "compress", "grid"], help="diagnostics: time one half of the e2e slab alone")
    p.add_argument("--e2e-up-gate", type=int, default=2, help="workers allowed at once in the upload-heavy call (compress)")
    p.add_argument("--e2e-down-gate", type=int, default=1, help="workers allowed at once in the download-heavy call (grid)")
    p.add_argument("--chunk-len", type=int, default=0, help="chunk length of the parallel segmentation (0 = automatic); tuning only")
    p.add_argument("--lane-warmup", type=int, default=-1, help="warm-up points of the lane engine (-1 = library default); tuning only")
    p.add_argument("--option", action="append", default=[], help="library tuning knob name=value (mdbcu_context_set_option); repeatable")
    p.add_argument("--fit-engine", type=int, default=0, help="compress engine (0 = automatic); tuning only, results are identical")
    p.add_argument("--cpu-seconds", type=float, default=20.0, help="rough budget of the CPU baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    args = p.parse_args()
    preset = CONFIGS[args.config]
    args.preset_overridden = any(getattr(args, k) is not None for k in ("series", "points", "eb", "kind"))
    for k in ("series", "points", "eb", "kind"):
        if getattr(args, k) is None:
            setattr(args, k, preset[k])
    args.scaling = preset["scaling"]
    return args


# BASELINE.json `configs` as slabs that fit one GPU next to their own reconstruction.  cfg2 / cfg3 / cfg5
# are tables of independent series streamed slab by slab: every rank works on its own slab (weak scaling).  cfg4 is ONE
# 10^9-point table: its 1000 series are sharded over the ranks (strong scaling) and the per-series aggregates gathered.
CONFIGS = {
    "cfg2": dict(series=1000, points=1_000_000, eb="rel:1.0", kind="sine", scaling="weak", name="BASELINE.json configs[1]",
                 what="10k series x 1M points, 1 % relative bound, compress + full grid, streamed in slabs"),
    "cfg3": dict(series=1000, points=1_000_000, eb="lossless", kind="walk", scaling="weak", name="BASELINE.json configs[2]",
                 what="high-entropy random walk, lossless (MacaqueV-heavy), compress + grid, streamed in slabs"),
    "cfg4": dict(series=1000, points=1_000_000, eb="rel:5.0", kind="sine", scaling="strong", name="BASELINE.json configs[3]",
                 what="1B-point table, 5 % relative bound, SUM/MIN/MAX/AVG GROUP BY series from segments, series sharded over the GPUs"),
    "cfg5": dict(series=100_000, points=10_000, eb="rel:1.0", kind="sine", scaling="weak", name="BASELINE.json configs[4]",
                 what="bulk ingest: 100k series x 10k points per slab (bounds 0 / 1 / 5 / 10 % with --eb)"),
}


def parse_eb(text):
    if text == "lossless":
        return (0, 0.0)
    kind, val = text.split(":")
    return ({"abs": 1, "rel": 2}[kind], float(val))


# ----------------------------------------------------------------------------------------------- data

SINE_RANGES = ((50.0, 150.0), (1.0, 20.0), (500.0, 2000.0))  # base, amplitude, period (set from --sine)


def set_sine_ranges(text):
    global SINE_RANGES
    SINE_RANGES = tuple(tuple(float(x) for x in part.split(":")) for part in text.split(","))


def gen_values_device(torch, n_series, n_points, seed, kind, device):
    """Per-series sine + noise (or random walk) as f32, generated on the device in chunks of series."""
    (b0, b1), (a0, a1), (p0, p1) = SINE_RANGES
    out = torch.empty(n_series * n_points, dtype=torch.float32, device=device)
    g = torch.Generator(device=device).manual_seed(seed)
    i = torch.arange(n_points, device=device, dtype=torch.float64)
    chunk = max(1, min(n_series, (64 << 20) // max(1, n_points)))
    for s0 in range(0, n_series, chunk):
        k = min(chunk, n_series - s0)
        if kind == "sine":
            base = b0 + (b1 - b0) * torch.rand(k, 1, device=device, generator=g, dtype=torch.float64)
            amp = a0 + (a1 - a0) * torch.rand(k, 1, device=device, generator=g, dtype=torch.float64)
            period = p0 + (p1 - p0) * torch.rand(k, 1, device=device, generator=g, dtype=torch.float64)
            phase = 6.28 * torch.rand(k, 1, device=device, generator=g, dtype=torch.float64)
            v = base + amp * torch.sin(2.0 * torch.pi * i / period + phase)
            v += 0.1 * torch.randn(k, n_points, device=device, generator=g, dtype=torch.float64)
        else:
            v = 100.0 + torch.cumsum(torch.randn(k, n_points, device=device, generator=g, dtype=torch.float64), dim=1)
        out[s0 * n_points:(s0 + k) * n_points] = v.to(torch.float32).reshape(-1)
        del v
    return out


def gen_values_host(n_series, n_points, seed, kind):
    from modelardb_rs_b200 import synthetic as syn
    rng = np.random.default_rng(seed)
    out = np.empty(n_series * n_points, np.float32)
    (b0, b1), (a0, a1), (p0, p1) = SINE_RANGES
    for s in range(n_series):
        sl = slice(s * n_points, (s + 1) * n_points)
        if kind == "sine":
            out[sl] = syn.sine_noise(n_points, seed + s, base=float(rng.uniform(b0, b1)), amp=float(rng.uniform(a0, a1)),
                                     period=float(rng.uniform(p0, p1)), phase=float(rng.uniform(0, 6.28)))
        else:
            out[sl] = syn.random_walk(n_points, seed + s)
    return out


def unit_offsets(n_series, n_points, units):
    from modelardb_rs_b200.compression import split_into_buffers
    if units == "series":
        return (np.arange(n_series + 1, dtype=np.uint64) * np.uint64(n_points)).astype(np.uint64)
    return split_into_buffers([n_points] * n_series)


# ----------------------------------------------------------------------------------------------- clocks

class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """The timed region starts here: samples that arrived before belong to the warm-up (nvidia-smi needs about a second
        before its first line, so it is started before the warm-up steps)."""
        self.first = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        first = getattr(self, "first", 0)
        timed = self.lines[first:]
        # a timed region shorter than nvidia-smi's sampling period: the last samples of the warm-up (same load), and say so
        source = "timed region" if timed else "warm-up steps right before the timed region (it was shorter than a sampling period)"
        for line in (timed or self.lines[-3:]):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "sampled_during": source, "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------- CPU arm

def cpu_port_inputs(n_series, n_points, kind, units, seed=4242):
    vals = gen_values_host(n_series, n_points, seed, kind)
    ts = np.tile(EPOCH_US + STEP_US * np.arange(n_points, dtype=np.int64), n_series)
    return ts, vals, unit_offsets(n_series, n_points, units)


def cpu_port_throughput(n_series, n_points, eb, kind, units, threads, seed=4242, inputs=None):
    """Times the CPU port of the reference (the oracle) on `n_series` series: compress -> grid -> aggregate,
    series partitioned statically over `threads` threads.  Returns (points/s, per-stage seconds)."""
    from oracle import mdb_oracle as oracle
    ts, vals, off = inputs if inputs is not None else cpu_port_inputs(n_series, n_points, kind, units, seed)
    t0 = time.perf_counter()
    seg = oracle.compress(ts, vals, off, eb=eb, n_threads=threads)
    t1 = time.perf_counter()
    oracle.grid(seg, n_threads=threads)
    t2 = time.perf_counter()
    oracle.aggregate(seg, seg.unit_seg_off, n_threads=threads)
    t3 = time.perf_counter()
    n = n_series * n_points
    return n / (t3 - t0), {"compress_s": t1 - t0, "grid_s": t2 - t1, "aggregate_s": t3 - t2}


def cpu_sample_size(n_points, threads, budget_s):
    # the port does ~8 M points/s/thread over compress+grid+aggregate; one series per thread at least
    per_thread_points = 6e6 * budget_s
    k = max(1, int(per_thread_points // n_points))
    return threads * k


def run_reference_arm(args):
    """--impl reference: the reference's own algorithm on the host cores.  The reference is Rust and cannot
    be built in this image (no cargo/rustc), so this is its C++ port (oracle/), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    eb = parse_eb(args.eb)
    threads = os.cpu_count() or 1
    # W warm-up steps and exactly K timed steps, each a bounded sample of the workload: the sample is sized so that the
    # whole run is about two minutes of CPU work (the inputs, larger than every cache, are generated once)
    warm, steps = max(0, args.warmup), max(1, args.steps)
    n_series = cpu_sample_size(args.points, threads, max(1.0, min(args.cpu_seconds, 120.0 / (warm + steps))))
    inputs = cpu_port_inputs(n_series, args.points, args.kind, args.units)
    times = []
    stages = None
    for step in range(warm + steps):
        rate, stages = cpu_port_throughput(n_series, args.points, eb, args.kind, args.units, threads, inputs=inputs)
        if step >= warm:
            times.append(n_series * args.points / rate)
    ms = 1000.0 * sum(times) / len(times)
    value = n_series * args.points / (ms / 1000.0)
    sample = f"{n_series} series x {args.points} points per step ({args.eb}, units={args.units})"
    line = {
        "impl": "reference", "metric": "compress+grid+aggregate data points/s", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": warm, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 values, f64 fitting, i64 timestamps",
        "data": "synthetic",
        "config": workload_config(args, per_gpu_series=n_series),
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": "port", "sample": sample,
                         "stages_s": stages},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def workload_config(args, per_gpu_series):
    preset = CONFIGS[args.config]
    if args.preset_overridden:  # name the shape that actually ran, not the preset it started from
        name = (f"custom shape (preset {args.config} overridden): {per_gpu_series} series x {args.points} points per GPU per step, "
                f"{args.kind}, error bound {args.eb}")
    else:
        name = f"{preset['name']} ({preset['what']}): {per_gpu_series} series x {args.points} points per GPU per step"
    gen = (f"sine + noise f32, per-series base/amplitude/period in {args.sine}" if args.kind == "sine" else "random walk f32 (100 + cumsum N(0,1))")
    return {
        "workload": f"{name}; {gen}, regular 1 ms timestamps at epoch scale, error bound {args.eb}; a step = compress -> full grid -> "
                    f"aggregates GROUP BY series of one slab",
        "config": args.config if not args.preset_overridden else "custom", "series_per_gpu_per_step": per_gpu_series,
        "points_per_series": args.points, "error_bound": args.eb, "kind": args.kind,
        "compress_units": args.units, "l2": "inputs (12 B/point x slab) far larger than the 126 MB L2; no flush needed",
        "parallelism": (f"series sharded over {args.gpus} GPU(s), no data-path collective; per-series aggregates all-gathered (one packed NCCL all-gather)"),
        "scaling": args.scaling,
    }


# ----------------------------------------------------------------------------------------------- GPU arm

def main():
    args = parse_args()
    set_sine_ranges(args.sine)
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    from modelardb_rs_b200 import compression as mc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(device))

    eb_t = parse_eb(args.eb)
    eb = mc.ErrorBound(*eb_t)
    from modelardb_rs_b200.sharding import Communicator, shard_units
    n_points = args.points
    if args.scaling == "strong":  # one table: its series are sharded over the ranks (mdbcu_shard_units)
        table_series = args.series
        s_lo, s_hi = shard_units(table_series, rank, world)
        n_series = s_hi - s_lo
    else:                         # independent slabs: every rank has its own
        n_series = args.series
        table_series = world * n_series
    n = n_series * n_points
    total_points_per_step = table_series * n_points

    ctx = mc.Context(local_rank)
    comm = None
    if world > 1:  # the library's own NCCL communicator (what a Rust host would use); torch.distributed only hands the id around
        uid = torch.zeros(128, dtype=torch.uint8, device=device)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(Communicator.unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        comm = Communicator.create(ctx, world, rank, bytes(uid.cpu().numpy().tobytes()))
    if args.chunk_len:
        ctx.set_chunk_len(args.chunk_len)
    if args.fit_engine:
        ctx.set_fit_engine(args.fit_engine)
    if args.lane_warmup >= 0:
        ctx.set_lane_warmup(args.lane_warmup)
    for opt in args.option:
        k_, v_ = opt.split("=")
        ctx.set_option(k_, int(v_))
    stream = torch.cuda.ExternalStream(ctx.stream, device=device)

    # ---- inputs resident in HBM
    vals = gen_values_device(torch, n_series, n_points, 1000 + rank, args.kind, device)
    ts = (EPOCH_US + STEP_US * torch.arange(n_points, device=device, dtype=torch.int64)).repeat(n_series)
    off_np = unit_offsets(n_series, n_points, args.units)
    off = torch.from_numpy(off_np.astype(np.int64)).to(device)
    n_units = len(off_np) - 1
    kinds = torch.full((n_units,), eb.kind, dtype=torch.uint8, device=device)
    ebv = torch.full((n_units,), eb.value, dtype=torch.float32, device=device)
    series_group_off = None  # group rows by series = by unit (units=series) or by runs of units
    ts_out = torch.empty(n, dtype=torch.int64, device=device)
    val_out = torch.empty(n, dtype=torch.float32, device=device)
    torch.cuda.synchronize()

    stage_ms = {"compress": [], "grid": [], "aggregate": []}
    seg_bytes_last = [0, 0]

    def step(record, measure_segments=False):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(stream)
        seg = mc.compress(ts, vals, off, (kinds, ebv), ctx)
        ev[1].record(stream)
        mc.grid(seg, ts_out, val_out, ctx)
        ev[2].record(stream)
        if args.units == "series":
            group = (seg.unit_seg_off_device_ptr(), n_units)
        else:
            group = series_group_off_of(seg)
        if comm is not None:  # only the per-series aggregate records travel: one packed all-gather inside the library
            gp = group if isinstance(group, tuple) else (group.data_ptr(), n_series)
            count, mn, mx, sm = comm.aggregate_sharded(seg, gp, n_series, table_series)
        else:
            count, mn, mx, sm = mc.aggregate(seg, group, ctx)
        ev[3].record(stream)
        if record:
            ev[3].synchronize()
            stage_ms["compress"].append(ev[0].elapsed_time(ev[1]))
            stage_ms["grid"].append(ev[1].elapsed_time(ev[2]))
            stage_ms["aggregate"].append(ev[2].elapsed_time(ev[3]))
        if measure_segments:  # (untimed steps only: a host copy of the segments is not part of the device-resident path)
            h = seg.to_host(copy=False)
            seg_bytes_last[0] = h.segment_bytes()
            seg_bytes_last[1] = len(h)
            del h
        seg.free()
        return count

    def series_group_off_of(seg):
        # units=buffers: a series is a run of consecutive units; its rows are [uso[first unit], uso[last unit + 1])
        per_series = n_units // n_series
        idx = torch.arange(0, n_units + 1, per_series, device=device)
        # view the library-owned array through torch without copying it to the host
        uso = _wrap_device_i64(torch, seg.unit_seg_off_device_ptr(), n_units + 1, device)
        with torch.cuda.stream(stream):
            return uso[idx].contiguous()

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    clocks.start()
    step(False, measure_segments=True)  # sizes of the compressed output, outside every timed region
    for _ in range(args.warmup):
        step(False)
    barrier()
    launches0 = ctx.launch_count
    clocks.mark()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step(True)
    e1.record(stream)
    e1.synchronize()
    barrier()
    clock_info = clocks.stop()
    total_ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    # per-kernel device times: a separate, untimed pass with the library's event bracketing switched on (two CUDA events per
    # launch would otherwise sit inside the timed region)
    prof_steps = max(1, min(2, args.steps))
    ctx.set_profiling(True)
    for _ in range(prof_steps):
        step(False)
    torch.cuda.synchronize()
    kstats = ctx.kernel_stats()
    ctx.set_profiling(False)

    t = torch.tensor([total_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = total_points_per_step * args.steps / (total_ms / 1000.0)

    # ---- e2e: same path through the C-ABI with HOST (pinned) buffers
    e2e = None
    if not args.no_e2e:
        # The user-level pattern for host data: slabs are independent, so a few worker threads -- each with its
        # own context (stream) and its own pinned buffers -- keep H2D of one slab, kernels of another and D2H
        # of a third in flight (the calls are blocking but release the GIL; PCIe is full duplex).
        es, esteps = min(args.e2e_series, n_series), args.e2e_steps
        # one GPU: up to 8 workers (sweep above); several ranks share the host's cores and links, and keep the 4 workers and the
        # slab sizes their lines were measured with
        workers = args.e2e_workers or (max(2, min(8, os.cpu_count() or 8)) if world == 1 else 4)
        if world > 1:
            es = min(n_series, 2 * es)
        if world > 2:  # every rank pins 24 B/point x workers of host memory: keep the box's total near the 2-rank figure
            es = max(8, es * 2 // world)
        en = es * n_points
        e_off = unit_offsets(es, n_points, args.units)
        e_units = len(e_off) - 1
        bufs = []
        for w in range(workers):
            b = {"ctx": mc.Context(local_rank),
                 "ts": torch.empty(en, dtype=torch.int64).pin_memory(), "vals": torch.empty(en, dtype=torch.float32).pin_memory(),
                 "ts_out": torch.empty(en, dtype=torch.int64).pin_memory(), "val_out": torch.empty(en, dtype=torch.float32).pin_memory()}
            b["ts"].copy_(ts[:en])
            b["vals"].copy_(vals[(w * en) % max(1, n - en + 1):][:en] if n > en else vals[:en])
            bufs.append(b)
        io = {"h2d": 0, "d2h": 0}

        phase_s = {"compress": 0.0, "to_host": 0.0, "grid": 0.0, "aggregate": 0.0}

        # A bounded number of workers at a time inside the upload-heavy call and inside the download-heavy call: bulk
        # copies in one direction only share that link (and concurrent bulk downloads were measured to slow each other
        # down: 4.3 -> 2.8 G points/s with four at once), while the other direction sits idle.
        up_gate = threading.Semaphore(args.e2e_up_gate)
        down_gate = threading.Semaphore(args.e2e_down_gate)

        def e2e_slab(b):
            p0 = time.perf_counter()
            if args.e2e_stages == "grid" and "seg" in b:  # diagnostics: the download-heavy half alone, on kept segments
                seg = b["seg"]
            else:
                with up_gate:
                    seg = mc.compress(b["ts"].numpy(), b["vals"].numpy(), e_off, [eb] * e_units, b["ctx"])
            p1 = time.perf_counter()
            host_seg = seg.to_host(copy=False)            # what the Rust caller gets back: the RecordBatch columns (host memory)
            p2 = time.perf_counter()
            if args.e2e_stages == "compress":             # diagnostics: the upload-heavy half alone
                del host_seg
                seg.free()
                phase_s["compress"] += p1 - p0
                phase_s["to_host"] += p2 - p1
                return None
            with down_gate:
                mc.grid(host_seg, b["ts_out"].numpy(), b["val_out"].numpy(), b["ctx"])
            p3 = time.perf_counter()
            uso = host_seg.unit_seg_off
            group = uso if args.units == "series" else uso[:: (e_units // es)]
            res = mc.aggregate(host_seg, group, b["ctx"])
            p4 = time.perf_counter()
            for key, d in (("compress", p1 - p0), ("to_host", p2 - p1), ("grid", p3 - p2), ("aggregate", p4 - p3)):
                phase_s[key] += d  # under the GIL; per-call wall time inside a worker, not exclusive time
            seg_b = host_seg.segment_bytes() + 3 * 8 * (len(host_seg) + 1)
            io["h2d"] = 12 * en + 2 * seg_b               # raw points in; segments in again for grid and aggregate
            io["d2h"] = seg_b + 12 * en + 24 * len(res[0])  # segments out; reconstructed points out; aggregates out
            del host_seg
            if args.e2e_stages == "grid":
                b["seg"] = seg
            else:
                seg.free()
            return res

        def worker(b, k):
            for _ in range(k):
                e2e_slab(b)

        def run_all(k):
            th = [threading.Thread(target=worker, args=(b, k)) for b in bufs]
            for t_ in th:
                t_.start()
            for t_ in th:
                t_.join()

        run_all(1)  # warm-up: every worker once
        for key in phase_s:
            phase_s[key] = 0.0
        barrier()
        t0 = time.perf_counter()
        run_all(esteps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        slabs = workers * esteps
        e2e = {"value": world * en * slabs / float(tt.item()), "unit": "points/s", "h2d_bytes_per_step": int(io["h2d"]),
               "d2h_bytes_per_step": int(io["d2h"]), "series_per_gpu_per_step": es, "steps": slabs, "workers": workers, "gate": [args.e2e_up_gate, args.e2e_down_gate],
               "call_ms_mean": {k: 1e3 * v / slabs for k, v in phase_s.items()},
               "note": "a step is one slab through compress -> to_host -> grid -> aggregate with numpy views of pinned host "
                       "tensors in MDBCU_HOST space; `workers` threads each own a context and pipeline slabs; wall clock, max over ranks"}
        for b in bufs:
            if "seg" in b:
                b.pop("seg").free()
            b["ctx"].close()
        del bufs

    if comm is not None:
        comm.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback of B200_PROFILING.md (6.65 TB/s)"
    seg_bytes, n_rows = seg_bytes_last
    algo_bytes = {  # per launch
        "k_spec_chain": 12 * n + 28 * n_rows,         # reads every point once; writes one 28 B model per accepted model
        "k_grid_tile": seg_bytes + 12 * n,            # reads the segments; writes 12 B per point
        "k_grid_sequential": seg_bytes + 12 * n,
        "k_agg_segments": seg_bytes,
        "k_agg_groups": seg_bytes,
        "k_agg_groups_warp": seg_bytes,
        "k_compress_emit": seg_bytes,
        "k_lanes_regular": 8 * n,                     # reads every timestamp once
    }
    # The chain kernel runs once per fixpoint round; its unit of work is one SLAB (one step), so every
    # kernel is accounted per step: algorithmic bytes of one slab / device time the kernel took in one step.
    algo_bytes["k_spec_chain_warp"] = algo_bytes["k_spec_async"] = algo_bytes["k_spec_lanes"] = algo_bytes["k_spec_chain"]
    algo_bytes["k_grid_tile_search"] = algo_bytes["k_grid_tile_tma"] = algo_bytes["k_grid_tile"]

    def algo_of(kernel):  # template instances report as e.g. "k_spec_async<WarpFit>"
        return algo_bytes.get(kernel.split("<")[0], 12 * n)
    # dram bytes per launch from the committed ncu --set full captures (default workload only), scaled by points
    traffic_of = {}
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_path) and args.kind == "sine" and args.eb == "rel:1.0" and args.units == "series":
        for k, v in json.load(open(traffic_path)).items():
            if isinstance(v, dict):
                traffic_of[k] = (v["dram_bytes"] * n / v["points"], v["source"])
    per_step_ms = {k: v[0] / prof_steps for k, v in kstats.items()}
    dominant = max(per_step_ms.items(), key=lambda kv: kv[1])[0] if per_step_ms else None
    roofline = None
    if dominant:
        ab = algo_of(dominant)
        achieved = ab / (per_step_ms[dominant] / 1000.0) / 1e9
        roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic_of.get(dominant, (None, None))[0], "traffic_source": traffic_of.get(dominant, (None, None))[1],
                    "algorithmic_bytes_per_step": ab, "kernel_ms_per_step": per_step_ms[dominant],
                    "launches_per_step": kstats[dominant][1] / prof_steps, "peak_source": peak_src,
                    "kernel_share_of_step": per_step_ms[dominant] / ms_per_step,
                    "algorithmic_bytes_note": "SURVEY.md 8(d) per-point figure of the stage (compress: 12 B read per point + 28 B per model); the "
                                              "screened chain kernel itself reads only the 4 B values, the 8 B timestamps are read once by "
                                              "k_lanes_regular (listed with its own fraction)" if dominant.startswith("k_spec_async<WarpFitScreen") else None,
                    "all_kernels_ms_per_step": dict(sorted(per_step_ms.items(), key=lambda kv: -kv[1]))}
        for k in per_step_ms:
            if k.split("<")[0] in ("k_grid_tile", "k_grid_tile_search", "k_grid_tile_tma", "k_spec_async", "k_spec_lanes", "k_spec_chain", "k_spec_chain_warp",
                                   "k_agg_segments", "k_agg_groups", "k_lanes_regular") \
                    and per_step_ms[k] > 0:
                a_ = algo_of(k) / (per_step_ms[k] / 1000.0) / 1e9
                roofline[f"{k}_GBps"] = a_
                roofline[f"{k}_frac"] = a_ / peak
        # what a caller sees: whole stages (median device time of the timed steps) against the same peak;
        # algorithmic bytes per stage as in SURVEY.md 8(d)
        med_ = {k: statistics.median(v) for k, v in stage_ms.items() if v}
        stage_bytes = {"compress": 12 * n + seg_bytes, "grid": seg_bytes + 12 * n, "aggregate": seg_bytes}
        stages = {}
        for k, b in stage_bytes.items():
            if k in med_ and med_[k] > 0:
                a_ = b / (med_[k] / 1000.0) / 1e9
                stages[k] = {"ms": med_[k], "algorithmic_bytes": b, "achieved": a_, "frac": a_ / peak}
        if "grid" in med_ and "aggregate" in med_:
            b, t_ = stage_bytes["grid"] + stage_bytes["aggregate"], med_["grid"] + med_["aggregate"]
            a_ = b / (t_ / 1000.0) / 1e9
            stages["grid+aggregate"] = {"ms": t_, "algorithmic_bytes": b, "achieved": a_, "frac": a_ / peak}
        roofline["stages"] = stages

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:  # (rank 0 at N = 1 only: the other ranks' processes would wait on it)
        threads = os.cpu_count() or 1
        cs = cpu_sample_size(n_points, threads, args.cpu_seconds)
        rate, stages = cpu_port_throughput(cs, n_points, eb_t, args.kind, args.units, threads)
        cpu_baseline = {"value": rate, "unit": "points/s", "cores": threads, "kind": "port",
                        "sample": f"{cs} series x {n_points} points, one pass compress+grid+aggregate, series split over {threads} threads "
                                  f"(the reference itself compresses on 1 thread: configuration.rs:116-129)",
                        "stages_s": stages}

    med = {k: statistics.median(v) for k, v in stage_ms.items() if v}
    line = {
        "metric": "compress+grid+aggregate data points/s", "value": value, "unit": "points/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32 values, f64 fitting, i64 timestamps", "data": "synthetic",
        "config": workload_config(args, n_series),
        "stage_ms_median": med,
        "stage_points_per_s": {k: total_points_per_step / (v / 1000.0) for k, v in med.items()},
        "compress_chain_rounds": ctx.last_compress_rounds,
        "segments": {"rows_per_slab": n_rows, "segment_bytes_per_point": seg_bytes / n if n else None,
                     "compression_ratio": 12 * n / seg_bytes if seg_bytes else None},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches, "clocks": clock_info,
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def _wrap_device_i64(torch, ptr, n, device):
    """A torch view of a library-owned device array (no copy), via the CUDA array interface."""
    class _Arr:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}
    return torch.as_tensor(_Arr(), device=device)


def _emit(line: dict):
    """The one JSON line, written to the process's ORIGINAL stdout (see _quarantine_stdout)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def _quarantine_stdout() -> int:
    """Native libraries print banners to file descriptor 1 (NCCL: "NCCL version ..." on communicator creation).  The
    contract is ONE JSON line on stdout, so fd 1 is pointed at stderr for the whole run and the line goes to a
    duplicate of the original descriptor."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


_REAL_STDOUT = 1

if __name__ == "__main__":
    _REAL_STDOUT = _quarantine_stdout()
    main()
