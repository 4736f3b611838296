import sys, ctypes as C, numpy as np, torch
sys.path.insert(0,'.')
import bench
from modelardb_rs_b200 import compression as mc, _native
ctx = mc.Context(0)
import os
if os.environ.get("CHUNK"): ctx.set_chunk_len(int(os.environ["CHUNK"]))
for ebs in ("rel:1.0","rel:5.0"):
    eb = mc.ErrorBound(*bench.parse_eb(ebs))
    ns, npnt = 400, 1_000_000
    vals = bench.gen_values_device(torch, ns, npnt, 1000, "sine", "cuda:0")
    ts = (bench.EPOCH_US + bench.STEP_US*torch.arange(npnt, device="cuda:0", dtype=torch.int64)).repeat(ns)
    off = torch.arange(ns+1, device="cuda:0", dtype=torch.int64)*npnt
    out = (C.c_uint64*16)()
    _native.lib().mdbcu_debug_counters(ctx._h, out)
    import time; torch.cuda.synchronize(); t0=time.time(); seg = mc.compress(ts, vals, off, eb, ctx); print("compress s", time.time()-t0)
    _native.lib().mdbcu_debug_counters(ctx._h, out)
    print(ebs, "rounds", ctx.last_compress_rounds, dict(zip(["fits","scalar","steps","quiet","spec","mismatch","pmc_inorder","wide","cyc_load","cyc_pmc","cyc_quiet","cyc_cand","cyc_scan","wide_fail"], list(out))))
    o=list(out); steps=max(1,o[2])
    print("  cycles per step:", {k: round(v/steps) for k,v in zip(["load","pmc","quiet","cand","scan"], o[8:13])})
    seg.free()
