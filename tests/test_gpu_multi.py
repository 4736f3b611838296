"""Multi-GPU behind the C-ABI (include/modelardb_cuda.h, "multi-GPU"): one process per GPU, each with its own context and
the library's own NCCL communicator -- no torch.distributed.  Units are sharded with mdbcu_shard_units, every rank
compresses and aggregates its own series, and ONE packed all-gather yields the table's GROUP BY series result on every
rank; the ungrouped aggregate is folded in rank order.  Needs at least two GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, tmp, n_series, n_points):
    import torch
    from modelardb_rs_b200 import compression as mc
    from modelardb_rs_b200 import synthetic as syn
    from modelardb_rs_b200.sharding import Communicator

    ctx = mc.Context(rank)
    id_path = os.path.join(tmp, "nccl_id.bin")
    if rank == 0:
        uid = Communicator.unique_id()
        with open(id_path + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(id_path + ".tmp", id_path)
    else:
        t0 = time.time()
        while not os.path.exists(id_path):
            assert time.time() - t0 < 120, "rank 0 never published the NCCL id"
            time.sleep(0.05)
        uid = open(id_path, "rb").read()
    comm = Communicator.create(ctx, world, rank, uid)
    assert (comm.world, comm.rank) == (world, rank)
    ts, vals, off = syn.multi_series(n_series, n_points, 21, "sine")
    lo, hi = comm.shard(n_series)
    a, b = int(off[lo]), int(off[hi])
    local_off = (off[lo:hi + 1] - off[lo]).astype(np.uint64)
    out = {}
    # host space: numpy in, numpy out
    seg = mc.compress(ts[a:b], vals[a:b], local_off, mc.ErrorBound.try_new_relative(1.0), ctx)
    host = seg.to_host()
    g = comm.aggregate_sharded(host, host.unit_seg_off, hi - lo, n_series)
    t = comm.aggregate_all_sharded(host)
    out.update(h_count=g[0], h_mn=g[1], h_mx=g[2], h_sm=g[3], ht_count=t[0], ht_mn=t[1], ht_mx=t[2], ht_sm=t[3])
    # device space: the owned batch and device outputs
    g = comm.aggregate_sharded(seg, seg.unit_seg_off_device_ptr(), hi - lo, n_series)
    t = comm.aggregate_all_sharded(seg)
    torch.cuda.synchronize()
    out.update(d_count=g[0].cpu().numpy(), d_mn=g[1].cpu().numpy(), d_mx=g[2].cpu().numpy(), d_sm=g[3].cpu().numpy(),
               dt_count=t[0].cpu().numpy(), dt_mn=t[1].cpu().numpy(), dt_mx=t[2].cpu().numpy(), dt_sm=t[3].cpu().numpy())
    np.savez(os.path.join(tmp, f"rank{rank}.npz"), **out)
    seg.free()
    comm.close()
    ctx.close()


@pytest.mark.parametrize("n_series", [7, 16])
def test_sharded_aggregates_over_the_library_communicator(oracle, tmp_path, n_series):
    import torch
    import torch.multiprocessing as mp
    from modelardb_rs_b200 import synthetic as syn
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    n_points = 20_000
    mp.spawn(_worker, args=(world, str(tmp_path), n_series, n_points), nprocs=world, join=True)
    ts, vals, off = syn.multi_series(n_series, n_points, 21, "sine")
    want_seg = oracle.compress(ts, vals, off, eb=(2, 1.0))
    wc, wmn, wmx, wsm = oracle.aggregate(want_seg, want_seg.unit_seg_off)
    tc, tmn, tmx, tsm = oracle.aggregate(want_seg, None)
    for rank in range(world):
        r = np.load(tmp_path / f"rank{rank}.npz")
        for p in ("h_", "d_"):
            assert np.array_equal(r[p + "count"], wc), (rank, p)
            assert np.array_equal(r[p + "mn"].view(np.uint32), wmn.view(np.uint32)) and np.array_equal(r[p + "mx"].view(np.uint32), wmx.view(np.uint32))
            assert (np.abs(r[p + "sm"] - wsm) <= 1e-12 * np.abs(wsm)).all()
        for p in ("ht_", "dt_"):
            assert r[p + "count"][0] == tc[0] and r[p + "mn"][0] == tmn[0] and r[p + "mx"][0] == tmx[0]
            assert abs(r[p + "sm"][0] - tsm[0]) <= 1e-12 * abs(tsm[0])
