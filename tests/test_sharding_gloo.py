"""The multi-GPU path on CPU: two processes over gloo, each owning its shard of the series.  The per-rank work is
done by the oracle here (no GPU in this container); what is under test is the sharding and the gather."""
import os
import socket

import numpy as np
import pytest

from modelardb_rs_b200.sharding import shard_units


def test_shard_units_cover_everything_once():
    for n in (0, 1, 2, 7, 8, 9, 1000):
        for world in (1, 2, 3, 8):
            ranges = [shard_units(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_units(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_series, n_points, out_dir):
    import torch
    import torch.distributed as dist
    from modelardb_rs_b200 import synthetic as syn
    from modelardb_rs_b200.sharding import gather_group_aggregates
    from oracle import mdb_oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ts, vals, off = syn.multi_series(n_series, n_points, 21, "sine")
        lo, hi = shard_units(n_series, rank, world)
        a, b = int(off[lo]), int(off[hi])
        seg = O.compress(ts[a:b], vals[a:b], off[lo:hi + 1] - off[lo], eb=(2, 1.0))
        count, mn, mx, sm = O.aggregate(seg, seg.unit_seg_off)
        g = gather_group_aggregates(torch.from_numpy(count), torch.from_numpy(mn), torch.from_numpy(mx), torch.from_numpy(sm), n_series)
        from modelardb_rs_b200.sharding import combine_global_aggregates
        c1, mn1, mx1, sm1 = O.aggregate(seg, None)
        t = combine_global_aggregates(torch.from_numpy(c1), torch.from_numpy(mn1), torch.from_numpy(mx1), torch.from_numpy(sm1))
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), count=g[0].numpy(), mn=g[1].numpy(), mx=g[2].numpy(), sm=g[3].numpy(),
                 tcount=t[0].numpy(), tmn=t[1].numpy(), tmx=t[2].numpy(), tsm=t[3].numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_series", [5, 8])
def test_two_ranks_over_gloo_give_the_single_process_aggregates(oracle, tmp_path, n_series):
    import torch.multiprocessing as mp
    from modelardb_rs_b200 import synthetic as syn
    n_points, world = 3000, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_series, n_points, str(tmp_path)), nprocs=world, join=True)
    ts, vals, off = syn.multi_series(n_series, n_points, 21, "sine")
    seg = oracle.compress(ts, vals, off, eb=(2, 1.0))
    count, mn, mx, sm = oracle.aggregate(seg, seg.unit_seg_off)
    for rank in range(world):  # every rank ends up with the complete result, in unit order
        got = np.load(os.path.join(str(tmp_path), f"rank{rank}.npz"))
        assert np.array_equal(got["count"], count)
        assert np.array_equal(got["mn"].view(np.uint32), mn.view(np.uint32))
        assert np.array_equal(got["mx"].view(np.uint32), mx.view(np.uint32))
        assert np.array_equal(got["sm"].view(np.uint64), sm.view(np.uint64))
        # ungrouped: the whole table as one group
        wc, wmn, wmx, wsm = oracle.aggregate(seg, None)
        assert np.array_equal(got["tcount"], wc)
        assert np.array_equal(got["tmn"].view(np.uint32), wmn.view(np.uint32))
        assert np.array_equal(got["tmx"].view(np.uint32), wmx.view(np.uint32))
        assert abs(got["tsm"][0] - wsm[0]) <= 1e-12 * abs(wsm[0])  # per-rank partial sums added in rank order
