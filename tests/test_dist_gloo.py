"""CPU, world_size 2, gloo: the host-side logic of the sharded paths (yael_b200/dist.py) -- shard
bounds, the all-gather layout the merge kernels expect, global ids -- with the oracle standing in
for the per-rank device search."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from oracle import bindings as ob
    from yael_b200 import dist as ydist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r = np.random.RandomState(0)
        base = r.randint(0, 5, (1001, 8)).astype(np.float32)  # ties across shards
        query = r.randint(0, 5, (37, 8)).astype(np.float32)
        k = 9
        lo, hi = ydist.shard_bounds(len(base), world)[rank]
        idx, dis = ob.orc_knn(base[lo:hi], query, k, canonical=True)
        idx = idx + lo  # global ids, as yb_knn_l2(id_offset=lo) returns them
        gi, gd = ydist.allgather_lists(dist, torch, torch.from_numpy(idx), torch.from_numpy(dis), world)
        gi, gd = gi.numpy(), gd.numpy()
        assert gi.shape == (world, 37, k)
        # merge by (distance, id): what yb_knn_merge does on the device
        out_i = np.empty((37, k), np.int32)
        out_d = np.empty((37, k), np.float32)
        for j in range(37):
            ii, dd = gi[:, j].ravel(), gd[:, j].ravel()
            order = np.lexsort((ii, dd))[:k]
            out_i[j], out_d[j] = ii[order], dd[order]
        widx, wdis = ob.orc_knn(base, query, k, canonical=True)
        ok = np.array_equal(out_i, widx) and np.array_equal(out_d, wdis)
        # the library's exchange is QUERY-PARTITIONED (yb_comm.cu: knn_sharded_impl): 37 queries over
        # 2 ranks = slices of 19 with one padding row; every rank merges only its slice of every
        # rank's lists, the merged slices are all-gathered.  Same answer as the full merge.
        sl, bounds = ydist.query_slices(37, world)
        ok = ok and sl == 19 and bounds == [(0, 19), (19, 37)]

        def merge(mi, md):   # [world][slice][k] -> [slice][k] by (distance, id); padding sorts last
            mi, md = mi.numpy(), md.numpy()
            oi = np.empty(mi.shape[1:], np.int32)
            od = np.empty(md.shape[1:], np.float32)
            for j in range(mi.shape[1]):
                ii, dd = mi[:, j].ravel(), md[:, j].ravel()
                key_d = np.where(ii < 0, np.inf, dd)
                order = np.lexsort((ii, key_d))[:k]
                oi[j], od[j] = ii[order], dd[order]
            return torch.from_numpy(oi), torch.from_numpy(od)

        pi, pd = ydist.partitioned_exchange_model(dist, torch, torch.from_numpy(idx), torch.from_numpy(dis),
                                                  rank, world, merge)
        ok = ok and np.array_equal(pi.numpy(), widx) and np.array_equal(pd.numpy(), wdis)
        # sharded k-means bookkeeping: all-reduced sums / counts == unsharded accumulation
        v = r.random_sample((600, 4)).astype(np.float32)
        cent = v[:5].copy()
        lo, hi = ydist.shard_bounds(600, world)[rank]
        _, _, assign, _, _ = ob.orc_kmeans_step(v[lo:hi], cent)
        sums = np.zeros((5, 4), np.float64)
        cnt = np.zeros(5, np.int64)
        np.add.at(sums, assign, v[lo:hi])
        np.add.at(cnt, assign, 1)
        ts, tc = torch.from_numpy(sums), torch.from_numpy(cnt)
        dist.all_reduce(ts)
        dist.all_reduce(tc)
        _, wc, wa, _, wn = ob.orc_kmeans_step(v, cent)
        ok = ok and np.array_equal(tc.numpy(), wn)
        ok = ok and np.allclose(ts.numpy() / np.maximum(tc.numpy(), 1)[:, None], wc, atol=1e-5)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_sharded_knn_and_kmeans_plumbing_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_bounds_cover_everything():
    from yael_b200.dist import shard_bounds
    for n in (0, 1, 7, 1000, 1001):
        for w in (1, 2, 3, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))


def test_query_slices_cover_everything():
    from yael_b200.dist import query_slices
    for nq in (1, 7, 37, 10000):
        for w in (1, 2, 3, 8):
            sl, b = query_slices(nq, w)
            assert sl * w >= nq and b[0][0] == 0 and b[-1][1] == nq
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert all(hi - lo <= sl for lo, hi in b)
