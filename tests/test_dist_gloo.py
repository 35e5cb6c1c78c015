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
        # the k-NN exchange proper: ids and distance bits of a rank travel as ONE [2][nq][k] block
        # (ShardedKnn._exchange); the merge reads shard g at gbuf + g * 2*nq*k (ids) and + nq*k (dis)
        import types
        buf = torch.empty((2, 37, k), dtype=torch.int32)
        buf[0] = torch.from_numpy(idx)
        buf[1] = torch.from_numpy(dis).view(torch.int32)
        stub = types.SimpleNamespace(torch=torch, world=world)
        gbuf = ydist.ShardedKnn._exchange(stub, buf, 37)
        flat = gbuf.numpy().reshape(-1)
        for g in range(world):
            ids_g = flat[g * 2 * 37 * k: g * 2 * 37 * k + 37 * k].reshape(37, k)
            dis_g = flat[g * 2 * 37 * k + 37 * k: (g + 1) * 2 * 37 * k].view(np.float32).reshape(37, k)
            ok = ok and np.array_equal(ids_g, gi[g]) and np.array_equal(dis_g, gd[g])
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
