"""2 GPUs, NCCL (skipped on a single-GPU box): the sharded paths of yael_b200/dist.py end to end --
database-sharded kNN and Hamming kNN (all-gather + merge) and point-sharded k-means (all-reduce
hook of the C host loop) -- against the single-GPU result."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import yael_b200
        return yael_b200.lib().yb_device_count()
    except Exception:
        return 0


def _worker(rank, world, port, q):
    import faulthandler
    faulthandler.dump_traceback_later(100, exit=True)  # a stuck collective must not hang the box
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    import yael_b200
    from yael_b200 import dist as ydist, ynumpy
    torch.cuda.set_device(rank)
    yael_b200.lib().yb_set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ok = True
    detail = {}
    try:
        r = np.random.RandomState(0)
        base = r.random_sample((40000, 64)).astype(np.float32)
        query = r.random_sample((500, 64)).astype(np.float32)
        k = 20
        lo, hi = ydist.shard_bounds(len(base), world)[rank]
        s = ydist.ShardedKnn(torch.from_numpy(base[lo:hi]).to(dev), k, rank=rank, world=world, id_offset=lo)
        idx, dis = s.search(torch.from_numpy(query).to(dev))
        widx, wdis = ynumpy.knn(query, base, k)
        detail['knn'] = bool(np.array_equal(idx.cpu().numpy(), widx) and np.array_equal(dis.cpu().numpy(), wdis))
        # end-to-end on host buffers: the shard is fed over PCIe while it is scanned
        big = r.random_sample((420000, 128)).astype(np.float32)
        bq = r.random_sample((64, 128)).astype(np.float32)
        lo, hi = ydist.shard_bounds(len(big), world)[rank]
        s2 = ydist.ShardedKnn(torch.empty((hi - lo, 128), dtype=torch.float32, device=dev), 10, rank=rank,
                              world=world, id_offset=lo)
        oi = np.empty((64, 10), np.int32)
        od = np.empty((64, 10), np.float32)
        s2.search_host(np.ascontiguousarray(big[lo:hi]), bq, oi, od)
        wi2, wd2 = ynumpy.knn(bq, big, 10)
        detail['knn_host'] = bool(np.array_equal(oi, wi2) and np.array_equal(od, wd2))
        del big, s2
        # Hamming: bit-identical for any shard count
        codes = r.randint(0, 256, (30000, 8)).astype(np.uint8)
        qc = r.randint(0, 256, (200, 8)).astype(np.uint8)
        lo, hi = ydist.shard_bounds(len(codes), world)[rank]
        h = ydist.ShardedHamming(torch.from_numpy(codes[lo:hi]).to(dev), 15, rank=rank, world=world, id_offset=lo)
        hi_, hd_ = h.search(torch.from_numpy(qc).to(dev))
        wi, wd = ynumpy.knn_hamming(qc, codes, 15)
        detail['hamming'] = bool(np.array_equal(hi_.cpu().numpy(), wi) and
                                 np.array_equal(hd_.cpu().numpy().view(np.uint16), wd))
        # k-means: sharded points, all-reduced sums; same assignments, centroids within 1e-4
        v = r.random_sample((20000, 32)).astype(np.float32)
        init = v[:64].copy()
        lo, hi = ydist.shard_bounds(len(v), world)[rank]
        cent, qerr, assign, nassign = ydist.sharded_kmeans(torch.from_numpy(v[lo:hi]).to(dev), 64, 5, init, len(v))
        wc, wq, _, wa, wn = ynumpy.kmeans(v, 64, niter=5, verbose=False, init=init, output="all")
        detail['kmeans_counts'] = bool(np.array_equal(nassign, wn) and np.array_equal(assign, wa[lo:hi]))
        detail['kmeans_cent_maxdiff'] = float(np.abs(cent - wc).max())
        detail['kmeans_qerr'] = (float(qerr), float(wq))
        # the round-1 exchange (torch.distributed all-gather + full merge) stays as the A/B arm
        s3 = ydist.ShardedKnn(s.base, k, rank=rank, world=world, id_offset=s.id_offset, exchange="torch")
        i3, d3 = s3.search(torch.from_numpy(query).to(dev))
        detail['knn_torch_exchange'] = bool(np.array_equal(i3.cpu().numpy(), widx) and
                                            np.array_equal(d3.cpu().numpy(), wdis))
        # ragged: nq not a multiple of the rank count (padding rows of the query partition)
        i4, d4 = s.search(torch.from_numpy(query[:333]).to(dev))
        detail['knn_ragged'] = bool(np.array_equal(i4.cpu().numpy(), widx[:333]) and
                                    np.array_equal(d4.cpu().numpy(), wdis[:333]))
        # uneven shards with the offsets derived from an all-gather of the shard sizes, and a shard
        # with FEWER rows than k (the reference only needs k <= n over all rows, nn.c:456)
        small = r.random_sample((150, 64)).astype(np.float32)
        cut = 30 if rank == 0 else 150
        part = small[:30] if rank == 0 else small[30:]
        s6 = ydist.ShardedKnn(torch.from_numpy(part).to(dev), 50, rank=rank, world=world)
        i6, d6 = s6.search(torch.from_numpy(query[:40]).to(dev))
        wi6, wd6 = ynumpy.knn(query[:40], small, 50)
        detail['knn_uneven_small_shard'] = bool(np.array_equal(i6.cpu().numpy(), wi6) and
                                                np.array_equal(d6.cpu().numpy(), wd6))
        del cut
        c5, q5, a5, n5 = ydist.sharded_kmeans(torch.from_numpy(v[lo:hi]).to(dev), 64, 5, init, len(v),
                                              exchange="torch")
        detail['kmeans_torch_exchange'] = bool(np.array_equal(n5, nassign) and np.array_equal(c5, cent))
        ok = detail['knn'] and detail['knn_host'] and detail['hamming'] and detail['kmeans_counts'] and \
            detail['kmeans_cent_maxdiff'] < 1e-4 and abs(qerr - wq) < 1e-4 * wq and \
            detail['knn_torch_exchange'] and detail['knn_ragged'] and detail['kmeans_torch_exchange'] and \
            detail['knn_uneven_small_shard']
        q.put((rank, True if ok else repr(detail)))
    except Exception as e:  # report instead of hanging the peer
        q.put((rank, "error: %r" % (e,)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_sharded_paths_match_single_gpu():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)], res


def _mgpu_worker(q):
    """ONE process, both GPUs: the drop-in calls themselves shard (yb_mgpu.cu, ncclCommInitAll)."""
    import faulthandler
    faulthandler.dump_traceback_later(150, exit=True)
    sys.path.insert(0, ROOT)
    os.environ["YAEL_B200_MGPU_MIN_WORK"] = "0"
    os.environ.pop("YAEL_GPU_DEVICES", None)
    import ctypes as C
    import yael_b200
    from yael_b200 import ynumpy
    L = yael_b200.lib()
    detail = {}
    try:
        r = np.random.RandomState(5)
        base = r.random_sample((300000, 64)).astype(np.float32)
        query = r.random_sample((777, 64)).astype(np.float32)
        detail['devices'] = L.yb_mgpu_device_count()
        i2, d2 = ynumpy.knn(query, base, 20)
        detail['knn_used'] = L.yb_mgpu_last_used()
        codes = r.randint(0, 256, (200000, 8)).astype(np.uint8)
        qc = r.randint(0, 256, (301, 8)).astype(np.uint8)
        hi2, hd2 = ynumpy.knn_hamming(qc, codes, 15)
        detail['hamming_used'] = L.yb_mgpu_last_used()
        v = r.random_sample((60000, 32)).astype(np.float32)
        km2 = ynumpy.kmeans(v, 64, niter=6, verbose=False, seed=7, output="all")
        detail['kmeans_used'] = L.yb_mgpu_last_used()
        kmone2 = ynumpy.kmeans(v, 64, niter=1, verbose=False, seed=7, output="all")
        kpp2 = ynumpy.kmeans(v[:20000], 16, niter=3, verbose=False, seed=3, init="kmeans++", output="all")
        one = (C.c_int * 1)(0)
        L.yb_mgpu_set_devices(1, one)      # the same calls on one GPU
        i1, d1 = ynumpy.knn(query, base, 20)
        detail['single_used'] = L.yb_mgpu_last_used()
        hi1, hd1 = ynumpy.knn_hamming(qc, codes, 15)
        km1 = ynumpy.kmeans(v, 64, niter=6, verbose=False, seed=7, output="all")
        kmone1 = ynumpy.kmeans(v, 64, niter=1, verbose=False, seed=7, output="all")
        kpp1 = ynumpy.kmeans(v[:20000], 16, niter=3, verbose=False, seed=3, init="kmeans++", output="all")
        detail['knn'] = bool(np.array_equal(i1, i2) and np.array_equal(d1, d2))
        detail['hamming'] = bool(np.array_equal(hi1, hi2) and np.array_equal(hd1, hd2))
        # centroids, qerr, dis, assign, nassign: same random init (replayed from the host rows), same
        # assignments; sums differ only by the order of the per-GPU partial sums
        # after ONE iteration: identical assignments and counts, centroids to rounding; after six the
        # two runs may have drifted apart at a few near-tie points, nothing more
        detail['kmeans_one_iter'] = bool(np.array_equal(kmone1[3], kmone2[3]) and np.array_equal(kmone1[4], kmone2[4])
                                         and np.abs(kmone1[0] - kmone2[0]).max() < 1e-5)
        detail['kmeans_one_iter_maxdiff'] = float(np.abs(kmone1[0] - kmone2[0]).max())
        detail['kmeans_assign_diff'] = int((km1[3] != km2[3]).sum())
        detail['kmeans_cent_maxdiff'] = float(np.abs(km1[0] - km2[0]).max())
        detail['kmeans'] = bool(detail['kmeans_one_iter'] and detail['kmeans_assign_diff'] <= len(v) // 200 and
                                abs(km1[1] - km2[1]) < 1e-4 * km1[1])
        detail['kmeanspp'] = bool(np.array_equal(kpp1[3], kpp2[3]) and np.abs(kpp1[0] - kpp2[0]).max() < 1e-4)
        ok = detail['devices'] >= 2 and detail['knn_used'] >= 2 and detail['hamming_used'] >= 2 and \
            detail['kmeans_used'] >= 2 and detail['single_used'] == 1 and detail['knn'] and \
            detail['hamming'] and detail['kmeans'] and detail['kmeanspp']
        q.put(True if ok else "mismatch: " + repr(detail))
    except Exception as e:
        q.put("error: %r %r" % (e, detail))


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_dropin_calls_shard_over_the_gpus_of_one_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_mgpu_worker, args=(q,))
    p.start()
    res = q.get(timeout=200)
    p.join(timeout=60)
    assert res is True, res
