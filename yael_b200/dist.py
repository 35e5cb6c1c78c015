"""One-process-per-GPU sharding of the hot path.

The reference is a single process; what follows is new plumbing around the same kernels:

  ShardedKnn / ShardedHamming   database rows split contiguously over the ranks, queries
        replicated; every rank computes a local top-k with GLOBAL ids, the lists are
        all-gathered (NCCL) and merged by (distance, id) on every rank, so the result does
        not depend on the number of ranks.
  sharded_kmeans                points split contiguously; centroids replicated; per
        iteration one all-reduce of (k*d sums, k counts, qerr) between the local
        accumulation and the scaling, through the yb_kmeans_comm_t hook of the C host loop.

torch is used for device memory, streams and torch.distributed only; all compute goes
through libyael_b200.so's device-level C ABI (include/yael_b200.h).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib


def shard_bounds(n, world):
    """Contiguous shards [n*r/world, n*(r+1)/world) -- the reference's own slicing rule for its
    OpenMP tasks (yael/nn.c:669-670)."""
    return [(n * r // world, n * (r + 1) // world) for r in range(world)]


def _stream_ptr(torch):
    """torch's current stream as the handle the C ABI expects.  Handle 0 would mean "the library's
    own stream" there, so the legacy default stream is passed as cudaStreamLegacy (0x1): our
    kernels and the NCCL collectives torch enqueues then share one stream order."""
    h = torch.cuda.current_stream().cuda_stream
    return C.c_void_p(h if h else 1)


def allgather_lists(dist, torch, idx, dis, world):
    """[nq][k] per rank -> [world][nq][k] on every rank (one collective per array)."""
    gi = torch.empty((world,) + tuple(idx.shape), dtype=idx.dtype, device=idx.device)
    gd = torch.empty((world,) + tuple(dis.shape), dtype=dis.dtype, device=dis.device)
    dist.all_gather_into_tensor(gi.view(-1), idx.reshape(-1).contiguous())
    # NCCL has no 16-bit integer type: ship the uint16 Hamming distances as raw bytes
    dist.all_gather_into_tensor(gd.view(-1).view(torch.uint8), dis.reshape(-1).contiguous().view(torch.uint8))
    return gi, gd


class ShardedKnn:
    """Exact L2 k-NN over a database sharded by rows.  `base` is THIS rank's shard (a CUDA
    float32 tensor [rows][d]); `id_offset` the global id of its first row."""

    def __init__(self, base, k, rank=0, world=1, id_offset=None):
        import torch
        self.torch = torch
        assert base.is_cuda and base.dtype == torch.float32 and base.is_contiguous()
        self.base, self.k, self.rank, self.world = base, k, rank, world
        if id_offset is None:
            id_offset = rank * base.shape[0]
        self.id_offset = int(id_offset)
        self.events = None   # set to [] to record (start, local, gathered, merged) CUDA events per search
        _lib.require_gpu()

    def phase_ms(self):
        """Average milliseconds of the three phases of the recorded searches (world > 1)."""
        if not self.events:
            return None
        self.torch.cuda.synchronize()
        n = len(self.events)
        return {"local_scan": sum(a.elapsed_time(b) for a, b, _, _ in self.events) / n,
                "all_gather": sum(b.elapsed_time(c) for _, b, c, _ in self.events) / n,
                "merge": sum(c.elapsed_time(d) for _, _, c, d in self.events) / n}

    def _local_into(self, query, idx, dis, id_offset):
        torch = self.torch
        assert query.is_cuda and query.dtype == torch.float32 and query.is_contiguous()
        nq, d = query.shape
        nb = self.base.shape[0]
        check(lib().yb_knn_l2(nq, nb, d, self.k, self.base.data_ptr(), query.data_ptr(), None,
                              idx.data_ptr(), dis.data_ptr(), int(id_offset),
                              _stream_ptr(torch)), "yb_knn_l2")

    def _local(self, query, idx, dis):
        self._local_into(query, idx, dis, self.id_offset)

    def _exchange(self, buf, nq):
        """buf: this rank's [2][nq][k] int32 block (ids, distance bits).  ONE all-gather, then the
        merge by (distance, id) on every rank."""
        torch = self.torch
        import torch.distributed as dist
        gbuf = torch.empty((self.world,) + tuple(buf.shape), dtype=torch.int32, device=buf.device)
        dist.all_gather_into_tensor(gbuf.view(-1), buf.view(-1))
        return gbuf

    def _merge(self, gbuf, nq):
        torch = self.torch
        oi = torch.empty((nq, self.k), dtype=torch.int32, device=gbuf.device)
        od = torch.empty((nq, self.k), dtype=torch.float32, device=gbuf.device)
        check(lib().yb_knn_merge_strided(nq, self.k, self.world, gbuf.data_ptr(), gbuf[0, 1].data_ptr(),
                                         2 * nq * self.k, oi.data_ptr(), od.data_ptr(),
                                         _stream_ptr(torch)), "yb_knn_merge_strided")
        return oi, od

    def search(self, query):
        torch = self.torch
        nq = query.shape[0]
        # ids and distances of this rank side by side: they travel in one collective
        buf = torch.empty((2, nq, self.k), dtype=torch.int32, device=query.device)
        idx, dis = buf[0], buf[1].view(torch.float32)
        ev = None
        if self.events is not None and self.world > 1:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
        self._local(query, idx, dis)
        if self.world == 1:
            return idx, dis
        if ev:
            ev[1].record()
        gbuf = self._exchange(buf, nq)
        if ev:
            ev[2].record()
        oi, od = self._merge(gbuf, nq)
        if ev:
            ev[3].record()
            self.events.append(tuple(ev))
        return oi, od

    def search_host(self, base_host, query_host, idx_out, dis_out):
        """End-to-end step on HOST buffers: host->device of the shard and the queries, search,
        device->host of the result.  world == 1 is exactly the drop-in C call."""
        torch = self.torch
        nq, d = query_host.shape
        if self.world == 1:
            f, i = C.POINTER(C.c_float), C.POINTER(C.c_int)
            lib().knn_full_thread(2, nq, base_host.shape[0], d, self.k,
                                  base_host.ctypes.data_as(f), query_host.ctypes.data_as(f), None,
                                  idx_out.ctypes.data_as(i), dis_out.ctypes.data_as(f), 1)
            return
        # the shard travels over PCIe while it is being scanned (yb_knn_l2_hostbase), then the
        # usual all-gather + merge
        q = torch.from_numpy(query_host).to(self.base.device, non_blocking=True)
        buf = torch.empty((2, nq, self.k), dtype=torch.int32, device=q.device)
        idx, dis = buf[0], buf[1].view(torch.float32)
        check(lib().yb_knn_l2_hostbase(nq, base_host.shape[0], d, self.k, base_host.ctypes.data,
                                       self.base.data_ptr(), q.data_ptr(), idx.data_ptr(),
                                       dis.data_ptr(), self.id_offset, _stream_ptr(torch)),
              "yb_knn_l2_hostbase")
        oi, od = self._merge(self._exchange(buf, nq), nq)
        torch.from_numpy(idx_out).copy_(oi, non_blocking=True)
        torch.from_numpy(dis_out).copy_(od, non_blocking=True)
        torch.cuda.synchronize()


class ShardedHamming:
    """nn_hamming over a code database sharded by rows; merged result is bit-identical for any
    number of ranks ((distance, id) order, global ids)."""

    def __init__(self, base, k, rank=0, world=1, id_offset=None):
        import torch
        self.torch = torch
        assert base.is_cuda and base.dtype == torch.uint8 and base.is_contiguous()
        self.base, self.k, self.rank, self.world = base, k, rank, world
        self.id_offset = int(rank * base.shape[0] if id_offset is None else id_offset)
        _lib.require_gpu()

    def search(self, query):
        torch = self.torch
        nq, nc = query.shape
        idx = torch.empty((nq, self.k), dtype=torch.int32, device=query.device)
        dis = torch.empty((nq, self.k), dtype=torch.int16, device=query.device)  # uint16 payload
        check(lib().yb_nn_hamming(nq, self.base.shape[0], nc, self.k, self.base.data_ptr(),
                                  query.data_ptr(), idx.data_ptr(), dis.data_ptr(), self.id_offset,
                                  _stream_ptr(torch)), "yb_nn_hamming")
        if self.world == 1:
            return idx, dis
        import torch.distributed as dist
        gi, gd = allgather_lists(dist, torch, idx, dis, self.world)
        oi, od = torch.empty_like(idx), torch.empty_like(dis)
        check(lib().yb_nn_hamming_merge(nq, self.k, self.world, gi.data_ptr(), gd.data_ptr(),
                                        oi.data_ptr(), od.data_ptr(), _stream_ptr(torch)),
              "yb_nn_hamming_merge")
        return oi, od


# C struct yb_kmeans_comm_t (include/yael_b200.h)
_ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long,
                            C.c_void_p, C.c_void_p)


class _KmeansComm(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("allreduce_sums", _ALLREDUCE_FN), ("n_total", C.c_long)]


def _wrap_device(torch, ptr, n, dtype, device):
    """A torch tensor aliasing `n` elements of device memory at `ptr` (no copy)."""
    itemsize = torch.empty((), dtype=dtype).element_size()

    class _Holder:  # __cuda_array_interface__ provider
        pass

    h = _Holder()
    typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.float64: "<f8"}[dtype]
    h.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                  "version": 2, "strides": None}
    del itemsize
    return torch.as_tensor(h, device=device)


def sharded_kmeans(v_shard, k, niter, init_centroids, n_total, flags=0, seed=0, group=None):
    """Lloyd's k-means (yael/kmeans.c:213-329) on points sharded by rows.  `v_shard`: this
    rank's CUDA float32 [n_local][d]; `init_centroids`: numpy [k][d], identical on every rank
    (KMEANS_INIT_USER).  Returns (centroids, qerr, assign_local, nassign)."""
    import torch
    import torch.distributed as dist
    from .ynumpy import KMEANS_INIT_USER, KMEANS_QUIET
    n, d = v_shard.shape
    dev = v_shard.device
    world = dist.get_world_size(group) if dist.is_initialized() else 1

    def allreduce(ctx, sums, nf, cnts, ni, qerr, stream):
        try:
            if world > 1:
                ts = _wrap_device(torch, sums, nf, torch.float32, dev)
                tc = _wrap_device(torch, cnts, ni, torch.int32, dev)
                tq = _wrap_device(torch, qerr, 1, torch.float64, dev)
                dist.all_reduce(ts, group=group)
                dist.all_reduce(tc, group=group)
                dist.all_reduce(tq, group=group)
            return 0
        except Exception as e:  # never unwind through C
            print("sharded_kmeans: all-reduce failed:", e)
            return 1

    cb = _ALLREDUCE_FN(allreduce)
    comm = _KmeansComm(None, cb, int(n_total))
    cent = np.ascontiguousarray(init_centroids, dtype=np.float32).copy()
    assign = np.empty(n, np.int32)
    nassign = np.empty(k, np.int32)
    f, i = C.POINTER(C.c_float), C.POINTER(C.c_int)
    qerr = lib().yb_kmeans_dev(d, n, k, niter, v_shard.data_ptr(), flags | KMEANS_INIT_USER | KMEANS_QUIET,
                               seed, 1, cent.ctypes.data_as(f), None, assign.ctypes.data_as(i),
                               nassign.ctypes.data_as(i), C.cast(C.pointer(comm), C.c_void_p),
                               _stream_ptr(torch))
    if qerr < 0:
        raise RuntimeError("kmeans: clustering failed. Is dataset diverse enough?")
    return cent, qerr, assign, nassign
