"""One-process-per-GPU sharding of the hot path -- a thin caller.

The exchange steps live in the library (csrc/kernels/yb_comm.cu: one NCCL communicator per GPU,
grouped send/recv, all-gather and all-reduce issued from C on the compute stream); this module
only creates the communicator from a torch.distributed group (the 128-byte NCCL id travels
through it) and hands torch tensors' device pointers to the C ABI:

  ShardedKnn / ShardedHamming   database rows split contiguously over the ranks, queries
        replicated; every rank computes a local top-k with GLOBAL ids; QUERY-PARTITIONED exchange:
        rank r receives queries [r*slice, (r+1)*slice) of every rank's lists, merges them by
        (distance, id), one all-gather distributes the merged slices.  The result does not depend
        on the number of ranks.
  sharded_kmeans                points split contiguously; centroids replicated; per iteration
        one grouped all-reduce of (k*d sums, k counts, qerr) between the local accumulation and
        the scaling, through the yb_kmeans_comm_t hook of the C host loop.

`exchange="torch"` keeps the round-1 path (torch.distributed all-gather of every list to every
rank + full merge) as the A/B arm and as the backend-agnostic model the CPU (gloo) tests run.
torch is used for device memory, streams and the rendezvous only.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib


def shard_bounds(n, world):
    """Contiguous shards [n*r/world, n*(r+1)/world) -- the reference's own slicing rule for its
    OpenMP tasks (yael/nn.c:669-670)."""
    return [(n * r // world, n * (r + 1) // world) for r in range(world)]


def query_slices(nq, world):
    """The query partition of the exchange: rank r merges queries [r*slice, min(nq, (r+1)*slice)),
    slice = ceil(nq / world) (yb_comm.cu: knn_sharded_impl)."""
    sl = (nq + world - 1) // world
    return sl, [(min(nq, r * sl), min(nq, (r + 1) * sl)) for r in range(world)]


def _stream_ptr(torch):
    """torch's current stream as the handle the C ABI expects.  Handle 0 would mean "the library's
    own stream" there, so the legacy default stream is passed as cudaStreamLegacy (0x1): our
    kernels and the NCCL collectives torch enqueues then share one stream order."""
    h = torch.cuda.current_stream().cuda_stream
    return C.c_void_p(h if h else 1)


class Comm:
    """The library's communicator (yb_comm) of this rank.  Collective: every rank of the
    torch.distributed group constructs it; rank 0's NCCL unique id is broadcast through the group."""

    def __init__(self, rank, world, device, group=None):
        import torch
        import torch.distributed as dist
        self.rank, self.world, self.handle = rank, world, None
        if world <= 1:
            return
        _lib.require_gpu()
        buf = C.create_string_buffer(128)
        if rank == 0:
            check(lib().yb_comm_unique_id(buf), "yb_comm_unique_id")
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(device)
        dist.broadcast(t, 0, group=group)
        ident = bytes(t.cpu().numpy().tobytes())
        with torch.cuda.device(device):
            self.handle = lib().yb_comm_create(ident, rank, world)
        if not self.handle:
            raise _lib.YaelB200Error("yb_comm_create failed: " + lib().yb_last_error().decode())

    def close(self):
        if self.handle:
            lib().yb_comm_destroy(self.handle)
            self.handle = None


_default_comm = None


def default_comm(rank, world, device):
    """One communicator per process, created on first use."""
    global _default_comm
    if _default_comm is None or _default_comm.world != world:
        _default_comm = Comm(rank, world, device)
    return _default_comm


def allgather_lists(dist, torch, idx, dis, world):
    """[nq][k] per rank -> [world][nq][k] on every rank (one collective per array)."""
    gi = torch.empty((world,) + tuple(idx.shape), dtype=idx.dtype, device=idx.device)
    gd = torch.empty((world,) + tuple(dis.shape), dtype=dis.dtype, device=dis.device)
    dist.all_gather_into_tensor(gi.view(-1), idx.reshape(-1).contiguous())
    # NCCL has no 16-bit integer type: ship the uint16 Hamming distances as raw bytes
    dist.all_gather_into_tensor(gd.view(-1).view(torch.uint8), dis.reshape(-1).contiguous().view(torch.uint8))
    return gi, gd


def partitioned_exchange_model(dist, torch, idx, dis, rank, world, merge):
    """Backend-agnostic model of the library's query-partitioned exchange (used by the CPU tests):
    pad the lists to slice*world queries, every rank takes ITS query slice of every rank's lists,
    merges it with `merge(ids[world][slice][k], dis[world][slice][k]) -> ([slice][k], [slice][k])`,
    and the merged slices are all-gathered.  Returns [nq][k] ids and distances."""
    nq, k = idx.shape
    sl, _ = query_slices(nq, world)
    pad = sl * world - nq
    if pad:
        idx = torch.cat([idx, torch.full((pad, k), -1, dtype=idx.dtype)])
        dis = torch.cat([dis, torch.full((pad, k), float("nan"), dtype=dis.dtype)])
    gi, gd = allgather_lists(dist, torch, idx, dis, world)       # stands in for the all-to-all
    mine_i, mine_d = gi[:, rank * sl:(rank + 1) * sl], gd[:, rank * sl:(rank + 1) * sl]
    mi, md = merge(mine_i, mine_d)
    oi, od = allgather_lists(dist, torch, mi, md, world)
    return oi.reshape(sl * world, k)[:nq], od.reshape(sl * world, k)[:nq]


def _shard_offset(torch, base, rank, world, id_offset):
    """Global id of this rank's first row.  Given explicitly, or derived from the shard sizes of all
    ranks (one all-gather of a scalar; shards may be uneven)."""
    if id_offset is not None:
        return int(id_offset)
    if world == 1:
        return 0
    import torch.distributed as dist
    if not dist.is_initialized():
        raise ValueError("id_offset is required when torch.distributed is not initialised")
    sizes = torch.zeros(world, dtype=torch.int64, device=base.device)
    mine = torch.tensor([base.shape[0]], dtype=torch.int64, device=base.device)
    dist.all_gather_into_tensor(sizes, mine)
    return int(sizes[:rank].sum().item())


class ShardedKnn:
    """Exact L2 k-NN over a database sharded by rows.  `base` is THIS rank's shard (a CUDA
    float32 tensor [rows][d]); `id_offset` the global id of its first row (mandatory for uneven
    shards; the default assumes equal shards)."""

    def __init__(self, base, k, rank=0, world=1, id_offset=None, comm=None, exchange="library"):
        import torch
        self.torch = torch
        assert base.is_cuda and base.dtype == torch.float32 and base.is_contiguous()
        assert exchange in ("library", "torch")
        self.base, self.k, self.rank, self.world = base, k, rank, world
        self.id_offset = _shard_offset(torch, base, rank, world, id_offset)
        self.exchange = exchange
        _lib.require_gpu()
        self.comm = comm
        if world > 1 and exchange == "library" and comm is None:
            self.comm = default_comm(rank, world, base.device)

    def _local_into(self, query, idx, dis, id_offset):
        torch = self.torch
        assert query.is_cuda and query.dtype == torch.float32 and query.is_contiguous()
        nq, d = query.shape
        nb = self.base.shape[0]
        check(lib().yb_knn_l2(nq, nb, d, self.k, self.base.data_ptr(), query.data_ptr(), None,
                              idx.data_ptr(), dis.data_ptr(), int(id_offset),
                              _stream_ptr(torch)), "yb_knn_l2")

    def _exchange_torch(self, buf, nq):
        """Round-1 exchange (A/B arm): ONE all-gather of every rank's [2][nq][k] block to every
        rank, then the merge of all queries on every rank."""
        torch = self.torch
        import torch.distributed as dist
        gbuf = torch.empty((self.world,) + tuple(buf.shape), dtype=torch.int32, device=buf.device)
        dist.all_gather_into_tensor(gbuf.view(-1), buf.view(-1))
        oi = torch.empty((nq, self.k), dtype=torch.int32, device=gbuf.device)
        od = torch.empty((nq, self.k), dtype=torch.float32, device=gbuf.device)
        check(lib().yb_knn_merge_strided(nq, self.k, self.world, gbuf.data_ptr(), gbuf[0, 1].data_ptr(),
                                         2 * nq * self.k, oi.data_ptr(), od.data_ptr(),
                                         _stream_ptr(torch)), "yb_knn_merge_strided")
        return oi, od

    def search(self, query):
        torch = self.torch
        nq, d = query.shape
        assert query.is_cuda and query.dtype == torch.float32 and query.is_contiguous()
        if self.world == 1 or self.exchange == "torch":
            buf = torch.empty((2, nq, self.k), dtype=torch.int32, device=query.device)
            idx, dis = buf[0], buf[1].view(torch.float32)
            self._local_into(query, idx, dis, self.id_offset)
            if self.world == 1:
                return idx, dis
            return self._exchange_torch(buf, nq)
        oi = torch.empty((nq, self.k), dtype=torch.int32, device=query.device)
        od = torch.empty((nq, self.k), dtype=torch.float32, device=query.device)
        check(lib().yb_knn_l2_sharded(self.comm.handle, nq, self.base.shape[0], d, self.k,
                                      self.base.data_ptr(), query.data_ptr(), self.id_offset,
                                      oi.data_ptr(), od.data_ptr(), _stream_ptr(torch)),
              "yb_knn_l2_sharded")
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
        # query-partitioned exchange
        q = torch.from_numpy(query_host).to(self.base.device, non_blocking=True)
        oi = torch.empty((nq, self.k), dtype=torch.int32, device=q.device)
        od = torch.empty((nq, self.k), dtype=torch.float32, device=q.device)
        if self.exchange == "torch":
            buf = torch.empty((2, nq, self.k), dtype=torch.int32, device=q.device)
            check(lib().yb_knn_l2_hostbase(nq, base_host.shape[0], d, self.k, base_host.ctypes.data,
                                           self.base.data_ptr(), q.data_ptr(), buf[0].data_ptr(),
                                           buf[1].data_ptr(), self.id_offset, _stream_ptr(torch)),
                  "yb_knn_l2_hostbase")
            oi, od = self._exchange_torch(buf, nq)
        else:
            check(lib().yb_knn_l2_sharded_hostbase(self.comm.handle, nq, base_host.shape[0], d, self.k,
                                                   base_host.ctypes.data, self.base.data_ptr(), q.data_ptr(),
                                                   self.id_offset, oi.data_ptr(), od.data_ptr(),
                                                   _stream_ptr(torch)), "yb_knn_l2_sharded_hostbase")
        torch.from_numpy(idx_out).copy_(oi, non_blocking=True)
        torch.from_numpy(dis_out).copy_(od, non_blocking=True)
        torch.cuda.synchronize()


class ShardedHamming:
    """nn_hamming over a code database sharded by rows; merged result is bit-identical for any
    number of ranks ((distance, id) order, global ids)."""

    def __init__(self, base, k, rank=0, world=1, id_offset=None, comm=None, exchange="library"):
        import torch
        self.torch = torch
        assert base.is_cuda and base.dtype == torch.uint8 and base.is_contiguous()
        self.base, self.k, self.rank, self.world = base, k, rank, world
        self.id_offset = _shard_offset(torch, base, rank, world, id_offset)
        self.exchange = exchange
        _lib.require_gpu()
        self.comm = comm
        if world > 1 and exchange == "library" and comm is None:
            self.comm = default_comm(rank, world, base.device)

    def search(self, query):
        torch = self.torch
        nq, nc = query.shape
        assert query.is_cuda and query.dtype == torch.uint8 and query.is_contiguous()
        idx = torch.empty((nq, self.k), dtype=torch.int32, device=query.device)
        dis = torch.empty((nq, self.k), dtype=torch.int16, device=query.device)  # uint16 payload
        if self.world > 1 and self.exchange == "library":
            check(lib().yb_nn_hamming_sharded(self.comm.handle, nq, self.base.shape[0], nc, self.k,
                                              self.base.data_ptr(), query.data_ptr(), self.id_offset,
                                              idx.data_ptr(), dis.data_ptr(), _stream_ptr(torch)),
                  "yb_nn_hamming_sharded")
            return idx, dis
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
    _fields_ = [("ctx", C.c_void_p), ("allreduce_sums", _ALLREDUCE_FN), ("n_total", C.c_long),
                ("v_host_all", C.c_void_p), ("rank", C.c_int)]


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


def sharded_kmeans(v_shard, k, niter, init_centroids, n_total, flags=0, seed=0, group=None,
                   comm=None, exchange="library"):
    """Lloyd's k-means (yael/kmeans.c:213-329) on points sharded by rows.  `v_shard`: this
    rank's CUDA float32 [n_local][d]; `init_centroids`: numpy [k][d], identical on every rank
    (KMEANS_INIT_USER).  Returns (centroids, qerr, assign_local, nassign).  exchange="library":
    the all-reduce is issued by the C host loop itself (yb_kmeans_sharded); "torch": the round-1
    path, a ctypes callback into torch.distributed (A/B arm)."""
    import torch
    import torch.distributed as dist
    from .ynumpy import KMEANS_INIT_USER, KMEANS_QUIET
    n, d = v_shard.shape
    dev = v_shard.device
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if exchange == "library":
        if world > 1 and comm is None:
            comm = default_comm(dist.get_rank(group), world, dev)
        cent = np.ascontiguousarray(init_centroids, dtype=np.float32).copy()
        assign = np.empty(n, np.int32)
        nassign = np.empty(k, np.int32)
        f, i = C.POINTER(C.c_float), C.POINTER(C.c_int)
        qerr = lib().yb_kmeans_sharded(comm.handle if comm else None, d, n, int(n_total), k, niter,
                                       v_shard.data_ptr(), flags | KMEANS_QUIET, seed,
                                       cent.ctypes.data_as(f), None, assign.ctypes.data_as(i),
                                       nassign.ctypes.data_as(i), _stream_ptr(torch))
        if qerr < 0:
            raise RuntimeError("kmeans: clustering failed. Is dataset diverse enough?")
        return cent, qerr, assign, nassign

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
    comm = _KmeansComm(None, cb, int(n_total), None, 0)
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
