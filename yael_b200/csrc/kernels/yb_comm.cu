// yb_comm.cu -- the exchange steps of the sharded hot path over NCCL, in the
// library itself: one communicator per GPU, usable from one process per GPU (torchrun: the id is
// broadcast by the caller) and from one host thread per GPU inside a single process (the drop-in
// layer's own multi-GPU mode, yb_mgpu.cu: ncclCommInitAll).
//
//   k-NN / Hamming   database rows split over the ranks, queries replicated.  Every rank scans its
//                    shard (global ids), then a QUERY-PARTITIONED exchange: rank r owns queries
//                    [r*slice, (r+1)*slice), receives that slice of every rank's lists
//                    (all-to-all: grouped ncclSend/ncclRecv, nq*k*8/G bytes per peer), merges them
//                    by (distance, id) -- 1/G of the merge work per rank -- and one all-gather
//                    distributes the merged slices.  An all-gather of the full lists (round 1)
//                    lands G times the bytes on every rank and makes every rank merge every query.
//   k-means          points split over the ranks; per iteration the all-reduces of the k*d float
//                    sums, the k int counts and the double qerr go out as ONE NCCL group (a single
//                    aggregated launch), issued from the C host loop on the compute stream.
//
// NCCL is resolved at run time (dlopen "libnccl.so.2": inside a PyTorch process that is the copy
// torch already loaded, so both use one library; in a plain C program the system copy), so the
// library still loads on a machine without NCCL.  The reference has no counterpart (a single
// process with OpenMP threads, yael/nn.c:665-699); the slicing rule is its own.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "yb_common.cuh"
#include "yb_internal.cuh"

// ---- the slice of the NCCL API used here (nccl.h is not needed to build)
extern "C" {
typedef struct ncclComm *ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;  // ncclSuccess = 0
enum { ybNcclInt8 = 0, ybNcclUint8 = 1, ybNcclInt32 = 2, ybNcclFloat32 = 7, ybNcclFloat64 = 8 };
enum { ybNcclSum = 0 };
}

namespace {

struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi g_nccl;
std::mutex g_nccl_mutex;

const NcclApi *nccl() {
  std::lock_guard<std::mutex> lk(g_nccl_mutex);
  if (g_nccl.ok) return &g_nccl;
  if (!g_nccl.h) {
    const char *names[] = {getenv("YAEL_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      if (!n || !*n) continue;
      g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (g_nccl.h) break;
    }
  }
  if (!g_nccl.h) return nullptr;
#define YB_SYM(field, name) *(void **)(&g_nccl.field) = dlsym(g_nccl.h, name)
  YB_SYM(GetUniqueId, "ncclGetUniqueId");
  YB_SYM(CommInitRank, "ncclCommInitRank");
  YB_SYM(CommInitAll, "ncclCommInitAll");
  YB_SYM(CommDestroy, "ncclCommDestroy");
  YB_SYM(AllReduce, "ncclAllReduce");
  YB_SYM(AllGather, "ncclAllGather");
  YB_SYM(Send, "ncclSend");
  YB_SYM(Recv, "ncclRecv");
  YB_SYM(GroupStart, "ncclGroupStart");
  YB_SYM(GroupEnd, "ncclGroupEnd");
  YB_SYM(GetErrorString, "ncclGetErrorString");
#undef YB_SYM
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommInitAll && g_nccl.CommDestroy &&
              g_nccl.AllReduce && g_nccl.AllGather && g_nccl.Send && g_nccl.Recv && g_nccl.GroupStart &&
              g_nccl.GroupEnd;
  return g_nccl.ok ? &g_nccl : nullptr;
}

}  // namespace

struct yb_comm {
  ncclComm_t nccl;
  int rank, world, dev;
};

#define YB_NCCL(expr)                                                                        \
  do {                                                                                       \
    ncclResult_t _r = (expr);                                                                \
    if (_r != 0)                                                                             \
      return yb::fail(8, "%s:%d: %s -> NCCL error %d (%s)", __FILE__, __LINE__, #expr, _r,   \
                      N->GetErrorString ? N->GetErrorString(_r) : "?");                      \
  } while (0)

namespace yb {

// rows [nq, nq_pad) of a result block: never selected (id -1, distance bits all ones)
__global__ void k_pad_rows_u32(unsigned *__restrict__ ids, unsigned *__restrict__ dis, long from, long to) {
  const long t = from + (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < to) {
    ids[t] = 0xffffffffu;
    dis[t] = 0xffffffffu;
  }
}
__global__ void k_pad_rows_u16(unsigned *__restrict__ ids, unsigned short *__restrict__ dis, long from, long to) {
  const long t = from + (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < to) {
    ids[t] = 0xffffffffu;
    dis[t] = 0xffffu;
  }
}

// every rank sends rows [p*slice, (p+1)*slice) of `send` to rank p and receives ITS rows from every
// rank into recv[p][slice] (row = row_bytes bytes)
static int all_to_all_rows(const NcclApi *N, yb_comm *c, const void *send, void *recv, long slice,
                           size_t row_bytes, cudaStream_t st) {
  const size_t chunk = (size_t)slice * row_bytes;
  YB_NCCL(N->GroupStart());
  for (int p = 0; p < c->world; p++) {
    YB_NCCL(N->Send((const char *)send + (size_t)p * chunk, chunk, ybNcclInt8, p, c->nccl, st));
    YB_NCCL(N->Recv((char *)recv + (size_t)p * chunk, chunk, ybNcclInt8, p, c->nccl, st));
  }
  YB_NCCL(N->GroupEnd());
  return 0;
}

}  // namespace yb

using namespace yb;

extern "C" int yb_comm_available(void) { return nccl() != nullptr; }

extern "C" int yb_comm_unique_id(void *id128) {
  const NcclApi *N = nccl();
  if (!N) return fail(8, "NCCL is not available (libnccl.so.2 could not be loaded: %s)", dlerror());
  ncclUniqueId id;
  YB_NCCL(N->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

// communicator of the CURRENT device for rank `rank` of `world` (collective: every rank calls it)
extern "C" yb_comm *yb_comm_create(const void *id128, int rank, int world) {
  const NcclApi *N = nccl();
  if (!N) {
    fail(8, "NCCL is not available (libnccl.so.2 could not be loaded)");
    return nullptr;
  }
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  yb_comm *c = new yb_comm();
  c->rank = rank;
  c->world = world;
  c->dev = dev_index();
  ncclResult_t r = N->CommInitRank(&c->nccl, world, id, rank);
  if (r != 0) {
    fail(8, "ncclCommInitRank failed with %d (%s)", r, N->GetErrorString ? N->GetErrorString(r) : "?");
    delete c;
    return nullptr;
  }
  return c;
}

// one communicator per listed device, all in THIS process (out[i] belongs to devs[i])
extern "C" int yb_comm_create_all(int ndev, const int *devs, yb_comm **out) {
  const NcclApi *N = nccl();
  if (!N) return fail(8, "NCCL is not available (libnccl.so.2 could not be loaded)");
  ncclComm_t comms[64];
  if (ndev < 1 || ndev > 64) return fail(3, "yb_comm_create_all: %d devices", ndev);
  YB_NCCL(N->CommInitAll(comms, ndev, devs));
  for (int i = 0; i < ndev; i++) {
    out[i] = new yb_comm();
    out[i]->nccl = comms[i];
    out[i]->rank = i;
    out[i]->world = ndev;
    out[i]->dev = devs[i];
  }
  return 0;
}

extern "C" void yb_comm_destroy(yb_comm *c) {
  if (!c) return;
  const NcclApi *N = nccl();
  if (N && c->nccl) N->CommDestroy(c->nccl);
  delete c;
}

extern "C" int yb_comm_rank(const yb_comm *c) { return c ? c->rank : 0; }
extern "C" int yb_comm_world(const yb_comm *c) { return c ? c->world : 1; }

extern "C" int yb_comm_allreduce_f32(yb_comm *c, float *buf, long n, yb_stream_t s) {
  if (!c || c->world <= 1 || n <= 0) return 0;
  const NcclApi *N = nccl();
  if (!N) return fail(8, "NCCL is not available");
  YB_NCCL(N->AllReduce(buf, buf, (size_t)n, ybNcclFloat32, ybNcclSum, c->nccl, stream_of(s)));
  return 0;
}

extern "C" int yb_comm_allgather(yb_comm *c, const void *send, void *recv, long bytes_per_rank,
                                 yb_stream_t s) {
  if (!c || bytes_per_rank <= 0) return 0;
  const NcclApi *N = nccl();
  if (!N) return fail(8, "NCCL is not available");
  YB_NCCL(N->AllGather(send, recv, (size_t)bytes_per_rank, ybNcclInt8, c->nccl, stream_of(s)));
  return 0;
}

// ------------------------------------------------------------------ sharded exact k-NN
// SPMD: every rank calls with ITS shard base[nb_local][d] (global id of its first row: id_offset) and
// the same queries.  assign / dis (device, [nq][k], may be NULL) receive the merged result of the
// whole database on every rank; slice_assign / slice_dis (HOST, may be NULL) receive only the rows
// this rank merged, i.e. queries [rank*slice, (rank+1)*slice) -- what the in-process multi-GPU mode
// wants: every GPU writes its share of the caller's output arrays and nothing is gathered.
// base_host != NULL: the shard is still in host memory and base is device scratch for it (the
// transfer overlaps the scan, yb_knn_l2_hostbase).  Phases 16 (exchange) and 17 (slice merge).
namespace yb {
int knn_sharded_impl(yb_comm *c, int nq, int nb_local, int d, int k, float *base,
                     const float *base_host, const float *query, int id_offset, int *assign,
                     float *dis, int *slice_assign, float *slice_dis, yb_stream_t s) {
  const NcclApi *N = nccl();
  if (!N) return fail(8, "NCCL is not available");
  if (k <= 0 || k > nb_local)
    return fail(3, "sharded k-NN: every shard needs at least k rows (k=%d, shard rows=%d)", k, nb_local);
  Guard g;
  cudaStream_t st = stream_of(s);
  const int G = c->world;
  const long slice = ((long)nq + G - 1) / G, nq_pad = slice * G;
  const size_t rows = (size_t)nq_pad * k;
  // pooled blocks (the local search reserves the device workspace itself)
  int *loc_i = (int *)yb_malloc(sizeof(int) * rows * 2);
  float *loc_d = (float *)(loc_i + rows);
  int *rcv_i = (int *)yb_malloc(sizeof(int) * rows * 2);
  float *rcv_d = (float *)(rcv_i + rows);
  int *mrg_i = (int *)yb_malloc(sizeof(int) * (size_t)slice * k * 2);
  float *mrg_d = (float *)(mrg_i + (size_t)slice * k);
  const bool gather = assign != nullptr && dis != nullptr;
  int *all_i = (gather && nq_pad != nq) ? (int *)yb_malloc(sizeof(int) * rows * 2) : nullptr;
  float *all_d = all_i ? (float *)(all_i + rows) : nullptr;
  int rc;
  if (base_host)
    rc = yb_knn_l2_hostbase(nq, nb_local, d, k, base_host, base, query, loc_i, loc_d, id_offset, s);
  else
    rc = yb_knn_l2(nq, nb_local, d, k, base, query, nullptr, loc_i, loc_d, id_offset, s);
  if (!rc && nq_pad > nq) {
    const long from = (long)nq * k, to = (long)nq_pad * k;
    k_pad_rows_u32<<<(unsigned)((to - from + 255) / 256), 256, 0, st>>>((unsigned *)loc_i, (unsigned *)loc_d, from, to);
    count_launch();
  }
  if (!rc) {
    ProfScope ps(16, st);
    rc = all_to_all_rows(N, c, loc_i, rcv_i, slice, sizeof(int) * (size_t)k, st);
    if (!rc) rc = all_to_all_rows(N, c, loc_d, rcv_d, slice, sizeof(float) * (size_t)k, st);
  }
  if (!rc) {
    ProfScope ps(17, st);
    rc = yb_knn_merge_strided((int)slice, k, G, rcv_i, rcv_d, slice * k, mrg_i, mrg_d, s);
  }
  if (!rc && gather) {
    ProfScope ps(16, st);
    if (nq_pad == nq) {  // straight into the caller's arrays
      rc = yb_comm_allgather(c, mrg_i, assign, (long)(sizeof(int) * (size_t)slice * k), s);
      if (!rc) rc = yb_comm_allgather(c, mrg_d, dis, (long)(sizeof(float) * (size_t)slice * k), s);
    } else {
      rc = yb_comm_allgather(c, mrg_i, all_i, (long)(sizeof(int) * (size_t)slice * k), s);
      if (!rc) rc = yb_comm_allgather(c, mrg_d, all_d, (long)(sizeof(float) * (size_t)slice * k), s);
      if (!rc) {
        cudaMemcpyAsync(assign, all_i, sizeof(int) * (size_t)nq * k, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(dis, all_d, sizeof(float) * (size_t)nq * k, cudaMemcpyDeviceToDevice, st);
      }
    }
  }
  if (!rc && slice_assign && slice_dis) {
    const long q0 = (long)c->rank * slice, q1 = q0 + slice < nq ? q0 + slice : nq;
    if (q1 > q0) {
      cudaMemcpyAsync(slice_assign, mrg_i, sizeof(int) * (size_t)(q1 - q0) * k, cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(slice_dis, mrg_d, sizeof(float) * (size_t)(q1 - q0) * k, cudaMemcpyDeviceToHost, st);
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) rc = fail(1, "sharded k-NN: device to host copy failed");
  }
  yb_free(loc_i); yb_free(rcv_i); yb_free(mrg_i);
  if (all_i) yb_free(all_i);
  return rc;
}
}  // namespace yb

extern "C" int yb_knn_l2_sharded(yb_comm *c, int nq, int nb_local, int d, int k, const float *base,
                                 const float *query, int id_offset, int *assign, float *dis,
                                 yb_stream_t s) {
  if (nq <= 0) return 0;
  if (!c || c->world <= 1)
    return yb_knn_l2(nq, nb_local, d, k, base, query, nullptr, assign, dis, id_offset, s);
  return knn_sharded_impl(c, nq, nb_local, d, k, (float *)base, nullptr, query, id_offset, assign, dis,
                          nullptr, nullptr, s);
}

extern "C" int yb_knn_l2_sharded_hostbase(yb_comm *c, int nq, int nb_local, int d, int k,
                                          const float *base_host, float *base_dev, const float *query,
                                          int id_offset, int *assign, float *dis, yb_stream_t s) {
  if (nq <= 0) return 0;
  if (!c || c->world <= 1)
    return yb_knn_l2_hostbase(nq, nb_local, d, k, base_host, base_dev, query, assign, dis, id_offset, s);
  return knn_sharded_impl(c, nq, nb_local, d, k, base_dev, base_host, query, id_offset, assign, dis,
                          nullptr, nullptr, s);
}

// the same for the Hamming k-NN (uint16 distances): merged result bit-identical for any rank count
namespace yb {
int hamming_sharded_impl(yb_comm *c, int nq, int nb_local, int ncodes, int k, const uint8_t *base,
                         const uint8_t *query, int id_offset, int *assign, uint16_t *dis,
                         int *slice_assign, uint16_t *slice_dis, yb_stream_t s) {
  const NcclApi *N = nccl();
  if (!N) return fail(8, "NCCL is not available");
  if (k <= 0 || k > nb_local)
    return fail(3, "sharded Hamming k-NN: every shard needs at least k rows (k=%d, shard rows=%d)", k, nb_local);
  Guard g;
  cudaStream_t st = stream_of(s);
  const int G = c->world;
  const long slice = ((long)nq + G - 1) / G, nq_pad = slice * G;
  const size_t rows = (size_t)nq_pad * k;
  int *loc_i = (int *)yb_malloc(sizeof(int) * rows + sizeof(uint16_t) * rows);
  uint16_t *loc_d = (uint16_t *)(loc_i + rows);
  int *rcv_i = (int *)yb_malloc(sizeof(int) * rows + sizeof(uint16_t) * rows);
  uint16_t *rcv_d = (uint16_t *)(rcv_i + rows);
  int *mrg_i = (int *)yb_malloc((sizeof(int) + sizeof(uint16_t)) * (size_t)slice * k);
  uint16_t *mrg_d = (uint16_t *)(mrg_i + (size_t)slice * k);
  const bool gather = assign != nullptr && dis != nullptr;
  int *all_i = (gather && nq_pad != nq) ? (int *)yb_malloc(sizeof(int) * rows + sizeof(uint16_t) * rows) : nullptr;
  uint16_t *all_d = all_i ? (uint16_t *)(all_i + rows) : nullptr;
  int rc = yb_nn_hamming(nq, nb_local, ncodes, k, base, query, loc_i, loc_d, id_offset, s);
  if (!rc && nq_pad > nq) {
    const long from = (long)nq * k, to = (long)nq_pad * k;
    k_pad_rows_u16<<<(unsigned)((to - from + 255) / 256), 256, 0, st>>>((unsigned *)loc_i, loc_d, from, to);
    count_launch();
  }
  if (!rc) {
    ProfScope ps(16, st);
    rc = all_to_all_rows(N, c, loc_i, rcv_i, slice, sizeof(int) * (size_t)k, st);
    if (!rc) rc = all_to_all_rows(N, c, loc_d, rcv_d, slice, sizeof(uint16_t) * (size_t)k, st);
  }
  if (!rc) {
    ProfScope ps(17, st);
    rc = yb_nn_hamming_merge((int)slice, k, G, rcv_i, rcv_d, mrg_i, mrg_d, s);
  }
  if (!rc && gather) {
    ProfScope ps(16, st);
    if (nq_pad == nq) {
      rc = yb_comm_allgather(c, mrg_i, assign, (long)(sizeof(int) * (size_t)slice * k), s);
      if (!rc) rc = yb_comm_allgather(c, mrg_d, dis, (long)(sizeof(uint16_t) * (size_t)slice * k), s);
    } else {
      rc = yb_comm_allgather(c, mrg_i, all_i, (long)(sizeof(int) * (size_t)slice * k), s);
      if (!rc) rc = yb_comm_allgather(c, mrg_d, all_d, (long)(sizeof(uint16_t) * (size_t)slice * k), s);
      if (!rc) {
        cudaMemcpyAsync(assign, all_i, sizeof(int) * (size_t)nq * k, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(dis, all_d, sizeof(uint16_t) * (size_t)nq * k, cudaMemcpyDeviceToDevice, st);
      }
    }
  }
  if (!rc && slice_assign && slice_dis) {
    const long q0 = (long)c->rank * slice, q1 = q0 + slice < nq ? q0 + slice : nq;
    if (q1 > q0) {
      cudaMemcpyAsync(slice_assign, mrg_i, sizeof(int) * (size_t)(q1 - q0) * k, cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(slice_dis, mrg_d, sizeof(uint16_t) * (size_t)(q1 - q0) * k, cudaMemcpyDeviceToHost, st);
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) rc = fail(1, "sharded Hamming k-NN: device to host copy failed");
  }
  yb_free(loc_i); yb_free(rcv_i); yb_free(mrg_i);
  if (all_i) yb_free(all_i);
  return rc;
}
}  // namespace yb

extern "C" int yb_nn_hamming_sharded(yb_comm *c, int nq, int nb_local, int ncodes, int k,
                                     const uint8_t *base, const uint8_t *query, int id_offset,
                                     int *assign, uint16_t *dis, yb_stream_t s) {
  if (nq <= 0) return 0;
  if (!c || c->world <= 1) return yb_nn_hamming(nq, nb_local, ncodes, k, base, query, assign, dis, id_offset, s);
  return hamming_sharded_impl(c, nq, nb_local, ncodes, k, base, query, id_offset, assign, dis, nullptr,
                              nullptr, s);
}

// ------------------------------------------------------------------ sharded k-means
// all-reduce hook of the C host loop (yb_kmeans_comm_t): the three reductions (float sums, int
// counts, double qerr) are issued as ONE NCCL group on the compute stream, which NCCL aggregates
// into a single launch -- no host round trip, no Python.
static int km_allreduce(void *ctx, float *sums, long n_float, int *nassign, long n_int, double *qerr,
                        yb_stream_t s) {
  yb_comm *c = (yb_comm *)ctx;
  if (!c || c->world <= 1) return 0;
  const NcclApi *N = nccl();
  if (!N) return fail(8, "NCCL is not available");
  cudaStream_t st = stream_of(s);
  ProfScope ps(16, st);
  YB_NCCL(N->GroupStart());
  YB_NCCL(N->AllReduce(sums, sums, (size_t)n_float, ybNcclFloat32, ybNcclSum, c->nccl, st));
  YB_NCCL(N->AllReduce(nassign, nassign, (size_t)n_int, ybNcclInt32, ybNcclSum, c->nccl, st));
  YB_NCCL(N->AllReduce(qerr, qerr, 1, ybNcclFloat64, ybNcclSum, c->nccl, st));
  YB_NCCL(N->GroupEnd());
  return 0;
}

// kmeans (yael/kmeans.c:332-447) on points sharded by rows over the ranks of `c`: v_dev is THIS
// rank's shard [n_local][d], n_total the global point count, centroids the k initial centroids
// (KMEANS_INIT_USER, identical on every rank) and the result.  assign / dis (host, may be NULL)
// receive this rank's points' assignment.
extern "C" float yb_kmeans_sharded(yb_comm *c, int d, int n_local, long n_total, int k, int niter,
                                   const float *v_dev, int flags, long seed, float *centroids,
                                   float *dis, int *assign, int *nassign, yb_stream_t s) {
  yb_kmeans_comm_t hook;
  hook.ctx = c;
  hook.allreduce_sums = km_allreduce;
  hook.n_total = n_total;
  hook.v_host_all = nullptr;
  hook.rank = c ? c->rank : 0;
  return yb_kmeans_dev(d, n_local, k, niter, v_dev, flags | 0x100000 /* KMEANS_INIT_USER */, seed, 1,
                       centroids, dis, assign, nassign, (c && c->world > 1) ? &hook : nullptr, s);
}

namespace yb {
yb_kmeans_comm_t kmeans_hook(yb_comm *c, long n_total, const float *v_host_all) {
  yb_kmeans_comm_t hook;
  hook.ctx = c;
  hook.allreduce_sums = km_allreduce;
  hook.n_total = n_total;
  hook.v_host_all = v_host_all;
  hook.rank = c ? c->rank : 0;
  return hook;
}
}  // namespace yb
