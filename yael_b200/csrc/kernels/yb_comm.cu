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

// Peer-memory segment of every rank (NVLink / NVSwitch): [4 KB of flags | data].  Mapped into every
// rank's address space -- cudaIpc handles between processes, cudaDeviceEnablePeerAccess inside one --
// so that the exchange steps are plain stores into the owner's memory and a flag barrier, with no
// NCCL kernel on the path (the grouped send / recv all-to-all cost 0.15 ms at 2 ranks but 0.8 - 2.8 ms
// at 8; the peer stores cost the NVLink transfer of 1 MB per peer).
constexpr int kP2pMaxWorld = 16;
constexpr size_t kP2pFlagBytes = 4096;
struct PeerPtrs {
  char *seg[kP2pMaxWorld];
};
struct yb_comm {
  ncclComm_t nccl;
  int rank, world, dev;
  bool p2p;            // the segments below are usable
  bool ipc;            // peers' segments were opened through cudaIpc (close them on destroy)
  PeerPtrs peers;      // seg[r]: base of rank r's segment as THIS rank addresses it
  size_t seg_bytes;    // bytes of data behind the flags
  unsigned epoch;      // barrier generation (same sequence on every rank: calls are collective)
  int *timeout_flag;   // device int: a barrier gave up waiting
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

// a shard with fewer than k rows: its lists [nq][kl] are widened to [nq][k], the tail never selected
__global__ void k_widen_u32(const unsigned *__restrict__ si, const unsigned *__restrict__ sd, long nq, int kl,
                            int k, unsigned *__restrict__ di, unsigned *__restrict__ dd) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq * k) return;
  const long q = t / k;
  const int j = (int)(t - q * k);
  di[t] = j < kl ? si[q * kl + j] : 0xffffffffu;
  dd[t] = j < kl ? sd[q * kl + j] : 0xffffffffu;
}
__global__ void k_widen_u16(const unsigned *__restrict__ si, const unsigned short *__restrict__ sd, long nq, int kl,
                            int k, unsigned *__restrict__ di, unsigned short *__restrict__ dd) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq * k) return;
  const long q = t / k;
  const int j = (int)(t - q * k);
  di[t] = j < kl ? si[q * kl + j] : 0xffffffffu;
  dd[t] = j < kl ? sd[q * kl + j] : (unsigned short)0xffffu;
}

// ---- peer-memory exchange
// rows [p*slice, (p+1)*slice) of my lists go to rank p's segment at [me][slice][k]: ids (4 bytes per
// entry) at off_i, distances (dsz = 4 or 2 bytes per entry) at off_d.  One block column per peer.
__global__ void __launch_bounds__(256)
k_p2p_scatter(PeerPtrs P, int rank, int world, const unsigned *__restrict__ loc_i,
              const unsigned char *__restrict__ loc_d, long slice_entries, int dsz, size_t off_i,
              size_t off_d) {
  const int p = blockIdx.y;
  unsigned *di = (unsigned *)(P.seg[p] + kP2pFlagBytes + off_i) + (size_t)rank * slice_entries;
  const unsigned *si = loc_i + (size_t)p * slice_entries;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < slice_entries; e += (long)gridDim.x * blockDim.x)
    di[e] = si[e];
  const size_t dbytes = (size_t)slice_entries * dsz;
  unsigned char *dd = (unsigned char *)(P.seg[p] + kP2pFlagBytes + off_d) + (size_t)rank * dbytes;
  const unsigned char *sd = loc_d + (size_t)p * dbytes;
  if (((dbytes | (size_t)(uintptr_t)dd | (size_t)(uintptr_t)sd) & 3) == 0) {
    const long n4 = (long)(dbytes >> 2);
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long)gridDim.x * blockDim.x)
      ((unsigned *)dd)[e] = ((const unsigned *)sd)[e];
  } else {
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)dbytes; e += (long)gridDim.x * blockDim.x)
      dd[e] = sd[e];
  }
}

// my merged slice goes to EVERY rank's result block at [me * slice ..): ids at off_i, distances at off_d
__global__ void __launch_bounds__(256)
k_p2p_publish(PeerPtrs P, int rank, int world, const unsigned *__restrict__ mrg_i,
              const unsigned char *__restrict__ mrg_d, long slice_entries, int dsz, size_t off_i,
              size_t off_d) {
  const int p = blockIdx.y;
  unsigned *di = (unsigned *)(P.seg[p] + kP2pFlagBytes + off_i) + (size_t)rank * slice_entries;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < slice_entries; e += (long)gridDim.x * blockDim.x)
    di[e] = mrg_i[e];
  const size_t dbytes = (size_t)slice_entries * dsz;
  unsigned char *dd = (unsigned char *)(P.seg[p] + kP2pFlagBytes + off_d) + (size_t)rank * dbytes;
  if (((dbytes | (size_t)(uintptr_t)dd | (size_t)(uintptr_t)mrg_d) & 3) == 0) {
    const long n4 = (long)(dbytes >> 2);
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long)gridDim.x * blockDim.x)
      ((unsigned *)dd)[e] = ((const unsigned *)mrg_d)[e];
  } else {
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)dbytes; e += (long)gridDim.x * blockDim.x)
      dd[e] = mrg_d[e];
  }
}

// flag barrier over the segments: thread p tells rank p "rank `rank` reached generation `epoch`"
// (release at system scope: the peer stores of the kernels before this one in the stream are visible
// to whoever acquires the flag) and waits for rank p's word in MY flags.  Gives up after
// YAEL_B200_P2P_TIMEOUT_S (default 20 s: a rank that died must not hang the others' GPUs) and traps.
__global__ void k_p2p_barrier(PeerPtrs P, int rank, int world, unsigned epoch, int *timeout_flag,
                              unsigned long long timeout_ns) {
  const int p = threadIdx.x;
  if (p >= world) return;
  __threadfence_system();
  if (p != rank) {
    unsigned *theirs = (unsigned *)P.seg[p] + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
    const unsigned *mine = (const unsigned *)P.seg[rank] + p;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int)(v - epoch) >= 0) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) {  // a peer died: fail loudly (sticky launch failure) instead of hanging
        *timeout_flag = 1;
        __threadfence_system();
        __trap();
      }
    }
  }
  __threadfence_system();
}

static int p2p_barrier(yb_comm *c, cudaStream_t st) {
  c->epoch++;
  static unsigned long long timeout_ns = 0;
  if (!timeout_ns) {
    const char *e = getenv("YAEL_B200_P2P_TIMEOUT_S");
    const double sec = e && atof(e) > 0 ? atof(e) : 20.0;
    timeout_ns = (unsigned long long)(sec * 1e9);
  }
  k_p2p_barrier<<<1, 32, 0, st>>>(c->peers, c->rank, c->world, c->epoch, c->timeout_flag, timeout_ns);
  YB_LAUNCH_CHECK();
  return 0;
}

// the query-partitioned exchange over peer memory.  loc_*: this rank's lists [nq_pad][k]; mrg_*:
// scratch for the merged slice [slice][k]; out_*: [nq][k] result on this rank (NULL: only the slice
// is wanted -- the caller reads mrg_*).  merge(recv_i, recv_d) merges [world][slice][k] into mrg_*.
// layout of the data region of a segment for one exchange (offsets in bytes, 256-aligned): the lists
// received from every rank, the gathered result, and -- so that the sharded calls allocate nothing
// and never touch the host between steps -- this rank's own lists and its merged slice
struct P2pLayout {
  size_t recv_i, recv_d, res_i, res_d, loc_i, loc_d, mrg_i, mrg_d, total;
};
static P2pLayout p2p_layout(long nq_pad, long slice, int k, int dsz) {
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  P2pLayout L;
  const size_t ni = (size_t)nq_pad * k * 4, nd = (size_t)nq_pad * k * dsz;
  const size_t si = (size_t)slice * k * 4, sd = (size_t)slice * k * dsz;
  L.recv_i = 0;
  L.recv_d = up(L.recv_i + ni);
  L.res_i = up(L.recv_d + nd);
  L.res_d = up(L.res_i + ni);
  L.loc_i = up(L.res_d + nd);
  L.loc_d = up(L.loc_i + ni);
  L.mrg_i = up(L.loc_d + nd);
  L.mrg_d = up(L.mrg_i + si);
  L.total = up(L.mrg_d + sd);
  return L;
}

template <typename MergeFn>
static int p2p_exchange(yb_comm *c, long nq, long slice, int k, int dsz, const void *loc_i, const void *loc_d,
                        void *mrg_i, void *mrg_d, void *out_i, void *out_d, MergeFn merge, cudaStream_t st) {
  const int G = c->world;
  const long slice_entries = slice * k, nq_pad = slice * G;
  const P2pLayout L = p2p_layout(nq_pad, slice, k, dsz);
  const size_t recv_i = L.recv_i, recv_d = L.recv_d, res_i = L.res_i, res_d = L.res_d;
  char *mine = c->peers.seg[c->rank] + kP2pFlagBytes;
  int rc;
  dim3 grid((unsigned)((slice_entries + 256 * 8 - 1) / (256 * 8)), (unsigned)G);
  if (grid.x > 64) grid.x = 64;
  if (grid.x < 1) grid.x = 1;
  {
    ProfScope ps(16, st);
    k_p2p_scatter<<<grid, 256, 0, st>>>(c->peers, c->rank, G, (const unsigned *)loc_i,
                                        (const unsigned char *)loc_d, slice_entries, dsz, recv_i, recv_d);
    YB_LAUNCH_CHECK();
    if ((rc = p2p_barrier(c, st))) return rc;
  }
  {
    ProfScope ps(17, st);
    if ((rc = merge(mine + recv_i, mine + recv_d))) return rc;
  }
  if (out_i && out_d) {
    ProfScope ps(16, st);
    k_p2p_publish<<<grid, 256, 0, st>>>(c->peers, c->rank, G, (const unsigned *)mrg_i,
                                        (const unsigned char *)mrg_d, slice_entries, dsz, res_i, res_d);
    YB_LAUNCH_CHECK();
    if ((rc = p2p_barrier(c, st))) return rc;
    YB_CUDA(cudaMemcpyAsync(out_i, mine + res_i, (size_t)nq * k * 4, cudaMemcpyDeviceToDevice, st));
    YB_CUDA(cudaMemcpyAsync(out_d, mine + res_d, (size_t)nq * k * dsz, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

static bool p2p_fits(const yb_comm *c, long nq_pad, long slice, int k, int dsz) {
  if (!c->p2p || getenv("YAEL_B200_NO_P2P")) return false;
  return p2p_layout(nq_pad, slice, k, dsz).total <= c->seg_bytes;
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

static size_t p2p_segment_bytes() {
  size_t mb = 64;
  if (const char *e = getenv("YAEL_B200_P2P_MB")) mb = (size_t)atol(e);
  return mb << 20;
}

static void p2p_reset(yb_comm *c) {
  c->p2p = false;
  c->ipc = false;
  c->seg_bytes = 0;
  c->epoch = 0;
  c->timeout_flag = nullptr;
  for (int i = 0; i < kP2pMaxWorld; i++) c->peers.seg[i] = nullptr;
}

// one process per GPU: allocate my segment, all-gather the cudaIpc handles through NCCL, map the
// peers' segments.  Every rank must end up with the same verdict: the flag is all-reduced (min).
static void p2p_setup_ipc(const NcclApi *N, yb_comm *c) {
  p2p_reset(c);
  if (c->world < 2 || c->world > kP2pMaxWorld || getenv("YAEL_B200_NO_P2P")) return;
  const size_t data = p2p_segment_bytes();
  if (data == 0) return;
  char *seg = nullptr;
  int ok = cudaMalloc(&seg, kP2pFlagBytes + data) == cudaSuccess;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) ok = cudaMemset(seg, 0, kP2pFlagBytes) == cudaSuccess && cudaIpcGetMemHandle(&mine, seg) == cudaSuccess;
  // staging: [world] handles + 1 int verdict, device memory
  char *stage = nullptr;
  const size_t hb = sizeof(cudaIpcMemHandle_t);
  if (cudaMalloc(&stage, hb * (c->world + 1) + 16) != cudaSuccess) {
    cudaGetLastError();
    if (seg) cudaFree(seg);
    return;  // (a rank that cannot even stage: the others time out in NCCL -- cannot happen in practice)
  }
  cudaMemcpy(stage + hb * c->world, &mine, hb, cudaMemcpyHostToDevice);
  bool nccl_ok = N->AllGather(stage + hb * c->world, stage, hb, ybNcclInt8, c->nccl, (cudaStream_t)0) == 0;
  nccl_ok = nccl_ok && cudaStreamSynchronize((cudaStream_t)0) == cudaSuccess;
  cudaIpcMemHandle_t all[kP2pMaxWorld];
  if (nccl_ok) cudaMemcpy(all, stage, hb * c->world, cudaMemcpyDeviceToHost);
  if (ok && nccl_ok) {
    for (int p = 0; p < c->world && ok; p++) {
      if (p == c->rank) {
        c->peers.seg[p] = seg;
        continue;
      }
      void *ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
      } else {
        c->peers.seg[p] = (char *)ptr;
      }
    }
  }
  // agree: everybody or nobody
  int *verdict = (int *)(stage + hb * (c->world + 1));
  int v = ok && nccl_ok ? 1 : 0;
  cudaMemcpy(verdict, &v, sizeof(int), cudaMemcpyHostToDevice);
  if (nccl_ok) {
    nccl_ok = N->AllReduce(verdict, verdict, 1, ybNcclInt32, 3 /* ncclMin */, c->nccl, (cudaStream_t)0) == 0;
    nccl_ok = nccl_ok && cudaStreamSynchronize((cudaStream_t)0) == cudaSuccess;
    cudaMemcpy(&v, verdict, sizeof(int), cudaMemcpyDeviceToHost);
  }
  if (v == 1 && nccl_ok) {
    c->p2p = true;
    c->ipc = true;
    c->seg_bytes = data;
    cudaMalloc(&c->timeout_flag, sizeof(int));
    cudaMemset(c->timeout_flag, 0, sizeof(int));
  } else {
    for (int p = 0; p < c->world; p++)
      if (p != c->rank && c->peers.seg[p]) cudaIpcCloseMemHandle(c->peers.seg[p]);
    if (seg) cudaFree(seg);
    p2p_reset(c);
  }
  cudaFree(stage);
  cudaGetLastError();
}

extern "C" int yb_comm_available(void) { return nccl() != nullptr; }
/* 1 when the exchange steps of this communicator run over peer memory (NVLink stores + flag barrier),
 * 0 when they use the NCCL send / recv path */
extern "C" int yb_comm_p2p(const yb_comm *c) { return c && c->p2p ? 1 : 0; }

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
  p2p_setup_ipc(N, c);  // peer-memory segments (falls back to the NCCL exchange when unavailable)
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
    p2p_reset(out[i]);
  }
  // peer-memory segments inside one process: plain peer access, the same pointers for everybody
  if (ndev >= 2 && ndev <= kP2pMaxWorld && !getenv("YAEL_B200_NO_P2P") && p2p_segment_bytes() > 0) {
    int cur = 0;
    cudaGetDevice(&cur);
    bool ok = true;
    char *seg[kP2pMaxWorld] = {};
    int *tf[kP2pMaxWorld] = {};
    const size_t data = p2p_segment_bytes();
    for (int i = 0; i < ndev && ok; i++) {
      cudaSetDevice(devs[i]);
      for (int j = 0; j < ndev && ok; j++) {
        if (j == i) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, devs[i], devs[j]);
        if (!can) ok = false;
        else {
          cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
          cudaGetLastError();
        }
      }
      if (ok) ok = cudaMalloc(&seg[i], kP2pFlagBytes + data) == cudaSuccess &&
                   cudaMemset(seg[i], 0, kP2pFlagBytes) == cudaSuccess &&
                   cudaMalloc(&tf[i], sizeof(int)) == cudaSuccess && cudaMemset(tf[i], 0, sizeof(int)) == cudaSuccess;
    }
    for (int i = 0; i < ndev; i++) {
      cudaSetDevice(devs[i]);
      cudaDeviceSynchronize();
      if (ok) {
        out[i]->p2p = true;
        out[i]->seg_bytes = data;
        out[i]->timeout_flag = tf[i];
        for (int j = 0; j < ndev; j++) out[i]->peers.seg[j] = seg[j];
      } else {
        if (seg[i]) cudaFree(seg[i]);
        if (tf[i]) cudaFree(tf[i]);
      }
    }
    cudaGetLastError();
    cudaSetDevice(cur);
  }
  return 0;
}

extern "C" void yb_comm_destroy(yb_comm *c) {
  if (!c) return;
  const NcclApi *N = nccl();
  if (c->p2p) {
    cudaDeviceSynchronize();
    if (c->ipc)
      for (int p = 0; p < c->world; p++)
        if (p != c->rank && c->peers.seg[p]) cudaIpcCloseMemHandle(c->peers.seg[p]);
    if (c->peers.seg[c->rank]) cudaFree(c->peers.seg[c->rank]);
    if (c->timeout_flag) cudaFree(c->timeout_flag);
    cudaGetLastError();
  }
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
  if (k <= 0 || nb_local <= 0)
    return fail(3, "sharded k-NN: k = %d, shard rows = %d (every rank needs at least one row)", k, nb_local);
  const int kl = k < nb_local ? k : nb_local;  // the reference only requires k <= n over ALL rows (nn.c:456)
  Guard g;
  cudaStream_t st = stream_of(s);
  const int G = c->world;
  const long slice = ((long)nq + G - 1) / G, nq_pad = slice * G;
  const size_t rows = (size_t)nq_pad * k;
  const bool gather = assign != nullptr && dis != nullptr;
  const bool p2p = p2p_fits(c, nq_pad, slice, k, 4);
  int *loc_i, *rcv_i = nullptr, *mrg_i, *all_i = nullptr;
  float *loc_d, *rcv_d = nullptr, *mrg_d, *all_d = nullptr;
  if (p2p) {  // everything lives in this rank's peer segment: no allocation, no host synchronisation
    const P2pLayout L = p2p_layout(nq_pad, slice, k, 4);
    char *mine = c->peers.seg[c->rank] + kP2pFlagBytes;
    loc_i = (int *)(mine + L.loc_i);
    loc_d = (float *)(mine + L.loc_d);
    mrg_i = (int *)(mine + L.mrg_i);
    mrg_d = (float *)(mine + L.mrg_d);
  } else {    // pooled blocks (the local search reserves the device workspace itself)
    loc_i = (int *)yb_malloc(sizeof(int) * rows * 2);
    loc_d = (float *)(loc_i + rows);
    rcv_i = (int *)yb_malloc(sizeof(int) * rows * 2);
    rcv_d = (float *)(rcv_i + rows);
    mrg_i = (int *)yb_malloc(sizeof(int) * (size_t)slice * k * 2);
    mrg_d = (float *)(mrg_i + (size_t)slice * k);
    all_i = (gather && nq_pad != nq) ? (int *)yb_malloc(sizeof(int) * rows * 2) : nullptr;
    all_d = all_i ? (float *)(all_i + rows) : nullptr;
  }
  int rc;
  int *tmp_i = kl < k ? (int *)yb_malloc(sizeof(int) * (size_t)nq * kl * 2) : nullptr;
  float *tmp_d = tmp_i ? (float *)(tmp_i + (size_t)nq * kl) : nullptr;
  if (base_host)
    rc = yb_knn_l2_hostbase(nq, nb_local, d, kl, base_host, base, query, tmp_i ? tmp_i : loc_i,
                            tmp_d ? tmp_d : loc_d, id_offset, s);
  else
    rc = yb_knn_l2(nq, nb_local, d, kl, base, query, nullptr, tmp_i ? tmp_i : loc_i, tmp_d ? tmp_d : loc_d,
                   id_offset, s);
  if (!rc && tmp_i) {
    k_widen_u32<<<(unsigned)(((long)nq * k + 255) / 256), 256, 0, st>>>((const unsigned *)tmp_i, (const unsigned *)tmp_d,
                                                                     nq, kl, k, (unsigned *)loc_i, (unsigned *)loc_d);
    count_launch();
  }
  if (tmp_i) yb_free(tmp_i);
  if (!rc && nq_pad > nq) {
    const long from = (long)nq * k, to = (long)nq_pad * k;
    k_pad_rows_u32<<<(unsigned)((to - from + 255) / 256), 256, 0, st>>>((unsigned *)loc_i, (unsigned *)loc_d, from, to);
    count_launch();
  }
  if (!rc && p2p) {
    rc = p2p_exchange(c, nq, slice, k, 4, loc_i, loc_d, mrg_i, mrg_d, gather ? assign : nullptr,
                      gather ? dis : nullptr,
                      [&](char *ri, char *rd) {
                        return yb_knn_merge_strided((int)slice, k, G, (const int *)ri, (const float *)rd,
                                                    slice * k, mrg_i, mrg_d, s);
                      },
                      st);
  }
  if (!rc && !p2p) {
    ProfScope ps(16, st);
    rc = all_to_all_rows(N, c, loc_i, rcv_i, slice, sizeof(int) * (size_t)k, st);
    if (!rc) rc = all_to_all_rows(N, c, loc_d, rcv_d, slice, sizeof(float) * (size_t)k, st);
  }
  if (!rc && !p2p) {
    ProfScope ps(17, st);
    rc = yb_knn_merge_strided((int)slice, k, G, rcv_i, rcv_d, slice * k, mrg_i, mrg_d, s);
  }
  if (!rc && gather && !p2p) {
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
  if (!p2p) {
    yb_free(loc_i); yb_free(rcv_i); yb_free(mrg_i);
    if (all_i) yb_free(all_i);
  }
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
  if (k <= 0 || nb_local <= 0)
    return fail(3, "sharded Hamming k-NN: k = %d, shard rows = %d (every rank needs at least one row)", k, nb_local);
  const int kl = k < nb_local ? k : nb_local;
  Guard g;
  cudaStream_t st = stream_of(s);
  const int G = c->world;
  const long slice = ((long)nq + G - 1) / G, nq_pad = slice * G;
  const size_t rows = (size_t)nq_pad * k;
  const bool gather = assign != nullptr && dis != nullptr;
  const bool p2p = p2p_fits(c, nq_pad, slice, k, 2);
  int *loc_i, *rcv_i = nullptr, *mrg_i, *all_i = nullptr;
  uint16_t *loc_d, *rcv_d = nullptr, *mrg_d, *all_d = nullptr;
  if (p2p) {
    const P2pLayout L = p2p_layout(nq_pad, slice, k, 2);
    char *mine = c->peers.seg[c->rank] + kP2pFlagBytes;
    loc_i = (int *)(mine + L.loc_i);
    loc_d = (uint16_t *)(mine + L.loc_d);
    mrg_i = (int *)(mine + L.mrg_i);
    mrg_d = (uint16_t *)(mine + L.mrg_d);
  } else {
    loc_i = (int *)yb_malloc(sizeof(int) * rows + sizeof(uint16_t) * rows);
    loc_d = (uint16_t *)(loc_i + rows);
    rcv_i = (int *)yb_malloc(sizeof(int) * rows + sizeof(uint16_t) * rows);
    rcv_d = (uint16_t *)(rcv_i + rows);
    mrg_i = (int *)yb_malloc((sizeof(int) + sizeof(uint16_t)) * (size_t)slice * k);
    mrg_d = (uint16_t *)(mrg_i + (size_t)slice * k);
    all_i = (gather && nq_pad != nq) ? (int *)yb_malloc(sizeof(int) * rows + sizeof(uint16_t) * rows) : nullptr;
    all_d = all_i ? (uint16_t *)(all_i + rows) : nullptr;
  }
  int *tmp_i = kl < k ? (int *)yb_malloc((sizeof(int) + sizeof(uint16_t)) * (size_t)nq * kl) : nullptr;
  uint16_t *tmp_d = tmp_i ? (uint16_t *)(tmp_i + (size_t)nq * kl) : nullptr;
  int rc = yb_nn_hamming(nq, nb_local, ncodes, kl, base, query, tmp_i ? tmp_i : loc_i, tmp_d ? tmp_d : loc_d,
                         id_offset, s);
  if (!rc && tmp_i) {
    k_widen_u16<<<(unsigned)(((long)nq * k + 255) / 256), 256, 0, st>>>((const unsigned *)tmp_i, tmp_d, nq, kl, k,
                                                                     (unsigned *)loc_i, loc_d);
    count_launch();
  }
  if (tmp_i) yb_free(tmp_i);
  if (!rc && nq_pad > nq) {
    const long from = (long)nq * k, to = (long)nq_pad * k;
    k_pad_rows_u16<<<(unsigned)((to - from + 255) / 256), 256, 0, st>>>((unsigned *)loc_i, loc_d, from, to);
    count_launch();
  }
  if (!rc && p2p) {
    rc = p2p_exchange(c, nq, slice, k, 2, loc_i, loc_d, mrg_i, mrg_d, gather ? assign : nullptr,
                      gather ? dis : nullptr,
                      [&](char *ri, char *rd) {
                        return yb_nn_hamming_merge((int)slice, k, G, (const int *)ri, (const uint16_t *)rd,
                                                   mrg_i, mrg_d, s);
                      },
                      st);
  }
  if (!rc && !p2p) {
    ProfScope ps(16, st);
    rc = all_to_all_rows(N, c, loc_i, rcv_i, slice, sizeof(int) * (size_t)k, st);
    if (!rc) rc = all_to_all_rows(N, c, loc_d, rcv_d, slice, sizeof(uint16_t) * (size_t)k, st);
  }
  if (!rc && !p2p) {
    ProfScope ps(17, st);
    rc = yb_nn_hamming_merge((int)slice, k, G, rcv_i, rcv_d, mrg_i, mrg_d, s);
  }
  if (!rc && gather && !p2p) {
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
  if (!p2p) {
    yb_free(loc_i); yb_free(rcv_i); yb_free(mrg_i);
    if (all_i) yb_free(all_i);
  }
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
