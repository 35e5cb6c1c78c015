// yb_mgpu.cu -- the drop-in layer's own multi-GPU mode: ONE process, one persistent host thread per
// GPU, one NCCL communicator per GPU (ncclCommInitAll), so that an unmodified program linked
// against the library (the reference's progs/knn.c:225, progs/kmeans.c:152) uses the whole box.
//
// The reference parallelises the same calls over OpenMP threads of one process by slicing rows
// [n*r/G, n*(r+1)/G) (yael/nn.c:665-699); here the slices are GPU shards:
//   knn_full / nn_hamming  database rows split, queries replicated; every GPU pulls ITS shard from
//                          the caller's host array (G PCIe links in parallel, overlapped with the
//                          scan), then the query-partitioned exchange of yb_comm.cu; every GPU
//                          writes its slice of the merged result straight into the caller's arrays.
//   kmeans                 points split; every thread runs the C host loop (identical rand_r
//                          draws, identical reduced counts), one grouped all-reduce per iteration.
// Which GPUs: yb_mgpu_set_devices, else YAEL_GPU_DEVICES ("all" or "0,1,.."; default all) -- unless
// the process pinned a device with yb_set_device (one process per GPU, the caller shards) -- and
// only for problems large enough to pay for the exchange (YAEL_B200_MGPU_MIN_WORK overrides).
#include <stdlib.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "yb_common.cuh"
#include "yb_internal.cuh"

extern "C" float yb_kmeans_dev(int d, int n, int k, int niter, const float *v_dev, int flags, long seed,
                               int redo, float *centroids, float *dis, int *assign, int *nassign,
                               const yb_kmeans_comm_t *comm, yb_stream_t s);

namespace {

constexpr int kMaxG = 64;

struct Pool {
  int ndev = 0;
  int devs[kMaxG];
  yb_comm *comms[kMaxG];
  std::vector<std::thread> threads;
  std::mutex m;
  std::condition_variable cv_work, cv_done;
  std::function<int(int)> job;
  long gen = 0;
  int pending = 0;
  int rc[kMaxG];
  bool stop = false;
};

std::mutex g_call_mutex;  // one sharded call at a time (the reference is re-entrant: callers queue)
Pool *g_pool = nullptr;
int g_override_n = 0;  // yb_mgpu_set_devices
int g_override[kMaxG];
thread_local int g_last_used = 1;

void worker(Pool *P, int i) {
  cudaSetDevice(P->devs[i]);
  long seen = 0;
  for (;;) {
    std::function<int(int)> job;
    {
      std::unique_lock<std::mutex> lk(P->m);
      P->cv_work.wait(lk, [&] { return P->stop || P->gen != seen; });
      if (P->stop) return;
      seen = P->gen;
      job = P->job;
    }
    const int rc = job(i);
    {
      std::lock_guard<std::mutex> lk(P->m);
      P->rc[i] = rc;
      if (--P->pending == 0) P->cv_done.notify_all();
    }
  }
}

int run_all(Pool *P, std::function<int(int)> job) {
  {
    std::lock_guard<std::mutex> lk(P->m);
    P->job = std::move(job);
    P->pending = P->ndev;
    P->gen++;
  }
  P->cv_work.notify_all();
  std::unique_lock<std::mutex> lk(P->m);
  P->cv_done.wait(lk, [&] { return P->pending == 0; });
  for (int i = 0; i < P->ndev; i++)
    if (P->rc[i]) return P->rc[i];
  return 0;
}

void pool_destroy() {
  Pool *P = g_pool;
  if (!P) return;
  {
    std::lock_guard<std::mutex> lk(P->m);
    P->stop = true;
  }
  P->cv_work.notify_all();
  for (auto &t : P->threads) t.join();
  int cur = 0;
  cudaGetDevice(&cur);
  for (int i = 0; i < P->ndev; i++) {
    cudaSetDevice(P->devs[i]);
    yb_comm_destroy(P->comms[i]);
  }
  cudaSetDevice(cur);
  delete P;
  g_pool = nullptr;
}

// the devices the next qualifying call uses (n = 1: the single-GPU path)
int resolve(int *devs) {
  if (g_override_n > 0) {
    memcpy(devs, g_override, sizeof(int) * g_override_n);
    return g_override_n;
  }
  if (yb::device_pinned()) return 1;
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess) {
    cudaGetLastError();
    return 1;
  }
  const char *e = getenv("YAEL_GPU_DEVICES");
  int n = 0;
  if (!e || !*e || !strcmp(e, "all")) {
    for (int i = 0; i < have && i < kMaxG; i++) devs[n++] = i;
  } else {
    const char *p = e;
    while (*p && n < kMaxG) {
      char *end;
      long v = strtol(p, &end, 10);
      if (end == p) break;
      if (v >= 0 && v < have) {
        bool dup = false;
        for (int i = 0; i < n; i++) dup |= devs[i] == (int)v;
        if (!dup) devs[n++] = (int)v;
      }
      p = *end ? end + 1 : end;
    }
  }
  return n > 0 ? n : 1;
}

// the pool for the resolved device set (rebuilt when the set changes); NULL: single GPU
Pool *pool_get() {
  int devs[kMaxG];
  const int n = resolve(devs);
  if (n <= 1 || !yb_comm_available()) return nullptr;
  if (g_pool && g_pool->ndev == n && !memcmp(g_pool->devs, devs, sizeof(int) * n)) return g_pool;
  pool_destroy();
  Pool *P = new Pool();
  P->ndev = n;
  memcpy(P->devs, devs, sizeof(int) * n);
  int cur = 0;
  cudaGetDevice(&cur);
  const int rc = yb_comm_create_all(n, devs, P->comms);
  cudaSetDevice(cur);
  if (rc) {
    fprintf(stderr, "yael_b200: multi-GPU mode disabled: %s\n", yb_last_error());
    delete P;
    g_override_n = 1;  // do not retry on every call
    g_override[0] = cur;
    return nullptr;
  }
  for (int i = 0; i < n; i++) P->threads.emplace_back(worker, P, i);
  g_pool = P;
  return P;
}

double min_work(double dflt) {
  const char *e = getenv("YAEL_B200_MGPU_MIN_WORK");
  return e && *e ? atof(e) : dflt;
}

}  // namespace

extern "C" int yb_mgpu_set_devices(int n, const int *devs) {
  std::lock_guard<std::mutex> lk(g_call_mutex);
  if (n < 0 || n > kMaxG) return yb::fail(3, "yb_mgpu_set_devices: %d devices", n);
  g_override_n = n;
  for (int i = 0; i < n; i++) g_override[i] = devs[i];
  if (n == 1) pool_destroy();
  return 0;
}

extern "C" int yb_mgpu_device_count(void) {
  std::lock_guard<std::mutex> lk(g_call_mutex);
  int devs[kMaxG];
  const int n = resolve(devs);
  return (n > 1 && yb_comm_available()) ? n : 1;
}

extern "C" int yb_mgpu_last_used(void) { return g_last_used; }

// knn_full (yael/nn.c:451-525), L2, host pointers
extern "C" int yb_mgpu_knn_full(int nq, int nb, int d, int k, const float *base, const float *query,
                                int *assign, float *dis) {
  g_last_used = 1;
  if ((double)nq * nb * d < min_work(2e11)) return -1;
  std::lock_guard<std::mutex> lk(g_call_mutex);
  Pool *P = pool_get();
  if (!P) return -1;
  const int G = P->ndev;
  if (nb / G < 4096) return -1;
  const long slice = ((long)nq + G - 1) / G;
  const int rc = run_all(P, [=](int i) -> int {
    const long lo = (long)nb * i / G, hi = (long)nb * (i + 1) / G;
    const int nbl = (int)(hi - lo);
    float *bdev = (float *)yb_malloc(sizeof(float) * (size_t)nbl * d);
    float *qdev = (float *)yb_malloc(sizeof(float) * (size_t)nq * d);
    int r = yb_h2d(qdev, query, sizeof(float) * (size_t)nq * d, nullptr);
    const long q0 = slice * i < nq ? slice * i : nq;
    if (!r)
      r = yb::knn_sharded_impl(P->comms[i], nq, nbl, d, k, bdev, base + (size_t)lo * d, qdev, (int)lo,
                               nullptr, nullptr, assign + (size_t)q0 * k, dis + (size_t)q0 * k, nullptr);
    if (!r) r = yb_sync(nullptr);
    yb_free(bdev);
    yb_free(qdev);
    return r;
  });
  if (!rc) g_last_used = G;
  return rc;
}

// nn_hamming (new; semantics of yael/hamming.c:177-219 + select), host pointers
extern "C" int yb_mgpu_nn_hamming(int nq, int nb, int ncodes, int k, const uint8_t *base,
                                  const uint8_t *query, int *assign, uint16_t *dis) {
  g_last_used = 1;
  if ((double)nq * nb * ncodes < min_work(1.6e11)) return -1;
  std::lock_guard<std::mutex> lk(g_call_mutex);
  Pool *P = pool_get();
  if (!P) return -1;
  const int G = P->ndev;
  if (nb / G < 4096) return -1;
  const long slice = ((long)nq + G - 1) / G;
  const int rc = run_all(P, [=](int i) -> int {
    const long lo = (long)nb * i / G, hi = (long)nb * (i + 1) / G;
    const int nbl = (int)(hi - lo);
    uint8_t *bdev = (uint8_t *)yb_malloc((size_t)nbl * ncodes);
    uint8_t *qdev = (uint8_t *)yb_malloc((size_t)nq * ncodes);
    int r = yb_h2d(bdev, base + (size_t)lo * ncodes, (size_t)nbl * ncodes, nullptr);
    if (!r) r = yb_h2d(qdev, query, (size_t)nq * ncodes, nullptr);
    const long q0 = slice * i < nq ? slice * i : nq;
    if (!r)
      r = yb::hamming_sharded_impl(P->comms[i], nq, nbl, ncodes, k, bdev, qdev, (int)lo, nullptr, nullptr,
                                   assign + (size_t)q0 * k, dis + (size_t)q0 * k, nullptr);
    if (!r) r = yb_sync(nullptr);
    yb_free(bdev);
    yb_free(qdev);
    return r;
  });
  if (!rc) g_last_used = G;
  return rc;
}

// kmeans (yael/kmeans.c:332-447), host pointers; *qerr_out receives the return value of kmeans()
extern "C" int yb_mgpu_kmeans(int d, int n, int k, int niter, const float *v, int flags, long seed,
                              int redo, float *centroids, float *dis, int *assign, int *nassign,
                              float *qerr_out) {
  g_last_used = 1;
  if ((double)n * k * d < min_work(2e11)) return -1;
  if (flags & (0x200000 | 0x400000)) return -1;  // KMEANS_L1 / KMEANS_CHI2: refused by the one-GPU path
  std::lock_guard<std::mutex> lk(g_call_mutex);
  Pool *P = pool_get();
  if (!P) return -1;
  const int G = P->ndev;
  if (n / G < 4096) return -1;
  if (seed == 0) seed = lrand48();  // yael/kmeans.c:379-380, drawn ONCE for all ranks
  const bool user_init = (flags & 0x100000) != 0;
  // KMEANS_INIT_USER: `centroids` is input and output; the other ranks read private copies
  std::vector<std::vector<float>> priv(G);
  if (user_init)
    for (int i = 1; i < G; i++) priv[i].assign(centroids, centroids + (size_t)k * d);
  float ret[kMaxG];
  const int rc = run_all(P, [&](int i) -> int {
    const long lo = (long)n * i / G, hi = (long)n * (i + 1) / G;
    const int nl = (int)(hi - lo);
    float *vdev = (float *)yb_malloc(sizeof(float) * (size_t)nl * d);
    int r = yb_h2d(vdev, v + (size_t)lo * d, sizeof(float) * (size_t)nl * d, nullptr);
    if (!r) {
      yb_kmeans_comm_t hook = yb::kmeans_hook(P->comms[i], n, v);
      float *cent = i == 0 ? centroids : (user_init ? priv[i].data() : nullptr);
      ret[i] = yb_kmeans_dev(d, nl, k, niter, vdev, flags, seed, redo, cent, dis ? dis + lo : nullptr,
                             assign ? assign + lo : nullptr, i == 0 ? nassign : nullptr, &hook, nullptr);
    }
    yb_free(vdev);
    return r;
  });
  if (!rc) {
    *qerr_out = ret[0];
    g_last_used = G;
  }
  return rc;
}
