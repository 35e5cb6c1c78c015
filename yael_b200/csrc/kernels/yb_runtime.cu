// yb_runtime.cu -- device context, workspace cache, error reporting.
//
// The reference has no device or global state (doc/index.rst:55-58: every function is
// re-entrant); the library therefore keeps one lazily created context per device behind a
// mutex and never requires an init call.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "yb_common.cuh"
#include "yb_internal.cuh"

namespace {

constexpr int kMaxDev = 64;

struct DevState {
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // host->device feeds that overlap the compute stream
  void *scratch = nullptr;
  size_t scratch_bytes = 0;
  cudaEvent_t scratch_ev = nullptr;  // completion of the last op that used the scratch
  cudaStream_t scratch_stream = nullptr;
  bool scratch_used = false;
  int sms = 0;
};

// A pooled block remembers where the streams of its device stood when it was freed: whoever gets
// it next waits (on the host) for those positions, so a kernel that was still reading the block
// on one stream can never see the writes of its next user on another.
struct PoolBlock {
  void *p;
  size_t bytes;
  int dev;
  bool free_;
  cudaEvent_t ev[3];
  int nev;
};
std::vector<PoolBlock> g_pool;
std::mutex g_pool_mutex;

DevState g_dev[kMaxDev];
// one lock per device: calls on different devices (one host thread per GPU) run concurrently,
// calls on the same device serialise their host-side submission
std::recursive_mutex g_dev_mutex[kMaxDev];
cudaStream_t g_last_caller_stream[kMaxDev];  // the last caller-provided stream seen per device
thread_local char g_err[512] = "";
std::atomic<long> g_launches{0};
std::atomic<bool> g_pinned{false};

int cur_dev() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) return 0;
  if (d >= kMaxDev) {
    fprintf(stderr, "yael_b200: device index %d exceeds the supported %d devices\n", d, kMaxDev);
    abort();
  }
  return d;
}


void pool_trim(int dev) {
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  cudaDeviceSynchronize();
  for (size_t i = 0; i < g_pool.size();) {
    if (g_pool[i].free_ && (dev < 0 || g_pool[i].dev == dev)) {
      for (int e = 0; e < 3; e++)
        if (g_pool[i].ev[e]) cudaEventDestroy(g_pool[i].ev[e]);
      cudaFree(g_pool[i].p);
      g_pool[i] = g_pool.back();
      g_pool.pop_back();
    } else {
      i++;
    }
  }
}

}  // namespace

namespace yb {

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

Guard::Guard() : dev(cur_dev()) { g_dev_mutex[dev].lock(); }
Guard::~Guard() { g_dev_mutex[dev].unlock(); }

int dev_index() { return cur_dev(); }
bool device_pinned() { return g_pinned.load(); }

cudaStream_t copy_stream() {
  const int dev = cur_dev();
  DevState &st = g_dev[dev];
  std::lock_guard<std::recursive_mutex> lk(g_dev_mutex[dev]);
  if (!st.copy_stream) cudaStreamCreateWithFlags(&st.copy_stream, cudaStreamNonBlocking);
  return st.copy_stream;
}

cudaStream_t stream_of(yb_stream_t s) {
  const int dev = cur_dev();
  if (s) {
    g_last_caller_stream[dev] = (cudaStream_t)s;
    return (cudaStream_t)s;
  }
  DevState &st = g_dev[dev];
  if (!st.stream) {  // lazy creation under the device lock (callers may not hold it)
    std::lock_guard<std::recursive_mutex> lk(g_dev_mutex[dev]);
    if (!st.stream && cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking) != cudaSuccess) {
      fprintf(stderr, "yael_b200: cannot create a CUDA stream: %s\n",
              cudaGetErrorString(cudaGetLastError()));
      abort();
    }
  }
  return st.stream;
}

void *scratch_reserve(size_t bytes, cudaStream_t on) {
  DevState &st = g_dev[cur_dev()];
  if (!st.scratch_ev) cudaEventCreateWithFlags(&st.scratch_ev, cudaEventDisableTiming);
  // an earlier op on ANOTHER stream may still be using the block: order after it
  if (st.scratch_used && st.scratch_stream != on) cudaStreamWaitEvent(on, st.scratch_ev, 0);
  if (bytes > st.scratch_bytes) {
    if (st.scratch) {
      cudaDeviceSynchronize();  // earlier work may still read the old block
      cudaFree(st.scratch);
      st.scratch = nullptr;
      st.scratch_bytes = 0;
    }
    size_t want = bytes + (bytes >> 3) + (1u << 20);
    cudaError_t e = cudaMalloc(&st.scratch, want);
    if (e != cudaSuccess) {
      // same convention as the reference's allocators (yael/vector.c:37-40): message + abort
      fprintf(stderr, "yael_b200: device workspace of %zu bytes: %s\n", want,
              cudaGetErrorString(e));
      abort();
    }
    st.scratch_bytes = want;
  }
  return st.scratch;
}

void scratch_done(cudaStream_t on) {
  DevState &st = g_dev[cur_dev()];
  if (!st.scratch_ev) cudaEventCreateWithFlags(&st.scratch_ev, cudaEventDisableTiming);
  cudaEventRecord(st.scratch_ev, on);
  st.scratch_stream = on;
  st.scratch_used = true;
}

int sm_count() {
  DevState &st = g_dev[cur_dev()];
  if (!st.sms) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, cur_dev());
    st.sms = v > 0 ? v : 148;
  }
  return st.sms;
}

void count_launch(long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- optional phase timing (bench.py roofline): CUDA event pairs on the launching stream
struct ProfSpan {
  int phase;
  cudaEvent_t a, b;
};
static bool g_prof_on = false;
static std::vector<ProfSpan> g_spans;
static std::vector<cudaEvent_t> g_ev_pool;
static std::mutex g_prof_mutex;

static cudaEvent_t prof_event() {
  if (!g_ev_pool.empty()) {
    cudaEvent_t e = g_ev_pool.back();
    g_ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

int prof_begin(int phase, cudaStream_t st) {
  if (!g_prof_on) return -1;
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  ProfSpan s = {phase, prof_event(), prof_event()};
  cudaEventRecord(s.a, st);
  g_spans.push_back(s);
  return (int)g_spans.size() - 1;
}

void prof_end(int handle, cudaStream_t st) {
  if (handle < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  if (handle < (int)g_spans.size()) cudaEventRecord(g_spans[handle].b, st);
}

}  // namespace yb

extern "C" {

const char *yb_version(void) { return "yael_b200 0.1 (sm_100a)"; }
const char *yb_last_error(void) { return g_err; }

int yb_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    yb::fail(1, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return -1;
  }
  return n;
}

int yb_set_device(int dev) {
  YB_CUDA(cudaSetDevice(dev));
  g_pinned.store(true);  // the caller manages devices: the drop-in layer stays on this one
  return 0;
}

int yb_sync(yb_stream_t s) {
  YB_CUDA(cudaStreamSynchronize(yb::stream_of(s)));
  return 0;
}

long yb_launch_count(int reset) {
  long v = g_launches.load();
  if (reset) g_launches.store(0);
  return v;
}

// caching pool: exact-ish reuse of freed blocks (a k-means run or a stream of knn calls asks
// for the same sizes again and again; cudaMalloc/cudaFree would serialise the device)
void *yb_malloc(size_t bytes) {
  if (bytes == 0) bytes = 1;
  bytes = (bytes + 511) & ~(size_t)511;
  int dev = cur_dev();
  {
    std::unique_lock<std::mutex> lk(g_pool_mutex);
    PoolBlock *best = nullptr;
    for (auto &b : g_pool)
      if (b.free_ && b.dev == dev && b.bytes >= bytes && b.bytes <= bytes + (bytes >> 2) + 4096 &&
          (!best || b.bytes < best->bytes))
        best = &b;
    if (best) {
      best->free_ = false;
      void *p = best->p;
      cudaEvent_t ev[3];
      const int nev = best->nev;
      for (int e = 0; e < nev; e++) ev[e] = best->ev[e];
      lk.unlock();
      // earlier users of the block (any stream of this device) are done before it is handed out
      for (int e = 0; e < nev; e++)
        if (cudaEventQuery(ev[e]) != cudaSuccess) {
          cudaGetLastError();
          cudaEventSynchronize(ev[e]);
        }
      return p;
    }
  }
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    pool_trim(dev);  // give cached blocks back and retry once
    e = cudaMalloc(&p, bytes);
  }
  if (e != cudaSuccess) {
    fprintf(stderr, "yael_b200: cudaMalloc(%zu): %s\n", bytes, cudaGetErrorString(e));
    abort();
  }
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  PoolBlock nb = {p, bytes, dev, false, {nullptr, nullptr, nullptr}, 0};
  g_pool.push_back(nb);
  return p;
}

void yb_free(void *p) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    for (auto &b : g_pool)
      if (b.p == p) {
        // stream positions at the time of the free: own stream, copy stream, last caller stream
        int cur = 0;
        const bool same_dev = cudaGetDevice(&cur) == cudaSuccess && cur == b.dev;
        if (!same_dev) cudaSetDevice(b.dev);
        const DevState &ds = g_dev[b.dev];
        cudaStream_t ss[3] = {ds.stream, ds.copy_stream, g_last_caller_stream[b.dev]};
        int n = 0;
        for (int i = 0; i < 3; i++) {
          if (!ss[i] || (i == 2 && (ss[2] == ss[0] || ss[2] == ss[1]))) continue;
          if (!b.ev[n] && cudaEventCreateWithFlags(&b.ev[n], cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            b.ev[n] = nullptr;
            continue;
          }
          if (cudaEventRecord(b.ev[n], ss[i]) != cudaSuccess) {
            cudaGetLastError();  // a caller stream that no longer exists: nothing can be pending on it
            continue;
          }
          n++;
        }
        b.nev = n;
        if (!same_dev) cudaSetDevice(cur);
        b.free_ = true;
        return;
      }
  }
  cudaFree(p);  // not ours
}

int yb_is_device_ptr(const void *p) {
  if (!p) return 0;
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

int yb_h2d(void *dst, const void *src, size_t bytes, yb_stream_t s) {
  YB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, yb::stream_of(s)));
  return 0;
}

int yb_d2h(void *dst, const void *src, size_t bytes, yb_stream_t s) {
  YB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, yb::stream_of(s)));
  return 0;
}

void yb_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(yb::g_prof_mutex);
  yb::g_prof_on = on != 0;
}

// total milliseconds (and number of spans via *count) recorded for a phase; reset drops them
double yb_prof_ms(int phase, long *count, int reset) {
  std::lock_guard<std::mutex> lk(yb::g_prof_mutex);
  double ms = 0.0;
  long n = 0;
  for (auto &s : yb::g_spans) {
    if (s.phase != phase) continue;
    cudaEventSynchronize(s.b);
    float t = 0.f;
    if (cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) {
      ms += t;
      n++;
    }
  }
  if (count) *count = n;
  if (reset) {
    for (auto &s : yb::g_spans) {
      yb::g_ev_pool.push_back(s.a);
      yb::g_ev_pool.push_back(s.b);
    }
    yb::g_spans.clear();
  }
  return ms;
}

void yb_release_scratch(void) {
  yb::Guard g;
  DevState &st = g_dev[cur_dev()];
  if (st.scratch) {
    cudaDeviceSynchronize();
    cudaFree(st.scratch);
    st.scratch = nullptr;
    st.scratch_bytes = 0;
  }
  pool_trim(cur_dev());
}

}  // extern "C"
