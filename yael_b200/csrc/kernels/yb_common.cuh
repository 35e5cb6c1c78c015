// yb_common.cuh -- shared device helpers and the runtime interface used by every kernel file.
// sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../../include/yael_b200.h"

// ---------------------------------------------------------------- runtime (yb_runtime.cu)
namespace yb {

// records a failure message for yb_last_error(); returns code so callers can `return fail(..)`
int fail(int code, const char *fmt, ...);
// stream to launch on: the caller's, or the library's own stream for the current device
cudaStream_t stream_of(yb_stream_t s);
cudaStream_t copy_stream();  // per-device stream for host->device feeds
// grow-only cached workspace of the current device; valid until the next reserve() call
// on the same device.  Calls on one device serialise their HOST-side submission behind that
// device's mutex (Guard); scratch_done() records the stream position after which the block may be reused
// from another stream.
void *scratch_reserve(size_t bytes, cudaStream_t on);
void scratch_done(cudaStream_t on);
int sm_count();
void count_launch(long n = 1);
// phase timing for bench.py (no-ops unless yb_prof_enable(1)); phases: see yael_b200.h
int prof_begin(int phase, cudaStream_t st);
void prof_end(int handle, cudaStream_t st);
struct ProfScope {
  int h;
  cudaStream_t st;
  ProfScope(int phase, cudaStream_t s) : h(prof_begin(phase, s)), st(s) {}
  ~ProfScope() { prof_end(h, st); }
};

struct Guard {  // RAII lock of the CURRENT DEVICE's mutex (the reference promises re-entrancy:
  int dev;      // doc/index.rst:55-58; callers may come from several host threads): calls on one
  Guard();      // device serialise their host-side submission, calls on different devices -- one
  ~Guard();     // host thread per GPU -- run concurrently
};
int dev_index();  // the current device
// run `f` (cudaFuncSetAttribute and the like) once per DEVICE: function attributes belong to the
// device that is current when they are set
template <typename F>
inline void once_per_device(bool (&done)[64], F f) {
  const int dv = dev_index();
  if (!done[dv]) {
    f();
    done[dv] = true;
  }
}

struct ScratchScope {  // reserve in the constructor, publish completion in the destructor
  cudaStream_t st;
  void *p;
  ScratchScope(size_t bytes, cudaStream_t on) : st(on), p(scratch_reserve(bytes, on)) {}
  ~ScratchScope() { scratch_done(st); }
};

// carve aligned pieces out of one scratch reservation
struct Carver {
  char *base;
  size_t off;
  explicit Carver(void *p) : base((char *)p), off(0) {}
  template <typename T>
  T *take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T *r = (T *)(base + off);
    off += n * sizeof(T);
    return r;
  }
  static size_t need(size_t bytes) { return ((bytes + 255) & ~(size_t)255) + 256; }
};

}  // namespace yb

#define YB_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return yb::fail(1, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

#define YB_LAUNCH_CHECK()                                                                  \
  do {                                                                                     \
    yb::count_launch();                                                                    \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess)                                                                 \
      return yb::fail(2, "%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
  } while (0)

// ---------------------------------------------------------------- device helpers
namespace yb {

// Monotone map float -> uint32 such that unsigned order == the order the reference's
// comparisons induce: -0.0 and +0.0 compare equal (canonicalised), NaN sorts after +inf
// (it is never selected: yael/binheap.c:144,149).
__device__ __forceinline__ uint32_t float_key(float f) {
  uint32_t u = __float_as_uint(f);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;  // NaN
  if (u == 0x80000000u) u = 0;                              // -0.0 -> +0.0
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ bool is_nan_key(uint32_t k) { return k == 0xffffffffu; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 128-bit streaming loads (read-once data: bypass L1 allocation)
__device__ __forceinline__ uint4 ld_stream_u4(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_stream_f4(const void *p) {
  uint4 r = ld_stream_u4(p);
  return make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z),
                     __uint_as_float(r.w));
}

// In-place bitonic sort of n_pad (power of two) 64-bit keys by the calling thread group
// (nthr threads, thread id tid), ascending.  `sync` separates the steps: __syncthreads for a
// CTA, __syncwarp for a warp.  Works on shared or global memory.
template <typename Sync>
__device__ __forceinline__ void bitonic_sort_u64(unsigned long long *a, int n_pad, int tid,
                                                 int nthr, Sync sync) {
  for (int size = 2; size <= n_pad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      sync();
      for (int t = tid; t < (n_pad >> 1); t += nthr) {
        int lo = 2 * t - (t & (stride - 1));  // index with bit `stride` cleared
        int hi = lo + stride;
        bool up = ((lo & size) == 0);
        unsigned long long x = a[lo], y = a[hi];
        if ((x > y) == up) {
          a[lo] = y;
          a[hi] = x;
        }
      }
    }
  }
  sync();
}

__host__ __device__ __forceinline__ int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace yb
