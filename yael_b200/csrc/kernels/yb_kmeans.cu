// yb_kmeans.cu -- centroid update of Lloyd's k-means (kmeans_core, yael/kmeans.c:247-288,310)
// as a sort + segmented reduction.
//
// The reference adds the points into their centroid serially in point order
// (yael/kmeans.c:278-283).  Here:
//   1. k_hist            nassign[] histogram (yael/kmeans.c:249-251)
//   2. stable LSD radix sort of the point ids by centroid id (8 bits per pass: hist / scan /
//      stable scatter) -> `order` lists every centroid's points in increasing point id
//   3. k_segsum          one warp per piece of <= P consecutive points of one centroid:
//      coalesced 128-bit row loads, FP32 adds in list order.  With one piece per centroid
//      (P >= the largest cluster) the sums equal the reference's bit for bit.
//   4. k_combine         adds the pieces of a centroid in order
//   5. k_sum_dis         qerr = sum(dis) as a fixed-shape double tree (yael/kmeans.c:310)
//   6. k_scale           c = (float)((double)c * (1.0/n)) (+ optional L2 normalisation)
// Everything is deterministic: no floating-point atomics anywhere.
#include "yb_common.cuh"
#include "yb_internal.cuh"

namespace yb {

// Device-side path selection without a host round trip: `guard` points at the largest cluster size
// of this iteration (NULL = no selection).  The short-segment kernels run when it is <= SG_LIMIT,
// the general (radix sort) kernels launched behind them when it is larger; the others return at once.
// 2048: a longer segment makes the short-segment kernel's id-range passes quadratic and ONE warp's
// serial walk the critical path -- measured at BASELINE configs[3]: the first iteration after a
// random-point initialisation has a 12 860-point cluster (p99.9: 3214, median 65) and the row stream
// took 9.9 ms against 0.9 ms from the second iteration on (largest cluster 316); the general path does
// that iteration in 2.0 ms.  (Was 65536: only truly degenerate clusterings switched.)
constexpr int kSegLimit = 2048;
__device__ __forceinline__ bool general_path_skips(const int *guard) {
  return guard != nullptr && *guard <= kSegLimit;
}

// out-of-range ids (the reference asserts on them, yael/kmeans.c:281) are not counted: the host
// loop notices that the histogram does not add up to n
__global__ void k_hist(const int *__restrict__ assign, long n, int k, int *__restrict__ counts) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long)gridDim.x * blockDim.x) {
    int a = assign[i];
    if (a >= 0 && a < k) atomicAdd(&counts[a], 1);
  }
}

// ------------------------------------------------------------------ stable radix sort
constexpr int RS_T = 256, RS_ITEMS = 16, RS_BLOCK = RS_T * RS_ITEMS;

// pass 0 reads keys from assign[] (payload = position); later passes read (key,payload) pairs
__device__ __forceinline__ void rs_load(const int *assign, const int2 *in, long i, int &key,
                                        int &pay) {
  if (in) {
    int2 v = in[i];
    key = v.x;
    pay = v.y;
  } else {
    key = assign[i];
    pay = (int)i;
  }
}

__global__ void __launch_bounds__(RS_T)
k_rs_hist(const int *__restrict__ assign, const int2 *__restrict__ in, long n, int shift,
          int nblocks, unsigned *__restrict__ ghist, const int *__restrict__ guard) {
  __shared__ unsigned h[256];
  if (general_path_skips(guard)) return;
  h[threadIdx.x] = 0;
  __syncthreads();
  long b0 = (long)blockIdx.x * RS_BLOCK;
  for (int r = 0; r < RS_ITEMS; r++) {
    long i = b0 + r * RS_T + threadIdx.x;
    if (i < n) {
      int key, pay;
      rs_load(assign, in, i, key, pay);
      atomicAdd(&h[(key >> shift) & 255], 1u);
    }
  }
  __syncthreads();
  ghist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];  // digit-major
}

// exclusive scan of `len` unsigned values by one CTA: tiles of 4096 values (coalesced 16-byte
// accesses), warp-shuffle scans inside the tile, a running carry between tiles
__global__ void __launch_bounds__(1024)
k_scan_u32(unsigned *__restrict__ v, long len, const int *__restrict__ guard) {
  __shared__ unsigned wsum[32];
  __shared__ unsigned tile_total;
  if (general_path_skips(guard)) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned carry = 0;
  const bool vec = (((uintptr_t)v) & 15) == 0;
  for (long base = 0; base < len; base += 4096) {
    const long i = base + (long)tid * 4;
    unsigned x[4] = {0u, 0u, 0u, 0u};
    if (vec && i + 4 <= len) {
      const uint4 q = *reinterpret_cast<const uint4 *>(v + i);
      x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (i + j < len) x[j] = v[i + j];
    }
    const unsigned s = x[0] + x[1] + x[2] + x[3];
    unsigned inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const unsigned w = wsum[lane];
      unsigned wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      wsum[lane] = wi - w;
      if (lane == 31) tile_total = wi;
    }
    __syncthreads();
    unsigned off = carry + wsum[warp] + inc - s;
    if (vec && i + 4 <= len) {
      uint4 q;
      q.x = off; q.y = off + x[0]; q.z = q.y + x[1]; q.w = q.z + x[2];
      *reinterpret_cast<uint4 *>(v + i) = q;
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (i + j < len) {
          v[i + j] = off;
          off += x[j];
        }
    }
    carry += tile_total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(RS_T)
k_rs_scatter(const int *__restrict__ assign, const int2 *__restrict__ in, long n, int shift,
             int nblocks, const unsigned *__restrict__ ghist, int2 *__restrict__ out,
             const int *__restrict__ guard) {
  __shared__ unsigned base[256];          // next output slot of each digit for this block
  __shared__ unsigned short wc[8][256];   // per-warp digit counts of the current round
  if (general_path_skips(guard)) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  base[tid] = ghist[(size_t)tid * nblocks + blockIdx.x];
  long b0 = (long)blockIdx.x * RS_BLOCK;
  for (int r = 0; r < RS_ITEMS; r++) {
    for (int x = tid; x < 8 * 256; x += RS_T) (&wc[0][0])[x] = 0;
    __syncthreads();
    long i = b0 + r * RS_T + tid;
    int key = 0, pay = 0, dig = -1;
    if (i < n) {
      rs_load(assign, in, i, key, pay);
      dig = (key >> shift) & 255;
    }
    unsigned peers = __match_any_sync(0xffffffffu, dig);
    int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (dig >= 0 && rank_in_warp == 0) wc[warp][dig] = (unsigned short)__popc(peers);
    __syncthreads();
    if (dig >= 0) {
      unsigned before = 0;
      for (int w = 0; w < warp; w++) before += wc[w][dig];
      out[base[dig] + before + rank_in_warp] = make_int2(key, pay);
    }
    __syncthreads();
    {  // advance the digit bases by this round's totals
      unsigned tot = 0;
#pragma unroll
      for (int w = 0; w < 8; w++) tot += wc[w][tid];
      base[tid] += tot;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ segmented sums
__global__ void k_piece_counts(const int *__restrict__ counts, int k, int P,
                               unsigned *__restrict__ seg_start, unsigned *__restrict__ piece_start,
                               const int *__restrict__ guard) {
  if (general_path_skips(guard)) return;
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < k) {
    seg_start[c] = (unsigned)counts[c];
    piece_start[c] = (unsigned)((counts[c] + P - 1) / P);
  }
  if (c == k) {
    seg_start[k] = 0;
    piece_start[k] = 0;
  }
}

// piece -> centroid map (binary search on piece_start), one thread per piece
__global__ void k_piece_map(const unsigned *__restrict__ piece_start, int k,
                            int *__restrict__ piece_cent, const int *__restrict__ guard) {
  if (general_path_skips(guard)) return;
  unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= piece_start[k]) return;
  int lo = 0, hi = k - 1;  // last c with piece_start[c] <= p
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (piece_start[mid] <= p) lo = mid; else hi = mid - 1;
  }
  piece_cent[p] = lo;
}

// one warp per piece; NV = 4: d % 4 == 0 and 16-byte aligned rows (128-bit loads), NV = 1 else
template <int NV>
__device__ __forceinline__ void load_nv(const float *p, float (&o)[NV]) {
  if constexpr (NV == 4) {
    float4 x = ld_stream_f4(p);
    o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w;
  } else {
    o[0] = __ldg(p);
  }
}

template <int NV>
__global__ void __launch_bounds__(128)
k_segsum(int d, const float *__restrict__ v, const int2 *__restrict__ order,
         const unsigned *__restrict__ seg_start, const unsigned *__restrict__ piece_start,
         const int *__restrict__ piece_cent, int k, int P, float *__restrict__ psums,
         const int *__restrict__ guard) {
  if (general_path_skips(guard)) return;
  const unsigned p = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= piece_start[k]) return;  // piece_start[k] = number of pieces
  const int c = piece_cent[p];
  const unsigned first = seg_start[c] + (p - piece_start[c]) * (unsigned)P;
  const unsigned last = min(seg_start[c + 1], first + (unsigned)P);
  float *out = psums + (size_t)p * d;
  constexpr int CH = 32 * NV;  // coordinates covered by the warp per chunk
  for (int t0 = 0; t0 < d; t0 += 4 * CH) {  // up to 4 chunks of accumulators in registers
    float acc[4][NV];
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int x = 0; x < NV; x++) acc[u][x] = 0.f;
    unsigned e = first;
    for (; e + 4 <= last; e += 4) {  // four rows in flight, added in list order
      float val[4][4][NV];
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const float *row = v + (size_t)order[e + r].y * d;
#pragma unroll
        for (int u = 0; u < 4; u++) {
          int t = t0 + u * CH + lane * NV;
          if (t < d) {
            load_nv<NV>(row + t, val[r][u]);
          } else {
#pragma unroll
            for (int x = 0; x < NV; x++) val[r][u][x] = 0.f;
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
          for (int x = 0; x < NV; x++) acc[u][x] = __fadd_rn(acc[u][x], val[r][u][x]);
    }
    for (; e < last; e++) {
      const float *row = v + (size_t)order[e].y * d;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        int t = t0 + u * CH + lane * NV;
        if (t < d) {
          float x[NV];
          load_nv<NV>(row + t, x);
#pragma unroll
          for (int y = 0; y < NV; y++) acc[u][y] = __fadd_rn(acc[u][y], x[y]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      int t = t0 + u * CH + lane * NV;
      if (t < d) {
        if constexpr (NV == 4)
          *reinterpret_cast<float4 *>(out + t) = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
        else
          out[t] = acc[u][0];
      }
    }
  }
}

// ------------------------------------------------------------------ short segments (large k)
// When the clusters are small (n / k <= SG_AVG: BASELINE config 4 has 153 points per centroid) the
// multi-pass radix sort above costs more than the row stream it prepares.  Instead:
//   1. k_scatter_ids   order[cursor[c]++] = i with an integer atomic on the centroid's cursor:
//      one pass, the points of a centroid land in its segment in ARBITRARY order;
//   2. k_segsum_sorted one warp per centroid sorts its segment's point ids in shared memory
//      (bitonic, <= SG_CAP ids at a time) and adds the rows in increasing point id -- the
//      reference's order (yael/kmeans.c:278-283), so the sums equal the reference's bit for bit
//      and do not depend on how the atomics were scheduled.  Segments longer than SG_CAP are
//      walked in passes over id ranges that hold <= SG_CAP ids each.
// A segment longer than SG_LIMIT (a degenerate clustering) would make that walk quadratic: the
// kernels below then do nothing (device-side flag maxc[0], no host round trip) and the general
// path, launched behind them with the opposite guard, does the work.
constexpr int SG_CAP = 1024;
constexpr int SG_WARPS = 4;
constexpr int SG_LIMIT = kSegLimit;
constexpr int SG_AVG = 1024;

__global__ void k_max_count(const int *__restrict__ counts, int k, int *__restrict__ maxc) {
  int m = 0;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < k; c += gridDim.x * blockDim.x) m = max(m, counts[c]);
#pragma unroll
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(maxc, m);
}

// seg_start[c] = cursor[c] = counts[c] (scanned afterwards); seg_start[k] = 0
__global__ void k_seg_counts(const int *__restrict__ counts, int k, unsigned *__restrict__ seg_start) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c <= k) seg_start[c] = c < k ? (unsigned)counts[c] : 0u;
}

__global__ void __launch_bounds__(256)
k_scatter_ids(const int *__restrict__ assign, long n, int k, const unsigned *__restrict__ seg_start,
              unsigned *__restrict__ cursor, int *__restrict__ order, const int *__restrict__ maxc) {
  if (*maxc > SG_LIMIT) return;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int a = assign[i];
    if (a >= 0 && a < k) order[seg_start[a] + atomicAdd(&cursor[a], 1u)] = (int)i;
  }
}

// sort n_pad (power of two <= SG_CAP) ints in shared memory, one warp
__device__ __forceinline__ void warp_bitonic_i32(int *a, int n_pad, int lane) {
  for (int size = 2; size <= n_pad; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncwarp();
      for (int t = lane; t < (n_pad >> 1); t += 32) {
        const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
        const bool up = (lo & size) == 0;
        const int x = a[lo], y = a[hi];
        if ((x > y) == up) {
          a[lo] = y;
          a[hi] = x;
        }
      }
    }
  __syncwarp();
}

// acc[u][x] += rows ids[0..cnt) in list order; R rows in flight
template <int NV, int NCH, int R>
__device__ __forceinline__ void seg_accumulate(const float *__restrict__ v, int d, int t0, int lane,
                                               const int *ids, int cnt, float (&acc)[NCH][NV]) {
  constexpr int CH = 32 * NV;
  int e = 0;
  for (; e + R <= cnt; e += R) {
    float val[R][NCH][NV];
#pragma unroll
    for (int r = 0; r < R; r++) {
      const float *row = v + (size_t)ids[e + r] * d;
#pragma unroll
      for (int u = 0; u < NCH; u++) {
        const int t = t0 + u * CH + lane * NV;
        if (t < d) {
          load_nv<NV>(row + t, val[r][u]);
        } else {
#pragma unroll
          for (int x = 0; x < NV; x++) val[r][u][x] = 0.f;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
      for (int u = 0; u < NCH; u++)
#pragma unroll
        for (int x = 0; x < NV; x++) acc[u][x] = __fadd_rn(acc[u][x], val[r][u][x]);
  }
  for (; e < cnt; e++) {
    const float *row = v + (size_t)ids[e] * d;
#pragma unroll
    for (int u = 0; u < NCH; u++) {
      const int t = t0 + u * CH + lane * NV;
      if (t < d) {
        float x[NV];
        load_nv<NV>(row + t, x);
#pragma unroll
        for (int y = 0; y < NV; y++) acc[u][y] = __fadd_rn(acc[u][y], x[y]);
      }
    }
  }
}

template <int NV, int NCH>
__global__ void __launch_bounds__(32 * SG_WARPS)
k_segsum_sorted(int d, long n, const float *__restrict__ v, const int *__restrict__ order,
                const unsigned *__restrict__ seg_start, int k, float *__restrict__ sums,
                const int *__restrict__ maxc) {
  __shared__ int ids_s[SG_WARPS][SG_CAP];
  if (*maxc > SG_LIMIT) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * SG_WARPS + warp;
  if (c >= k) return;
  int *ids = ids_s[warp];
  const unsigned first = seg_start[c], last = seg_start[c + 1];
  const int m = (int)(last - first);
  constexpr int CH = 32 * NV;
  constexpr int R = NCH == 1 ? 8 : 4;
  const unsigned lt = (1u << lane) - 1u;
  float *out = sums + (size_t)c * d;
  for (int t0 = 0; t0 < d; t0 += NCH * CH) {
    float acc[NCH][NV];
#pragma unroll
    for (int u = 0; u < NCH; u++)
#pragma unroll
      for (int x = 0; x < NV; x++) acc[u][x] = 0.f;
    if (m <= SG_CAP) {
      if (t0 == 0) {  // the sorted ids stay in shared memory for the later column blocks
        const int n_pad = pow2_ceil(m < 2 ? 2 : m);
        for (int e = lane; e < n_pad; e += 32) ids[e] = e < m ? order[first + e] : 0x7fffffff;
        warp_bitonic_i32(ids, n_pad, lane);
      }
      seg_accumulate<NV, NCH, R>(v, d, t0, lane, ids, m, acc);
    } else {
      // passes over id ranges [lo, hi) that hold at most SG_CAP of the segment's ids
      const long passes = (m + SG_CAP / 2 - 1) / (SG_CAP / 2);
      const long w0 = (n + passes - 1) / passes;
      long lo = 0;
      while (lo < n) {
        long w = w0, hi;
        int cnt;
        for (;;) {
          hi = lo + w < n ? lo + w : n;
          cnt = 0;
          for (unsigned e = first + lane; e < last; e += 32) {
            const int id = order[e];
            cnt += (id >= lo && id < hi);
          }
          cnt = warp_sum(cnt);
          if (cnt <= SG_CAP) break;
          w = (w + 1) / 2;  // ids are distinct: a range of w <= SG_CAP ids always fits
        }
        if (cnt > 0) {
          int pos = 0;
          for (unsigned e0 = first; e0 < last; e0 += 32) {
            const int id = e0 + lane < last ? order[e0 + lane] : -1;
            const bool in = id >= lo && id < hi;
            const unsigned b = __ballot_sync(0xffffffffu, in);
            if (in) ids[pos + __popc(b & lt)] = id;
            pos += __popc(b);
          }
          const int n_pad = pow2_ceil(cnt < 2 ? 2 : cnt);
          __syncwarp();
          for (int e = cnt + lane; e < n_pad; e += 32) ids[e] = 0x7fffffff;
          warp_bitonic_i32(ids, n_pad, lane);
          seg_accumulate<NV, NCH, R>(v, d, t0, lane, ids, cnt, acc);
          __syncwarp();
        }
        lo = hi;
      }
    }
#pragma unroll
    for (int u = 0; u < NCH; u++) {
      const int t = t0 + u * CH + lane * NV;
      if (t < d) {
        if constexpr (NV == 4)
          *reinterpret_cast<float4 *>(out + t) = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
        else
          out[t] = acc[u][0];
      }
    }
  }
}

__global__ void k_combine(int d, int k, const unsigned *__restrict__ piece_start,
                          const float *__restrict__ psums, float *__restrict__ sums,
                          const int *__restrict__ guard) {
  if (general_path_skips(guard)) return;
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)k * d) return;
  int c = (int)(t / d), x = (int)(t - (long)c * d);
  unsigned p0 = piece_start[c], p1 = piece_start[c + 1];
  float s = 0.f;
  if (p0 < p1) {
    s = psums[(size_t)p0 * d + x];
    for (unsigned p = p0 + 1; p < p1; p++) s = __fadd_rn(s, psums[(size_t)p * d + x]);
  }
  sums[t] = s;
}

// fixed-shape double reduction: 1024 partials (contiguous slices), then one thread in order
__global__ void __launch_bounds__(256) k_sum_dis_partial(const float *__restrict__ x, long n,
                                                          double *__restrict__ part) {
  __shared__ double sh[256];
  long per = (n + gridDim.x - 1) / gridDim.x;
  long b = (long)blockIdx.x * per, e = min(n, b + per);
  double s = 0.0;
  for (long i = b + threadIdx.x; i < e; i += 256) s += (double)x[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
__global__ void k_sum_dis_final(const double *__restrict__ part, int np, double *out) {
  double s = 0.0;
  for (int i = 0; i < np; i++) s += part[i];
  *out = s;
}

__global__ void k_scale(int d, int k, const float *__restrict__ sums, const int *__restrict__ cnt,
                        float *__restrict__ cent, int normalize) {
  // thread per centroid row when normalising (sequential double norm, yael/vector.c:2180-2199),
  // thread per element otherwise
  if (!normalize) {
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)k * d) return;
    int c = (int)(t / d);
    double f = 1.0 / (double)cnt[c];  // inf for an empty cluster: 0 * inf = NaN, as the reference
    cent[t] = (float)((double)sums[t] * f);
  } else {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k) return;
    double f = 1.0 / (double)cnt[c];
    double nr = 0.0;
    for (int x = 0; x < d; x++) {
      float val = (float)((double)sums[(size_t)c * d + x] * f);
      cent[(size_t)c * d + x] = val;
      nr += (double)__fmul_rn(val, val);
    }
    double g = 1.0 / sqrt(nr);
    for (int x = 0; x < d; x++)
      cent[(size_t)c * d + x] = (float)((double)cent[(size_t)c * d + x] * g);
  }
}

}  // namespace yb

using namespace yb;

extern "C" int yb_kmeans_accumulate(int d, int n, int k, const float *v, const int *assign,
                                     const float *dis, float *sums, int *nassign, double *qerr,
                                     int exact_order, yb_stream_t s) {
  if (k <= 0 || d <= 0) return 0;
  Guard g;
  cudaStream_t st = stream_of(s);
  // piece length: one piece per centroid reproduces the reference's summation order; when
  // that would leave the machine idle (few, large clusters) shorter pieces trade the last
  // ulp for parallelism (still deterministic).
  int P;
  bool many = false;
  if (exact_order) {
    P = 1 << 30;
  } else {
    long target = (long)sm_count() * 32;  // warps wanted in flight
    long p = ((long)n + target - 1) / target;
    P = (int)(p < 64 ? 64 : p > 4096 ? 4096 : p);
    // many centroids: the short-segment path below takes the iterations whose clusters are all
    // small (strict order); the general path only runs when a cluster exceeds kSegLimit -- the first
    // iteration after a random-point initialisation: 12 860 points in one cluster at BASELINE
    // configs[3] -- and then walks such clusters in pieces instead of one warp per cluster (9 ms)
    if ((long)k >= target) P = kSegLimit;
    many = (long)k >= target;
  }
  // short segments (many centroids): one-pass scatter + per-segment sort; the general path is
  // still launched behind it, guarded on the device by the largest cluster size
  const bool fast = n > 0 && k >= 1024 && (long)n / k <= SG_AVG && (exact_order || many) &&
                    !getenv("YAEL_B200_KMEANS_GENERAL_UPDATE");
  const int nblocks = (int)(((long)n + RS_BLOCK - 1) / RS_BLOCK);
  int passes = 1;
  while (passes < 4 && ((long)k - 1) >> (8 * passes)) passes++;
  const unsigned max_pieces = (unsigned)(k + (P >= (1 << 30) ? 0 : ((long)n + P - 1) / P)) + 1;
  size_t need = 2 * Carver::need(sizeof(int2) * (size_t)(n > 0 ? n : 1)) +
                Carver::need(4ull * 256 * (nblocks > 0 ? nblocks : 1)) +
                3 * Carver::need(4ull * (k + 1)) + Carver::need(4ull * max_pieces) +
                Carver::need(4ull * (size_t)max_pieces * d) + Carver::need(8 * 1024 + 64) + Carver::need(64) +
                (fast ? Carver::need(sizeof(int) * (size_t)n) : 0);
  ScratchScope ws(need, st);
  ProfScope ps(8, st);
  Carver c(ws.p);
  int2 *bufA = c.take<int2>(n > 0 ? n : 1);
  int2 *bufB = c.take<int2>(n > 0 ? n : 1);
  unsigned *ghist = c.take<unsigned>(256ull * (nblocks > 0 ? nblocks : 1));
  unsigned *seg_start = c.take<unsigned>(k + 1);
  unsigned *piece_start = c.take<unsigned>(k + 1);
  unsigned *cursor = c.take<unsigned>(k + 1);
  int *piece_cent = c.take<int>(max_pieces);
  float *psums = c.take<float>((size_t)max_pieces * d);
  double *part = c.take<double>(1024 + 8);
  int *maxc = c.take<int>(16);
  int *forder = fast ? c.take<int>(n) : nullptr;  // the short-segment path's id list

  YB_CUDA(cudaMemsetAsync(nassign, 0, sizeof(int) * (size_t)k, st));
  const bool vec = (d % 4 == 0) && ((((uintptr_t)v) & 15) == 0);
  const int *guard = nullptr;
  if (n > 0) {
    k_hist<<<4 * sm_count(), 256, 0, st>>>(assign, n, k, nassign);
    YB_LAUNCH_CHECK();
  }
  if (fast) {
    int *order = forder;
    YB_CUDA(cudaMemsetAsync(maxc, 0, 64, st));
    YB_CUDA(cudaMemsetAsync(cursor, 0, 4ull * (k + 1), st));
    k_max_count<<<(k + 1023) / 1024 < 64 ? (k + 1023) / 1024 : 64, 1024, 0, st>>>(nassign, k, maxc);
    YB_LAUNCH_CHECK();
    k_seg_counts<<<(k + 1 + 255) / 256, 256, 0, st>>>(nassign, k, seg_start);
    YB_LAUNCH_CHECK();
    k_scan_u32<<<1, 1024, 0, st>>>(seg_start, k + 1, nullptr);
    YB_LAUNCH_CHECK();
    k_scatter_ids<<<8 * sm_count(), 256, 0, st>>>(assign, n, k, seg_start, cursor, order, maxc);
    YB_LAUNCH_CHECK();
    const unsigned grid = (unsigned)((k + SG_WARPS - 1) / SG_WARPS);
    ProfScope pseg(18, st);  // the row stream alone (the update's dominant kernel)
    if (vec && d <= 128)
      k_segsum_sorted<4, 1><<<grid, 32 * SG_WARPS, 0, st>>>(d, n, v, order, seg_start, k, sums, maxc);
    else if (vec)
      k_segsum_sorted<4, 4><<<grid, 32 * SG_WARPS, 0, st>>>(d, n, v, order, seg_start, k, sums, maxc);
    else
      k_segsum_sorted<1, 4><<<grid, 32 * SG_WARPS, 0, st>>>(d, n, v, order, seg_start, k, sums, maxc);
    YB_LAUNCH_CHECK();
    guard = maxc;  // the general path below runs only if a cluster exceeded SG_LIMIT
  }
  if (n > 0) {
    const int2 *in = nullptr;
    int2 *out = bufA;
    for (int p = 0; p < passes; p++) {
      k_rs_hist<<<nblocks, RS_T, 0, st>>>(assign, in, n, 8 * p, nblocks, ghist, guard);
      YB_LAUNCH_CHECK();
      k_scan_u32<<<1, 1024, 0, st>>>(ghist, 256L * nblocks, guard);
      YB_LAUNCH_CHECK();
      k_rs_scatter<<<nblocks, RS_T, 0, st>>>(assign, in, n, 8 * p, nblocks, ghist, out, guard);
      YB_LAUNCH_CHECK();
      in = out;
      out = (out == bufA) ? bufB : bufA;
    }
    const int2 *order = in;
    // (the general path gets its own segment offsets: the short-segment kernels may be reading
    // seg_start while these are enqueued)
    unsigned *gseg = fast ? cursor : seg_start;
    k_piece_counts<<<(k + 1 + 255) / 256, 256, 0, st>>>(nassign, k, P, gseg, piece_start, guard);
    YB_LAUNCH_CHECK();
    k_scan_u32<<<1, 1024, 0, st>>>(gseg, k + 1, guard);
    YB_LAUNCH_CHECK();
    k_scan_u32<<<1, 1024, 0, st>>>(piece_start, k + 1, guard);
    YB_LAUNCH_CHECK();
    // the number of pieces is data dependent (piece_start[k]); launch for the upper bound,
    // surplus threads / warps exit
    k_piece_map<<<(max_pieces + 255) / 256, 256, 0, st>>>(piece_start, k, piece_cent, guard);
    YB_LAUNCH_CHECK();
    if (vec)
      k_segsum<4><<<(max_pieces + 3) / 4, 128, 0, st>>>(d, v, order, gseg, piece_start, piece_cent, k, P,
                                                       psums, guard);
    else
      k_segsum<1><<<(max_pieces + 3) / 4, 128, 0, st>>>(d, v, order, gseg, piece_start, piece_cent, k, P,
                                                       psums, guard);
    YB_LAUNCH_CHECK();
  } else {
    YB_CUDA(cudaMemsetAsync(piece_start, 0, 4ull * (k + 1), st));
  }
  long tot = (long)k * d;
  k_combine<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d, k, piece_start, psums, sums, guard);
  YB_LAUNCH_CHECK();
  if (qerr) {
    k_sum_dis_partial<<<1024, 256, 0, st>>>(dis, n, part);
    YB_LAUNCH_CHECK();
    k_sum_dis_final<<<1, 1, 0, st>>>(part, 1024, qerr);
    YB_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int yb_kmeans_scale(int d, int k, const float *sums, const int *nassign,
                                float *centroids, int normalize, yb_stream_t s) {
  if (k <= 0 || d <= 0) return 0;
  Guard g;
  cudaStream_t st = stream_of(s);
  if (!normalize) {
    long tot = (long)k * d;
    k_scale<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d, k, sums, nassign, centroids, 0);
  } else {
    k_scale<<<(k + 127) / 128, 128, 0, st>>>(d, k, sums, nassign, centroids, 1);
  }
  YB_LAUNCH_CHECK();
  return 0;
}
