// yb_distance.cu -- exact FP32 distance kernels (the value-defining path).
//
//   k_row_norms_seq   squared norms in the reference's accumulation order (yael/nn.c:108-120)
//   k_l2_simt         dist2 = fl32(fl64 |b|^2 + fl32 |a|^2) - 2<a,b>, dot = sequential FP32 FMA
//                     chain over the coordinates (what the reference's sgemm micro-kernel does
//                     for each output element; verified bit-for-bit against the compiled
//                     reference, see DESIGN.md "numerics")
//   k_distances_1     one-vs-many with both norms in double (yael/nn.c:132-154)
//   k_alt_pairs       compute_cross_distances_alt_nonpacked (yael/nn.c:178-216, 280-350)
//
// These kernels define the VALUES the library returns.  The tcgen05 TF32 kernel
// (yb_knn_tf32.cu) only ever produces a shortlist that is re-ranked with the same
// arithmetic as here (yb_knn.cu: k_rerank).
#include <stdlib.h>

#include "yb_common.cuh"
#include "yb_internal.cuh"

namespace yb {

// ------------------------------------------------------------------------------------
// A warp walks 32 rows at once: coordinates are staged 32 at a time through a padded
// shared tile with coalesced 128-byte reads, then every lane consumes ITS row left to
// right, so per-row arithmetic is strictly sequential (the reference's order) while global
// traffic stays coalesced.  rowptr is per lane (nullptr = no row).
// ------------------------------------------------------------------------------------
template <typename F>
__device__ __forceinline__ void warp_rows_seq(const float *rowptr, int d, float (*tile)[33],
                                              F f) {
  const int lane = threadIdx.x & 31;
  for (int t0 = 0; t0 < d; t0 += 32) {
    const int w = min(32, d - t0);
#pragma unroll 4
    for (int r = 0; r < 32; r++) {
      const float *p = (const float *)__shfl_sync(0xffffffffu, (unsigned long long)rowptr, r);
      tile[r][lane] = (p != nullptr && lane < w) ? __ldg(p + t0 + lane) : 0.0f;
    }
    __syncwarp();
    if (rowptr != nullptr) {
      for (int t = 0; t < w; t++) f(t0 + t, tile[lane][t]);
    }
    __syncwarp();
  }
}

// out_f[i] = float-accumulated sum of squares (a-side, yael/nn.c:108-114)
// out_d[i] = double-accumulated sum of float products (b-side, yael/nn.c:116-120)
__global__ void __launch_bounds__(128) k_row_norms_seq(const float *__restrict__ x, long n, int d,
                                                        long ld, float *__restrict__ out_f,
                                                        double *__restrict__ out_d) {
  __shared__ float tile[4][32][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long row = ((long)blockIdx.x * 4 + warp) * 32 + lane;
  const float *p = row < n ? x + row * ld : nullptr;
  float sf = 0.0f;
  double sd = 0.0;
  warp_rows_seq(p, d, tile[warp], [&](int, float v) {
    float sq = __fmul_rn(v, v);
    sf = __fadd_rn(sf, sq);
    sd += (double)sq;
  });
  if (row < n) {
    if (out_f) out_f[row] = sf;
    if (out_d) out_d[row] = sd;
  }
}

// ------------------------------------------------------------------------------------
// SIMT tile kernel: 64 a-rows x 64 b-rows per CTA, 16 coordinates per stage, 256 threads,
// 4x4 outputs per thread.  The k loop runs over ascending coordinates with one fmaf per
// step, i.e. the per-element sequential FMA chain.
// ------------------------------------------------------------------------------------
constexpr int SB = 64, SK = 16;

__device__ __forceinline__ void load_tile_k4(const float *__restrict__ m, long nrows, int d,
                                             long ld, long row, int t, bool vec_ok,
                                             float out[4]) {
  if (row < nrows && vec_ok && t + 4 <= d) {
    float4 v = __ldg(reinterpret_cast<const float4 *>(m + row * ld + t));
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
  } else {
#pragma unroll
    for (int u = 0; u < 4; u++)
      out[u] = (row < nrows && t + u < d) ? __ldg(m + row * ld + t + u) : 0.0f;
  }
}

// MODE 0: L2 (needs an_f, bn_d); MODE 1: plain dot product (type 16)
template <int MODE>
__global__ void __launch_bounds__(256)
k_l2_simt(int d, long na, long nb, const float *__restrict__ a, long lda,
          const float *__restrict__ b, long ldb, const float *__restrict__ an_f,
          const double *__restrict__ bn_d, const float *__restrict__ a_weights,
          float *__restrict__ out, long ldd, int a_vec, int b_vec, int o_vec) {
  __shared__ __align__(16) float As[SK][SB + 4];
  __shared__ __align__(16) float Bs[SK][SB + 4];
  const int tid = threadIdx.x;
  const long i0 = (long)blockIdx.x * SB, j0 = (long)blockIdx.y * SB;
  const int tx = tid & 15, ty = tid >> 4;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;

  float acc[4][4];
#pragma unroll
  for (int jj = 0; jj < 4; jj++)
#pragma unroll
    for (int ii = 0; ii < 4; ii++) acc[jj][ii] = 0.0f;

  float ra[4], rb[4];
  load_tile_k4(a, na, d, lda, i0 + lrow, lk, a_vec, ra);
  load_tile_k4(b, nb, d, ldb, j0 + lrow, lk, b_vec, rb);
  for (int t0 = 0; t0 < d; t0 += SK) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      As[lk + u][lrow] = ra[u];
      Bs[lk + u][lrow] = rb[u];
    }
    __syncthreads();
    if (t0 + SK < d) {  // prefetch the next stage while this one is consumed
      load_tile_k4(a, na, d, lda, i0 + lrow, t0 + SK + lk, a_vec, ra);
      load_tile_k4(b, nb, d, ldb, j0 + lrow, t0 + SK + lk, b_vec, rb);
    }
#pragma unroll
    for (int kk = 0; kk < SK; kk++) {
      float4 av = *reinterpret_cast<const float4 *>(&As[kk][tx * 4]);
      float4 bv = *reinterpret_cast<const float4 *>(&Bs[kk][ty * 4]);
      const float avv[4] = {av.x, av.y, av.z, av.w};
      const float bvv[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int jj = 0; jj < 4; jj++)
#pragma unroll
        for (int ii = 0; ii < 4; ii++) acc[jj][ii] = fmaf(avv[ii], bvv[jj], acc[jj][ii]);
    }
    __syncthreads();
  }

  const long ib = i0 + tx * 4;
  float anv[4] = {0, 0, 0, 0}, wv[4] = {1, 1, 1, 1};
  if (MODE == 0) {
#pragma unroll
    for (int ii = 0; ii < 4; ii++)
      if (ib + ii < na) {
        anv[ii] = an_f[ib + ii];
        if (a_weights) wv[ii] = a_weights[ib + ii];
      }
  }
#pragma unroll
  for (int jj = 0; jj < 4; jj++) {
    const long j = j0 + ty * 4 + jj;
    if (j >= nb) continue;
    float r[4];
    if (MODE == 0) {
      const double bn = bn_d[j];
#pragma unroll
      for (int ii = 0; ii < 4; ii++) {
        float base = (float)(bn + (double)anv[ii]);                        // nn.c:123
        float v = __fadd_rn(base, __fmul_rn(-2.0f, acc[jj][ii]));          // nn.c:64, beta=1
        r[ii] = a_weights ? __fmul_rn(v, wv[ii]) : v;                      // nn.c:497-500
      }
    } else {
#pragma unroll
      for (int ii = 0; ii < 4; ii++) r[ii] = acc[jj][ii];
    }
    float *o = out + j * ldd + ib;
    if (o_vec && ib + 4 <= na) {
      *reinterpret_cast<float4 *>(o) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int ii = 0; ii < 4; ii++)
        if (ib + ii < na) o[ii] = r[ii];
    }
  }
}

// one query against nb rows; both norms double (yael/nn.c:132-154)
__global__ void __launch_bounds__(128) k_distances_1(int d, long nb, const float *__restrict__ q,
                                                      const float *__restrict__ b, long ldb,
                                                      float *__restrict__ out) {
  __shared__ float tile[4][32][33];
  extern __shared__ float qs[];
  for (int t = threadIdx.x; t < d; t += blockDim.x) qs[t] = q[t];
  __syncthreads();
  double qn = 0.0;
  for (int t = 0; t < d; t++) qn += (double)__fmul_rn(qs[t], qs[t]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long row = ((long)blockIdx.x * 4 + warp) * 32 + lane;
  const float *p = row < nb ? b + row * ldb : nullptr;
  double bn = 0.0;
  float dot = 0.0f;
  warp_rows_seq(p, d, tile[warp], [&](int t, float v) {
    bn += (double)__fmul_rn(v, v);
    dot = fmaf(qs[t], v, dot);  // sgemv as a sequential FMA chain
  });
  if (row < nb) out[row] = __fadd_rn((float)(bn + qn), __fmul_rn(-2.0f, dot));
}

// thread per (i, j) pair, the reference's scalar loops with their accumulator types
__global__ void __launch_bounds__(256)
k_alt_pairs(int type, int d, long na, long nb, const float *__restrict__ a, long lda,
            const float *__restrict__ b, long ldb, float *__restrict__ out, long ldd) {
  const long i = (long)blockIdx.x * 32 + (threadIdx.x & 31);
  const long j = (long)blockIdx.y * 8 + (threadIdx.x >> 5);
  if (i >= na || j >= nb) return;
  const float *x = a + i * lda, *y = b + j * ldb;
  float res;
  if ((type == 1 || type == 3) && (d % 4 == 0)) {
    // SSE2 variants (yael/nn.c:178-216): four float partial sums, combined left to right
    float p[4] = {0, 0, 0, 0};
    for (int t = 0; t < d; t += 4) {
#pragma unroll
      for (int l = 0; l < 4; l++) {
        float av = x[t + l], bv = y[t + l];
        float diff = __fsub_rn(av, bv);
        float term;
        if (type == 1) {
          term = fabsf(diff);
        } else {
          float sum = __fadd_rn(av, bv);
          term = (av != bv) ? __fdiv_rn(__fmul_rn(diff, diff), sum) : 0.0f;
        }
        p[l] = __fadd_rn(p[l], term);
      }
    }
    res = __fadd_rn(__fadd_rn(__fadd_rn(p[0], p[1]), p[2]), p[3]);
  } else {
    double s = 0.0;
    for (int t = 0; t < d; t++) {
      float av = x[t], bv = y[t];
      switch (type) {
        case 1: s += fabs((double)__fsub_rn(av, bv)); break;
        case 2: { double df = (double)__fsub_rn(av, bv); s += df * df; } break;
        case 3: {
          float sm = __fadd_rn(av, bv);
          if (sm != 0.0f) { double df = (double)__fsub_rn(av, bv); s += df * df / (double)sm; }
        } break;
        case 4: {
          float den = fabsf(__fadd_rn(av, bv));
          if (den != 0.0f) { double df = (double)__fsub_rn(av, bv); s += df * df / (double)den; }
        } break;
        case 5: s += (double)(av < bv ? av : bv); break;
        case 6: s += (double)__fmul_rn(av, bv); break;
        default: break;
      }
    }
    res = (float)s;
  }
  out[j * ldd + i] = res;
}

// ------------------------------------------------------------------------------------ host
static inline bool vec4_ok(const void *p, long ld) {
  return (((uintptr_t)p) & 15) == 0 && (ld % 4) == 0;
}

size_t l2_ws_bytes(long na, long nb) {
  return Carver::need(sizeof(float) * (size_t)na) + Carver::need(sizeof(double) * (size_t)nb);
}

int row_norms_seq(const float *x, long n, int d, long ld, float *out_f, double *out_d,
                  cudaStream_t st) {
  if (n <= 0) return 0;
  long blocks = (n + 127) / 128;
  k_row_norms_seq<<<(unsigned)blocks, 128, 0, st>>>(x, n, d, ld, out_f, out_d);
  YB_LAUNCH_CHECK();
  return 0;
}

int l2_matrix(int d, long na, long nb, const float *a, long lda, const float *b, long ldb,
              const float *an_f, const double *bn_d, const float *a_weights, float *out,
              long ldd, cudaStream_t st) {
  if (na <= 0 || nb <= 0) return 0;
  // grid.y is limited to 65535 blocks: walk the b side in slabs
  const long slab = 65535L * SB;
  for (long j0 = 0; j0 < nb; j0 += slab) {
    long nbj = nb - j0 < slab ? nb - j0 : slab;
    dim3 grid((unsigned)((na + SB - 1) / SB), (unsigned)((nbj + SB - 1) / SB));
    k_l2_simt<0><<<grid, 256, 0, st>>>(d, na, nbj, a, lda, b + j0 * ldb, ldb, an_f, bn_d + j0,
                                       a_weights, out + j0 * ldd, ldd, vec4_ok(a, lda),
                                       vec4_ok(b + j0 * ldb, ldb), vec4_ok(out + j0 * ldd, ldd));
    YB_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace yb

using namespace yb;

// engine of yb_cross_distances_l2: 0 = exact FP32 (k_l2_simt: the reference's rounding sequence,
// bit for bit), 1 = tensor cores with split-precision FP16 operands (within the north star's 1e-5
// relative; large packed problems), -1 = automatic (YAEL_B200_CROSS_ENGINE overrides)
static int g_cross_engine = -1;
static thread_local int g_last_cross_engine = 0;
extern "C" void yb_set_cross_engine(int engine) { g_cross_engine = engine; }
extern "C" int yb_last_cross_engine(void) { return g_last_cross_engine; }

extern "C" int yb_cross_distances_l2(int d, int na, int nb, const float *a, int lda,
                                      const float *b, int ldb, float *dist2, int ldd,
                                      yb_stream_t s) {
  if (na <= 0 || nb <= 0) return 0;
  if (d < 0 || lda < d || ldb < d || ldd < na) return fail(3, "yb_cross_distances_l2: bad leading dimension");
  Guard g;
  cudaStream_t st = stream_of(s);
  g_last_cross_engine = 0;
  {
    int engine = g_cross_engine;
    if (const char *e = getenv("YAEL_B200_CROSS_ENGINE")) engine = atoi(e);
    const bool packed = lda == d && ldb == d && d >= 1;
    const bool big = (double)na * nb * d >= 2e9 && na >= 256 && nb >= 1024;
    if (packed && (engine == 1 || (engine < 0 && big))) {
      const int rc = cross_l2_tensor(d, na, nb, a, b, dist2, ldd, st);
      if (rc == 0) {
        g_last_cross_engine = 1;
        return 0;
      }
      if (rc != -1000 && rc != -1001) return rc;  // does not qualify / out of FP16 range: exact engine
    }
  }
  ScratchScope ws(l2_ws_bytes(na, nb), st);
  Carver c(ws.p);
  float *an = c.take<float>(na);
  double *bn = c.take<double>(nb);
  int rc;
  if ((rc = row_norms_seq(a, na, d, lda, an, nullptr, st))) return rc;
  if ((rc = row_norms_seq(b, nb, d, ldb, nullptr, bn, st))) return rc;
  return l2_matrix(d, na, nb, a, lda, b, ldb, an, bn, nullptr, dist2, ldd, st);
}

extern "C" int yb_distances_1(int d, int nb, const float *a, const float *b, int ldb,
                               float *dist2, yb_stream_t s) {
  if (nb <= 0) return 0;
  Guard g;
  cudaStream_t st = stream_of(s);
  long blocks = ((long)nb + 127) / 128;
  k_distances_1<<<(unsigned)blocks, 128, sizeof(float) * (size_t)(d > 0 ? d : 1), st>>>(
      d, nb, a, b, ldb, dist2);
  YB_LAUNCH_CHECK();
  return 0;
}

extern "C" int yb_cross_distances_alt(int type, int d, int na, int nb, const float *a, int lda,
                                       const float *b, int ldb, float *dist2, int ldd,
                                       yb_stream_t s) {
  if (na <= 0 || nb <= 0) return 0;
  if (type == 12) return yb_cross_distances_l2(d, na, nb, a, lda, b, ldb, dist2, ldd, s);
  Guard g;
  cudaStream_t st = stream_of(s);
  if (type == 16) {  // mat_product (yael/nn.c:266-275): sgemm alpha=1 beta=0
    const long slab = 65535L * SB;
    for (long j0 = 0; j0 < nb; j0 += slab) {
      long nbj = nb - j0 < slab ? nb - j0 : slab;
      dim3 grid((unsigned)((na + SB - 1) / SB), (unsigned)((nbj + SB - 1) / SB));
      k_l2_simt<1><<<grid, 256, 0, st>>>(d, na, nbj, a, lda, b + j0 * (long)ldb, ldb, nullptr,
                                         nullptr, nullptr, dist2 + j0 * (long)ldd, ldd,
                                         vec4_ok(a, lda), vec4_ok(b + j0 * (long)ldb, ldb),
                                         vec4_ok(dist2 + j0 * (long)ldd, ldd));
      YB_LAUNCH_CHECK();
    }
    return 0;
  }
  if (type < 1 || type > 6) return fail(3, "yb_cross_distances_alt: unknown distance_type %d", type);
  const long slab = 65535L * 8;
  for (long j0 = 0; j0 < nb; j0 += slab) {
    long nbj = nb - j0 < slab ? nb - j0 : slab;
    dim3 grid((unsigned)((na + 31) / 32), (unsigned)((nbj + 7) / 8));
    k_alt_pairs<<<grid, 256, 0, st>>>(type, d, na, nbj, a, lda, b + j0 * (long)ldb, ldb,
                                      dist2 + j0 * (long)ldd, ldd);
    YB_LAUNCH_CHECK();
  }
  return 0;
}

// ------------------------------------------------------------------ bring-up: store patterns
// The output of compute_cross_distances is dist2[row * ld + query] (yael/nn.c:100-129): a CTA that
// owns 128 queries writes 512 contiguous bytes per database row, rows ld floats apart.  This
// micro-benchmark writes a whole [nb][ld] matrix with the access patterns an epilogue can use,
// work items in the tensor kernel's order (all query tiles of one 256-row tile run concurrently):
//   mode 0  lane = query, one 4-byte store per row: 128 bytes per warp instruction (the TMEM-lane
//           layout of the accumulators written as they come)
//   mode 1  lane = 4 queries, one 16-byte store per row: a warp instruction covers the CTA's 512
//           bytes of a row (what a tile staged through shared memory can do)
//   mode 2  as 1, two rows per instruction half-warp each (256 bytes per row piece)
namespace yb {
__global__ void __launch_bounds__(256)
k_dbg_store_pattern(float *__restrict__ out, long ld, int nq, int nb, int mode) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_q = (nq + 127) / 128, tiles_b = (nb + 255) / 256;
  for (long item = blockIdx.x; item < (long)tiles_q * tiles_b; item += gridDim.x) {
    const int qt = (int)(item % tiles_q), jt = (int)(item / tiles_q);
    if (mode == 0) {
      const int q = qt * 128 + (warp & 3) * 32 + lane;
      const long r0 = (long)jt * 256 + (warp >> 2) * 128;
      if (q < nq)
#pragma unroll 8
        for (int c = 0; c < 128; c++)
          if (r0 + c < nb) out[(r0 + c) * ld + q] = (float)c;
    } else if (mode == 1) {
      const int q = qt * 128 + lane * 4;
      const long r0 = (long)jt * 256 + warp * 32;
      if (q + 3 < nq)
#pragma unroll 8
        for (int c = 0; c < 32; c++)
          if (r0 + c < nb)
            *reinterpret_cast<float4 *>(out + (r0 + c) * ld + q) = make_float4((float)c, 1.f, 2.f, 3.f);
    } else {
      const int q = qt * 128 + (lane & 15) * 4 + (warp & 1) * 64;
      const long r0 = (long)jt * 256 + (warp >> 1) * 64 + (lane >> 4);
      if (q + 3 < nq)
#pragma unroll 8
        for (int c = 0; c < 64; c += 2)
          if (r0 + c < nb)
            *reinterpret_cast<float4 *>(out + (r0 + c) * ld + q) = make_float4((float)c, 1.f, 2.f, 3.f);
    }
  }
}
}  // namespace yb

extern "C" double yb_debug_store_pattern_gbs(float *out, long ld, int nq, int nb, int mode, yb_stream_t s) {
  Guard g;
  cudaStream_t st = stream_of(s);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  yb::k_dbg_store_pattern<<<sm_count(), 256, 0, st>>>(out, ld, nq, nb, mode);
  cudaEventRecord(e0, st);
  for (int i = 0; i < 3; i++) yb::k_dbg_store_pattern<<<sm_count(), 256, 0, st>>>(out, ld, nq, nb, mode);
  cudaEventRecord(e1, st);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  count_launch(4);
  if (ms <= 0.f) return 0.0;
  return 3.0 * 4.0 * (double)nq * nb / (ms * 1e-3) / 1e9;
}
