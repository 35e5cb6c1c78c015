// yb_gmm.cu -- the GMM E-step (yael/gmm.c:211-367): posteriors p(c_j | x_i) of
// a diagonal-covariance mixture for n points.  The reference computes the squared Mahalanobis
// distances as TWO sgemm calls of the same shape onto a matrix pre-filled with sum mu^2/sigma
// (gmm.c:221-254):
//
//     m[i][j] = fl32(sum_l mu_jl^2 / sigma_jl)  +  <1/sigma_j, x_i^2>  +  (-2) <mu_j/sigma_j, x_i>
//
// then the log-domain combination logdet_j - 0.5 m + log w_j in double (gmm.c:357) and a
// max-shifted softmax per point (gmm.c:262-300).
//
// k_gmm_logp is the contraction: a 64 x 64 (points x components) tile per CTA, the K loop staged
// through shared memory 32 coordinates at a time, 4 x 4 outputs per thread, both dot products as
// sequential FP32 FMA chains in coordinate order (the oracle's ORC_DOT_F32_SEQ order: the value is
// DEFINED, not whatever a BLAS does) and the reference's two rounded adds onto mu2; the epilogue
// writes the log-domain value.  k_gmm_softmax is one warp per point: float max, exp in double
// rounded to float, the sum SEQUENTIAL in component order on one lane (the reference's order),
// (float)(1.0 / s) scaling.  Bound: FP32 FMA pipe (4 n k d FLOP); the softmax is one pass over p.
#include "yb_common.cuh"
#include "yb_internal.cuh"

namespace yb {

constexpr int GT = 64;   // tile edge (points and components)
constexpr int GK = 32;   // coordinates per stage

__global__ void __launch_bounds__(256)
k_gmm_logp(int n, int k, int d, const float *__restrict__ v, const float *__restrict__ inv_sigma,
           const float *__restrict__ mu_sigma, const float *__restrict__ mu2,
           const float *__restrict__ logdetnr, const float *__restrict__ lg, float *__restrict__ p) {
  // [coordinate][row]: a thread's 4 rows are consecutive words, the 16 threads of a half warp
  // read 64 consecutive words (no conflicts), rows of the other operand are broadcasts
  __shared__ float sv[GK][GT + 4], si[GK][GT + 4], sm[GK][GT + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // tx: components, ty: points
  const long i0 = (long)blockIdx.y * GT;
  const int j0 = blockIdx.x * GT;
  float a1[4][4], a2[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) a1[a][b] = a2[a][b] = 0.f;
  for (int t0 = 0; t0 < d; t0 += GK) {
    // stage: 64 rows x 32 coordinates of each operand; thread -> (row = tid / 4 + 0 | 32 ..., 8 floats)
    for (int e = threadIdx.x; e < GT * GK; e += 256) {
      const int r = e / GK, t = e % GK;
      const bool tin = t0 + t < d;
      const long pi = i0 + r;
      const int cj = j0 + r;
      sv[t][r] = (tin && pi < n) ? __ldg(v + (size_t)pi * d + t0 + t) : 0.f;
      si[t][r] = (tin && cj < k) ? __ldg(inv_sigma + (size_t)cj * d + t0 + t) : 0.f;
      sm[t][r] = (tin && cj < k) ? __ldg(mu_sigma + (size_t)cj * d + t0 + t) : 0.f;
    }
    __syncthreads();
    const int tc = d - t0 < GK ? d - t0 : GK;   // the chains stop at d: padding zeros would still
    for (int t = 0; t < tc; t++) {               // be exact, but -0 + 0 could flip a sign bit
      float x[4], x2[4], is[4], ms[4];
#pragma unroll
      for (int a = 0; a < 4; a++) {
        x[a] = sv[t][ty * 4 + a];
        x2[a] = __fmul_rn(x[a], x[a]);           // v2 = v * v in float (gmm.c:235-236)
      }
#pragma unroll
      for (int b = 0; b < 4; b++) {
        is[b] = si[t][tx * 4 + b];
        ms[b] = sm[t][tx * 4 + b];
      }
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
          a1[a][b] = fmaf(is[b], x2[a], a1[a][b]);
          a2[a][b] = fmaf(ms[b], x[a], a2[a][b]);
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const long pi = i0 + ty * 4 + a;
    if (pi >= n) continue;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int cj = j0 + tx * 4 + b;
      if (cj >= k) continue;
      float m = __ldg(mu2 + cj);
      m = __fadd_rn(m, a1[a][b]);                          // sgemm 1: C += A'B          (gmm.c:244)
      m = __fadd_rn(m, __fmul_rn(-2.0f, a2[a][b]));        // sgemm 2: C += -2 A'B       (gmm.c:254)
      // logdetnr[j] - 0.5 * p + lg[j] in double, stored as float                 (gmm.c:357)
      const double lp = (double)__ldg(logdetnr + cj) - 0.5 * (double)m + (double)__ldg(lg + cj);
      p[(size_t)pi * k + cj] = (float)lp;
    }
  }
}

// softmax_ref (gmm.c:262-300), in place, one warp per point
__global__ void __launch_bounds__(128)
k_gmm_softmax(long n, int k, float *__restrict__ p, float *__restrict__ coeffs) {
  const long i = (long)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  float *row = p + (size_t)i * k;
  const float norm_to_0 = 16.636f;  // log(2^24)
  float mx = -1e30f;
  for (int l = lane; l < k; l += 32) {
    const float f = row[l];
    if (f > mx) mx = f;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, mx, o);
    if (t > mx) mx = t;
  }
  const float lim = __fsub_rn(mx, norm_to_0);
  for (int l = lane; l < k; l += 32) {
    const float f = row[l];
    row[l] = f >= lim ? (float)exp((double)__fsub_rn(f, mx)) : 0.f;
  }
  __syncwarp();
  float s = 0.f;
  if (lane == 0)
    for (int l = 0; l < k; l++) s = __fadd_rn(s, row[l]);   // the reference's order: l = 0, 1, 2, ...
  s = __shfl_sync(0xffffffffu, s, 0);
  if (coeffs && lane == 0) coeffs[i] = (float)(log((double)s) + (double)mx);
  if (s != 0.f) {
    const float is = (float)(1.0 / (double)s);
    for (int l = lane; l < k; l += 32) row[l] = __fmul_rn(row[l], is);
  }
}

}  // namespace yb

using namespace yb;

// p[n][k] = posteriors.  inv_sigma = (float)(1.0 / sigma), mu_sigma = mu / sigma ([k][d]), mu2[k] =
// (float) sum_l mu^2 / sigma (double sum), logdetnr[k], lg[k] (log weights or zeros): the O(k d)
// tables the reference prepares on the host (gmm.c:221-226,239-250,318-349); coeffs (may be NULL)
// receives log(s) + max per point (softmax_ref's optional output).  All device pointers.
extern "C" int yb_gmm_posteriors(long n, int k, int d, const float *v, const float *inv_sigma,
                                 const float *mu_sigma, const float *mu2, const float *logdetnr,
                                 const float *lg, float *p, float *coeffs, yb_stream_t s) {
  if (n <= 0 || k <= 0) return 0;
  if (d <= 0) return fail(3, "yb_gmm_posteriors: d = %d", d);
  Guard g;
  cudaStream_t st = stream_of(s);
  const long ty = (n + GT - 1) / GT;
  if (ty > 65535L * 1024) return fail(3, "yb_gmm_posteriors: n = %ld is too large", n);
  // grid.y is limited to 65535: walk the points in slabs
  for (long y0 = 0; y0 < ty; y0 += 65535) {
    const long ny = ty - y0 < 65535 ? ty - y0 : 65535;
    const long r0 = y0 * GT;
    const long rows = n - r0 < ny * GT ? n - r0 : ny * GT;
    dim3 grid((unsigned)((k + GT - 1) / GT), (unsigned)ny);
    k_gmm_logp<<<grid, 256, 0, st>>>((int)rows, k, d, v + (size_t)r0 * d, inv_sigma, mu_sigma, mu2,
                                     logdetnr, lg, p + (size_t)r0 * k);
    YB_LAUNCH_CHECK();
  }
  k_gmm_softmax<<<(unsigned)((n + 3) / 4), 128, 0, st>>>(n, k, p, coeffs);
  YB_LAUNCH_CHECK();
  return 0;
}
