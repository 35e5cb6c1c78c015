// yb_vlad.cu -- the aggregation step of the consumers of the k = 1 search:
// VLAD residual sums and bag-of-features histograms (yael/vlad.c:10-139).  The assignment itself
// is yb_knn_l2 with k = 1 (the tensor-core margin mode); what is left is a segmented reduction
// whose ORDER defines the result: the reference adds fl32(v_i - c) to desc[assign_i] for
// i = 0, 1, 2, ... (vlad.c:20-23), or in list order for the *_subsets variants (vlad.c:66-73).
//
// k_vlad_rows: one warp per centroid; the warp scans the (listed) points 32 at a time, ballots
// the ones assigned to its centroid and adds their residual rows in scan order -- the reference's
// order, so the descriptor equals the reference's bit for bit.  The scan costs n * k / 32 loads
// of `assign` from L2, which is nothing for the codebook sizes VLAD / BoF use (k <= a few
// thousand; refused above 16384: the k-means update path is the tool for large k).
#include "yb_common.cuh"
#include "yb_internal.cuh"

namespace yb {

template <int NV>
__device__ __forceinline__ void vlad_load(const float *p, float (&o)[NV]) {
  if constexpr (NV == 4) {
    const float4 x = __ldg(reinterpret_cast<const float4 *>(p));
    o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w;
  } else {
    o[0] = __ldg(p);
  }
}

template <int NV>
__global__ void __launch_bounds__(128)
k_vlad_rows(int k, int d, const float *__restrict__ cent, long n_list, const int *__restrict__ list,
            const float *__restrict__ v, const int *__restrict__ assign,
            const float *__restrict__ weights, float *__restrict__ desc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 4 + warp;
  if (c >= k) return;
  for (int t0 = 0; t0 < d; t0 += 32 * NV) {
    const int t = t0 + lane * NV;
    const bool active = t < d;
    float cc[NV], acc[NV];
#pragma unroll
    for (int x = 0; x < NV; x++) acc[x] = cc[x] = 0.f;
    if (active) vlad_load<NV>(cent + (size_t)c * d + t, cc);
    for (long ii0 = 0; ii0 < n_list; ii0 += 32) {
      const long ii = ii0 + lane;
      const int i = ii < n_list ? (list ? list[ii] : (int)ii) : -1;
      const int a = i >= 0 ? assign[i] : -1;
      unsigned m = __ballot_sync(0xffffffffu, a == c);
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const int pi = __shfl_sync(0xffffffffu, i, b);
        if (active) {
          float x[NV];
          vlad_load<NV>(v + (size_t)pi * d + t, x);
          const float w = weights ? __ldg(weights + pi) : 1.f;
#pragma unroll
          for (int y = 0; y < NV; y++) {
            float r = __fsub_rn(x[y], cc[y]);         // fl32(v - c)           (vlad.c:22)
            if (weights) r = __fmul_rn(r, w);         // fl32((v - c) * w)     (vlad.c:45)
            acc[y] = __fadd_rn(acc[y], r);
          }
        }
      }
    }
    if (active) {
#pragma unroll
      for (int y = 0; y < NV; y++) desc[(size_t)c * d + t + y] = acc[y];
    }
  }
}

__global__ void k_bof_hist(long n_list, const int *__restrict__ list, const int *__restrict__ assign, long n_assign,
                           int k, int *__restrict__ desc) {
  for (long ii = (long)blockIdx.x * blockDim.x + threadIdx.x; ii < n_list; ii += (long)gridDim.x * blockDim.x) {
    const long i = list ? list[ii] : ii;
    if (i < 0 || i >= n_assign) continue;
    const int a = assign[i];
    if (a >= 0 && a < k) atomicAdd(&desc[a], 1);
  }
}

__global__ void k_i2f(const int *__restrict__ in, long n, float *__restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

}  // namespace yb

using namespace yb;

// desc[k][d] = sum over the listed points (list == NULL: points 0 .. n_list-1), in list order, of
// fl32(v_i - centroids[assign_i]) (times weights[i] when given): vlad_compute /
// vlad_compute_weighted / one subset of vlad_compute_subsets (yael/vlad.c:10-79)
extern "C" int yb_vlad_accumulate(int k, int d, const float *centroids, long n_list, const int *list,
                                  const float *v, const int *assign, const float *weights, float *desc,
                                  yb_stream_t s) {
  if (k <= 0 || d <= 0) return 0;
  if (k > 16384) return fail(3, "yb_vlad_accumulate: k = %d (codebooks of up to 16384 centroids)", k);
  Guard g;
  cudaStream_t st = stream_of(s);
  const bool v4 = (d & 3) == 0 && ((((uintptr_t)v) | ((uintptr_t)centroids)) & 15) == 0;
  const unsigned blocks = (unsigned)((k + 3) / 4);
  if (v4)
    k_vlad_rows<4><<<blocks, 128, 0, st>>>(k, d, centroids, n_list, list, v, assign, weights, desc);
  else
    k_vlad_rows<1><<<blocks, 128, 0, st>>>(k, d, centroids, n_list, list, v, assign, weights, desc);
  YB_LAUNCH_CHECK();
  return 0;
}

// desc[k] (+)= number of listed entries of assign[0 .. n_assign) per centroid: bof_compute,
// bof_compute_ma, one subset of bof_compute_subsets (yael/vlad.c:82-139).  desc_f != NULL: the
// counts are also written as floats (the subsets variant returns floats).
extern "C" int yb_bof_accumulate(int k, long n_list, const int *list, const int *assign, long n_assign,
                                 int *desc, float *desc_f, yb_stream_t s) {
  if (k <= 0) return 0;
  Guard g;
  cudaStream_t st = stream_of(s);
  YB_CUDA(cudaMemsetAsync(desc, 0, sizeof(int) * (size_t)k, st));
  if (n_list > 0) {
    long blocks = (n_list + 255) / 256;
    if (blocks > 4L * sm_count()) blocks = 4L * sm_count();
    k_bof_hist<<<(unsigned)blocks, 256, 0, st>>>(n_list, list, assign, n_assign, k, desc);
    YB_LAUNCH_CHECK();
  }
  if (desc_f) {
    k_i2f<<<(unsigned)((k + 255) / 256), 256, 0, st>>>(desc, k, desc_f);
    YB_LAUNCH_CHECK();
  }
  return 0;
}
