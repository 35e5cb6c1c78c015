/* yael_nn.c -- the drop-in nearest-neighbour API (include/yael/nn.h) on top of the yb_ C ABI.
 *
 * Host side only: argument checks in the reference's style (assert on precondition
 * violations, yael/nn.c:456), staging of host buffers, the calls into the device layer.
 * Each function names the reference lines it replaces.
 */
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../../include/yael/nn.h"
#include "../../../include/yael/vector.h"
#include "yb_host.h"

/* yael/nn.c:451-525 (k > 1) and 383-446 (k == 1) */
void knn_full(int distance_type, int nq, int nb, int d, int k, const float *b, const float *q,
              const float *b_weights, int *assign, float *dis) {
  assert(k <= nb); /* yael/nn.c:456 */
  if (nq <= 0 || k <= 0) return;
  /* L2 on a host-resident database: the transfer is overlapped with the scan inside
   * yb_knn_l2_hostbase (which falls back to copy-then-scan for shapes that do not qualify) */
  const int l2 = (distance_type == 2 || distance_type == 12);
  const int feed_host = l2 && !b_weights && !ybh_is_device_ptr(b);
  /* large problems on host buffers: sharded over the box's GPUs (yb_mgpu.cu; -1 = does not
   * qualify, take the one-GPU path below) */
  if (feed_host && !ybh_is_device_ptr(q) && !ybh_is_device_ptr(assign) && !ybh_is_device_ptr(dis)) {
    int rc = yb_mgpu_knn_full(nq, nb, d, k, b, q, assign, dis);
    if (rc == 0) return;
    if (rc > 0) ybh_die("yb_mgpu_knn_full", rc);
  }
  ybh_arg ab;
  if (feed_host) {
    ab = ybh_out((void *)b, sizeof(float) * (size_t)nb * d); /* device scratch, no copy */
  } else {
    ab = ybh_in(b, sizeof(float) * (size_t)nb * d);
  }
  ybh_arg aq = ybh_in(q, sizeof(float) * (size_t)nq * d);
  ybh_arg aw = ybh_in(b_weights, sizeof(float) * (size_t)nb);
  ybh_arg oa = ybh_out(assign, sizeof(int) * (size_t)nq * k);
  ybh_arg od = ybh_out(dis, sizeof(float) * (size_t)nq * k);
  if (feed_host) {
    YBH_CHECK(yb_knn_l2_hostbase(nq, nb, d, k, b, (float *)ab.dev, (const float *)aq.dev,
                                 (int *)oa.dev, (float *)od.dev, 0, NULL));
  } else if (l2) {
    YBH_CHECK(yb_knn_l2(nq, nb, d, k, (const float *)ab.dev, (const float *)aq.dev,
                        (const float *)aw.dev, (int *)oa.dev, (float *)od.dev, 0, NULL));
  } else if ((distance_type >= 1 && distance_type <= 6) || distance_type == 16) {
    YBH_CHECK(yb_knn_alt(distance_type, nq, nb, d, k, (const float *)ab.dev, (const float *)aq.dev,
                         (const float *)aw.dev, (int *)oa.dev, (float *)od.dev, NULL));
  } else {
    fprintf(stderr, "yael_b200: knn_full: unknown distance_type %d\n", distance_type);
    abort();
  }
  ybh_finish(&oa, 1);
  ybh_finish(&od, 1);
  ybh_finish(&ab, 0);
  ybh_finish(&aq, 0);
  ybh_finish(&aw, 0);
  ybh_sync();
}

/* yael/nn.c:679-699: the reference splits the queries over n_thread OpenMP tasks and its
 * results do not depend on the split; one device call covers all queries. */
void knn_full_thread(int distance_type, int nq, int nb, int d, int k, const float *b,
                     const float *q, const float *b_weights, int *assign, float *dis,
                     int n_thread) {
  (void)n_thread;
  knn_full(distance_type, nq, nb, d, k, b, q, b_weights, assign, dis);
}

/* yael/nn.c:608-621 */
double nn(int npt, int nclust, int d, const float *codebook, const float *coords, int *vw) {
  float *vwdis = fvec_new(npt);
  knn_full(2, npt, nclust, d, 1, codebook, coords, NULL, vw, vwdis);
  double toterr = fvec_sum(vwdis, npt);
  free(vwdis);
  return toterr;
}

/* yael/nn.c:624-632: returned block is the caller's to free */
float *knn(int npt, int nclust, int d, int k, const float *codebook, const float *coords,
           int *vw) {
  float *vwdis = fvec_new((long)npt * k);
  knn_full(2, npt, nclust, d, k, codebook, coords, NULL, vw, vwdis);
  return vwdis;
}

/* yael/nn.c:704-711 */
float *knn_thread(int npt, int nclust, int d, int k, const float *codebook, const float *coords,
                  int *vw, int n_thread) {
  float *vwdis = fvec_new((long)k * npt);
  knn_full_thread(2, npt, nclust, d, k, codebook, coords, NULL, vw, vwdis, n_thread);
  return vwdis;
}

/* yael/nn.c:715-726 */
double nn_thread(int npt, int nclust, int d, const float *codebook, const float *coords, int *vw,
                 int n_thread) {
  float *vwdis = fvec_new(npt);
  knn_full_thread(2, npt, nclust, d, 1, codebook, coords, NULL, vw, vwdis, n_thread);
  double toterr = fvec_sum(vwdis, npt);
  free(vwdis);
  return toterr;
}

/* yael/nn.c:528-580 */
void knn_reorder_shortlist(int n, int nb, int d, int k, const float *b, const float *v, int *idx,
                           float *dis) {
  if (n <= 0 || k <= 0) return;
  ybh_arg ab = ybh_in(b, sizeof(float) * (size_t)nb * d);
  ybh_arg av = ybh_in(v, sizeof(float) * (size_t)n * d);
  ybh_arg ai = ybh_in(idx, sizeof(int) * (size_t)n * k); /* in/out */
  ybh_arg od = ybh_out(dis, sizeof(float) * (size_t)n * k);
  /* entries past the first negative id are left untouched by the reference: preload dis */
  if (od.owned) YBH_CHECK(yb_h2d(od.dev, dis, od.bytes, NULL));
  YBH_CHECK(yb_knn_reorder_shortlist(n, nb, d, k, (const float *)ab.dev, (const float *)av.dev,
                                     (int *)ai.dev, (float *)od.dev, NULL));
  ybh_finish(&od, 1);
  ybh_finish(&ai, 1);
  ybh_finish(&ab, 0);
  ybh_finish(&av, 0);
  ybh_sync();
}

/* yael/nn.c:583-600: a handful of scattered exact distances against a partial base that is
 * streamed from disk by the caller: pointer chasing over host data, a few thousand flops; kept
 * on the host exactly as the reference computes it (double accumulation of squared
 * differences, yael/vector.c:2348-2359). */
void knn_recompute_exact_dists(int nq, int nb, int d, int k, const float *b, const float *v,
                               int label0, int *kp, const int *idx, float *dis) {
  long q, i;
  for (q = 0; q < nq; q++) {
    const float *vq = v + (long)d * q;
    for (i = kp[q]; i < k; i++) {
      long j = idx[q * k + i] - label0;
      assert(j >= 0);
      if (j >= nb) break;
      dis[q * k + i] = (float)fvec_distance_L2sqr(vq, b + j * d, d);
    }
    kp[q] = (int)i;
  }
}

/* yael/nn.c:100-129 */
void compute_cross_distances_nonpacked(int d, int na, int nb, const float *a, int lda,
                                       const float *b, int ldb, float *dist2, int ldd) {
  if (na <= 0 || nb <= 0) return;
  ybh_arg aa = ybh_in(a, sizeof(float) * ((size_t)(na - 1) * lda + d));
  ybh_arg ab = ybh_in(b, sizeof(float) * ((size_t)(nb - 1) * ldb + d));
  /* non-packed output: the gaps between lines belong to the caller, so stage a packed
   * matrix and copy line by line when ldd > na */
  int packed = (ldd == na) || ybh_is_device_ptr(dist2);
  if (packed) {
    ybh_arg od = ybh_out(dist2, sizeof(float) * ((size_t)(nb - 1) * ldd + na));
    YBH_CHECK(yb_cross_distances_l2(d, na, nb, (const float *)aa.dev, lda, (const float *)ab.dev,
                                    ldb, (float *)od.dev, ldd, NULL));
    ybh_finish(&od, 1);
  } else {
    float *tmp = fvec_new((long)na * nb);
    ybh_arg od = ybh_out(tmp, sizeof(float) * (size_t)na * nb);
    YBH_CHECK(yb_cross_distances_l2(d, na, nb, (const float *)aa.dev, lda, (const float *)ab.dev,
                                    ldb, (float *)od.dev, na, NULL));
    ybh_finish(&od, 1);
    for (long j = 0; j < nb; j++)
      memcpy(dist2 + j * ldd, tmp + j * na, sizeof(float) * (size_t)na);
    free(tmp);
  }
  ybh_finish(&aa, 0);
  ybh_finish(&ab, 0);
  ybh_sync();
}

/* yael/nn.c:92-97 */
void compute_cross_distances(int d, int na, int nb, const float *a, const float *b,
                             float *dist2) {
  compute_cross_distances_nonpacked(d, na, nb, a, d, b, d, dist2, na);
}

/* yael/nn.c:777-792 (the slice layout does not change any value) */
void compute_cross_distances_thread(int d, int na, int nb, const float *a, const float *b,
                                    float *dist2, int nt) {
  (void)nt;
  compute_cross_distances(d, na, nb, a, b, dist2);
}

/* yael/nn.c:280-350 */
void compute_cross_distances_alt_nonpacked(int distance_type, int d, int na, int nb,
                                           const float *a, int lda, const float *b, int ldb,
                                           float *dist2, int ldd) {
  if (na <= 0 || nb <= 0) return;
  if (distance_type == 12) {
    compute_cross_distances_nonpacked(d, na, nb, a, lda, b, ldb, dist2, ldd);
    return;
  }
  if (!((distance_type >= 1 && distance_type <= 6) || distance_type == 16)) {
    /* the reference silently writes zeros for an unknown type (nn.c:316-344); be loud */
    fprintf(stderr, "yael_b200: compute_cross_distances_alt: unknown distance_type %d\n",
            distance_type);
    abort();
  }
  ybh_arg aa = ybh_in(a, sizeof(float) * ((size_t)(na - 1) * lda + d));
  ybh_arg ab = ybh_in(b, sizeof(float) * ((size_t)(nb - 1) * ldb + d));
  float *tmp = NULL;
  int direct = (ldd == na) || ybh_is_device_ptr(dist2);
  ybh_arg od;
  if (direct) {
    od = ybh_out(dist2, sizeof(float) * ((size_t)(nb - 1) * ldd + na));
  } else {
    tmp = fvec_new((long)na * nb);
    od = ybh_out(tmp, sizeof(float) * (size_t)na * nb);
  }
  YBH_CHECK(yb_cross_distances_alt(distance_type, d, na, nb, (const float *)aa.dev, lda,
                                   (const float *)ab.dev, ldb, (float *)od.dev,
                                   direct ? ldd : na, NULL));
  ybh_finish(&od, 1);
  if (!direct) {
    for (long j = 0; j < nb; j++)
      memcpy(dist2 + j * ldd, tmp + j * na, sizeof(float) * (size_t)na);
    free(tmp);
  }
  ybh_finish(&aa, 0);
  ybh_finish(&ab, 0);
  ybh_sync();
}

/* yael/nn.c:352-356 */
void compute_cross_distances_alt(int distance_type, int d, int na, int nb, const float *a,
                                 const float *b, float *dist2) {
  compute_cross_distances_alt_nonpacked(distance_type, d, na, nb, a, d, b, d, dist2, na);
}

/* yael/nn.c:795-810 */
void compute_cross_distances_alt_thread(int distance_type, int d, int na, int nb, const float *a,
                                        const float *b, float *dist2, int nt) {
  (void)nt;
  compute_cross_distances_alt(distance_type, d, na, nb, a, b, dist2);
}

/* yael/nn.c:132-154 */
void compute_distances_1_nonpacked(int d, int nb, const float *a, const float *b, int ldb,
                                   float *dist2) {
  if (nb <= 0) return;
  ybh_arg aa = ybh_in(a, sizeof(float) * (size_t)d);
  ybh_arg ab = ybh_in(b, sizeof(float) * ((size_t)(nb - 1) * ldb + d));
  ybh_arg od = ybh_out(dist2, sizeof(float) * (size_t)nb);
  YBH_CHECK(yb_distances_1(d, nb, (const float *)aa.dev, (const float *)ab.dev, ldb,
                           (float *)od.dev, NULL));
  ybh_finish(&od, 1);
  ybh_finish(&aa, 0);
  ybh_finish(&ab, 0);
  ybh_sync();
}

/* yael/nn.c:156-162 */
void compute_distances_1(int d, int nb, const float *a, const float *b, float *dist2) {
  compute_distances_1_nonpacked(d, nb, a, b, d, dist2);
}

/* yael/nn.c:830-860 */
void compute_distances_1_thread(int d, int nb, const float *a, const float *b, float *dist2,
                                int n_thread) {
  (void)n_thread;
  compute_distances_1_nonpacked(d, nb, a, b, d, dist2);
}

void compute_distances_1_nonpacked_thread(int d, int nb, const float *a, const float *b, int ldb,
                                          float *dist2, int n_thread) {
  (void)n_thread;
  compute_distances_1_nonpacked(d, nb, a, b, ldb, dist2);
}
