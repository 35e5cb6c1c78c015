/* yael_vlad.c -- include/yael/vlad.h on top of the yb_ C ABI (yael/vlad.c): the assignment is the
 * library's k = 1 / k = ma search on the device, the aggregation yb_vlad_accumulate /
 * yb_bof_accumulate; nothing but staging happens on the host. */
#include <assert.h>
#include <stdlib.h>

#include "../../../include/yael/vlad.h"
#include "yb_host.h"

typedef struct {
  ybh_arg cent, v;
  int *assign;  /* device [n * ma] */
  float *dis;   /* device [n * ma] */
} vlad_ctx;

/* nn (n, k, d, centroids, v, assign) (yael/nn.c:608-621) with everything left on the device */
static vlad_ctx assign_points(int k, int d, const float *centroids, int n, const float *v, int ma) {
  vlad_ctx c;
  c.cent = ybh_in(centroids, sizeof(float) * (size_t)k * d);
  c.v = ybh_in(v, sizeof(float) * (size_t)n * d);
  c.assign = (int *)yb_malloc(sizeof(int) * (size_t)(n > 0 ? n : 1) * ma);
  c.dis = (float *)yb_malloc(sizeof(float) * (size_t)(n > 0 ? n : 1) * ma);
  if (n > 0)
    YBH_CHECK(yb_knn_l2(n, k, d, ma, (const float *)c.cent.dev, (const float *)c.v.dev, NULL, c.assign,
                        c.dis, 0, NULL));
  return c;
}

static void release(vlad_ctx *c) {
  yb_free(c->assign);
  yb_free(c->dis);
  ybh_finish(&c->cent, 0);
  ybh_finish(&c->v, 0);
}

void vlad_compute_weighted(int k, int d, const float *centroids, int n, const float *v,
                           const float *weights, float *desc) {
  vlad_ctx c = assign_points(k, d, centroids, n, v, 1);
  ybh_arg aw = ybh_in(weights, sizeof(float) * (size_t)n);
  ybh_arg od = ybh_out(desc, sizeof(float) * (size_t)k * d);
  YBH_CHECK(yb_vlad_accumulate(k, d, (const float *)c.cent.dev, n, NULL, (const float *)c.v.dev, c.assign,
                               (const float *)aw.dev, (float *)od.dev, NULL));
  ybh_finish(&od, 1);
  ybh_finish(&aw, 0);
  ybh_sync();
  release(&c);
}

void vlad_compute(int k, int d, const float *centroids, int n, const float *v, float *desc) {
  vlad_compute_weighted(k, d, centroids, n, v, NULL, desc);
}

void vlad_compute_subsets(int k, int d, const float *centroids, int n, const float *v, int n_subset,
                          const int *subset_indexes, const int *subset_ends, float *desc) {
  vlad_ctx c = assign_points(k, d, centroids, n, v, 1);
  const int total = n_subset > 0 ? subset_ends[n_subset - 1] : 0;
  ybh_arg ai = ybh_in(subset_indexes, sizeof(int) * (size_t)total);
  ybh_arg od = ybh_out(desc, sizeof(float) * (size_t)k * d * n_subset);
  int ss, begin = 0;
  for (ss = 0; ss < n_subset; ss++) {
    const int end = subset_ends[ss];
    YBH_CHECK(yb_vlad_accumulate(k, d, (const float *)c.cent.dev, end - begin,
                                 (const int *)ai.dev + begin, (const float *)c.v.dev, c.assign, NULL,
                                 (float *)od.dev + (size_t)ss * k * d, NULL));
    begin = end;
  }
  ybh_finish(&od, 1);
  ybh_finish(&ai, 0);
  ybh_sync();
  release(&c);
}

void bof_compute_ma(int k, int d, const float *centroids, int n, const float *v, int *desc, int ma,
                    float alpha, int nt) {
  (void)alpha;
  (void)nt;
  assert(ma >= 1 && ma <= k);
  vlad_ctx c = assign_points(k, d, centroids, n, v, ma);
  ybh_arg od = ybh_out(desc, sizeof(int) * (size_t)k);
  YBH_CHECK(yb_bof_accumulate(k, (long)n * ma, NULL, c.assign, (long)n * ma, (int *)od.dev, NULL, NULL));
  ybh_finish(&od, 1);
  ybh_sync();
  release(&c);
}

void bof_compute(int k, int d, const float *centroids, int n, const float *v, int *desc) {
  bof_compute_ma(k, d, centroids, n, v, desc, 1, 0.f, 1);
}

void bof_compute_subsets(int k, int d, const float *centroids, int n, const float *v, int n_subset,
                         const int *subset_indexes, const int *subset_ends, float *desc) {
  vlad_ctx c = assign_points(k, d, centroids, n, v, 1);
  const int total = n_subset > 0 ? subset_ends[n_subset - 1] : 0;
  ybh_arg ai = ybh_in(subset_indexes, sizeof(int) * (size_t)total);
  ybh_arg od = ybh_out(desc, sizeof(float) * (size_t)k * n_subset);
  int *cnt = (int *)yb_malloc(sizeof(int) * (size_t)k);
  int ss, begin = 0;
  for (ss = 0; ss < n_subset; ss++) {
    const int end = subset_ends[ss];
    YBH_CHECK(yb_bof_accumulate(k, end - begin, (const int *)ai.dev + begin, c.assign, n, cnt,
                                (float *)od.dev + (size_t)ss * k, NULL));
    begin = end;
  }
  ybh_finish(&od, 1);
  ybh_finish(&ai, 0);
  ybh_sync();
  yb_free(cnt);
  release(&c);
}
