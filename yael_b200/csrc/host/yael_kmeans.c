/* yael_kmeans.c -- Lloyd's k-means driver (include/yael/kmeans.h) on top of the yb_ C ABI.
 *
 * What stays on the host, in C, exactly as the reference orders it (yael/kmeans.c:332-447,
 * 213-329): the run / iteration control flow, every rand_r draw (init seed, core seed, one
 * draw per iteration, the empty-cluster splits), random / k-means++ selection, the
 * empty-cluster split itself, the stopping rule and the progress messages.
 * What runs on the device every iteration: the assignment (yb_knn_l2, k = 1), the histogram,
 * the centroid sums, qerr, and the scaling (yb_kmeans_accumulate / yb_kmeans_scale).
 */
#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../../include/yael/kmeans.h"
#include "../../../include/yael/machinedeps.h"
#include "../../../include/yael/nn.h"
#include "../../../include/yael/vector.h"
#include "yb_host.h"

static double drand_r(unsigned int *seed) { /* yael/kmeans.c:22-24 */
  return rand_r(seed) / ((double)RAND_MAX + 1.0);
}

/* yael/kmeans.c:166-209, on host copies of the centroids; touches rows only when a cluster
 * is empty */
static int reassign_empty(int d, int k, float *centroids, const int *nassign,
                          unsigned int seed) {
  int c, j, moved = 0;
  float *proba = fvec_new(k);
  float *eps = fvec_new(d);
  for (c = 0; c < k; c++)
    proba[c] = (nassign[c] < 2 ? 0 : nassign[c] * nassign[c] - 1);
  fvec_normalize(proba, k, 1);
  for (c = 0; c < k; c++) {
    if (nassign[c] != 0) continue;
    moved++;
    double rd = drand_r(&seed);
    for (j = 0; j < k - 1; j++) {
      rd -= proba[j];
      if (rd < 0) break;
    }
    fvec_cpy(centroids + (size_t)c * d, centroids + (size_t)j * d, d);
    double s = fvec_norm(centroids + (size_t)j * d, d, 2) * 0.0000001;
    fvec_randn_r(eps, d, rand_r(&seed));
    fvec_mul_by(eps, d, s);
    fvec_add(centroids + (size_t)j * d, eps, d);
    fvec_sub(centroids + (size_t)c * d, eps, d);
    proba[j] = 0;
    fvec_normalize(proba, k, 1);
  }
  free(proba);
  free(eps);
  return moved;
}

typedef struct {
  int d, n, k;
  const float *v;   /* device */
  float *cent;      /* device [k][d] */
  float *sums;      /* device [k][d] */
  int *assign;      /* device [n] */
  float *dis;       /* device [n] */
  int *nassign;     /* device [k] */
  double *qerr;     /* device */
  const yb_kmeans_comm_t *comm;
  yb_stream_t s;
  long n_total;
  int exact_order;
} km_dev;

/* yael/kmeans.c:213-329 */
static int kmeans_core_dev(km_dev *K, int niter, int flags, int verbose, float *cent_host,
                           int *nassign_host, unsigned int seed, double *qerr_out,
                           long *iter_tot) {
  const int d = K->d, n = K->n, k = K->k;
  double qerr = HUGE_VAL, qerr_old;
  int tot_moved = 0, iter;
  /* YAEL_B200_KM_TRACE=1: host-side timeline of every iteration on stderr (ms since the iteration
   * started: calls returned = work QUEUED, not finished, until the sync) */
  const int trace = getenv("YAEL_B200_KM_TRACE") != NULL && (!K->comm || K->comm->rank == 0);
  for (iter = 1; iter <= niter; iter++) {
    double t0 = trace ? getmillisecs() : 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    (*iter_tot)++;
    /* assignment: knn_full_thread(2, n, k, d, 1, centroids, v, ...) (kmeans.c:242-244) */
    YBH_CHECK(yb_knn_l2(n, k, d, 1, K->cent, K->v, NULL, K->assign, K->dis, 0, K->s));
    if (trace) t1 = getmillisecs();
    /* histogram + sums + qerr (kmeans.c:249-251, 278-283, 310) */
    YBH_CHECK(yb_kmeans_accumulate(d, n, k, K->v, K->assign, K->dis, K->sums, K->nassign, K->qerr,
                                   K->exact_order, K->s));
    if (trace) t2 = getmillisecs();
    if (K->comm && K->comm->allreduce_sums) {
      int rc = K->comm->allreduce_sums(K->comm->ctx, K->sums, (long)k * d, K->nassign, k, K->qerr,
                                       K->s);
      if (rc) ybh_die("kmeans: allreduce hook", rc);
    }
    if (trace) t3 = getmillisecs();
    /* normalise by the counts (kmeans.c:286-288) and optionally to unit norm (291-293) */
    YBH_CHECK(yb_kmeans_scale(d, k, K->sums, K->nassign, K->cent,
                              (flags & KMEANS_NORMALIZE_CENTS) ? 1 : 0, K->s));
    double q_new = 0;
    YBH_CHECK(yb_d2h(nassign_host, K->nassign, sizeof(int) * (size_t)k, K->s));
    YBH_CHECK(yb_d2h(&q_new, K->qerr, sizeof(double), K->s));
    YBH_CHECK(yb_sync(K->s));
    if (trace) {
      t4 = getmillisecs();
      fprintf(stderr, "kmeans trace iter %d: assign queued %.3f, accumulate queued %.3f, all-reduce queued "
                      "%.3f, synced %.3f ms\n", iter, t1 - t0, t2 - t0, t3 - t0, t4 - t0);
    }
    long tot = 0;
    int empties = 0, c;
    for (c = 0; c < k; c++) {
      tot += nassign_host[c];
      empties += nassign_host[c] == 0;
    }
    if (tot != K->n_total) {
      /* the reference asserts here (kmeans.c:281) */
      fprintf(stderr, "yael_b200: kmeans: %ld of %ld points could not be assigned. "
                      "Something wrong in input. Maybe there are NaNs?\n",
              K->n_total - tot, K->n_total);
      abort();
    }
    /* manage empty clusters; the seed draw happens every iteration (kmeans.c:296) */
    unsigned int split_seed = rand_r(&seed);
    int moved = 0;
    if (empties) {
      YBH_CHECK(yb_d2h(cent_host, K->cent, sizeof(float) * (size_t)k * d, K->s));
      YBH_CHECK(yb_sync(K->s));
      moved = reassign_empty(d, k, cent_host, nassign_host, split_seed);
      YBH_CHECK(yb_h2d(K->cent, cent_host, sizeof(float) * (size_t)k * d, K->s));
    }
    if (moved > 0 && verbose)
      fprintf(stderr, "# kmeans warning: %d empty clusters -> split\n", moved);
    tot_moved += moved;
    if (tot_moved > K->n_total / 100 && tot_moved > 1000) { /* kmeans.c:302-306 */
      fprintf(stderr, "# kmeans: reassigned %d times, abandoning\n", tot_moved);
      return -1;
    }
    qerr_old = qerr;
    qerr = q_new;
    if (qerr_old == qerr && moved == 0) break; /* kmeans.c:312-313 */
    if (verbose) {
      printf(" -> %.3f", qerr / K->n_total);
      fflush(stdout);
    }
  }
  if (verbose) printf("\n");
  *qerr_out = qerr;
  return 0;
}

/* yael/kmeans.c:27-82; distances on the device (compute_distances_1), draws on the host */
static void kmeanspp_init_dev(int d, int n, int k, const float *v_dev, int *sel, int verbose,
                              unsigned int seed, yb_stream_t s) {
  long i, j;
  float *best = fvec_new_set(n, HUGE_VAL);
  float *tmp = fvec_new(n);
  float *tmp_dev = (float *)yb_malloc(sizeof(float) * (size_t)n);
  sel[0] = rand_r(&seed) % k;
  for (i = 1; i < k; i++) {
    int cur = sel[i - 1];
    if (verbose && i % 10 == 0) {
      printf("%d/%d\r", (int)i, k);
      fflush(stdout);
    }
    YBH_CHECK(yb_distances_1(d, n, v_dev + (size_t)d * cur, v_dev, d, tmp_dev, s));
    YBH_CHECK(yb_d2h(tmp, tmp_dev, sizeof(float) * (size_t)n, s));
    YBH_CHECK(yb_sync(s));
    for (j = 0; j < n; j++)
      if (tmp[j] < best[j]) best[j] = tmp[j];
    memcpy(tmp, best, n * sizeof(*tmp));
    fvec_normalize(tmp, n, 1);
    double rd = drand_r(&seed);
    for (j = 0; j < n - 1; j++) {
      rd -= tmp[j];
      if (rd < 0) break;
    }
    sel[i] = (int)j;
  }
  if (verbose) printf("\n");
  free(best);
  free(tmp);
  yb_free(tmp_dev);
}

float yb_kmeans_dev(int d, int n, int k, int niter, const float *v_dev, int flags, long seed_in,
                    int redo, float *centroids_out, float *dis_out, int *assign_out,
                    int *nassign_out, const yb_kmeans_comm_t *comm, yb_stream_t s) {
  long run, iter_tot = 0;
  int verbose = !(flags & KMEANS_QUIET) && !(comm && comm->rank != 0);
  if (flags & (KMEANS_L1 | KMEANS_CHI2)) {
    fprintf(stderr, "yael_b200: kmeans: KMEANS_L1 / KMEANS_CHI2 are outside the B200 hot path "
                    "(medians / Newton per coordinate, no contraction) and are not provided\n");
    return -1;
  }
  niter = (niter == 0 ? 1000000 : niter); /* kmeans.c:344 */
  int is_user_init = (flags & KMEANS_INIT_USER) ? 1 : 0;
  if (is_user_init) {
    assert(centroids_out != NULL);
    redo = 1;
  }
  const float *v_all = comm ? comm->v_host_all : NULL; /* all points, host (sharded run) */
  if (comm && !is_user_init && !v_all) {
    fprintf(stderr, "yael_b200: sharded kmeans needs KMEANS_INIT_USER (the caller gathers the "
                    "initial centroids across ranks) or yb_kmeans_comm_t.v_host_all\n");
    abort();
  }
  long n_total = (comm && comm->n_total > 0) ? comm->n_total : n;
  assert(k <= n_total || !"better to have fewer clusters than points"); /* kmeans.c:377 */

  km_dev K;
  K.d = d; K.n = n; K.k = k; K.v = v_dev; K.comm = comm; K.s = s; K.n_total = n_total;
  K.cent = (float *)yb_malloc(sizeof(float) * (size_t)k * d);
  /* room behind the sums for the packed counts / qerr of a one-collective all-reduce hook */
  K.sums = (float *)yb_malloc(sizeof(float) * ((size_t)k * d + 2 * (size_t)k + 4));
  K.assign = (int *)yb_malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  K.dis = (float *)yb_malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  K.nassign = (int *)yb_malloc(sizeof(int) * (size_t)k);
  K.qerr = (double *)yb_malloc(sizeof(double));
  {
    const char *e = getenv("YAEL_B200_EXACT_UPDATE");
    K.exact_order = e ? atoi(e) : 0;
  }
  float *cent_host = fvec_new((long)k * d);
  int *nassign_host = ivec_new(k);
  int *selected = ivec_new(k);
  double qerr = HUGE_VAL, qerr_best = HUGE_VAL;

  if (seed_in == 0) seed_in = lrand48(); /* kmeans.c:379-380 */
  unsigned int seed = (unsigned int)seed_in;
  int core_ret = 0;

  for (run = 0; run < redo; run++) {
    if (verbose) printf("<><><><> kmeans / run %d <><><><><>\n", (int)run);
    if (is_user_init) {
      YBH_CHECK(yb_h2d(K.cent, centroids_out, sizeof(float) * (size_t)k * d, s));
    } else {
      /* a sharded run replays the selection on every rank from the host copy of ALL points */
      const int n_init = (int)n_total;
      if (flags & KMEANS_INIT_BERKELEY) {
        int nsubset = n_init;
        if (n_init > k * 8 && n_init > 8192) {
          nsubset = k * 8;
          if (verbose) printf("Restricting k-means++ initialization to %d points\n", nsubset);
        }
        if (v_all) {
          float *sub = (float *)yb_malloc(sizeof(float) * (size_t)nsubset * d);
          YBH_CHECK(yb_h2d(sub, v_all, sizeof(float) * (size_t)nsubset * d, s));
          kmeanspp_init_dev(d, nsubset, k, sub, selected, verbose, rand_r(&seed), s);
          yb_free(sub);
        } else {
          kmeanspp_init_dev(d, nsubset, k, v_dev, selected, verbose, rand_r(&seed), s);
        }
      } else {
        /* random_init (kmeans.c:15-20): first k of a seeded Fisher-Yates permutation */
        int *perm = ivec_new_random_perm_r(n_init, rand_r(&seed));
        ivec_cpy(selected, perm, k);
        free(perm);
      }
      if (v_all) {
        for (long c = 0; c < k; c++)
          memcpy(cent_host + (size_t)c * d, v_all + (size_t)selected[c] * d, sizeof(float) * (size_t)d);
        YBH_CHECK(yb_h2d(K.cent, cent_host, sizeof(float) * (size_t)k * d, s));
        YBH_CHECK(yb_sync(s));
      } else {
        int *sel_dev = (int *)yb_malloc(sizeof(int) * (size_t)k);
        YBH_CHECK(yb_h2d(sel_dev, selected, sizeof(int) * (size_t)k, s));
        YBH_CHECK(yb_gather_rows(v_dev, sel_dev, k, d, K.cent, s));
        YBH_CHECK(yb_sync(s));
        yb_free(sel_dev);
      }
    }
    core_ret = kmeans_core_dev(&K, niter, flags, verbose, cent_host, nassign_host, rand_r(&seed),
                               &qerr, &iter_tot);
    if (core_ret < 0) break;
    if (qerr < qerr_best) { /* kmeans.c:417-428 */
      qerr_best = qerr;
      if (centroids_out)
        YBH_CHECK(yb_d2h(centroids_out, K.cent, sizeof(float) * (size_t)k * d, s));
      if (dis_out && n > 0) YBH_CHECK(yb_d2h(dis_out, K.dis, sizeof(float) * (size_t)n, s));
      if (assign_out && n > 0) YBH_CHECK(yb_d2h(assign_out, K.assign, sizeof(int) * (size_t)n, s));
      if (nassign_out) memcpy(nassign_out, nassign_host, sizeof(int) * (size_t)k);
      YBH_CHECK(yb_sync(s));
    }
  }
  if (verbose && core_ret >= 0) { /* kmeans.c:431-434 */
    printf("Total number of iterations: %d\n", (int)iter_tot);
    printf("Unbalanced factor of last iteration: %g\n", ivec_unbalanced_factor(nassign_host, k));
  }
  YBH_CHECK(yb_sync(s));
  free(selected);
  free(cent_host);
  free(nassign_host);
  yb_free(K.cent); yb_free(K.sums); yb_free(K.assign); yb_free(K.dis); yb_free(K.nassign);
  yb_free(K.qerr);
  if (core_ret < 0) return -1;
  return (float)(qerr_best / n_total);
}

/* yael/kmeans.c:332-447 */
float kmeans(int d, int n, int k, int niter, const float *v, int flags, long seed, int redo,
             float *centroids, float *dis, int *assign, int *nassign) {
  /* large problems on host buffers: points sharded over the box's GPUs (yb_mgpu.cu) */
  if (!ybh_is_device_ptr(v) && !(centroids && ybh_is_device_ptr(centroids)) &&
      !(dis && ybh_is_device_ptr(dis)) && !(assign && ybh_is_device_ptr(assign)) &&
      !(nassign && ybh_is_device_ptr(nassign))) {
    float q = 0;
    int rc = yb_mgpu_kmeans(d, n, k, niter, v, flags, seed, redo, centroids, dis, assign, nassign, &q);
    if (rc == 0) return q;
    if (rc > 0) ybh_die("yb_mgpu_kmeans", rc);
  }
  ybh_arg av = ybh_in(v, sizeof(float) * (size_t)n * d);
  /* outputs may be device pointers too: stage through host blocks in that case */
  float *c_h = centroids, *d_h = dis;
  int *a_h = assign, *n_h = nassign;
  int c_dev = centroids && ybh_is_device_ptr(centroids), d_dev = dis && ybh_is_device_ptr(dis);
  int a_dev = assign && ybh_is_device_ptr(assign), n_dev = nassign && ybh_is_device_ptr(nassign);
  if (c_dev) {
    c_h = fvec_new((long)k * d);
    if (flags & KMEANS_INIT_USER) {
      YBH_CHECK(yb_d2h(c_h, centroids, sizeof(float) * (size_t)k * d, NULL));
      YBH_CHECK(yb_sync(NULL));
    }
  }
  if (d_dev) d_h = fvec_new(n);
  if (a_dev) a_h = ivec_new(n);
  if (n_dev) n_h = ivec_new(k);
  float ret = yb_kmeans_dev(d, n, k, niter, (const float *)av.dev, flags, seed, redo, c_h, d_h,
                            a_h, n_h, NULL, NULL);
  if (c_dev) { YBH_CHECK(yb_h2d(centroids, c_h, sizeof(float) * (size_t)k * d, NULL)); }
  if (d_dev) { YBH_CHECK(yb_h2d(dis, d_h, sizeof(float) * (size_t)n, NULL)); }
  if (a_dev) { YBH_CHECK(yb_h2d(assign, a_h, sizeof(int) * (size_t)n, NULL)); }
  if (n_dev) { YBH_CHECK(yb_h2d(nassign, n_h, sizeof(int) * (size_t)k, NULL)); }
  ybh_sync();
  if (c_dev) free(c_h);
  if (d_dev) free(d_h);
  if (a_dev) free(a_h);
  if (n_dev) free(n_h);
  ybh_finish(&av, 0);
  return ret;
}

/* yael/kmeans.c:452-482 */
float *clustering_kmeans_assign_with_score(int n, int di, const float *points, int k,
                                           int nb_iter_max, double normalize, int n_thread,
                                           double *score, int **clust_assign_out) {
  (void)normalize;
  (void)score;
  long d = di;
  float *centroids = fvec_new(k * d);
  int *ca = clust_assign_out ? ivec_new(n) : NULL;
  int nredo = 1;
  if (nb_iter_max / 100000 != 0) {
    nredo = nb_iter_max / 100000;
    nb_iter_max = nb_iter_max % 100000;
  }
  float ret = kmeans(di, n, k, nb_iter_max, points, n_thread | KMEANS_INIT_RANDOM, 0, nredo,
                     centroids, NULL, ca, NULL);
  if (ret >= 0) {
    if (clust_assign_out) *clust_assign_out = ca;
    return centroids;
  }
  free(centroids);
  free(ca);
  if (clust_assign_out) *clust_assign_out = NULL;
  return NULL;
}

/* yael/kmeans.c:484-490 */
float *clustering_kmeans_assign(int n, int d, const float *points, int k, int nb_iter_max,
                                double normalize, int **clust_assign_out) {
  return clustering_kmeans_assign_with_score(n, d, points, k, nb_iter_max, normalize,
                                             count_cpu(), NULL, clust_assign_out);
}

/* yael/kmeans.c:492-504 */
float *clustering_kmeans(int n, int d, const float *points, int k, int nb_iter_max,
                         double normalize) {
  int *clust_assign = NULL;
  float *centroids = clustering_kmeans_assign_with_score(n, d, points, k, nb_iter_max, normalize,
                                                         count_cpu(), NULL, &clust_assign);
  free(clust_assign);
  return centroids;
}
