/* yael_gmm.c -- include/yael/gmm.h (yael/gmm.c:30-49, 211-367, 810-869).  The O(k d) tables the
 * reference prepares before its two sgemm calls are prepared here the same way, on the host, with
 * the same types (double sums and logs rounded to float where the source stores floats); the
 * O(n k d) contraction, the log-domain combination and the softmax run on the device
 * (yb_gmm_posteriors). */
#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../../include/yael/gmm.h"
#include "../../../include/yael/vector.h"
#include "yb_host.h"

void gmm_compute_p(int n, const float *v, const gmm_t *g, float *p, int flags) {
  const long d = g->d, k = g->k;
  float *tab, *logdetnr, *mu2, *lg, *inv_sigma, *mu_sigma;
  long i, j, l, chunk, i0;
  ybh_arg at;
  if (n <= 0) return; /* gmm.c:310 */
  tab = fvec_new(3 * k + 2 * k * d);
  logdetnr = tab;
  mu2 = tab + k;
  lg = tab + 2 * k;
  inv_sigma = tab + 3 * k;
  mu_sigma = inv_sigma + k * d;
  for (j = 0; j < k; j++) {
    double dtmp = 0;
    logdetnr[j] = -d / 2.0 * log(2 * M_PI);                                   /* gmm.c:320 */
    for (i = 0; i < d; i++) logdetnr[j] -= 0.5 * log(g->sigma[j * d + i]);     /* gmm.c:321-322 */
    for (l = 0; l < d; l++) {                                                  /* gmm.c:221-226 */
      const double m = g->mu[j * d + l];
      dtmp += m * m / g->sigma[j * d + l];
    }
    mu2[j] = dtmp;
    lg[j] = (flags & GMM_FLAGS_W) ? log(g->w[j]) : 0.f;                        /* gmm.c:346-351 */
  }
  for (i = 0; i < k * d; i++) {
    inv_sigma[i] = 1.0 / g->sigma[i];                                          /* gmm.c:239-240 */
    mu_sigma[i] = g->mu[i] / g->sigma[i];                                      /* gmm.c:249-250 */
  }
  at = ybh_in(tab, sizeof(float) * (size_t)(3 * k + 2 * k * d));
  /* points in slabs of at most 2^28 posteriors (1 GB) when they have to be staged */
  chunk = ((long)1 << 28) / (k > 0 ? k : 1);
  if (chunk < 1024) chunk = 1024;
  if (ybh_is_device_ptr(p) && ybh_is_device_ptr(v)) chunk = n;
  for (i0 = 0; i0 < n; i0 += chunk) {
    const long m = n - i0 < chunk ? n - i0 : chunk;
    const float *td = (const float *)at.dev;
    ybh_arg av = ybh_in(v + (size_t)i0 * d, sizeof(float) * (size_t)m * d);
    ybh_arg ap = ybh_out(p + (size_t)i0 * k, sizeof(float) * (size_t)m * k);
    YBH_CHECK(yb_gmm_posteriors(m, (int)k, (int)d, (const float *)av.dev, td + 3 * k, td + 3 * k + k * d,
                                td + k, td, td + 2 * k, (float *)ap.dev, NULL, NULL));
    ybh_finish(&ap, 1);
    ybh_finish(&av, 0);
  }
  ybh_sync();
  ybh_finish(&at, 0);
  free(tab);
}

void gmm_compute_p_thread(int n, const float *v, const gmm_t *g, float *p, int flags, int n_thread) {
  (void)n_thread;
  gmm_compute_p(n, v, g, p, flags);
}

void gmm_delete(gmm_t *g) {
  free(g->w);
  free(g->mu);
  free(g->sigma);
  free(g);
}

void gmm_write(const gmm_t *g, FILE *f) {
  const size_t k = (size_t)g->k, kd = (size_t)g->k * g->d;
  if (fwrite(&g->d, sizeof(int), 1, f) != 1 || fwrite(&g->k, sizeof(int), 1, f) != 1 ||
      fwrite(g->w, sizeof(float), k, f) != k || fwrite(g->mu, sizeof(float), kd, f) != kd ||
      fwrite(g->sigma, sizeof(float), kd, f) != kd) {
    perror("gmm_write");
    abort();
  }
}

gmm_t *gmm_read(FILE *f) {
  int d, k;
  gmm_t *g;
  size_t kd;
  if (fread(&d, sizeof(int), 1, f) != 1 || fread(&k, sizeof(int), 1, f) != 1) {
    perror("gmm_read");
    abort();
  }
  g = (gmm_t *)malloc(sizeof(*g));
  assert(g);
  g->d = d;
  g->k = k;
  kd = (size_t)k * d;
  g->w = fvec_new(k);
  g->mu = fvec_new((long)kd);
  g->sigma = fvec_new((long)kd);
  if (fread(g->w, sizeof(float), (size_t)k, f) != (size_t)k || fread(g->mu, sizeof(float), kd, f) != kd ||
      fread(g->sigma, sizeof(float), kd, f) != kd) {
    perror("gmm_read");
    abort();
  }
  return g;
}
