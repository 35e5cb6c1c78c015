/* gmm.h -- the GMM E-step: posteriors of a diagonal-covariance mixture.
 * Structure, flags and prototypes as in the reference's yael/gmm.h:20-33,63-66,121-131.  The
 * learning loop (gmm_learn) and the Fisher-vector code are consumers outside the path and are not
 * provided; a mixture is built by the caller or loaded with gmm_read. */
#ifndef YAEL_B200_GMM_H
#define YAEL_B200_GMM_H

#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* yael/gmm.h:20-26 (host memory) */
typedef struct gmm_s {
  int d;         /* vector dimension */
  int k;         /* number of mixture components */
  float *w;      /* weights [k] */
  float *mu;     /* means [k][d] */
  float *sigma;  /* diagonals of the covariance matrices [k][d] */
} gmm_t;

/* yael/gmm.h:29-33: take the weights into account / leave the log-likelihoods unnormalised (the
 * reference declares the second flag and never reads it; neither does this library) */
#define GMM_FLAGS_W 1
#define GMM_FLAGS_NO_NORM 2

/* yael/gmm.c:305-367: p[n][k] = p(c_j | v_i).  The squared Mahalanobis distances are the
 * reference's two contractions (gmm.c:211-258) on the device, the log-domain combination and the
 * max-shifted softmax (gmm.c:262-300) follow in the reference's arithmetic.  v and p may be host or
 * device pointers. */
void gmm_compute_p(int n, const float *v, const gmm_t *g, float *p, int flags);
/* yael/gmm.c:862-869: the same (n_thread sliced the points over CPU threads; accepted, unused) */
void gmm_compute_p_thread(int n, const float *v, const gmm_t *g, float *p, int flags, int n_thread);
/* yael/gmm.c:43-49 */
void gmm_delete(gmm_t *g);
/* yael/gmm.c:810-836: d, k (int32), then w[k], mu[k][d], sigma[k][d] (float32) */
void gmm_write(const gmm_t *g, FILE *f);
gmm_t *gmm_read(FILE *f);

#ifdef __cplusplus
}
#endif
#endif
