/*
 * yael_oracle.c -- CPU restatement of the reference hot path.  See yael_oracle.h.
 *
 * TEST INFRASTRUCTURE ONLY (checker + timed CPU baseline); never linked into
 * the product library.  Written from the reference's behaviour, not its text;
 * each function names the reference lines it follows (paths relative to
 * /root/reference).
 */
#define _GNU_SOURCE
#include "yael_oracle.h"

#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_BLOCK 256 /* yael/nn.c:371-372 */

static void *xmalloc(size_t n) {
  void *p = malloc(n ? n : 1);
  if (!p) {
    fprintf(stderr, "oracle: out of memory (%zu bytes)\n", n);
    abort();
  }
  return p;
}

/* ------------------------------------------------------------------ */
/* BLAS stand-in: -2 * <x,y> with a defined accumulation order          */
/* ------------------------------------------------------------------ */

static inline float dot_f32_seq(const float *x, const float *y, int d) {
  float acc = 0.0f;
  for (int t = 0; t < d; t++) acc = fmaf(x[t], y[t], acc);
  return acc;
}

static inline float dot_f64(const float *x, const float *y, int d) {
  double acc = 0.0;
  for (int t = 0; t < d; t++) acc += (double)x[t] * (double)y[t];
  return (float)acc;
}

static inline float orc_dot(const float *x, const float *y, int d, int mode) {
  return mode == ORC_DOT_F64 ? dot_f64(x, y, d) : dot_f32_seq(x, y, d);
}

/* ------------------------------------------------------------------ */
/* distances                                                           */
/* ------------------------------------------------------------------ */

/* yael/nn.c:100-129.  a-side squared norm: float accumulator (nn.c:108-114);
 * b-side: double accumulator of float products (nn.c:116-120; `dl[j]*dl[j]` is a
 * float*float product promoted to double for the add); the sum is rounded to float
 * once when stored (nn.c:123); then sgemm with alpha=-2, beta=1 adds -2<a_i,b_j>
 * onto it (nn.c:54-67,126). */
void orc_cross_distances_nonpacked(int d, int na, int nb, const float *a, int lda,
                                   const float *b, int ldb, float *dist2, int ldd,
                                   int dot_mode) {
  float *an = (float *)xmalloc(sizeof(float) * (size_t)na);
  for (long i = 0; i < na; i++) {
    const float *row = a + (size_t)lda * i;
    float s = 0;
    for (int t = 0; t < d; t++) s += row[t] * row[t];
    an[i] = s;
  }
  for (long j = 0; j < nb; j++) {
    const float *row = b + (size_t)ldb * j;
    double bn = 0;
    for (int t = 0; t < d; t++) {
      float p = row[t] * row[t];
      bn += p;
    }
    float *out = dist2 + (size_t)ldd * j;
    for (long i = 0; i < na; i++) {
      float base = (float)(bn + an[i]);
      float dp = orc_dot(a + (size_t)lda * i, row, d, dot_mode);
      out[i] = base + (-2.0f) * dp; /* alpha*dot + beta*C */
    }
  }
  free(an);
}

void orc_cross_distances(int d, int na, int nb, const float *a, const float *b,
                         float *dist2, int dot_mode) {
  /* yael/nn.c:92-97 */
  orc_cross_distances_nonpacked(d, na, nb, a, d, b, d, dist2, na, dot_mode);
}

/* yael/nn.c:132-154: one-vs-many; BOTH norms accumulated in double, then sgemv. */
void orc_distances_1(int d, int nb, const float *a, const float *b, int ldb, float *dist2,
                     int dot_mode) {
  double an = 0;
  for (int t = 0; t < d; t++) {
    float p = a[t] * a[t];
    an += p;
  }
  for (long j = 0; j < nb; j++) {
    const float *row = b + (size_t)ldb * j;
    double bn = 0;
    for (int t = 0; t < d; t++) {
      float p = row[t] * row[t];
      bn += p;
    }
    float base = (float)(bn + an);
    dist2[j] = base + (-2.0f) * orc_dot(a, row, d, dot_mode);
  }
}

/* ------------------------------------------------------------------ */
/* sort helper: yael/sorting.c:286-316 -- permutation ordering by        */
/* (value, position).  The reference comparator tests `tab[i]-tab[j]`    */
/* for non-zero, so two NaNs or a NaN and a number compare through the   */
/* position only when the difference is exactly zero; we reproduce that  */
/* comparator literally because fbinheap_sort depends on it.             */
/* ------------------------------------------------------------------ */

static __thread const float *g_sort_tab;

static int cmp_val_then_pos(const void *pa, const void *pb) {
  int ia = *(const int *)pa, ib = *(const int *)pb;
  float diff = g_sort_tab[ia] - g_sort_tab[ib];
  if (diff) return diff > 0 ? 1 : -1;
  return ia - ib;
}

void orc_fvec_sort_index(const float *tab, int n, int *perm) {
  for (int i = 0; i < n; i++) perm[i] = i;
  g_sort_tab = tab;
  qsort(perm, (size_t)n, sizeof(int), cmp_val_then_pos);
}

/* ------------------------------------------------------------------ */
/* max-heap selector                                                   */
/* ------------------------------------------------------------------ */

orc_heap_t *orc_heap_new(int maxk) {
  /* yael/binheap.c:11-35: nodes are addressed from 1 */
  orc_heap_t *h = (orc_heap_t *)xmalloc(sizeof(*h));
  h->k = 0;
  h->maxk = maxk;
  h->val = (float *)xmalloc(sizeof(float) * ((size_t)maxk + 1));
  h->label = (int *)xmalloc(sizeof(int) * ((size_t)maxk + 1));
  return h;
}

void orc_heap_free(orc_heap_t *h) {
  if (!h) return;
  free(h->val);
  free(h->label);
  free(h);
}

/* yael/binheap.c:85-103: append at the bottom, bubble up while the parent is
 * strictly smaller (a parent that is >= stops the climb). */
static void heap_push(orc_heap_t *h, int label, float val) {
  assert(h->k < h->maxk);
  int pos = ++h->k;
  while (pos > 1) {
    int up = pos >> 1;
    if (h->val[up] >= val) break;
    h->val[pos] = h->val[up];
    h->label[pos] = h->label[up];
    pos = up;
  }
  h->val[pos] = val;
  h->label[pos] = label;
}

/* yael/binheap.c:48-82: remove the root; the last element sinks from the top.
 * At each level the left child is preferred only when it is strictly larger than
 * the right one (or the right one does not exist); sinking stops as soon as the
 * moving value is strictly larger than the chosen child. */
static void heap_pop(orc_heap_t *h) {
  assert(h->k > 0);
  int last = h->k;
  float moving = h->val[last];
  int pos = 1;
  for (;;) {
    int l = pos << 1, r = l + 1;
    if (l > last) break;
    int c = (r == last + 1 || h->val[l] > h->val[r]) ? l : r;
    if (moving > h->val[c]) break;
    h->val[pos] = h->val[c];
    h->label[pos] = h->label[c];
    pos = c;
  }
  h->val[pos] = h->val[last];
  h->label[pos] = h->label[last];
  h->k--;
}

void orc_heap_add(orc_heap_t *h, int label, float val) {
  /* yael/binheap.c:106-117: no NaN test on this entry point */
  if (h->k < h->maxk) {
    heap_push(h, label, val);
    return;
  }
  if (val < h->val[1]) {
    heap_pop(h);
    heap_push(h, label, val);
  }
}

void orc_heap_addn_range(orc_heap_t *h, int n, int label0, const float *v) {
  /* yael/binheap.c:139-156: fill phase skips NaN; steady phase admits strictly
   * below the root, which a NaN never is. */
  int i = 0;
  for (; i < n && h->k < h->maxk; i++)
    if (!isnan(v[i])) heap_push(h, label0 + i, v[i]);
  float root = h->val[1];
  for (; i < n; i++) {
    if (v[i] < root) {
      heap_pop(h);
      heap_push(h, label0 + i, v[i]);
      root = h->val[1];
    }
  }
}

void orc_heap_sorted(const orc_heap_t *h, int *labels, float *vals) {
  /* yael/binheap.c:201-211: order by (value, heap slot) */
  int *perm = (int *)xmalloc(sizeof(int) * (size_t)(h->k + 1));
  orc_fvec_sort_index(h->val + 1, h->k, perm);
  for (int i = 0; i < h->k; i++) {
    int slot = perm[i] + 1;
    labels[i] = h->label[slot];
    if (vals) vals[i] = h->val[slot];
  }
  free(perm);
}

/* ------------------------------------------------------------------ */
/* k-NN                                                                */
/* ------------------------------------------------------------------ */

/* yael/nn.c:383-446: k==1.  Running (argmin,min) initialised to (-1, 1e30),
 * strict '<' so the lowest id wins exact ties. */
static void knn_k1_slice(int nq, int nb, int d, const float *b, const float *q,
                         const float *w, int *assign, float *dis, int dot_mode) {
  int s1 = nq < ORC_BLOCK ? nq : ORC_BLOCK, s2 = nb < ORC_BLOCK ? nb : ORC_BLOCK;
  float *blk = (float *)xmalloc(sizeof(float) * (size_t)s1 * s2);
  for (long q0 = 0; q0 < nq; q0 += s1) {
    int m1 = (int)(nq - q0 < s1 ? nq - q0 : s1);
    for (int j = 0; j < m1; j++) {
      assign[q0 + j] = -1;
      dis[q0 + j] = 1e30f;
    }
    for (long b0 = 0; b0 < nb; b0 += s2) {
      int m2 = (int)(nb - b0 < s2 ? nb - b0 : s2);
      orc_cross_distances(d, m2, m1, b + b0 * d, q + q0 * d, blk, dot_mode);
      if (w)
        for (int j = 0; j < m1; j++)
          for (int i = 0; i < m2; i++) blk[(size_t)j * m2 + i] *= w[b0 + i];
      for (int j = 0; j < m1; j++) {
        const float *line = blk + (size_t)j * m2;
        int best = assign[q0 + j];
        float bestd = dis[q0 + j];
        for (int i = 0; i < m2; i++)
          if (line[i] < bestd) {
            bestd = line[i];
            best = (int)(b0 + i);
          }
        assign[q0 + j] = best;
        dis[q0 + j] = bestd;
      }
    }
  }
  free(blk);
}

/* yael/nn.c:451-525 */
static void knn_slice(int nq, int nb, int d, int k, const float *b, const float *q,
                      const float *w, int *assign, float *dis, int dot_mode) {
  assert(k <= nb);
  if (k == 1) {
    knn_k1_slice(nq, nb, d, b, q, w, assign, dis, dot_mode);
    return;
  }
  int s1 = nq < ORC_BLOCK ? nq : ORC_BLOCK, s2 = nb < ORC_BLOCK ? nb : ORC_BLOCK;
  float *blk = (float *)xmalloc(sizeof(float) * (size_t)s1 * s2);
  orc_heap_t **heaps = (orc_heap_t **)xmalloc(sizeof(*heaps) * (size_t)s1);
  for (int j = 0; j < s1; j++) heaps[j] = orc_heap_new(k);

  for (long q0 = 0; q0 < nq; q0 += s1) {
    int m1 = (int)(nq - q0 < s1 ? nq - q0 : s1);
    for (int j = 0; j < m1; j++) heaps[j]->k = 0;
    for (long b0 = 0; b0 < nb; b0 += s2) {
      int m2 = (int)(nb - b0 < s2 ? nb - b0 : s2);
      orc_cross_distances(d, m2, m1, b + b0 * d, q + q0 * d, blk, dot_mode);
      if (w)
        for (int j = 0; j < m1; j++)
          for (int i = 0; i < m2; i++) blk[(size_t)j * m2 + i] *= w[b0 + i];
      for (int j = 0; j < m1; j++)
        orc_heap_addn_range(heaps[j], m2, (int)b0, blk + (size_t)j * m2);
    }
    for (int j = 0; j < m1; j++) {
      orc_heap_t *h = heaps[j];
      int *ao = assign + (size_t)(q0 + j) * k;
      float *dd = dis + (size_t)(q0 + j) * k;
      orc_heap_sorted(h, ao, dd);
      if (h->k < k) { /* yael/nn.c:515-518: pad with all-ones bytes */
        memset(ao + h->k, 0xff, sizeof(int) * (size_t)(k - h->k));
        memset(dd + h->k, 0xff, sizeof(float) * (size_t)(k - h->k));
      }
    }
  }
  for (int j = 0; j < s1; j++) orc_heap_free(heaps[j]);
  free(heaps);
  free(blk);
}

/* yael/nn.c:665-699: nt contiguous query slices [nq*i/nt, nq*(i+1)/nt). */
void orc_knn_full(int nq, int nb, int d, int k, const float *b, const float *q,
                  const float *b_weights, int *assign, float *dis, int dot_mode,
                  int n_thread) {
  if (n_thread < 1) n_thread = 1;
  if (nq < n_thread || n_thread == 1) {
    knn_slice(nq, nb, d, k, b, q, b_weights, assign, dis, dot_mode);
    return;
  }
#pragma omp parallel for schedule(dynamic) num_threads(n_thread)
  for (int i = 0; i < n_thread; i++) {
    long n0 = (long)nq * i / n_thread, n1 = (long)nq * (i + 1) / n_thread;
    knn_slice((int)(n1 - n0), nb, d, k, b, q + n0 * d, b_weights, assign + n0 * k,
              dis + n0 * k, dot_mode);
  }
}

typedef struct {
  float v;
  int id;
} orc_pair_t;

static int cmp_pair(const void *pa, const void *pb) {
  const orc_pair_t *a = (const orc_pair_t *)pa, *b = (const orc_pair_t *)pb;
  if (a->v < b->v) return -1;
  if (a->v > b->v) return 1;
  return (a->id > b->id) - (a->id < b->id);
}

void orc_knn_canonical(int nq, int nb, int d, int k, const float *b, const float *q,
                       int *assign, float *dis, int dot_mode, int n_thread) {
  if (n_thread < 1) n_thread = 1;
#pragma omp parallel num_threads(n_thread)
  {
    float *row = (float *)xmalloc(sizeof(float) * (size_t)nb);
    orc_pair_t *pairs = (orc_pair_t *)xmalloc(sizeof(orc_pair_t) * (size_t)nb);
#pragma omp for schedule(dynamic, 4)
    for (int j = 0; j < nq; j++) {
      /* distances of query j to every base row, same formula as nn.c:100-129 with the
       * base as the a-operand and the query as the b-operand (nn.c:493) */
      orc_cross_distances(d, nb, 1, b, q + (size_t)j * d, row, dot_mode);
      int m = 0;
      for (int i = 0; i < nb; i++)
        if (!isnan(row[i])) {
          pairs[m].v = row[i];
          pairs[m].id = i;
          m++;
        }
      qsort(pairs, (size_t)m, sizeof(orc_pair_t), cmp_pair);
      for (int r = 0; r < k; r++) {
        if (r < m) {
          assign[(size_t)j * k + r] = pairs[r].id;
          dis[(size_t)j * k + r] = pairs[r].v;
        } else {
          assign[(size_t)j * k + r] = -1;
          memset(&dis[(size_t)j * k + r], 0xff, sizeof(float));
        }
      }
    }
    free(row);
    free(pairs);
  }
}

/* yael/nn.c:528-580 */
void orc_knn_reorder_shortlist(int n, int nb, int d, int k, const float *b, const float *v,
                               int *idx, float *dis, int dot_mode) {
  (void)nb;
  float *rows = (float *)xmalloc(sizeof(float) * (size_t)k * d);
  float *tmpd = (float *)xmalloc(sizeof(float) * (size_t)k);
  int *perm = (int *)xmalloc(sizeof(int) * (size_t)k);
  int *tmpi = (int *)xmalloc(sizeof(int) * (size_t)k);
  for (long i = 0; i < n; i++) {
    int *ids = idx + i * k;
    float *out = dis + i * k;
    int ki = 0;
    while (ki < k && ids[ki] >= 0) {
      memcpy(rows + (size_t)ki * d, b + (size_t)ids[ki] * d, sizeof(float) * (size_t)d);
      ki++;
    }
    orc_distances_1(d, ki, v + i * d, rows, d, tmpd, dot_mode);
    orc_fvec_sort_index(tmpd, ki, perm);
    memcpy(tmpi, ids, sizeof(int) * (size_t)ki);
    for (int j = 0; j < ki; j++) {
      out[j] = tmpd[perm[j]];
      ids[j] = tmpi[perm[j]];
    }
  }
  free(rows);
  free(tmpd);
  free(perm);
  free(tmpi);
}

/* ------------------------------------------------------------------ */
/* k smallest of an array                                              */
/* ------------------------------------------------------------------ */

int orc_fvec_arg_min(const float *f, long n) {
  /* yael/sorting.c:778-789: first index among equal minima */
  assert(n > 0);
  long best = 0;
  float m = f[0];
  for (long i = 1; i < n; i++)
    if (f[i] < m) {
      m = f[i];
      best = i;
    }
  return (int)best;
}

void orc_fvec_k_min_canonical(const float *val, int n, int *idx, int k) {
  orc_pair_t *p = (orc_pair_t *)xmalloc(sizeof(orc_pair_t) * (size_t)n);
  for (int i = 0; i < n; i++) {
    p[i].v = val[i];
    p[i].id = i;
  }
  qsort(p, (size_t)n, sizeof(orc_pair_t), cmp_pair);
  for (int i = 0; i < k; i++) idx[i] = p[i].id;
  free(p);
}

void orc_fvec_k_min(const float *val, int n, int *idx, int k) {
  /* yael/sorting.c:239-255 */
  assert(k <= n);
  if (n == 0 || k == 0) return;
  if (k == 1) {
    idx[0] = orc_fvec_arg_min(val, n);
    return;
  }
  if (n > 20 * k) {
    /* yael/sorting.c:225-236: heap over every element, then labels by (value, slot) */
    orc_heap_t *h = orc_heap_new(k);
    for (int i = 0; i < n; i++) orc_heap_add(h, i, val[i]);
    orc_heap_sorted(h, idx, NULL);
    orc_heap_free(h);
    return;
  }
  /* yael/sorting.c:202-221 (quickselect + qsort with a comparator that never returns 0):
   * the selected VALUES are the k smallest, ascending; which of several equal values is
   * reported is unspecified (sorting.c:174-181), so the restatement fixes (value,index). */
  orc_fvec_k_min_canonical(val, n, idx, k);
}

void orc_fvecs_k_min(const float *val, long m, long n, int *idx, int k) {
  /* yael/sorting.c:191-196: n arrays of length m, serial */
  for (long i = 0; i < n; i++) orc_fvec_k_min(val + m * i, (int)m, idx + (size_t)k * i, k);
}

/* ------------------------------------------------------------------ */
/* RNG                                                                 */
/* ------------------------------------------------------------------ */

double orc_drand_r(unsigned int *seed) {
  /* yael/kmeans.c:22-24, yael/vector.c:135-137 */
  return rand_r(seed) / ((double)RAND_MAX + 1.0);
}

double orc_gaussrand_r(unsigned int *seed) {
  /* yael/vector.c:141-154: ratio-of-uniforms rejection; u1, u2 and the squared
   * quarter are held in FLOAT variables, the returned deviate is double. */
  const double magic = 1.71552776992141;
  for (;;) {
    float u1 = (float)orc_drand_r(seed);
    float u2 = (float)orc_drand_r(seed);
    double z = magic * (u1 - .5) / u2;
    float zz = (float)(z * z / 4.0);
    if (zz < -log(u2)) return z;
  }
}

void orc_fvec_randn_r(float *v, long n, unsigned int seed) {
  /* yael/vector.c:184-189 */
  for (long i = 0; i < n; i++) v[i] = (float)orc_gaussrand_r(&seed);
}

int *orc_random_perm_r(int n, unsigned int seed) {
  /* yael/vector.c:226-253: Fisher-Yates, n-1 swaps, j = i + rand_r % (n-i) */
  int *p = (int *)xmalloc(sizeof(int) * (size_t)n);
  for (int i = 0; i < n; i++) p[i] = i;
  for (int i = 0; i < n - 1; i++) {
    int j = i + rand_r(&seed) % (n - i);
    int t = p[i];
    p[i] = p[j];
    p[j] = t;
  }
  return p;
}

/* ------------------------------------------------------------------ */
/* k-means                                                             */
/* ------------------------------------------------------------------ */

static double vec_norm2(const float *v, long n) {
  /* yael/vector.c:2180-2199 (norm==2): double accumulation of float products */
  double s = 0;
  for (long i = 0; i < n; i++) {
    float p = v[i] * v[i];
    s += p;
  }
  return sqrt(s);
}

static double vec_norm1(const float *v, long n) {
  double s = 0;
  for (long i = 0; i < n; i++) s += fabs(v[i]);
  return s;
}

static void vec_scale(float *v, long n, double f) {
  /* yael/vector.c:1792-1797: float *= double */
  for (long i = 0; i < n; i++) v[i] = (float)(v[i] * f);
}

/* yael/kmeans.c:166-209 */
int orc_kmeans_reassign_empty(int d, int n, int k, float *centroids, int *assign,
                              int *nassign, unsigned int seed) {
  (void)n;
  (void)assign;
  int moved = 0;
  float *p = (float *)xmalloc(sizeof(float) * (size_t)k);
  float *eps = (float *)xmalloc(sizeof(float) * (size_t)d);
  for (int c = 0; c < k; c++)
    p[c] = nassign[c] < 2 ? 0 : (float)(nassign[c] * nassign[c] - 1);
  vec_scale(p, k, 1.0 / vec_norm1(p, k));

  for (int c = 0; c < k; c++) {
    if (nassign[c] != 0) continue;
    moved++;
    double r = orc_drand_r(&seed);
    int j = 0;
    for (; j < k - 1; j++) {
      r -= p[j];
      if (r < 0) break;
    }
    float *cj = centroids + (size_t)j * d, *cc = centroids + (size_t)c * d;
    memcpy(cc, cj, sizeof(float) * (size_t)d);
    double s = vec_norm2(cj, d) * 0.0000001;
    orc_fvec_randn_r(eps, d, rand_r(&seed));
    vec_scale(eps, d, s);
    for (int t = 0; t < d; t++) cj[t] += eps[t];
    for (int t = 0; t < d; t++) cc[t] -= eps[t];
    p[j] = 0;
    vec_scale(p, k, 1.0 / vec_norm1(p, k));
  }
  free(p);
  free(eps);
  return moved;
}

double orc_kmeans_step(int d, int n, int k, const float *v, const float *centroids_in,
                       float *centroids_out, int *assign, float *dis, int *nassign,
                       int dot_mode, int n_thread) {
  /* yael/kmeans.c:242-288, 310 */
  orc_knn_full(n, k, d, 1, centroids_in, v, NULL, assign, dis, dot_mode, n_thread);
  memset(nassign, 0, sizeof(int) * (size_t)k);
  for (long i = 0; i < n; i++) nassign[assign[i]]++;
  memset(centroids_out, 0, sizeof(float) * (size_t)k * d);
  for (long i = 0; i < n; i++) {
    float *c = centroids_out + (size_t)assign[i] * d;
    const float *x = v + (size_t)i * d;
    for (int t = 0; t < d; t++) c[t] += x[t];
  }
  for (int c = 0; c < k; c++) vec_scale(centroids_out + (size_t)c * d, d, 1.0 / nassign[c]);
  double q = 0;
  for (long i = 0; i < n; i++) q += dis[i];
  return q;
}

/* yael/kmeans.c:213-329 */
static int kmeans_core(int d, int n, int k, int niter, int nt, int flags, int verbose,
                       float *centroids, const float *v, unsigned int seed, int *assign,
                       int *nassign, float *dis, double *qerr_out, long *iter_tot,
                       int dot_mode) {
  double qerr = HUGE_VAL, qerr_old;
  int tot_moved = 0;
  float *next = (float *)xmalloc(sizeof(float) * (size_t)k * d);
  for (int iter = 1; iter <= niter; iter++) {
    (*iter_tot)++;
    double q = orc_kmeans_step(d, n, k, v, centroids, next, assign, dis, nassign, dot_mode, nt);
    memcpy(centroids, next, sizeof(float) * (size_t)k * d);
    if (flags & ORC_KMEANS_NORMALIZE_CENTS)
      for (int c = 0; c < k; c++) {
        float *row = centroids + (size_t)c * d;
        vec_scale(row, d, 1.0 / vec_norm2(row, d));
      }
    int moved = orc_kmeans_reassign_empty(d, n, k, centroids, assign, nassign, rand_r(&seed));
    if (moved > 0 && verbose)
      fprintf(stderr, "# kmeans warning: %d empty clusters -> split\n", moved);
    tot_moved += moved;
    if (tot_moved > n / 100 && tot_moved > 1000) {
      fprintf(stderr, "# kmeans: reassigned %d times, abandoning\n", tot_moved);
      free(next);
      return -1;
    }
    qerr_old = qerr;
    qerr = q;
    if (qerr_old == qerr && moved == 0) break;
    if (verbose) {
      printf(" -> %.3f", qerr / n);
      fflush(stdout);
    }
  }
  if (verbose) printf("\n");
  *qerr_out = qerr;
  free(next);
  return 0;
}

/* yael/kmeans.c:27-82 */
static void kmeanspp_init(long d, int n, int k, const float *v, int *sel, int verbose,
                          unsigned int seed, int dot_mode) {
  float *best = (float *)xmalloc(sizeof(float) * (size_t)n);
  float *tmp = (float *)xmalloc(sizeof(float) * (size_t)n);
  for (int j = 0; j < n; j++) best[j] = HUGE_VALF;
  sel[0] = rand_r(&seed) % k;
  for (long i = 1; i < k; i++) {
    int cur = sel[i - 1];
    if (verbose && i % 10 == 0) {
      printf("%d/%d\r", (int)i, k);
      fflush(stdout);
    }
    orc_distances_1((int)d, n, v + d * cur, v, (int)d, tmp, dot_mode);
    for (int j = 0; j < n; j++)
      if (tmp[j] < best[j]) best[j] = tmp[j];
    memcpy(tmp, best, sizeof(float) * (size_t)n);
    vec_scale(tmp, n, 1.0 / vec_norm1(tmp, n));
    double r = orc_drand_r(&seed);
    int j = 0;
    for (; j < n - 1; j++) {
      r -= tmp[j];
      if (r < 0) break;
    }
    sel[i] = j;
  }
  if (verbose) printf("\n");
  free(best);
  free(tmp);
}

/* yael/kmeans.c:332-447 */
float orc_kmeans(int di, int n, int k, int niter, const float *v, int flags, long seed_in,
                 int redo, float *centroids_out, float *dis_out, int *assign_out,
                 int *nassign_out, int dot_mode) {
  long d = di, iter_tot = 0;
  int nt = flags & 0xffff;
  if (nt == 0) nt = 1;
  int verbose = !(flags & ORC_KMEANS_QUIET);
  if (niter == 0) niter = 1000000;
  int user_init = (flags & ORC_KMEANS_INIT_USER) != 0;
  if (user_init) {
    assert(centroids_out != NULL);
    redo = 1;
  }
  float *centroids = (float *)xmalloc(sizeof(float) * (size_t)k * d);
  float *dis = (float *)xmalloc(sizeof(float) * (size_t)n);
  int *assign = (int *)xmalloc(sizeof(int) * (size_t)n);
  int *nassign = (int *)xmalloc(sizeof(int) * (size_t)k);
  int *sel = (int *)xmalloc(sizeof(int) * (size_t)k);
  double qerr = HUGE_VAL, qerr_best = HUGE_VAL;
  assert(k <= n);
  if (seed_in == 0) seed_in = lrand48();
  unsigned int seed = (unsigned int)seed_in;
  int core_ret = 0;

  for (int run = 0; run < redo; run++) {
    if (verbose) printf("<><><><> kmeans / run %d <><><><><>\n", run);
    if (user_init) {
      memcpy(centroids, centroids_out, sizeof(float) * (size_t)k * d);
    } else {
      if (flags & ORC_KMEANS_INIT_BERKELEY) {
        int nsub = n;
        if (n > k * 8 && n > 8192) {
          nsub = k * 8;
          if (verbose) printf("Restricting k-means++ initialization to %d points\n", nsub);
        }
        kmeanspp_init(d, nsub, k, v, sel, verbose, rand_r(&seed), dot_mode);
      } else {
        int *perm = orc_random_perm_r(n, rand_r(&seed)); /* kmeans.c:15-20 */
        memcpy(sel, perm, sizeof(int) * (size_t)k);
        free(perm);
      }
      /* note: the reference indexes with an int product here (kmeans.c:405) */
      for (long i = 0; i < k; i++)
        memcpy(centroids + i * d, v + (size_t)sel[i] * d, sizeof(float) * (size_t)d);
    }
    core_ret = kmeans_core((int)d, n, k, niter, nt, flags, verbose, centroids, v,
                           rand_r(&seed), assign, nassign, dis, &qerr, &iter_tot, dot_mode);
    if (core_ret < 0) break;
    if (qerr < qerr_best) {
      qerr_best = qerr;
      if (centroids_out) memcpy(centroids_out, centroids, sizeof(float) * (size_t)k * d);
      if (dis_out) memcpy(dis_out, dis, sizeof(float) * (size_t)n);
      if (assign_out) memcpy(assign_out, assign, sizeof(int) * (size_t)n);
      if (nassign_out) memcpy(nassign_out, nassign, sizeof(int) * (size_t)k);
    }
  }
  if (verbose && core_ret >= 0) {
    double tot = 0, uf = 0; /* yael/vector.c:2301-2314 */
    for (int c = 0; c < k; c++) {
      tot += nassign[c];
      uf += nassign[c] * (double)nassign[c];
    }
    printf("Total number of iterations: %d\n", (int)iter_tot);
    printf("Unbalanced factor of last iteration: %g\n", uf * k / (tot * tot));
  }
  free(sel);
  free(centroids);
  free(dis);
  free(assign);
  free(nassign);
  return core_ret < 0 ? -1.0f : (float)(qerr_best / n);
}

/* ------------------------------------------------------------------ */
/* Hamming                                                             */
/* ------------------------------------------------------------------ */

uint16_t orc_hamming(const uint8_t *a, const uint8_t *b, int ncodes) {
  /* yael/hamming.c:66-78 (byte LUT) == popcount of the XOR; :20-24 (SSE4.2 popcnt) */
  unsigned h = 0;
  for (int i = 0; i < ncodes; i++) h += (unsigned)__builtin_popcount((unsigned)(a[i] ^ b[i]));
  return (uint16_t)h;
}

static inline unsigned ham_words(const uint8_t *a, const uint8_t *b, int ncodes) {
  unsigned h = 0;
  int i = 0;
  for (; i + 8 <= ncodes; i += 8) {
    uint64_t x, y;
    memcpy(&x, a + i, 8);
    memcpy(&y, b + i, 8);
    h += (unsigned)__builtin_popcountll(x ^ y);
  }
  for (; i < ncodes; i++) h += (unsigned)__builtin_popcount((unsigned)(a[i] ^ b[i]));
  return h;
}

void orc_compute_hamming(uint16_t *dis, const uint8_t *a, const uint8_t *b, int na, int nb,
                         int ncodes) {
  /* yael/hamming.c:177-219: dis[j*na + i] = ham(a_i, b_j) for every size class */
  for (long j = 0; j < nb; j++)
    for (long i = 0; i < na; i++)
      dis[j * na + i] =
          (uint16_t)ham_words(a + (size_t)i * ncodes, b + (size_t)j * ncodes, ncodes);
}

void orc_nn_hamming(int nq, int nb, int ncodes, int k, const uint8_t *b, const uint8_t *q,
                    int *assign, uint16_t *dis, int n_thread) {
  /* composed oracle: compute_hamming row + counting select in id
   * order == stable (distance, id) order. */
  int nbits = ncodes * 8;
  if (n_thread < 1) n_thread = 1;
#pragma omp parallel num_threads(n_thread)
  {
    uint16_t *row = (uint16_t *)xmalloc(sizeof(uint16_t) * (size_t)nb);
    long *hist = (long *)xmalloc(sizeof(long) * (size_t)(nbits + 2));
#pragma omp for schedule(dynamic, 4)
    for (int j = 0; j < nq; j++) {
      const uint8_t *qj = q + (size_t)j * ncodes;
      memset(hist, 0, sizeof(long) * (size_t)(nbits + 2));
      for (long i = 0; i < nb; i++) {
        row[i] = (uint16_t)ham_words(b + (size_t)i * ncodes, qj, ncodes);
        hist[row[i]]++;
      }
      /* offsets of each distance bucket in the output */
      long acc = 0;
      for (int h = 0; h <= nbits; h++) {
        long c = hist[h];
        hist[h] = acc;
        acc += c;
      }
      int *ao = assign + (size_t)j * k;
      uint16_t *dd = dis + (size_t)j * k;
      for (int r = 0; r < k; r++) {
        ao[r] = -1;
        dd[r] = 0xffff;
      }
      for (long i = 0; i < nb; i++) {
        long pos = hist[row[i]]++;
        if (pos < k) {
          ao[pos] = (int)i;
          dd[pos] = row[i];
        }
      }
    }
    free(row);
    free(hist);
  }
}

void orc_match_hamming_count(const uint8_t *bs1, const uint8_t *bs2, int n1, int n2, int ht,
                             int ncodes, size_t *nptr) {
  /* yael/hamming.c:224-300: score <= ht */
  size_t cnt = 0;
  for (long i = 0; i < n1; i++)
    for (long j = 0; j < n2; j++)
      if ((int)ham_words(bs1 + (size_t)i * ncodes, bs2 + (size_t)j * ncodes, ncodes) <= ht)
        cnt++;
  *nptr = cnt;
}

size_t orc_match_hamming_thres_prealloc(const uint8_t *bs1, const uint8_t *bs2, int n1,
                                        int n2, int ht, int ncodes, int *idx,
                                        uint16_t *hams) {
  /* yael/hamming.c:563-700: idx receives (i, j) pairs interleaved */
  size_t cnt = 0;
  for (long i = 0; i < n1; i++)
    for (long j = 0; j < n2; j++) {
      unsigned h = ham_words(bs1 + (size_t)i * ncodes, bs2 + (size_t)j * ncodes, ncodes);
      if ((int)h <= ht) {
        idx[2 * cnt] = (int)i;
        idx[2 * cnt + 1] = (int)j;
        hams[cnt] = (uint16_t)h;
        cnt++;
      }
    }
  return cnt;
}

void orc_crossmatch_hamming_count(const uint8_t *dbs, int n, int ht, int ncodes, size_t *nptr) {
  /* yael/hamming.c:310-395: pairs i < j of one set, score <= ht */
  size_t cnt = 0;
  for (long i = 0; i < n; i++)
    for (long j = i + 1; j < n; j++)
      if ((int)ham_words(dbs + (size_t)i * ncodes, dbs + (size_t)j * ncodes, ncodes) <= ht) cnt++;
  *nptr = cnt;
}

size_t orc_crossmatch_hamming_prealloc(const uint8_t *dbs, long n, int ht, int ncodes, int *idx,
                                       uint16_t *hams) {
  /* yael/hamming.c:793-829: (i, j) interleaved, i outer / j = i+1.. inner */
  size_t cnt = 0;
  for (long i = 0; i < n; i++)
    for (long j = i + 1; j < n; j++) {
      unsigned h = ham_words(dbs + (size_t)i * ncodes, dbs + (size_t)j * ncodes, ncodes);
      if ((int)h <= ht) {
        idx[2 * cnt] = (int)i;
        idx[2 * cnt + 1] = (int)j;
        hams[cnt] = (uint16_t)h;
        cnt++;
      }
    }
  return cnt;
}

/* ---- consumers of the k = 1 search: VLAD / bag of features (yael/vlad.c:10-139) ---- */
/* yael/vlad.c:10-49: assign = nn() (yael/nn.c:608-621 -> knn_full, k = 1), then, for the points in
 * increasing order, desc[assign_i] += fl32(v_i - c) (times w_i: vlad.c:45); floats throughout */
void orc_vlad_compute(int k, int d, const float *centroids, int n, const float *v,
                      const float *weights, float *desc, int dot_mode) {
  int *assign = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  float *dis = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  orc_knn_full(n, k, d, 1, centroids, v, NULL, assign, dis, dot_mode, 1);
  memset(desc, 0, sizeof(float) * (size_t)k * d);
  for (long i = 0; i < n; i++) {
    const float *c = centroids + (size_t)assign[i] * d;
    float *o = desc + (size_t)assign[i] * d;
    for (int j = 0; j < d; j++) {
      float r = v[i * d + j] - c[j];
      if (weights) r = r * weights[i];
      o[j] += r;
    }
  }
  free(assign);
  free(dis);
}

/* yael/vlad.c:52-79: one descriptor per subset, summed in LIST order */
void orc_vlad_compute_subsets(int k, int d, const float *centroids, int n, const float *v,
                              int n_subset, const int *subset_indexes, const int *subset_ends,
                              float *desc, int dot_mode) {
  int *assign = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  float *dis = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  orc_knn_full(n, k, d, 1, centroids, v, NULL, assign, dis, dot_mode, 1);
  memset(desc, 0, sizeof(float) * (size_t)k * d * n_subset);
  int begin = 0;
  for (int ss = 0; ss < n_subset; ss++) {
    float *dss = desc + (size_t)ss * k * d;
    for (int ii = begin; ii < subset_ends[ss]; ii++) {
      const long i = subset_indexes[ii];
      for (int j = 0; j < d; j++)
        dss[(size_t)assign[i] * d + j] += v[i * d + j] - centroids[(size_t)assign[i] * d + j];
    }
    begin = subset_ends[ss];
  }
  free(assign);
  free(dis);
}

/* yael/vlad.c:110-138: histogram of the ma nearest centroids of every point (ma = 1: bof_compute) */
void orc_bof_compute_ma(int k, int d, const float *centroids, int n, const float *v, int *desc, int ma,
                        int dot_mode) {
  int *assign = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1) * ma);
  float *dis = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1) * ma);
  orc_knn_full(n, k, d, ma, centroids, v, NULL, assign, dis, dot_mode, 1);
  memset(desc, 0, sizeof(int) * (size_t)k);
  for (long i = 0; i < (long)n * ma; i++) desc[assign[i]]++;
  free(assign);
  free(dis);
}

/* yael/vlad.c:82-107: one float histogram per subset */
void orc_bof_compute_subsets(int k, int d, const float *centroids, int n, const float *v, int n_subset,
                             const int *subset_indexes, const int *subset_ends, float *desc,
                             int dot_mode) {
  int *assign = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  float *dis = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  orc_knn_full(n, k, d, 1, centroids, v, NULL, assign, dis, dot_mode, 1);
  memset(desc, 0, sizeof(float) * (size_t)k * n_subset);
  int begin = 0;
  for (int ss = 0; ss < n_subset; ss++) {
    for (int ii = begin; ii < subset_ends[ss]; ii++) desc[(size_t)ss * k + assign[subset_indexes[ii]]] += 1.0f;
    begin = subset_ends[ss];
  }
  free(assign);
  free(dis);
}

/* ------------------------------------------------------------------ */
/* hierarchical k-means quantiser (yael/hkm.c:144-162)                  */
/* ------------------------------------------------------------------ */

/* yael/hkm.c:144-162: every point walks the tree; at level l the bf children of its current node
 * are searched with nn() (yael/nn.c:608-621 -> knn_full, k = 1: lowest id on exact ties) and
 * vw = vw * bf + child.  centroids = the levels' tables concatenated: level l holds bf^(l+1) rows. */
void orc_hkm_quantize(int nlevel, int bf, int d, const float *centroids, int n, const float *v,
                      int *idx, int dot_mode) {
  for (long i = 0; i < n; i++) {
    int vw = 0;
    const float *level = centroids;
    long rows = bf;
    for (int l = 0; l < nlevel; l++) {
      int child;
      float dis;
      orc_knn_full(1, bf, d, 1, level + (size_t)vw * d * bf, v + (size_t)d * i, NULL, &child, &dis,
                   dot_mode, 1);
      vw = vw * bf + child;
      level += (size_t)rows * d;
      rows *= bf;
    }
    idx[i] = vw;
  }
}

/* ------------------------------------------------------------------ */
/* GMM E-step (yael/gmm.c:211-258 Mahalanobis, :262-300 softmax,        */
/* :305-367 gmm_compute_p)                                             */
/* ------------------------------------------------------------------ */

#define ORC_GMM_FLAGS_W 1

/* p[i*k + j] = posterior of mixture component j for point i.  The two sgemm calls of
 * compute_mahalanobis_sqr (gmm.c:244,254: C += A'B, then C += -2 A'B) are restated with the
 * selectable dot order; everything around them follows the source's types: double sums for
 * mu^2/sigma rounded to float (gmm.c:221-226), float v^2 (gmm.c:235-236), (float)(1.0/sigma)
 * (gmm.c:239-240), float mu/sigma (gmm.c:249-250), the log-domain combination in double
 * (gmm.c:357), exp in double stored as float and a sequential float sum (gmm.c:279-293). */
void orc_gmm_compute_p(int n, int d, int k, const float *w, const float *mu, const float *sigma,
                       const float *v, float *p, int flags, int dot_mode) {
  if (n == 0) return;
  float *logdetnr = (float *)xmalloc(sizeof(float) * (size_t)k);
  float *mu2 = (float *)xmalloc(sizeof(float) * (size_t)k);
  float *lg = (float *)xmalloc(sizeof(float) * (size_t)k);
  float *is = (float *)xmalloc(sizeof(float) * (size_t)k * d);
  float *ms = (float *)xmalloc(sizeof(float) * (size_t)k * d);
  float *v2 = (float *)xmalloc(sizeof(float) * (size_t)d);
  for (long j = 0; j < k; j++) {
    logdetnr[j] = -(long)d / 2.0 * log(2 * M_PI);
    for (long i = 0; i < d; i++) logdetnr[j] -= 0.5 * log(sigma[j * d + i]);
    double dt = 0;
    for (long l = 0; l < d; l++) {
      double m = mu[j * d + l];
      dt += m * m / sigma[j * d + l];
    }
    mu2[j] = dt;
    lg[j] = (flags & ORC_GMM_FLAGS_W) ? log(w[j]) : 0.f;
  }
  for (long i = 0; i < (long)k * d; i++) {
    is[i] = 1.0 / sigma[i];
    ms[i] = mu[i] / sigma[i];
  }
  const float norm_to_0 = 16.636;
  for (long i = 0; i < n; i++) {
    const float *vi = v + (size_t)i * d;
    float *pi = p + (size_t)i * k;
    for (int l = 0; l < d; l++) v2[l] = vi[l] * vi[l];
    for (long j = 0; j < k; j++) {
      float c = mu2[j];
      c = c + orc_dot(is + j * d, v2, d, dot_mode);
      c = c + (-2.0f) * orc_dot(ms + j * d, vi, d, dot_mode);
      pi[j] = logdetnr[j] - 0.5 * c + lg[j];
    }
    float maxval = -1e30;
    for (long l = 0; l < k; l++)
      if (pi[l] > maxval) maxval = pi[l];
    float s = 0.0;
    for (long l = 0; l < k; l++) {
      if (pi[l] >= maxval - norm_to_0) {
        pi[l] = exp(pi[l] - maxval);
        s += pi[l];
      } else
        pi[l] = 0;
    }
    if (s != 0) {
      float inv = 1.0 / s;
      for (long l = 0; l < k; l++) pi[l] *= inv;
    }
  }
  free(logdetnr); free(mu2); free(lg); free(is); free(ms); free(v2);
}
