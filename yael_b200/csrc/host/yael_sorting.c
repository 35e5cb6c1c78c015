/* yael_sorting.c -- the k-smallest family of include/yael/sorting.h on the device
 * (yael/sorting.c:153-255), plus two host helpers callers pair with it. */
#include <assert.h>
#include <stdlib.h>

#include "../../../include/yael/sorting.h"
#include "yb_host.h"

static void k_select(const float *val, long m, long n, int *idx, int k, int sign) {
  if (m == 0 || k == 0 || n == 0) return; /* yael/sorting.c:243-244 */
  ybh_arg av = ybh_in(val, sizeof(float) * (size_t)m * n);
  ybh_arg oi = ybh_out(idx, sizeof(int) * (size_t)k * n);
  YBH_CHECK(yb_k_min_rows((const float *)av.dev, m, m, n, k, sign, (int *)oi.dev, NULL, NULL));
  ybh_finish(&oi, 1);
  ybh_finish(&av, 0);
  ybh_sync();
}

void fvec_k_min(const float *v, int n, int *mins, int k) { /* yael/sorting.c:239-255 */
  assert(k <= n);
  k_select(v, n, 1, mins, k, +1);
}
void fvec_k_max(const float *v, int n, int *maxes, int k) { /* yael/sorting.c:153-170 */
  assert(k <= n);
  k_select(v, n, 1, maxes, k, -1);
}
void fvecs_k_min(const float *val, long m, long n, int *idx, int k) { /* sorting.c:191-196 */
  assert(k <= m);
  k_select(val, m, n, idx, k, +1);
}
void fvecs_k_max(const float *val, long m, long n, int *idx, int k) { /* sorting.c:184-189 */
  assert(k <= m);
  k_select(val, m, n, idx, k, -1);
}

/* ---- host helpers ---- */
static __thread const float *sort_tab;
static int by_value_then_index(const void *a, const void *b) { /* sorting.c:282-299 */
  int ia = *(const int *)a, ib = *(const int *)b;
  float dt = sort_tab[ia] - sort_tab[ib];
  if (dt) return dt > 0 ? 1 : -1;
  return ia - ib;
}
void fvec_sort_index(const float *tab, int n, int *perm) { /* sorting.c:304-316 */
  for (int i = 0; i < n; i++) perm[i] = i;
  sort_tab = tab;
  qsort(perm, (size_t)n, sizeof(int), by_value_then_index);
}
int fvec_arg_min(const float *f, long n) { /* sorting.c:778-789 */
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
