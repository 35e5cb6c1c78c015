/* yael_hkm.c -- include/yael/hkm.h (yael/hkm.c).  hkm_learn is the reference's level-by-level
 * loop on the host -- points grouped by node in (node, index) order, one kmeans() call per node
 * with the reference's flags and seed 0 -- around this library's kmeans() (device); hkm_quantize
 * walks every point down the tree on the device (yb_hkm_quantize: one exact k = 1 search among the
 * bf children per level).  Tables live in host memory, as the structure promises. */
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../../include/yael/hkm.h"
#include "../../../include/yael/kmeans.h"
#include "../../../include/yael/machinedeps.h"
#include "../../../include/yael/vector.h"
#include "yb_host.h"

/* yael/hkm.c:18-32 */
static hkm_t *hkm_alloc(int d, int nlevel, int bf) {
  hkm_t *h = (hkm_t *)malloc(sizeof(*h));
  long rows = 1;
  int l;
  assert(h);
  h->nlevel = nlevel;
  h->bf = bf;
  h->d = d;
  h->centroids = (float **)malloc(sizeof(float *) * (size_t)(nlevel > 0 ? nlevel : 1));
  for (l = 0; l < nlevel; l++) {
    rows *= bf;
    h->centroids[l] = fvec_new(rows * d);
  }
  h->k = (int)rows;
  return h;
}

hkm_t *hkm_learn(int n, int d, int nlevel, int bf, const float *points, int nb_iter_max, int nt,
                 int verbose, int **clust_assign_out) {
  hkm_t *h = hkm_alloc(d, nlevel, bf);
  int *node = ivec_new_0(n);           /* node of every point at the current level */
  int *order = ivec_new(n);            /* points by (node, index): ivec_sort_index (hkm.c:54-55) */
  float *grouped = fvec_new((long)n * d);
  long nodes = 1;
  int l;
  (void)nt; /* the reference overrides it with count_cpu() as well (hkm.c:86) */
  for (l = 0; l < nlevel; l++) {
    long *begin = (long *)calloc((size_t)nodes + 1, sizeof(long));
    long parent, i;
    assert(begin);
    for (i = 0; i < n; i++) begin[node[i] + 1]++;
    for (parent = 0; parent < nodes; parent++) begin[parent + 1] += begin[parent];
    {
      long *fill = (long *)malloc(sizeof(long) * (size_t)nodes);
      assert(fill);
      memcpy(fill, begin, sizeof(long) * (size_t)nodes);
      for (i = 0; i < n; i++) order[fill[node[i]]++] = (int)i;
      free(fill);
    }
    for (i = 0; i < n; i++) memcpy(grouped + (size_t)d * i, points + (size_t)d * order[i], sizeof(float) * (size_t)d);
    for (parent = 0; parent < nodes; parent++) {
      const long pos = begin[parent];
      const int cnt = (int)(begin[parent + 1] - pos);
      int *sub;
      float err;
      if (verbose) fprintf(stderr, "[Level %d | Parent %ld] nassign=%d | pos=%ld", l, parent, cnt, pos);
      if (cnt == 0) { /* hkm.c:78-81 */
        fprintf(stderr, "# Problem2: no enough vectors in a node\n");
        exit(1);
      }
      sub = ivec_new(cnt);
      /* hkm.c:86-89: random init, quiet, seed 0 (kmeans draws lrand48), one run */
      err = kmeans(d, cnt, bf, nb_iter_max, grouped + (size_t)d * pos,
                   count_cpu() | KMEANS_INIT_RANDOM | KMEANS_QUIET, 0, 1,
                   h->centroids[l] + (size_t)d * parent * bf, NULL, sub, NULL);
      if (verbose) fprintf(stderr, "-> err = %.3f\n", err);
      for (i = 0; i < cnt; i++) {
        const int p = order[pos + i];
        node[p] = node[p] * bf + sub[i];
      }
      free(sub);
    }
    free(begin);
    nodes *= bf;
  }
  if (clust_assign_out) {
    *clust_assign_out = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    memcpy(*clust_assign_out, node, sizeof(int) * (size_t)n);
  }
  free(node);
  free(order);
  free(grouped);
  return h;
}

void hkm_delete(hkm_t *h) { /* yael/hkm.c:121-128 */
  int l;
  for (l = 0; l < h->nlevel; l++) free(h->centroids[l]);
  free(h->centroids);
  free(h);
}

void hkm_quantize(const hkm_t *h, int npt, const float *v, int *idx) { /* yael/hkm.c:144-162 */
  const int nlevel = h->nlevel, bf = h->bf, d = h->d;
  ybh_arg *tab;
  const float **dev;
  long rows = 1, chunk, i0;
  int l;
  if (npt <= 0) return;
  tab = (ybh_arg *)malloc(sizeof(ybh_arg) * (size_t)(nlevel > 0 ? nlevel : 1));
  dev = (const float **)malloc(sizeof(float *) * (size_t)(nlevel > 0 ? nlevel : 1));
  assert(tab && dev);
  for (l = 0; l < nlevel; l++) {
    rows *= bf;
    tab[l] = ybh_in(h->centroids[l], sizeof(float) * (size_t)rows * d);
    dev[l] = (const float *)tab[l].dev;
  }
  /* the candidate table is npt x bf ints: walk the points in chunks of at most 2^26 candidates */
  chunk = ((long)1 << 26) / (bf > 0 ? bf : 1);
  if (chunk < 1024) chunk = 1024;
  for (i0 = 0; i0 < npt; i0 += chunk) {
    const long m = npt - i0 < chunk ? npt - i0 : chunk;
    ybh_arg av = ybh_in(v + (size_t)i0 * d, sizeof(float) * (size_t)m * d);
    ybh_arg ai = ybh_out(idx + i0, sizeof(int) * (size_t)m);
    YBH_CHECK(yb_hkm_quantize(nlevel, bf, d, dev, m, (const float *)av.dev, (int *)ai.dev, NULL));
    ybh_finish(&ai, 1);
    ybh_finish(&av, 0);
  }
  ybh_sync();
  for (l = 0; l < nlevel; l++) ybh_finish(&tab[l], 0);
  free(tab);
  free(dev);
}

float *hkm_get_centroids(const hkm_t *h, int l, int no) { /* yael/hkm.c:166-169 */
  return h->centroids[l] + (size_t)h->d * h->bf * no;
}

/* File format (yael/hkm.c:181-232): int32 nlevel, bf, d, then level l's table as ONE vector in the
 * fvecs framing (int32 length, then the floats). */
void hkm_write(const char *filename, const hkm_t *h) {
  FILE *f = fopen(filename, "w");
  long rows = 1;
  int l;
  assert(f);
  if (fwrite(&h->nlevel, sizeof(int), 1, f) != 1 || fwrite(&h->bf, sizeof(int), 1, f) != 1 ||
      fwrite(&h->d, sizeof(int), 1, f) != 1)
    goto bad;
  for (l = 0; l < h->nlevel; l++) {
    int len;
    rows *= h->bf;
    len = (int)(rows * h->d);
    if (fwrite(&len, sizeof(int), 1, f) != 1 || fwrite(h->centroids[l], sizeof(float), (size_t)len, f) != (size_t)len)
      goto bad;
  }
  fclose(f);
  return;
bad:
  fprintf(stderr, "# Unable to write the hkm file %s\n", filename);
  fclose(f);
}

hkm_t *hkm_read(const char *filename) {
  FILE *f = fopen(filename, "r");
  int nlevel, bf, d, l;
  long rows = 1;
  hkm_t *h;
  if (!f) {
    fprintf(stderr, "# Unable to read the hkm file %s\n", filename);
    return NULL;
  }
  if (fread(&nlevel, sizeof(int), 1, f) != 1 || fread(&bf, sizeof(int), 1, f) != 1 ||
      fread(&d, sizeof(int), 1, f) != 1) {
    fprintf(stderr, "# Unable to read the hkm file %s\n", filename);
    fclose(f);
    return NULL;
  }
  h = hkm_alloc(d, nlevel, bf);
  for (l = 0; l < nlevel; l++) {
    int len = 0;
    rows *= bf;
    if (fread(&len, sizeof(int), 1, f) != 1 || len != rows * d ||
        fread(h->centroids[l], sizeof(float), (size_t)len, f) != (size_t)len) {
      fprintf(stderr, "# Unable to read the hkm file %s\n", filename);
      fclose(f);
      hkm_delete(h);
      return NULL;
    }
  }
  fclose(f);
  return h;
}
