/* yael_binheap.c -- host fbinheap of include/yael/binheap.h (yael/binheap.c:11-211): a
 * fixed-capacity max-heap that keeps the maxk smallest values of a stream.  Kept because it is
 * part of the library surface callers link against; the device path does not use it. */
#include <assert.h>
#include <math.h>
#include <string.h>

#include "../../../include/yael/binheap.h"
#include "../../../include/yael/sorting.h"

size_t fbinheap_sizeof(int maxk) { /* binheap.c:11-14 */
  return sizeof(fbinheap_t) + (size_t)maxk * (sizeof(float) + sizeof(int));
}

void fbinheap_init(fbinheap_t *bh, int maxk) { /* binheap.c:17-26: arrays follow the header */
  char *mem = (char *)bh + sizeof(fbinheap_t);
  bh->k = 0;
  bh->maxk = maxk;
  bh->val = (float *)mem - 1; /* 1-based */
  bh->label = (int *)(mem + (size_t)maxk * sizeof(float)) - 1;
}

fbinheap_t *fbinheap_new(int maxk) { /* binheap.c:29-35 */
  fbinheap_t *bh = (fbinheap_t *)malloc(fbinheap_sizeof(maxk));
  fbinheap_init(bh, maxk);
  return bh;
}
void fbinheap_reset(fbinheap_t *bh) { bh->k = 0; }
void fbinheap_delete(fbinheap_t *bh) { free(bh); }

static void sift_in(fbinheap_t *bh, int label, float val) { /* binheap.c:85-103 */
  assert(bh->k < bh->maxk);
  int pos = ++bh->k;
  for (; pos > 1; pos >>= 1) {
    int up = pos >> 1;
    if (bh->val[up] >= val) break;
    bh->val[pos] = bh->val[up];
    bh->label[pos] = bh->label[up];
  }
  bh->val[pos] = val;
  bh->label[pos] = label;
}

void fbinheap_pop(fbinheap_t *bh) { /* binheap.c:48-82 */
  assert(bh->k > 0);
  const int last = bh->k;
  const float moving = bh->val[last];
  int pos = 1;
  for (;;) {
    int l = 2 * pos, r = l + 1, c;
    if (l > last) break;
    c = (r == last + 1 || bh->val[l] > bh->val[r]) ? l : r;
    if (moving > bh->val[c]) break;
    bh->val[pos] = bh->val[c];
    bh->label[pos] = bh->label[c];
    pos = c;
  }
  bh->val[pos] = bh->val[last];
  bh->label[pos] = bh->label[last];
  bh->k--;
}

void fbinheap_add(fbinheap_t *bh, int label, float val) { /* binheap.c:106-117 */
  if (bh->k < bh->maxk) {
    sift_in(bh, label, val);
  } else if (val < bh->val[1]) {
    fbinheap_pop(bh);
    sift_in(bh, label, val);
  }
}

void fbinheap_addn(fbinheap_t *bh, int n, const int *label, const float *v) { /* :120-136 */
  int i = 0;
  for (; i < n && bh->k < bh->maxk; i++)
    if (!isnan(v[i])) sift_in(bh, label[i], v[i]);
  float root = bh->val[1];
  for (; i < n; i++)
    if (v[i] < root) {
      fbinheap_pop(bh);
      sift_in(bh, label[i], v[i]);
      root = bh->val[1];
    }
}

void fbinheap_addn_label_range(fbinheap_t *bh, int n, int label0, const float *v) { /* :139-156 */
  int i = 0;
  for (; i < n && bh->k < bh->maxk; i++)
    if (!isnan(v[i])) sift_in(bh, label0 + i, v[i]);
  float root = bh->val[1];
  for (; i < n; i++)
    if (v[i] < root) {
      fbinheap_pop(bh);
      sift_in(bh, label0 + i, v[i]);
      root = bh->val[1];
    }
}

void fbinheap_sort_labels(fbinheap_t *bh, int *perm) { /* binheap.c:168-174 */
  fvec_sort_index(bh->val + 1, bh->k, perm);
  for (int i = 0; i < bh->k; i++) perm[i] = bh->label[perm[i] + 1];
}

static int cmp_float(const void *a, const void *b) {
  float x = *(const float *)a, y = *(const float *)b;
  return x == y ? 0 : (x > y ? 1 : -1);
}
void fbinheap_sort_values(fbinheap_t *bh, float *v) { /* binheap.c:177-181 */
  memcpy(v, bh->val + 1, sizeof(float) * (size_t)bh->k);
  qsort(v, (size_t)bh->k, sizeof(float), cmp_float);
}

void fbinheap_sort(fbinheap_t *bh, int *labels, float *v) { /* binheap.c:201-211 */
  fvec_sort_index(bh->val + 1, bh->k, labels);
  for (int i = 0; i < bh->k; i++) {
    int slot = labels[i] + 1;
    labels[i] = bh->label[slot];
    v[i] = bh->val[slot];
  }
}
