/* yael_hamming.c -- include/yael/hamming.h on top of the yb_ C ABI (yael/hamming.c). */
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../../include/yael/hamming.h"
#include "yb_host.h"

uint16 hamming(const uint8 *bs1, const uint8 *bs2, int ncodes) { /* hamming.c:66-78 */
  unsigned h = 0;
  for (int i = 0; i < ncodes; i++) h += (unsigned)__builtin_popcount((unsigned)(bs1[i] ^ bs2[i]));
  return (uint16)h;
}

void compute_hamming(uint16 *dis, const uint8 *a, const uint8 *b, int na, int nb, int ncodes) {
  /* hamming.c:177-219 */
  if (na <= 0 || nb <= 0) return;
  ybh_arg aa = ybh_in(a, (size_t)na * ncodes);
  ybh_arg ab = ybh_in(b, (size_t)nb * ncodes);
  ybh_arg od = ybh_out(dis, sizeof(uint16) * (size_t)na * nb);
  YBH_CHECK(yb_compute_hamming((uint16_t *)od.dev, (const uint8_t *)aa.dev, (const uint8_t *)ab.dev,
                               na, nb, ncodes, NULL));
  ybh_finish(&od, 1);
  ybh_finish(&aa, 0);
  ybh_finish(&ab, 0);
  ybh_sync();
}

void nn_hamming(int nq, int nb, int ncodes, int k, const uint8 *b, const uint8 *q, int *assign,
                uint16 *dis) {
  assert(k <= nb); /* same precondition as knn_full, yael/nn.c:456 */
  if (nq <= 0 || k <= 0) return;
  if (!ybh_is_device_ptr(b) && !ybh_is_device_ptr(q) && !ybh_is_device_ptr(assign) &&
      !ybh_is_device_ptr(dis)) { /* large problems: sharded over the box's GPUs (yb_mgpu.cu) */
    int rc = yb_mgpu_nn_hamming(nq, nb, ncodes, k, b, q, assign, dis);
    if (rc == 0) return;
    if (rc > 0) ybh_die("yb_mgpu_nn_hamming", rc);
  }
  ybh_arg ab = ybh_in(b, (size_t)nb * ncodes);
  ybh_arg aq = ybh_in(q, (size_t)nq * ncodes);
  ybh_arg oa = ybh_out(assign, sizeof(int) * (size_t)nq * k);
  ybh_arg od = ybh_out(dis, sizeof(uint16) * (size_t)nq * k);
  YBH_CHECK(yb_nn_hamming(nq, nb, ncodes, k, (const uint8_t *)ab.dev, (const uint8_t *)aq.dev,
                          (int *)oa.dev, (uint16_t *)od.dev, 0, NULL));
  ybh_finish(&oa, 1);
  ybh_finish(&od, 1);
  ybh_finish(&ab, 0);
  ybh_finish(&aq, 0);
  ybh_sync();
}

void match_hamming_count(const uint8 *bs1, const uint8 *bs2, int n1, int n2, int ht, int ncodes,
                         size_t *nptr) { /* hamming.c:283-300 */
  ybh_arg a1 = ybh_in(bs1, (size_t)n1 * ncodes);
  ybh_arg a2 = ybh_in(bs2, (size_t)n2 * ncodes);
  unsigned long long *cnt = (unsigned long long *)yb_malloc(sizeof(unsigned long long));
  unsigned long long h = 0;
  YBH_CHECK(yb_match_hamming_count((const uint8_t *)a1.dev, (const uint8_t *)a2.dev, n1, n2, ht,
                                   ncodes, cnt, NULL));
  YBH_CHECK(yb_d2h(&h, cnt, sizeof(h), NULL));
  ybh_sync();
  yb_free(cnt);
  ybh_finish(&a1, 0);
  ybh_finish(&a2, 0);
  *nptr = (size_t)h;
}

size_t match_hamming_thres_prealloc(const uint8 *bs1, const uint8 *bs2, int n1, int n2, int ht,
                                    int ncodes, int *idx, uint16 *hams) { /* hamming.c:704-748 */
  size_t n = 0;
  match_hamming_count(bs1, bs2, n1, n2, ht, ncodes, &n);
  if (n == 0) return 0;
  ybh_arg a1 = ybh_in(bs1, (size_t)n1 * ncodes);
  ybh_arg a2 = ybh_in(bs2, (size_t)n2 * ncodes);
  ybh_arg oi = ybh_out(idx, sizeof(int) * 2 * n);
  ybh_arg oh = ybh_out(hams, sizeof(uint16) * n);
  unsigned long long *cnt = (unsigned long long *)yb_malloc(sizeof(unsigned long long));
  YBH_CHECK(yb_match_hamming_thres((const uint8_t *)a1.dev, (const uint8_t *)a2.dev, n1, n2, ht,
                                   ncodes, (int *)oi.dev, (uint16_t *)oh.dev, cnt, NULL));
  ybh_finish(&oi, 1);
  ybh_finish(&oh, 1);
  yb_free(cnt);
  ybh_finish(&a1, 0);
  ybh_finish(&a2, 0);
  ybh_sync();
  return n;
}

void match_hamming_thres(const uint8 *bs1, const uint8 *bs2, int n1, int n2, int ht, int ncodes,
                         size_t bufsize, hammatch_t **hmptr, size_t *nptr) { /* hamming.c:516-560 */
  (void)bufsize;
  size_t n = 0;
  match_hamming_count(bs1, bs2, n1, n2, ht, ncodes, &n);
  hammatch_t *hm = (hammatch_t *)malloc(sizeof(hammatch_t) * (n ? n : 1));
  if (n) {
    int *idx = (int *)malloc(sizeof(int) * 2 * n);
    uint16 *hams = (uint16 *)malloc(sizeof(uint16) * n);
    match_hamming_thres_prealloc(bs1, bs2, n1, n2, ht, ncodes, idx, hams);
    for (size_t i = 0; i < n; i++) {
      hm[i].qid = idx[2 * i];
      hm[i].bid = idx[2 * i + 1];
      hm[i].score = hams[i];
    }
    free(idx);
    free(hams);
  }
  *hmptr = hm;
  *nptr = n;
}

void compute_hamming_thread(uint16 *dis, const uint8 *a, const uint8 *b, int na, int nb,
                            int ncodes) { /* hamming.c:832-843 */
  compute_hamming(dis, a, b, na, nb, ncodes);
}

void crossmatch_hamming_count(const uint8 *dbs, int n, int ht, int ncodes, size_t *nptr) {
  /* hamming.c:368-395 */
  ybh_arg a1 = ybh_in(dbs, (size_t)(n > 0 ? n : 0) * ncodes);
  unsigned long long *cnt = (unsigned long long *)yb_malloc(sizeof(unsigned long long));
  unsigned long long h = 0;
  YBH_CHECK(yb_crossmatch_hamming_count((const uint8_t *)a1.dev, n, ht, ncodes, cnt, NULL));
  YBH_CHECK(yb_d2h(&h, cnt, sizeof(h), NULL));
  ybh_sync();
  yb_free(cnt);
  ybh_finish(&a1, 0);
  *nptr = (size_t)h;
}

size_t crossmatch_hamming_prealloc(const uint8 *dbs, long n, int ht, int ncodes, int *idx,
                                   uint16 *hams) { /* hamming.c:793-829 */
  size_t m = 0;
  assert(n <= 0x7fffffffL); /* ids are int in the reference too */
  crossmatch_hamming_count(dbs, (int)n, ht, ncodes, &m);
  if (m == 0) return 0;
  ybh_arg a1 = ybh_in(dbs, (size_t)n * ncodes);
  ybh_arg oi = ybh_out(idx, sizeof(int) * 2 * m);
  ybh_arg oh = ybh_out(hams, sizeof(uint16) * m);
  unsigned long long *cnt = (unsigned long long *)yb_malloc(sizeof(unsigned long long));
  YBH_CHECK(yb_crossmatch_hamming((const uint8_t *)a1.dev, (int)n, ht, ncodes, (int *)oi.dev,
                                  (uint16_t *)oh.dev, cnt, NULL));
  ybh_finish(&oi, 1);
  ybh_finish(&oh, 1);
  yb_free(cnt);
  ybh_finish(&a1, 0);
  ybh_sync();
  return m;
}

void crossmatch_hamming(const uint8 *dbs, long n, int ht, int ncodes, long bufsize,
                        hammatch_t **hmptr, size_t *nptr) { /* hamming.c:751-790 */
  (void)bufsize;
  size_t m = 0;
  assert(n <= 0x7fffffffL);
  crossmatch_hamming_count(dbs, (int)n, ht, ncodes, &m);
  hammatch_t *hm = (hammatch_t *)malloc(sizeof(hammatch_t) * (m ? m : 1));
  if (m) {
    int *idx = (int *)malloc(sizeof(int) * 2 * m);
    uint16 *hams = (uint16 *)malloc(sizeof(uint16) * m);
    crossmatch_hamming_prealloc(dbs, n, ht, ncodes, idx, hams);
    for (size_t i = 0; i < m; i++) {
      hm[i].qid = idx[2 * i];
      hm[i].bid = idx[2 * i + 1];
      hm[i].score = hams[i];
    }
    free(idx);
    free(hams);
  }
  *hmptr = hm;
  *nptr = m;
}
