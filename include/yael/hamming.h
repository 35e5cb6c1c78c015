/* yael/hamming.h -- drop-in prototypes (replaces /root/reference/yael/hamming.h:6-66) plus
 * the NEW nn_hamming (BASELINE.json north_star; absent from the reference).
 * ncodes is in BYTES (yael/hamming.h:21-24).  Codes of up to 64 bytes are supported. */
#ifndef YAEL_B200_HAMMING_H
#define YAEL_B200_HAMMING_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef unsigned char uint8;
typedef unsigned short uint16;
typedef unsigned int uint32;
typedef unsigned long long uint64;
typedef long long int64;

typedef struct hammatch_s { /* yael/hamming.h:14-18 */
  int qid;
  int bid;
  uint16 score;
} hammatch_t;

/* yael/hamming.c:66-78 (host scalar) */
uint16 hamming(const uint8 *bs1, const uint8 *bs2, int ncodes);
/* yael/hamming.c:177-219: dis[j*na + i] */
void compute_hamming(uint16 *dis, const uint8 *a, const uint8 *b, int na, int nb, int ncodes);
/* NEW: k smallest Hamming distances of each of nq queries among nb base codes, ordered by
 * (distance, id); assign[nq][k], dis[nq][k]; padding id -1 / 0xffff when k > nb. */
void nn_hamming(int nq, int nb, int ncodes, int k, const uint8 *b, const uint8 *q, int *assign,
                uint16 *dis);
/* yael/hamming.c:283-300: number of pairs with distance <= ht */
void match_hamming_count(const uint8 *bs1, const uint8 *bs2, int n1, int n2, int ht, int ncodes,
                         size_t *nptr);
/* yael/hamming.c:516-560: *hmptr is malloc'd here (bufsize is only the reference's initial
 * guess and is ignored), caller frees */
void match_hamming_thres(const uint8 *bs1, const uint8 *bs2, int n1, int n2, int ht, int ncodes,
                         size_t bufsize, hammatch_t **hmptr, size_t *nptr);
/* yael/hamming.c:704-748: idx receives (qid, bid) interleaved, hams the scores */
size_t match_hamming_thres_prealloc(const uint8 *bs1, const uint8 *bs2, int n1, int n2, int ht,
                                    int ncodes, int *idx, uint16 *hams);
/* yael/hamming.c:368-395: number of pairs i < j of one set with distance <= ht */
void crossmatch_hamming_count(const uint8 *dbs, int n, int ht, int ncodes, size_t *nptr);
/* yael/hamming.c:751-790: the pairs themselves (qid = i, bid = j, i < j, in (i, j) order);
 * *hmptr is malloc'd here (bufsize: the reference's initial guess, ignored), caller frees */
void crossmatch_hamming(const uint8 *dbs, long n, int ht, int ncodes, long bufsize,
                        hammatch_t **hmptr, size_t *nptr);
/* yael/hamming.c:793-829: the same into caller memory sized by crossmatch_hamming_count;
 * idx receives (i, j) interleaved; returns the number of pairs */
size_t crossmatch_hamming_prealloc(const uint8 *dbs, long n, int ht, int ncodes, int *idx,
                                   uint16 *hams);
/* yael/hamming.c:832-843: the reference's OpenMP variant (built there only under _OPENMP,
 * yael/hamming.h:69-75); the GPU path has no thread count to honour -- same result as
 * compute_hamming.  (match_hamming_thres_nt, hamming.c:846-903, is NOT provided: it is compiled
 * out of the reference's default build and returns block-local ids from a block index that
 * only covers all pairs when n1 and n2 span equally many 128-blocks.) */
void compute_hamming_thread(uint16 *dis, const uint8 *a, const uint8 *b, int na, int nb,
                            int ncodes);
#ifdef __cplusplus
}
#endif
#endif
