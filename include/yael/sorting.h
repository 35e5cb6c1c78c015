/* yael/sorting.h -- drop-in prototypes of the k-smallest family (replaces
 * /root/reference/yael/sorting.h:19-40; the rest of that header -- ranks, medians, merges --
 * is host utility code outside the hot path). */
#ifndef YAEL_B200_SORTING_H
#define YAEL_B200_SORTING_H
#ifdef __cplusplus
extern "C" {
#endif
/* yael/sorting.c:153-170: indices of the k largest, descending */
void fvec_k_max(const float *v, int n, int *maxes, int k);
/* yael/sorting.c:239-255: indices of the k smallest, ascending; ties by index */
void fvec_k_min(const float *v, int n, int *mins, int k);
/* yael/sorting.c:184-196: n arrays of length m -> idx[n][k] */
void fvecs_k_max(const float *val, long m, long n, int *idx, int k);
void fvecs_k_min(const float *val, long m, long n, int *idx, int k);
/* yael/sorting.c:304-316, 778-789 (host helpers used by callers of the above) */
void fvec_sort_index(const float *tab, int n, int *perm);
int fvec_arg_min(const float *f, long n);
#ifdef __cplusplus
}
#endif
#endif
