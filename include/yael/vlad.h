/* vlad.h -- consumers of the k = 1 search: VLAD descriptors and bag-of-features histograms.
 * Same prototypes as the reference's yael/vlad.h:9-49 (host or device pointers).  Assignment =
 * nn() / knn_thread() (the tensor-core k-NN of this library), aggregation on the device in the
 * reference's summation order (bit-identical descriptors). */
#ifndef YAEL_B200_VLAD_H
#define YAEL_B200_VLAD_H

#ifdef __cplusplus
extern "C" {
#endif

/* yael/vlad.c:10-27: desc[k][d] = sum_i (v_i - centroids[nn(v_i)]), points in increasing order */
void vlad_compute(int k, int d, const float *centroids, int n, const float *v, float *desc);
/* yael/vlad.c:30-49: every residual times weights[i] */
void vlad_compute_weighted(int k, int d, const float *centroids, int n, const float *v,
                           const float *weights, float *desc);
/* yael/vlad.c:52-79: one descriptor per subset; subset ss lists the points
 * subset_indexes[subset_ends[ss-1] .. subset_ends[ss]) and is summed in list order */
void vlad_compute_subsets(int k, int d, const float *centroids, int n, const float *v, int n_subset,
                          const int *subset_indexes, const int *subset_ends, float *desc);
/* yael/vlad.c:110-122: desc[k] = histogram of the nearest centroids */
void bof_compute(int k, int d, const float *centroids, int n, const float *v, int *desc);
/* yael/vlad.c:125-138: multiple assignment: the ma nearest centroids of every point are counted
 * (alpha is unused by the reference as well) */
void bof_compute_ma(int k, int d, const float *centroids, int n, const float *v, int *desc, int ma,
                    float alpha, int nt);
/* yael/vlad.c:82-107: one (float) histogram per subset */
void bof_compute_subsets(int k, int d, const float *centroids, int n, const float *v, int n_subset,
                         const int *subset_indexes, const int *subset_ends, float *desc);

#ifdef __cplusplus
}
#endif
#endif
