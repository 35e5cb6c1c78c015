/* yael/nn.h -- drop-in prototypes of the reference's nearest-neighbour API
 * (replaces /root/reference/yael/nn.h:41-214; same names, argument order and meaning).
 * Vectors are rows of a row-major [n][d] float array ("column-major (d,n)" in the
 * reference's Fortran wording, yael/nn.h:15-23).  All pointers may be host pointers (the
 * reference's contract) or CUDA device pointers (detected with cudaPointerGetAttributes).
 * Implemented by libyael_b200.so on sm_100a; there is no CPU fallback. */
#ifndef YAEL_B200_NN_H
#define YAEL_B200_NN_H
#ifdef __cplusplus
extern "C" {
#endif

/* yael/nn.h:41-45, yael/nn.c:451-525.  Base first, query second; assign[nq][k], dis[nq][k]
 * ascending.  Ties are ordered by id (the reference: by heap slot). */
void knn_full(int distance_type, int nq, int nb, int d, int k, const float *b, const float *q,
              const float *b_weights, int *assign, float *dis);
/* yael/nn.h:49-54, yael/nn.c:679-699.  n_thread is accepted and ignored: results of the
 * reference do not depend on it. */
void knn_full_thread(int distance_type, int nq, int nb, int d, int k, const float *b,
                     const float *q, const float *b_weights, int *assign, float *dis,
                     int n_thread);
/* yael/nn.h:59-81, yael/nn.c:608-632,704-726 */
double nn(int n, int nb, int d, const float *b, const float *v, int *assign);
double nn_thread(int n, int nb, int d, const float *b, const float *v, int *assign, int n_thread);
float *knn(int n, int nb, int d, int k, const float *b, const float *v, int *assign);
float *knn_thread(int nq, int nb, int d, int k, const float *b, const float *v, int *assign,
                  int n_thread);
/* yael/nn.h:100-102, yael/nn.c:528-580 */
void knn_reorder_shortlist(int n, int nb, int d, int k, const float *b, const float *v, int *idx,
                           float *dis);
/* yael/nn.h:125-128, yael/nn.c:583-600 */
void knn_recompute_exact_dists(int n, int nb, int d, int k, const float *b, const float *v,
                               int label0, int *kp, const int *idx, float *dis);
/* yael/nn.h:141-160, yael/nn.c:92-129,777-792: dist2[i + na*j] = |a_i - b_j|^2 */
void compute_cross_distances(int d, int na, int nb, const float *a, const float *b, float *dist2);
void compute_cross_distances_nonpacked(int d, int na, int nb, const float *a, int lda,
                                       const float *b, int ldb, float *dist2, int ldd);
void compute_cross_distances_thread(int d, int na, int nb, const float *a, const float *b,
                                    float *dist2, int nt);
/* yael/nn.h:173-188, yael/nn.c:280-356,795-810: distance_type 1 L1, 2 L2, 3 chi2, 4 chi2 abs,
 * 5 histogram intersection, 6 dot product, 12 / 16 the sgemm forms of 2 / 6 */
void compute_cross_distances_alt(int distance_type, int d, int na, int nb, const float *a,
                                 const float *b, float *dist2);
void compute_cross_distances_alt_nonpacked(int distance_type, int d, int na, int nb,
                                           const float *a, int lda, const float *b, int ldb,
                                           float *dist2, int ldd);
void compute_cross_distances_alt_thread(int distance_type, int d, int na, int nb, const float *a,
                                        const float *b, float *dist2, int nt);
/* yael/nn.h:191-211, yael/nn.c:132-162,830-860 */
void compute_distances_1(int d, int nb, const float *a, const float *b, float *dist2);
void compute_distances_1_nonpacked(int d, int nb, const float *a, const float *b, int ldb,
                                   float *dist2);
void compute_distances_1_thread(int d, int nb, const float *a, const float *b, float *dist2,
                                int n_thread);
void compute_distances_1_nonpacked_thread(int d, int nb, const float *a, const float *b, int ldb,
                                          float *dist2, int n_thread);
#ifdef __cplusplus
}
#endif
#endif
