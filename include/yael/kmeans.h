/* yael/kmeans.h -- drop-in prototypes (replaces /root/reference/yael/kmeans.h:11-66). */
#ifndef YAEL_B200_KMEANS_H
#define YAEL_B200_KMEANS_H
#ifdef __cplusplus
extern "C" {
#endif

/* yael/kmeans.h:11-18; flags & 0xffff = requested CPU threads (accepted, unused) */
#define KMEANS_QUIET 0x10000
#define KMEANS_INIT_BERKELEY 0x20000
#define KMEANS_NORMALIZE_CENTS 0x40000
#define KMEANS_INIT_RANDOM 0x80000
#define KMEANS_INIT_USER 0x100000
#define KMEANS_L1 0x200000   /* not a contraction: rejected loudly */
#define KMEANS_CHI2 0x400000 /* idem */

/* yael/kmeans.h:41-44, yael/kmeans.c:332-447.  v[n][d]; centroids[k][d] (input too under
 * KMEANS_INIT_USER); dis[n], assign[n], nassign[k] may be NULL.  Returns qerr/n, or -1 after
 * the "reassigned ... abandoning" message (yael/kmeans.c:302-306). */
float kmeans(int d, int n, int k, int niter, const float *v, int flags, long seed, int redo,
             float *centroids, float *dis, int *assign, int *nassign);

/* yael/kmeans.h:49-66, yael/kmeans.c:452-504: forward-compatibility wrappers; the returned
 * centroid block (and *clust_assign_out) is malloc'd, the caller frees it */
float *clustering_kmeans(int n, int d, const float *points, int k, int nb_iter_max,
                         double normalize);
float *clustering_kmeans_assign(int n, int d, const float *points, int k, int nb_iter_max,
                                double normalize, int **clust_assign_out);
float *clustering_kmeans_assign_with_score(int n, int d, const float *points, int k,
                                           int nb_iter_max, double normalize, int n_thread,
                                           double *score_out, int **clust_assign_out);
#ifdef __cplusplus
}
#endif
#endif
