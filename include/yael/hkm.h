/* hkm.h -- hierarchical k-means quantiser.  Same structure and prototypes as
 * the reference's yael/hkm.h:11-40.  hkm_learn runs the reference's level-by-level loop on the
 * host with this library's kmeans() (device) per node; hkm_quantize walks all points down the tree
 * on the device, one exact k = 1 search among the bf children per level (nn() semantics:
 * yael/nn.c:608-621, lowest id on exact ties). */
#ifndef YAEL_B200_HKM_H
#define YAEL_B200_HKM_H

#ifdef __cplusplus
extern "C" {
#endif

/* yael/hkm.h:11-17 */
typedef struct hkm_s {
  int nlevel;         /* number of levels */
  int bf;             /* the branching factor */
  int k;              /* the number of leaves (bf^nlevel) */
  int d;              /* dimension of the input vectors */
  float **centroids;  /* centroids[l]: the bf^(l+1) centroids of level l (host memory) */
} hkm_t;

/* yael/hkm.c:35-118: learn the tree; *clust_assign_out (may be NULL) receives a malloc'd array
 * with the leaf of every learning point */
hkm_t *hkm_learn(int n, int d, int nlevel, int bf, const float *v, int nb_iter_max, int nt,
                 int verbose, int **clust_assign_out);
/* yael/hkm.c:121-128 */
void hkm_delete(hkm_t *hkm);
/* yael/hkm.c:144-162: idx[i] = leaf of v_i; v and idx may be host or device pointers */
void hkm_quantize(const hkm_t *hkm, int n, const float *v, int *idx);
/* yael/hkm.c:181-232: file format = nlevel, bf, d (int32), then every level's table as one vector
 * in the fvecs framing (int32 length + floats) */
void hkm_write(const char *filename, const hkm_t *hkm);
hkm_t *hkm_read(const char *filename);
/* yael/hkm.c:166-169: centroids of node `no` at level l */
float *hkm_get_centroids(const hkm_t *hkm, int l, int no);

#ifdef __cplusplus
}
#endif
#endif
