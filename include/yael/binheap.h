/* yael/binheap.h -- drop-in fbinheap (replaces /root/reference/yael/binheap.h:20-87).
 * Host-side utility kept for callers of the reference API; on the GPU path its role (the
 * streaming k-smallest selector of knn_full, yael/nn.c:470-519) is played by the kernels in
 * yael_b200/csrc/kernels.  Same struct layout, so fbinheap_sizeof / fbinheap_init work on
 * caller-provided memory exactly as in the reference. */
#ifndef YAEL_B200_BINHEAP_H
#define YAEL_B200_BINHEAP_H
#include <stdlib.h>
#ifdef __cplusplus
extern "C" {
#endif
struct fbinheap_s {
  float *val; /* valid entries are val[1..k] */
  int *label;
  int k;
  int maxk;
};
typedef struct fbinheap_s fbinheap_t;

struct fbinheap_s *fbinheap_new(int maxk);
size_t fbinheap_sizeof(int maxk);
void fbinheap_init(fbinheap_t *bh, int maxk);
void fbinheap_delete(fbinheap_t *bh);
void fbinheap_reset(fbinheap_t *bh);
void fbinheap_add(fbinheap_t *bh, int label, float val);
void fbinheap_pop(fbinheap_t *bh);
void fbinheap_addn(fbinheap_t *bh, int n, const int *labels, const float *v);
void fbinheap_addn_label_range(fbinheap_t *bh, int n, int label0, const float *v);
void fbinheap_sort_labels(fbinheap_t *bh, int *perm);
void fbinheap_sort_values(fbinheap_t *bh, float *v);
void fbinheap_sort(fbinheap_t *bh, int *labels, float *v);
#ifdef __cplusplus
}
#endif
#endif
