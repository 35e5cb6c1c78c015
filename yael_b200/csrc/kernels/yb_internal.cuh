// yb_internal.cuh -- cross-file internal entry points (not exported).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace yb {

// yb_distance.cu
size_t l2_ws_bytes(long na, long nb);
int row_norms_seq(const float *x, long n, int d, long ld, float *out_f, double *out_d,
                  cudaStream_t st);
int l2_matrix(int d, long na, long nb, const float *a, long lda, const float *b, long ldb,
              const float *an_f, const double *bn_d, const float *a_weights, float *out,
              long ldd, cudaStream_t st);

// yb_select.cu
size_t kmin_ws_bytes(long nrow, int k);
int kmin_rows(const float *val, long n, long ld, long nrow, int k, int sign, int *idx,
              float *vals, int id_offset, int flags, void *ws, cudaStream_t st);

// yb_knn_tf32.cu: tcgen05 shortlist kernel
struct Tf32Plan {
  int ok;          // 0: shape not supported by the tensor-core path
  int kprime;      // shortlist length per (query, split)
  int cap;         // append-buffer capacity per (query, split)
  int splits;      // database ranges per query tile
  int lists;       // shortlists produced per query (2 per range: one per column half)
  int ctas;        // persistent grid size
  size_t ws_bytes; // workspace for buffers + shortlists
};
Tf32Plan tf32_plan(int nq, int nb, int d, int k);
// Produces, for every query, `lists` shortlists of `kprime` candidates: out_score[q][s][e] =
// |b|^2 - 2<q,b> evaluated with TF32 operands, out_id[q][s][e] the row id (unused slots:
// +inf / -1).  Every database row that is NOT listed for (q, s) has a TF32 score >= the
// largest listed score of a full list.  bnorm_padded: |b|^2 for tf32_padded_rows(nb) rows,
// the padding filled with +inf.
int tf32_shortlist(const Tf32Plan &plan, int nq, int nb, int d, const float *base,
                   const float *query, const float *bnorm_padded, float *out_score, int *out_id,
                   void *ws, cudaStream_t st);
long tf32_padded_rows(int nb);
int fill_f32(float *p, long n, float v, cudaStream_t st);

}  // namespace yb
