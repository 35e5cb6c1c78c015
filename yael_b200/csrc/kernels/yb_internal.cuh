// yb_internal.cuh -- cross-file internal entry points (not exported).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../../include/yael_b200.h"

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
  int pair;        // CTAs launched as clusters of 2 sharing the database stream
  int stream;      // query chunks travel through the ring with the database chunks (any d; 2-SM kernel)
  int nka;         // resident query tile: chunk slots (4; 5 .. 7 = the 2-SM kernel's wide layout; 0: streamed)
  size_t ws_bytes; // workspace for buffers + shortlists
  int kind;        // operand kind: 0 = FP32 rows read as TF32 (kind::tf32); 1 = E4M3 bytes
                   // (kind::f8f6f4; `base` / `query` point to [rows][4*d] BYTE matrices -- the Hamming
                   // path); 2 = FP16 (kind::f16; [rows][d] half matrices, d % 8 == 0; k = 1 mode);
                   // 5 = E4M3 bytes whose |b|^2 is one constant (score_c0; Hamming, one row per
                   // accumulator): no |b|^2 tiles, padding rows are E4M3 NaN;
                   // 3 = FP16 with folded norms (the LAST 16 elements of a row carry |b|^2
                   // (database) resp. 2^15 (queries), see center_operands_h; the database copy is
                   // padded to tf32_padded_rows(nb) rows; top-k', sampling and dump modes)
  const float *acc_scale;  // device scalar a: score = acc * a + |b|^2 (NULL: a = -2)
  // kind 1 only: > 1 = packed Hamming passes (yb_hamming_tc.cu): ham_slots consecutive database
  // rows share one accumulator; ham_nb real rows; ham_magic = 2^23 + (bits/2)(1 + 2^8 [+ 2^16])
  int ham_slots, ham_nb;
  float ham_magic;
  float score_c0;  // kind 5 (E4M3, constant |b|^2): score = acc * a + score_c0
};
Tf32Plan tf32_plan(int nq, int nb, int d, int k, int kind = 0);
Tf32Plan tf32_plan_tiles(int nq, int nbt_logical, int d, int kprime, int kind = 0);
int tf32_kprime_for(int k);
int tf32_pair_mode(int kind, int tiles_q);
// One pass of the tensor-core kernel over the logical tiles 0..nbt_logical-1, logical tile j being
// database tile j*tile_stride (256 rows each).  Produces, for every query, `lists` shortlists of
// `kprime` candidates: out_score[q][l][e] = |b|^2 - 2<q,b> evaluated with TF32 operands,
// out_id[q][l][e] the row id (unused slots: +inf / -1), and out_thr[q][l]: every visited row of
// the list's range that is NOT listed has a TF32 score >= out_thr[q][l].  thr_init[q] (may be
// NULL) is the initial admission threshold.  bnorm_padded: |b|^2 for tf32_padded_rows(nb) rows,
// the padding filled with +inf.
// Optional output placement of a shortlist pass: several passes (row ranges of one database) can
// publish into one set of arrays [nq][lists_ld][kprime]; out_cnt[q][l] receives the number of
// entries of every list, which lets the merge skip the empty slots.
struct Tf32Out {
  int *out_cnt;
  int lists_ld;
  int list0;
  int id0;
  float *gmin;   // group-minimum mode (tf32_group_min)
  long gmin_ld;
  int gsize;
  int cross;     // dump mode writes out[row * ld + query] (tf32_cross)
};
int tf32_cross(const Tf32Plan &plan, int nq, int nb, int d, const float *base, const float *query,
               float *out, long ld, void *ws, cudaStream_t st);
// yb_knn.cu: compute_cross_distances on the tensor cores (split-precision FP16 operands, both norms
// folded into the contraction); -1000: shape does not qualify, -1001: a value left FP16's range
int cross_l2_tensor(int d, int na, int nb, const float *a, const float *b, float *out, long ldd,
                    cudaStream_t st);
int tf32_group_min(const Tf32Plan &plan, int nq, int nb, int d, int nbt_logical, int tile_stride,
                   const float *base, const float *query, const float *bnorm_padded, float *gmin,
                   long ld, int gsize, void *ws, cudaStream_t st);
int tf32_shortlist(const Tf32Plan &plan, int nq, int nb, int d, int nbt_logical, int tile_stride,
                   const float *base, const float *query, const float *bnorm_padded,
                   const float *thr_init, float *out_score, int *out_id, float *out_thr, void *ws,
                   cudaStream_t st, const Tf32Out *oo = nullptr);
int tf32_scores(const Tf32Plan &plan, int nq, int nb, int d, int nbt_logical, int tile_stride,
                const float *base, const float *query, const float *bnorm_padded, float *scores,
                long ld, void *ws, cudaStream_t st);
Tf32Plan tf32_plan_nearest(int nq, int nb, int d, int kind = 0);
int tf32_nearest(const Tf32Plan &plan, int nq, int nb, int d, const float *base, const float *query,
                 const float *bnorm_padded, const float *k1_margin, float *out_score, int *out_id,
                 float *out_thr, void *ws, cudaStream_t st);
// yb_knn.cu: thr[r] = the j-th smallest of vals[r][0..n) for nrow rows (n <= row_kth_max_n())
int row_kth(const float *vals, long ld, int nrow, int n, int j, float *thr, cudaStream_t st);
int row_kth_max_n();

// yb_hamming_tc.cu: nn_hamming as an exact E4M3 contraction on the tensor cores
bool hamming_tc_supported(int nq, int nb, int W, int k);
int hamming_tc(int nq, int nb, int W, int k, const unsigned long long *pb,
               const unsigned long long *pq, int *assign, uint16_t *dis, int id_offset,
               int **flag_list_out, int *n_flag_out, cudaStream_t st);
int hamming_tc_scores(int nq, int nb, int W, const unsigned long long *pb,
                      const unsigned long long *pq, float *scores, cudaStream_t st);

int hamming_tc_packed_dump(int nq, int nb, int W, int slots, const unsigned long long *pb,
                           const unsigned long long *pq, float *out, cudaStream_t st);
// yb_runtime.cu: the process pinned a device with yb_set_device (one process per GPU)
bool device_pinned();
// yb_comm.cu: sharded bodies shared by the SPMD entry points and the in-process multi-GPU mode
int knn_sharded_impl(yb_comm *c, int nq, int nb_local, int d, int k, float *base, const float *base_host,
                     const float *query, int id_offset, int *assign, float *dis, int *slice_assign,
                     float *slice_dis, yb_stream_t s);
int hamming_sharded_impl(yb_comm *c, int nq, int nb_local, int ncodes, int k, const uint8_t *base,
                         const uint8_t *query, int id_offset, int *assign, uint16_t *dis,
                         int *slice_assign, uint16_t *slice_dis, yb_stream_t s);
yb_kmeans_comm_t kmeans_hook(yb_comm *c, long n_total, const float *v_host_all);

long tf32_padded_rows(int nb);
int tf32_tiles(int nb);
int fill_f32(float *p, long n, float v, cudaStream_t st);

}  // namespace yb
