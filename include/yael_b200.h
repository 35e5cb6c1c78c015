/*
 * yael_b200.h -- device-level C ABI of libyael_b200.so.
 *
 * Two layers are exported by the library:
 *
 *  1. The DROP-IN layer: yael's own C prototypes (include/yael/ headers -- knn_full,
 *     knn_full_thread, compute_cross_distances, fvec_k_min, kmeans, compute_hamming,
 *     nn_hamming ...).  Host pointers in, host pointers out, exactly what a program
 *     linked against the reference's libyael.so binds (yael/Makefile:31-39).
 *
 *  2. This layer (prefix yb_): the same operations on DEVICE pointers and a caller
 *     stream, used by the drop-in layer itself (yael_b200/csrc/host/), by the
 *     multi-GPU plumbing (yael_b200/dist.py hands in torch tensors' data_ptr()) and by
 *     bench.py's HBM-resident timing.  Plain C: pointers, sizes, an opaque stream
 *     handle (cudaStream_t passed as void*).  No C++ or torch types.
 *
 * Conventions: every yb_ function returns 0 on success and a non-zero code on failure
 * (yb_last_error() describes it); nothing here falls back to the CPU.  Vectors are
 * row-major [n][d] float (yael/vector.h:32-42, yael/nn.h:15-23).  Indices are int.
 */
#ifndef YAEL_B200_H
#define YAEL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *yb_stream_t; /* cudaStream_t; NULL = the library's own per-device (non-blocking)
                              stream; pass cudaStreamLegacy (0x1) for the legacy default stream */

/* ---- runtime ------------------------------------------------------------------ */
const char *yb_version(void);
const char *yb_last_error(void);
int yb_device_count(void);            /* <0 if the CUDA runtime is unusable            */
int yb_set_device(int dev);           /* device used by subsequent calls of this thread */
int yb_sync(yb_stream_t s);           /* cudaStreamSynchronize on s (or the own stream) */
/* number of kernels this library launched since the last reset (bench.py: gpu_launches) */
long yb_launch_count(int reset);
/* which kNN engine the last yb_knn_l2 call used: 1 = tcgen05 TF32 shortlist + FP32 re-rank,
 * 0 = exact FP32 SIMT path; and how many queries failed the shortlist certificate and were
 * re-done by the exact path */
/* bring-up / tests: raw scores |b|^2 - 2<q,b> of the FP16-operand tensor pass (the default
 * operand kind of the resident k-NN / k-means passes; YAEL_B200_OPERANDS=tf32 selects FP32 rows
 * read as TF32).  scores[nq][nb]; returns 7 if a value overflowed FP16 even after scaling. */
int yb_debug_f16_scores(int nq, int nb, int d, const float *base, const float *query, float *scores,
                        yb_stream_t s);
int yb_last_knn_engine(void);
/* operand kind of the last resident tensor pass: 0 = TF32, 2 = FP16 */
int yb_last_knn_operands(void);
long yb_last_knn_uncertified(void);
/* force an engine for testing: -1 auto (default), 0 exact SIMT only, 1 TF32 whenever legal */
void yb_set_knn_engine(int engine);

/* phase timing with CUDA events on the launching stream (off by default).  Phases:
 * 0 row norms, 1 tcgen05 TF32 shortlist kernel, 2 shortlist merge-select, 3 exact FP32 re-rank,
 * 4 exact-engine fallback, 5 exact distance slab (k_l2_simt), 6 per-row select (k_kmin_rows),
 * 7 Hamming popcount scan, 8 k-means accumulate (sort + segmented sums), 9 k-means scale,
 * 10 / 11 k-NN admission-threshold sampling, 12 Hamming code expansion (+-1 E4M3), 13 Hamming
 * threshold sampling, 14 Hamming tcgen05 E4M3 pass, 15 Hamming order + certify, 16 sharded
 * exchange (NCCL all-to-all / all-gather / all-reduce), 17 merge of a query slice, 18 the row stream
 * of the k-means update alone (k_segsum_sorted, inside phase 8) */
void yb_prof_enable(int on);
double yb_prof_ms(int phase, long *count, int reset);

/* device memory helpers for C callers.  yb_malloc/yb_free go through a small caching pool
 * (freed blocks are kept for reuse; yb_release_scratch() returns them to the driver);
 * allocation failure prints a message and aborts, as the reference's allocators do
 * (yael/vector.c:37-40). */
void *yb_malloc(size_t bytes);
void yb_free(void *p);
int yb_is_device_ptr(const void *p); /* 1: device/managed memory, 0: host memory */
/* gather rows: dst[i][:] = src[rows[i]][:] (all device pointers) */
int yb_gather_rows(const float *src, const int *rows, int n, int d, float *dst, yb_stream_t s);
int yb_h2d(void *dst, const void *src, size_t bytes, yb_stream_t s);
int yb_d2h(void *dst, const void *src, size_t bytes, yb_stream_t s);
void yb_release_scratch(void); /* drop the cached workspace of the current device */

/* ---- distances: yael/nn.c:92-162 ---------------------------------------------- */
/* dist2[i + ldd*j] = fl32(fl64(|b_j|^2) + fl32(|a_i|^2)) - 2 <a_i,b_j>, the dot product
 * accumulated as a sequential FP32 FMA chain over the coordinates (the order the
 * reference's sgemm micro-kernel uses).  Replaces
 * compute_cross_distances_nonpacked (yael/nn.c:100-129). */
int yb_cross_distances_l2(int d, int na, int nb, const float *a, int lda, const float *b,
                          int ldb, float *dist2, int ldd, yb_stream_t s);
/* engine of yb_cross_distances_l2 / compute_cross_distances: 0 = exact FP32 on CUDA cores (the
 * reference's rounding sequence, bit for bit), 1 = tcgen05 with split-precision FP16 operands and
 * both norms folded into the contraction (within 1e-5 relative of the exact engine; packed
 * matrices, na > 128), -1 = automatic: tensor cores for large problems (default;
 * YAEL_B200_CROSS_ENGINE overrides).  yb_last_cross_engine: what the last call of this thread used. */
void yb_set_cross_engine(int engine);
int yb_last_cross_engine(void);
/* dist2[j] for one query a[d] against nb rows, both norms in double:
 * compute_distances_1_nonpacked (yael/nn.c:132-154). */
int yb_distances_1(int d, int nb, const float *a, const float *b, int ldb, float *dist2,
                   yb_stream_t s);
/* alternative distances, compute_cross_distances_alt_nonpacked (yael/nn.c:280-350):
 * 1 L1, 2 L2 (double accumulation), 3 chi2, 4 chi2 |.|, 5 hist. intersection, 6 dot,
 * 12 = yb_cross_distances_l2, 16 = dot product as FP32 FMA chain. */
int yb_cross_distances_alt(int distance_type, int d, int na, int nb, const float *a, int lda,
                           const float *b, int ldb, float *dist2, int ldd, yb_stream_t s);

/* ---- exact k-NN: knn_full (yael/nn.c:451-525), nn_single_full (yael/nn.c:383-446) -- */
/* base[nb][d], query[nq][d] -> assign[nq][k] (ids + id_offset), dis[nq][k] ascending by
 * (distance, id); NaN distances are never selected; short lists are padded with id -1 and
 * distance bits 0xffffffff (yael/nn.c:515-518).  b_weights (device, may be NULL) multiplies
 * the squared distances per base vector (yael/nn.c:497-500).  k == 1 follows nn_single_full:
 * strict '<' from (-1, 1e30), lowest id on ties. */
int yb_knn_l2(int nq, int nb, int d, int k, const float *base, const float *query,
              const float *b_weights, int *assign, float *dis, int id_offset, yb_stream_t s);
/* knn_full for the other distance types (1 L1, 3/4 chi2, 5 histogram intersection, 6/16 dot
 * product, 2/12 L2 with double / FP32 accumulation): compute_cross_distances_alt
 * (yael/nn.c:280-350) in query chunks + the per-base weights of yael/nn.c:497-500 + the per-row
 * select; k == 1 follows nn_single_full (start (-1, 1e30f), strict '<') for every type. */
int yb_knn_alt(int distance_type, int nq, int nb, int d, int k, const float *base,
               const float *query, const float *b_weights, int *assign, float *dis, yb_stream_t s);
/* yb_knn_l2 for a database that is still in HOST memory (what knn_full() receives,
 * yael/nn.c:451): the host->device transfer is overlapped with the scan -- the sample tiles the
 * admission thresholds are computed from travel first, the rest follows in large 2-D copies and
 * every chunk is scanned as soon as it has landed.  base_dev: device scratch for nb*d floats
 * (holds the database on return).  query/assign/dis are device pointers.  Same results as
 * yb_knn_l2. */
int yb_knn_l2_hostbase(int nq, int nb, int d, int k, const float *base_host, float *base_dev,
                       const float *query, int *assign, float *dis, int id_offset, yb_stream_t s);
/* merge G per-shard results laid out [G][nq][k] into [nq][k] by (distance, id); padded
 * entries (id < 0) sort last.  The exchange step of the sharded k-NN. */
int yb_knn_merge(int nq, int k, int G, const int *assign_in, const float *dis_in,
                 int *assign_out, float *dis_out, yb_stream_t s);
/* the same with an explicit distance (in elements) between the lists of consecutive shards:
 * 2 * nq * k when each shard ships ids and distances in ONE [2][nq][k] buffer (one collective) */
int yb_knn_merge_strided(int nq, int k, int G, const int *assign_in, const float *dis_in,
                         long shard_stride, int *assign_out, float *dis_out, yb_stream_t s);
/* knn_reorder_shortlist (yael/nn.c:528-580): exact one-vs-many distances for the listed
 * ids (stop at the first id < 0), re-ordered ascending by (distance, position). */
int yb_knn_reorder_shortlist(int nq, int nb, int d, int k, const float *base,
                             const float *query, int *idx, float *dis, yb_stream_t s);

/* Test / bring-up entry of the tensor-core engine: scores[q][n] = |b_n|^2 - 2 <q, b_n> with TF32
 * operands for every pair (the fused top-k switched off).  d must be a multiple of 4, <= 128. */
int yb_debug_tf32_scores(int nq, int nb, int d, const float *base, const float *query,
                         float *scores, yb_stream_t s);

/* bring-up: clock64() attribution of the last tensor pass run with YAEL_B200_TF32_DEBUG bit 512
 * (out[cta][16]: issuer waits for accumulator / operands / extras, issuer total, epilogue warp 0
 * waits for an accumulator / drains it / hands it back, epilogue total, tiles) */
int yb_debug_tf32_clocks(long long *out, int n_cta);

/* ---- k smallest: fvec_k_min / fvecs_k_min (yael/sorting.c:191-255) ---------------- */
/* nrow arrays of length n (row stride ld) -> idx[nrow][k] (+ optional vals[nrow][k]),
 * ascending by (value, index); sign = +1 for k-min, -1 for k-max (values negated as the
 * reference does, yael/sorting.c:153). */
int yb_k_min_rows(const float *val, long n, long ld, long nrow, int k, int sign, int *idx,
                  float *vals, yb_stream_t s);

/* ---- k-means building blocks: kmeans_core (yael/kmeans.c:213-329) ----------------- */
/* nassign[k] histogram + centroid sums[k][d] (strict point order inside each centroid when
 * exact_order != 0, as yael/kmeans.c:278-283) + qerr = sum of dis as double
 * (yael/kmeans.c:310).  Outputs are UNSCALED sums so a sharded run can all-reduce them. */
int yb_kmeans_accumulate(int d, int n, int k, const float *v, const int *assign,
                         const float *dis, float *sums, int *nassign, double *qerr,
                         int exact_order, yb_stream_t s);
/* centroids[j][t] = (float)((double)sums[j][t] * (1.0 / nassign[j])) (yael/kmeans.c:286-288,
 * yael/vector.c:1792-1797); optional L2 normalisation (yael/kmeans.c:291-293). */
int yb_kmeans_scale(int d, int k, const float *sums, const int *nassign, float *centroids,
                    int normalize, yb_stream_t s);

/* hooks for the sharded k-means: called by the host loop after the local accumulation
 * with DEVICE pointers; must leave the global sums in place on every rank */
typedef struct {
  void *ctx;
  /* returns 0 on success.  `sums` has room for 2 * n_int + 4 floats behind its n_float sums (a
   * hook may pack the counts and qerr there and reduce everything in ONE collective) */
  int (*allreduce_sums)(void *ctx, float *sums, long n_float, int *nassign, long n_int,
                        double *qerr, yb_stream_t s);
  long n_total; /* global number of points (0 = n) */
  /* optional (may be NULL): ALL n_total points in host memory.  With it the sharded run may use
   * every initialisation of the reference (random / k-means++ draws are replayed identically on
   * every rank from these rows) and redo > 1; without it the init must be KMEANS_INIT_USER. */
  const float *v_host_all;
  int rank; /* rank 0 prints the progress messages of a verbose run */
} yb_kmeans_comm_t;

/* kmeans (yael/kmeans.c:332-447) on a device-resident v[n][d]; every other argument as the
 * reference's (host pointers; centroids is also the input under KMEANS_INIT_USER).  With
 * comm != NULL the points are this rank's shard and the init must be KMEANS_INIT_USER. */
float yb_kmeans_dev(int d, int n, int k, int niter, const float *v_dev, int flags, long seed,
                    int redo, float *centroids, float *dis, int *assign, int *nassign,
                    const yb_kmeans_comm_t *comm, yb_stream_t s);

/* ---- consumers of the k = 1 search: VLAD / bag of features (yael/vlad.c:10-139) ------- */
/* desc[k][d] = sum over the listed points (list == NULL: points 0 .. n_list-1), in LIST ORDER, of
 * fl32(v_i - centroids[assign_i]) (times weights[i] when given): the reference's summation order,
 * so the descriptor is bit-identical to vlad_compute / _weighted / one subset of _subsets.
 * k <= 16384. */
int yb_vlad_accumulate(int k, int d, const float *centroids, long n_list, const int *list,
                       const float *v, const int *assign, const float *weights, float *desc,
                       yb_stream_t s);
/* desc[k] = how many listed entries of assign[0 .. n_assign) name each centroid (bof_compute,
 * bof_compute_ma, one subset of bof_compute_subsets); desc_f (may be NULL) receives the counts as
 * floats */
int yb_bof_accumulate(int k, long n_list, const int *list, const int *assign, long n_assign, int *desc,
                      float *desc_f, yb_stream_t s);

/* ---- Hamming: yael/hamming.c:66-219 ------------------------------------------------ */
/* dis[j*na + i] = popcount(a_i xor b_j), uint16 (compute_hamming, yael/hamming.c:177-219) */
int yb_compute_hamming(uint16_t *dis, const uint8_t *a, const uint8_t *b, int na, int nb,
                       int ncodes, yb_stream_t s);
/* NEW (not in the reference): k smallest Hamming distances per query ordered by
 * (distance, id); padding id -1 / distance 0xffff. */
int yb_nn_hamming(int nq, int nb, int ncodes, int k, const uint8_t *base, const uint8_t *query,
                  int *assign, uint16_t *dis, int id_offset, yb_stream_t s);
/* engine of yb_nn_hamming: 0 = popcount scan (CUDA cores), 1 = exact E4M3 contraction on the
 * tensor cores (falls back per query to the scan), -1 = automatic (default; also
 * YAEL_B200_HAMMING_ENGINE).  yb_last_hamming_engine / _fallbacks describe the last call. */
void yb_set_hamming_engine(int engine);
int yb_last_hamming_engine(void);
long yb_last_hamming_fallbacks(void);
/* bring-up / tests: raw scores of the E4M3 pass, scores[q][n] = 4 * hamming(query q, base n) */
int yb_debug_hamming_tc_scores(int nq, int nb, int ncodes, const uint8_t *base,
                               const uint8_t *query, float *scores, yb_stream_t s);
/* bring-up / tests: the PACKED pass (`slots` consecutive database rows share one accumulator):
 * out[q][c] = -2 * (dot_0 + 2^8 dot_1 + 2^16 dot_2), dot_i = bits - 2 hamming(q, row slots*c+i),
 * c < ceil(nb / slots); absent rows contribute 0 */
int yb_debug_hamming_tc_packed(int nq, int nb, int ncodes, int slots, const uint8_t *base,
                               const uint8_t *query, float *out, yb_stream_t s);
int yb_nn_hamming_merge(int nq, int k, int G, const int *assign_in, const uint16_t *dis_in,
                        int *assign_out, uint16_t *dis_out, yb_stream_t s);
/* micro-benchmark: measured 64-bit xor+popcount pair rate of the whole GPU (the ceiling the
 * Hamming scan is reported against) */
double yb_debug_popc_pairs_per_s(yb_stream_t s);
/* bring-up: GB/s written into out[nb][ld] (queries fastest, the layout of compute_cross_distances)
 * with the store pattern `mode` (yb_distance.cu: 0 = 4 bytes per lane, 1 / 2 = 16 bytes per lane) */
double yb_debug_store_pattern_gbs(float *out, long ld, int nq, int nb, int mode, yb_stream_t s);
/* match_hamming_count / match_hamming_thres_prealloc (yael/hamming.c:283-300, 563-700):
 * pairs with distance <= ht, emitted query-major / base-ascending.  count is a device
 * size_t; idx receives (qid, bid) interleaved. */
int yb_match_hamming_count(const uint8_t *bs1, const uint8_t *bs2, int n1, int n2, int ht,
                           int ncodes, unsigned long long *count, yb_stream_t s);
int yb_match_hamming_thres(const uint8_t *bs1, const uint8_t *bs2, int n1, int n2, int ht,
                           int ncodes, int *idx, uint16_t *hams, unsigned long long *count,
                           yb_stream_t s);
/* crossmatch_hamming_count / crossmatch_hamming_prealloc (yael/hamming.c:368-395, 793-829):
 * pairs i < j of one code set with distance <= ht, emitted as (i, j) in (i, j) order. */
int yb_crossmatch_hamming_count(const uint8_t *dbs, int n, int ht, int ncodes,
                                unsigned long long *count, yb_stream_t s);
int yb_crossmatch_hamming(const uint8_t *dbs, int n, int ht, int ncodes, int *idx,
                          uint16_t *hams, unsigned long long *count, yb_stream_t s);

/* ---- further consumers of the path ---------------------------------- */
/* hkm_quantize (yael/hkm.c:144-162): idx[i] = leaf of point i after nlevel exact k = 1 searches
 * among the bf children of its current node (nn() semantics, lowest id on exact ties).  levels is a
 * HOST array of nlevel DEVICE pointers, level l's table being [bf^(l+1)][d]; v, idx: device. */
int yb_hkm_quantize(int nlevel, int bf, int d, const float *const *levels, long n, const float *v,
                    int *idx, yb_stream_t s);
/* GMM E-step (yael/gmm.c:211-367): p[n][k] = posteriors.  inv_sigma = (float)(1.0 / sigma) and
 * mu_sigma = mu / sigma ([k][d]), mu2[k] = (float) sum_l mu^2 / sigma, logdetnr[k], lg[k] (log
 * weights or zeros) are the O(k d) tables the reference prepares before its two sgemm calls;
 * coeffs (may be NULL) receives log(sum) + max per point.  Device pointers. */
int yb_gmm_posteriors(long n, int k, int d, const float *v, const float *inv_sigma,
                      const float *mu_sigma, const float *mu2, const float *logdetnr, const float *lg,
                      float *p, float *coeffs, yb_stream_t s);

/* ---- sharded hot path: the exchange steps of SURVEY.md 8(e) inside the library ------- */
/* One NCCL communicator per GPU (NCCL is resolved at run time: libnccl.so.2).  Two ways to get
 * one: (a) one PROCESS per GPU (torchrun): rank 0 calls yb_comm_unique_id, the caller broadcasts
 * the 128 bytes, every rank calls yb_comm_create on its device; (b) one host THREAD per GPU in a
 * single process: yb_comm_create_all (ncclCommInitAll) -- what the drop-in layer's own multi-GPU
 * mode below uses.  The reference has no counterpart: it is one process with OpenMP threads
 * (yael/nn.c:665-699), whose slicing rule [n*r/G, n*(r+1)/G) is the sharding rule here. */
typedef struct yb_comm yb_comm;
int yb_comm_available(void);              /* 1 when NCCL could be loaded */
int yb_comm_unique_id(void *id128);       /* 128 bytes */
yb_comm *yb_comm_create(const void *id128, int rank, int world); /* collective; NULL on failure */
int yb_comm_create_all(int ndev, const int *devs, yb_comm **out);
void yb_comm_destroy(yb_comm *c);
/* 1 when the exchange steps of this communicator run over peer memory (every rank's segment mapped
 * into every other's address space: NVLink stores + a flag barrier, no NCCL kernel on the path;
 * YAEL_B200_NO_P2P=1 / YAEL_B200_P2P_MB size the segments), 0 when they use NCCL send / recv */
int yb_comm_p2p(const yb_comm *c);
int yb_comm_rank(const yb_comm *c);
int yb_comm_world(const yb_comm *c);
int yb_comm_allreduce_f32(yb_comm *c, float *buf, long n, yb_stream_t s);
int yb_comm_allgather(yb_comm *c, const void *send, void *recv, long bytes_per_rank, yb_stream_t s);
/* exact k-NN over a database sharded by rows (SPMD: every rank calls with ITS shard
 * base[nb_local][d], the global id of its first row and the same queries): local search, then a
 * QUERY-PARTITIONED exchange -- rank r receives queries [r*slice, (r+1)*slice) of every rank's
 * lists (all-to-all), merges them by (distance, id), and one all-gather distributes the merged
 * slices.  assign / dis [nq][k] receive the result for the whole database on every rank,
 * identical for any number of ranks.  (yael/nn.c:451-525 semantics.) */
int yb_knn_l2_sharded(yb_comm *c, int nq, int nb_local, int d, int k, const float *base,
                      const float *query, int id_offset, int *assign, float *dis, yb_stream_t s);
/* the same with the shard still in HOST memory (transfer overlapped with the scan,
 * yb_knn_l2_hostbase); base_dev: device scratch for nb_local * d floats */
int yb_knn_l2_sharded_hostbase(yb_comm *c, int nq, int nb_local, int d, int k,
                               const float *base_host, float *base_dev, const float *query,
                               int id_offset, int *assign, float *dis, yb_stream_t s);
int yb_nn_hamming_sharded(yb_comm *c, int nq, int nb_local, int ncodes, int k, const uint8_t *base,
                          const uint8_t *query, int id_offset, int *assign, uint16_t *dis,
                          yb_stream_t s);
/* kmeans (yael/kmeans.c:332-447) on points sharded by rows: v_dev is this rank's shard
 * [n_local][d], centroids the k initial centroids (identical on every rank) and the result; ONE
 * all-reduce of (sums | counts | qerr) per iteration on the compute stream.  assign / dis (host,
 * may be NULL) receive this rank's points' assignment. */
float yb_kmeans_sharded(yb_comm *c, int d, int n_local, long n_total, int k, int niter,
                        const float *v_dev, int flags, long seed, float *centroids, float *dis,
                        int *assign, int *nassign, yb_stream_t s);

/* ---- the drop-in layer's own multi-GPU mode (one process, one host thread per GPU) ---- */
/* knn_full, nn_hamming and kmeans called with HOST pointers shard large problems over the GPUs
 * named by YAEL_GPU_DEVICES ("all", or a comma-separated list of ordinals; default: all) -- unless
 * the process pinned a device with yb_set_device (one process per GPU: the caller shards).
 * yb_mgpu_set_devices overrides the environment (n = 0: back to it; n = 1: single GPU).
 * yb_mgpu_device_count: how many GPUs the next qualifying call would use. */
int yb_mgpu_set_devices(int n, const int *devs);
int yb_mgpu_device_count(void);
/* how many GPUs the last knn_full / nn_hamming / kmeans call of this thread used */
int yb_mgpu_last_used(void);
/* the sharded bodies: return 0 when done, -1 when the call does not qualify (single GPU, small
 * problem, k > rows per shard; the caller then takes the one-GPU path), > 0 on failure.  Host
 * pointers throughout. */
int yb_mgpu_knn_full(int nq, int nb, int d, int k, const float *base, const float *query,
                     int *assign, float *dis);
int yb_mgpu_nn_hamming(int nq, int nb, int ncodes, int k, const uint8_t *base, const uint8_t *query,
                       int *assign, uint16_t *dis);
int yb_mgpu_kmeans(int d, int n, int k, int niter, const float *v, int flags, long seed, int redo,
                   float *centroids, float *dis, int *assign, int *nassign, float *qerr_out);

#ifdef __cplusplus
}
#endif
#endif
