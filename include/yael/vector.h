/* yael/vector.h -- the subset of the reference's vector.h (allocation, RNG, .fvecs/.ivecs/
 * .bvecs I/O, a few BLAS-1 loops) that callers of the hot path use, so that progs/knn.c and
 * progs/kmeans.c link unmodified.  Same prototypes as
 * /root/reference/yael/vector.h:47-634.  All host code. */
#ifndef YAEL_B200_VECTOR_H
#define YAEL_B200_VECTOR_H
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif
float *fvec_new(long n);              /* vector.c:34-42: memalign(16), abort on OOM */
int *ivec_new(long n);                /* vector.c:56-64 */
unsigned char *bvec_new(long n);
float *fvec_new_0(long n);
int *ivec_new_0(long n);
float *fvec_new_set(long n, float val);
float *fvec_new_cpy(const float *v, long n);
int *ivec_new_cpy(const int *v, long n);
void fvec_0(float *v, long n);
void ivec_0(int *v, long n);
void fvec_cpy(float *dst, const float *src, long n);
void ivec_cpy(int *dst, const int *src, long n);
/* RNG: glibc rand_r sequences, vector.c:135-253 */
void fvec_randn_r(float *v, long n, unsigned int seed);
void fvec_rand_r(float *v, long n, unsigned int seed);
float *fvec_new_rand_r(long n, unsigned int seed);
float *fvec_new_randn_r(long n, unsigned int seed);
int *ivec_new_random_idx_r(int n, int k, unsigned int seed);
int *ivec_new_random_perm_r(int n, unsigned int seed);
/* BLAS-1 style loops, vector.c:1792-1830, 2016-2026, 2066-2075, 2180-2213 */
void fvec_mul_by(float *v, long n, double scal);
void fvec_add(float *v1, const float *v2, long n);
void fvec_sub(float *v1, const float *v2, long n);
double fvec_sum(const float *v, long n);
double fvec_norm(const float *v, long n, double norm);
double fvec_normalize(float *v, long n, double norm);
long fvec_purge_nans(float *v, long n, float replace_value); /* vector.c:1955-1964 */
double ivec_unbalanced_factor(const int *hist, long n);      /* vector.c:2301-2314 */
double fvec_distance_L2sqr(const float *v1, const float *v2, long n); /* vector.c:2348-2359 */
/* file format (doc/file_format.rst:4-20): per vector an int32 dimension then d values */
long fvecs_fsize(const char *fname, int *d_out, int *n_out); /* vector.c:593-640 */
long ivecs_fsize(const char *fname, int *d_out, int *n_out);
long bvecs_fsize(const char *fname, int *d_out, int *n_out);
int fvecs_read(const char *fname, int d, int n, float *v);   /* vector.c:882-920 */
int fvecs_new_read(const char *fname, int *d_out, float **vf);
int ivecs_new_read(const char *fname, int *d_out, int **vi);
int bvecs_new_read(const char *fname, int *d_out, unsigned char **v_out);
int fvecs_write(const char *fname, int d, int n, const float *vf); /* vector.c:1459-1472 */
int fvecs_read_txt(const char *fname, int d, int n, float *v);     /* vector.c:939-964 */
int b2fvecs_read(const char *fname, int d, int n, float *v);        /* vector.c:923-936 */
int fvecs_write_txt(const char *fname, int d, int n, const float *vf); /* vector.c:1475-1491 */
int ivecs_write_txt(const char *fname, int d, int n, const int *v);    /* vector.c:1279-1295 */
int ivecs_write(const char *fname, int d, int n, const int *v);    /* vector.c:1494-1537 */
#ifdef __cplusplus
}
#endif
#endif
