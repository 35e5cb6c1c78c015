/* yael_vector.c -- host utilities of include/yael/vector.h: allocation, the rand_r based RNG,
 * BLAS-1 style loops and the .fvecs/.ivecs/.bvecs file format.  Behaviour follows
 * /root/reference/yael/vector.c (lines cited per function); host code, nothing here is on
 * the device path except that the k-means driver depends on the RNG sequences bit for bit.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <malloc.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../../include/yael/vector.h"

static void *checked(void *p, const char *what, long n) {
  if (!p) { /* vector.c:37-40 */
    fprintf(stderr, "%s %ld : out of memory\n", what, n);
    abort();
  }
  return p;
}

float *fvec_new(long n) { /* vector.c:34-42 */
  return (float *)checked(memalign(16, sizeof(float) * (size_t)(n > 0 ? n : 1)), "fvec_new", n);
}
int *ivec_new(long n) { /* vector.c:56-64 */
  return (int *)checked(malloc(sizeof(int) * (size_t)(n > 0 ? n : 1)), "ivec_new", n);
}
unsigned char *bvec_new(long n) {
  return (unsigned char *)checked(malloc((size_t)(n > 0 ? n : 1)), "bvec_new", n);
}
float *fvec_new_0(long n) {
  float *v = fvec_new(n);
  fvec_0(v, n);
  return v;
}
int *ivec_new_0(long n) {
  int *v = ivec_new(n);
  ivec_0(v, n);
  return v;
}
float *fvec_new_set(long n, float val) {
  float *v = fvec_new(n);
  for (long i = 0; i < n; i++) v[i] = val;
  return v;
}
float *fvec_new_cpy(const float *v, long n) {
  float *r = fvec_new(n);
  memcpy(r, v, sizeof(float) * (size_t)n);
  return r;
}
int *ivec_new_cpy(const int *v, long n) {
  int *r = ivec_new(n);
  memcpy(r, v, sizeof(int) * (size_t)n);
  return r;
}
void fvec_0(float *v, long n) { memset(v, 0, sizeof(float) * (size_t)n); }
void ivec_0(int *v, long n) { memset(v, 0, sizeof(int) * (size_t)n); }
void fvec_cpy(float *dst, const float *src, long n) { memmove(dst, src, sizeof(float) * (size_t)n); }
void ivec_cpy(int *dst, const int *src, long n) { memmove(dst, src, sizeof(int) * (size_t)n); }

/* ---- RNG ---- */
static double unit_r(unsigned int *seed) { /* vector.c:135-137 */
  return rand_r(seed) / ((double)RAND_MAX + 1.0);
}

/* vector.c:141-154: ratio-of-uniforms normal deviate; the two uniforms and the acceptance
 * statistic live in float variables, which decides accept/reject at the margin */
static double normal_r(unsigned int *seed) {
  const double c = 1.71552776992141;
  double z;
  for (;;) {
    float u1 = (float)unit_r(seed);
    float u2 = (float)unit_r(seed);
    z = c * (u1 - .5) / u2;
    float quarter_sq = (float)(z * z / 4.0);
    if (quarter_sq < -log(u2)) break;
  }
  return z;
}

void fvec_rand_r(float *v, long n, unsigned int seed) { /* vector.c:170-175 */
  for (long i = 0; i < n; i++) v[i] = (float)unit_r(&seed);
}
void fvec_randn_r(float *v, long n, unsigned int seed) { /* vector.c:184-189 */
  for (long i = 0; i < n; i++) v[i] = (float)normal_r(&seed);
}
float *fvec_new_rand_r(long n, unsigned int seed) {
  float *v = fvec_new(n);
  fvec_rand_r(v, n, seed);
  return v;
}
float *fvec_new_randn_r(long n, unsigned int seed) {
  float *v = fvec_new(n);
  fvec_randn_r(v, n, seed);
  return v;
}

int *ivec_new_random_idx_r(int n, int k, unsigned int seed) { /* vector.c:226-243 */
  int *idx = ivec_new(n);
  for (int i = 0; i < n; i++) idx[i] = i;
  for (int i = 0; i < k; i++) {
    int j = i + rand_r(&seed) % (n - i);
    int t = idx[i];
    idx[i] = idx[j];
    idx[j] = t;
  }
  return idx;
}
int *ivec_new_random_perm_r(int n, unsigned int seed) { /* vector.c:250-253 */
  return ivec_new_random_idx_r(n, n - 1, seed);
}

/* ---- BLAS-1 style ---- */
void fvec_mul_by(float *v, long n, double scal) { /* vector.c:1792-1797: float *= double */
  for (long i = 0; i < n; i++) v[i] = (float)(v[i] * scal);
}
void fvec_add(float *v1, const float *v2, long n) { /* vector.c:1811-1816 */
  for (long i = 0; i < n; i++) v1[i] += v2[i];
}
void fvec_sub(float *v1, const float *v2, long n) { /* vector.c:1824-1829 */
  for (long i = 0; i < n; i++) v1[i] -= v2[i];
}
double fvec_sum(const float *v, long n) { /* vector.c:2066-2074 */
  double s = 0;
  for (long i = 0; i < n; i++) s += v[i];
  return s;
}
double fvec_norm(const float *v, long n, double norm) { /* vector.c:2180-2213 */
  if (norm == 0) return n;
  double s = 0;
  long i;
  if (norm == 1) {
    for (i = 0; i < n; i++) s += fabs(v[i]);
    return s;
  }
  if (norm == 2) {
    for (i = 0; i < n; i++) {
      float p = v[i] * v[i];
      s += p;
    }
    return sqrt(s);
  }
  if (norm == -1) {
    for (i = 0; i < n; i++)
      if (fabs(v[i]) > s) s = fabs(v[i]);
    return s;
  }
  for (i = 0; i < n; i++) s += pow(v[i], norm);
  return pow(s, 1 / norm);
}
double fvec_normalize(float *v, long n, double norm) { /* vector.c:2016-2026 */
  if (norm == 0) return 0;
  double nr = fvec_norm(v, n, norm);
  fvec_mul_by(v, n, 1. / nr);
  return nr;
}
long fvec_purge_nans(float *v, long n, float replace_value) { /* vector.c:1955-1964 */
  long count = 0;
  for (long i = 0; i < n; i++)
    if (isnan(v[i])) {
      count++;
      v[i] = replace_value;
    }
  return count;
}
double ivec_unbalanced_factor(const int *hist, long n) { /* vector.c:2301-2314 */
  double tot = 0, uf = 0;
  for (long i = 0; i < n; i++) {
    tot += hist[i];
    uf += hist[i] * (double)hist[i];
  }
  return uf * n / (tot * tot);
}
double fvec_distance_L2sqr(const float *v1, const float *v2, long n) { /* vector.c:2348-2359 */
  double dis = 0;
  for (long i = 0; i < n; i++) {
    double a = (double)v1[i] - v2[i];
    dis += a * a;
  }
  return dis;
}

/* ---- file format: [int32 d][d values] per vector (doc/file_format.rst:4-20) ---- */
static long vecs_fsize(long unit, const char *fname, int *d_out, int *n_out) {
  /* vector.c:593-626 */
  *d_out = -1;
  *n_out = -1;
  FILE *f = fopen(fname, "r");
  if (!f) {
    fprintf(stderr, "xvecs_fsize %s: %s\n", fname, strerror(errno));
    return -1;
  }
  int d;
  if (fread(&d, sizeof(d), 1, f) == 0) {
    *n_out = 0;
    fclose(f);
    return 0;
  }
  fseek(f, 0, SEEK_END);
  long nbytes = ftell(f);
  fclose(f);
  if (nbytes % (unit * d + 4) != 0) {
    fprintf(stderr, "xvecs_size %s: weird file size %ld for vectors of dimension %d\n", fname,
            nbytes, d);
    return -1;
  }
  *d_out = d;
  *n_out = (int)(nbytes / (unit * d + 4));
  return nbytes;
}
long fvecs_fsize(const char *fname, int *d_out, int *n_out) {
  return vecs_fsize(sizeof(float), fname, d_out, n_out);
}
long ivecs_fsize(const char *fname, int *d_out, int *n_out) {
  return vecs_fsize(sizeof(int), fname, d_out, n_out);
}
long bvecs_fsize(const char *fname, int *d_out, int *n_out) {
  return vecs_fsize(1, fname, d_out, n_out);
}

int fvecs_read(const char *fname, int d, int n, float *a) { /* vector.c:882-920 */
  FILE *f = fopen(fname, "r");
  if (!f) {
    fprintf(stderr, "fvecs_read: could not open %s\n", fname);
    perror("");
    return -1;
  }
  long i;
  for (i = 0; i < n; i++) {
    int new_d;
    if (fread(&new_d, sizeof(int), 1, f) != 1) {
      if (feof(f)) break;
      perror("fvecs_read error 1");
      fclose(f);
      return -1;
    }
    if (new_d != d) {
      fprintf(stderr, "fvecs_read error 2: unexpected vector dimension\n");
      fclose(f);
      return -1;
    }
    if (fread(a + d * i, sizeof(float), d, f) != (size_t)d) {
      fprintf(stderr, "fvecs_read error 3\n");
      fclose(f);
      return -1;
    }
  }
  fclose(f);
  return (int)i;
}

/* reads every vector of a file into one malloc'd block; unit = bytes per component */
static int vecs_new_read(const char *fname, long unit, int *d_out, void **out, const char *who) {
  int d, n;
  long nbytes = vecs_fsize(unit, fname, &d, &n);
  if (nbytes < 0) {
    *d_out = -1;
    return -1;
  }
  *d_out = d;
  if (n == 0) {
    *out = NULL;
    return 0;
  }
  FILE *f = fopen(fname, "r");
  if (!f) {
    fprintf(stderr, "%s: could not open %s\n", who, fname);
    return -1;
  }
  char *buf = (char *)checked(memalign(16, (size_t)unit * d * n), who, (long)d * n);
  for (long i = 0; i < n; i++) {
    int new_d;
    if (fread(&new_d, sizeof(int), 1, f) != 1 || new_d != d ||
        fread(buf + (size_t)unit * d * i, unit, d, f) != (size_t)d) {
      fprintf(stderr, "%s: non-uniform vectors sizes or short read in %s\n", who, fname);
      free(buf);
      fclose(f);
      return -1;
    }
  }
  fclose(f);
  *out = buf;
  return n;
}
int fvecs_new_read(const char *fname, int *d_out, float **vf) { /* vector.c:650-668 */
  return vecs_new_read(fname, sizeof(float), d_out, (void **)vf, "fvecs_new_read");
}
int ivecs_new_read(const char *fname, int *d_out, int **vi) {
  return vecs_new_read(fname, sizeof(int), d_out, (void **)vi, "ivecs_new_read");
}
int bvecs_new_read(const char *fname, int *d_out, unsigned char **v_out) {
  return vecs_new_read(fname, 1, d_out, (void **)v_out, "bvecs_new_read");
}

static int vecs_write(const char *fname, long unit, int d, int n, const void *v, const char *who) {
  FILE *f = fopen(fname, "w");
  if (!f) {
    fprintf(stderr, "%s: cannot open %s for writing", who, fname);
    perror("");
    return -1;
  }
  for (long i = 0; i < n; i++) {
    if (fwrite(&d, sizeof(d), 1, f) != 1 ||
        fwrite((const char *)v + (size_t)unit * d * i, unit, d, f) != (size_t)d) {
      perror(who);
      fclose(f);
      return -1;
    }
  }
  fclose(f);
  return n;
}
int fvecs_write(const char *fname, int d, int n, const float *vf) { /* vector.c:1459-1472 */
  return vecs_write(fname, sizeof(float), d, n, vf, "fvecs_write");
}
int ivecs_write(const char *fname, int d, int n, const int *v) { /* vector.c:1521-1534 */
  return vecs_write(fname, sizeof(int), d, n, v, "ivecs_write");
}

/* ---- text / byte-vector variants used by progs/knn.c and progs/kmeans.c ---- */
int fvecs_read_txt(const char *fname, int d, int n, float *v) { /* vector.c:939-964 */
  FILE *f = fopen(fname, "r");
  if (!f) {
    fprintf(stderr, "fvecs_read_txt: could not open %s\n", fname);
    perror("");
    return -1;
  }
  long i;
  for (i = 0; i < (long)n * d; i++) {
    if (fscanf(f, "%f", v + i) != 1) {
      if (feof(f)) break;
      perror("fvecs_read_txt error 1");
      fclose(f);
      return -1;
    }
  }
  fclose(f);
  return (int)(i / d);
}

int b2fvecs_read(const char *fname, int d, int n, float *v) { /* vector.c:923-936: bytes -> floats */
  int d_file, n_file;
  if (bvecs_fsize(fname, &d_file, &n_file) < 0 || d_file != d || n > n_file) {
    fprintf(stderr, "b2fvecs_read %s: expected %d vectors of dimension %d, file has %d x %d\n", fname,
            n, d, n_file, d_file);
    abort();
  }
  FILE *f = fopen(fname, "r");
  if (!f) {
    fprintf(stderr, "b2fvecs_read: Unable to open %s\n", fname);
    abort();
  }
  unsigned char *row = bvec_new(d);
  for (long i = 0; i < n; i++) {
    int dd;
    if (fread(&dd, sizeof(int), 1, f) != 1 || dd != d || fread(row, 1, d, f) != (size_t)d) {
      fprintf(stderr, "b2fvecs_read %s: short read\n", fname);
      abort();
    }
    for (int j = 0; j < d; j++) v[i * d + j] = row[j];
  }
  free(row);
  fclose(f);
  return n;
}

int fvecs_write_txt(const char *fname, int d, int n, const float *vf) { /* vector.c:1475-1491 */
  int ret = 0;
  FILE *fo = fopen(fname, "w");
  if (!fo) {
    perror("fvecs_write_txt: cannot open file");
    return -1;
  }
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < d; j++) fprintf(fo, "%f ", vf[(long)i * d + j]);
    ret += fprintf(fo, "\n");
  }
  fclose(fo);
  return ret;
}

int ivecs_write_txt(const char *fname, int d, int n, const int *v) { /* vector.c:1279-1295 */
  int ret = 0;
  FILE *fo = fopen(fname, "w");
  if (!fo) {
    perror("ivecs_write_txt: cannot open file");
    return -1;
  }
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < d; j++) fprintf(fo, "%d ", v[(long)i * d + j]);
    ret += fprintf(fo, "\n");
  }
  fclose(fo);
  return ret;
}
